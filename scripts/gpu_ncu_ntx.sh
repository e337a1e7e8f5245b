mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ntx_mlp" -s 1 -c 1 -o gpurun_out/prof_neutex python scripts/profile_target.py neutex 1 > gpurun_out/ncu_neutex.log 2>&1; echo "ncu-neutex rc=$?"; tail -2 gpurun_out/ncu_neutex.log; ls -la gpurun_out/*.ncu-rep
