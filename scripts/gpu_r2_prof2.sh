mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ngf_colour" -s 2 -c 1 -o gpurun_out/r2_prof_dense_tma -f python scripts/profile_target.py dense 3 > gpurun_out/r2_ncu_dense_tma.log 2>&1; echo "ncu-dense rc=$?"; tail -2 gpurun_out/r2_ncu_dense_tma.log
