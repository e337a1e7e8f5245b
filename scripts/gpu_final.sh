# end-of-round GPU pass: full GPU suite, bench line, launch list, ncu captures of the hot kernels
bash scripts/gpu_round.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ntx_" -s 3 -c 3 -o gpurun_out/prof_neutex python scripts/profile_target.py neutex 1 > gpurun_out/ncu_neutex.log 2>&1; echo "ncu-neutex rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ngf_(march|colour)_kernel" -s 4 -c 2 -o gpurun_out/prof_hull python scripts/profile_target.py hull 4 > gpurun_out/ncu_hull.log 2>&1; echo "ncu-hull rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ngf_(march|colour)_kernel" -s 2 -c 2 -o gpurun_out/prof_dense python scripts/profile_target.py dense 2 > gpurun_out/ncu_dense.log 2>&1; echo "ncu-dense rc=$?"
