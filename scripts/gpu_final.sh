# end-of-round GPU pass: full GPU suite, bench line, launch list, ncu captures of the three hot kernels
bash scripts/gpu_round.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"ntx_" -s 3 -c 3 -o gpurun_out/prof_neutex python scripts/profile_target.py neutex 1 > gpurun_out/ncu_neutex.log 2>&1; echo "ncu-neutex rc=$?"
