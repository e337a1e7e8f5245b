# round 2: ncu --set full of the march + colour kernels on the bench's 16-pose workload (sparse and dense regimes)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ngf_(march|colour)_kernel" -s 36 -c 2 -o gpurun_out/r2_prof_hull -f python scripts/profile_target.py hull 20 > gpurun_out/r2_ncu_hull.log 2>&1; echo "ncu-hull rc=$?"; tail -2 gpurun_out/r2_ncu_hull.log
ncu --set full --clock-control none --import-source on -k regex:"ngf_(march|colour)_kernel" -s 4 -c 2 -o gpurun_out/r2_prof_dense -f python scripts/profile_target.py dense 3 > gpurun_out/r2_ncu_dense.log 2>&1; echo "ncu-dense rc=$?"; tail -2 gpurun_out/r2_ncu_dense.log
ncu --set full --clock-control none --import-source on -k regex:"ngf_(march|colour)_kernel" -s 8 -c 2 -o gpurun_out/r2_prof_infoinv -f python scripts/profile_target.py infoinv 6 > gpurun_out/r2_ncu_infoinv.log 2>&1; echo "ncu-infoinv rc=$?"; tail -2 gpurun_out/r2_ncu_infoinv.log
