mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_neutex.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_neutex.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_neutex.log
timeout 300 python scripts/profile_target.py neutex 3
timeout 300 python scripts/diag_neutex.py neutex_white | tail -6
