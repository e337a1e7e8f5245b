mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_neutex.py -m gpu -q > gpurun_out/pytest_neutex.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_neutex.log
timeout 300 python scripts/profile_target.py neutex 3
