# round 2, call B (1 GPU): the new gradient tests (verbose)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "gradients or training_loop or zero_jitter" > gpurun_out/r2b_grad.log 2>&1; echo "grad rc=$?"; grep -E "^\{|Error|assert |passed|failed" gpurun_out/r2b_grad.log | head -40
