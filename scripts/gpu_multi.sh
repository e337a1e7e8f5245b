mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "bench1 rc=$?"; cat gpurun_out/bench1.log; tail -3 gpurun_out/bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; cat gpurun_out/bench2.log; tail -5 gpurun_out/bench2.err
