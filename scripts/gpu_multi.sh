mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; tail -4 gpurun_out/bench2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench2.log').read().strip().splitlines() if l.startswith('{')][-1])
print('N=2 value %.3e e2e %.3e ms/step %.3f march %.3f colour %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['colour_kernel']['kernel_ms']))
PY
