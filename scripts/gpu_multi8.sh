mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/bench$n.log 2> gpurun_out/bench$n.err; echo "bench$n rc=$?"; tail -3 gpurun_out/bench$n.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench$n.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('N=$n value %.3e e2e %.3e ms/step %.3f march %.3f colour %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['colour_kernel']['kernel_ms']))
except Exception as e: print('parse fail', e)
PY
done
