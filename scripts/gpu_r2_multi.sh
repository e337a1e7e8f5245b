# round 2 (N GPUs): sharded-frame parity for every exchange variant + the bench with the chosen variants
N=${1:-2}
MODES=${2:-copy}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_gpu$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu$N.log
for mode in $MODES; do
  NGF_BENCH_COMM=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r2_bench_n${N}_$mode.log 2> gpurun_out/r2_bench_n${N}_$mode.err; echo "bench $mode rc=$?"; tail -3 gpurun_out/r2_bench_n${N}_$mode.err | grep -v "OMP\|\*\*\*"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_n${N}_$mode.log').read().strip().splitlines() if l.startswith('{')][-1])
    r=d['roofline']; o=r['other_kernel']
    print('N=$N $mode value %.3e e2e %.3e (floor %.3e) e2e_cam %.3e ms/step %.3f parity %s (%.1e/%.1e) %s %.3f | %s %.3f launches %d'%(d['value'], d['e2e']['value'], d['e2e']['transport_floor']['value'], (d.get('e2e_camera') or {}).get('value', 0), d['ms_per_step'], d.get('parity_ok'), d['parity']['rgb_max_abs'], d['parity']['depth_max_abs'], r['kernel'][:16], r['kernel_ms'], o['kernel'][:16], o['kernel_ms'], d['gpu_launches']), d['config'].get('collective_note'))
except Exception as e:
    print('parse failed', e)
PY
done
