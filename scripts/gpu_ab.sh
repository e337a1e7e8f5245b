# quick A/B of the TriPlane render path: golden parity, then the bench line's kernel times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or pointwise or camera" > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ab.log
timeout 600 python bench.py --no-cpu-baseline --no-dense > gpurun_out/bench_ab.log 2> gpurun_out/bench_ab.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ab.log').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e ms/step %.4f parity_ok %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('parity_ok')))
r=d['roofline']; print('roofline kernel', r.get('kernel'), 'ms', r.get('kernel_ms'), 'frac', r.get('frac'), 'other', r.get('other_kernel'))
dr=d.get('dense_regime')
if dr: print('dense:', {k: dr[k] for k in dr if k != 'roofline'}, dr['roofline'].get('frac'))
oc=d.get('other_configs', {})
for k,v in oc.items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
