# round 2, call A (1 GPU): whole GPU suite + a short bench with every new leg
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 200 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2a_bench.log').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e e2e %.3e e2e_cam %.3e ms/step %.3f'%(d['value'], d['e2e']['value'], d['e2e_camera']['value'], d['ms_per_step']))
print('dominant', r['kernel'], r['kernel_ms'], r['frac'], 'other', r['other_kernel']['kernel'], r['other_kernel']['kernel_ms'])
print('cpu', d.get('cpu_baseline')); print('torch_cuda', d.get('torch_cuda_baseline'))
for k,v in d['other_configs'].items(): print(k, {a:b for a,b in v.items() if a not in ('workload',)})
PY
