mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e ms/step %.3f march %.3f colour %.3f (tensor frac %.3f) launches %d clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['colour_kernel']['kernel_ms'], d['roofline']['colour_kernel']['frac'], d['gpu_launches'], d['clocks']))
dr=d['dense_regime']; print('dense: rays/s %.3e march %.2f colour %.2f frac %.3f exec %.1f'%(dr['rays_per_s'], dr['march_kernel_ms'], dr['colour_kernel_ms'], dr['roofline']['frac'], dr['roofline']['executed_tflops']))
print('infoinv', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['other_configs']['infoinv'].items() if k!='workload'})
nt=d['other_configs']['neutex']; print('neutex', {k:(round(v,4) if isinstance(v,float) else v) for k,v in nt.items() if k not in('workload','roofline')}, nt['roofline']['frac'])
print('cpu', d['cpu_baseline'])
PY
