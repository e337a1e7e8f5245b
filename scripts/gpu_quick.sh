mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e ms/step %.3f march %.3f colour %.3f (tensor frac %.3f)'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['colour_kernel']['kernel_ms'], d['roofline']['colour_kernel']['frac']))
dr=d['dense_regime']; print('dense: rays/s %.3e march %.2f colour %.2f frac %.3f exec %.1f'%(dr['rays_per_s'], dr['march_kernel_ms'], dr['colour_kernel_ms'], dr['roofline']['frac'], dr['roofline']['executed_tflops']))
PY
python scripts/diag_parity.py 2>&1 | tail -5
