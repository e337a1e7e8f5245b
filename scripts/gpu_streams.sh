mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streams or golden or camera" > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ab.log
timeout 300 python bench.py --no-cpu-baseline --no-dense --no-extra > gpurun_out/bench_s.log 2> gpurun_out/bench_s.err; tail -3 gpurun_out/bench_s.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s.log').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms/step %.4f single %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('value_single_stream')))
print(d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['roofline']['other_kernel']['kernel_ms'], d['roofline']['other_kernel']['kernel_share_of_step'], d['gpu_launches'])
PY
