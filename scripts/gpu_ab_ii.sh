# A/B of the InfoInv render path: goldens + full-size test, then the bench's InfoInv side number
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "infoinv or ii_ or golden" > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
timeout 600 python bench.py --no-cpu-baseline --no-dense --steps 300 > gpurun_out/bench_ab.log 2> gpurun_out/bench_ab.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ab.log').read().strip().splitlines()[-1])
v=d['other_configs']['infoinv']
print('infoinv rays/s %.4e ms %.4f march %.4f colour %.4f e2e %.4e'%(v['rays_per_s'], v['ms_per_frame'], v['march_kernel_ms'], v['colour_kernel_ms'], v['e2e']['value']))
print('triplane value %.4e'%d['value'])
PY
