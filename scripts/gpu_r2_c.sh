# round 2, call C (1 GPU): parity of the TMA-staged colour kernel, then A/B bench against the direct gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c_pytest.log
for tma in 1 0; do
NGF_COLOUR_TMA=$tma timeout 600 python bench.py --steps 300 --no-extra --no-cpu-baseline > gpurun_out/r2c_bench_tma$tma.log 2> gpurun_out/r2c_bench_tma$tma.err; echo "bench tma=$tma rc=$?"; tail -3 gpurun_out/r2c_bench_tma$tma.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench_tma$tma.log').read().strip().splitlines()[-1])
r=d['roofline']; o=r['other_kernel']
print('tma=$tma value %.3e e2e %.3e cam %.3e ms/step %.3f | %s %.4f | %s %.4f'%(d['value'], d['e2e']['value'], d['e2e_camera']['value'], d['ms_per_step'], r['kernel'][:18], r['kernel_ms'], o['kernel'][:18], o['kernel_ms']))
dr=d['dense_regime']; print('   dense: rays/s %.3e march %.2f colour %.2f frac %.3f'%(dr['rays_per_s'], dr['march_kernel_ms'], dr['colour_kernel_ms'], dr['roofline']['frac']))
PY
done
