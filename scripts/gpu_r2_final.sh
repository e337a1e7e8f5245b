# round 2 evidence (1 GPU): the default bench line, then the ncu launch list of a short bench run
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/r2_bench_ref.err; cut -c1-600 gpurun_out/r2_bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu-list rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.log').read().strip().splitlines()[-1])
r=d['roofline']; o=r['other_kernel']
print('value %.3e e2e %.3e (floor %.3e) cam %.3e ms/step %.4f | %s %.4f frac %.3f | %s %.4f | launches %d clocks %s'%(d['value'], d['e2e']['value'], d['e2e']['transport_floor']['value'], d['e2e_camera']['value'], d['ms_per_step'], r['kernel'][:18], r['kernel_ms'], r['frac'], o['kernel'][:18], o['kernel_ms'], d['gpu_launches'], d['clocks']))
dr=d['dense_regime']; print('dense: rays/s %.3e march %.2f colour %.2f frac %.3f'%(dr['rays_per_s'], dr['march_kernel_ms'], dr['colour_kernel_ms'], dr['roofline']['frac']))
for k,v in d['other_configs'].items(): print(k, 'rays/s %.3e e2e %.3e cpu %s'%(v['rays_per_s'], v['e2e']['value'], v.get('cpu_baseline')))
print('cpu', d['cpu_baseline']); print('torch_cuda', d['torch_cuda_baseline'])
PY
