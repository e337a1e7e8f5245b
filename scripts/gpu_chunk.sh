mkdir -p gpurun_out
for c in 160000 640000; do
NGF_HOST_CHUNK_ASYNC=$c timeout 300 python bench.py --no-cpu-baseline --no-dense --no-extra > gpurun_out/bench_c$c.log 2> gpurun_out/bench_c$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c$c.log').read().strip().splitlines()[-1])
print('chunk $c value %.4e e2e %.4e (floor %.4e) cam %.4e'%(d['value'], d['e2e']['value'], d['e2e']['transport_floor']['value'], d['e2e_camera']['value']))
PY
done
