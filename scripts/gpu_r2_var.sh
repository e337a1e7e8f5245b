# A/B of library variants built with different macros (variants/lib_*.so), short bench each
mkdir -p gpurun_out
for v in "$@"; do
  NGF_COLOUR_TMA=0 NGF_B200_LIB=$PWD/variants/lib_$v.so timeout 600 python bench.py --steps 400 --no-extra --no-dense --no-cpu-baseline > gpurun_out/r2_var_$v.log 2> gpurun_out/r2_var_$v.err; echo "variant $v rc=$?"; tail -2 gpurun_out/r2_var_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_var_$v.log').read().strip().splitlines()[-1])
r=d['roofline']; o=r['other_kernel']
print('$v value %.3e e2e %.3e cam %.3e ms/step %.4f | %s %.4f | %s %.4f'%(d['value'], d['e2e']['value'], d['e2e_camera']['value'], d['ms_per_step'], r['kernel'][:18], r['kernel_ms'], o['kernel'][:18], o['kernel_ms']))
PY
done
