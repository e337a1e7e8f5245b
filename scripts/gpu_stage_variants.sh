# NeuTex weight ring geometry: libraries built with -DNTX_STAGE_BYTES / -DNTX_STAGES under neural-gauge-fields_b200/variants/
mkdir -p gpurun_out
for lib in neural-gauge-fields_b200/libngf_b200.so neural-gauge-fields_b200/variants/*.so; do
  export NGF_B200_LIB=$PWD/$lib
  echo "== $lib"
  timeout 240 python -m pytest tests/test_gpu_neutex.py -m gpu -q -x 2>&1 | tail -1
  NGF_NTX_DBG=4 timeout 150 python scripts/ntx_trace.py 2>&1 | grep "256-wide\|total cycles"
  timeout 150 python scripts/profile_target.py neutex 5 2>&1 | tail -1
done
