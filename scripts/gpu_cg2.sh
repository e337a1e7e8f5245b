# NeuTex MLP kernel: parity + layer timeline + frame timing for NGF_NTX_CG = 1 (default) and 2 (CTA pairs)
mkdir -p gpurun_out
for cg in ${CGS:-1 2}; do
  export NGF_NTX_CG=$cg
  timeout 240 python -m pytest tests/test_gpu_neutex.py -m gpu -q -x > gpurun_out/pytest_cg$cg.log 2>&1; echo "pytest cg$cg rc=$?"; tail -2 gpurun_out/pytest_cg$cg.log
  NGF_NTX_DBG=4 timeout 150 python scripts/ntx_trace.py > gpurun_out/trace_cg${cg}_dbg4.txt 2>&1; echo "trace rc=$?"; grep "256-wide\|total cycles" gpurun_out/trace_cg${cg}_dbg4.txt
  timeout 150 python scripts/profile_target.py neutex 5 2>&1 | tail -1
done
