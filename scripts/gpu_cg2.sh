mkdir -p gpurun_out
for cg in ${CGS:-1}; do
  NGF_NTX_CG=$cg timeout 240 python -m pytest tests/test_gpu_neutex.py -m gpu -q -x > gpurun_out/pytest_cg$cg.log 2>&1; echo "pytest cg$cg rc=$?"; tail -4 gpurun_out/pytest_cg$cg.log
  for dbg in 4; do
    NGF_NTX_CG=$cg NGF_NTX_DBG=$dbg timeout 150 python scripts/ntx_trace.py > gpurun_out/trace_cg${cg}_dbg$dbg.txt 2>&1; echo "trace cg$cg dbg$dbg rc=$?"; tail -42 gpurun_out/trace_cg${cg}_dbg$dbg.txt
  done
  NGF_NTX_CG=$cg timeout 150 python scripts/profile_target.py neutex 5 2>&1 | tail -1
done
