mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|neutex|tp_fog|ii_fog" gpurun_out/sanitize_$tool.log | tail -5
done
NGF_NTX_CG=2 timeout 200 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/sanitize_memcheck_cg2.log 2>&1; echo "memcheck cg2 rc=$?"; grep -E "ERROR SUMMARY|neutex" gpurun_out/sanitize_memcheck_cg2.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
