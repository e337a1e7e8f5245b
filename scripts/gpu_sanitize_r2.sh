mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool python scripts/sanitize_target_r2.py > gpurun_out/r2_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "SUMMARY|ok|neutex|train_" gpurun_out/r2_sanitize_$tool.log | tail -8
done
NGF_COLOUR_TMA=1 NGF_INFOINV_TC=1 timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/r2_sanitize_memcheck_optin.log 2>&1; echo "memcheck opt-in rc=$?"; grep -E "ERROR SUMMARY|tp_fog|ii_fog" gpurun_out/r2_sanitize_memcheck_optin.log | tail -4
NGF_COLOUR_TMA=1 NGF_INFOINV_TC=1 timeout 400 compute-sanitizer --tool synccheck python scripts/sanitize_target.py > gpurun_out/r2_sanitize_synccheck_optin.log 2>&1; echo "synccheck opt-in rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/r2_sanitize_synccheck_optin.log | tail -2
NGF_INFOINV_PHASED=1 timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/r2_sanitize_memcheck_phased.log 2>&1; echo "memcheck phased rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/r2_sanitize_memcheck_phased.log | tail -2
