# round 2 final evidence on one GPU: bench line, reference arm, ncu launch list, ncu --set full (hull, dense, InfoInv),
# then the whole GPU test suite, smoke() and compute-sanitizer on the same binary
bash scripts/gpu_r2_final.sh
bash scripts/gpu_r2_prof.sh
bash scripts/gpu_suite.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
bash scripts/gpu_sanitize_r2.sh
