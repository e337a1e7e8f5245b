# round 2 final evidence on one GPU: bench line, reference arm, ncu launch list, ncu --set full (hull + dense)
bash scripts/gpu_r2_final.sh
bash scripts/gpu_r2_prof.sh
