# round 2, call E (1 GPU): InfoInv tensor-core march: parity, then A/B against the in-march MLP
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ii_ or infoinv or pointwise or train" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2e_pytest.log
for tc in 1 0; do
NGF_INFOINV_TC=$tc timeout 600 python - <<'PY'
import os, sys, json, types
sys.path.insert(0, os.getcwd())
import torch, bench, ngf_b200
from ngf_b200 import synth
dev = torch.device("cuda", 0)
host = [synth.config_rays("C2", p).pin_memory() for p in range(16)]
dev_rays = [h.to(dev) for h in host]
args = types.SimpleNamespace(no_cpu_baseline=True, cpu_seconds=1.0)
r = bench.infoinv_config(ngf_b200, synth, dev, bench.peaks(), dev_rays, host, args)
print("tc", os.environ["NGF_INFOINV_TC"], {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("workload", "roofline", "e2e")}, "e2e %.3e" % r["e2e"]["value"])
PY
done
