# round 2, call D (1 GPU): GPU suite, then the InfoInv side configuration with the phased march and with the old one
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for ph in 1 0; do
NGF_INFOINV_PHASED=$ph timeout 600 python - <<'PY'
import os, sys, json, types
sys.path.insert(0, os.getcwd())
import torch, bench, ngf_b200
from ngf_b200 import synth
dev = torch.device("cuda", 0)
host = [synth.config_rays("C2", p).pin_memory() for p in range(16)]
dev_rays = [h.to(dev) for h in host]
args = types.SimpleNamespace(no_cpu_baseline=True, cpu_seconds=1.0)
r = bench.infoinv_config(ngf_b200, synth, dev, bench.peaks(), dev_rays, host, args)
print("phased", os.environ["NGF_INFOINV_PHASED"], {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("workload", "roofline", "e2e")}, "e2e %.3e" % r["e2e"]["value"])
PY
done
