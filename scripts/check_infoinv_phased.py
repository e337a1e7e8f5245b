"""Run with NGF_INFOINV_PHASED=1 or NGF_INFOINV_TC=1: the InfoInv march then runs as the three-phase cooperative kernel
(csrc/ngf_infoinv_march.cuh: find samples / density MLP per lane / composite, in rounds) or as find / tensor-core density /
composite (csrc/ngf_infoinv_tc.cuh: split-fp16 tcgen05 MMAs).  Every InfoInv golden (render and training-time forward) must
still match."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs, load_golden
assert os.environ.get("NGF_INFOINV_PHASED") == "1" or os.environ.get("NGF_INFOINV_TC") == "1"
ok = True
for case in list(K.CASES) + list(K.TRAIN_CASES):
    if case.variant != "infoinv":
        continue
    gold = load_golden(case.name)
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    train = "jitter" in gold
    with torch.no_grad():
        if train:
            out = f(rays.cuda(), white_bg=bool(gold["white_used"]), is_train=True, N_samples=case.n_samples,
                    jitter=torch.from_numpy(gold["jitter"]), **forward_kwargs(case))
        else:
            out = f(rays.cuda(), white_bg=case.white_bg, N_samples=case.n_samples, image_width=64 if rays.shape[0] == 4096 else 0,
                    **forward_kwargs(case))
    e = float(np.abs(out["rgb_map"].cpu().numpy() - gold["rgb"]).max()); d = float(np.abs(out["depth_map"].cpu().numpy() - gold["depth"]).max())
    print(f"{case.name}: rgb {e:.2e} depth {d:.2e} {f.last_stats()}")
    ok = ok and e < 1e-3 and d < 2e-3
sys.exit(0 if ok else 1)
