"""Run with NGF_QUEUE_MIB=1: the colour queue then holds only 32 Ki work items, so a 4096-ray x 64-sample render is split
into many ray batches (render_dev) and a fog field overflows nothing.  Compares with the reference golden frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs, load_golden
assert os.environ.get("NGF_QUEUE_MIB") == "1"
ok = True
for name in ("tp_fog_c1", "ii_fog_c1"):
    case = K.CASE_BY_NAME[name]
    gold = load_golden(name)
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    for width in (0, 64):
        out = f(rays.cuda(), white_bg=case.white_bg, N_samples=case.n_samples, image_width=width, **forward_kwargs(case))
        e = float(np.abs(out["rgb_map"].cpu().numpy() - gold["rgb"]).max()); d = float(np.abs(out["depth_map"].cpu().numpy() - gold["depth"]).max())
        print(f"{name} image_width={width}: rgb {e:.2e} depth {d:.2e}")
        ok = ok and e < 1e-3 and d < 2e-3
    rgb_h, dep_h = f.render_host(rays.pin_memory(), white_bg=case.white_bg, N_samples=case.n_samples, image_width=64, **forward_kwargs(case))
    ok = ok and float(np.abs(rgb_h.numpy() - gold["rgb"]).max()) < 1e-3
sys.exit(0 if ok else 1)
