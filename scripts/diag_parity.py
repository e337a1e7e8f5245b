"""Print the CUDA-vs-golden error of every render case (max abs rgb / depth)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs, load_golden
worst = (0, 0)
for case in K.CASES:
    gold = load_golden(case.name)
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    out = f(rays.cuda(), white_bg=case.white_bg, N_samples=case.n_samples, **forward_kwargs(case))
    e = np.abs(out["rgb_map"].cpu().numpy() - gold["rgb"]).max(); d = np.abs(out["depth_map"].cpu().numpy() - gold["depth"]).max()
    worst = (max(worst[0], e), max(worst[1], d))
    print(f"{case.name:22s} rgb {e:.2e} depth {d:.2e}")
print("worst rgb %.2e depth %.2e" % worst)
