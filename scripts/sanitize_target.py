"""Tiny renders of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs
import ngf_b200
for name in ("tp_fog_c1", "ii_fog_c1"):
    case = K.CASE_BY_NAME[name]
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    out = f(rays[:1500].cuda(), white_bg=True, N_samples=64, **forward_kwargs(case))
    torch.cuda.synchronize()
    print(name, float(out["rgb_map"].mean()), f.last_stats())
case = K.NEUTEX_BY_NAME["neutex_white"]
state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
m = ngf_b200.NeuTex(device="cuda"); m.load_state_dict(state)
out = m(campos.cuda(), raydir[:, :300].cuda(), bg.cuda(), noise=noise[:, :300].cuda())
torch.cuda.synchronize()
print("neutex", float(out["color"].mean()), m.last_valid_samples())
