"""Diagnostic (GPU): per-parameter gradient errors of the CUDA backward against the reference goldens."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs, load_golden

for name in K.GRAD_CASES:
    case = K.TRAIN_BY_NAME[K.GRAD_CASES[name]]
    gold = load_golden(name)
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    u = torch.from_numpy(gold["jitter"])
    target = K.grad_target(rays.shape[0]).cuda()
    f.zero_grad()
    out = f(rays.cuda(), white_bg=True, is_train=True, N_samples=case.n_samples, jitter=u, **forward_kwargs(case))
    loss = torch.mean((out["rgb_map"] - target) ** 2)
    loss.backward()
    print(name, "loss", float(loss.detach()), float(gold["loss"]))
    for k, p in f.named_parameters():
        key = k.replace(".", "__")
        flat = p.grad.detach().reshape(-1).cpu().numpy()
        if "full__" + key in gold:
            ref = gold["full__" + key]; mine = flat
        else:
            idx = gold["idx__" + key]; ref = gold["val__" + key]; mine = flat[idx]
        err = np.abs(mine - ref)
        j = int(err.argmax())
        msg = f"  {k:34s} scale {np.abs(ref).max():.3e} maxerr {err.max():.3e} rel {err.max()/max(np.abs(ref).max(),1e-30):.2e} at ref {ref[j]:.3e} mine {mine[j]:.3e}"
        if "idx__" + key in gold:
            C = p.shape[1]; hw = p.shape[2] * p.shape[3]
            ch = gold["idx__" + key] // hw
            dc = 16 if C == 64 else (24 if C == 96 else 0)
            if C > 2:
                ed, ea = err[ch < dc], err[ch >= dc]
                msg += f" | dens-ch maxerr {ed.max() if ed.size else 0:.2e} app-ch maxerr {ea.max() if ea.size else 0:.2e} ch@max {int(ch[j])}"
            msg += f" | sum {flat.astype(np.float64).sum():.6e} ref {float(gold['sum__'+key]):.6e} abs {np.abs(flat).astype(np.float64).sum():.6e} ref {float(gold['abs__'+key]):.6e}"
        print(msg)
