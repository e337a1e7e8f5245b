"""Run with NGF_COLOUR_TMA=1: the TriPlane colour kernel then stages the plane texels through TMA boxes
(csrc/ngf_colour_tma.cuh) instead of gathering them with ld.global.  Every TriPlane render golden must still match, and
the kernel must have staged a part of the patches through the TMA (direct_patches < 12 per tile)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs, load_golden
assert os.environ.get("NGF_COLOUR_TMA") == "1"
ok = True
for case in K.CASES:
    if case.variant != "triplane":
        continue
    gold = load_golden(case.name)
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    for width in (0, 64):
        out = f(rays.cuda(), white_bg=case.white_bg, N_samples=case.n_samples, image_width=width if rays.shape[0] == 4096 else 0,
                **forward_kwargs(case))
        e = float(np.abs(out["rgb_map"].cpu().numpy() - gold["rgb"]).max()); d = float(np.abs(out["depth_map"].cpu().numpy() - gold["depth"]).max())
        st = f.last_stats()
        print(f"{case.name} image_width={width}: rgb {e:.2e} depth {d:.2e} tiles {st['mlp_tiles']} direct patches {st['direct_patches']}")
        ok = ok and e < 1e-3 and d < 2e-3
# the goldens are 64x64-ray frames (12.5 pixels of the 800x800 frame per ray): no 32-sample group fits a 5x5 patch there, so
# they exercise the direct-gather fallback.  A 96x64-pixel crop of the full-resolution frame has the bench's footprints:
# patches are staged through the TMA; compare with the oracle.
from oracle import restate_field as R
from helpers import oracle_spec
case = K.Case("c2_fog_crop", kind="fog", config="C2", n_samples=192)
state, kw, occ, rays = K.build_inputs(case)
crop = rays.view(800, 800, 6)[352:416, 352:448].reshape(-1, 6).contiguous()
f = build_cuda_field(case, state, kw, occ)
out = f(crop.cuda(), white_bg=True, N_samples=192, image_width=96, **forward_kwargs(case))
o_rgb, o_depth = R.render(oracle_spec(case, state, kw, occ), crop, white_bg=True, N_samples=192)
e = float((out["rgb_map"].cpu() - o_rgb).abs().max()); d = float((out["depth_map"].cpu() - o_depth).abs().max())
st = f.last_stats()
print(f"c2_fog_crop: rgb {e:.2e} depth {d:.2e} tiles {st['mlp_tiles']} direct patches {st['direct_patches']} of {12 * st['mlp_tiles']}")
ok = ok and e < 1e-3 and d < 2e-3 and st["mlp_tiles"] > 0 and st["direct_patches"] < 0.9 * 12 * st["mlp_tiles"]
sys.exit(0 if ok else 1)
