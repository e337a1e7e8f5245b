"""TEST INFRASTRUCTURE — generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (imported from
``/root/reference`` through ``oracle/ref_loader.py``) on the seeded cases of ``oracle/cases.py``.

Run in the build container only:   python -m oracle.make_golden [case ...]

The reference ships no tests or golden vectors, so these files are the parity pin that travels
to the GPU box: each holds the reference's ``rgb_map`` / ``depth_map`` for one case, the per-case sample counts, a
SHA-256 of the regenerated inputs (so a drifted generator is detected instead of silently compared), and the torch
version that produced them.  ``pointwise_*.npz`` hold the reference's ``compute_gauge`` / ``compute_density`` /
``compute_rgb`` / ``sample_alpha`` / ``sample_ray`` outputs on seeded points.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import cases as K
from . import ref_loader


def build_reference_field(case: K.Case):
    """Instantiate the reference's model class for ``case`` on the CPU and load the synthetic state into it."""
    state, kw, occ, rays = K.build_inputs(case)
    if case.variant == "triplane":
        pkg = ref_loader.triplane_models()
        extra = dict(gauge_start=0)
    else:
        pkg = ref_loader.infoinv_models()
        extra = {}
    field = ref_loader.quiet(pkg.Field.TriPlane, kw["aabb"], kw["gridSize"], "cpu", near_far=kw["near_far"],
                             step_ratio=kw["step_ratio"], distance_scale=kw["distance_scale"],
                             rayMarch_weight_thres=kw["rayMarch_weight_thres"], alphaMask_thres=kw["alphaMask_thres"],
                             **extra)
    K.synth.load_into(field, state)
    if occ is not None:
        field.alphaMask = pkg.FieldBase.AlphaGridMask("cpu", K.mask_aabb(), occ)
    return field, state, kw, occ, rays


def forward_kwargs(case: K.Case) -> dict:
    if case.variant == "triplane":
        return {"iteration": 30001 if case.gauge_on else -1}      # Field.py:58 with gauge_start=0
    return {"infoinv": case.infoinv}


@torch.no_grad()
def make_case(case: K.Case) -> str:
    field, state, kw, occ, rays = build_reference_field(case)
    rgb, depth = ref_loader.reference_renderer(rays, field, chunk=4096, N_samples=case.n_samples,
                                               white_bg=case.white_bg, **forward_kwargs(case))
    path = K.golden_path(case.name)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, rgb=rgb.numpy().astype(np.float32), depth=depth.numpy().astype(np.float32),
                        n_samples=np.int64(field.nSamples), step_size=np.float32(field.stepSize.item()),
                        fingerprint=K.fingerprint(state, rays, occ), torch_version=torch.__version__)
    return path


@torch.no_grad()
def make_pointwise(variant: str) -> str:
    case = K.Case(f"pointwise_{variant}", variant=variant, kind="rand")
    field, state, kw, occ, rays = build_reference_field(case)
    xyz, dirs, world = K.pointwise_inputs()
    out = {}
    if variant == "triplane":
        xy, yz, xz = field.compute_gauge(xyz.clone(), iteration=1)
        out["sigma"] = field.compute_density(xy, yz, xz).numpy()
        out["rgb"] = field.compute_rgb(xy, yz, xz, dirs).numpy()
        xy0, yz0, xz0 = field.compute_gauge(xyz.clone(), iteration=-1)
        out.update(xy0=xy0.numpy(), yz0=yz0.numpy(), xz0=xz0.numpy())
    else:
        xy, yz, xz = field.transform(xyz.clone())
        out["sigma"] = field.compute_density(xy, yz, xz, infoinv=True).numpy()
        out["rgb"] = field.compute_rgb(xy, yz, xz, dirs, infoinv=True).numpy()
        out["sigma_noinv"] = field.compute_density(xy, yz, xz, infoinv=False).numpy()
        out["rgb_noinv"] = field.compute_rgb(xy, yz, xz, dirs, infoinv=False).numpy()
    out.update(xy=xy.numpy(), yz=yz.numpy(), xz=xz.numpy())
    out["alpha_keep"] = (field.alphaMask.sample_alpha(world) > 0).numpy()
    out["alpha"] = field.compute_alpha(world, field.stepSize).numpy()          # FieldBase.py:140-159
    pts, t, inside = field.sample_ray(rays[:64, :3], rays[:64, 3:6], is_train=False, N_samples=48)
    out.update(march_pts=pts.numpy(), march_t=t.numpy(), march_inside=inside.numpy())
    path = K.golden_path(case.name)
    np.savez_compressed(path, fingerprint=K.fingerprint(state, rays, occ), torch_version=torch.__version__,
                        **{k: np.asarray(v) for k, v in out.items()})
    return path


def main(argv):
    if not ref_loader.available():
        raise SystemExit("reference tree not found: golden vectors can only be generated in the build container")
    names = argv or [c.name for c in K.CASES] + ["pointwise_triplane", "pointwise_infoinv"]
    for n in names:
        if n.startswith("pointwise_"):
            p = make_pointwise(n.split("_", 1)[1])
        else:
            p = make_case(K.CASE_BY_NAME[n])
        print(f"{n}: {os.path.getsize(p) / 1024:.1f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
