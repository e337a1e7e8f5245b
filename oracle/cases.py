"""TEST INFRASTRUCTURE — the parity cases shared by ``oracle/make_golden.py`` (which runs the imported reference
on them in the build container) and ``tests/`` (which run the oracle restatement and the CUDA path on them).

Every case is regenerated from seeds by ``ngf_b200.synth``; a golden file stores only the reference's outputs and a
checksum of the inputs, so fixtures stay small (tests/golden/*.npz).
"""
from __future__ import annotations

import functools
import hashlib
import importlib.util
import os
from dataclasses import dataclass, field as dc_field
from typing import Optional, Tuple

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_synth():
    # loaded by path so that importing the cases never imports the product package (and its CUDA library)
    spec = importlib.util.spec_from_file_location("_ngf_synth", os.path.join(_ROOT, "neural-gauge-fields_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synth = _load_synth()


@dataclass
class Case:
    name: str
    variant: str = "triplane"            # "triplane" | "infoinv"
    kind: str = "hull"                   # synth.field_state kind
    config: str = "C1"
    pose: int = 0
    n_samples: int = 64                  # forward(N_samples=...); -1 = field.nSamples
    white_bg: bool = True
    mask: bool = True                    # attach an alpha mask
    gauge_on: bool = True                # TriPlane: iteration >= gauge_start
    infoinv: bool = True                 # InfoInv: forward(infoinv=...)
    res: Tuple[int, int, int] = (256, 256, 256)       # plane resolution (X, Y, Z)
    aabb: Optional[list] = None          # model box (default synth.AABB); the mask keeps the default box
    step_ratio: Optional[float] = None
    seed: int = 1234
    ray_cols: int = 6                    # 8: two extra columns (the last one feeds the depth background term)
    max_rays: int = 0                    # > 0: keep a strided subset of the frame's rays
    extra: dict = dc_field(default_factory=dict)


CASES = [
    Case("tp_hull_c1"),
    Case("tp_fog_c1", kind="fog"),
    Case("tp_rand_c1", kind="rand"),
    Case("tp_hull_nomask", mask=False),
    Case("tp_hull_nogauge", gauge_on=False, pose=3),
    Case("tp_fog_blackbg", kind="fog", white_bg=False, pose=5),
    Case("tp_fog_full_march", kind="fog", n_samples=-1, step_ratio=3.0, pose=7, max_rays=1024),
    # after up_sampling / shrink (Field.py:108-132): non-square planes, tighter model box, mask keeps its own box
    Case("tp_fog_shrunk", kind="fog", res=(200, 168, 232), aabb=[[-1.1, -0.9, -1.2], [1.0, 0.95, 1.25]], pose=11),
    Case("tp_rand_cols8", kind="rand", ray_cols=8, pose=2, max_rays=2048),
    Case("ii_hull_c1", variant="infoinv"),
    Case("ii_fog_c1", variant="infoinv", kind="fog", pose=4),
    Case("ii_rand_noinv", variant="infoinv", kind="rand", infoinv=False, pose=6, max_rays=2048),
    Case("ii_fog_shrunk", variant="infoinv", kind="fog", res=(180, 256, 140), aabb=[[-1.0, -1.2, -0.8], [1.2, 1.0, 1.1]],
         pose=9, white_bg=False),
]
CASE_BY_NAME = {c.name: c for c in CASES}

# forward(is_train=True): jittered sampling (FieldBase.py:128-130), one 4096-ray chunk each so that the reference draws
# its random numbers exactly once (u [R,1], then the background coin when white_bg is False).  Forward only.
TRAIN_SEED = 4242
TRAIN_CASES = [
    Case("train_tp_hull"),
    Case("train_tp_fog_blackbg", kind="fog", white_bg=False, pose=5),
    Case("train_ii_fog", variant="infoinv", kind="fog", pose=4),
]
TRAIN_BY_NAME = {c.name: c for c in TRAIN_CASES}

# one training step's loss and gradients (main.py:272-283) on the same chunks: the pin of the backward-pass oracle
GRAD_CASES = {"grad_tp_hull": "train_tp_hull", "grad_ii_fog": "train_ii_fog"}
GRAD_TARGET_SEED = 99
GRAD_SAMPLES = 4096            # entries of every large gradient tensor kept in the golden file


def grad_target(n_rays: int) -> torch.Tensor:
    """Synthetic rgb_train: the gradients do not care what the target is, only that both sides use the same."""
    return torch.rand((n_rays, 3), generator=torch.Generator().manual_seed(GRAD_TARGET_SEED))


@functools.lru_cache(maxsize=8)
def _field_state(variant, kind, seed, res):
    return synth.field_state(variant, kind, seed=seed, res=res)       # treated as read-only by every user


@functools.lru_cache(maxsize=4)
def _occupancy(kind):
    return synth.occupancy_volume(kind)


def build_inputs(case: Case):
    """-> (state_dict, ctor kwargs, occupancy volume or None, rays [R, ray_cols])."""
    kw = synth.field_kwargs(case.config)
    if case.step_ratio is not None:
        kw["step_ratio"] = case.step_ratio
    grid = list(case.res)
    kw["gridSize"] = grid
    if case.aabb is not None:
        kw["aabb"] = torch.tensor(case.aabb, dtype=torch.float32)
    state = _field_state(case.variant, case.kind, case.seed, tuple(case.res))
    occ = _occupancy(case.kind) if case.mask else None
    rays = synth.config_rays(case.config, case.pose)
    if case.max_rays and rays.shape[0] > case.max_rays:
        rays = rays[:: rays.shape[0] // case.max_rays][: case.max_rays].contiguous()
    if case.ray_cols > 6:
        g = torch.Generator().manual_seed(case.seed + 1)
        rays = torch.cat([rays, torch.rand((rays.shape[0], case.ray_cols - 6), generator=g) * 4 + 2], 1).contiguous()
    return state, kw, occ, rays


def mask_aabb() -> torch.Tensor:
    """The alpha mask always covers the default box (it is built before ``shrink`` in the reference's training
    schedule, main.py:333-342), so shrunk cases exercise a mask box different from the model box."""
    return torch.tensor(synth.AABB, dtype=torch.float32)


def fingerprint(state, rays, occ) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(np.ascontiguousarray(state[k].numpy()).tobytes())
    h.update(np.ascontiguousarray(rays.numpy()).tobytes())
    if occ is not None:
        # {0,1} occupancy volumes hash as bits (the same digest as before); other tensors (textures) as raw bytes
        binary = bool(((occ == 0) | (occ == 1)).all())
        h.update(np.packbits(occ.numpy() > 0).tobytes() if binary else np.ascontiguousarray(occ.numpy()).tobytes())
    return h.hexdigest()


def golden_path(name: str) -> str:
    return os.path.join(_ROOT, "tests", "golden", f"{name}.npz")


# point-wise fixture: inputs for compute_gauge / compute_density / compute_rgb / sample_alpha / sample_ray
def pointwise_inputs(n: int = 2048, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand((n, 3), generator=g) * 2.2 - 1.1          # some points outside [-1,1]^3 (zero padding)
    xyz[:16] = torch.tensor([-1.0, 0.0, 1.0])[torch.randint(0, 3, (16, 3), generator=g)]   # exact lattice points
    dirs = torch.randn((n, 3), generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    world = (torch.rand((n, 3), generator=g) * 2 - 1) * 1.6
    lattice = torch.linspace(-1.5, 1.5, 256)
    world[:64] = lattice[torch.randint(0, 256, (64, 3), generator=g)]                      # exact voxel centres
    return xyz.contiguous(), dirs.contiguous(), world.contiguous()


# ---------------------------------------------------------------------------------------------------------------
# UV-Mapping (NeuTex) cases
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class NeutexCase:
    name: str
    pose: int = 0
    max_rays: int = 2048
    background: Optional[Tuple[float, float, float]] = (1.0, 1.0, 1.0)
    texture_channels: int = 0            # 0: learned texture; 3 / 4: synthetic edited texture (target_texture branch)
    seed: int = 0
    gain: float = 1.0
    noise_seed: int = 11
    sample_num: int = 64                 # opt.sample_num
    primitive: str = "square"          # 'sphere': 3-d unit-vector uv (gauge_fields.py:55-56,71-74)


NEUTEX_CASES = [
    NeutexCase("neutex_white"),
    NeutexCase("neutex_black", pose=5, background=(0.0, 0.0, 0.0), seed=2, max_rays=1024),
    NeutexCase("neutex_nobg", pose=9, background=None, max_rays=777, noise_seed=4),
    NeutexCase("neutex_texture_rgb", pose=3, texture_channels=3, max_rays=1024),
    NeutexCase("neutex_texture_rgba", pose=7, texture_channels=4, max_rays=1024, background=(0.2, 0.4, 0.6)),
    NeutexCase("neutex_sphere", pose=2, max_rays=1024, seed=3, primitive="sphere"),
    NeutexCase("neutex_s48", pose=6, max_rays=768, seed=1, sample_num=48, noise_seed=9),
]
NEUTEX_BY_NAME = {c.name: c for c in NEUTEX_CASES}


@functools.lru_cache(maxsize=4)
def _neutex_state(seed, gain, primitive="square"):
    return synth.neutex_state(seed, gain, primitive)


def build_neutex_inputs(case: NeutexCase):
    """-> (state_dict, texture or None, campos [1,3], raydir [1,R,3], background [1,3] or None, noise [1,R,64])."""
    state = _neutex_state(case.seed, case.gain, case.primitive)
    tex = synth.neutex_texture(channels=case.texture_channels) if case.texture_channels else None
    campos, raydir = synth.neutex_camera(case.pose, max_rays=case.max_rays)
    bg = None if case.background is None else torch.tensor([case.background], dtype=torch.float32)
    noise = synth.neutex_noise(raydir.shape[1], samples=case.sample_num, seed=case.noise_seed)
    return state, tex, campos, raydir, bg, noise
