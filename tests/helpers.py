"""Shared helpers of the parity tests."""
from __future__ import annotations

import numpy as np
import torch

from oracle import cases as K
from oracle import restate_field as R

RGB_TOL = 1e-3        # BASELINE.json north_star: per-pixel max abs < 1e-3 (fp32 reference)
DEPTH_TOL = 2e-3      # depth is sum(w * t) with t in [2, 6]: same relative budget


def load_golden(name: str):
    g = np.load(K.golden_path(name), allow_pickle=False)
    return {k: g[k] for k in g.files}


def oracle_spec(case: K.Case, state, kw, occ) -> R.FieldSpec:
    return R.spec_from_state(case.variant, state, alpha_volume=occ, alpha_aabb=K.mask_aabb() if occ is not None else None,
                             gauge_on=case.gauge_on, infoinv=case.infoinv, **kw)


def build_cuda_field(case: K.Case, state, kw, occ, device="cuda"):
    """The drop-in model class on the GPU with the synthetic state loaded (calls the C ABI lazily)."""
    import ngf_b200
    cls = ngf_b200.TriPlane if case.variant == "triplane" else ngf_b200.InfoInvTriPlane
    extra = dict(gauge_start=0) if case.variant == "triplane" else {}
    f = cls(kw["aabb"], kw["gridSize"], device, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
            distance_scale=kw["distance_scale"], rayMarch_weight_thres=kw["rayMarch_weight_thres"],
            alphaMask_thres=kw["alphaMask_thres"], **extra)
    K.synth.load_into(f, state)
    if occ is not None:
        f.alphaMask = ngf_b200.AlphaGridMask(f.device, K.mask_aabb(), occ.to(f.device))
    return f


def forward_kwargs(case: K.Case) -> dict:
    if case.variant == "triplane":
        return {"iteration": 30001 if case.gauge_on else -1}
    return {"infoinv": case.infoinv}


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """TriPlane/main.py:105-106."""
    mse = torch.mean((a - b) ** 2)
    return float(-10.0 * torch.log(mse) / np.log(10.0))
