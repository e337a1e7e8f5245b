"""Parameter containers with the reference's module / parameter names, so ``state_dict`` round-trips
(``rgb_decoder.basis.weight``, ``rgb_decoder.mlp.{0,2,4}.{weight,bias}``, ``density_decoder.mlp.{0,2,4}.*``).

Mirrors ``TriPlane/models/networks.py:12-32`` (rgb_decoder) and ``InfoInv/models/networks.py:34-54``
(density_decoder).  The arithmetic of these networks runs inside the fused CUDA kernels
(``csrc/ngf_mlp.cuh``, ``csrc/ngf_common.cuh``); these classes only own and initialise the weights.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class rgb_decoder(nn.Module):
    """basis (bias-free feat_dim x feat_dim) then MLP (feat_dim+3+2*view_pe*3) -> middle -> middle -> 3.
    Reference: TriPlane/models/networks.py:12-23 (default nn.Linear init, last bias zeroed)."""

    def __init__(self, feat_dim: int, view_pe: int = 6, middle_dim: int = 128):
        super().__init__()
        self.input_dim = feat_dim + 3 + 2 * view_pe * 3
        self.view_pe = view_pe
        self.basis = nn.Linear(feat_dim, feat_dim, bias=False)
        self.mlp = nn.Sequential(nn.Linear(self.input_dim, middle_dim), nn.ReLU(inplace=True),
                                 nn.Linear(middle_dim, middle_dim), nn.ReLU(inplace=True),
                                 nn.Linear(middle_dim, 3))
        nn.init.constant_(self.mlp[-1].bias, 0)

    def forward(self, features, view_dirs):
        raise NotImplementedError(
            "rgb_decoder runs fused with the plane gather on the GPU: call field.compute_rgb(xy, yz, xz, viewdirs)")


class density_decoder(nn.Module):
    """feat_dim -> middle -> middle -> 1 (InfoInv/models/networks.py:34-45)."""

    def __init__(self, feat_dim: int, middle_dim: int = 32):
        super().__init__()
        self.input_dim = feat_dim
        self.mlp = nn.Sequential(nn.Linear(feat_dim, middle_dim), nn.ReLU(inplace=True),
                                 nn.Linear(middle_dim, middle_dim), nn.ReLU(inplace=True),
                                 nn.Linear(middle_dim, 1))
        nn.init.constant_(self.mlp[-1].bias, 0)

    def forward(self, features):
        raise NotImplementedError(
            "density_decoder runs fused with the plane gather on the GPU: call field.compute_density(xy, yz, xz)")


def xavier_uniform_linear_(m: nn.Linear, gain: float = 1.0) -> None:
    """The reference's home-grown xavier init for the TriPlane density head
    (TriPlane/models/networks.py:143-156,169-190: uniform(+-gain*sqrt(2/(n_in+n_out))*sqrt(3)), bias 0)."""
    std = gain * math.sqrt(2.0 / (m.in_features + m.out_features))
    with torch.no_grad():
        m.weight.uniform_(-std * math.sqrt(3.0), std * math.sqrt(3.0))
        if m.bias is not None:
            m.bias.zero_()
