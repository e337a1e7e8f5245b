"""Drop-in for the reference's ``InfoInv/models/Field.py`` ``TriPlane`` class (96-channel planes whose features are
multiplied by a sinusoidal phase code of the sample position, plus a 3-layer density MLP), executing on
hand-written sm_100a CUDA through ``libngf_b200.so``.

``from ngf_b200.infoinv import *`` gives the names ``InfoInv/main.py`` star-imports from ``models.Field``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .field_base import AlphaGridMask, Base
from .networks import density_decoder, rgb_decoder

__all__ = ["TriPlane", "AlphaGridMask", "Base", "rgb_decoder", "density_decoder"]


class TriPlane(Base):
    VARIANT = _lib.NGF_INFOINV

    def __init__(self, aabb, gridSize, device, **kargs):
        super().__init__(aabb, gridSize, device, **kargs)

    # Reference: InfoInv/models/Field.py:14-24
    def init_model(self, res=256, dim=96, scale=0.1, device=None):
        self.plane_xy = torch.nn.Parameter(scale * torch.randn((1, dim, res, res), device=device))
        self.plane_yz = torch.nn.Parameter(scale * torch.randn((1, dim, res, res), device=device))
        self.plane_xz = torch.nn.Parameter(scale * torch.randn((1, dim, res, res), device=device))
        self.density_dim = 24
        self.rgb_dim = dim - self.density_dim
        self.density_decoder = density_decoder(feat_dim=self.density_dim * 3, middle_dim=32).to(device)
        self.rgb_decoder = rgb_decoder(feat_dim=self.rgb_dim * 3, view_pe=2, middle_dim=64).to(device)
        self._infoinv_flag = True

    # Reference: InfoInv/models/Field.py:27-37
    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        return [{'params': self.plane_xy, 'lr': lr_init_spatialxyz},
                {'params': self.plane_yz, 'lr': lr_init_spatialxyz},
                {'params': self.plane_xz, 'lr': lr_init_spatialxyz},
                {'params': self.rgb_decoder.parameters(), 'lr': lr_init_network},
                {'params': self.density_decoder.parameters(), 'lr': lr_init_network}]

    def _fill_desc(self, d, keep):
        d.variant = _lib.NGF_INFOINV
        d.density_c = int(self.density_dim)
        self._fill_common(d, keep)
        d.gauge_on = 0
        layers = []
        for i in (0, 2, 4):
            m = self.density_decoder.mlp[i]
            w, b = m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous()
            keep += [w, b]
            layers.append(_lib.NgfLinear(w.data_ptr(), b.data_ptr(), m.in_features, m.out_features))
        d.dens_l1, d.dens_l2, d.dens_l3 = layers
        d.infoinv = 1

    def _grad_parameters(self):
        m, dm = self.rgb_decoder.mlp, self.density_decoder.mlp
        return [self.plane_xy, self.plane_yz, self.plane_xz, None, None, None,
                self.rgb_decoder.basis.weight, m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias,
                dm[0].weight, dm[0].bias, dm[2].weight, dm[2].bias, dm[4].weight, dm[4].bias]

    def _set_switches(self, lib, h, infoinv=True, **_):
        _lib.check(lib.ngf_field_set_infoinv(h, int(bool(infoinv))))

    def _apply_alpha_kw(self, infoinv=True, **_):
        _lib.check(_lib.load().ngf_field_set_infoinv(self._ensure_handle(), int(bool(infoinv))))

    def forward(self, rays_chunk, white_bg=True, is_train=False, N_samples=-1, infoinv=True, image_width=0, jitter=None):
        return super().forward(rays_chunk, white_bg=white_bg, is_train=is_train, N_samples=N_samples,
                               image_width=image_width, infoinv=infoinv, jitter=jitter)

    def feature2density(self, density_features, density_shift=-10):
        return F.softplus(density_features + density_shift)

    # Reference: InfoInv/models/Field.py:43-50
    def transform(self, valid_xyz):
        return self._coords(valid_xyz, gauge_on=False)

    # Reference: InfoInv/models/Field.py:52-70
    def compute_density(self, xy, yz, xz, infoinv=True):
        self._apply_alpha_kw(infoinv=infoinv)
        return self._density(xy, yz, xz)

    # Reference: InfoInv/models/Field.py:72-89
    def compute_rgb(self, xy, yz, xz, view_sampled, infoinv=True):
        self._apply_alpha_kw(infoinv=infoinv)
        return self._rgb(xy, yz, xz, view_sampled)

    # upstream TensoRF names listed by BASELINE.json's north_star (see triplane.py)
    def compute_densityfeature(self, xyz_sampled, infoinv=True):
        xy, yz, xz = self.transform(xyz_sampled)
        return self.compute_density(xy, yz, xz, infoinv)

    def compute_appfeature(self, xyz_sampled, viewdirs, infoinv=True):
        xy, yz, xz = self.transform(xyz_sampled)
        return self.compute_rgb(xy, yz, xz, viewdirs, infoinv)

    # Reference: InfoInv/models/Field.py:107-110
    def density_L1(self):
        return torch.mean(torch.abs(self.plane_xy)) + torch.mean(torch.abs(self.plane_yz)) \
            + torch.mean(torch.abs(self.plane_xz))
