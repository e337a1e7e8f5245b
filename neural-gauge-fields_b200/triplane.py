"""Drop-in for the reference's ``TriPlane/models/Field.py`` ``TriPlane`` class (learned gauge planes + 64-channel
feature planes), executing on hand-written sm_100a CUDA through ``libngf_b200.so``.

``from ngf_b200.triplane import *`` gives the names ``TriPlane/main.py:13`` star-imports from ``models.Field``
(``TriPlane``, ``AlphaGridMask``), so ``eval(args.model_name)(**kwargs)`` (main.py:37,226,230) keeps working.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .field_base import AlphaGridMask, Base
from .networks import rgb_decoder, xavier_uniform_linear_

__all__ = ["TriPlane", "AlphaGridMask", "Base", "rgb_decoder"]


class TriPlane(Base):
    VARIANT = _lib.NGF_TRIPLANE

    def __init__(self, aabb, gridSize, device, **kargs):
        super().__init__(aabb, gridSize, device, **kargs)

    # Reference: TriPlane/models/Field.py:17-32
    def init_model(self, res=256, dim=64, scale=0.1, device=None, gauge_start=0):
        self.plane_xy = torch.nn.Parameter(scale * torch.randn((1, dim, res, res), device=device))
        self.plane_yz = torch.nn.Parameter(scale * torch.randn((1, dim, res, res), device=device))
        self.plane_xz = torch.nn.Parameter(scale * torch.randn((1, dim, res, res), device=device))
        gauge_res = 256
        self.gauge_xy = torch.nn.Parameter(torch.zeros((1, 2, gauge_res, gauge_res), device=device))
        self.gauge_yz = torch.nn.Parameter(torch.zeros((1, 2, gauge_res, gauge_res), device=device))
        self.gauge_xz = torch.nn.Parameter(torch.zeros((1, 2, gauge_res, gauge_res), device=device))
        self.rgb_decoder = rgb_decoder(feat_dim=48 * 3, view_pe=2, middle_dim=64).to(device)
        self.density_decoder = torch.nn.Linear(16 * 3, 1).to(device)
        xavier_uniform_linear_(self.density_decoder)
        self.gauge_start = gauge_start

    # Reference: Field.py:34-46
    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        return [{'params': self.plane_xy, 'lr': lr_init_spatialxyz},
                {'params': self.plane_yz, 'lr': lr_init_spatialxyz},
                {'params': self.plane_xz, 'lr': lr_init_spatialxyz},
                {'params': self.rgb_decoder.parameters(), 'lr': lr_init_network},
                {'params': self.density_decoder.parameters(), 'lr': lr_init_network},
                {'params': self.gauge_xy, 'lr': lr_init_network * 0.1},
                {'params': self.gauge_yz, 'lr': lr_init_network * 0.1},
                {'params': self.gauge_xz, 'lr': lr_init_network * 0.1}]

    def _fill_desc(self, d, keep):
        d.variant = _lib.NGF_TRIPLANE
        d.density_c = 16
        self._fill_common(d, keep)
        for i, g in enumerate((self.gauge_xy, self.gauge_yz, self.gauge_xz)):
            t = g.detach().float().contiguous()
            keep.append(t)
            d.gauge[i] = t.data_ptr()
            d.gauge_h[i], d.gauge_w[i] = t.shape[2], t.shape[3]
        d.gauge_on = 1
        w = self.density_decoder.weight.detach().float().contiguous()
        b = self.density_decoder.bias.detach().float().contiguous()
        keep += [w, b]
        d.dens_l1 = _lib.NgfLinear(w.data_ptr(), b.data_ptr(), 48, 1)
        d.infoinv = 0

    def _grad_parameters(self):
        m = self.rgb_decoder.mlp
        return [self.plane_xy, self.plane_yz, self.plane_xz, self.gauge_xy, self.gauge_yz, self.gauge_xz,
                self.rgb_decoder.basis.weight, m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias,
                self.density_decoder.weight, self.density_decoder.bias]

    def _set_switches(self, lib, h, iteration=0, **_):
        # Field.py:58: the gauge offsets apply once iteration >= gauge_start (main.py:67 passes 30001 at eval)
        _lib.check(lib.ngf_field_set_gauge(h, int(iteration >= self.gauge_start)))

    def forward(self, rays_chunk, white_bg=True, is_train=False, N_samples=-1, iteration=0, image_width=0, jitter=None):
        return super().forward(rays_chunk, white_bg=white_bg, is_train=is_train, N_samples=N_samples,
                               image_width=image_width, iteration=iteration, jitter=jitter)

    # Reference: Field.py:48-50
    def feature2density(self, density_features, density_shift=-10):
        return F.softplus(density_features + density_shift)

    # Reference: Field.py:53-75
    def compute_gauge(self, valid_xyz, iteration=0):
        return self._coords(valid_xyz, gauge_on=iteration >= self.gauge_start)

    # Reference: Field.py:77-91
    def compute_density(self, xy, yz, xz):
        return self._density(xy, yz, xz)

    # Reference: Field.py:93-105
    def compute_rgb(self, xy, yz, xz, view_sampled):
        return self._rgb(xy, yz, xz, view_sampled)

    # Names of the upstream TensoRF API that BASELINE.json's north_star lists; this reference renamed them
    # (compute_density / compute_rgb, which already include feature2density and the rgb decoder).
    def compute_densityfeature(self, xyz_sampled, iteration=30001):
        """xyz_sampled: normalised coordinates [N,3] -> density [N] (compute_gauge + compute_density, Field.py:53-91)."""
        xy, yz, xz = self.compute_gauge(xyz_sampled, iteration)
        return self.compute_density(xy, yz, xz)

    def compute_appfeature(self, xyz_sampled, viewdirs, iteration=30001):
        """xyz_sampled [N,3] normalised, viewdirs [N,3] -> rgb [N,3] (compute_gauge + compute_rgb, Field.py:53-75,93-105)."""
        xy, yz, xz = self.compute_gauge(xyz_sampled, iteration)
        return self.compute_rgb(xy, yz, xz, viewdirs)

    # Reference: Field.py:108-114
    @torch.no_grad()
    def up_sampling(self, res):
        def up(p, size):
            return torch.nn.Parameter(F.interpolate(p.data, size=size, mode='bilinear', align_corners=True))
        self.plane_xy = up(self.plane_xy, (res[1], res[0]))
        self.plane_yz = up(self.plane_yz, (res[2], res[1]))
        self.plane_xz = up(self.plane_xz, (res[2], res[0]))
        self.init_para(res)

    # Reference: Field.py:117-132
    @torch.no_grad()
    def shrink(self, new_aabb):
        xyz_min, xyz_max = new_aabb
        t_l, b_r = (xyz_min - self.aabb[0]) / self.units, (xyz_max - self.aabb[0]) / self.units
        t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
        b_r = torch.stack([b_r, self.gridSize]).amin(0)
        self.plane_xy = torch.nn.Parameter(self.plane_xy.data[..., t_l[1]:b_r[1], t_l[0]:b_r[0]])
        self.plane_yz = torch.nn.Parameter(self.plane_yz.data[..., t_l[2]:b_r[2], t_l[1]:b_r[1]])
        self.plane_xz = torch.nn.Parameter(self.plane_xz.data[..., t_l[2]:b_r[2], t_l[0]:b_r[0]])
        newSize = b_r - t_l
        self.aabb = new_aabb
        self.init_para((newSize[0], newSize[1], newSize[2]))

    # Reference: Field.py:149-152
    def density_L1(self):
        return torch.mean(torch.abs(self.plane_xy)) + torch.mean(torch.abs(self.plane_yz)) \
            + torch.mean(torch.abs(self.plane_xz))
