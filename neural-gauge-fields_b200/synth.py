"""Seeded synthetic inputs for tests and ``bench.py`` (no dataset or checkpoint exists offline).

Everything here is plain CPU torch with explicit generators, so the same seed gives the same tensors in the build
container and on the GPU box.  Nothing in this file computes the render path.

  camera_rays        Synthetic-NeRF-like pinhole camera on a sphere looking at the origin -> rays [H*W, 6]
                     (layout of the reference's dataset classes: origin xyz, unit direction xyz;
                     TriPlane/dataLoader/ray_utils.py:24-42,66-87, blender.py:46-52)
  field_state        a ``state_dict`` in the reference's parameter names for a scene-like field:
                       "hull"  visual-hull solid (sparse regime: < 1 colour sample / ray on average)
                       "fog"   same silhouette, low density (dense-ish regime: tens of colour samples / ray)
                       "rand"  random planes, moderate density (adversarial for mask decisions)
  occupancy_volume   {0,1} alpha-mask volume for those fields (the silhouettes' intersection, dilated)
  field_kwargs       constructor kwargs (aabb, gridSize, step_ratio ...) for the BASELINE.json configurations
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------------------
# configurations (C1 = 64x64 rays / 64 samples, C2 = 800x800 / 192 samples)
# ---------------------------------------------------------------------------------------------------------------
AABB = [[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]]
NEAR_FAR = [2.0, 6.0]
CAM_RADIUS = 4.031
FOCAL_800 = 0.5 * 800 / math.tan(0.5 * 0.6911112)      # blender.py:46 with camera_angle_x of the lego scene

CONFIGS = {
    # name: (H, W, focal, gridSize, step_ratio, N_samples)
    "C1": (64, 64, FOCAL_800 * 64 / 800, (256, 256, 256), 7.0, 64),
    "C2": (800, 800, FOCAL_800, (256, 256, 256), 2.305, 192),
}


def field_kwargs(config: str = "C2") -> dict:
    _, _, _, grid, step_ratio, _ = CONFIGS[config]
    return dict(aabb=torch.tensor(AABB, dtype=torch.float32), gridSize=list(grid), near_far=list(NEAR_FAR),
                step_ratio=step_ratio, distance_scale=25, rayMarch_weight_thres=1e-4, alphaMask_thres=1e-4)


# ---------------------------------------------------------------------------------------------------------------
# cameras
# ---------------------------------------------------------------------------------------------------------------
def pose_angles(index: int, seed: int = 0) -> Tuple[float, float]:
    """(azimuth, elevation) in radians of pose ``index`` of the seeded 200-pose test list."""
    g = torch.Generator().manual_seed(seed)
    az = torch.rand(200, generator=g) * 2 * math.pi
    el = (torch.rand(200, generator=g) * 50.0 + 10.0) * math.pi / 180.0
    return float(az[index % 200]), float(el[index % 200])


def look_at_c2w(az: float, el: float, radius: float = CAM_RADIUS) -> torch.Tensor:
    """OpenCV-convention camera-to-world [3,4]: +z looks at the origin, +y is image-down."""
    eye = torch.tensor([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az),
                        radius * math.sin(el)], dtype=torch.float64)
    fwd = -eye / eye.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    return torch.stack([right, down, fwd, eye], 1).float()


def camera_rays(H: int, W: int, focal: float, pose: int = 0, seed: int = 0) -> torch.Tensor:
    """Row-major pixel rays [H*W, 6] fp32: directions through pixel centres ((i+0.5-W/2)/f, (j+0.5-H/2)/f, 1),
    rotated by c2w and normalised (ray_utils.py:34-40,80-87; blender.py:52)."""
    az, el = pose_angles(pose, seed)
    c2w = look_at_c2w(az, el)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    d = torch.stack([(i + 0.5 - W / 2) / focal, (j + 0.5 - H / 2) / focal, torch.ones_like(i)], -1)
    d = d @ c2w[:, :3].T
    d = d / d.norm(dim=-1, keepdim=True)
    o = c2w[:, 3].expand_as(d)
    return torch.cat([o, d], -1).reshape(-1, 6).contiguous()


def config_rays(config: str = "C2", pose: int = 0) -> torch.Tensor:
    H, W, focal, _, _, _ = CONFIGS[config]
    return camera_rays(H, W, focal, pose)


# ---------------------------------------------------------------------------------------------------------------
# fields
# ---------------------------------------------------------------------------------------------------------------
def _silhouette(res_h: int, res_w: int, kind: int) -> torch.Tensor:
    """Smooth {~-1 outside, ~+1 inside} 2-D shape on [-1,1]^2: a disc, a rounded box and a two-lobe blob, so the
    three-plane intersection is not symmetric."""
    v, u = torch.meshgrid(torch.linspace(-1, 1, res_h), torch.linspace(-1, 1, res_w), indexing="ij")
    if kind == 0:
        s = torch.sqrt(u * u + v * v) / 0.55
    elif kind == 1:
        s = ((u.abs() / 0.5) ** 4 + (v.abs() / 0.42) ** 4) ** 0.25
    else:
        a = torch.sqrt((u - 0.18) ** 2 + (v + 0.1) ** 2) / 0.4
        b = torch.sqrt((u + 0.22) ** 2 + (v - 0.12) ** 2) / 0.36
        s = torch.minimum(a, b)
    return torch.tanh(12.0 * (1.0 - s))


def _smooth_noise(g: torch.Generator, c: int, h: int, w: int, amp: float, k: int) -> torch.Tensor:
    x = amp * torch.randn((1, c, h, w), generator=g)
    return F.avg_pool2d(x, k, stride=1, padding=k // 2, count_include_pad=False)


def field_state(variant: str = "triplane", kind: str = "hull", seed: int = 1234, res: Tuple[int, int, int] = (256,) * 3,
                gauge_res: int = 256) -> Dict[str, torch.Tensor]:
    """state_dict (reference parameter names) of a synthetic field.  ``res`` = (X, Y, Z) plane resolution:
    plane_xy is [1,C,Y,X], plane_yz [1,C,Z,Y], plane_xz [1,C,Z,X] (Field.py:19-21,108-114)."""
    assert variant in ("triplane", "infoinv") and kind in ("hull", "fog", "rand")
    g = torch.Generator().manual_seed(seed)
    C, DC = (64, 16) if variant == "triplane" else (96, 24)
    AC = C - DC
    X, Y, Z = res
    shapes = {"plane_xy": (Y, X), "plane_yz": (Z, Y), "plane_xz": (Z, X)}
    st: Dict[str, torch.Tensor] = {}
    for n, (name, (h, w)) in enumerate(shapes.items()):
        p = 3.0 * _smooth_noise(g, C, h, w, 0.3, 5)
        if kind != "rand":
            p[0, 0 if variant == "triplane" else 12] = _silhouette(h, w, n)
        st[name] = p.contiguous()
    if variant == "triplane":
        for name in ("gauge_xy", "gauge_yz", "gauge_xz"):
            st[name] = (4.0 * _smooth_noise(g, 2, gauge_res, gauge_res, 0.05, 9)).contiguous()

    def linear(out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand((out_f, in_f), generator=g) * 2 - 1) * bound
        b = (torch.rand((out_f,), generator=g) * 2 - 1) * bound if bias else None
        return w, b

    Fd = 3 * AC
    st["rgb_decoder.basis.weight"], _ = linear(Fd, Fd, bias=False)
    for idx, (o, i) in zip((0, 2, 4), ((64, Fd + 15), (64, 64), (3, 64))):
        w, b = linear(o, i)
        st[f"rgb_decoder.mlp.{idx}.weight"], st[f"rgb_decoder.mlp.{idx}.bias"] = w * (2.0 if idx else 1.0), b
    if variant == "triplane":
        w, _ = linear(1, 3 * DC)
        w = 0.5 * w
        if kind == "hull":
            w[0, 0] = w[0, DC] = w[0, 2 * DC] = 10.0
            b = torch.tensor([-10.0])
        elif kind == "fog":
            w[0, 0] = w[0, DC] = w[0, 2 * DC] = 1.5
            b = torch.tensor([5.0])
        else:
            w = 4.0 * w
            b = torch.tensor([9.0])
        st["density_decoder.weight"], st["density_decoder.bias"] = w, b
    else:
        # 72 -> 32 -> 32 -> 1 as a visual hull: the silhouette sits in channel 12 of each plane's density block
        # (the cos(x*2^0) slot of the 4-band phase code, positive over the whole box).  Hidden unit k of layer 1 is
        # an "outside silhouette k" indicator, unit 0 of layer 2 is relu(1 - sum) = "inside all three".
        w1, b1 = linear(32, 3 * DC)
        w2, b2 = linear(32, 32)
        w3, b3 = linear(1, 32)
        if kind != "rand":
            w1[:3] = 0.0
            b1[:3] = 0.0
            for k in range(3):
                w1[k, k * DC + 12] = -10.0
            w2[0] = 0.0
            w2[0, :3] = -1.0
            b2[0] = 1.0
            w3 = 0.1 * w3
            w3[0, 0] = 30.0 if kind == "hull" else 20.0
            b3 = torch.tensor([-10.0])
        else:
            b3 = torch.tensor([10.0])
        for idx, (w, b) in zip((0, 2, 4), ((w1, b1), (w2, b2), (w3, b3))):
            st[f"density_decoder.mlp.{idx}.weight"], st[f"density_decoder.mlp.{idx}.bias"] = w, b
    return {k: v.float().contiguous() for k, v in st.items()}


def occupancy_volume(kind: str = "hull", res: int = 256, dilate: int = 7) -> torch.Tensor:
    """{0,1} fp32 volume [D(z), H(y), W(x)] over the full box: intersection of the three silhouettes, dilated by a
    ``dilate``-wide max-pool (the gauge offsets move the surface by a few cells).  For kind == "rand" a seeded
    random blob pattern uncorrelated with the density (the adversarial case of SURVEY.md §7)."""
    if kind == "rand":
        g = torch.Generator().manual_seed(99)
        v = F.avg_pool3d(torch.randn((1, 1, res, res, res), generator=g), 5, stride=1, padding=2)
        return (v[0, 0] > 0.05).float().contiguous()
    sxy, syz, sxz = (_silhouette(res, res, n) > -0.5 for n in range(3))      # [y,x], [z,y], [z,x]
    vol = sxy[None, :, :] & syz[:, :, None] & sxz[:, None, :]
    vol = F.max_pool3d(vol.float()[None, None], dilate, stride=1, padding=dilate // 2)[0, 0]
    return vol.contiguous()


def load_into(field, state: Dict[str, torch.Tensor], occupancy: torch.Tensor | None = None, mask_cls=None):
    """Copy a synthetic state into a field module (reference class or drop-in): parameters by name, planes
    re-allocated when their shape differs, optional alpha mask over the field's box."""
    with torch.no_grad():
        for name in ("plane_xy", "plane_yz", "plane_xz", "gauge_xy", "gauge_yz", "gauge_xz"):
            if name in state and hasattr(field, name) and getattr(field, name).shape != state[name].shape:
                setattr(field, name, torch.nn.Parameter(torch.empty_like(state[name], device=getattr(field, name).device)))
        own = dict(field.named_parameters())
        for k, v in state.items():
            own[k].copy_(v.to(own[k].device))
    if occupancy is not None:
        dev = field.device
        field.alphaMask = mask_cls(dev, field.aabb.to(dev), occupancy.to(dev))
    return field


# ---------------------------------------------------------------------------------------------------------------
# UV-Mapping (NeuTex) synthetic inputs — BASELINE configs[3]
# ---------------------------------------------------------------------------------------------------------------
NEUTEX_H, NEUTEX_W = 600, 800                 # DTU test render, UV-Mapping/data/dtu.py (no_crop)
NEUTEX_FOCAL = (1446.2, 1441.6)               # scan83 intrinsics (in_camFocal.npy / in_camPrincpt.npy, rounded)
NEUTEX_PRINCPT = (411.6, 309.5)
NEUTEX_SAMPLES = 64                           # dtu_test.sh: --sample_num 64

# layer shapes (out, in) in the reference's module order
NEUTEX_LAYERS = {
    "net_geometry_decoder.block": [(256, 63)] + [(256, 256)] * 10 + [(1, 256)],          # decoder.py:201-217
    "net_texture.block1": [(256, 42)] + [(256, 256)] * 5,                                # decoder.py:20-26
    "net_texture.block2": [(256, 295)] + [(256, 256)] * 3 + [(3, 256)],                  # decoder.py:29-36
}


def neutex_state(seed: int = 0, gain: float = 1.0, primitive: str = "square") -> Dict[str, torch.Tensor]:
    """state_dict (reference NeuTex parameter names) with xavier-uniform weights like the reference's init_seq /
    init_weights (util.py:390-425) and small random biases.  ``primitive='sphere'``: the gauge network ends in 3 outputs
    (gauge_fields.py:55-56) and the texture network starts from [uv3, PE(uv3, 10)] = 63 inputs (model.py:22); those two
    layers are drawn from their own generator so the 'square' state of the same seed is unchanged."""
    g = torch.Generator().manual_seed(seed)
    st: Dict[str, torch.Tensor] = {}

    def lin(name, o, i, act_gain):
        bound = gain * act_gain * math.sqrt(6.0 / (i + o))
        st[name + ".weight"] = (torch.rand((o, i), generator=g) * 2 - 1) * bound
        st[name + ".bias"] = (torch.rand((o,), generator=g) * 2 - 1) * 0.05

    relu, leaky = math.sqrt(2.0), math.sqrt(2.0 / (1 + 0.2 ** 2))
    for prefix, shapes in NEUTEX_LAYERS.items():
        ag = relu if "geometry" in prefix else leaky
        for n, (o, i) in enumerate(shapes):
            lin(f"{prefix}.{2 * n}", o, i, ag if o > 3 else 1.0)
    lin("net_texture.color1", 3, 256, 1.0)
    enc = "gauge_transform.encoder"
    lin(f"{enc}.linear1", 64, 63, 1.0)
    lin(f"{enc}.linear2", 128, 64, 1.0)
    lin(f"{enc}.linear_list.0", 128, 128, 1.0)
    lin(f"{enc}.linear_list.1", 128, 128, 1.0)
    lin(f"{enc}.last_linear", 2, 128, 1.0)
    # a denser object in the middle of the cube so rays see structure: bias the density head
    st["net_geometry_decoder.block.22.bias"] = torch.tensor([1.5])
    if primitive == "sphere":
        g = torch.Generator().manual_seed(seed + 1000)
        lin(f"{enc}.last_linear", 3, 128, 1.0)
        lin("net_texture.block1.0", 256, 63, leaky)
    elif primitive != "square":
        raise ValueError(primitive)
    return {k: v.float().contiguous() for k, v in st.items()}


def neutex_camera(pose: int = 0, H: int = NEUTEX_H, W: int = NEUTEX_W, radius: float = 2.6, max_rays: int = 0):
    """-> campos [1,3], raydir [1,R,3] (unit), like DtuDataset.get_item (data/dtu.py:119-182): pixel-centre rays
    through a pinhole with the scan83 intrinsics, camera on a sphere looking at the unit cube."""
    az, el = pose_angles(pose, seed=3)
    c2w = look_at_c2w(az, el, radius)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    fx, fy = NEUTEX_FOCAL[0] * W / NEUTEX_W, NEUTEX_FOCAL[1] * H / NEUTEX_H
    # keep the object framed whatever the crop: principal point scaled with the image
    cx, cy = NEUTEX_PRINCPT[0] * W / NEUTEX_W, NEUTEX_PRINCPT[1] * H / NEUTEX_H
    d = torch.stack([(i + 0.5 - cx) / fx * 3.0, (j + 0.5 - cy) / fy * 3.0, torch.ones_like(i)], -1)   # x3: wider view
    d = d @ c2w[:, :3].T
    d = d / (d.norm(dim=-1, keepdim=True) + 1e-5)                                                       # dtu.py:35
    d = d.reshape(1, -1, 3)
    if max_rays and d.shape[1] > max_rays:
        d = d[:, :: d.shape[1] // max_rays][:, :max_rays]
    return c2w[:, 3].reshape(1, 3).contiguous(), d.contiguous()


def scan83_camera(view: int = 33, max_rays: int = 0):
    """The real DTU scan83 camera ``view`` as DtuDataset.get_item builds it for a full-image test render (data/dtu.py:119-182
    with no_crop: integer pixel coordinates, get_rays_dir dtu.py:27-37): -> campos [1,3], raydir [1, 600*800, 3].  The 64
    cameras ship with the reference (trainData/in_cam*.npy) and are kept in data/scan83_cameras.npz; 33 is the reference's
    centre camera (dtu.py:113-114)."""
    import os
    import numpy as np
    c = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "scan83_cameras.npz"))
    focal, princpt, rot = c["focal"][view], c["princpt"][view], c["extrinsics"][view][0:3, 0:3]
    px, py = np.meshgrid(np.arange(NEUTEX_W).astype(np.float32), np.arange(NEUTEX_H).astype(np.float32))
    x = (px - princpt[0]) / focal[0]
    y = (py - princpt[1]) / focal[1]
    dirs = np.stack([x, y, np.ones_like(x)], axis=-1)
    dirs = np.sum(rot[None, None, :, :] * dirs[..., None], axis=-2)
    dirs = dirs / (np.linalg.norm(dirs, axis=-1, keepdims=True) + 1e-5)
    d = torch.from_numpy(np.reshape(dirs, (-1, 3))).float().reshape(1, -1, 3)
    if max_rays and d.shape[1] > max_rays:
        d = d[:, :: d.shape[1] // max_rays][:, :max_rays]
    return torch.from_numpy(c["campos"][view]).float().reshape(1, 3).contiguous(), d.contiguous()


def neutex_noise(n_rays: int, samples: int = NEUTEX_SAMPLES, seed: int = 11) -> torch.Tensor:
    """The U[0,1) jitter tensor cube_ray_generation draws with torch.rand (renderer.py:113-118), made explicit so
    both implementations consume the same numbers."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((1, n_rays, samples), generator=g)


def neutex_texture(h: int = 96, w: int = 128, channels: int = 3, seed: int = 5) -> torch.Tensor:
    """A synthetic edited texture [h, w, channels] in [0,1] (stands in for UV-Mapping/data/texture*.png after
    load_square: flipped vertically, /255; util.py:270-274)."""
    g = torch.Generator().manual_seed(seed)
    v, u = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    base = torch.stack([0.5 + 0.5 * torch.sin(12 * u + 3 * c) * torch.cos(9 * v - c) for c in range(channels)], -1)
    return (0.8 * base + 0.2 * torch.rand((h, w, channels), generator=g)).clamp(0, 1).float().contiguous()
