"""TEST INFRASTRUCTURE — CPU (PyTorch fp32) restatement of the reference's TriPlane / InfoInv render path.

Never imported by the product package.  Follows, function by function (paths relative to /root/reference):

  march            TriPlane/models/FieldBase.py:118-137   (Base.sample_ray, eval branch)
  alpha_keep       TriPlane/models/FieldBase.py:33-40, 261-267 (AlphaGridMask.sample_alpha + use in forward)
  to_unit_cube     TriPlane/models/FieldBase.py:88-89     (Base.normalize_coord)
  gauge_coords     TriPlane/models/Field.py:53-75         (TriPlane.compute_gauge)
                   InfoInv/models/Field.py:43-50          (TriPlane.transform: identity split)
  sigma            TriPlane/models/Field.py:77-91, 48-50  (compute_density, feature2density)
                   InfoInv/models/Field.py:52-70 + InfoInv/models/networks.py:34-54
  colour           TriPlane/models/Field.py:93-105 + TriPlane/models/networks.py:12-32, 205-216
                   InfoInv/models/Field.py:72-89
  transmittance    TriPlane/models/FieldBase.py:12-19     (raw2alpha)
  render_chunk     TriPlane/models/FieldBase.py:251-312   (Base.forward, is_train=False)
                   InfoInv/models/FieldBase.py:228-282
  render           TriPlane/main.py:60-71                 (renderer chunk loop)

The arithmetic that is NOT in /root/reference is PyTorch's (the reference pins no version, README.md:23;
this container and the GPU box both have torch 2.11.0): ``F.grid_sample`` (bilinear, ``align_corners=True``,
zero padding), ``nn.Linear``, ``softplus``, ``cumprod``, ``sin/cos/exp/sigmoid``.  The restatement calls the same
torch ops in the same order on the decision-critical chain (t, p, bbox test, alpha-mask test) so masks are
bit-identical to the reference; ``bilinear_explicit`` / ``alpha_keep_bits`` restate the published grid_sample
algorithm gather-by-gather — that is the formulation the CUDA kernels implement — and are tested against
``F.grid_sample`` in ``tests/test_oracle.py``.

Like the reference, work is compacted with boolean masks per 4096-ray chunk (so CPU timings of this port are
comparable with the reference's own).  Parity pin: ``tests/test_oracle_vs_reference.py`` (reference imported
here) and ``tests/golden/*.npz`` (generated from the reference by ``oracle/make_golden.py``).
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class FieldSpec:
    """Everything the render path reads, as plain tensors / numbers (reference attribute names in comments)."""
    variant: str                      # "triplane" | "infoinv"
    aabb: torch.Tensor                # [2,3]            Base.aabb
    step_size: torch.Tensor           # 0-d fp32         Base.stepSize
    n_samples: int                    #                  Base.nSamples
    near: float                       #                  Base.near_far[0]
    far: float
    distance_scale: float             #                  Base.distance_scale
    weight_thres: float               #                  Base.rayMarch_weight_thres
    planes: List[torch.Tensor]        # 3 x [1,C,H,W]    plane_xy, plane_yz, plane_xz
    density_c: int                    # 16 (TriPlane) | 24 (InfoInv)
    basis_w: torch.Tensor             # [F,F]            rgb_decoder.basis.weight
    rgb_layers: List[Tuple[torch.Tensor, torch.Tensor]]       # rgb_decoder.mlp.{0,2,4}
    density_layers: List[Tuple[torch.Tensor, torch.Tensor]]   # density_decoder (1 layer) | density_decoder.mlp.{0,2,4}
    view_pe: int = 2
    density_shift: float = -10.0      # Field.py:48 default argument
    gauge: Optional[List[torch.Tensor]] = None   # 3 x [1,2,Hg,Wg]  gauge_xy, gauge_yz, gauge_xz
    gauge_on: bool = False            # iteration >= gauge_start (Field.py:58)
    infoinv: bool = True              # forward(..., infoinv=True), InfoInv only
    alpha_volume: Optional[torch.Tensor] = None  # [1,1,D,H,W] {0,1}   alphaMask.alpha_volume
    alpha_aabb: Optional[torch.Tensor] = None    # [2,3]               alphaMask.aabb
    stats: dict = dc_field(default_factory=dict)


def spec_from_module(m, *, iteration: int = 30001, infoinv: bool = True) -> FieldSpec:
    """Read a FieldSpec out of a reference ``TriPlane`` module (either sub-project) or a drop-in with the
    same attribute names."""
    is_info = not hasattr(m, "gauge_xy")
    if is_info:
        dl = [(m.density_decoder.mlp[i].weight.detach(), m.density_decoder.mlp[i].bias.detach()) for i in (0, 2, 4)]
        dc = int(m.density_dim)
    else:
        dl = [(m.density_decoder.weight.detach(), m.density_decoder.bias.detach())]
        dc = 16
    am = getattr(m, "alphaMask", None)
    return FieldSpec(
        variant="infoinv" if is_info else "triplane",
        aabb=m.aabb.detach().float().cpu(),
        step_size=m.stepSize.detach().float().cpu(),
        n_samples=int(m.nSamples),
        near=float(m.near_far[0]), far=float(m.near_far[1]),
        distance_scale=float(m.distance_scale), weight_thres=float(m.rayMarch_weight_thres),
        planes=[p.detach() for p in (m.plane_xy, m.plane_yz, m.plane_xz)],
        density_c=dc,
        basis_w=m.rgb_decoder.basis.weight.detach(),
        rgb_layers=[(m.rgb_decoder.mlp[i].weight.detach(), m.rgb_decoder.mlp[i].bias.detach()) for i in (0, 2, 4)],
        density_layers=dl,
        view_pe=int(m.rgb_decoder.view_pe),
        gauge=None if is_info else [g.detach() for g in (m.gauge_xy, m.gauge_yz, m.gauge_xz)],
        gauge_on=(not is_info) and iteration >= int(m.gauge_start),
        infoinv=infoinv,
        alpha_volume=None if am is None else am.alpha_volume.detach(),
        alpha_aabb=None if am is None else am.aabb.detach().float().cpu(),
    )


def grid_bookkeeping(aabb: torch.Tensor, grid_size, step_ratio: float):
    """Base.init_para (FieldBase.py:63-74): -> (stepSize 0-d fp32, nSamples).  Same torch ops, same order."""
    aabb = aabb.float().cpu()
    size = aabb[1] - aabb[0]
    grid = torch.LongTensor(list(grid_size))
    units = size / (grid - 1)
    step = torch.mean(units) * step_ratio
    diag = torch.sqrt(torch.sum(torch.square(size)))
    return step, int((diag / step).item()) + 1


def spec_from_state(variant: str, state: dict, *, aabb, gridSize, step_ratio=2.0, near_far=(2.0, 6.0),
                    distance_scale=25, rayMarch_weight_thres=1e-4, gauge_on=True, infoinv=True,
                    alpha_volume: Optional[torch.Tensor] = None, alpha_aabb: Optional[torch.Tensor] = None,
                    **_unused) -> FieldSpec:
    """Build a FieldSpec from a state_dict in the reference's parameter names plus the Base constructor arguments
    (no module needed, so it also works where /root/reference does not exist)."""
    aabb = torch.as_tensor(aabb, dtype=torch.float32).cpu()
    step, n_samples = grid_bookkeeping(aabb, gridSize, step_ratio)
    is_info = variant == "infoinv"
    if is_info:
        dl = [(state[f"density_decoder.mlp.{i}.weight"], state[f"density_decoder.mlp.{i}.bias"]) for i in (0, 2, 4)]
    else:
        dl = [(state["density_decoder.weight"], state["density_decoder.bias"])]
    vol = None
    if alpha_volume is not None:
        vol = alpha_volume.float().cpu().view(1, 1, *alpha_volume.shape[-3:])
    return FieldSpec(
        variant=variant, aabb=aabb, step_size=step, n_samples=n_samples,
        near=float(near_far[0]), far=float(near_far[1]), distance_scale=float(distance_scale),
        weight_thres=float(rayMarch_weight_thres),
        planes=[state[k].float().cpu() for k in ("plane_xy", "plane_yz", "plane_xz")],
        density_c=24 if is_info else 16,
        basis_w=state["rgb_decoder.basis.weight"].float().cpu(),
        rgb_layers=[(state[f"rgb_decoder.mlp.{i}.weight"].float().cpu(), state[f"rgb_decoder.mlp.{i}.bias"].float().cpu())
                    for i in (0, 2, 4)],
        density_layers=[(w.float().cpu(), b.float().cpu()) for w, b in dl],
        gauge=None if is_info else [state[k].float().cpu() for k in ("gauge_xy", "gauge_yz", "gauge_xz")],
        gauge_on=(not is_info) and bool(gauge_on), infoinv=bool(infoinv),
        alpha_volume=vol,
        alpha_aabb=None if vol is None else (aabb if alpha_aabb is None else torch.as_tensor(alpha_aabb).float().cpu()),
    )


def spec_to_device(spec: FieldSpec, device) -> FieldSpec:
    """The same spec with every tensor on ``device`` — the reference's normal device is the GPU (TriPlane/main.py:18);
    bench.py times this as the torch-CUDA baseline.  The restated ops are device-agnostic; the CPU is the parity pin."""
    import dataclasses
    mv = lambda x: x.to(device) if isinstance(x, torch.Tensor) else x
    kw = {}
    for fld in dataclasses.fields(spec):
        v = getattr(spec, fld.name)
        if isinstance(v, torch.Tensor):
            v = mv(v)
        elif isinstance(v, list):
            v = [tuple(mv(y) for y in x) if isinstance(x, tuple) else mv(x) for x in v]
        kw[fld.name] = v
    kw["stats"] = {}
    return FieldSpec(**kw)


# --------------------------------------------------------------------------------------------------------------
# §8f rank 1  camera rays (TriPlane/dataLoader/ray_utils.py:24-42,66-87; blender.py:46-52; main.py:155-159)
# --------------------------------------------------------------------------------------------------------------
def reference_rays(H: int, W: int, focal, c2w: torch.Tensor, center=None) -> torch.Tensor:
    """-> rays [H*W, 6]: get_ray_directions (pixel-centre directions; kornia.create_meshgrid restated as index grids
    with x fastest) / their norm (blender.py:52), get_rays (directions @ c2w[:3,:3].T, origin c2w[:3,3]).
    kornia is absent offline, so this function is pinned by reading, not by running the reference."""
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    i, j = xs + 0.5, ys + 0.5
    fx, fy = focal if isinstance(focal, (tuple, list)) else (focal, focal)
    cent = center if center is not None else [W / 2, H / 2]
    dirs = torch.stack([(i - cent[0]) / fx, (j - cent[1]) / fy, torch.ones_like(i)], -1)
    dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    rays_d = dirs @ c2w[:3, :3].T
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return torch.cat([rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)], 1).contiguous()


# --------------------------------------------------------------------------------------------------------------
# a2  ray march (FieldBase.py:118-137)
# --------------------------------------------------------------------------------------------------------------
def march(spec: FieldSpec, o: torch.Tensor, d: torch.Tensor, S: int, jitter=None):
    """-> p [R,S,3], t [R,S], inside [R,S] bool.  Every op is a separate fp32 rounding, as in eager torch.
    jitter [R,1] (is_train, FieldBase.py:128-130): rng = arange(S) + u, one u per ray."""
    lo, hi = spec.aabb[0], spec.aabb[1]
    safe_d = torch.where(d == 0, torch.full_like(d, 1e-6), d)
    ta = (hi - o) / safe_d
    tb = (lo - o) / safe_d
    t0 = torch.minimum(ta, tb).amax(-1).clamp(min=spec.near, max=spec.far)
    k = torch.arange(S, device=d.device)[None].float()
    if jitter is not None:
        k = k.repeat(d.shape[-2], 1)
        k += jitter.reshape(-1, 1)
    t = t0[:, None] + spec.step_size * k                     # mul, then add
    p = o[:, None, :] + d[:, None, :] * t[..., None]          # mul, then add
    outside = ((lo > p) | (p > hi)).any(dim=-1)
    return p, t, ~outside


# --------------------------------------------------------------------------------------------------------------
# a3  occupancy ("alpha mask") test (FieldBase.py:33-40, 261-267)
# --------------------------------------------------------------------------------------------------------------
def alpha_keep(spec: FieldSpec, pts: torch.Tensor) -> torch.Tensor:
    """pts [N,3] world -> bool [N]: trilinear sample of the {0,1} volume is > 0."""
    lo = spec.alpha_aabb[0]
    inv = 1.0 / (spec.alpha_aabb[1] - spec.alpha_aabb[0]) * 2          # FieldBase.py:29
    q = (pts - lo) * inv - 1
    v = F.grid_sample(spec.alpha_volume, q.view(1, -1, 1, 1, 3), align_corners=True).view(-1)
    return v > 0


def alpha_keep_bits(spec: FieldSpec, pts: torch.Tensor) -> torch.Tensor:
    """Same decision without interpolation arithmetic — the form the CUDA kernel uses on a bit-packed grid.

    ATen grid_sampler_3d (align_corners=True, zeros): i = ((q+1)/2)*(size-1); corner weights are products of
    (floor(i)+1-i) [always > 0] and (i-floor(i)) [> 0 iff i is not an integer]; out-of-range corners
    contribute 0.  With a {0,1} volume the sample is > 0 iff some in-range corner with non-zero weight is 1.
    """
    vol = spec.alpha_volume[0, 0]
    D, H, W = vol.shape
    lo = spec.alpha_aabb[0]
    inv = 1.0 / (spec.alpha_aabb[1] - spec.alpha_aabb[0]) * 2
    q = (pts - lo) * inv - 1
    keep = torch.zeros(pts.shape[0], dtype=torch.bool)
    idx = []
    for a, n in ((0, W), (1, H), (2, D)):
        i = ((q[:, a] + 1) / 2) * (n - 1)
        f = torch.floor(i)
        idx.append((f.long(), i != f, n))
    (x0, xf, W_), (y0, yf, H_), (z0, zf, D_) = idx
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ok = torch.ones_like(keep)
                if dx: ok &= xf
                if dy: ok &= yf
                if dz: ok &= zf
                x, y, z = x0 + dx, y0 + dy, z0 + dz
                ok &= (x >= 0) & (x < W_) & (y >= 0) & (y < H_) & (z >= 0) & (z < D_)
                bit = vol[z.clamp(0, D_ - 1), y.clamp(0, H_ - 1), x.clamp(0, W_ - 1)] > 0
                keep |= ok & bit
    return keep


# --------------------------------------------------------------------------------------------------------------
# a4  world -> [-1,1]^3 (FieldBase.py:88-89)
# --------------------------------------------------------------------------------------------------------------
def to_unit_cube(spec: FieldSpec, p: torch.Tensor) -> torch.Tensor:
    inv = 2.0 / (spec.aabb[1] - spec.aabb[0])                           # FieldBase.py:67
    return (p - spec.aabb[0]) * inv - 1


def bilinear(plane: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """plane [1,C,H,W], uv [N,2] (u -> W, v -> H) -> [N,C].  torch's grid_sample, as the reference calls it
    (Field.py:59-61,79-83,97-101)."""
    n = uv.shape[0]
    return F.grid_sample(plane, uv.view(1, n, 1, 2), align_corners=True).view(plane.shape[1], n).t()


def bilinear_explicit(plane: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """Gather-by-gather restatement of grid_sampler_2d (bilinear, align_corners=True, zeros) — what the CUDA
    gather implements: i = (c+1)/2*(size-1); taps floor/floor+1; weights (x1-ix)(y1-iy)...; OOB taps = 0."""
    _, C, H, W = plane.shape
    P = plane[0].permute(1, 2, 0)                                       # [H,W,C] channels-last
    ix = ((uv[:, 0] + 1) / 2) * (W - 1)
    iy = ((uv[:, 1] + 1) / 2) * (H - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    out = torch.zeros(uv.shape[0], C, dtype=plane.dtype)
    for xx, yy, w in ((x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)),
                      (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0))):
        ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        tap = P[yy.clamp(0, H - 1).long(), xx.clamp(0, W - 1).long()]
        out = out + torch.where(ok, w, torch.zeros_like(w))[:, None] * tap
    return out


# --------------------------------------------------------------------------------------------------------------
# a5  gauge transform (TriPlane/models/Field.py:53-75) / identity split (InfoInv/models/Field.py:43-50)
# --------------------------------------------------------------------------------------------------------------
def gauge_coords(spec: FieldSpec, n: torch.Tensor):
    """n [N,3] in [-1,1]^3 -> (xy', yz', xz') each [N,2]."""
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    xy, yz, xz = torch.stack([x, y], 1), torch.stack([y, z], 1), torch.stack([x, z], 1)
    if spec.variant != "triplane" or not spec.gauge_on:
        return xy, yz, xz
    gxy, gyz, gxz = (bilinear(g, c) for g, c in zip(spec.gauge, (xy, yz, xz)))
    # association order matters in fp32: (coord + own plane's offset) + neighbour plane's offset
    xy2 = torch.stack([(x + gxy[:, 0]) + gxz[:, 0], (y + gxy[:, 1]) + gyz[:, 0]], 1)
    yz2 = torch.stack([(y + gyz[:, 0]) + gxy[:, 1], (z + gyz[:, 1]) + gxz[:, 1]], 1)
    xz2 = torch.stack([(x + gxz[:, 0]) + gxy[:, 0], (z + gxz[:, 1]) + gyz[:, 1]], 1)
    return xy2, yz2, xz2


def phase_code(xyz: torch.Tensor, n_freq: int) -> torch.Tensor:
    """positional_encoding (networks.py:205-216 / InfoInv networks.py:227-237): [N,D] -> [N,2*D*F];
    layout [sin(x0*2^0..2^(F-1)), sin(x1*...), ..., cos(same order)]."""
    bands = 2 ** torch.arange(n_freq, device=xyz.device).float()
    a = (xyz[..., None] * bands).reshape(xyz.shape[0], -1)
    return torch.cat([torch.sin(a), torch.cos(a)], -1)


def _xyz_from_coords(xy, yz):
    return torch.cat([xy, yz[:, 1:]], -1)                                # InfoInv Field.py:54,74


# --------------------------------------------------------------------------------------------------------------
# a6  density (Field.py:77-91 | InfoInv Field.py:52-70)
# --------------------------------------------------------------------------------------------------------------
def sigma(spec: FieldSpec, xy, yz, xz) -> torch.Tensor:
    dc = spec.density_c
    feats = [bilinear(p[:, :dc], c) for p, c in zip(spec.planes, (xy, yz, xz))]
    if spec.variant == "infoinv":
        if spec.infoinv:
            pe = phase_code(_xyz_from_coords(xy, yz), 4)                 # 24 = density_c
            feats = [f * pe for f in feats]
        h = torch.cat(feats, -1)
        (w1, b1), (w2, b2), (w3, b3) = spec.density_layers
        h = torch.relu(F.linear(h, w1, b1))
        h = torch.relu(F.linear(h, w2, b2))
        raw = F.linear(h, w3, b3).reshape(-1)
    else:
        (w, b), = spec.density_layers
        raw = F.linear(torch.cat(feats, -1), w, b).reshape(-1)
    return F.softplus(raw + spec.density_shift)


# --------------------------------------------------------------------------------------------------------------
# a8/a9  colour (Field.py:93-105, networks.py:12-32 | InfoInv Field.py:72-89)
# --------------------------------------------------------------------------------------------------------------
def colour(spec: FieldSpec, xy, yz, xz, viewdir) -> torch.Tensor:
    dc = spec.density_c
    feats = [bilinear(p[:, dc:], c) for p, c in zip(spec.planes, (xy, yz, xz))]
    if spec.variant == "infoinv" and spec.infoinv:
        pe = phase_code(_xyz_from_coords(xy, yz), 12)                    # 72 = C - density_c
        feats = [f * pe for f in feats]
    f = F.linear(torch.cat(feats, -1), spec.basis_w)                     # bias-free basis
    h = torch.cat([f, viewdir, phase_code(viewdir, spec.view_pe)], -1)
    (w1, b1), (w2, b2), (w3, b3) = spec.rgb_layers
    h = torch.relu(F.linear(h, w1, b1))
    h = torch.relu(F.linear(h, w2, b2))
    return torch.sigmoid(F.linear(h, w3, b3))


# --------------------------------------------------------------------------------------------------------------
# a7  alpha / transmittance / weights (FieldBase.py:12-19)
# --------------------------------------------------------------------------------------------------------------
def transmittance(sig: torch.Tensor, delta: torch.Tensor):
    a = 1.0 - torch.exp(-sig * delta)
    T = torch.cumprod(torch.cat([torch.ones(a.shape[0], 1, device=a.device), 1.0 - a + 1e-10], -1), -1)
    return a, a * T[:, :-1]


# --------------------------------------------------------------------------------------------------------------
# forward (FieldBase.py:251-312); jitter = the per-ray u of is_train=True, passed in (forward only)
# --------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def render_chunk(spec: FieldSpec, rays: torch.Tensor, white_bg: bool = True, N_samples: int = -1, jitter=None):
    """rays [R,>=6] fp32 CPU -> rgb [R,3], depth [R].  Accumulates n_valid / n_active into spec.stats."""
    return _render_chunk(spec, rays, white_bg, N_samples, jitter)


def _render_chunk(spec: FieldSpec, rays: torch.Tensor, white_bg: bool = True, N_samples: int = -1, jitter=None):
    """Body of render_chunk; differentiable when the spec's parameter tensors require gradients (loss_and_grads)."""
    S = N_samples if N_samples > 0 else spec.n_samples
    o, d = rays[:, :3], rays[:, 3:6]
    p, t, live = march(spec, o, d, S, jitter)
    delta = torch.cat([t[:, 1:] - t[:, :-1], torch.zeros_like(t[:, :1])], -1)
    if spec.alpha_volume is not None:
        keep = alpha_keep(spec, p[live])
        live = live.clone()
        live[live.clone()] = keep
    R = rays.shape[0]
    dev = rays.device
    sig = torch.zeros(R, S, device=dev)
    cxy, cyz, cxz = torch.zeros(R, S, 2, device=dev), torch.zeros(R, S, 2, device=dev), torch.zeros(R, S, 2, device=dev)
    if live.any():
        n = to_unit_cube(spec, p)
        a, b, c = gauge_coords(spec, n[live])
        sig[live] = sigma(spec, a, b, c)
        cxy[live], cyz[live], cxz[live] = a, b, c
    _, w = transmittance(sig, delta * spec.distance_scale)
    hot = w > spec.weight_thres
    rgb = torch.zeros(R, S, 3, device=dev)
    if hot.any():
        dirs = d[:, None, :].expand(R, S, 3)
        rgb[hot] = colour(spec, cxy[hot], cyz[hot], cxz[hot], dirs[hot])
    acc = w.sum(-1)
    out = (w[..., None] * rgb).sum(-2)
    if white_bg:
        out = out + (1.0 - acc[:, None])
    out = out.clamp(0, 1)
    with torch.no_grad():                                                # FieldBase.py:304-306
        depth = (w * t).sum(-1) + (1.0 - acc) * rays[:, -1]              # NB last ray column (FieldBase.py:306)
    st = spec.stats
    st["rays"] = st.get("rays", 0) + R
    st["n_valid"] = st.get("n_valid", 0) + int(live.sum())
    st["n_active"] = st.get("n_active", 0) + int(hot.sum())
    return out, depth


@torch.no_grad()
def render(spec: FieldSpec, rays: torch.Tensor, chunk: int = 4096, white_bg: bool = True, N_samples: int = -1,
           jitter=None):
    """Chunk loop of TriPlane/main.py:60-71."""
    rgbs, depths = [], []
    for s in range(0, rays.shape[0], chunk):
        r, z = render_chunk(spec, rays[s:s + chunk], white_bg=white_bg, N_samples=N_samples,
                            jitter=None if jitter is None else jitter.reshape(-1, 1)[s:s + chunk])
        rgbs.append(r)
        depths.append(z)
    return torch.cat(rgbs), torch.cat(depths)


# --------------------------------------------------------------------------------------------------------------
# §8f rank 3  one training step's loss and gradients (TriPlane/main.py:272-283), the oracle of the backward pass that
# the CUDA path does not have yet: forward(is_train=True) with the given jitter, loss = mean((rgb_map - rgb_train)^2),
# autograd through the same restated ops.  Parameter names are the reference's state_dict keys.
# --------------------------------------------------------------------------------------------------------------
def param_tensors(spec: FieldSpec) -> dict:
    p = {"plane_xy": spec.planes[0], "plane_yz": spec.planes[1], "plane_xz": spec.planes[2],
         "rgb_decoder.basis.weight": spec.basis_w}
    if spec.gauge is not None:
        p.update({"gauge_xy": spec.gauge[0], "gauge_yz": spec.gauge[1], "gauge_xz": spec.gauge[2]})
    for i, (w, b) in zip((0, 2, 4), spec.rgb_layers):
        p[f"rgb_decoder.mlp.{i}.weight"], p[f"rgb_decoder.mlp.{i}.bias"] = w, b
    if spec.variant == "infoinv":
        for i, (w, b) in zip((0, 2, 4), spec.density_layers):
            p[f"density_decoder.mlp.{i}.weight"], p[f"density_decoder.mlp.{i}.bias"] = w, b
    else:
        p["density_decoder.weight"], p["density_decoder.bias"] = spec.density_layers[0]
    return p


def loss_and_grads(spec: FieldSpec, rays: torch.Tensor, target: torch.Tensor, jitter, white_bg: bool = True,
                   N_samples: int = -1):
    """-> (loss 0-d, {state_dict key: gradient}) for one chunk of training rays.  ``spec`` is not modified."""
    import dataclasses
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in param_tensors(spec).items()}
    s2 = dataclasses.replace(
        spec, planes=[leaf["plane_xy"], leaf["plane_yz"], leaf["plane_xz"]], basis_w=leaf["rgb_decoder.basis.weight"],
        gauge=None if spec.gauge is None else [leaf["gauge_xy"], leaf["gauge_yz"], leaf["gauge_xz"]],
        rgb_layers=[(leaf[f"rgb_decoder.mlp.{i}.weight"], leaf[f"rgb_decoder.mlp.{i}.bias"]) for i in (0, 2, 4)],
        density_layers=([(leaf[f"density_decoder.mlp.{i}.weight"], leaf[f"density_decoder.mlp.{i}.bias"]) for i in (0, 2, 4)]
                        if spec.variant == "infoinv" else [(leaf["density_decoder.weight"], leaf["density_decoder.bias"])]),
        stats={})
    with torch.enable_grad():
        rgb, _ = _render_chunk(s2, rays, white_bg, N_samples, jitter)
        loss = torch.mean((rgb - target) ** 2)
        loss.backward()
    return loss.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}


# --------------------------------------------------------------------------------------------------------------
# point-wise queries used by occupancy maintenance (FieldBase.py:140-159 compute_alpha)
# --------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def sigma_at(spec: FieldSpec, pts: torch.Tensor, use_gauge: bool = False) -> torch.Tensor:
    """World points [N,3] -> sigma [N] (0 where the alpha mask rejects).  compute_alpha evaluates the field
    with the gauge OFF (iteration=-1, FieldBase.py:154)."""
    keep = alpha_keep(spec, pts) if spec.alpha_volume is not None else torch.ones(pts.shape[0], dtype=torch.bool)
    out = torch.zeros(pts.shape[0])
    if keep.any():
        saved = spec.gauge_on
        spec.gauge_on = saved and use_gauge
        try:
            a, b, c = gauge_coords(spec, to_unit_cube(spec, pts[keep]))
            out[keep] = sigma(spec, a, b, c)
        finally:
            spec.gauge_on = saved
    return out
