"""TEST INFRASTRUCTURE — CPU (PyTorch fp32) restatement of the reference's UV-Mapping (NeuTex) render path.

Never imported by the product package.  Follows (paths relative to /root/reference/UV-Mapping):

  encode           util.py:427-438                 positional_encoding; inputs are cat([x, PE(x)]) at every call site
  raygen           model/renderer.py:79-141        cube_ray_generation (jitter noise passed in instead of torch.rand)
  geometry         model/decoder.py:219-237        GeometryMlpDecoder.forward
  gauge            model/gauge_fields.py:37-46,60-74   GaugeNetwork.forward + GaugeTransform.forward ('square': tanh)
  texture          model/decoder.py:56-121         TextureMlpDecoder.forward incl. the target_texture branch, mode 0
  sample_square    util.py:277-282                 bilinear, align_corners=False, border padding
  march            model/renderer.py:176-247,4-11  ray_march with radiance_render / alpha_blend, simple_tone_map
  render           model/model.py:27-59            NeuTex.forward minus the loss-only inverse-gauge lines (35-36, 56),
                                                   which cannot run as shipped

The arithmetic that is not under /root/reference is PyTorch's (nn.Linear, softplus, tanh, leaky_relu, cumsum, cumprod,
grid_sample); the same torch ops are called in the same order, so on the same torch build the restatement is
bit-identical to the reference modules (tests/test_oracle_neutex.py, tests/golden/neutex_*.npz).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F


@dataclass
class NeuTexSpec:
    state: Dict[str, torch.Tensor]              # reference parameter names (NeuTex.state_dict())
    sample_num: int = 64
    jitter: float = 0.05                        # hard-coded at model.py:30
    texture: Optional[torch.Tensor] = None      # [h, w, c] edited texture (TextureMlpDecoder.cubemap_) or None


def encode(x: torch.Tensor, freqs: int) -> torch.Tensor:
    bands = (2 ** torch.arange(freqs).float())
    pts = (x[..., None] * bands).reshape(x.shape[:-1] + (freqs * x.shape[-1],))
    return torch.cat([x, torch.sin(pts), torch.cos(pts)], dim=-1)


def raygen(campos, raydir, S: int, noise: torch.Tensor, jitter: float = 0.05, domain: float = 1.0):
    """-> raypos [N,R,S,3], seg [N,R,S], valid [N,R,S] (uint8), mid_ts [N,R,S]."""
    t1 = (-domain - campos[:, None, :]) / raydir
    t2 = (domain - campos[:, None, :]) / raydir
    tmin = torch.max(torch.min(t1[..., 0], t2[..., 0]),
                     torch.max(torch.min(t1[..., 1], t2[..., 1]), torch.min(t1[..., 2], t2[..., 2])))
    tmax = torch.min(torch.max(t1[..., 0], t2[..., 0]),
                     torch.min(torch.max(t1[..., 1], t2[..., 1]), torch.max(t1[..., 2], t2[..., 2])))
    hit = tmin < tmax
    t = torch.where(hit, tmin, torch.zeros_like(tmin)).clamp(min=0.0)
    dt = domain * 2 / S
    seg = dt + dt * jitter * (noise - 0.5)
    ends = torch.cumsum(seg, dim=2)
    ends = torch.cat([torch.zeros((ends.shape[0], ends.shape[1], 1)), ends], dim=2)
    ends = t[:, :, None] + ends
    mid = (ends[:, :, :-1] + ends[:, :, 1:]) / 2
    pos = campos[:, None, None, :] + raydir[:, :, None, :] * mid[:, :, :, None]
    valid = torch.prod(torch.gt(pos, -domain) * torch.lt(pos, domain), dim=-1).byte()
    return pos, seg, valid, mid


def _seq(st, prefix, n_layers, x, act, last_act=True):
    for n in range(n_layers):
        x = F.linear(x, st[f"{prefix}.{2 * n}.weight"], st[f"{prefix}.{2 * n}.bias"])
        if n < n_layers - 1 or last_act:
            x = act(x)
    return x


def geometry(spec: NeuTexSpec, pts: torch.Tensor) -> torch.Tensor:
    """-> density [N,R,S] = softplus(block([p, PE(p,10)]))."""
    raw = _seq(spec.state, "net_geometry_decoder.block", 12, encode(pts, 10), torch.relu, last_act=False)[..., 0]
    return F.softplus(raw)


def gauge(spec: NeuTexSpec, pts: torch.Tensor) -> torch.Tensor:
    """-> uv [N,R,S,2] = tanh(GaugeNetwork([p, PE(p,10)])) for the square primitive, [N,R,S,3] = normalize(...) for the
    sphere (gauge_fields.py:60-74); the primitive is read off the last layer's width."""
    st, e = spec.state, "gauge_transform.encoder"
    shape = pts.shape
    x = pts.reshape(shape[0], -1, 3)
    x = torch.relu(F.linear(encode(x, 10), st[f"{e}.linear1.weight"], st[f"{e}.linear1.bias"]))
    x = torch.relu(F.linear(x, st[f"{e}.linear2.weight"], st[f"{e}.linear2.bias"]))
    for i in range(2):
        x = torch.relu(F.linear(x, st[f"{e}.linear_list.{i}.weight"], st[f"{e}.linear_list.{i}.bias"]))
    x = F.linear(x, st[f"{e}.last_linear.weight"], st[f"{e}.last_linear.bias"])
    out_dim = st[f"{e}.last_linear.weight"].shape[0]
    x = x.reshape(shape[:-1] + (out_dim,))
    return torch.tanh(x) if out_dim == 2 else F.normalize(x, dim=-1)


def sample_square(square: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    return F.grid_sample(square.permute(2, 0, 1)[None], uv.reshape((1, -1, 1, 2)), padding_mode="border",
                         align_corners=False).permute(0, 2, 3, 1).reshape(uv.shape[:-1] + (square.shape[-1],))


def texture(spec: NeuTexSpec, uv: torch.Tensor, view_dir: torch.Tensor) -> torch.Tensor:
    """uv [N,R,S,2], view_dir [N,R,1,3] -> [N,R,S,3 or texture channels]."""
    st = spec.state
    leaky = lambda v: F.leaky_relu(v, 0.2)
    out = _seq(st, "net_texture.block1", 6, encode(uv, 10), leaky)
    c1 = F.softplus(F.linear(out, st["net_texture.color1.weight"], st["net_texture.color1.bias"]))
    vd = view_dir.expand(out.shape[:-1] + (3,))
    h = torch.cat([out, encode(vd, 6)], dim=-1)
    c2 = _seq(st, "net_texture.block2", 5, h, leaky, last_act=False)
    if spec.texture is None:
        return (c1 + c2).clamp(min=0)
    orig = ((c1 + c2) * 8).clamp(min=0, max=1)
    return sample_square(spec.texture, uv) * orig.mean(dim=-1, keepdim=True)


def march(density, radiance, seg, valid):
    """-> ray_color [N,R,3], background transmittance [N,R], blend weights [N,R,S]."""
    sigma = density * valid.float()
    opacity = 1 - torch.exp(-sigma * seg)
    acc = torch.cumprod(1.0 - opacity + 1e-10, dim=-1)
    bg_t = acc[:, :, -1]
    acc = torch.cat([torch.ones(opacity.shape[0:2] + (1,)), acc[:, :, :-1]], dim=-1)
    w = opacity * acc
    return torch.sum(radiance * w[..., None], dim=-2), bg_t, w


def tone_map(color, gamma=2.2, exposure=1):
    return torch.pow(color * exposure + 1e-5, 1 / gamma).clamp(0, 1)


@torch.no_grad()
def render(spec: NeuTexSpec, campos, raydir, background, noise, chunk: int = 1024):
    """campos [1,3], raydir [1,R,3], background [1,3] or None, noise [1,R,S] -> color [1,R,3], transmittance [1,R].
    Chunked over rays like UV-Mapping/test.py:108-114 (random_sample_size^2 = 1024 rays per call)."""
    cols, trs = [], []
    for s in range(0, raydir.shape[1], chunk):
        rd, nz = raydir[:, s:s + chunk], noise[:, s:s + chunk]
        pos, seg, valid, _ = raygen(campos, rd, spec.sample_num, nz, spec.jitter)
        dens = geometry(spec, pos)
        uv = gauge(spec, pos)
        rad = texture(spec, uv, rd[:, :, None, :])[..., :3]
        color, bg_t, _ = march(dens, rad, seg, valid)
        if background is not None:
            color = color + background[:, None, :] * bg_t[:, :, None]
        cols.append(tone_map(color))
        trs.append(bg_t)
    return torch.cat(cols, 1), torch.cat(trs, 1)
