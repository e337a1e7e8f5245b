"""Drop-in for the render path of the reference's ``UV-Mapping/model/model.py`` ``NeuTex`` module
(``primitive_type`` 'square' or 'sphere'), executing on hand-written sm_100a CUDA through ``libngf_b200.so``.

Sub-modules keep the reference's names and parameter shapes so ``state_dict`` round-trips with ``strict=False``
(the reference's own ``load_networks`` uses ``strict=False``, model/model.py:215-230):
``net_geometry_decoder.block.{0..22}``, ``gauge_transform.encoder.{linear1,linear2,linear_list.{0,1},last_linear}``,
``net_texture.{block1.{0..10},color1,block2.{0..8}}``.  The loss-only inverse gauge network
(model.py:35-36,56; broken as shipped) is not part of the render path and is not built.

There is no CPU implementation here: a module living on the CPU raises.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib


def _seq(dims, act):
    layers = []
    for n, (i, o) in enumerate(dims):
        layers.append(nn.Linear(i, o))
        if act is not None and n < len(dims) - 1:
            layers.append(act())
    return nn.Sequential(*layers)


class GeometryMlpDecoder(nn.Module):
    """Parameters of decoder.py:201-217 (63 -> 256 -> 10 x 256 -> 1)."""

    def __init__(self, pos_freqs=10, hidden_size=256, num_layers=10):
        super().__init__()
        dims = [(3 + 6 * pos_freqs, hidden_size)] + [(hidden_size, hidden_size)] * num_layers + [(hidden_size, 1)]
        self.block = _seq(dims, nn.ReLU)


class GaugeNetwork(nn.Module):
    """Parameters of gauge_fields.py:8-35 (63 -> 64 -> 128 -> 128 -> 128 -> 2 | 3)."""

    def __init__(self, input_dim=3, output_dim=2, mid_size=64, hidden_size=128, num_layers=2):
        super().__init__()
        self.linear1 = nn.Linear(input_dim + 2 * input_dim * 10, mid_size)
        self.linear2 = nn.Linear(mid_size, hidden_size)
        self.linear_list = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(num_layers)])
        self.last_linear = nn.Linear(hidden_size, output_dim)


class GaugeTransform(nn.Module):
    """gauge_fields.py:49-74: 'square' -> 2 outputs, uv = tanh; 'sphere' -> 3 outputs, uv = normalize."""

    def __init__(self, primitive_type="square"):
        super().__init__()
        if primitive_type not in ("square", "sphere"):
            raise Exception("Unknown primitive type {}".format(primitive_type))          # as gauge_fields.py:57-58
        self.primitive_type = primitive_type
        self.output_dim = 2 if primitive_type == "square" else 3
        self.encoder = GaugeNetwork(3, self.output_dim)


class TextureMlpDecoder(nn.Module):
    """Parameters of decoder.py:11-36 (42 -> 256 -> 5 x 256; color1 256 -> 3; 295 -> 256 -> 3 x 256 -> 3)."""

    def __init__(self, width=256, layers=(5, 3), uv_dim=2):
        super().__init__()
        # Linear at even Sequential indices, as in the reference; input = [uv, PE(uv, 10)] (uv_dim * 21)
        mods = []
        for i, o in [(uv_dim * 21, width)] + [(width, width)] * layers[0]:
            mods += [nn.Linear(i, o), nn.LeakyReLU(0.2)]
        self.block1 = nn.Sequential(*mods)
        self.color1 = nn.Linear(width, 3)
        mods = []
        for i, o in [(width + 3 + 36, width)] + [(width, width)] * layers[1]:
            mods += [nn.Linear(i, o), nn.LeakyReLU(0.2)]
        mods.append(nn.Linear(width, 3))
        self.block2 = nn.Sequential(*mods)
        self.cubemap_ = None          # [h, w, c] edited texture (util.load_square) or None
        self.cubemap_mode_ = 0


class NeuTex(nn.Module):
    """Reference: UV-Mapping/model/model.py:11-59."""

    def __init__(self, opt=None, device="cuda"):
        super().__init__()
        self.opt = opt if opt is not None else SimpleNamespace(sample_num=64, primitive_type="square", target_texture="None")
        prim = getattr(self.opt, "primitive_type", "square")
        self.net_geometry_decoder = GeometryMlpDecoder(pos_freqs=10, hidden_size=256, num_layers=10)
        self.gauge_transform = GaugeTransform(prim)
        self.net_texture = TextureMlpDecoder(uv_dim=2 if prim == "square" else 3)          # model.py:22
        self._handle = None
        self._handle_sig = None
        self.to(device)

    # ------------------------------------------------------------------ handle management
    @property
    def device(self):
        return next(self.parameters()).device

    def set_texture(self, texture):
        """``texture``: [h, w, c] float tensor in [0,1] as ``util.load_square`` returns it (vertically flipped image),
        or None to restore the learned texture (decoder.py:48-52, 79-103 mode 0)."""
        self.net_texture.cubemap_ = None if texture is None else texture.detach().float().contiguous()
        self._handle_sig = None

    def _layers(self):
        g = [self.net_geometry_decoder.block[i] for i in range(0, 23, 2)]
        e = self.gauge_transform.encoder
        ga = [e.linear1, e.linear2, e.linear_list[0], e.linear_list[1], e.last_linear]
        t1 = [self.net_texture.block1[i] for i in range(0, 11, 2)]
        t2 = [self.net_texture.block2[i] for i in range(0, 9, 2)]
        return g, ga, t1, self.net_texture.color1, t2

    def _signature(self):
        tex = self.net_texture.cubemap_
        return [(p.data_ptr(), p._version) for p in self.parameters()] + [None if tex is None else (tex.data_ptr(), tex._version)]

    def _free_handle(self):
        if getattr(self, "_handle", None):
            _lib.load().ngf_neutex_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free_handle()
        except Exception:
            pass

    def _ensure_handle(self):
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("ngf_b200 NeuTex runs on a CUDA device only (there is no CPU fallback); "
                               f"this module is on {dev}")
        sig = self._signature()
        if self._handle is not None and sig == self._handle_sig:
            return self._handle
        lib = _lib.load()
        d = _lib.NgfNeutexDesc()
        keep = []

        def lin(m):
            w, b = m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous()
            keep.extend([w, b])
            return _lib.NgfLinear(w.data_ptr(), b.data_ptr(), m.in_features, m.out_features)

        g, ga, t1, c1, t2 = self._layers()
        for i, m in enumerate(g):
            d.geometry[i] = lin(m)
        for i, m in enumerate(ga):
            d.gauge[i] = lin(m)
        for i, m in enumerate(t1):
            d.tex_block1[i] = lin(m)
        d.tex_color1 = lin(c1)
        for i, m in enumerate(t2):
            d.tex_block2[i] = lin(m)
        d.sample_num = int(getattr(self.opt, "sample_num", 64))
        d.jitter = 0.05                                                   # model.py:30
        d.primitive = 0 if self.gauge_transform.primitive_type == "square" else 1
        tex = self.net_texture.cubemap_
        if tex is not None:
            t = tex.to(dev).float().contiguous()
            keep.append(t)
            d.texture = t.data_ptr()
            d.tex_h, d.tex_w, d.tex_c = t.shape
        torch.cuda.synchronize(dev)
        self._free_handle()
        h = C.c_void_p()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _lib.check(lib.ngf_neutex_pack(C.byref(d), idx, C.byref(h)), "ngf_neutex_pack")
        self._handle, self._handle_sig = h, sig
        if getattr(self, "_precision", None) is not None:             # set_precision() survives a re-pack
            _lib.check(lib.ngf_neutex_set_precision(h, 1 if self._precision == "fp32" else 0), "ngf_neutex_set_precision")
        return h

    # ------------------------------------------------------------------ render
    @torch.no_grad()
    def forward(self, camera_position=None, ray_direction=None, background_color=None, noise=None, seed=None):
        """camera_position [N,3], ray_direction [N,R,3], background_color [N,3] or None ->
        {'color': [N,R,3], 'transmittance': [N,R]} (model.py:27-59).  The jitter the reference draws with ``torch.rand``
        inside ``cube_ray_generation``: ``noise`` [N,R,sample_num] U[0,1) numbers when given; otherwise drawn inside the
        kernels from ``seed`` (camera n uses seed + n; ``noise_for`` returns the same numbers), a fresh seed from torch's
        generator when that is omitted too."""
        h = self._ensure_handle()
        lib, dev = _lib.load(), self.device
        cam = camera_position.to(dev).float().contiguous()
        rd = ray_direction.to(dev).float().contiguous()
        N, R = rd.shape[0], rd.shape[1]
        if noise is None and seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        nz = None if noise is None else noise.to(dev).float().contiguous()
        bg = None if background_color is None else background_color.to(dev).float().contiguous()
        color = torch.empty((N, R, 3), dtype=torch.float32, device=dev)
        trans = torch.empty((N, R), dtype=torch.float32, device=dev)
        stream = int(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            for n in range(N):
                bgp = None if bg is None else bg[n].data_ptr()
                if nz is not None:
                    _lib.check(lib.ngf_neutex_render(h, cam[n].data_ptr(), rd[n].data_ptr(), bgp, nz[n].data_ptr(), R,
                                                     color[n].data_ptr(), trans[n].data_ptr(), stream), "ngf_neutex_render")
                else:
                    _lib.check(lib.ngf_neutex_render_seeded(h, cam[n].data_ptr(), rd[n].data_ptr(), bgp, int(seed) + n, 0, R,
                                                            color[n].data_ptr(), trans[n].data_ptr(), stream),
                               "ngf_neutex_render_seeded")
        return {"color": color, "transmittance": trans}

    def noise_for(self, seed: int, n_rays: int, first_ray: int = 0) -> torch.Tensor:
        """[1, n_rays, sample_num] jitter numbers a render with ``seed`` draws for frame rays first_ray.. (ngf_neutex_noise)."""
        h = self._ensure_handle()
        S = int(getattr(self.opt, "sample_num", 64))
        out = torch.empty((1, n_rays, S), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_neutex_noise(h, int(seed), first_ray, n_rays, out.data_ptr(),
                                                    int(torch.cuda.current_stream(self.device).cuda_stream)), "ngf_neutex_noise")
        return out

    @torch.no_grad()
    def render_host(self, camera_position, ray_direction, background_color, noise=None, color_host=None, trans_host=None,
                    seed=None):
        """One camera through HOST buffers (ngf_neutex_render_host[_seeded]): [1,3], [1,R,3], [1,3] or None CPU tensors
        (pinned for full copy bandwidth), jitter as [1,R,sample_num] ``noise`` or drawn on the device from ``seed`` ->
        (color [1,R,3], transmittance [1,R]) CPU tensors."""
        h = self._ensure_handle()
        R = ray_direction.shape[1]
        if (noise is None) == (seed is None):
            raise ValueError("render_host takes either noise or seed")
        for t in (camera_position, ray_direction) + (() if noise is None else (noise,)):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("render_host takes contiguous fp32 CPU tensors")
        if color_host is None:
            color_host = torch.empty((1, R, 3)).pin_memory()
        if trans_host is None:
            trans_host = torch.empty((1, R)).pin_memory()
        bgp = None if background_color is None else background_color.data_ptr()
        if noise is not None:
            _lib.check(_lib.load().ngf_neutex_render_host(h, camera_position.data_ptr(), ray_direction.data_ptr(), bgp,
                                                          noise.data_ptr(), R, color_host.data_ptr(), trans_host.data_ptr()),
                       "ngf_neutex_render_host")
        else:
            _lib.check(_lib.load().ngf_neutex_render_host_seeded(h, camera_position.data_ptr(), ray_direction.data_ptr(), bgp,
                                                                 int(seed), R, color_host.data_ptr(), trans_host.data_ptr()),
                       "ngf_neutex_render_host_seeded")
        return color_host, trans_host

    def set_precision(self, mode: str):
        """"tc" (default): tcgen05 fp16 / split-fp16 tensor-core arithmetic; "fp32": the CUDA-core fp32 kernel over the
        unpacked parameters (ngf_neutex_set_precision) — slower, for checkpoints the self-check flags."""
        if mode not in ("tc", "fp32"):
            raise ValueError('precision is "tc" or "fp32"')
        self._precision = mode
        _lib.check(_lib.load().ngf_neutex_set_precision(self._ensure_handle(), 1 if mode == "fp32" else 0),
                   "ngf_neutex_set_precision")

    def self_check(self, n_points: int = 4096, seed: int = 0) -> dict:
        """Deviation of the tensor-core path from the fp32 path on seeded random in-cube points (ngf_neutex_self_check):
        {"sigma_rel": max |d sigma| / (1 + |sigma|), "rgb_max": max |d rgb|, "rgb_mean": ..., "rgb_range": max |rgb|}."""
        rep = (C.c_float * 4)()
        _lib.check(_lib.load().ngf_neutex_self_check(self._ensure_handle(), n_points, seed, rep,
                                                     int(torch.cuda.current_stream(self.device).cuda_stream)),
                   "ngf_neutex_self_check")
        return {"sigma_rel": rep[0], "rgb_max": rep[1], "rgb_mean": rep[2], "rgb_range": rep[3]}

    def last_valid_samples(self) -> int:
        n = C.c_uint64()
        _lib.check(_lib.load().ngf_neutex_last_valid_samples(self._ensure_handle(), C.byref(n),
                                                             int(torch.cuda.current_stream(self.device).cuda_stream)))
        return int(n.value)

    def last_samples(self, n_rays: int):
        """-> (sigma_rgb [n_rays,64,4], valid [n_rays,64] bool) of the first ``n_rays`` rays of the last render."""
        import numpy as np
        out = torch.empty((n_rays, 64, 4), dtype=torch.float32)
        mask = np.zeros(n_rays, dtype=np.uint64)
        _lib.check(_lib.load().ngf_neutex_copy_samples(self._ensure_handle(), 0, n_rays * 64, out.data_ptr(),
                                                       mask.ctypes.data, 0, n_rays))
        valid = torch.from_numpy(((mask[:, None] >> np.arange(64, dtype=np.uint64)[None]) & np.uint64(1)).astype(np.bool_))
        return out, valid

    def kernel_timing(self, capacity: int):
        _lib.check(_lib.load().ngf_neutex_timing_begin(self._ensure_handle(), int(capacity)))

    def kernel_timing_read(self):
        """-> (renders timed, raygen ms, mlp ms, march ms)."""
        n, a, b, c = C.c_int32(), C.c_double(), C.c_double(), C.c_double()
        _lib.check(_lib.load().ngf_neutex_timing_read(self._ensure_handle(), C.byref(n), C.byref(a), C.byref(b), C.byref(c)))
        return int(n.value), float(a.value), float(b.value), float(c.value)
