"""Host-side mirror of the reference's ``Base`` field class (``TriPlane/models/FieldBase.py:44-312`` and
``InfoInv/models/FieldBase.py``): same constructor, attributes, parameter names and method signatures, with the
render path executed by ``libngf_b200.so`` (hand-written sm_100a CUDA) through the C ABI in
``include/ngf_b200.h``.

There is deliberately no CPU implementation here: a field living on the CPU, a missing shared library or a
non-Blackwell GPU raises.  ``is_train=True`` runs the reference's jittered sampling (FieldBase.py:128-130); with autograd
enabled the call is recorded as one ``torch.autograd.Function`` whose backward is ``ngf_field_backward`` (csrc/ngf_train.cu),
so the reference's training step (TriPlane/main.py:272-302: forward, MSE, ``loss.backward()``, Adam) runs on these classes.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib


def _cuda_stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class AlphaGridMask(torch.nn.Module):
    """Binary occupancy volume with its own box (reference: FieldBase.py:22-40).  ``sample_alpha`` answers through
    the owning field's packed bit grid (ngf_field_alpha_value): the trilinear value of the {0,1} volume, as the
    reference's ``F.grid_sample`` returns it."""

    def __init__(self, device, aabb, alpha_volume):
        super().__init__()
        self.device = device
        self.aabb = aabb.to(self.device)
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.invgridSize = 1.0 / self.aabbSize * 2
        self.alpha_volume = alpha_volume.view(1, 1, *alpha_volume.shape[-3:])
        self.gridSize = torch.LongTensor(
            [alpha_volume.shape[-1], alpha_volume.shape[-2], alpha_volume.shape[-3]]).to(self.device)
        self._owner = None      # set by Base when the mask is attached

    def normalize_coord(self, xyz_sampled):
        return (xyz_sampled - self.aabb[0]) * self.invgridSize - 1

    def sample_alpha(self, xyz_sampled):
        owner = self._owner() if self._owner is not None else None
        if owner is None or owner.alphaMask is not self:
            raise RuntimeError("this AlphaGridMask is not the mask of a live field (assign it to field.alphaMask first); "
                               "a detached or replaced mask would answer from another mask's packed bits")
        return owner._alpha_value(xyz_sampled)


class _RenderTrain(torch.autograd.Function):
    """forward(is_train=True) as one autograd node: ngf_field_render_jitter forward, ngf_field_backward backward."""

    @staticmethod
    def forward(ctx, field, rays, jitter, white_bg, N_samples, image_width, *params):
        with torch.no_grad():
            out = field._forward(rays, white_bg, True, N_samples, image_width, jitter=jitter, _white_decided=True,
                                 **field._train_fwd_kw)
        ctx.field, ctx.white_bg, ctx.N_samples = field, white_bg, N_samples
        ctx.save_for_backward(rays, jitter)
        ctx.handle_sig = field._handle_sig
        ctx.mark_non_differentiable(out['depth_map'])                     # FieldBase.py:304-306: under torch.no_grad()
        return out['rgb_map'], out['depth_map']

    @staticmethod
    def backward(ctx, g_rgb, _g_depth):
        field = ctx.field
        rays, jitter = ctx.saved_tensors
        if field._signature() != ctx.handle_sig:
            raise RuntimeError("the field's parameters changed between forward(is_train=True) and backward(): the "
                               "backward re-marches the rays on the packed parameters of the forward")
        grads = field._backward(rays, jitter, ctx.white_bg, ctx.N_samples, g_rgb)
        params = field._grad_parameters()
        out = [g for g, p in zip(grads, params) if p is not None]
        return (None, None, None, None, None, None, *out)


class Base(torch.nn.Module):
    VARIANT = _lib.NGF_TRIPLANE

    def __init__(self, aabb, gridSize, device, alphaMask=None, near_far=[2.0, 6.0], alphaMask_thres=0.001,
                 distance_scale=25, rayMarch_weight_thres=0.0001, step_ratio=2.0, **model_kw):
        super().__init__()
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.aabb = aabb.to(self.device) if isinstance(aabb, torch.Tensor) else torch.tensor(aabb, device=self.device)
        self._alphaMask = None
        self._handle = None
        self._handle_sig = None
        self._mlp_impl = _lib.MLP_TCGEN05
        self.alphaMask = alphaMask
        self.alphaMask_thres = alphaMask_thres
        self.distance_scale = distance_scale
        self.rayMarch_weight_thres = rayMarch_weight_thres
        self.near_far = near_far
        self.step_ratio = step_ratio
        self.init_para(gridSize)
        self.init_model(device=self.device, **model_kw)

    # ------------------------------------------------------------------ bookkeeping (FieldBase.py:63-74)
    def init_para(self, gridSize):
        # Evaluated on the CPU whatever the field's device: a CUDA torch.mean can round differently from the CPU
        # one by an ulp, and stepSize feeds every sample position (one ulp there moves samples across mask
        # boundaries).  The reference's CPU path is the parity target.
        aabb = self.aabb.detach().float().cpu()
        size = aabb[1] - aabb[0]
        grid = torch.LongTensor([int(g) for g in gridSize])
        units = size / (grid - 1)
        step = torch.mean(units) * self.step_ratio
        diag = torch.sqrt(torch.sum(torch.square(size)))
        self.aabbSize = size.to(self.device)
        self.invaabbSize = (2.0 / size).to(self.device)
        self.gridSize = grid.to(self.device)
        self.units = units.to(self.device)
        self.stepSize = step.to(self.device)
        self.aabbDiag = diag.to(self.device)
        self.nSamples = int((diag / step).item()) + 1
        self._invalidate()

    def init_model(self, **kw):
        raise NotImplementedError

    @property
    def alphaMask(self):
        return self._alphaMask

    def __setattr__(self, name, value):
        # nn.Module.__setattr__ would register an AlphaGridMask as a sub-module and bypass a property setter;
        # intercept it so the mask is attached to this field and the packed handle is refreshed.
        if name == "alphaMask":
            old = self.__dict__.get("_alphaMask")
            if old is not None and old is not value:
                old._owner = None                       # a replaced mask must not answer from this field's packed bits
            object.__setattr__(self, "_alphaMask", value)
            if value is not None:
                value._owner = weakref.ref(self)
            self._invalidate()
            return
        super().__setattr__(name, value)

    def normalize_coord(self, xyz_sampled):
        return (xyz_sampled - self.aabb[0]) * self.invaabbSize - 1

    def get_optparam_groups(self, lr_init_spatial=0.02, lr_init_network=0.001):
        raise NotImplementedError

    # ------------------------------------------------------------------ checkpoint format (FieldBase.py:94-116)
    def get_kwargs(self):
        return {'aabb': self.aabb, 'gridSize': self.gridSize.tolist(), 'alphaMask_thres': self.alphaMask_thres,
                'distance_scale': self.distance_scale, 'rayMarch_weight_thres': self.rayMarch_weight_thres,
                'near_far': self.near_far, 'step_ratio': self.step_ratio}

    def save(self, path):
        ckpt = {'kwargs': self.get_kwargs(), 'state_dict': self.state_dict()}
        if self.alphaMask is not None:
            alpha_volume = self.alphaMask.alpha_volume.bool().cpu().numpy()
            ckpt.update({'alphaMask.shape': alpha_volume.shape})
            ckpt.update({'alphaMask.mask': np.packbits(alpha_volume.reshape(-1))})
            ckpt.update({'alphaMask.aabb': self.alphaMask.aabb.cpu()})
        torch.save(ckpt, path)

    def load(self, ckpt):
        if 'alphaMask.aabb' in ckpt.keys():
            length = int(np.prod(ckpt['alphaMask.shape']))
            alpha_volume = torch.from_numpy(
                np.unpackbits(ckpt['alphaMask.mask'])[:length].reshape(ckpt['alphaMask.shape']))
            self.alphaMask = AlphaGridMask(self.device, ckpt['alphaMask.aabb'].to(self.device),
                                           alpha_volume.float().to(self.device))
        self.load_state_dict(ckpt['state_dict'])
        self._invalidate()

    # ------------------------------------------------------------------ C-ABI handle management
    def _invalidate(self):
        self._handle_sig = None

    def _free_handle(self):
        if getattr(self, "_handle", None):
            _lib.load().ngf_field_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free_handle()
        except Exception:
            pass

    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("ngf_b200 fields run on a CUDA device only (there is no CPU fallback); "
                               f"this field was constructed with device={self.device}")

    def _signature(self):
        """Cheap staleness key (no device reads): parameter storage + in-place version counters, the mask, and the
        python-side scalars.  aabb / stepSize / nSamples only change through init_para(), which invalidates."""
        sig = [(p.data_ptr(), p._version) for p in self.parameters()]
        am = self.alphaMask
        sig.append(None if am is None else (id(am), am.alpha_volume.data_ptr(), am.alpha_volume._version))
        sig.append((tuple(float(x) for x in self.near_far), float(self.distance_scale),
                    float(self.rayMarch_weight_thres)))
        return sig

    def _fill_common(self, d: _lib.NgfFieldDesc, keep: list):
        def lin(mod, need_bias=True):
            w = _f32c(mod.weight)
            keep.append(w)
            b = None
            if mod.bias is not None:
                b = _f32c(mod.bias)
                keep.append(b)
            return _lib.NgfLinear(w.data_ptr(), b.data_ptr() if b is not None else None, mod.in_features,
                                  mod.out_features)

        planes = [self.plane_xy, self.plane_yz, self.plane_xz]
        for i, p in enumerate(planes):
            t = _f32c(p)
            keep.append(t)
            d.plane[i] = t.data_ptr()
            d.plane_h[i], d.plane_w[i] = t.shape[2], t.shape[3]
        d.plane_c = planes[0].shape[1]
        d.rgb_basis = lin(self.rgb_decoder.basis)
        d.rgb_l1, d.rgb_l2, d.rgb_l3 = (lin(self.rgb_decoder.mlp[i]) for i in (0, 2, 4))
        d.view_pe = int(self.rgb_decoder.view_pe)
        d.density_shift = -10.0
        aabb = self.aabb.detach().float().cpu()
        inv = self.invaabbSize.detach().float().cpu()
        for k in range(3):
            d.aabb[k] = float(aabb[0, k])
            d.aabb[3 + k] = float(aabb[1, k])
            d.inv_aabb_size[k] = float(inv[k])
        d.step_size = float(self.stepSize.detach().float().cpu())
        d.n_samples = int(self.nSamples)
        d.near_t, d.far_t = float(self.near_far[0]), float(self.near_far[1])
        d.distance_scale = float(self.distance_scale)
        d.weight_thres = float(self.rayMarch_weight_thres)
        am = self.alphaMask
        if am is not None:
            vol = _f32c(am.alpha_volume)
            keep.append(vol)
            d.alpha_volume = vol.data_ptr()
            D_, H_, W_ = vol.shape[-3:]
            d.alpha_dims[0], d.alpha_dims[1], d.alpha_dims[2] = W_, H_, D_
            ab = am.aabb.detach().float().cpu()
            ai = am.invgridSize.detach().float().cpu()
            for k in range(3):
                d.alpha_aabb[k] = float(ab[0, k])
                d.alpha_aabb[3 + k] = float(ab[1, k])
                d.alpha_inv[k] = float(ai[k])
        else:
            d.alpha_volume = None

    def _fill_desc(self, d: _lib.NgfFieldDesc, keep: list):
        raise NotImplementedError

    def _shape_signature(self):
        am = self.alphaMask
        return ([tuple(p.shape) for p in self.parameters()],
                None if am is None else tuple(am.alpha_volume.shape))

    def invalidate(self):
        """Tell the field that parameters or attributes were changed in a way the version counters cannot see (e.g.
        ``plane.data.mul_()``, an in-place edit of ``aabb`` or of the mask volume): the next render re-packs."""
        self._invalidate()

    refresh = invalidate

    def _ensure_handle(self):
        self._require_cuda()
        sig = self._signature()
        if self._handle is not None and sig == self._handle_sig:
            return self._handle
        lib = _lib.load()
        d = _lib.NgfFieldDesc()
        keep: list = []
        self._fill_desc(d, keep)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        shapes = self._shape_signature()
        torch.cuda.synchronize(self.device)
        if self._handle is not None and shapes == getattr(self, "_handle_shapes", None):
            # same shapes (an optimizer step, load_state_dict): refresh the packed shadows in place — allocations,
            # streams and outstanding host-path tickets of the handle survive
            _lib.check(lib.ngf_field_repack(self._handle, C.byref(d)), "ngf_field_repack")
            self._handle_sig = sig
            return self._handle
        self._free_handle()
        h = C.c_void_p()
        _lib.check(lib.ngf_field_pack(C.byref(d), dev_index, C.byref(h)), "ngf_field_pack")
        self._handle = h
        self._handle_sig = sig
        self._handle_shapes = shapes
        return h

    def _live_handle(self):
        """The handle as it is (no staleness check): for waits, counters and timers, which must not re-pack."""
        return self._handle if self._handle is not None else self._ensure_handle()

    def set_mlp_impl(self, name: str):
        """'tcgen05' (default) or 'simt' (CUDA-core cross-check of the same packed weights)."""
        self._mlp_impl = {"tcgen05": _lib.MLP_TCGEN05, "simt": _lib.MLP_SIMT}[name]

    # ------------------------------------------------------------------ the render path
    def _set_switches(self, lib, h, **fwd_kw):
        pass

    def forward(self, rays_chunk, white_bg=True, is_train=False, N_samples=-1, image_width=0, **fwd_kw):
        """Reference: Base.forward (FieldBase.py:251-312).  Returns {'rgb_map': [R,3], 'depth_map': [R]} on the
        field's device.  ``image_width`` (optional, not in the reference) tells the kernel that the rays are the
        row-major pixels of an image so warps can take 8x4 pixel blocks.

        ``is_train=True`` (forward only): every ray's samples are shifted by one ``u ~ U[0,1)`` step, drawn exactly as
        the reference does (``torch.rand_like`` of a CPU ``[R,1]`` tensor, FieldBase.py:128-130) or passed in as
        ``jitter=`` ([R] or [R,1]); a non-white background is made white with probability 1/2 (FieldBase.py:299)."""
        if is_train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(rays_chunk, white_bg, N_samples, image_width, **fwd_kw)
        with torch.no_grad():
            return self._forward(rays_chunk, white_bg, is_train, N_samples, image_width, **fwd_kw)

    def _grad_parameters(self):
        """Parameters in the order of _lib.NgfFieldGrads: planes, gauge planes (or None), rgb_decoder, density head."""
        raise NotImplementedError

    def _forward_train(self, rays_chunk, white_bg, N_samples, image_width, jitter=None, **fwd_kw):
        """The training-time forward under autograd (TriPlane/main.py:272): one autograd node whose backward is
        ngf_field_backward."""
        h = self._ensure_handle()
        self._set_switches(_lib.load(), h, **fwd_kw)
        rays = _f32c(rays_chunk.to(self.device))
        R = rays.shape[0]
        if jitter is None:
            jitter = torch.rand_like(torch.empty((R, 1), dtype=torch.float32))          # CPU draw, as the reference
        jitter = _f32c(jitter.reshape(-1).to(self.device))
        white = bool(white_bg) or bool(torch.rand((1,)) < 0.5)                           # FieldBase.py:299
        params = self._grad_parameters()
        self._train_fwd_kw = dict(fwd_kw)               # per-call switches (iteration / infoinv), re-applied in the node
        rgb, depth = _RenderTrain.apply(self, rays, jitter, white, int(N_samples), int(image_width),
                                        *[p for p in params if p is not None])
        return {'rgb_map': rgb, 'depth_map': depth}

    def _backward(self, rays, jitter, white_bg, N_samples, grad_rgb):
        """-> list of gradient tensors (or None) aligned with _grad_parameters()."""
        lib = _lib.load()
        h = self._live_handle()
        params = self._grad_parameters()
        grads = [None if (p is None or not p.requires_grad) else torch.zeros_like(p, dtype=torch.float32,
                                                                                 memory_format=torch.contiguous_format)
                 for p in params]
        # the library adds into every buffer it is given; parameters that need no gradient get a scratch buffer
        scratch = [g if g is not None else (None if p is None else torch.zeros_like(p, dtype=torch.float32))
                   for g, p in zip(grads, params)]
        ptr = lambda t: None if t is None else t.data_ptr()
        G = _lib.NgfFieldGrads()
        keep = []
        for i in range(3):
            G.plane[i] = ptr(scratch[i])
            G.gauge[i] = ptr(scratch[3 + i])
            keep.append(_f32c(params[i]))                       # the fp32 parameter: exact features for the ReLU masks
            G.plane_param[i] = keep[-1].data_ptr()
        (G.rgb_basis, G.rgb_l1_w, G.rgb_l1_b, G.rgb_l2_w, G.rgb_l2_b, G.rgb_l3_w, G.rgb_l3_b) = (ptr(t) for t in scratch[6:13])
        dens = [ptr(t) for t in scratch[13:]] + [None] * 6
        (G.dens_l1_w, G.dens_l1_b, G.dens_l2_w, G.dens_l2_b, G.dens_l3_w, G.dens_l3_b) = dens[:6]
        g = _f32c(grad_rgb)
        with torch.cuda.device(self.device):
            _lib.check(lib.ngf_field_backward(h, rays.data_ptr(), rays.shape[0], rays.shape[1], int(N_samples),
                                              int(bool(white_bg)), jitter.data_ptr(), g.data_ptr(), C.byref(G),
                                              _cuda_stream_ptr(self.device)), "ngf_field_backward")
        return grads

    def _forward(self, rays_chunk, white_bg, is_train, N_samples, image_width, jitter=None, _white_decided=False,
                 **fwd_kw):
        h = self._ensure_handle()
        lib = _lib.load()
        self._set_switches(lib, h, **fwd_kw)
        rays = rays_chunk
        if rays.device != self.device:
            rays = rays.to(self.device, non_blocking=True)
        rays = _f32c(rays)
        if rays.dim() != 2 or rays.shape[1] < 6:
            raise ValueError(f"rays must be [R, >=6], got {tuple(rays.shape)}")
        R = rays.shape[0]
        rgb = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        depth = torch.empty((R,), dtype=torch.float32, device=self.device)
        acc = torch.empty((R,), dtype=torch.float32, device=self.device)
        if is_train:
            if jitter is None:
                jitter = torch.rand_like(torch.empty((R, 1), dtype=torch.float32))      # CPU draw, as the reference
            jitter = _f32c(jitter.reshape(-1).to(self.device))
            if jitter.numel() != R:
                raise ValueError(f"jitter must have one value per ray ({R}), got {jitter.numel()}")
            if not _white_decided:
                white_bg = bool(white_bg) or bool(torch.rand((1,)) < 0.5)
        with torch.cuda.device(self.device):
            if is_train:
                _lib.check(lib.ngf_field_render_jitter(h, rays.data_ptr(), R, rays.shape[1], int(N_samples),
                                                       int(bool(white_bg)), int(image_width), jitter.data_ptr(),
                                                       rgb.data_ptr(), depth.data_ptr(), acc.data_ptr(), self._mlp_impl,
                                                       _cuda_stream_ptr(self.device)), "ngf_field_render_jitter")
            else:
                _lib.check(lib.ngf_field_render(h, rays.data_ptr(), R, rays.shape[1], int(N_samples),
                                                int(bool(white_bg)), int(image_width), rgb.data_ptr(), depth.data_ptr(),
                                                acc.data_ptr(), self._mlp_impl, _cuda_stream_ptr(self.device)),
                           "ngf_field_render")
        self._last_acc = acc
        return {'rgb_map': rgb, 'depth_map': depth}

    @torch.no_grad()
    def render_host(self, rays_host, rgb_host=None, depth_host=None, white_bg=True, N_samples=-1, image_width=0,
                    **fwd_kw):
        """Whole-frame render through HOST buffers (ngf_field_render_host): H2D, kernels and D2H are chunked and
        overlapped inside the library.  ``rays_host`` should be pinned for full copy bandwidth."""
        h = self._ensure_handle()
        lib = _lib.load()
        self._set_switches(lib, h, **fwd_kw)
        if rays_host.device.type != "cpu" or rays_host.dtype != torch.float32 or not rays_host.is_contiguous():
            raise ValueError("rays_host must be a contiguous fp32 CPU tensor")
        R = rays_host.shape[0]
        if rgb_host is None:
            rgb_host = torch.empty((R, 3), dtype=torch.float32).pin_memory()
        if depth_host is None:
            depth_host = torch.empty((R,), dtype=torch.float32).pin_memory()
        _lib.check(lib.ngf_field_render_host(h, rays_host.data_ptr(), R, rays_host.shape[1], int(N_samples),
                                             int(bool(white_bg)), int(image_width), rgb_host.data_ptr(),
                                             depth_host.data_ptr(), self._mlp_impl), "ngf_field_render_host")
        return rgb_host, depth_host

    @torch.no_grad()
    def render_host_async(self, rays_host, rgb_host, depth_host, white_bg=True, N_samples=-1, image_width=0, **fwd_kw):
        """Enqueue a whole-frame host-buffer render and return a ticket (ngf_field_render_host_async).  The three
        pinned CPU tensors must stay untouched until ``host_wait(ticket)``; consecutive frames overlap on the device."""
        h = self._ensure_handle()
        lib = _lib.load()
        self._set_switches(lib, h, **fwd_kw)
        for t in (rays_host, rgb_host, depth_host):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("render_host_async takes contiguous fp32 CPU tensors")
        ticket = C.c_uint64()
        _lib.check(lib.ngf_field_render_host_async(h, rays_host.data_ptr(), rays_host.shape[0], rays_host.shape[1],
                                                   int(N_samples), int(bool(white_bg)), int(image_width),
                                                   rgb_host.data_ptr(), depth_host.data_ptr(), self._mlp_impl,
                                                   C.byref(ticket)), "ngf_field_render_host_async")
        return int(ticket.value)

    @staticmethod
    def _camera(c2w, H, W, focal, center=None):
        cam = _lib.NgfCamera()
        m = torch.as_tensor(c2w, dtype=torch.float32).cpu().reshape(-1)[:12]
        for k in range(12):
            cam.c2w[k] = float(m[k])
        fx, fy = (focal if isinstance(focal, (tuple, list)) else (focal, focal))
        cx, cy = center if center is not None else (W / 2, H / 2)           # ray_utils.py:38
        cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height = float(fx), float(fy), float(cx), float(cy), int(W), int(H)
        return cam

    @torch.no_grad()
    def render_camera(self, c2w, H, W, focal, center=None, white_bg=True, N_samples=-1, **fwd_kw):
        """Render the H x W frame of a pinhole camera with its rays generated on the device (ngf_field_render_camera):
        what evaluation_path does with get_rays(directions, c2w) + renderer (TriPlane/main.py:155-161).
        -> {'rgb_map': [H*W,3], 'depth_map': [H*W]} on the field's device."""
        h = self._ensure_handle()
        lib = _lib.load()
        self._set_switches(lib, h, **fwd_kw)
        cam = self._camera(c2w, H, W, focal, center)
        R = H * W
        rgb = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        depth = torch.empty((R,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.ngf_field_render_camera(h, C.byref(cam), int(N_samples), int(bool(white_bg)), rgb.data_ptr(),
                                                   depth.data_ptr(), None, self._mlp_impl,
                                                   _cuda_stream_ptr(self.device)), "ngf_field_render_camera")
        return {'rgb_map': rgb, 'depth_map': depth}

    @torch.no_grad()
    def render_camera_host_async(self, c2w, H, W, focal, rgb_host, depth_host, center=None, white_bg=True, N_samples=-1,
                                 **fwd_kw):
        """As render_camera but into pinned CPU tensors, pipelined across frames; returns a ticket for host_wait."""
        h = self._ensure_handle()
        lib = _lib.load()
        self._set_switches(lib, h, **fwd_kw)
        cam = self._camera(c2w, H, W, focal, center)
        ticket = C.c_uint64()
        _lib.check(lib.ngf_field_render_camera_host_async(h, C.byref(cam), int(N_samples), int(bool(white_bg)),
                                                          rgb_host.data_ptr(), depth_host.data_ptr(), self._mlp_impl,
                                                          C.byref(ticket)), "ngf_field_render_camera_host_async")
        return int(ticket.value)

    @torch.no_grad()
    def render_camera_u8_host_async(self, c2w, H, W, focal, u8_host, depth_host=None, center=None, white_bg=True,
                                    N_samples=-1, **fwd_kw):
        """evaluation_path's per-frame job (TriPlane/main.py:155-161,116): pose in, uint8 image [H*W,3] (pinned CPU
        tensor) out, depth optional; pipelined across frames; returns a ticket for host_wait."""
        h = self._ensure_handle()
        lib = _lib.load()
        self._set_switches(lib, h, **fwd_kw)
        if u8_host.dtype != torch.uint8 or u8_host.device.type != "cpu" or not u8_host.is_contiguous():
            raise ValueError("u8_host must be a contiguous uint8 CPU tensor")
        cam = self._camera(c2w, H, W, focal, center)
        ticket = C.c_uint64()
        _lib.check(lib.ngf_field_render_camera_u8_host_async(h, C.byref(cam), int(N_samples), int(bool(white_bg)),
                                                             u8_host.data_ptr(),
                                                             None if depth_host is None else depth_host.data_ptr(),
                                                             self._mlp_impl, C.byref(ticket)),
                   "ngf_field_render_camera_u8_host_async")
        return int(ticket.value)

    def host_wait(self, ticket: int):
        _lib.check(_lib.load().ngf_field_host_wait(self._live_handle(), int(ticket)), "ngf_field_host_wait")

    def last_stats(self) -> dict:
        """Counters of the last device-side render (forces a stream sync)."""
        st = _lib.NgfStats()
        _lib.check(_lib.load().ngf_field_stats(self._live_handle(), C.byref(st), _cuda_stream_ptr(self.device)))
        return {k: int(getattr(st, k)) for k, _ in st._fields_}

    def kernel_timing(self, capacity: int):
        """Arm (capacity > 0) or disarm (0) CUDA-event timing of the march / colour kernels (ngf_field_timing_begin)."""
        _lib.check(_lib.load().ngf_field_timing_begin(self._live_handle(), int(capacity)))

    def kernel_timing_read(self):
        """-> (march+colour pairs timed, march milliseconds, colour milliseconds) since the last read."""
        n, a, b = C.c_int32(), C.c_double(), C.c_double()
        _lib.check(_lib.load().ngf_field_timing_read(self._live_handle(), C.byref(n), C.byref(a), C.byref(b)))
        return int(n.value), float(a.value), float(b.value)

    # ------------------------------------------------------------------ point-wise API parity
    @torch.no_grad()
    def sample_ray(self, rays_o, rays_d, is_train=True, N_samples=-1, jitter=None):
        """Reference: Base.sample_ray (FieldBase.py:118-137).  -> (rays_pts [R,S,3], interpx [R,S],
        ~mask_outbbox [R,S] bool).  ``is_train``: one ``u ~ U[0,1)`` per ray shifts its samples, drawn as the
        reference does (CPU ``torch.rand_like``) unless ``jitter`` ([R] or [R,1]) is given."""
        h = self._ensure_handle()
        S = N_samples if N_samples > 0 else self.nSamples
        rays = _f32c(torch.cat([rays_o, rays_d], -1).to(self.device))
        R = rays.shape[0]
        pts = torch.empty((R, S, 3), dtype=torch.float32, device=self.device)
        t = torch.empty((R, S), dtype=torch.float32, device=self.device)
        inside = torch.empty((R, S), dtype=torch.uint8, device=self.device)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            if is_train:
                if jitter is None:
                    jitter = torch.rand_like(torch.empty((R, 1), dtype=torch.float32))
                jitter = _f32c(jitter.reshape(-1).to(self.device))
                if jitter.numel() != R:
                    raise ValueError(f"jitter must have one value per ray ({R}), got {jitter.numel()}")
                _lib.check(lib.ngf_field_sample_ray_jitter(h, rays.data_ptr(), R, 6, S, jitter.data_ptr(), pts.data_ptr(),
                                                           t.data_ptr(), inside.data_ptr(),
                                                           _cuda_stream_ptr(self.device)))
            else:
                _lib.check(lib.ngf_field_sample_ray(h, rays.data_ptr(), R, 6, S, pts.data_ptr(), t.data_ptr(),
                                                    inside.data_ptr(), _cuda_stream_ptr(self.device)))
        return pts, t, inside.bool()

    @torch.no_grad()
    def _alpha_keep(self, xyz):
        h = self._ensure_handle()
        pts = _f32c(xyz.to(self.device)).view(-1, 3)
        keep = torch.empty((pts.shape[0],), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_field_alpha_keep(h, pts.data_ptr(), pts.shape[0], keep.data_ptr(),
                                                        _cuda_stream_ptr(self.device)))
        return keep.bool()

    @torch.no_grad()
    def _alpha_value(self, xyz):
        h = self._ensure_handle()
        pts = _f32c(xyz.to(self.device)).view(-1, 3)
        out = torch.empty((pts.shape[0],), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_field_alpha_value(h, pts.data_ptr(), pts.shape[0], out.data_ptr(),
                                                         _cuda_stream_ptr(self.device)))
        return out.view(xyz.shape[:-1])

    @torch.no_grad()
    def _coords(self, valid_xyz, gauge_on: bool):
        h = self._ensure_handle()
        xyz = _f32c(valid_xyz.to(self.device)).view(-1, 3)
        n = xyz.shape[0]
        outs = [torch.empty((n, 2), dtype=torch.float32, device=self.device) for _ in range(3)]
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_field_gauge(h, xyz.data_ptr(), n, int(gauge_on), outs[0].data_ptr(),
                                                   outs[1].data_ptr(), outs[2].data_ptr(),
                                                   _cuda_stream_ptr(self.device)))
        return tuple(outs)

    @torch.no_grad()
    def _density(self, xy, yz, xz):
        h = self._ensure_handle()
        a, b, c = (_f32c(t.to(self.device)).view(-1, 2) for t in (xy, yz, xz))
        n = a.shape[0]
        out = torch.empty((n,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_field_density(h, a.data_ptr(), b.data_ptr(), c.data_ptr(), n, out.data_ptr(),
                                                     _cuda_stream_ptr(self.device)))
        return out

    @torch.no_grad()
    def _rgb(self, xy, yz, xz, view_sampled):
        h = self._ensure_handle()
        a, b, c = (_f32c(t.to(self.device)).view(-1, 2) for t in (xy, yz, xz))
        v = _f32c(view_sampled.to(self.device)).view(-1, 3)
        n = a.shape[0]
        out = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_field_rgb(h, a.data_ptr(), b.data_ptr(), c.data_ptr(), v.data_ptr(), n,
                                                 out.data_ptr(), self._mlp_impl, _cuda_stream_ptr(self.device)))
        return out

    @torch.no_grad()
    def _sigma_world(self, xyz_locs, use_gauge=False):
        h = self._ensure_handle()
        pts = _f32c(xyz_locs.to(self.device)).view(-1, 3)
        out = torch.empty((pts.shape[0],), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngf_field_sigma_world(h, pts.data_ptr(), pts.shape[0], int(use_gauge),
                                                         out.data_ptr(), _cuda_stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ occupancy maintenance (FieldBase.py:140-246)
    # These run on the same point-wise kernels as the renderer (ngf_field_sigma_world / sample_ray / alpha_keep); the whole
    # lattice or ray set goes to the device in one call instead of the reference's per-slice / per-chunk Python loops.
    @torch.no_grad()
    def compute_alpha(self, xyz_locs, length=1, **kw):
        """Reference: Base.compute_alpha (FieldBase.py:140-159): alpha = 1 - exp(-sigma * length), field evaluated with
        the gauge off and the current alpha mask applied."""
        self._apply_alpha_kw(**kw)
        sigma = self._sigma_world(xyz_locs.reshape(-1, 3), use_gauge=False)
        return (1 - torch.exp(-sigma * length)).view(xyz_locs.shape[:-1])

    def _apply_alpha_kw(self, **kw):
        pass

    def _lattice(self, gs):
        """World positions of the gs[0] x gs[1] x gs[2] lattice spanning the box, [gx, gy, gz, 3] on the device.  The
        per-axis blend aabb0*(1-s) + aabb1*s with s = torch.linspace(0, 1, g) is evaluated on the CPU exactly as
        FieldBase.py:165-171 does, then broadcast."""
        lo, hi = self.aabb[0].detach().float().cpu(), self.aabb[1].detach().float().cpu()
        axes = []
        for k in range(3):
            t = torch.linspace(0, 1, gs[k])
            axes.append((lo[k] * (1 - t) + hi[k] * t).to(self.device))
        gx, gy, gz = torch.meshgrid(axes[0], axes[1], axes[2], indexing="ij")
        return torch.stack((gx, gy, gz), -1)

    @torch.no_grad()
    def getDenseAlpha(self, gridSize=None, **kw):
        """Reference: Base.getDenseAlpha (FieldBase.py:161-177) -> (alpha [gx,gy,gz], dense_xyz [gx,gy,gz,3])."""
        gs = [int(g) for g in (self.gridSize if gridSize is None else gridSize)]
        dense_xyz = self._lattice(gs)
        alpha = self.compute_alpha(dense_xyz.view(-1, 3), self.stepSize, **kw).view(gs)
        return alpha, dense_xyz

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=(200, 200, 200), **kw):
        """Reference: Base.updateAlphaMask (FieldBase.py:179-215): dense alpha, 3x3x3 dilation, threshold at
        alphaMask_thres -> new AlphaGridMask over the current box; returns the tight box of the occupied lattice points."""
        gs = [int(g) for g in gridSize]
        alpha, dense_xyz = self.getDenseAlpha(gs, **kw)
        # the mask volume is indexed [z, y, x] (FieldBase.py:183-184)
        vol = alpha.clamp(0, 1).permute(2, 1, 0).contiguous()
        vol = F.max_pool3d(vol[None, None], kernel_size=3, padding=1, stride=1)[0, 0]
        occupied = vol >= self.alphaMask_thres
        self.alphaMask = AlphaGridMask(self.device, self.aabb, occupied.float())
        pts = dense_xyz.permute(2, 1, 0, 3)[occupied]
        return torch.stack((pts.amin(0), pts.amax(0)))

    @torch.no_grad()
    def filtering_rays(self, all_rays, all_rgbs, N_samples=256, chunk=10240 * 5, bbox_only=False):
        """Reference: Base.filtering_rays (FieldBase.py:217-246): keep the rays that cross the box (bbox_only) or that
        have at least one sample the alpha mask keeps."""
        flat = all_rays.reshape(-1, all_rays.shape[-1])
        keep = torch.empty(flat.shape[0], dtype=torch.bool)
        step = int(chunk) * 8                                  # the device takes far larger pieces than the reference's
        for s in range(0, flat.shape[0], step):
            r = flat[s:s + step].to(self.device)
            o, d = r[:, :3], r[:, 3:6]
            if bbox_only:
                safe = torch.where(d == 0, torch.full_like(d, 1e-6), d)
                ta, tb = (self.aabb[1] - o) / safe, (self.aabb[0] - o) / safe
                hit = torch.maximum(ta, tb).amin(-1) > torch.minimum(ta, tb).amax(-1)
            else:
                pts, _, _ = self.sample_ray(o, d, N_samples=N_samples, is_train=False)
                hit = self._alpha_keep(pts.view(-1, 3)).view(pts.shape[:-1]).any(-1)
            keep[s:s + step] = hit.cpu()
        keep = keep.view(all_rgbs.shape[:-1])
        return all_rays[keep], all_rgbs[keep]
