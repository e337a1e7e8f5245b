"""``renderer`` with the reference's signature (TriPlane/main.py:60-71, InfoInv/main.py:61-72) and the multi-GPU
ray-sharded frame render (interleaved ray blocks per rank + one all-gather of the frame).
"""
from __future__ import annotations

import torch

from . import _lib
from .field_base import _cuda_stream_ptr


def renderer(rays, field, chunk=4096, N_samples=-1, white_bg=True, is_train=False, device='cuda', image_width=0,
             **fwd_kw):
    """Reference: ``renderer`` (TriPlane/main.py:60-71).  The reference slices the frame into ``chunk``-ray pieces
    and calls the field once per piece; the fused kernel takes the whole frame in one launch, so ``chunk`` is
    accepted for signature compatibility and ignored.  ``fwd_kw`` defaults to what the reference passes
    (``iteration=30001`` for TriPlane, main.py:67; ``infoinv`` for InfoInv).  Gradients are never recorded (the field's
    forward runs under ``torch.no_grad()``); ``is_train=True`` with autograd enabled is refused by the field, as the
    training loop (main.py:272) would otherwise get detached tensors."""
    if not fwd_kw and hasattr(field, "gauge_start"):
        fwd_kw = {"iteration": 30001}
    out = field(rays, is_train=is_train, white_bg=white_bg, N_samples=N_samples, image_width=image_width, **fwd_kw)
    return out['rgb_map'], out['depth_map']


render_rays = renderer          # the upstream TensoRF-style name BASELINE.json's north_star uses for the same call


@torch.no_grad()
def render_frames(frames, field, N_samples=-1, white_bg=True, image_width=0, depth=2, **fwd_kw):
    """Render a sequence of frames whose rays live in pinned CPU memory, as the reference's ``evaluation`` loop does
    (TriPlane/main.py:89-100: ``renderer`` per frame, then ``.cpu()``), but pipelined: while frame k is on the device,
    frame k+1 is being uploaded and frame k-1 downloaded (ngf_field_render_host_async).  ``frames`` yields [R, C] fp32
    CPU tensors (pinned for full copy bandwidth); yields (rgb [R,3], depth [R]) pinned CPU tensors in order.  A yielded
    pair is reused ``depth + 1`` frames later, so consume (or copy) it before advancing that far."""
    if not fwd_kw and hasattr(field, "gauge_start"):
        fwd_kw = {"iteration": 30001}
    bufs, pending = [], []
    for k, rays in enumerate(frames):
        if len(bufs) <= k % (depth + 1):
            bufs.append((torch.empty((rays.shape[0], 3)).pin_memory(), torch.empty((rays.shape[0],)).pin_memory()))
        rgb, dep = bufs[k % (depth + 1)]
        if rgb.shape[0] != rays.shape[0]:
            rgb, dep = torch.empty((rays.shape[0], 3)).pin_memory(), torch.empty((rays.shape[0],)).pin_memory()
            bufs[k % (depth + 1)] = (rgb, dep)
        t = field.render_host_async(rays, rgb, dep, white_bg=white_bg, N_samples=N_samples, image_width=image_width,
                                    **fwd_kw)
        pending.append((t, rgb, dep))
        if len(pending) > depth - 1:
            t0, r0, d0 = pending.pop(0)
            field.host_wait(t0)
            yield r0, d0
    for t0, r0, d0 in pending:
        field.host_wait(t0)
        yield r0, d0


@torch.no_grad()
def frame_post(rgb_map, gt_rgb=None, want_u8=True):
    """Device-side tail of the reference's ``evaluation`` (TriPlane/main.py:99-116): -> (uint8 image or None, PSNR or
    None).  ``rgb_map`` / ``gt_rgb``: CUDA fp32 tensors of equal shape."""
    import math
    rgb = rgb_map.contiguous()
    n = rgb.numel()
    u8 = torch.empty(rgb.shape, dtype=torch.uint8, device=rgb.device) if want_u8 else None
    sse = torch.zeros(1, dtype=torch.float64, device=rgb.device) if gt_rgb is not None else None
    gt = None if gt_rgb is None else gt_rgb.to(rgb.device).float().contiguous()
    with torch.cuda.device(rgb.device):
        _lib.check(_lib.load().ngf_frame_post(rgb.data_ptr(), None if gt is None else gt.data_ptr(), n,
                                              None if u8 is None else u8.data_ptr(),
                                              None if sse is None else sse.data_ptr(), _cuda_stream_ptr(rgb.device)))
    psnr = None
    if sse is not None:
        psnr = -10.0 * math.log(float(sse.item()) / n) / math.log(10.0)
    return u8, psnr


@torch.no_grad()
def visualize_depth(depth_map, minmax=None):
    """Device-side ``visualize_depth_numpy`` (TriPlane/utils.py:32-47; evaluation() calls it with near_far, main.py:102):
    depth [H, W] CUDA fp32 -> (uint8 [H, W, 3] in OpenCV's B, G, R order, [mi, ma])."""
    d = depth_map.contiguous().float()
    if minmax is None:
        x = torch.nan_to_num(d)
        mi, ma = float(x[x > 0].min()), float(x.max())
    else:
        mi, ma = float(minmax[0]), float(minmax[1])
    out = torch.empty(tuple(d.shape) + (3,), dtype=torch.uint8, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().ngf_depth_colormap(d.data_ptr(), d.numel(), mi, ma, out.data_ptr(), _cuda_stream_ptr(d.device)),
                   "ngf_depth_colormap")
    return out, [mi, ma]


# ---------------------------------------------------------------------------------------------------------------
# ray sharding: global ray g belongs to rank (g // block) % world; local index (g // (block*world))*block + g % block
# ---------------------------------------------------------------------------------------------------------------
def shard_count(n_rays: int, block: int, rank: int, world: int) -> int:
    n = _lib.load().ngf_shard_count(n_rays, block, rank, world)
    if n < 0:
        raise ValueError("bad shard arguments")
    return int(n)


def shard_index(n_rays: int, block: int, rank: int, world: int) -> torch.Tensor:
    """Global indices of the rays rank owns, in local order (host-side index arithmetic, used by CPU tests and as
    the specification of ngf_shard_gather/scatter)."""
    g = torch.arange(n_rays)
    return g[((g // block) % world) == rank]


def shard_rays(rays: torch.Tensor, block: int, rank: int, world: int) -> torch.Tensor:
    """Rows of ``rays`` [R, C] this rank renders.  CUDA tensors go through ngf_shard_gather."""
    if world == 1:
        return rays
    R, Cn = rays.shape
    if rays.is_cuda:
        n = shard_count(R, block, rank, world)
        out = torch.empty((n, Cn), dtype=torch.float32, device=rays.device)
        src = rays.contiguous()
        with torch.cuda.device(rays.device):
            _lib.check(_lib.load().ngf_shard_gather(src.data_ptr(), R, Cn, block, rank, world, out.data_ptr(),
                                                    _cuda_stream_ptr(rays.device)))
        return out
    return rays[shard_index(R, block, rank, world)]


def unshard_frame(gathered: torch.Tensor, n_rays: int, block: int, world: int) -> torch.Tensor:
    """Inverse of the all-gather: ``gathered`` [world, max_shard, C] (rank-major) -> [n_rays, C] in frame order."""
    world_, max_shard, Cn = gathered.shape
    assert world_ == world
    if gathered.is_cuda:
        out = torch.empty((n_rays, Cn), dtype=torch.float32, device=gathered.device)
        src = gathered.contiguous()
        with torch.cuda.device(gathered.device):
            _lib.check(_lib.load().ngf_shard_scatter(src.data_ptr(), n_rays, Cn, block, world, max_shard,
                                                     out.data_ptr(), _cuda_stream_ptr(gathered.device)))
        return out
    out = torch.empty((n_rays, Cn), dtype=gathered.dtype)
    for r in range(world):
        idx = shard_index(n_rays, block, r, world)
        out[idx] = gathered[r, :idx.numel()]
    return out


def frame_allgather(local: torch.Tensor, n_rays: int, block: int, group=None) -> torch.Tensor:
    """All-gather the per-rank [n_local, C] results of a ray-sharded frame and restore frame order.
    One collective per frame (NCCL on GPU; gloo in the CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    max_shard = -(-(-(-n_rays // block)) // world) * block          # ceil(ceil(n/block)/world)*block
    buf = torch.zeros((max_shard, local.shape[1]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = torch.empty((world, max_shard, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * max_shard, -1), buf, group=group)
    return unshard_frame(out, n_rays, block, world)


@torch.no_grad()
def render_frame_sharded(rays, field, block=2048, N_samples=-1, white_bg=True, group=None, **fwd_kw):
    """Render one frame with its rays dealt to the ranks of ``group`` in interleaved blocks and return the full
    [R, 4] (rgb, depth) frame on every rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    R = rays.shape[0]
    mine = shard_rays(rays, block, rank, world)
    rgb, depth = renderer(mine, field, N_samples=N_samples, white_bg=white_bg, **fwd_kw)
    local = torch.cat([rgb, depth[:, None]], 1)
    return frame_allgather(local, R, block, group)


class _DevView:
    """A [rows, cols] fp32 device array owned by the library, exposed through __cuda_array_interface__."""

    def __init__(self, ptr: int, rows: int, cols: int):
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class FrameComm:
    """Ray-sharded frames with the all-gather inside the C ABI (ngf_comm_* / ngf_field_render_sharded /
    ngf_frame_allgather, include/ngf_b200.h): every rank renders its interleaved ``block``-ray blocks
    of a ``n_rays_total``-ray batch straight into frame order, and the rows travel to the other ranks over NVLink peer
    mappings — ``mode="copy"``: one strided copy per peer on the copy engines (no SM); ``mode="store"``: stored by the
    render's last kernel itself.  torch.distributed is used once, to exchange the 128-byte buffer handles.

        comm = FrameComm(field, n_rays_total, block)               # collective: every rank of `group` calls it
        t = comm.submit(my_rays_dev, N_samples=192, image_width=800)
        frame = comm.result(t)                                     # [n_rays_total, 4] (r, g, b, depth), stream-ordered
        ...; comm.release(t)                                       # peers may overwrite the buffer
    """

    def __init__(self, field, n_rays_total: int, block: int, group=None, mode: str = "copy", n_slots: int = 3):
        import ctypes as C
        import torch.distributed as dist
        lib = _lib.load()
        self.field, self.n, self.block, self.group = field, int(n_rays_total), int(block), group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        self.mode = mode
        dev = field.device
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        _lib.check(lib.ngf_comm_init(self.rank, self.world, dev_index, self.n, self.block, int(n_slots),
                                     {"copy": _lib.COMM_COPY, "store": _lib.COMM_STORE}[mode], C.byref(h)), "ngf_comm_init")
        self._h = h
        self.n_local = int(lib.ngf_comm_local_rays(h))
        if self.world > 1:
            nb = int(lib.ngf_comm_handle_bytes())
            mine = C.create_string_buffer(nb)
            _lib.check(lib.ngf_comm_export(h, mine), "ngf_comm_export")
            blobs = [None] * self.world
            dist.all_gather_object(blobs, bytes(mine.raw), group=group)
            _lib.check(lib.ngf_comm_connect(h, b"".join(blobs)), "ngf_comm_connect")
            dist.barrier(group)             # nobody starts writing into a peer that has not mapped / zeroed its flags

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().ngf_comm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _field_handle(self, fwd_kw):
        fh = self.field._ensure_handle()
        self.field._set_switches(_lib.load(), fh, **fwd_kw)
        return fh

    @torch.no_grad()
    def submit(self, rays_local, N_samples=-1, white_bg=True, image_width=0, **fwd_kw):
        """Render this rank's rays (CUDA tensor [n_local, C >= 6], local order) of the next batch on the current stream
        and start the exchange.  -> ticket."""
        import ctypes as C
        if not fwd_kw and hasattr(self.field, "gauge_start"):
            fwd_kw = {"iteration": 30001}
        fh = self._field_handle(fwd_kw)
        rays = rays_local
        if rays.dtype != torch.float32 or not rays.is_contiguous() or rays.device != self.field.device:
            rays = rays.to(self.field.device, torch.float32).contiguous()
        t = C.c_uint64()
        with torch.cuda.device(self.field.device):
            _lib.check(_lib.load().ngf_field_render_sharded(fh, self._h, rays.data_ptr(), rays.shape[0], rays.shape[1],
                                                            int(N_samples), int(bool(white_bg)), int(image_width),
                                                            self.field._mlp_impl, _cuda_stream_ptr(self.field.device),
                                                            C.byref(t)), "ngf_field_render_sharded")
        return int(t.value)

    def result(self, ticket: int) -> torch.Tensor:
        """Make the current stream wait for every rank's rows of ``ticket`` and return the gathered [n, 4] batch (a view of
        the library's frame buffer: valid until ``release(ticket)`` and n_slots - 1 further submits)."""
        import ctypes as C
        p = C.c_void_p()
        with torch.cuda.device(self.field.device):
            _lib.check(_lib.load().ngf_frame_allgather(self._h, int(ticket), _cuda_stream_ptr(self.field.device),
                                                       C.byref(p)), "ngf_frame_allgather")
        return torch.as_tensor(_DevView(p.value, self.n, 4), device=self.field.device)

    def release(self, ticket: int):
        with torch.cuda.device(self.field.device):
            _lib.check(_lib.load().ngf_frame_release(self._h, int(ticket), _cuda_stream_ptr(self.field.device)),
                       "ngf_frame_release")

    @torch.no_grad()
    def submit_host(self, rays_local_host, frame_host, first_row=0, n_rows=None, N_samples=-1, white_bg=True,
                    image_width=0, **fwd_kw):
        """Host-buffer pipeline (ngf_field_render_sharded_host_async): pinned rays of this rank in, rows
        [first_row, first_row + n_rows) of the gathered batch out into the pinned ``frame_host`` [n_rows, 4].  -> ticket
        for ``wait``; consecutive batches overlap upload / render / exchange / download."""
        import ctypes as C
        if not fwd_kw and hasattr(self.field, "gauge_start"):
            fwd_kw = {"iteration": 30001}
        fh = self._field_handle(fwd_kw)
        for t_ in (rays_local_host, frame_host):
            if t_.device.type != "cpu" or t_.dtype != torch.float32 or not t_.is_contiguous():
                raise ValueError("submit_host takes contiguous fp32 CPU tensors")
        n_rows = frame_host.shape[0] if n_rows is None else int(n_rows)
        t = C.c_uint64()
        _lib.check(_lib.load().ngf_field_render_sharded_host_async(
            fh, self._h, rays_local_host.data_ptr(), rays_local_host.shape[0], rays_local_host.shape[1], int(N_samples),
            int(bool(white_bg)), int(image_width), self.field._mlp_impl, frame_host.data_ptr(), int(first_row), n_rows,
            C.byref(t)), "ngf_field_render_sharded_host_async")
        return int(t.value)

    def wait(self, ticket: int):
        _lib.check(_lib.load().ngf_comm_wait(self._h, int(ticket)), "ngf_comm_wait")

    def _poses(self, poses):
        p = torch.as_tensor(poses, dtype=torch.float32).reshape(-1, 12)
        return p.contiguous()

    @torch.no_grad()
    def submit_camera(self, poses, H, W, focal, center=None, N_samples=-1, white_bg=True, **fwd_kw):
        """A batch of ``len(poses)`` camera frames (poses [F,3,4] camera-to-world, one pinhole model), this rank's rays
        generated in the march kernel (ngf_field_render_sharded_camera).  -> ticket for result() / release()."""
        import ctypes as C
        if not fwd_kw and hasattr(self.field, "gauge_start"):
            fwd_kw = {"iteration": 30001}
        fh = self._field_handle(fwd_kw)
        p = self._poses(poses).to(self.field.device)
        cam = self.field._camera(p[0].cpu(), H, W, focal, center)
        t = C.c_uint64()
        with torch.cuda.device(self.field.device):
            _lib.check(_lib.load().ngf_field_render_sharded_camera(fh, self._h, C.byref(cam), p.data_ptr(), p.shape[0], int(N_samples),
                                                                   int(bool(white_bg)), self.field._mlp_impl,
                                                                   _cuda_stream_ptr(self.field.device), C.byref(t)),
                       "ngf_field_render_sharded_camera")
        self._keep_poses = p
        return int(t.value)

    @torch.no_grad()
    def submit_camera_host(self, poses_host, H, W, focal, u8_host, first_row=0, n_rows=None, center=None, N_samples=-1,
                           white_bg=True, **fwd_kw):
        """Host pipeline of a camera batch (ngf_field_render_sharded_camera_u8_host_async): poses [F,3,4] (pinned CPU fp32) in,
        rows [first_row, first_row + n_rows) of the gathered batch out as uint8 rgb in the pinned ``u8_host``; -> ticket for
        ``wait``."""
        import ctypes as C
        if not fwd_kw and hasattr(self.field, "gauge_start"):
            fwd_kw = {"iteration": 30001}
        fh = self._field_handle(fwd_kw)
        if poses_host.device.type != "cpu" or poses_host.dtype != torch.float32 or not poses_host.is_contiguous():
            raise ValueError("poses_host must be a contiguous fp32 CPU tensor [F, 3, 4]")
        if u8_host.device.type != "cpu" or u8_host.dtype != torch.uint8 or not u8_host.is_contiguous():
            raise ValueError("u8_host must be a contiguous uint8 CPU tensor")
        n_frames = poses_host.numel() // 12
        n_rows = u8_host.shape[0] if n_rows is None else int(n_rows)
        cam = self.field._camera(poses_host.reshape(-1, 12)[0], H, W, focal, center)
        t = C.c_uint64()
        _lib.check(_lib.load().ngf_field_render_sharded_camera_u8_host_async(
            fh, self._h, C.byref(cam), poses_host.data_ptr(), n_frames, int(N_samples), int(bool(white_bg)), self.field._mlp_impl,
            u8_host.data_ptr(), int(first_row), n_rows, C.byref(t)), "ngf_field_render_sharded_camera_u8_host_async")
        return int(t.value)


class ShardedFrameRenderer:
    """Ray-sharded multi-GPU rendering, one all-gather of the rendered batch per step, overlapped with the next batch.

    ``mode="copy"`` / ``"store"`` (default copy): the exchange runs inside the C ABI over NVLink peer memory (FrameComm).
    ``mode="nccl"``: the render is followed by one ``torch.distributed.all_gather_into_tensor`` (NCCL) and a re-ordering
    kernel on a side stream, double-buffered — kept as the library baseline the peer-memory path is compared with.

        r = ShardedFrameRenderer(field, n_rays_total, block)
        t = r.submit(my_shard_of_rays)          # returns at once
        frame = r.result(t)                      # [n_rays_total, 4] on this rank's device, ordered as the input batch
        r.release(t)                             # done reading (peer-memory modes; no-op for nccl)
    """

    def __init__(self, field, n_rays_total: int, block: int = 2000, group=None, mode: str = "copy"):
        import torch.distributed as dist
        self.field, self.n, self.block, self.group = field, int(n_rays_total), int(block), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.mode = mode
        self.comm = None
        if mode != "nccl":
            self.comm = FrameComm(field, n_rays_total, block, group, mode)
            return
        dev = field.device
        self.max_shard = -(-(-(-self.n // block)) // self.world) * block
        self.side = torch.cuda.Stream(device=dev)
        self.local = [torch.zeros((self.max_shard, 4), device=dev) for _ in range(2)]
        self.gathered = [torch.empty((self.world, self.max_shard, 4), device=dev) for _ in range(2)]
        self.frame = [torch.empty((self.n, 4), device=dev) for _ in range(2)]
        self.rendered = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.read = [torch.cuda.Event() for _ in range(2)]
        self.d2h = torch.cuda.Stream(device=dev)
        self.count = 0

    @torch.no_grad()
    def submit(self, rays_local, N_samples=-1, white_bg=True, **fwd_kw):
        if self.comm is not None:
            return self.comm.submit(rays_local, N_samples=N_samples, white_bg=white_bg, **fwd_kw)
        import torch.distributed as dist
        b = self.count % 2
        cur = torch.cuda.current_stream(self.field.device)
        cur.wait_event(self.done[b])                       # buffers of this parity are free again
        rgb, depth = renderer(rays_local, self.field, N_samples=N_samples, white_bg=white_bg, **fwd_kw)
        n = rgb.shape[0]
        torch.cat([rgb, depth[:, None]], 1, out=self.local[b][:n])
        self.rendered[b].record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.rendered[b])
            self.side.wait_event(self.read[b])             # a pending download of this parity's frame buffer
            dist.all_gather_into_tensor(self.gathered[b].view(self.world * self.max_shard, 4), self.local[b],
                                        group=self.group)
            src = self.gathered[b]
            _lib.check(_lib.load().ngf_shard_scatter(src.data_ptr(), self.n, 4, self.block, self.world, self.max_shard,
                                                     self.frame[b].data_ptr(), int(self.side.cuda_stream)))
            self.done[b].record(self.side)
        self.count += 1
        return self.count - 1

    def download(self, ticket: int, out_host: torch.Tensor, n_rows: int = None):
        """nccl mode: copy the first ``n_rows`` rows (default: all) of batch ``ticket``'s gathered frame into the pinned CPU
        tensor ``out_host`` on a dedicated stream, without stalling the compute stream; returns the event to wait on."""
        if self.comm is not None:
            raise RuntimeError("download() belongs to mode='nccl'; the peer-memory modes use FrameComm.submit_host")
        if ticket < self.count - 2 or ticket >= self.count:
            raise ValueError("only the last two submitted batches are still buffered")
        b = ticket % 2
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.done[b])
            out_host.copy_(self.frame[b][:n_rows] if n_rows is not None else self.frame[b], non_blocking=True)
            self.read[b].record(self.d2h)
        return self.read[b]

    def result(self, ticket: int):
        """Make the current stream wait for batch ``ticket`` (it must be one of the last two submitted) and return its
        frame buffer; the buffer is overwritten two submissions later."""
        if self.comm is not None:
            return self.comm.result(ticket)
        if ticket < self.count - 2 or ticket >= self.count:
            raise ValueError("only the last two submitted batches are still buffered")
        torch.cuda.current_stream(self.field.device).wait_event(self.done[ticket % 2])
        return self.frame[ticket % 2]

    def release(self, ticket: int):
        if self.comm is not None:
            self.comm.release(ticket)
