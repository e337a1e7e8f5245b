"""Launched under torchrun (one rank per GPU): render the golden case tp_fog_c1 with its rays sharded across the
ranks and compare the gathered frame with the reference golden vector on every rank, for every exchange variant:
  nccl   interleaved blocks + one torch.distributed NCCL all-gather (ngf_b200.render.render_frame_sharded), plain and
         double-buffered;
  copy   the C ABI's peer-memory all-gather on the copy engines (ngf_comm_* / ngf_field_render_sharded /
         ngf_frame_allgather), device-resident and through the host-buffer pipeline;
  store  the same with the rows stored into the peers' buffers by the render's last kernel.
Several batches are kept in flight so that slot reuse and the freed / arrived flags are exercised."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from oracle import cases as K
from helpers import build_cuda_field, load_golden
import ngf_b200
from ngf_b200.render import shard_rays

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
case = K.CASE_BY_NAME["tp_fog_c1"]
state, kw, occ, rays = K.build_inputs(case)
f = build_cuda_field(case, state, kw, occ, device=dev)
gold = load_golden("tp_fog_c1")
BLOCK = 96
N = rays.shape[0]
report = []


def errs(fr):
    fr = fr.detach().cpu().numpy()
    return float(np.abs(fr[:, :3] - gold["rgb"]).max()), float(np.abs(fr[:, 3] - gold["depth"]).max())


def check(tag, fr):
    e_rgb, e_dep = errs(fr)
    good = e_rgb < 1e-3 and e_dep < 2e-3
    report.append(f"{tag}: rgb {e_rgb:.2e} depth {e_dep:.2e} {'ok' if good else 'MISMATCH'}")
    return good


ok = True
frame = ngf_b200.render_frame_sharded(rays.to(dev), f, block=BLOCK, N_samples=case.n_samples, white_bg=True, iteration=30001)
torch.cuda.synchronize()
ok &= check("nccl plain", frame)
mine = shard_rays(rays.to(dev), BLOCK, rank, world)
mine_host = mine.cpu().pin_memory()

for mode in ("nccl", "copy", "store"):
    sr = ngf_b200.ShardedFrameRenderer(f, N, block=BLOCK, mode=mode)
    # 7 batches, the consumer trails the producer by one batch (as bench.py does)
    pending = None
    for k in range(7):
        t = sr.submit(mine, N_samples=case.n_samples, white_bg=True, iteration=30001)
        if pending is not None:
            fr = sr.result(pending).clone()
            sr.release(pending)
            ok &= check(f"{mode} batch {k - 1}", fr)
        pending = t
    fr = sr.result(pending).clone()
    sr.release(pending)
    torch.cuda.synchronize()
    ok &= check(f"{mode} batch 6", fr)
    if sr.comm is not None:
        # host-buffer pipeline: every rank downloads a different row range of the gathered frame
        rows = N // world
        outs = [torch.zeros((rows, 4)).pin_memory() for _ in range(3)]
        tickets = []
        for k in range(6):
            tickets.append(sr.comm.submit_host(mine_host, outs[k % 3], first_row=rank * rows, N_samples=case.n_samples,
                                               white_bg=True, iteration=30001))
            if len(tickets) > 2:
                sr.comm.wait(tickets.pop(0))
        for t in tickets:
            sr.comm.wait(t)
        for k, o in enumerate(outs):
            e_rgb = float(np.abs(o[:, :3].numpy() - gold["rgb"][rank * rows:(rank + 1) * rows]).max())
            e_dep = float(np.abs(o[:, 3].numpy() - gold["depth"][rank * rows:(rank + 1) * rows]).max())
            good = e_rgb < 1e-3 and e_dep < 2e-3
            ok &= good
            report.append(f"{mode} host buffer {k}: rgb {e_rgb:.2e} depth {e_dep:.2e} {'ok' if good else 'MISMATCH'}")
        dist.barrier()
        sr.comm.close()
# camera batches through the host pipeline: `world` frames per batch, every rank downloads its own frame as uint8
H, W = 48, 80
focal = K.synth.FOCAL_800 * 64 / 800
poses = torch.stack([K.synth.look_at_c2w(*K.synth.pose_angles(p)) for p in range(4, 4 + world)]).contiguous()
per = H * W
o = f.render_camera(poses[rank], H, W, focal, white_bg=True, N_samples=64, iteration=30001)
want = (o["rgb_map"].cpu().numpy() * 255).astype("uint8")
for mode in ("copy", "store"):
    comm = ngf_b200.FrameComm(f, world * per, block=4 * W, mode=mode)
    u8 = [torch.zeros((per, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
    ph = poses.pin_memory()
    tickets = []
    for k in range(6):
        tickets.append(comm.submit_camera_host(ph, H, W, focal, u8[k % 3], first_row=rank * per, N_samples=64, white_bg=True,
                                               iteration=30001))
        if len(tickets) > 2:
            comm.wait(tickets.pop(0))
    for t in tickets:
        comm.wait(t)
    for k, b in enumerate(u8):
        d = np.abs(b.numpy().astype(np.int16) - want.astype(np.int16))
        good = d.max() <= 1 and (d > 0).mean() < 1e-3
        ok &= bool(good)
        report.append(f"{mode} camera u8 buffer {k}: max diff {int(d.max())} {'ok' if good else 'MISMATCH'}")
    dist.barrier()
    comm.close()
print(f"rank {rank}/{world}: " + "; ".join(report) + f" => {'OK' if ok else 'MISMATCH'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
