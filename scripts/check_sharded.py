"""Launched under torchrun (one rank per GPU): render the golden case tp_fog_c1 with its rays sharded across the
ranks (interleaved blocks + one NCCL all-gather of the frame, ngf_b200.render.render_frame_sharded) and compare the
gathered frame with the reference golden vector on every rank."""
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

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
case = K.CASE_BY_NAME["tp_fog_c1"]
state, kw, occ, rays = K.build_inputs(case)
f = build_cuda_field(case, state, kw, occ, device=dev)
frame = ngf_b200.render_frame_sharded(rays.to(dev), f, block=96, N_samples=case.n_samples, white_bg=True, iteration=30001)
torch.cuda.synchronize()
gold = load_golden("tp_fog_c1")
e_rgb = float(np.abs(frame[:, :3].cpu().numpy() - gold["rgb"]).max())
e_dep = float(np.abs(frame[:, 3].cpu().numpy() - gold["depth"]).max())
ok = e_rgb < 1e-3 and e_dep < 2e-3
# the overlapped, double-buffered variant: three batches in a row, each must equal the same golden frame
from ngf_b200.render import shard_rays
sr = ngf_b200.ShardedFrameRenderer(f, rays.shape[0], block=96)
mine = shard_rays(rays.to(dev), 96, dist.get_rank(), dist.get_world_size())
tickets = []
for k in range(3):
    tickets.append(sr.submit(mine, N_samples=case.n_samples, white_bg=True, iteration=30001))
    fr = sr.result(tickets[-1]).clone()
    torch.cuda.synchronize()
    ok = ok and float(np.abs(fr[:, :3].cpu().numpy() - gold["rgb"]).max()) < 1e-3 and \
        float(np.abs(fr[:, 3].cpu().numpy() - gold["depth"]).max()) < 2e-3
print(f"rank {dist.get_rank()}/{dist.get_world_size()}: rgb {e_rgb:.2e} depth {e_dep:.2e} {'OK' if ok else 'MISMATCH'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
