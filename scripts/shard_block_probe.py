"""GPU probe (one GPU): how do the march / colour kernels of ONE rank depend on the shard block size?  For world = 8 the
local ray set of a rank (its interleaved blocks of an 8-frame batch) is rendered through the plain single-GPU path for
several block sizes and ranks; a full frame is the reference point."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ngf_b200
from ngf_b200 import synth

dev = torch.device("cuda", 0)
kw = synth.field_kwargs("C2")
f = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
                      distance_scale=25, rayMarch_weight_thres=1e-4, gauge_start=0)
synth.load_into(f, synth.field_state("triplane", "hull"), synth.occupancy_volume("hull"), ngf_b200.AlphaGridMask)
os.environ.setdefault("NGF_COLOUR_TMA", "0")
frames = [synth.config_rays("C2", p) for p in range(16)]
R = 640000
world = 8


def shard(pose0, block, rank):
    g = torch.arange(world * R)
    mine = g[((g // block) % world) == rank]
    out = torch.empty((mine.numel(), 6))
    fidx = mine // R
    for fr in fidx.unique().tolist():
        sel = fidx == fr
        out[sel] = frames[(pose0 + fr) % 16][mine[sel] - fr * R]
    return out.contiguous()


def run(tag, sets):
    dev_sets = [s.to(dev) for s in sets]
    for s in dev_sets[:3]:
        f(s, white_bg=True, N_samples=192, image_width=800, iteration=30001)
    torch.cuda.synchronize()
    steps = 96
    f.kernel_timing(steps)
    for i in range(steps):
        f(dev_sets[i % len(dev_sets)], white_bg=True, N_samples=192, image_width=800, iteration=30001)
    k, m_ms, c_ms = f.kernel_timing_read()
    f.kernel_timing(0)
    print(f"{tag:34s} march {m_ms / k:.4f} ms  colour {c_ms / k:.4f} ms  sum {(m_ms + c_ms) / k:.4f}", flush=True)


run("full frames (16 poses)", frames)
for rows in (4, 20, 40, 100):
    for rank in (0, 3):
        run(f"block {rows:3d} rows, rank {rank}", [shard(p, rows * 800, rank) for p in range(0, 16, 2)])
