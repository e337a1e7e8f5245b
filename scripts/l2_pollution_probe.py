"""GPU probe: do copy-engine writes streaming through L2 (what inbound peer copies of the frame exchange look like to the
receiving GPU) slow the march / colour kernels down?  Renders the bench workload on one GPU with and without a
concurrent device-to-device copy of MB megabytes per frame on a side stream."""
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
rays = [synth.config_rays("C2", p).to(dev) for p in range(16)]
side = torch.cuda.Stream()
for mb in (0, 10, 72, 144):
    n = mb * (1 << 20) // 4
    src = torch.empty(max(n, 1), device=dev)
    dst = [torch.empty(max(n, 1), device=dev) for _ in range(3)]
    for i in range(5):
        f(rays[i], white_bg=True, N_samples=192, image_width=800, iteration=30001)
    torch.cuda.synchronize()
    steps = 200
    f.kernel_timing(steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if mb:
            with torch.cuda.stream(side):
                dst[i % 3].copy_(src, non_blocking=True)
        f(rays[i % 16], white_bg=True, N_samples=192, image_width=800, iteration=30001)
    e1.record()
    torch.cuda.synchronize()
    k, m_ms, c_ms = f.kernel_timing_read()
    f.kernel_timing(0)
    print(f"side copy {mb:4d} MB/frame: {e0.elapsed_time(e1) / steps:.4f} ms/frame, march {m_ms / k:.4f} ms, colour {c_ms / k:.4f} ms")
