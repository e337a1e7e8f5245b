import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, ngf_b200
from ngf_b200 import synth
dev = torch.device("cuda", 0)
kw = synth.field_kwargs("C2")
f = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"], distance_scale=25, rayMarch_weight_thres=1e-4, gauge_start=0)
synth.load_into(f, synth.field_state("triplane", "hull"), synth.occupancy_volume("hull"), ngf_b200.AlphaGridMask)
host = [synth.config_rays("C2", p).pin_memory() for p in range(8)]
rgb = torch.empty((640000, 3)).pin_memory(); dep = torch.empty((640000,)).pin_memory()
for i in range(5): f.render_host(host[i % 8], rgb, dep, N_samples=192, image_width=800, iteration=30001)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 40
for i in range(n): f.render_host(host[i % 8], rgb, dep, N_samples=192, image_width=800, iteration=30001)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print(f"chunk={os.environ.get('NGF_HOST_CHUNK','default')}: {dt*1e3:.3f} ms/frame  {640000/dt:.3e} rays/s")
outs = [(torch.empty((640000, 3)).pin_memory(), torch.empty((640000,)).pin_memory()) for _ in range(3)]
for depth in (1, 2, 3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); tickets = []
    for i in range(n):
        r, d = outs[i % 3]
        tickets.append(f.render_host_async(host[i % 8], r, d, N_samples=192, image_width=800, iteration=30001))
        if len(tickets) > depth - 1: f.host_wait(tickets.pop(0))
    for t in tickets: f.host_wait(t)
    dt = (time.perf_counter() - t0) / n
    print(f"chunk={os.environ.get('NGF_HOST_CHUNK','default')} async depth {depth}: {dt*1e3:.3f} ms/frame  {640000/dt:.3e} rays/s")
# raw copy speeds
d = torch.empty((640000, 6), device=dev); o = torch.empty((640000, 4), device=dev); oh = torch.empty((640000, 4)).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): d.copy_(host[i % 8], non_blocking=True)
torch.cuda.synchronize(); print(f"H2D 15.4MB: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
t0 = time.perf_counter()
for i in range(20): oh.copy_(o, non_blocking=True)
torch.cuda.synchronize(); print(f"D2H 10.2MB: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
