"""GPU probe: end-to-end rays/s of ngf_field_render_host_async (pinned rays in, rgb + depth out, 3 frames in flight) as a
function of the host-path chunk size (NGF_HOST_CHUNK_ASYNC is read once per process: one process per value)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch, ngf_b200
from ngf_b200 import synth
dev = torch.device("cuda", 0)
kw = synth.field_kwargs("C2")
f = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"], distance_scale=25,
                      rayMarch_weight_thres=1e-4, gauge_start=0)
synth.load_into(f, synth.field_state("triplane", "hull"), synth.occupancy_volume("hull"), ngf_b200.AlphaGridMask)
host = [synth.config_rays("C2", p).pin_memory() for p in range(16)]
outs = [(torch.empty((640000, 3)).pin_memory(), torch.empty((640000,)).pin_memory()) for _ in range(3)]
pend = []
def step(i):
    r, d = outs[i %% 3]
    pend.append(f.render_host_async(host[i %% 16], r, d, white_bg=True, N_samples=192, image_width=800, iteration=30001))
    if len(pend) > 2: f.host_wait(pend.pop(0))
for i in range(6): step(i)
while pend: f.host_wait(pend.pop(0))
t0 = time.perf_counter()
for i in range(400): step(i)
while pend: f.host_wait(pend.pop(0))
dt = time.perf_counter() - t0
print("chunk %%s: %%.4f ms/frame, %%.3e rays/s" %% (os.environ.get("NGF_HOST_CHUNK_ASYNC"), dt / 400 * 1e3, 640000 * 400 / dt))
''' % ROOT
for chunk in sys.argv[1:] or ["80000", "160000", "214400", "320000", "640000"]:
    env = dict(os.environ, NGF_HOST_CHUNK_ASYNC=chunk)
    print(subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
