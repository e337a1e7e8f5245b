import os, sys, ctypes as C
os.environ["NGF_NTX_DBG"] = "4"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, ngf_b200
from ngf_b200 import synth, _lib
m = ngf_b200.NeuTex(device="cuda"); m.load_state_dict(synth.neutex_state(0))
campos, raydir = synth.neutex_camera(0); R = raydir.shape[1]
noise = synth.neutex_noise(R).cuda(); campos, raydir, bg = campos.cuda(), raydir.cuda(), torch.ones(1, 3).cuda()
for _ in range(2): m(campos, raydir, bg, noise=noise)
t = np.zeros((25, 4), dtype=np.int64)
_lib.check(_lib.load().ngf_neutex_debug_trace(m._ensure_handle(), t.ctypes.data))
t0 = t[0, 0]
print("layer  a_ready  issued(+)  acc_seen(+)  epi_done(+)   | next a_ready - epi_done")
for l in range(25):
    nxt = t[l + 1, 0] - t[l, 3] if l < 24 else 0
    print(f"{l:3d}  {t[l,0]-t0:8d}  {t[l,1]-t[l,0]:8d}  {t[l,2]-t[l,0]:8d}  {t[l,3]-t[l,2]:8d}   | {nxt:6d}")
print("total cycles for the tile:", t[24, 3] - t0)
