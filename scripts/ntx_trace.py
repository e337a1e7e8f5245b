import os, sys, ctypes as C
os.environ["NGF_NTX_DBG"] = str(int(os.environ.get("NGF_NTX_DBG", "4")) | 4)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, ngf_b200
from ngf_b200 import synth, _lib
m = ngf_b200.NeuTex(device="cuda"); m.load_state_dict(synth.neutex_state(0))
campos, raydir = synth.neutex_camera(0); R = raydir.shape[1]
noise = synth.neutex_noise(R).cuda(); campos, raydir, bg = campos.cuda(), raydir.cuda(), torch.ones(1, 3).cuda()
for _ in range(2): m(campos, raydir, bg, noise=noise)
tt = np.zeros(25 * 4 + 64, dtype=np.int64)
_lib.check(_lib.load().ngf_neutex_debug_trace(m._ensure_handle(), tt.ctypes.data))
t = tt[:100].reshape(25, 4); r = tt[100:]
t0 = t[0, 0]
print("layer  a_ready  issued(+)  acc_seen(+)  epi_done(+)   | next a_ready - epi_done")
for l in range(25):
    nxt = t[l + 1, 0] - t[l, 3] if l < 24 else 0
    print(f"{l:3d}  {t[l,0]-t0:8d}  {t[l,1]-t[l,0]:8d}  {t[l,2]-t[l,0]:8d}  {t[l,3]-t[l,2]:8d}   | {nxt:6d}")
print("total cycles for the tile:", t[24, 3] - t0)
wide = [l for l in range(25) if (1 <= l <= 10) or (16 <= l <= 20) or (21 <= l <= 24)]
print("256-wide layers: mean a_ready->acc_seen", np.mean([t[l, 2] - t[l, 0] for l in wide]), "cycles; mean epilogue",
      np.mean([t[l, 3] - t[l, 2] for l in wide]), "cycles; CG", os.environ.get("NGF_NTX_CG", "1"), "DBG", os.environ["NGF_NTX_DBG"])
print("layer 5 ring: stage  full_seen  issued   | producer: empty_seen (all relative to the layer's a_ready)")
for i in range(12):
    if r[2 * i] or r[32 + i]:
        print(f"  {i:2d}  {r[2*i]-t[5,0] if r[2*i] else 0:8d} {r[2*i+1]-t[5,0] if r[2*i+1] else 0:8d}   | {r[32+i]-t[5,0] if r[32+i] else 0:8d}")
