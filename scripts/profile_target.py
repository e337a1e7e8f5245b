"""Small ncu target: N whole-frame renders of BASELINE configs[1] in the sparse ("hull") or dense regime.
Usage: python scripts/profile_target.py hull|dense|infoinv [n_renders]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ngf_b200
from ngf_b200 import synth

regime = sys.argv[1] if len(sys.argv) > 1 else "hull"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
if regime == "neutex":
    m = ngf_b200.NeuTex(device=dev)
    m.load_state_dict(synth.neutex_state(0))
    campos, raydir = synth.neutex_camera(0)
    R = raydir.shape[1]
    noise = synth.neutex_noise(R).to(dev)
    campos, raydir, bg = campos.to(dev), raydir.to(dev), torch.ones(1, 3, device=dev)
    m(campos, raydir, bg, noise=noise)
    m.kernel_timing(n)
    for i in range(n):
        m(campos, raydir, bg, noise=noise)
    k, a, b, c = m.kernel_timing_read()
    nv = m.last_valid_samples()
    flops = nv * 2 * (63 * 256 + 10 * 256 * 256 + 256 + 63 * 64 + 64 * 128 + 2 * 128 * 128 + 256 + 42 * 256 + 5 * 256 * 256 + 768
                      + 295 * 256 + 3 * 256 * 256 + 768)
    print(f"neutex: {R} rays, {nv} in-cube samples ({nv / R:.1f}/ray); raygen {a / k:.3f} ms, mlp {b / k:.3f} ms, march {c / k:.3f} ms;"
          f" {R / ((a + b + c) / k * 1e-3):.3e} rays/s; mlp {flops / (b / k * 1e-3) / 1e12:.1f} TFLOP/s (reference FLOPs of evaluated samples)")
    sys.exit(0)
kw = synth.field_kwargs("C2")
if regime == "infoinv":
    f = ngf_b200.InfoInvTriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
                                 distance_scale=25, rayMarch_weight_thres=1e-4)
    synth.load_into(f, synth.field_state("infoinv", "hull"), synth.occupancy_volume("hull"), ngf_b200.AlphaGridMask)
    fwd = dict(infoinv=True)
else:
    thres = -1.0 if regime == "dense" else 1e-4
    f = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], dev, near_far=kw["near_far"], step_ratio=kw["step_ratio"],
                          distance_scale=25, rayMarch_weight_thres=thres, gauge_start=0)
    st = synth.field_state("triplane", "fog" if regime == "dense" else "hull")
    if regime == "dense":
        st["density_decoder.bias"] = st["density_decoder.bias"] - 12.0
        synth.load_into(f, st)
    else:
        synth.load_into(f, st, synth.occupancy_volume("hull"), ngf_b200.AlphaGridMask)
    fwd = dict(iteration=30001)
# the bench's workload: 16 poses in rotation (246 MB of rays > L2), so a captured launch reads its rays from HBM
rays = [synth.config_rays("C2", p).to(dev) for p in range(16)]
for i in range(n):
    f(rays[i % 16], white_bg=True, N_samples=192, image_width=800, **fwd)
torch.cuda.synchronize()
print(regime, f.last_stats())
