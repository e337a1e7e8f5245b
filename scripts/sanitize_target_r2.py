"""Tiny runs of every kernel family added in round 2, for compute-sanitizer (memcheck / racecheck / synccheck):
backward pass (TriPlane + InfoInv), FusedAdam, sharded render with three same-process ranks (copy and store exchange),
sharded camera batch + uint8 host path, depth colormap, NeuTex sphere; with NGF_COLOUR_TMA=1 / NGF_INFOINV_TC=1 /
NGF_INFOINV_PHASED=1 in the environment the opt-in kernels run instead of the default ones."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import cases as K
from helpers import build_cuda_field, forward_kwargs
import ngf_b200
from ngf_b200 import _lib
from ngf_b200.render import shard_index

for name in ("train_tp_hull", "train_ii_fog"):
    case = K.TRAIN_BY_NAME[name]
    state, kw, occ, rays = K.build_inputs(case)
    f = build_cuda_field(case, state, kw, occ)
    opt = ngf_b200.FusedAdam(f.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    r = rays[::5][:700].contiguous().cuda()
    out = f(r, white_bg=True, is_train=True, N_samples=case.n_samples, **forward_kwargs(case))
    loss = torch.mean((out["rgb_map"] - 0.5) ** 2)
    loss.backward()
    opt.step()
    ev = f(r, white_bg=True, N_samples=case.n_samples, **forward_kwargs(case))
    torch.cuda.synchronize()
    print(name, float(loss.detach()), float(ev["rgb_map"].mean()), f.last_stats())

case = K.CASE_BY_NAME["tp_fog_c1"]
state, kw, occ, rays = K.build_inputs(case)
f = build_cuda_field(case, state, kw, occ)
rays = rays[:1536]
n = rays.shape[0]
lib = _lib.load()
nb = int(lib.ngf_comm_handle_bytes())
fh = f._ensure_handle()
for mode in (_lib.COMM_COPY, _lib.COMM_STORE):
    hs, blobs = [], b""
    for rk in range(3):
        h = C.c_void_p()
        _lib.check(lib.ngf_comm_init(rk, 3, 0, n, 96, 3, mode, C.byref(h)))
        buf = C.create_string_buffer(nb)
        _lib.check(lib.ngf_comm_export(h, buf))
        hs.append(h); blobs += bytes(buf.raw)
    for h in hs:
        _lib.check(lib.ngf_comm_connect(h, blobs))
    mine = [rays[shard_index(n, 96, rk, 3)].cuda().contiguous() for rk in range(3)]
    t = C.c_uint64()
    for k in range(4):
        tk = []
        for rk in range(3):
            _lib.check(lib.ngf_field_render_sharded(fh, hs[rk], mine[rk].data_ptr(), mine[rk].shape[0], 6, 64, 1, 0, 0, None, C.byref(t)))
            tk.append(int(t.value))
        for rk in range(3):
            p = C.c_void_p()
            _lib.check(lib.ngf_frame_allgather(hs[rk], tk[rk], None, C.byref(p)))
            _lib.check(lib.ngf_frame_release(hs[rk], tk[rk], None))
    torch.cuda.synchronize()
    for h in hs:
        lib.ngf_comm_free(h)
    print("sharded mode", mode, "ok")
H, W = 24, 40
poses = torch.stack([K.synth.look_at_c2w(*K.synth.pose_angles(p)) for p in (4, 9)])
comm = ngf_b200.FrameComm(f, 2 * H * W, block=4 * W)
u8 = torch.zeros((H * W, 3), dtype=torch.uint8).pin_memory()
for k in range(3):
    comm.wait(comm.submit_camera_host(poses.contiguous().pin_memory(), H, W, K.synth.FOCAL_800 * 64 / 800, u8, first_row=H * W, N_samples=64,
                                      white_bg=True, iteration=30001))
comm.close()
out = f.render_camera(poses[0], H, W, K.synth.FOCAL_800 * 64 / 800, white_bg=True, N_samples=64, iteration=30001)
img, _ = ngf_b200.visualize_depth(out["depth_map"].view(H, W), [2.0, 6.0])
print("camera batch + depth colormap ok", int(u8.sum()), int(img.sum()))
case = K.NEUTEX_BY_NAME["neutex_sphere"]
state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
from types import SimpleNamespace
m = ngf_b200.NeuTex(SimpleNamespace(sample_num=64, primitive_type="sphere", target_texture="None"), device="cuda")
m.load_state_dict(state)
o = m(campos.cuda(), raydir[:, :300].cuda(), bg.cuda(), noise=noise[:, :300].cuda())
torch.cuda.synchronize()
print("neutex sphere", float(o["color"].mean()), m.last_valid_samples())
# later in round 2: seeded jitter, the fp32 NeuTex path and the self-check; renders of one field on three streams
o2 = m(campos.cuda(), raydir[:, :300].cuda(), bg.cuda(), seed=5)
nz = m.noise_for(5, 300)
m.set_precision("fp32")
o3 = m(campos.cuda(), raydir[:, :300].cuda(), bg.cuda(), noise=nz)
rep = m.self_check(512, seed=1)
torch.cuda.synchronize()
print("neutex seeded / fp32 / self-check", float((o2["color"] - o3["color"]).abs().max()), rep["rgb_max"])
case = K.CASE_BY_NAME["tp_hull_c1"]
state, kw, occ, rays = K.build_inputs(case)
f = build_cuda_field(case, state, kw, occ)
r = rays.cuda()
alone = f(r, white_bg=True, N_samples=case.n_samples, **forward_kwargs(case))["rgb_map"]
streams = [torch.cuda.Stream() for _ in range(3)]
outs = []
for st in streams:
    st.wait_stream(torch.cuda.current_stream())
for i in range(6):
    with torch.cuda.stream(streams[i % 3]):
        outs.append(f(r, white_bg=True, N_samples=case.n_samples, **forward_kwargs(case))["rgb_map"])
torch.cuda.synchronize()
print("multi-stream renders ok", max(float((o - alone).abs().max()) for o in outs))
