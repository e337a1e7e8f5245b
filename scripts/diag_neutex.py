import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases as K, restate_neutex as U
import ngf_b200
name = sys.argv[1] if len(sys.argv) > 1 else "neutex_black_gain"
case = K.NEUTEX_BY_NAME[name]
state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
m = ngf_b200.NeuTex(device="cuda"); m.load_state_dict(state); m.set_texture(tex)
out = m(campos.cuda(), raydir.cuda(), None if bg is None else bg.cuda(), noise=noise.cuda())
R = raydir.shape[1]
sr, valid = m.last_samples(R)
spec = U.NeuTexSpec(state, texture=tex)
pos, seg, v_o, _ = U.raygen(campos, raydir, 64, noise)
print("valid eq", torch.equal(valid, v_o[0].bool()))
dens = U.geometry(spec, pos)[0]; uv = U.gauge(spec, pos); rad = U.texture(spec, uv, raydir[:, :, None, :])[0][..., :3]
vm = valid
print("sigma err max", (sr[..., 0] - dens)[vm].abs().max().item(), "rel", ((sr[..., 0] - dens)[vm].abs() / dens[vm].abs().clamp(min=1e-3)).max().item(), "sigma range", dens[vm].min().item(), dens[vm].max().item())
print("rad err max", (sr[..., 1:] - rad)[vm].abs().max().item(), "rad range", rad[vm].min().item(), rad[vm].max().item())
col, bt, w = U.march(dens[None], rad[None], seg, v_o)
col_g, bt_g, _ = U.march((sr[..., 0] * vm)[None], (sr[..., 1:] * vm[..., None])[None], seg, v_o)
print("pre-tonemap color err (oracle march on GPU samples)", (col - col_g).abs().max().item(), "T err", (bt - bt_g).abs().max().item())
o_c, o_t = U.render(spec, campos, raydir, bg, noise)
e = (out["color"].cpu() - o_c).abs()
i = int(e.max(-1).values.argmax())
print("final err", e.max().item(), "at ray", i, "oracle", o_c[0, i].tolist(), "gpu", out["color"][0, i].tolist(), "pre", col[0, i].tolist())
print("T final err", (out["transmittance"].cpu() - o_t).abs().max().item())
