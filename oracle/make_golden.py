"""TEST INFRASTRUCTURE — generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (imported from
``/root/reference`` through ``oracle/ref_loader.py``) on the seeded cases of ``oracle/cases.py``.

Run in the build container only:   python -m oracle.make_golden [case ...]

The reference ships no tests or golden vectors, so these files are the parity pin that travels
to the GPU box: each holds the reference's ``rgb_map`` / ``depth_map`` for one case, the per-case sample counts, a
SHA-256 of the regenerated inputs (so a drifted generator is detected instead of silently compared), and the torch
version that produced them.  ``pointwise_*.npz`` hold the reference's ``compute_gauge`` / ``compute_density`` /
``compute_rgb`` / ``sample_alpha`` / ``sample_ray`` outputs on seeded points.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import cases as K
from . import ref_loader


def build_reference_field(case: K.Case):
    """Instantiate the reference's model class for ``case`` on the CPU and load the synthetic state into it."""
    state, kw, occ, rays = K.build_inputs(case)
    if case.variant == "triplane":
        pkg = ref_loader.triplane_models()
        extra = dict(gauge_start=0)
    else:
        pkg = ref_loader.infoinv_models()
        extra = {}
    field = ref_loader.quiet(pkg.Field.TriPlane, kw["aabb"], kw["gridSize"], "cpu", near_far=kw["near_far"],
                             step_ratio=kw["step_ratio"], distance_scale=kw["distance_scale"],
                             rayMarch_weight_thres=kw["rayMarch_weight_thres"], alphaMask_thres=kw["alphaMask_thres"],
                             **extra)
    K.synth.load_into(field, state)
    if occ is not None:
        field.alphaMask = pkg.FieldBase.AlphaGridMask("cpu", K.mask_aabb(), occ)
    return field, state, kw, occ, rays


def forward_kwargs(case: K.Case) -> dict:
    if case.variant == "triplane":
        return {"iteration": 30001 if case.gauge_on else -1}      # Field.py:58 with gauge_start=0
    return {"infoinv": case.infoinv}


@torch.no_grad()
def make_case(case: K.Case) -> str:
    field, state, kw, occ, rays = build_reference_field(case)
    rgb, depth = ref_loader.reference_renderer(rays, field, chunk=4096, N_samples=case.n_samples,
                                               white_bg=case.white_bg, **forward_kwargs(case))
    path = K.golden_path(case.name)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, rgb=rgb.numpy().astype(np.float32), depth=depth.numpy().astype(np.float32),
                        n_samples=np.int64(field.nSamples), step_size=np.float32(field.stepSize.item()),
                        fingerprint=K.fingerprint(state, rays, occ), torch_version=torch.__version__)
    return path


@torch.no_grad()
def make_train_case(case: K.Case) -> str:
    """The reference's forward(is_train=True) on one chunk.  Its random draws (FieldBase.py:130, :299) are recorded by
    drawing the same numbers from the same seed first: u = rand_like([R,1]), then the background coin rand(1)."""
    field, state, kw, occ, rays = build_reference_field(case)
    assert rays.shape[0] <= 4096
    torch.manual_seed(K.TRAIN_SEED)
    u = torch.rand_like(torch.empty((rays.shape[0], 1), dtype=torch.float32))
    coin = float(torch.rand((1,)))
    torch.manual_seed(K.TRAIN_SEED)
    out = field(rays, is_train=True, white_bg=case.white_bg, N_samples=case.n_samples, **forward_kwargs(case))
    path = K.golden_path(case.name)
    np.savez_compressed(path, rgb=out["rgb_map"].numpy().astype(np.float32),
                        depth=out["depth_map"].numpy().astype(np.float32), jitter=u.numpy().astype(np.float32),
                        coin=np.float32(coin), white_used=np.bool_(case.white_bg or coin < 0.5),
                        seed=np.int64(K.TRAIN_SEED), fingerprint=K.fingerprint(state, rays, occ),
                        torch_version=torch.__version__)
    return path


def make_grad_case(name: str) -> str:
    """The reference's own training step on one chunk (main.py:272-283): forward(is_train=True), loss = mean squared
    error against rgb_train, loss.backward().  Small gradients are stored whole; of the plane and gauge-plane gradients
    (sparse, up to 4 M entries) the sum, the absolute sum, the number of non-zeros and GRAD_SAMPLES of the non-zero
    entries."""
    case = K.TRAIN_BY_NAME[K.GRAD_CASES[name]]
    field, state, kw, occ, rays = build_reference_field(case)
    target = K.grad_target(rays.shape[0])
    torch.manual_seed(K.TRAIN_SEED)
    u = torch.rand_like(torch.empty((rays.shape[0], 1), dtype=torch.float32))
    torch.manual_seed(K.TRAIN_SEED)
    field.zero_grad()
    out = field(rays, is_train=True, white_bg=True, N_samples=case.n_samples, **forward_kwargs(case))
    loss = torch.mean((out["rgb_map"] - target) ** 2)
    loss.backward()
    rec = dict(loss=np.float32(loss.item()), jitter=u.numpy(), fingerprint=K.fingerprint(state, rays, occ),
               torch_version=torch.__version__)
    g = torch.Generator().manual_seed(7)
    for k, p in field.named_parameters():
        gr = p.grad.detach().reshape(-1)
        key = k.replace(".", "__")
        if gr.numel() <= 65536:
            rec["full__" + key] = gr.numpy()
        else:
            nz = torch.nonzero(gr).reshape(-1)
            pick = nz[torch.randperm(nz.numel(), generator=g)[:K.GRAD_SAMPLES]].sort().values
            rec["idx__" + key] = pick.numpy().astype(np.int64)
            rec["val__" + key] = gr[pick].numpy()
            rec["sum__" + key] = np.float64(gr.double().sum().item())
            rec["abs__" + key] = np.float64(gr.double().abs().sum().item())
            rec["nnz__" + key] = np.int64(nz.numel())
    path = K.golden_path(name)
    np.savez_compressed(path, **rec)
    return path


@torch.no_grad()
def make_pointwise(variant: str) -> str:
    case = K.Case(f"pointwise_{variant}", variant=variant, kind="rand")
    field, state, kw, occ, rays = build_reference_field(case)
    xyz, dirs, world = K.pointwise_inputs()
    out = {}
    if variant == "triplane":
        xy, yz, xz = field.compute_gauge(xyz.clone(), iteration=1)
        out["sigma"] = field.compute_density(xy, yz, xz).numpy()
        out["rgb"] = field.compute_rgb(xy, yz, xz, dirs).numpy()
        xy0, yz0, xz0 = field.compute_gauge(xyz.clone(), iteration=-1)
        out.update(xy0=xy0.numpy(), yz0=yz0.numpy(), xz0=xz0.numpy())
    else:
        xy, yz, xz = field.transform(xyz.clone())
        out["sigma"] = field.compute_density(xy, yz, xz, infoinv=True).numpy()
        out["rgb"] = field.compute_rgb(xy, yz, xz, dirs, infoinv=True).numpy()
        out["sigma_noinv"] = field.compute_density(xy, yz, xz, infoinv=False).numpy()
        out["rgb_noinv"] = field.compute_rgb(xy, yz, xz, dirs, infoinv=False).numpy()
    out.update(xy=xy.numpy(), yz=yz.numpy(), xz=xz.numpy())
    out["alpha_keep"] = (field.alphaMask.sample_alpha(world) > 0).numpy()
    out["alpha"] = field.compute_alpha(world, field.stepSize).numpy()          # FieldBase.py:140-159
    pts, t, inside = field.sample_ray(rays[:64, :3], rays[:64, 3:6], is_train=False, N_samples=48)
    out.update(march_pts=pts.numpy(), march_t=t.numpy(), march_inside=inside.numpy())
    path = K.golden_path(case.name)
    np.savez_compressed(path, fingerprint=K.fingerprint(state, rays, occ), torch_version=torch.__version__,
                        **{k: np.asarray(v) for k, v in out.items()})
    return path


@torch.no_grad()
def make_alphamask() -> str:
    """Reference occupancy maintenance on a 48^3 lattice: updateAlphaMask (FieldBase.py:179-215), then filtering_rays
    (FieldBase.py:217-246) with and without the mask."""
    case = K.Case("alphamask_triplane", kind="hull", mask=False)
    field, state, kw, occ, rays = build_reference_field(case)
    new_aabb = ref_loader.quiet(field.updateAlphaMask, (48, 48, 48))
    vol = field.alphaMask.alpha_volume[0, 0]
    rgbs = torch.zeros(rays.shape[0], 3)
    kept_mask = ref_loader.quiet(field.filtering_rays, rays, rgbs, N_samples=64)[0]
    kept_bbox = ref_loader.quiet(field.filtering_rays, rays, rgbs, bbox_only=True)[0]
    path = K.golden_path(case.name)
    np.savez_compressed(path, volume_bits=np.packbits(vol.numpy() > 0), volume_shape=np.array(vol.shape),
                        new_aabb=new_aabb.numpy(), n_kept_mask=np.int64(kept_mask.shape[0]),
                        n_kept_bbox=np.int64(kept_bbox.shape[0]), kept_mask_first=kept_mask[:16].numpy(),
                        fingerprint=K.fingerprint(state, rays, None), torch_version=torch.__version__)
    return path


def build_reference_neutex(case: K.NeutexCase):
    """The reference's UV-Mapping sub-modules (decoder.py, gauge_fields.py) with the synthetic state loaded.
    NeuTex.forward itself cannot run as shipped, so the modules are wired exactly as
    model/model.py:30-50 does, minus the loss-only inverse-gauge lines."""
    pkg = ref_loader.uvmapping_model()
    state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
    geo = ref_loader.quiet(pkg.decoder.GeometryMlpDecoder, pos_freqs=10, hidden_size=256, num_layers=10)
    gt = pkg.gauge_fields.GaugeTransform(case.primitive)
    texnet = ref_loader.quiet(pkg.decoder.TextureMlpDecoder, 3, 10, 6, uv_dim=2 if case.primitive == "square" else 3,
                              layers=[5, 3], width=256, clamp=False, primitive_type=case.primitive, target_texture="None")
    for mod, prefix in ((geo, "net_geometry_decoder"), (gt, "gauge_transform"), (texnet, "net_texture")):
        mod.load_state_dict({k[len(prefix) + 1:]: v for k, v in state.items() if k.startswith(prefix + ".")}, strict=True)
    if tex is not None:
        texnet.cubemap_ = tex.clone()                       # what __init__ does with load_square(target_texture)
    return pkg, geo, gt, texnet, (state, tex, campos, raydir, bg, noise)


@torch.no_grad()
def make_neutex(case: K.NeutexCase) -> str:
    pkg, geo, gt, texnet, (state, tex, campos, raydir, bg, noise) = build_reference_neutex(case)
    ren = pkg.renderer
    cols, trs, dens, uvs = [], [], [], []
    for s in range(0, raydir.shape[1], 1024):               # test.py:108-114 chunking
        rd, nz = raydir[:, s:s + 1024], noise[:, s:s + 1024]
        # cube_ray_generation draws its jitter with torch.rand((N, R, S)); feed it the case's numbers
        real_rand = torch.rand
        torch.rand = lambda *a, **k: nz.clone()
        try:
            pos, seg, valid, _ = ren.cube_ray_generation(campos, rd, case.sample_num, jitter=0.05)
        finally:
            torch.rand = real_rand
        density = geo(pos)["density"][..., None]
        uv = gt(pos)
        feat = texnet(uv, rd[:, :, None, :])
        bsdf = torch.cat([density, feat[..., :3]], dim=-1)
        out = ren.ray_march(rd, pos, seg, valid, bsdf, None, None, ren.radiance_render, ren.alpha_blend)
        color, bg_w = out[0], out[6]
        if bg is not None:
            color = color + bg[:, None, :] * bg_w[:, :, None]
        cols.append(ren.simple_tone_map(color))
        trs.append(bg_w)
        dens.append(density[..., 0])
        uvs.append(uv)
    path = K.golden_path(case.name)
    np.savez_compressed(path, color=torch.cat(cols, 1).numpy(), transmittance=torch.cat(trs, 1).numpy(),
                        density_head=torch.cat(dens, 1)[0, :8].numpy(), uv_head=torch.cat(uvs, 1)[0, :8].numpy(),
                        fingerprint=K.fingerprint(state, torch.cat([raydir[0], noise[0]], 1), tex),
                        torch_version=torch.__version__)
    return path


def main(argv):
    if not ref_loader.available():
        raise SystemExit("reference tree not found: golden vectors can only be generated in the build container")
    names = argv or [c.name for c in K.CASES] + ["pointwise_triplane", "pointwise_infoinv", "alphamask_triplane"] + \
        [c.name for c in K.NEUTEX_CASES] + [c.name for c in K.TRAIN_CASES] + list(K.GRAD_CASES)
    for n in names:
        if n in K.GRAD_CASES:
            p = make_grad_case(n)
        elif n in K.TRAIN_BY_NAME:
            p = make_train_case(K.TRAIN_BY_NAME[n])
        elif n == "alphamask_triplane":
            p = make_alphamask()
        elif n in K.NEUTEX_BY_NAME:
            p = make_neutex(K.NEUTEX_BY_NAME[n])
        elif n.startswith("pointwise_"):
            p = make_pointwise(n.split("_", 1)[1])
        else:
            p = make_case(K.CASE_BY_NAME[n])
        print(f"{n}: {os.path.getsize(p) / 1024:.1f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
