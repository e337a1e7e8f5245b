"""GPU (B200): the UV-Mapping (NeuTex) CUDA path through the C ABI against the reference golden vectors and the CPU
oracle.  Tolerance: per-pixel max abs < 1e-3 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import cases as K
from oracle import restate_neutex as U
from helpers import load_golden, psnr

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _build(state, tex, sample_num=64):
    import ngf_b200
    from types import SimpleNamespace
    sphere = state["gauge_transform.encoder.last_linear.weight"].shape[0] == 3
    opt = SimpleNamespace(sample_num=sample_num, primitive_type="sphere" if sphere else "square", target_texture="None")
    m = ngf_b200.NeuTex(opt, device="cuda")
    m.load_state_dict(state, strict=True)
    m.set_texture(tex)
    return m


@pytest.mark.parametrize("name", [c.name for c in K.NEUTEX_CASES])
def test_neutex_matches_reference_golden(name):
    case = K.NEUTEX_BY_NAME[name]
    gold = load_golden(name)
    state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
    m = _build(state, tex, case.sample_num)
    out = m(campos.cuda(), raydir.cuda(), None if bg is None else bg.cuda(), noise=noise.cuda())
    torch.cuda.synchronize()
    e_c = np.abs(out["color"].cpu().numpy() - gold["color"]).max()
    e_t = np.abs(out["transmittance"].cpu().numpy() - gold["transmittance"]).max()
    assert m.last_valid_samples() > 0
    assert e_c < TOL, f"{name}: color max-abs {e_c:.3e}"
    assert e_t < TOL, f"{name}: transmittance max-abs {e_t:.3e}"
    g = torch.Generator().manual_seed(7)
    ref = torch.from_numpy(gold["color"])
    gt = (ref + 0.02 * torch.randn(ref.shape, generator=g)).clamp(0, 1)
    assert abs(psnr(out["color"].cpu(), gt) - psnr(ref, gt)) < 0.05


def test_neutex_valid_sample_count_is_exact_and_host_path_agrees():
    case = K.NEUTEX_BY_NAME["neutex_white"]
    state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
    m = _build(state, tex)
    out = m(campos.cuda(), raydir.cuda(), bg.cuda(), noise=noise.cuda())
    _, _, valid, _ = U.raygen(campos, raydir, 64, noise)
    assert m.last_valid_samples() == int(valid.sum())                  # integer decision chain: bit-exact
    c_h, t_h = m.render_host(campos, raydir.pin_memory(), bg, noise.pin_memory())
    assert torch.equal(c_h, out["color"].cpu()) and torch.equal(t_h, out["transmittance"].cpu())


def test_neutex_ragged_and_miss_rays():
    case = K.NEUTEX_BY_NAME["neutex_white"]
    state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
    m = _build(state, tex)
    gold = load_golden("neutex_white")
    for n in (1, 3, 255, 257):
        out = m(campos.cuda(), raydir[:, :n].cuda(), bg.cuda(), noise=noise[:, :n].cuda())
        assert np.abs(out["color"].cpu().numpy() - gold["color"][:, :n]).max() < TOL
    away = -raydir[:, :64]                                             # rays pointing away from the cube
    out = m(campos.cuda(), away.cuda(), bg.cuda(), noise=noise[:, :64].cuda())
    spec = U.NeuTexSpec(state)
    o_c, o_t = U.render(spec, campos, away, bg, noise[:, :64])
    assert torch.equal(out["transmittance"].cpu(), o_t) and (out["color"].cpu() - o_c).abs().max() < 1e-6


def test_neutex_full_frame_properties():
    """BASELINE config 4 size (600x800 rays): finite, in range, a strided subset equals the oracle."""
    from ngf_b200 import synth
    state = synth.neutex_state(0)
    campos, raydir = synth.neutex_camera(0)
    R = raydir.shape[1]
    assert R == 480000
    noise = synth.neutex_noise(R)
    bg = torch.ones(1, 3)
    m = _build(state, None)
    out = m(campos.cuda(), raydir.cuda(), bg.cuda(), noise=noise.cuda())
    color = out["color"].cpu()
    assert torch.isfinite(color).all() and float(color.min()) >= 0 and float(color.max()) <= 1
    idx = torch.arange(0, R, 469)[:1024]
    o_c, o_t = U.render(U.NeuTexSpec(state), campos, raydir[:, idx], bg, noise[:, idx])
    assert (color[:, idx] - o_c).abs().max() < TOL
    assert (out["transmittance"].cpu()[:, idx] - o_t).abs().max() < TOL


def test_neutex_cta_pair_variant_matches_golden(monkeypatch):
    """NGF_NTX_CG=2 (read when the network is packed): CTA pairs run every layer as cta_group::2 M=256 MMAs, each CTA
    streaming half of the weights.  Same tolerance as the default one-CTA kernel, and the two agree closely."""
    case = K.NEUTEX_BY_NAME["neutex_white"]
    gold = load_golden("neutex_white")
    state, tex, campos, raydir, bg, noise = K.build_neutex_inputs(case)
    args = (campos.cuda(), raydir.cuda(), bg.cuda())
    ref = _build(state, tex)(*args, noise=noise.cuda())["color"].cpu()
    monkeypatch.setenv("NGF_NTX_CG", "2")
    m = _build(state, tex)
    out = m(*args, noise=noise.cuda())
    torch.cuda.synchronize()
    assert np.abs(out["color"].cpu().numpy() - gold["color"]).max() < TOL
    assert np.abs(out["transmittance"].cpu().numpy() - gold["transmittance"]).max() < TOL
    assert float((out["color"].cpu() - ref).abs().max()) < 2e-4       # same operands, different MMA shapes
    for n in (1, 300):                                                 # one pair with an idle second CTA; two tiles
        o = m(campos.cuda(), raydir[:, :n].cuda(), bg.cuda(), noise=noise[:, :n].cuda())
        assert np.abs(o["color"].cpu().numpy() - gold["color"][:, :n]).max() < TOL
