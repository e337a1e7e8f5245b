"""CPU: libngf_b200.so builds for sm_100a, loads, and exports every symbol include/ngf_b200.h declares; argument
validation that needs no GPU returns the documented error codes.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ngf_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("ngf_field_pack", "ngf_field_render", "ngf_field_render_host", "ngf_field_density", "ngf_field_rgb",
                 "ngf_field_gauge", "ngf_field_sample_ray", "ngf_field_alpha_keep", "ngf_shard_gather",
                 "ngf_shard_scatter", "ngf_last_error", "ngf_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    import ngf_b200
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ngf_b200.h but not exported"
        assert name in ngf_b200._lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.ngf_abi_version() == 1


def test_library_is_sm100a_and_uses_tcgen05(lib):
    import ngf_b200
    out = subprocess.run(["cuobjdump", "-lelf", ngf_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3ngf17ngf_colour_kernelILi0ELi0EEEvNS_8FieldDevENS_10RenderArgsE",
                           ngf_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "colour kernel has no tcgen05 MMA in its SASS"


def test_argument_validation_without_gpu(lib):
    import ngf_b200
    L = ngf_b200._lib
    h = C.c_void_p()
    assert lib.ngf_field_pack(None, 0, C.byref(h)) == L.NGF_EINVAL
    assert b"NULL" in lib.ngf_last_error()
    d = L.NgfFieldDesc()
    d.variant = 7
    assert lib.ngf_field_pack(C.byref(d), 0, C.byref(h)) == L.NGF_EINVAL
    d.variant = L.NGF_TRIPLANE
    d.plane_c, d.density_c = 32, 8
    assert lib.ngf_field_pack(C.byref(d), 0, C.byref(h)) == L.NGF_EUNSUPPORTED
    assert lib.ngf_field_render(None, None, 0, 6, 0, 1, 0, None, None, None, 0, None) == L.NGF_EINVAL
    # round-2 entry points: argument errors are reported before any CUDA call
    assert lib.ngf_field_backward(None, None, 0, 6, 0, 1, None, None, None, None) == L.NGF_EINVAL
    assert lib.ngf_adam_step(None, None, None, None, 4, 1e-3, 0.9, 0.99, 1e-8, 1, None) == L.NGF_EINVAL
    assert lib.ngf_depth_colormap(None, 4, 2.0, 6.0, None, None) == L.NGF_EINVAL
    comm = C.c_void_p()
    assert lib.ngf_comm_init(5, 2, 0, 1000, 64, 3, L.COMM_COPY, C.byref(comm)) == L.NGF_EINVAL      # rank outside the world
    assert lib.ngf_comm_init(0, 2, 0, 1000, 64, 9, L.COMM_COPY, C.byref(comm)) == L.NGF_EINVAL      # too many frame buffers
    assert lib.ngf_comm_init(0, 2, 0, 1000, 64, 3, 7, C.byref(comm)) == L.NGF_EINVAL                # unknown exchange mode
    assert lib.ngf_comm_handle_bytes() == 128 and lib.ngf_comm_local_rays(None) == -1
    assert lib.ngf_frame_allgather(None, 1, None, None) == L.NGF_EINVAL
    t = C.c_uint64()
    assert lib.ngf_field_render_sharded(None, None, None, 0, 6, 0, 1, 0, 0, None, C.byref(t)) == L.NGF_EINVAL
    assert lib.ngf_shard_count(10, 0, 0, 1) == -1
    assert lib.ngf_shard_count(100, 8, 1, 4) == 24         # blocks 1, 5, 9
    assert lib.ngf_shard_count(100, 8, 0, 4) == 24 + 4     # blocks 0, 4, 8 + the 4-ray tail block 12


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import ngf_b200
    L = ngf_b200._lib
    monkeypatch.setattr(L, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.load(str(tmp_path / "nope.so"))


def test_cpu_field_refuses_to_render():
    import torch
    import ngf_b200
    from ngf_b200 import synth
    kw = synth.field_kwargs("C1")
    f = ngf_b200.TriPlane(kw["aabb"], kw["gridSize"], "cpu", res=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f(torch.zeros(4, 6))
