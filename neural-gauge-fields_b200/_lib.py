"""ctypes binding of ``libngf_b200.so`` (C ABI declared in ``include/ngf_b200.h``).

There is no CPU implementation behind these calls: if the shared library is missing or a call fails, a
``RuntimeError`` carrying ``ngf_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libngf_b200.so")

NGF_OK, NGF_EINVAL, NGF_ECUDA, NGF_EUNSUPPORTED, NGF_ENOMEM, NGF_ECOMM = 0, -1, -2, -3, -4, -5
COMM_COPY, COMM_STORE = 0, 1
NGF_TRIPLANE, NGF_INFOINV = 0, 1
MLP_TCGEN05, MLP_SIMT = 0, 1

c_float_p = C.POINTER(C.c_float)


class NgfLinear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("in_dim", C.c_int32), ("out_dim", C.c_int32)]


class NgfFieldDesc(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("plane", C.c_void_p * 3), ("plane_h", C.c_int32 * 3), ("plane_w", C.c_int32 * 3),
        ("plane_c", C.c_int32), ("density_c", C.c_int32),
        ("gauge", C.c_void_p * 3), ("gauge_h", C.c_int32 * 3), ("gauge_w", C.c_int32 * 3), ("gauge_on", C.c_int32),
        ("rgb_basis", NgfLinear), ("rgb_l1", NgfLinear), ("rgb_l2", NgfLinear), ("rgb_l3", NgfLinear),
        ("view_pe", C.c_int32),
        ("dens_l1", NgfLinear), ("dens_l2", NgfLinear), ("dens_l3", NgfLinear),
        ("density_shift", C.c_float), ("infoinv", C.c_int32),
        ("aabb", C.c_float * 6), ("inv_aabb_size", C.c_float * 3), ("step_size", C.c_float),
        ("n_samples", C.c_int32), ("near_t", C.c_float), ("far_t", C.c_float),
        ("distance_scale", C.c_float), ("weight_thres", C.c_float),
        ("alpha_volume", C.c_void_p), ("alpha_dims", C.c_int32 * 3), ("alpha_aabb", C.c_float * 6),
        ("alpha_inv", C.c_float * 3),
    ]


class NgfFieldGrads(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("gauge", C.c_void_p * 3), ("plane_param", C.c_void_p * 3),
                ("rgb_basis", C.c_void_p), ("rgb_l1_w", C.c_void_p), ("rgb_l1_b", C.c_void_p), ("rgb_l2_w", C.c_void_p),
                ("rgb_l2_b", C.c_void_p), ("rgb_l3_w", C.c_void_p), ("rgb_l3_b", C.c_void_p),
                ("dens_l1_w", C.c_void_p), ("dens_l1_b", C.c_void_p), ("dens_l2_w", C.c_void_p), ("dens_l2_b", C.c_void_p),
                ("dens_l3_w", C.c_void_p), ("dens_l3_b", C.c_void_p)]


class NgfNeutexDesc(C.Structure):
    _fields_ = [("geometry", NgfLinear * 12), ("gauge", NgfLinear * 5), ("tex_block1", NgfLinear * 6),
                ("tex_color1", NgfLinear), ("tex_block2", NgfLinear * 5), ("sample_num", C.c_int32),
                ("jitter", C.c_float), ("texture", C.c_void_p), ("tex_h", C.c_int32), ("tex_w", C.c_int32),
                ("tex_c", C.c_int32), ("primitive", C.c_int32)]


class NgfCamera(C.Structure):
    _fields_ = [("c2w", C.c_float * 12), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32)]


class NgfStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples_in_box", C.c_uint64), ("samples_density", C.c_uint64),
                ("samples_colour", C.c_uint64), ("mlp_tiles", C.c_uint64), ("direct_patches", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/ngf_b200.h declares
SIGNATURES = {
    "ngf_abi_version": (C.c_int, []),
    "ngf_last_error": (C.c_char_p, []),
    "ngf_launch_count": (C.c_uint64, []),
    "ngf_field_pack": (C.c_int, [C.POINTER(NgfFieldDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "ngf_field_repack": (C.c_int, [C.c_void_p, C.POINTER(NgfFieldDesc)]),
    "ngf_field_free": (None, [C.c_void_p]),
    "ngf_field_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ngf_field_render_jitter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ngf_field_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.POINTER(NgfFieldGrads), C.c_void_p]),
    "ngf_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_int64, C.c_void_p]),
    "ngf_field_render_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "ngf_field_render_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]),
    "ngf_field_render_camera": (C.c_int, [C.c_void_p, C.POINTER(NgfCamera), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int32, C.c_void_p]),
    "ngf_field_render_camera_host_async": (C.c_int, [C.c_void_p, C.POINTER(NgfCamera), C.c_int32, C.c_int32, C.c_void_p,
                                                     C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]),
    "ngf_field_render_camera_u8_host_async": (C.c_int, [C.c_void_p, C.POINTER(NgfCamera), C.c_int32, C.c_int32, C.c_void_p,
                                                        C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]),
    "ngf_field_host_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "ngf_field_set_gauge": (C.c_int, [C.c_void_p, C.c_int32]),
    "ngf_field_set_infoinv": (C.c_int, [C.c_void_p, C.c_int32]),
    "ngf_field_stats": (C.c_int, [C.c_void_p, C.POINTER(NgfStats), C.c_void_p]),
    "ngf_field_timing_begin": (C.c_int, [C.c_void_p, C.c_int32]),
    "ngf_field_timing_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]),
    "ngf_field_sample_ray": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngf_field_sample_ray_jitter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngf_field_alpha_keep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ngf_field_alpha_value": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ngf_field_gauge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "ngf_field_density": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]),
    "ngf_field_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_int32, C.c_void_p]),
    "ngf_field_sigma_world": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "ngf_neutex_pack": (C.c_int, [C.POINTER(NgfNeutexDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "ngf_neutex_free": (None, [C.c_void_p]),
    "ngf_neutex_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "ngf_neutex_render_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p]),
    "ngf_neutex_last_valid_samples": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "ngf_neutex_render_seeded": (C.c_int, [C.c_void_p] * 4 + [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngf_neutex_render_host_seeded": (C.c_int, [C.c_void_p] * 4 + [C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ngf_neutex_noise": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ngf_neutex_set_precision": (C.c_int, [C.c_void_p, C.c_int32]),
    "ngf_neutex_self_check": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_float), C.c_void_p]),
    "ngf_neutex_copy_samples": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "ngf_neutex_debug_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ngf_neutex_timing_begin": (C.c_int, [C.c_void_p, C.c_int32]),
    "ngf_neutex_timing_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]),
    "ngf_frame_post": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngf_depth_colormap": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "ngf_shard_count": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "ngf_shard_gather": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p]),
    "ngf_shard_scatter": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                    C.c_void_p]),
    "ngf_comm_init": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                C.POINTER(C.c_void_p)]),
    "ngf_comm_handle_bytes": (C.c_int64, []),
    "ngf_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ngf_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ngf_comm_free": (None, [C.c_void_p]),
    "ngf_comm_local_rays": (C.c_int64, [C.c_void_p]),
    "ngf_field_render_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_uint64)]),
    "ngf_frame_allgather": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ngf_frame_release": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "ngf_field_render_sharded_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                                      C.POINTER(C.c_uint64)]),
    "ngf_comm_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "ngf_field_render_sharded_camera": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(NgfCamera), C.c_void_p, C.c_int32, C.c_int32,
                                                  C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_uint64)]),
    "ngf_field_render_sharded_camera_u8_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(NgfCamera), C.c_void_p,
                                                                C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                                C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]),
}

_lib = None


def load(path: str | None = None):
    """Load (once) and return the ctypes handle of libngf_b200.so.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("NGF_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `python neural-gauge-fields_b200/build.py`). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ngf_abi_version() != 1:
        raise RuntimeError(f"libngf_b200.so ABI version {lib.ngf_abi_version()} != 1")
    _lib = lib
    return lib


def last_error() -> str:
    return load().ngf_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != NGF_OK:
        name = {NGF_EINVAL: "NGF_EINVAL", NGF_ECUDA: "NGF_ECUDA", NGF_EUNSUPPORTED: "NGF_EUNSUPPORTED",
                NGF_ENOMEM: "NGF_ENOMEM", NGF_ECOMM: "NGF_ECOMM"}.get(rc, str(rc))
        raise RuntimeError(f"{what or 'libngf_b200'} failed with {name}: {last_error()}")


def launch_count() -> int:
    return int(load().ngf_launch_count())
