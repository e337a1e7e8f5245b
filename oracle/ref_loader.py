"""TEST INFRASTRUCTURE — import the UNMODIFIED reference modules from ``/root/reference``.

Only usable in the build container (the GPU box has no ``/root/reference``).  Used to
(1) pin ``oracle/restate_*.py`` to the real reference and (2) generate ``tests/golden/*``.

The reference's model files import plotting packages they never use
(``TriPlane/models/FieldBase.py:7``, ``TriPlane/models/Field.py:6,8``, ``InfoInv/models/FieldBase.py:7``);
those are satisfied with empty stub modules.  TriPlane and InfoInv both call their package ``models``,
so each is imported and then re-registered under a private name to let both live in one process.
Importing the reference re-seeds torch/numpy (``FieldBase.py:9-10``); the RNG state is restored afterwards.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np
import torch

_CANDIDATES = [os.environ.get("NGF_REFERENCE", ""), "/root/reference"]
_CACHE: dict[str, types.ModuleType] = {}


def reference_root() -> str | None:
    for c in _CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "TriPlane", "models")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def _install_stubs() -> None:
    def stub(name: str, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__ngf_stub__ = True
        sys.modules[name] = m
        return m

    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = stub("matplotlib")
        mpl.pyplot = stub("matplotlib.pyplot")
    try:
        from mpl_toolkits.mplot3d import axes3d  # noqa: F401
    except Exception:
        tk = stub("mpl_toolkits")
        m3 = stub("mpl_toolkits.mplot3d")
        m3.axes3d = types.SimpleNamespace()
        tk.mplot3d = m3


def _import_models(subproject: str, top_pkg: str) -> types.ModuleType:
    """Import ``<root>/<subproject>/<top_pkg>`` and park it under ``ngfref_<subproject>``."""
    key = f"{subproject}:{top_pkg}"
    if key in _CACHE:
        return _CACHE[key]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set NGF_REFERENCE or mount /root/reference)")
    _install_stubs()
    path = os.path.join(root, subproject)
    rng_t, rng_n = torch.get_rng_state(), np.random.get_state()
    clash = {k: v for k, v in sys.modules.items() if k == top_pkg or k.startswith(top_pkg + ".")}
    for k in clash:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            pkg = importlib.import_module(top_pkg)
            for sub in _SUBMODULES[key]:
                importlib.import_module(f"{top_pkg}.{sub}")
    finally:
        sys.path.remove(path)
        mine = {k: v for k, v in sys.modules.items() if k == top_pkg or k.startswith(top_pkg + ".")}
        for k, v in mine.items():
            del sys.modules[k]
            sys.modules[f"ngfref_{subproject.replace('-', '_')}." + k] = v
        sys.modules.update(clash)
        torch.set_rng_state(rng_t)
        np.random.set_state(rng_n)
    _CACHE[key] = pkg
    return pkg


_SUBMODULES = {
    "TriPlane:models": ["FieldBase", "networks", "Field"],
    "InfoInv:models": ["FieldBase", "networks", "Field"],
    "UV-Mapping:model": ["renderer", "decoder", "gauge_fields"],
}


def triplane_models():
    """-> the reference ``TriPlane/models`` package (``.Field.TriPlane``, ``.FieldBase.AlphaGridMask`` ...)."""
    return _import_models("TriPlane", "models")


def infoinv_models():
    """-> the reference ``InfoInv/models`` package."""
    return _import_models("InfoInv", "models")


def uvmapping_model():
    """-> the reference ``UV-Mapping/model`` package (renderer, decoder, gauge_fields; ``model.model`` needs
    the repo-level ``util`` module and is imported lazily by the UV oracle)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found")
    # UV-Mapping/model/*.py do `from util import ...`-style absolute imports of UV-Mapping/util.py
    sys.path.insert(0, os.path.join(root, "UV-Mapping"))
    try:
        return _import_models("UV-Mapping", "model")
    finally:
        sys.path.remove(os.path.join(root, "UV-Mapping"))


def quiet(fn, *a, **k):
    """Call ``fn`` swallowing the reference's constructor prints (``FieldBase.py:64-65,73-74``)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def reference_renderer(rays, field, chunk=4096, N_samples=-1, white_bg=True, is_train=False, device="cpu",
                       **fwd_kw):
    """The 12-line chunk loop of ``TriPlane/main.py:60-71`` (``main.py`` itself is not importable here:
    it needs configargparse / imageio / kornia).  ``fwd_kw`` carries ``iteration=30001`` (TriPlane,
    main.py:67) or ``infoinv=...`` (InfoInv, main.py:68)."""
    rgbs, depths = [], []
    n = rays.shape[0]
    for c in range(n // chunk + int(n % chunk > 0)):
        out = field(rays[c * chunk:(c + 1) * chunk].to(device), is_train=is_train, white_bg=white_bg,
                    N_samples=N_samples, **fwd_kw)
        rgbs.append(out["rgb_map"])
        depths.append(out["depth_map"])
    return torch.cat(rgbs), torch.cat(depths)
