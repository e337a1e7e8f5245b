"""pytest configuration: the ``gpu`` marker (tests that need a B200) and import paths.

``-m "not gpu"`` runs here on the CPU: oracle vs the committed golden vectors (and vs the imported reference when
``/root/reference`` exists), host-side logic, C-ABI symbol checks, world-size-2 gloo tests.
``-m gpu`` runs on the GPU box: the CUDA path, called through the C ABI, against the oracle and the golden vectors.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ngf_build", os.path.join(ROOT, "neural-gauge-fields_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    import ngf_b200
    return ngf_b200._lib.load()
