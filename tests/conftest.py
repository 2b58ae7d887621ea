import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device."""
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _ensure_native_built():
    """tests import the in-tree C-ABI library and the oracle; build them when a fresh checkout has none (the
    artefacts are git-ignored).  Same commands as __graft_entry__.build().  Without nvcc only the oracle is built:
    the host-side tests that need libdfx.so then fail on import with the library's own message."""
    from difflexmm_b200 import build_native
    if build_native.have_nvcc():
        build_native.build()
    import oracle
    oracle.build()


_ensure_native_built()
