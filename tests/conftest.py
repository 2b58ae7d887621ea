import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _ensure_native_built():
    """tests import the in-tree C-ABI library and the oracle; build them when a fresh checkout has none (the
    artefacts are git-ignored).  Same commands as __graft_entry__.build()."""
    import subprocess
    so = os.path.join(ROOT, "difflexmm_b200", "libdfx.so")
    src_dir = os.path.join(ROOT, "difflexmm_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir))
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, "include", "dfx.h")))
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-Xcompiler", "-fPIC", "-shared", "-o", so, os.path.join(src_dir, "dfx_api.cu")])
    import oracle
    oracle.build()


_ensure_native_built()
