"""Builds difflexmm_b200/libdfx.so in-tree: every translation unit of csrc/ is compiled for sm_100a by its own nvcc
process (in parallel) and the objects are linked into one shared library.  Used by __graft_entry__.build() and by
tests/conftest.py; the library itself never falls back to anything when this has not been run (see _lib.py)."""

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(HERE, "libdfx.so")
UNITS = ["dfx_api.cu", "dfx_adjoint3.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newest_source():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cc"))]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "dfx.h"))
    return max(os.path.getmtime(p) for p in paths)


def up_to_date():
    return os.path.exists(SO) and os.path.getmtime(SO) >= _newest_source()


def have_nvcc():
    return shutil.which("nvcc") is not None


def build(force=False, extra_flags=(), output=None):
    """Compile and link.  `extra_flags` / `output` build an experimental variant of the same ABI (selected at run time
    with DFX_LIB)."""
    so = output or SO
    if not force and output is None and up_to_date():
        return so
    obj_dir = OBJ if output is None else OBJ + "_" + os.path.basename(output).replace(".", "_")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_unit(unit):
        obj = os.path.join(obj_dir, unit.replace(".cu", ".o"))
        subprocess.check_call(["nvcc", *NVCC_FLAGS, *extra_flags, "-c", "-o", obj, os.path.join(CSRC, unit)])
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
        objs = list(pool.map(compile_unit, UNITS))
    subprocess.check_call(["nvcc", "-shared", "-o", so, *objs])
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
