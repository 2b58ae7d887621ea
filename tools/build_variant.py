#!/usr/bin/env python
"""Builds an experimental variant of libdfx.so: only csrc/dfx_adjoint3.cu is recompiled (bench instance only, with the
given -D flags), the other objects are reused from the last full build.
  python tools/build_variant.py name [-DFLAG ...]   ->  gpurun_variants/libdfx_<name>.so   (select with DFX_LIB)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from difflexmm_b200 import build_native as bn  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "gpurun_variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"a3_{name}.o")
src = os.environ.get("A3_SRC", os.path.join(bn.CSRC, "dfx_adjoint3.cu"))
subprocess.check_call(["nvcc", *bn.NVCC_FLAGS, "-DDFX_A3_MAIN_ONLY", "-I", bn.CSRC, *flags, "-Xptxas", "-v", "-c", "-o", obj, src])
so = os.path.join(out_dir, f"libdfx_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-o", so, os.path.join(bn.OBJ, "dfx_api.o"), obj])
print(so)
