"""The C-ABI library loads on a CPU-only box and exports every entry point that include/dfx.h declares
(no compute calls here: those need a GPU)."""

import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "dfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = _declared_functions()
    for required in ("dfx_topology_create", "dfx_topology_destroy", "dfx_forward", "dfx_adjoint", "dfx_expand_fields",
                     "dfx_forward_workspace_bytes", "dfx_adjoint_workspace_bytes", "dfx_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from difflexmm_b200 import _lib  # raises if libdfx.so is missing: there is no fallback
    for name in _declared_functions():
        assert hasattr(_lib.lib, name), f"libdfx.so does not export {name}"
    assert b"sm_100a" in _lib.lib.dfx_version()
    assert [_lib.lib.dfx_drive_n_params(k) for k in range(5)] == [0, 3, 3, 2, 5]


def test_ctypes_struct_layout_matches_header_sizes():
    from difflexmm_b200 import _abi
    assert C.sizeof(_abi.DfxStats) == 40
    assert C.sizeof(_abi.DfxLeaf) == 16
    assert C.sizeof(_abi.DfxOptions) == 24
    assert C.sizeof(_abi.DfxParamGrads) == 10 * 8
    # DfxParams: 5 leaves, int[3] (+pad), leaf, int (+pad), 3 leaves
    assert C.sizeof(_abi.DfxParams) == 5 * 16 + 16 + 16 + 8 + 4 * 16


def test_topology_create_reports_errors_instead_of_crashing():
    import numpy as np
    import torch
    from difflexmm_b200 import _abi, _lib
    spec = _abi.TopologySpec(2, 4, [[0, 6]], [])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="dfx_topology_create failed"):
            _lib.Topology(spec, 0)
    # a vertex shared by two bonds is rejected by the slot-based assembly, with a message
    desc = _abi.TopologySpec(2, 4, [[0, 6], [0, 7]], []).to_desc()
    h = C.c_void_p()
    rc = _lib.lib.dfx_topology_create(C.byref(desc), 0, C.byref(h))
    assert rc != 0 and b"more than one bond" in _lib.lib.dfx_last_error()


def test_header_is_plain_c_and_struct_sizes_match_ctypes(tmp_path):
    """include/dfx.h must compile as C99 (the boundary is a C ABI, no C++ types) and the ctypes mirrors in
    difflexmm_b200/_abi.py must have the sizes the C compiler gives the structs"""
    import subprocess
    from difflexmm_b200 import _abi
    names = ["DfxTopologyDesc", "DfxLeaf", "DfxParams", "DfxParamGrads", "DfxOptions", "DfxStats", "DfxObjective", "DfxGeometryDesc", "DfxConstraintDesc"]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "dfx.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert int(sizes[n]) == C.sizeof(getattr(_abi, n)), n


def test_jax_ffi_shim_type_checks_against_the_ffi_surface():
    """difflexmm_b200/csrc/jax_ffi_shim.cc (the XLA FFI handlers of the C ABI) cannot be built here -- no jaxlib headers --
    but it must at least be valid C++ against the FFI surface it uses: tests/stubs/xla/ffi/api/ffi.h mirrors that surface
    and its handler macro refuses bindings whose context / attribute / argument / result types do not match the
    implementation's signature.  A deliberately broken binding must be rejected (the check has teeth)."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(root, "difflexmm_b200", "csrc", "jax_ffi_shim.cc")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(root, "tests", "stubs"), "-I", os.path.join(root, "include"),
           "-I", "/usr/local/cuda/include"]
    ok = subprocess.run(cmd + [shim], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr[-3000:]
    src = open(shim).read()
    bad = src.replace(".Ret<ffi::Buffer<ffi::U8>>());", ".Ret<ffi::Buffer<ffi::F64>>());", 1)
    assert bad != src
    with tempfile.NamedTemporaryFile("w", suffix=".cc", delete=False) as f:
        f.write(bad)
    try:
        r = subprocess.run(cmd + [f.name], capture_output=True, text=True)
    finally:
        os.unlink(f.name)
    assert r.returncode != 0 and "static assertion" in r.stderr
