"""TEST INFRASTRUCTURE ONLY: loader for the CPU oracle (`oracle/dfx_oracle.cpp`).

May be imported by tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs only -- never by `difflexmm_b200`.  Parity status: unpinned against
real JAX (not installable here); see the header of `dfx_oracle.cpp`.
"""

import ctypes as C
import os
import subprocess

import numpy as np

from difflexmm_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """builds the checker (literal arithmetic: -O2, no FMA contraction) and the baseline build of the same source
    (-O3 -march=x86-64-v3 with FMA: the fastest honest CPU port, used by bench.py's CPU legs only)"""
    so = os.path.join(_HERE, "libdfx_oracle.so")
    fast = os.path.join(_HERE, "libdfx_oracle_fast.so")
    src = os.path.join(_HERE, "dfx_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "dfx.h")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr)) if os.path.exists(src) else 0.0
    if force or not os.path.exists(so) or not os.path.exists(fast) or min(os.path.getmtime(so), os.path.getmtime(fast)) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return so


_FAST = False


def use_fast_build(on=True):
    """bench.py's CPU legs: switch to the -O3 / AVX2 / FMA build when the host supports it (results differ from the
    checker build at round-off level).  Returns whether the fast build is in use."""
    global _LIB, _FAST
    ok = False
    if on:
        try:
            flags = open("/proc/cpuinfo").read()
            ok = " avx2" in flags and " fma" in flags and os.path.exists(os.path.join(_HERE, "libdfx_oracle_fast.so"))
        except OSError:
            ok = False
    if ok != _FAST:
        _LIB, _FAST = None, ok
    return ok


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        _LIB = C.CDLL(os.path.join(_HERE, "libdfx_oracle_fast.so") if _FAST else so)
        _LIB.dfxo_energy.restype = C.c_double
    return _LIB


def _f64(a):
    a = np.asarray(a, dtype=np.float64)
    return a if a.flags.c_contiguous else np.array(a, order="C")  # keeps 0-d leaves 0-d


class Oracle:
    """numpy front-end of the oracle for one topology."""

    def __init__(self, spec: _abi.TopologySpec):
        self.spec = spec
        self.desc = spec.to_desc()
        self.n_free = spec.n_free

    def params(self, batch, leaves, per_bond=(), damping_per_dof=False):
        return _abi.ParamSet(self.spec, batch, {k: _f64(v) for k, v in leaves.items() if v is not None},
                             per_bond, damping_per_dof)

    def _opt(self, variant, max_steps=0):
        return _abi.DfxOptions(int(variant), 0, int(max_steps))

    def forward(self, ps, y0, ts, rtol, atol, variant=0, n_threads=1):
        B, N = ps.batch, 2 * self.n_free
        y0, ts = _f64(y0), _f64(ts)
        n_t = ts.shape[-1]
        ys = np.empty((B, n_t, N))
        stats = np.zeros(B, dtype=_abi.STATS_DTYPE)
        p, o = ps.to_struct(), self._opt(variant)
        rc = lib().dfxo_forward(C.byref(self.desc), C.byref(p), B, y0.ctypes.data_as(C.c_void_p),
                                C.c_int64(N if y0.ndim == 2 else 0), ts.ctypes.data_as(C.c_void_p),
                                C.c_int64(n_t if ts.ndim == 2 else 0), n_t, C.c_double(rtol), C.c_double(atol),
                                C.byref(o), ys.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p),
                                int(n_threads))
        assert rc == 0
        return ys, stats

    def adjoint(self, ps, ys, ts, g, rtol, atol, aug_size=0, variant=0, n_threads=1):
        B, N = ps.batch, 2 * self.n_free
        ys, ts, g = _f64(ys), _f64(ts), _f64(g)
        n_t = ts.shape[-1]
        out = {n: np.zeros((B,) + ps.base_shapes[n]) for n in ps.leaves}
        gr = _abi.DfxParamGrads()
        for n, a in out.items():
            setattr(gr, n, a.ctypes.data)
        y0_bar, ts_bar = np.zeros((B, N)), np.zeros((B, n_t))
        stats = np.zeros(B, dtype=_abi.STATS_DTYPE)
        p, o = ps.to_struct(), self._opt(variant)
        rc = lib().dfxo_adjoint(C.byref(self.desc), C.byref(p), B, ys.ctypes.data_as(C.c_void_p),
                                ts.ctypes.data_as(C.c_void_p), C.c_int64(n_t if ts.ndim == 2 else 0), n_t,
                                g.ctypes.data_as(C.c_void_p), C.c_double(rtol), C.c_double(atol), C.c_int64(aug_size),
                                C.byref(o), y0_bar.ctypes.data_as(C.c_void_p), ts_bar.ctypes.data_as(C.c_void_p),
                                C.byref(gr), stats.ctypes.data_as(C.c_void_p), int(n_threads))
        assert rc == 0
        return y0_bar, ts_bar, out, stats

    def expand_fields(self, ps, ys, ts):
        B = ps.batch
        ys, ts = _f64(ys), _f64(ts)
        n_t = ts.shape[-1]
        fields = np.zeros((B, n_t, 2, self.spec.n_blocks, 3))
        p = ps.to_struct()
        lib().dfxo_expand_fields(C.byref(self.desc), C.byref(p), B, ys.ctypes.data_as(C.c_void_p),
                                 ts.ctypes.data_as(C.c_void_p), C.c_int64(n_t if ts.ndim == 2 else 0), n_t,
                                 fields.ctypes.data_as(C.c_void_p))
        return fields

    def rhs(self, ps, y, t):
        y = _f64(y)
        out = np.empty_like(y)
        p = ps.to_struct()
        lib().dfxo_rhs(C.byref(self.desc), C.byref(p), y.ctypes.data_as(C.c_void_p), C.c_double(t),
                       out.ctypes.data_as(C.c_void_p))
        return out

    def aug_size(self, ps):
        p = ps.to_struct()
        return lib().dfxo_aug_size(C.byref(self.desc), C.byref(p))

    def aug_rhs(self, ps, z, s):
        z = _f64(z)
        out = np.empty_like(z)
        p = ps.to_struct()
        lib().dfxo_aug_rhs(C.byref(self.desc), C.byref(p), z.ctypes.data_as(C.c_void_p), C.c_double(s),
                           out.ctypes.data_as(C.c_void_p))
        return out

    def energy(self, ps, U):
        U = _f64(U)
        p = ps.to_struct()
        return lib().dfxo_energy(C.byref(self.desc), C.byref(p), U.ctypes.data_as(C.c_void_p))
