"""ctypes binding of libdfx.so (the C ABI declared in include/dfx.h).

Importing this module loads the CUDA library or raises: the solver has no CPU or eager fallback.
Device buffers come from torch (allocation, streams); the library itself never sees a torch type.
"""

import ctypes as C
import os

import numpy as np
import torch

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdfx.so")
if os.environ.get("DFX_LIB"):  # experiments: load another build of the same ABI
    _SO = os.environ["DFX_LIB"]


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> libdfx.so for sm_100a (nvcc cross-compiles without a GPU); see build_native.py."""
    from . import build_native
    return build_native.build(force=force, extra_flags=("-Xptxas=-v",) if verbose else ())


if not os.path.exists(_SO):
    raise ImportError(
        f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(difflexmm_b200 has no CPU path)")

lib = C.CDLL(_SO)
lib.dfx_last_error.restype = C.c_char_p
lib.dfx_version.restype = C.c_char_p
lib.dfx_forward_workspace_bytes.restype = C.c_size_t
lib.dfx_adjoint_workspace_bytes.restype = C.c_size_t
lib.dfx_fp64_peak.restype = C.c_double
lib.dfx_adjoint_plan.restype = C.c_char_p

EXPORTS = ("dfx_topology_create", "dfx_topology_destroy", "dfx_topology_n_free", "dfx_drive_n_params",
           "dfx_forward_workspace_bytes", "dfx_adjoint_workspace_bytes", "dfx_forward", "dfx_adjoint",
           "dfx_expand_fields", "dfx_objective", "dfx_adjoint_objective",
           "dfx_geometry_create", "dfx_geometry_destroy", "dfx_geometry_forward", "dfx_geometry_vjp",
           "dfx_rotated_square_forward", "dfx_rotated_square_vjp",
           "dfx_constraints_create", "dfx_constraints_destroy", "dfx_constraints_rows", "dfx_constraints_columns", "dfx_constraints_eval",
           "dfx_fp64_peak", "dfx_math_selftest", "dfx_sincos_selftest", "dfx_adjoint_plan", "dfx_last_error", "dfx_version")


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib.dfx_last_error().decode()}")


class Topology:
    """Owns one `DfxTopology*` on one device."""

    def __init__(self, spec: _abi.TopologySpec, device_index: int):
        self.spec, self.device_index = spec, int(device_index)
        self._desc = spec.to_desc()
        self._h = C.c_void_p()
        _check(lib.dfx_topology_create(C.byref(self._desc), self.device_index, C.byref(self._h)), "dfx_topology_create")
        assert lib.dfx_topology_n_free(self._h) == spec.n_free

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.dfx_topology_destroy(h)


class GeometryHandle:
    """Owns one `DfxGeometry*`: polygon vertices = base_nodes + design[node_design]."""

    def __init__(self, n_blocks, n_npb, n_design, base_nodes, node_design, device_index):
        import numpy as np
        self.n_blocks, self.n_npb, self.n_design = int(n_blocks), int(n_npb), int(n_design)
        self._base = np.ascontiguousarray(base_nodes, dtype=np.float64).reshape(self.n_blocks * self.n_npb, 2)
        self._nd = np.ascontiguousarray(node_design, dtype=np.int32).reshape(self.n_blocks * self.n_npb)
        desc = _abi.DfxGeometryDesc(self.n_blocks, self.n_npb, self.n_design, self._base.ctypes.data, self._nd.ctypes.data)
        self._h = C.c_void_p()
        _check(lib.dfx_geometry_create(C.byref(desc), int(device_index), C.byref(self._h)), "dfx_geometry_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.dfx_geometry_destroy(h)


def rotated_square_forward(n1_blocks, n2_blocks, half_side, angle, density):
    """angle (B,), density (B,) or () -> cnv (B, n_blocks, 4, 2), inertia (B, n_blocks, 3)"""
    dev, B, nb = angle.device, angle.shape[0], n1_blocks * n2_blocks
    angle, density = angle.contiguous(), density.contiguous()
    cnv = torch.empty((B, nb, 4, 2), dtype=torch.float64, device=dev)
    inertia = torch.empty((B, nb, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _check(lib.dfx_rotated_square_forward(int(n1_blocks), int(n2_blocks), C.c_double(half_side), B, C.c_void_p(angle.data_ptr()),
                                              C.c_void_p(density.data_ptr()), C.c_int64(1 if density.dim() == 1 else 0),
                                              C.c_void_p(cnv.data_ptr()), C.c_void_p(inertia.data_ptr()), _stream_ptr(dev)),
               "dfx_rotated_square_forward")
    return cnv, inertia


def rotated_square_vjp(n1_blocks, n2_blocks, half_side, angle, density, cnv_bar, inertia_bar, want_density_bar=False):
    dev, B = angle.device, angle.shape[0]
    angle, density = angle.contiguous(), density.contiguous()
    cnv_bar, inertia_bar = [None if t is None else t.contiguous() for t in (cnv_bar, inertia_bar)]
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)
    angle_bar = torch.empty((B,), dtype=torch.float64, device=dev)
    density_bar = torch.empty((B,), dtype=torch.float64, device=dev) if want_density_bar else None
    with torch.cuda.device(dev):
        _check(lib.dfx_rotated_square_vjp(int(n1_blocks), int(n2_blocks), C.c_double(half_side), B, C.c_void_p(angle.data_ptr()),
                                          C.c_void_p(density.data_ptr()), C.c_int64(1 if density.dim() == 1 else 0), ptr(cnv_bar),
                                          ptr(inertia_bar), C.c_void_p(angle_bar.data_ptr()), ptr(density_bar), _stream_ptr(dev)),
               "dfx_rotated_square_vjp")
    return angle_bar, density_bar


class ConstraintsHandle:
    """Owns one `DfxConstraints*`: angle / edge-length inequality rows of a lattice geometry and their column table."""

    def __init__(self, geo: GeometryHandle, bonds, boundary_nodes, angles, edges, min_void_angle, min_block_angle, min_edge_length):
        import numpy as np
        self.geo = geo  # keeps the geometry alive
        self._bonds = np.ascontiguousarray(bonds, dtype=np.int32).reshape(-1, 2)
        self._boundary = np.ascontiguousarray(boundary_nodes if boundary_nodes is not None else [], dtype=np.int32).reshape(-1)
        desc = _abi.DfxConstraintDesc(len(self._bonds), self._bonds.ctypes.data, len(self._boundary),
                                      self._boundary.ctypes.data if len(self._boundary) else None, int(bool(angles)), int(bool(edges)),
                                      float(min_void_angle), float(min_block_angle), float(min_edge_length))
        self._h = C.c_void_p()
        _check(lib.dfx_constraints_create(geo._h, C.byref(desc), C.byref(self._h)), "dfx_constraints_create")
        n_angle = C.c_int32(0)
        self.n_rows = int(lib.dfx_constraints_rows(self._h, C.byref(n_angle)))
        self.n_angle_rows = int(n_angle.value)
        self.columns = np.empty((self.n_rows, 4), dtype=np.int32)
        _check(lib.dfx_constraints_columns(self._h, C.c_void_p(self.columns.ctypes.data)), "dfx_constraints_columns")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.dfx_constraints_destroy(h)


def constraints_eval(con: ConstraintsHandle, design, want_jacobian=True):
    """design (B, n_design, 2) -> values (B, rows), jac (B, rows, 4, 2) or None: one launch for the whole batch"""
    dev, B = design.device, design.shape[0]
    design = design.contiguous()
    values = torch.empty((B, con.n_rows), dtype=torch.float64, device=dev)
    jac = torch.empty((B, con.n_rows, 4, 2), dtype=torch.float64, device=dev) if want_jacobian else None
    with torch.cuda.device(dev):
        _check(lib.dfx_constraints_eval(con._h, B, C.c_void_p(design.data_ptr()), C.c_void_p(values.data_ptr()),
                                        C.c_void_p(jac.data_ptr() if jac is not None else None), _stream_ptr(dev)), "dfx_constraints_eval")
    return values, jac


def geometry_forward(geo: GeometryHandle, design, density):
    """design (B, n_design, 2), density (B,) or () -> cnv (B, n_nodes, 2), centroid_shift (B, n_blocks, 2), inertia (B, n_blocks, 3)"""
    dev, B = design.device, design.shape[0]
    design, density = design.contiguous(), density.contiguous()
    cnv = torch.empty((B, geo.n_blocks * geo.n_npb, 2), dtype=torch.float64, device=dev)
    cen = torch.empty((B, geo.n_blocks, 2), dtype=torch.float64, device=dev)
    inertia = torch.empty((B, geo.n_blocks, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _check(lib.dfx_geometry_forward(geo._h, B, C.c_void_p(design.data_ptr()), C.c_void_p(density.data_ptr()),
                                        C.c_int64(1 if density.dim() == 1 else 0), C.c_void_p(cnv.data_ptr()),
                                        C.c_void_p(cen.data_ptr()), C.c_void_p(inertia.data_ptr()), _stream_ptr(dev)),
               "dfx_geometry_forward")
    return cnv, cen, inertia


def geometry_vjp(geo: GeometryHandle, design, density, cnv_bar, centroid_bar, inertia_bar, want_density_bar=False):
    dev, B = design.device, design.shape[0]
    design, density = design.contiguous(), density.contiguous()
    cnv_bar, centroid_bar, inertia_bar = [None if t is None else t.contiguous() for t in (cnv_bar, centroid_bar, inertia_bar)]
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)
    design_bar = torch.empty((B, geo.n_design, 2), dtype=torch.float64, device=dev)
    density_bar = torch.empty((B,), dtype=torch.float64, device=dev) if want_density_bar else None
    with torch.cuda.device(dev):
        _check(lib.dfx_geometry_vjp(geo._h, B, C.c_void_p(design.data_ptr()), C.c_void_p(density.data_ptr()),
                                    C.c_int64(1 if density.dim() == 1 else 0), ptr(cnv_bar), ptr(centroid_bar), ptr(inertia_bar),
                                    C.c_void_p(design_bar.data_ptr()), ptr(density_bar), _stream_ptr(dev)), "dfx_geometry_vjp")
    return design_bar, density_bar


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev(t):
    return t.device


def forward(topo: Topology, ps: _abi.ParamSet, y0, ts, rtol, atol, options: _abi.DfxOptions):
    """-> ys (B, n_t, 2 n_free) device tensor, stats (B,) numpy structured array (synchronises)."""
    spec, B = topo.spec, ps.batch
    N = 2 * spec.n_free
    dev = y0.device
    assert y0.is_cuda and y0.dtype == torch.float64 and ts.dtype == torch.float64
    n_t = ts.shape[-1]
    y0 = y0.contiguous()
    ts = ts.contiguous()
    if y0.shape[-1] != N:
        raise ValueError(f"state0 has {y0.shape[-1]} free entries, expected {N}")
    ys = torch.empty((B, n_t, N), dtype=torch.float64, device=dev)
    stats = torch.zeros((B, _abi.STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    p = ps.to_struct()
    with torch.cuda.device(dev):
        ws_bytes = lib.dfx_forward_workspace_bytes(topo._h, B)
        ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
        _check(lib.dfx_forward(topo._h, C.byref(p), B, C.c_void_p(y0.data_ptr()), C.c_int64(N if y0.dim() == 2 else 0),
                               C.c_void_p(ts.data_ptr()), C.c_int64(n_t if ts.dim() == 2 else 0), n_t,
                               C.c_double(rtol), C.c_double(atol), C.byref(options),
                               C.c_void_p(ys.data_ptr()), C.c_void_p(stats.data_ptr()),
                               C.c_void_p(ws.data_ptr()), C.c_size_t(ws_bytes), _stream_ptr(dev)), "dfx_forward")
    return ys, _Stats(stats)


def longest_first(stats: "_Stats", options: _abi.DfxOptions, min_batch: int = 149):
    """Options for the adjoint launch that follows a forward solve: the designs are launched in decreasing order of
    their forward step count (a good predictor of the adjoint's), so that a launch with more designs than SMs does not end
    with one long design on an otherwise idle GPU.  Returns (options, order tensor to keep alive); small batches (one
    wave) keep the plain order.  No host synchronisation."""
    steps = stats.steps_device()
    if steps.numel() < min_batch:
        return options, None
    order = torch.argsort(steps, descending=True, stable=True).to(torch.int32)
    opt = _abi.DfxOptions(options.init_step_variant, options.threads, options.max_steps, order.data_ptr())
    return opt, order


def adjoint_plan(topo: Topology, ps: _abi.ParamSet) -> str:
    """name of the adjoint kernel the library launches for this topology / these leaf forms / this batch (diagnostic)"""
    p = ps.to_struct()
    return lib.dfx_adjoint_plan(topo._h, C.byref(p), ps.batch).decode()


def adjoint(topo: Topology, ps: _abi.ParamSet, ys, ts, g, rtol, atol, aug_size, options: _abi.DfxOptions):
    """-> y0_bar (B, 2 n_free), ts_bar (B, n_t), grads {leaf: (B,)+shape}, stats."""
    spec, B = topo.spec, ps.batch
    N = 2 * spec.n_free
    dev = ys.device
    n_t = ts.shape[-1]
    ys, ts, g = ys.contiguous(), ts.contiguous(), g.contiguous()
    y0_bar = torch.zeros((B, N), dtype=torch.float64, device=dev)
    ts_bar = torch.zeros((B, n_t), dtype=torch.float64, device=dev)
    grads = {n: torch.zeros((B,) + ps.base_shapes[n], dtype=torch.float64, device=dev) for n in ps.leaves}
    gs = _abi.DfxParamGrads()
    for n, t in grads.items():
        setattr(gs, n, t.data_ptr())
    stats = torch.zeros((B, _abi.STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    p = ps.to_struct()
    with torch.cuda.device(dev):
        ws_bytes = lib.dfx_adjoint_workspace_bytes(topo._h, B)
        ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
        _check(lib.dfx_adjoint(topo._h, C.byref(p), B, C.c_void_p(ys.data_ptr()), C.c_void_p(ts.data_ptr()),
                               C.c_int64(n_t if ts.dim() == 2 else 0), n_t, C.c_void_p(g.data_ptr()),
                               C.c_double(rtol), C.c_double(atol), C.c_int64(int(aug_size)), C.byref(options),
                               C.c_void_p(y0_bar.data_ptr()), C.c_void_p(ts_bar.data_ptr()), C.byref(gs),
                               C.c_void_p(stats.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws_bytes),
                               _stream_ptr(dev)), "dfx_adjoint")
    return y0_bar, ts_bar, grads, _Stats(stats)


def _objective_struct(kind, ids, weights, arm, B):
    """-> (DfxObjective, tensors to keep alive)"""
    keep = [ids]
    obj = _abi.DfxObjective()
    obj.kind, obj.n_target, obj.target_free_ids = int(kind), ids.numel(), ids.data_ptr()
    if weights is not None:
        keep.append(weights)
        obj.weights = weights.data_ptr()
    if arm is not None:
        keep.append(arm)
        obj.arm, obj.arm_bstride = arm.data_ptr(), (arm[0].numel() if arm.dim() == 3 else 0)
    return obj, keep


def objective_value(topo: Topology, ps: _abi.ParamSet, ys, target_free_ids, kind=_abi.DFX_OBJ_KINETIC, arm=None):
    """-> J (B,), dJ/d(inertia) (B, n_free), dJ/d(arm) (B, n_target/3, 2) or None.
    kinetic: J = sum_t sum_target 1/2 m v^2; angular: sum_t sum_blocks (arm + u) x (m v) + I omega (include/dfx.h)."""
    spec, B = topo.spec, ps.batch
    dev = ys.device
    n_t = ys.shape[1]
    ys = ys.contiguous()
    ids = target_free_ids.to(device=dev, dtype=torch.int32).contiguous()
    arm = None if arm is None else arm.to(device=dev, dtype=torch.float64).contiguous()
    value = torch.empty((B,), dtype=torch.float64, device=dev)
    ibar = torch.empty((B, spec.n_free), dtype=torch.float64, device=dev)
    abar = torch.empty((B, ids.numel() // 3, 2), dtype=torch.float64, device=dev) if kind == _abi.DFX_OBJ_ANGULAR else None
    obj, keep = _objective_struct(kind, ids, None, arm, B)
    p = ps.to_struct()
    with torch.cuda.device(dev):
        _check(lib.dfx_objective(topo._h, C.byref(p), B, C.c_void_p(ys.data_ptr()), n_t, C.byref(obj),
                                 C.c_void_p(value.data_ptr()), C.c_void_p(ibar.data_ptr()),
                                 C.c_void_p(abar.data_ptr() if abar is not None else None), _stream_ptr(dev)), "dfx_objective")
    return value, ibar, abar


def adjoint_objective(topo: Topology, ps: _abi.ParamSet, ys, ts, target_free_ids, weights, rtol, atol, aug_size,
                      options: _abi.DfxOptions, kind=_abi.DFX_OBJ_KINETIC, arm=None):
    """`adjoint` with the cotangent of the device objective generated in the kernel (weights[b] = dL/dJ_b)."""
    spec, B = topo.spec, ps.batch
    N = 2 * spec.n_free
    dev = ys.device
    n_t = ts.shape[-1]
    ys, ts = ys.contiguous(), ts.contiguous()
    ids = target_free_ids.to(device=dev, dtype=torch.int32).contiguous()
    w = weights.to(device=dev, dtype=torch.float64).expand(B).contiguous()
    arm = None if arm is None else arm.to(device=dev, dtype=torch.float64).contiguous()
    y0_bar = torch.zeros((B, N), dtype=torch.float64, device=dev)
    ts_bar = torch.zeros((B, n_t), dtype=torch.float64, device=dev)
    grads = {n: torch.zeros((B,) + ps.base_shapes[n], dtype=torch.float64, device=dev) for n in ps.leaves}
    gs = _abi.DfxParamGrads()
    for n, t in grads.items():
        setattr(gs, n, t.data_ptr())
    stats = torch.zeros((B, _abi.STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    obj, keep = _objective_struct(kind, ids, w, arm, B)
    p = ps.to_struct()
    with torch.cuda.device(dev):
        ws_bytes = lib.dfx_adjoint_workspace_bytes(topo._h, B)
        ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
        _check(lib.dfx_adjoint_objective(topo._h, C.byref(p), B, C.c_void_p(ys.data_ptr()), C.c_void_p(ts.data_ptr()),
                                         C.c_int64(n_t if ts.dim() == 2 else 0), n_t, C.byref(obj),
                                         C.c_double(rtol), C.c_double(atol), C.c_int64(int(aug_size)), C.byref(options),
                                         C.c_void_p(y0_bar.data_ptr()), C.c_void_p(ts_bar.data_ptr()), C.byref(gs),
                                         C.c_void_p(stats.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws_bytes),
                                         _stream_ptr(dev)), "dfx_adjoint_objective")
    return y0_bar, ts_bar, grads, _Stats(stats)


def expand_fields(topo: Topology, ps: _abi.ParamSet, ys, ts):
    spec, B = topo.spec, ps.batch
    dev = ys.device
    n_t = ts.shape[-1]
    fields = torch.empty((B, n_t, 2, spec.n_blocks, 3), dtype=torch.float64, device=dev)
    p = ps.to_struct()
    with torch.cuda.device(dev):
        _check(lib.dfx_expand_fields(topo._h, C.byref(p), B, C.c_void_p(ys.contiguous().data_ptr()),
                                     C.c_void_p(ts.contiguous().data_ptr()), C.c_int64(n_t if ts.dim() == 2 else 0), n_t,
                                     C.c_void_p(fields.data_ptr()), _stream_ptr(dev)), "dfx_expand_fields")
    return fields


class _Stats:
    """Per-design solver statistics; stays on the device until read (reading synchronises)."""

    def __init__(self, raw):
        self._raw = raw

    def numpy(self):
        return self._raw.cpu().numpy().view(_abi.STATS_DTYPE).reshape(-1)

    def steps_device(self):
        """attempted step counts as an int64 tensor on the device (no synchronisation)"""
        return self._raw.view(torch.int64)[:, 0]

    def __getitem__(self, k):
        return self.numpy()[k]
