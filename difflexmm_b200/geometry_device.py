"""Design -> (centroid_node_vectors, block_centroids, inertia) on the device with a hand-written VJP
(libdfx `dfx_geometry_forward` / `dfx_geometry_vjp`).

Works for the lattices whose polygon vertices are `base + one design 2-vector` -- the reference's QuadGeometry and
KagomeGeometry (`geometry.py:607-952`).  The vertex -> design-variable table is not re-derived by hand: it is read off
the (tested) torch map `reference_node_vectors`, which is affine with 0/1 coefficients, by probing it once on the CPU.
"""
from typing import Sequence

import numpy as np
import torch

_F64 = torch.float64


def design_vertex_table(geometry):
    """-> (base (n_blocks, n_npb, 2), node_design (n_blocks, n_npb) int64 with -1 = none, shapes, sizes): every polygon
    vertex of the lattice is `base + design_flat[node_design]`, design_flat = the design arrays concatenated.  Read off
    the torch map `reference_node_vectors` by probing it once (host only, no GPU needed)."""
    if not hasattr(geometry, "reference_node_vectors"):
        geometry.compute_geometry()
    if not hasattr(geometry, "reference_node_vectors"):
        raise TypeError(f"{type(geometry).__name__} has no per-vertex design shifts (use its torch maps)")
    shapes = [tuple(s) for s in geometry.design_shapes]
    sizes = [int(np.prod(s[:-1])) for s in shapes]
    base = geometry.reference_node_vectors(*[torch.zeros(s, dtype=_F64) for s in shapes])
    # probe: design 2-vector k = (k + 1, -(k + 1)) -> the vertex offset identifies k
    probe, k0 = [], 0
    for s, n in zip(shapes, sizes):
        k = torch.arange(k0 + 1, k0 + n + 1, dtype=_F64).reshape(s[:-1])
        probe.append(torch.stack([k, -k], -1))
        k0 += n
    off = geometry.reference_node_vectors(*probe) - base
    idx = torch.round(off[..., 0]).to(torch.int64)
    tol = 1e-9 * max(1.0, float(base.abs().max()))
    if (off[..., 0] - idx).abs().max() > tol or (off[..., 1] + idx).abs().max() > tol or idx.min() < 0:
        raise ValueError("reference_node_vectors is not `base + one design vector per vertex`")
    return base, idx - 1, shapes, sizes


class _DesignToParams(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dg, design, density):
        from . import _lib
        cnv, cen, inertia = _lib.geometry_forward(dg.handle, design, density)
        ctx.dg = dg
        ctx.save_for_backward(design, density)
        return cnv, cen, inertia

    @staticmethod
    def backward(ctx, cnv_bar, cen_bar, inertia_bar):
        design, density = ctx.saved_tensors
        want_rho = ctx.needs_input_grad[2]
        from . import _lib
        design_bar, rho_bar = _lib.geometry_vjp(ctx.dg.handle, design, density, cnv_bar, cen_bar, inertia_bar, want_rho)
        if want_rho and density.dim() == 0:
            rho_bar = rho_bar.sum()
        return None, design_bar, rho_bar


class DeviceGeometry:
    """Device-side design map of one lattice geometry (`QuadGeometry` or `KagomeGeometry` of `difflexmm_b200.geometry`)."""

    def __init__(self, geometry, device="cuda"):
        from . import _lib  # raises if the CUDA library is missing
        base, node_design, self.shapes, self.sizes = design_vertex_table(geometry)
        self.geometry = geometry
        self.device = torch.device(device)
        self.n_design = sum(self.sizes)
        self.n_blocks, self.n_npb = base.shape[0], base.shape[1]
        self.handle = _lib.GeometryHandle(self.n_blocks, self.n_npb, self.n_design, base.numpy(), node_design.numpy(),
                                          self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.reference_points = geometry.reference_points.to(self.device)

    def flatten(self, design: Sequence[torch.Tensor]):
        """tuple of design arrays (optionally with a leading batch axis) -> (B, n_design, 2), batched flag"""
        parts = [torch.as_tensor(d, dtype=_F64, device=self.device) for d in design]
        batched = parts[0].dim() == len(self.shapes[0]) + 1
        B = parts[0].shape[0] if batched else 1
        flat = torch.cat([p.reshape(B, n, 2) for p, n in zip(parts, self.sizes)], dim=1)
        return flat, batched

    def __call__(self, design, density):
        """-> centroid_node_vectors ([B,] n_blocks, n_npb, 2), block_centroids ([B,] n_blocks, 2), inertia ([B,] n_blocks, 3);
        differentiable w.r.t. the design arrays and the density."""
        flat, batched = self.flatten(design)
        rho = torch.as_tensor(density, dtype=_F64, device=self.device)
        cnv, cen, inertia = _DesignToParams.apply(self, flat, rho)
        cnv = cnv.reshape(flat.shape[0], self.n_blocks, self.n_npb, 2)
        cen = cen + self.reference_points
        if not batched:
            return cnv[0], cen[0], inertia[0]
        return cnv, cen, inertia


class _AngleToParams(torch.autograd.Function):
    @staticmethod
    def forward(ctx, drs, angle, density):
        from . import _lib
        cnv, inertia = _lib.rotated_square_forward(drs.n1_blocks, drs.n2_blocks, drs.half_side, angle, density)
        ctx.drs = drs
        ctx.save_for_backward(angle, density)
        return cnv, inertia

    @staticmethod
    def backward(ctx, cnv_bar, inertia_bar):
        angle, density = ctx.saved_tensors
        want_rho = ctx.needs_input_grad[2]
        from . import _lib
        d = ctx.drs
        angle_bar, rho_bar = _lib.rotated_square_vjp(d.n1_blocks, d.n2_blocks, d.half_side, angle, density, cnv_bar, inertia_bar, want_rho)
        if want_rho and density.dim() == 0:
            rho_bar = rho_bar.sum()
        return None, angle_bar, rho_bar


class DeviceRotatedSquare:
    """Device-side design map of `RotatedSquareGeometry` (reference `geometry.py:354-443`): the design is one angle per
    lattice (libdfx `dfx_rotated_square_forward` / `_vjp`); the block centroids do not depend on it."""

    def __init__(self, geometry, device="cuda"):
        from . import _lib  # noqa: F401  (raises if the CUDA library is missing)
        if not hasattr(geometry, "_computed"):
            geometry.compute_geometry()
        self.geometry, self.device = geometry, torch.device(device)
        self.n1_blocks, self.n2_blocks, self.n_blocks = geometry.n1_blocks, geometry.n2_blocks, geometry.n_blocks
        self.half_side = (geometry.spacing - geometry.bond_length) / 2
        self.centroids = geometry.block_centroids().to(self.device)

    def __call__(self, angle, density):
        """angle () or (B,) -> centroid_node_vectors ([B,] n_blocks, 4, 2), block_centroids (n_blocks, 2), inertia ([B,] n_blocks, 3);
        differentiable w.r.t. the angle and the density"""
        a = torch.as_tensor(angle, dtype=_F64, device=self.device)
        batched = a.dim() == 1
        rho = torch.as_tensor(density, dtype=_F64, device=self.device)
        cnv, inertia = _AngleToParams.apply(self, a.reshape(-1), rho)
        if not batched:
            return cnv[0], self.centroids, inertia[0]
        return cnv, self.centroids, inertia


class DeviceConstraints:
    """Angle and edge-length inequality constraints of a lattice design (`<= 0` when satisfied) with their sparse
    Jacobian, one kernel launch per batch of designs (libdfx `dfx_constraints_eval`; reference
    `problems/quads_focusing.py:473-544` and the `jit(jacobian(...))` of the nlopt callbacks, `:585-588, :613-616`).

    Rows: the reference's `angle_constraints` rows (when both `min_void_angle` and `min_block_angle` are given, as in
    `run_optimization_nlopt`), then its `edge_length_constraints` rows (when `min_edge_length` is given).  A row depends
    on at most 4 design 2-vectors: `jac[b, row, slot, xy]` belongs to the flat design variable `2 * columns[row, slot] + xy`
    (flat design = the design arrays concatenated, as in `DeviceGeometry.flatten`); `columns < 0` marks an empty slot."""

    def __init__(self, device_geometry: DeviceGeometry, min_void_angle=None, min_block_angle=None, min_edge_length=None,
                 boundary_angle_constraint=False):
        from . import _lib
        geometry = device_geometry.geometry
        angles = min_void_angle is not None and min_block_angle is not None
        edges = min_edge_length is not None
        if not angles and not edges:
            raise ValueError("no constraint requested (give min_void_angle and min_block_angle, and / or min_edge_length)")
        boundary = None
        if angles and boundary_angle_constraint:
            from .optimization import quad_boundary_node_ids
            boundary = quad_boundary_node_ids(geometry.n1_blocks, geometry.n2_blocks)
        self.dg = device_geometry
        self.handle = _lib.ConstraintsHandle(device_geometry.handle, np.asarray(geometry.bond_connectivity()), boundary, angles, edges,
                                             min_void_angle if angles else 0.0, min_block_angle if angles else 0.0,
                                             min_edge_length if edges else 0.0)
        self.n_rows, self.n_angle_rows = self.handle.n_rows, self.handle.n_angle_rows
        self.columns = torch.from_numpy(self.handle.columns.astype(np.int64)).to(device_geometry.device)
        # flat scalar column of every Jacobian entry; empty slots point at column 0 and carry exact zeros
        col = self.columns.clamp_min(0)
        self.scalar_columns = torch.stack([2 * col, 2 * col + 1], dim=-1).reshape(self.n_rows, 8)

    def __call__(self, design_flat, want_jacobian=True):
        """design_flat (B, n_design, 2) or (B, 2 n_design) on the device -> values (B, rows), jac (B, rows, 4, 2) | None"""
        from . import _lib
        x = torch.as_tensor(design_flat, dtype=_F64, device=self.dg.device).reshape(-1, self.dg.n_design, 2)
        return _lib.constraints_eval(self.handle, x, want_jacobian)

    def dense_jacobian(self, jac):
        """(B, rows, 4, 2) -> (B, rows, 2 n_design) (tests, small lattices)"""
        B = jac.shape[0]
        out = torch.zeros((B, self.n_rows, 2 * self.dg.n_design), dtype=_F64, device=jac.device)
        out.scatter_add_(2, self.scalar_columns.expand(B, -1, -1), jac.reshape(B, self.n_rows, 8))
        return out
