"""Lattice geometries and polygon inertia (host side, torch float64, differentiable).

These are the design -> parameter maps that sit immediately *before* the solver
boundary.  They keep the reference's class names, constructor
arguments, block/node/bond numbering and `get_parametrization()` protocol
(`difflexmm/geometry.py:256-952`) so a `forward(design)` closure written for the
reference produces the same `ControlParams` here.  Everything is vectorised torch;
gradients w.r.t. the design variables come from torch autograd.
"""

import math
from typing import Callable, Tuple

import numpy as np
import torch

_F64 = torch.float64


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x.to(_F64)
    dev = like.device if isinstance(like, torch.Tensor) else None
    return torch.as_tensor(x, dtype=_F64, device=dev)


def rotation_matrix(angle):
    """2x2 rotation (reference `geometry.py:17-23`)."""
    angle = _t(angle)
    c, s = torch.cos(angle), torch.sin(angle)
    return torch.stack([torch.stack([c, -s]), torch.stack([s, c])])


# ----------------------------------------------------------------------------------
# Polygon properties, batched over a leading block axis (reference geometry.py:71-160)
# ----------------------------------------------------------------------------------

def _shoelace_terms(vertices):
    prev = torch.roll(vertices, shifts=1, dims=-2)
    cross = prev[..., 0] * vertices[..., 1] - prev[..., 1] * vertices[..., 0]
    return prev, cross


def polygon_area(vertices):
    """|signed shoelace area| of polygons `(..., n_vertices, 2)` (reference :71-84)."""
    _, cross = _shoelace_terms(vertices)
    return torch.abs(cross.sum(-1) / 2)


def polygon_centroid(vertices):
    """Centroid of polygons `(..., n_vertices, 2)` (reference :87-105)."""
    prev, cross = _shoelace_terms(vertices)
    area = torch.abs(cross.sum(-1) / 2)
    num = ((prev + vertices) * cross[..., None]).sum(-2)
    return num / (6 * area)[..., None]


def polygon_polar_moment(vertices):
    """Polar moment of area about the centroid (reference :108-127)."""
    c = polygon_centroid(vertices)[..., None, :]
    v2 = vertices - c
    v1 = torch.roll(vertices, shifts=1, dims=-2) - c
    cross = v1[..., 0] * v2[..., 1] - v1[..., 1] * v2[..., 0]
    quad = (v1[..., 0] ** 2 + v1[..., 0] * v2[..., 0] + v2[..., 0] ** 2
            + v1[..., 1] ** 2 + v1[..., 1] * v2[..., 1] + v2[..., 1] ** 2)
    return torch.abs((cross * quad).sum(-1) / 12)


def compute_inertia(vertices, density):
    """`[rho*A, rho*A, rho*J]` per block, shape (n_blocks, 3) (reference :144-160)."""
    vertices = _t(vertices)
    density = _t(density, vertices)
    area = polygon_area(vertices)
    polar = polygon_polar_moment(vertices)
    m = density * area
    return torch.stack([m, m, density * polar], dim=-1)


def DOFsInfo(n_blocks: int, constrained_block_DOF_pairs):
    """free / constrained / all DOF ids (reference `geometry.py:163-178`).

    Constrained ids keep the order of the input pairs, free ids are ascending.  The
    reference builds the free list with a Python `in` test per DOF; here it is one
    vectorised mask, which matters at 100x100 blocks.
    """
    pairs = np.asarray(
        constrained_block_DOF_pairs.cpu() if isinstance(constrained_block_DOF_pairs, torch.Tensor)
        else constrained_block_DOF_pairs)
    if pairs.size == 0:
        constrained = np.zeros((0,), dtype=np.int64)
    else:
        pairs = pairs.reshape(-1, 2).astype(np.int64)
        constrained = pairs[:, 0] * 3 + pairs[:, 1]
    all_ids = np.arange(n_blocks * 3, dtype=np.int64)
    mask = np.ones(n_blocks * 3, dtype=bool)
    mask[constrained] = False
    return all_ids[mask], constrained, all_ids


def compute_edge_lengths(centroid_node_vectors):
    """Edge lengths per block (reference `geometry.py:205-218`)."""
    cnv = _t(centroid_node_vectors)
    return torch.linalg.norm(torch.roll(cnv, 1, dims=1) - cnv, dim=2)


def angle_between_unit_vectors(u1, u2):
    """Signed angle from u1 to u2, in [-pi, pi] (reference `geometry.py:221-231`); vectorised over leading axes."""
    return torch.atan2(u1[..., 0] * u2[..., 1] - u1[..., 1] * u2[..., 0], u1[..., 0] * u2[..., 0] + u1[..., 1] * u2[..., 1])


def compute_edge_unit_vectors(current_block_nodes, node_id):
    """Unit vectors from vertex `node_id` (global ids, any shape) to the next / previous vertex of its block
    (reference `geometry.py:181-202`)."""
    nodes = _t(current_block_nodes)
    npb = nodes.shape[-2]
    node_id = torch.as_tensor(node_id, dtype=torch.int64, device=nodes.device)
    blk, l = node_id // npb, node_id % npb
    here = nodes[blk, l]
    e1 = nodes[blk, (l + 1) % npb] - here
    e2 = nodes[blk, (l - 1) % npb] - here
    return e1 / torch.linalg.norm(e1, dim=-1, keepdim=True), e2 / torch.linalg.norm(e2, dim=-1, keepdim=True)


def compute_edge_angles(current_block_nodes, nodes):
    """(void_angle_1, void_angle_2, block_angle_1, block_angle_2) of the bonds `nodes` (n_bonds, 2)
    (reference `geometry.py:234-253`, vectorised over the bonds)."""
    nodes = torch.as_tensor(np.asarray(nodes), dtype=torch.int64)
    b1n1, b1n2 = compute_edge_unit_vectors(current_block_nodes, nodes[..., 0])
    b2n1, b2n2 = compute_edge_unit_vectors(current_block_nodes, nodes[..., 1])
    return (angle_between_unit_vectors(b2n2, b1n1), angle_between_unit_vectors(b1n2, b2n1),
            angle_between_unit_vectors(b1n1, b1n2), angle_between_unit_vectors(b2n1, b2n2))


# ----------------------------------------------------------------------------------
# Geometry classes
# ----------------------------------------------------------------------------------

class Geometry:
    """Protocol of the reference's `Geometry` (`geometry.py:272-327`)."""

    n_blocks: int
    n_nodes: int
    n_npb: int
    block_centroids: Callable
    centroid_node_vectors: Callable
    bond_connectivity: Callable
    reference_bond_vectors: Callable

    def compute_geometry(self):
        raise NotImplementedError("Child classes should implement this method.")

    def get_parametrization(self) -> Tuple[Callable, Callable, Callable, Callable]:
        self.compute_geometry()
        return (self.block_centroids, self.centroid_node_vectors,
                self.bond_connectivity, self.reference_bond_vectors)

    def get_reference_geometry(self, *args):
        if not hasattr(self, "_computed"):
            self.compute_geometry()
        return self.centroid_node_vectors(*args) + self.block_centroids(*args)[:, None, :]


class LatticeGeometry(Geometry):
    def __init__(self, n1_cells: int, n2_cells: int, n_bpc: int, direct_basis=None):
        self.n1_cells = n1_cells
        self.n2_cells = n2_cells
        self.n_bpc = n_bpc
        self.n_cells = n1_cells * n2_cells
        self.n_blocks = self.n_cells * n_bpc
        self.direct_basis = _t(np.eye(2) if direct_basis is None else direct_basis)


def _grid_row_major(n1: int, n2: int):
    """(n1 index, n2 index) of every cell/block in the reference's row-wise numbering
    (`meshgrid(arange(n1), arange(n2))` flattened: n1 runs fastest)."""
    i2, i1 = np.divmod(np.arange(n1 * n2), n1)
    return i1, i2


def _square_grid_bonds(n1_blocks: int, n2_blocks: int):
    """node0 of (i,j) -- node2 of (i+1,j); node1 of (i,j) -- node3 of (i,j+1)
    (reference `geometry.py:414-421`, `:897-904`)."""
    j, i = np.meshgrid(np.arange(n2_blocks), np.arange(n1_blocks - 1), indexing="ij")
    h = np.stack([(n1_blocks * j + i) * 4, (n1_blocks * j + i + 1) * 4 + 2], -1).reshape(-1, 2)
    j, i = np.meshgrid(np.arange(n2_blocks - 1), np.arange(n1_blocks), indexing="ij")
    v = np.stack([(n1_blocks * j + i) * 4 + 1, (n1_blocks * (j + 1) + i) * 4 + 3], -1).reshape(-1, 2)
    return np.concatenate([h, v]).astype(np.int64)


def _square_grid_reference_bonds(n1_blocks, n2_blocks, bond_length):
    h = np.tile(np.array([[1., 0.]]) * bond_length, ((n1_blocks - 1) * n2_blocks, 1))
    v = np.tile(np.array([[0., 1.]]) * bond_length, ((n2_blocks - 1) * n1_blocks, 1))
    return torch.from_numpy(np.concatenate([h, v]))


_QUARTER_TURNS = [0., math.pi / 2, math.pi, 3 * math.pi / 2]  # linspace(0, 3pi/2, 4)


class RotatedSquareGeometry(LatticeGeometry):
    """Rotated-square lattice, design variable = one angle (reference :354-443)."""

    def __init__(self, n1_cells: int, n2_cells: int, spacing: float = 1., bond_length: float = 0.1):
        super().__init__(n1_cells, n2_cells, n_bpc=4, direct_basis=spacing * np.eye(2))
        self.spacing = spacing
        self.bond_length = bond_length
        self.n1_blocks = 2 * n1_cells
        self.n2_blocks = 2 * n2_cells
        self.n_npb = 4
        self.n_nodes = self.n_npb * self.n_blocks

    def compute_geometry(self):
        i1, i2 = _grid_row_major(self.n1_blocks, self.n2_blocks)
        parity = torch.from_numpy(((-1.0) ** (i1 + i2)))
        half = (self.spacing - self.bond_length) / 2
        quarter = torch.tensor(_QUARTER_TURNS, dtype=_F64)

        def centroid_node_vectors(angle):
            a = parity * _t(angle)
            # v0 = half / cos(a) * (cos a, sin a), then rotated by the quarter turns
            v0 = torch.stack([torch.cos(a), torch.sin(a)], -1) * (half / torch.cos(a))[:, None]
            c, s = torch.cos(quarter), torch.sin(quarter)
            x = c[None, :] * v0[:, None, 0] - s[None, :] * v0[:, None, 1]
            y = s[None, :] * v0[:, None, 0] + c[None, :] * v0[:, None, 1]
            return torch.stack([x, y], -1)

        def block_centroids(angle=None):
            g = torch.from_numpy(np.stack([i1, i2], -1).astype(np.float64))
            return g @ self.direct_basis

        self.centroid_node_vectors = centroid_node_vectors
        self.block_centroids = block_centroids
        self.bond_connectivity = lambda: _square_grid_bonds(self.n1_blocks, self.n2_blocks)
        self.reference_bond_vectors = lambda: _square_grid_reference_bonds(
            self.n1_blocks, self.n2_blocks, self.bond_length)
        self._computed = True


class QuadGeometry(LatticeGeometry):
    """Aperiodic quadrilateral lattice; design = (horizontal_shift (n1+1,n2,2),
    vertical_shift (n1,n2+1,2)) (reference :804-952)."""

    def __init__(self, n1_blocks: int, n2_blocks: int, spacing: float = 1.0, bond_length: float = 0.1):
        super().__init__(n1_blocks, n2_blocks, n_bpc=1, direct_basis=spacing * np.eye(2))
        self.spacing = spacing
        self.bond_length = bond_length
        self.n1_blocks = n1_blocks
        self.n2_blocks = n2_blocks
        self.n_npb = 4
        self.n_nodes = self.n_npb * self.n_blocks

    def compute_geometry(self):
        i1, i2 = _grid_row_major(self.n1_blocks, self.n2_blocks)
        half = (self.spacing - self.bond_length) / 2
        quarter = torch.tensor(_QUARTER_TURNS, dtype=_F64)
        v0s = torch.stack([torch.cos(quarter) * half, torch.sin(quarter) * half], -1)  # (4,2)
        ref_points = torch.from_numpy(np.stack([i1, i2], -1).astype(np.float64)) @ self.direct_basis
        i1_t, i2_t = torch.from_numpy(i1), torch.from_numpy(i2)

        def reference_node_vectors(horizontal_shift, vertical_shift):
            hs, vs = _t(horizontal_shift), _t(vertical_shift)
            dev = hs.device
            a, b = i1_t.to(dev), i2_t.to(dev)
            shifts = torch.stack([hs[a + 1, b], vs[a, b + 1], hs[a, b], vs[a, b]], 1)  # (n_blocks,4,2)
            return v0s.to(dev)[None] + shifts

        def centroid_node_vectors(horizontal_shift, vertical_shift):
            ref = reference_node_vectors(horizontal_shift, vertical_shift)
            return ref - polygon_centroid(ref)[:, None, :]

        def block_centroids(horizontal_shift, vertical_shift):
            ref = reference_node_vectors(horizontal_shift, vertical_shift)
            return ref_points.to(ref.device) + polygon_centroid(ref)

        self.centroid_node_vectors = centroid_node_vectors
        self.block_centroids = block_centroids
        self.reference_node_vectors = reference_node_vectors
        self.reference_points = ref_points
        self.design_shapes = [(self.n1_blocks + 1, self.n2_blocks, 2), (self.n1_blocks, self.n2_blocks + 1, 2)]
        self.bond_connectivity = lambda: _square_grid_bonds(self.n1_blocks, self.n2_blocks)
        self.reference_bond_vectors = lambda: _square_grid_reference_bonds(
            self.n1_blocks, self.n2_blocks, self.bond_length)
        self._computed = True

    def get_design_from_rotated_square(self, angle):
        """Shifts reproducing a rotated-square lattice of the given angle (reference :928-952)."""
        half = (self.spacing - self.bond_length) / 2

        def shift(n1, n2):
            a = (-1.0) ** (n1 + n2) * angle
            return (half / np.cos(a))[..., None] * np.stack([np.cos(a), np.sin(a)], -1) - np.array([1., 0.]) * half

        n1, n2 = np.meshgrid(np.arange(self.n1_blocks + 1), np.arange(self.n2_blocks), indexing="ij")
        hs = shift(n1, n2)
        n1, n2 = np.meshgrid(np.arange(self.n1_blocks), np.arange(self.n2_blocks + 1), indexing="ij")
        base = shift(n1, n2)
        c, s = np.cos(np.pi / 2), np.sin(np.pi / 2)
        vs = np.stack([c * base[..., 0] - s * base[..., 1], s * base[..., 0] + c * base[..., 1]], -1)
        return torch.from_numpy(hs), torch.from_numpy(vs)


class KagomeGeometry(LatticeGeometry):
    """Non-periodic kagome lattice of triangles, 2 blocks per cell; design =
    (shifts_1 (n1+1,n2,2), shifts_2 (n1,n2+1,2), shifts_3 (n1,n2,2)) (reference :607-801)."""

    def __init__(self, n1_cells: int, n2_cells: int, direct_basis=None, bond_length: float = 0.1):
        if direct_basis is None:
            direct_basis = np.array([[1., 0.], [math.cos(math.pi / 3), math.sin(math.pi / 3)]])
        super().__init__(n1_cells, n2_cells, n_bpc=2, direct_basis=direct_basis)
        self.bond_length = bond_length
        self.n_npb = 3
        self.n_nodes = self.n_npb * self.n_blocks

    def compute_geometry(self):
        bl = self.bond_length
        rv_int = bl * np.array([math.cos(math.pi / 6), math.sin(math.pi / 6)])
        rv_b1 = bl * np.array([0., -1.])
        rv_b2 = bl * np.array([-math.cos(math.pi / 6), math.sin(math.pi / 6)])
        a1, a2 = self.direct_basis[0], self.direct_basis[1]
        n1c, n2c = self.n1_cells, self.n2_cells

        base_1 = torch.stack([a1 / 2, a1 / 2 + a2 / 2, a2 / 2]) \
            - 0.5 * torch.from_numpy(np.stack([rv_b1, rv_int, rv_b2]))
        base_2 = torch.stack([a1 / 2 + a2 / 2, a1 + a2 / 2, a1 / 2 + a2]) \
            + 0.5 * torch.from_numpy(np.stack([rv_int, rv_b2, rv_b1]))

        def reference_node_vectors(shifts_1=None, shifts_2=None, shifts_3=None):
            s1 = torch.zeros((n1c + 1, n2c, 2), dtype=_F64) if shifts_1 is None else _t(shifts_1)
            s2 = torch.zeros((n1c, n2c + 1, 2), dtype=_F64) if shifts_2 is None else _t(shifts_2)
            s3 = torch.zeros((n1c, n2c, 2), dtype=_F64) if shifts_3 is None else _t(shifts_3)
            dev = s1.device
            # per cell (n1, n2): block_1 nodes get (shift_2_1, shift_3, shift_1_1),
            # block_2 nodes get (shift_3, shift_1_2, shift_2_2)
            b1 = base_1.to(dev) + torch.stack([s2[:, :-1], s3, s1[:-1]], dim=2)
            b2 = base_2.to(dev) + torch.stack([s3, s1[1:], s2[:, 1:]], dim=2)
            cells = torch.stack([b1, b2], dim=2)  # (n1, n2, bpc, npb, 2)
            return cells.transpose(0, 1).reshape(self.n_blocks, self.n_npb, 2)

        def centroid_node_vectors(shifts_1=None, shifts_2=None, shifts_3=None):
            ref = reference_node_vectors(shifts_1, shifts_2, shifts_3)
            return ref - polygon_centroid(ref)[:, None, :]

        i1, i2 = _grid_row_major(n1c, n2c)
        cell_points = torch.from_numpy(np.stack([i1, i2], -1).astype(np.float64)) @ self.direct_basis
        ref_points = cell_points.repeat_interleave(self.n_bpc, dim=0)

        def block_centroids(shifts_1=None, shifts_2=None, shifts_3=None):
            ref = reference_node_vectors(shifts_1, shifts_2, shifts_3)
            return ref_points.to(ref.device) + polygon_centroid(ref)

        def bond_connectivity():
            npc = self.n_npb * self.n_bpc
            cell = lambda n1, n2: (n2 * n1c + n1) * npc
            i1, i2 = _grid_row_major(n1c, n2c)
            internal = np.stack([1 + cell(i1, i2), 3 + cell(i1, i2)], -1)
            i1, i2 = _grid_row_major(n1c, n2c - 1)
            boundary1 = np.stack([0 + cell(i1, i2 + 1), 5 + cell(i1, i2)], -1)
            i1, i2 = _grid_row_major(n1c - 1, n2c)
            boundary2 = np.stack([2 + cell(i1 + 1, i2), 4 + cell(i1, i2)], -1)
            return np.concatenate([internal, boundary1, boundary2]).astype(np.int64)

        def reference_bond_vectors():
            return torch.from_numpy(np.concatenate([
                np.tile(rv_int, (self.n_cells, 1)),
                np.tile(rv_b1, (n1c * (n2c - 1), 1)),
                np.tile(rv_b2, ((n1c - 1) * n2c, 1)),
            ]))

        self.centroid_node_vectors = centroid_node_vectors
        self.block_centroids = block_centroids
        self.reference_node_vectors = reference_node_vectors
        self.reference_points = ref_points
        self.design_shapes = [(n1c + 1, n2c, 2), (n1c, n2c + 1, 2), (n1c, n2c, 2)]
        self.bond_connectivity = bond_connectivity
        self.reference_bond_vectors = reference_bond_vectors
        self._computed = True
