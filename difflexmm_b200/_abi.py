"""ctypes mirror of `include/dfx.h` (struct layouts and enum values only).

Shared by the product binding (`_lib.py`, device pointers into libdfx.so) and by the test
oracle loader (`oracle/__init__.py`, host pointers into libdfx_oracle.so).
"""

import ctypes as C

import numpy as np

DFX_BOND_LIGAMENT, DFX_BOND_LINEARIZED, DFX_BOND_SPRING = 0, 1, 2
DFX_DRIVE_ZERO, DFX_DRIVE_PULSE, DFX_DRIVE_HARMONIC, DFX_DRIVE_RAMP, DFX_DRIVE_STATIC_PULSE, DFX_DRIVE_TABLE = range(6)
DFX_LOAD_NONE, DFX_LOAD_RAMP, DFX_LOAD_SECH2 = range(3)
DFX_MAX_DRIVE_PARAMS = 5
DFX_MAX_LOAD_CONSTS = 4
DFX_OK = 0
DFX_STATUS_OK, DFX_STATUS_MAX_STEPS, DFX_STATUS_DT_UNDERFLOW, DFX_STATUS_NONFINITE = 0, 1, 2, 4

DRIVE_PARAM_NAMES = {
    DFX_DRIVE_ZERO: (),
    DFX_DRIVE_PULSE: ("amplitude", "loading_rate", "input_delay"),
    DFX_DRIVE_HARMONIC: ("amplitude", "loading_rate", "input_delay"),
    DFX_DRIVE_RAMP: ("amplitude", "loading_rate"),
    DFX_DRIVE_STATIC_PULSE: ("amplitude", "loading_rate", "compressive_strain",
                             "compressive_strain_rate", "input_delay"),
    DFX_DRIVE_TABLE: (),
}

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class DfxTopologyDesc(C.Structure):
    _fields_ = [
        ("n_blocks", C.c_int32), ("n_npb", C.c_int32), ("n_bonds", C.c_int32),
        ("bond_nodes", _i32p),
        ("n_constrained", C.c_int32), ("constrained_dofs", _i32p),
        ("bond_energy", C.c_int32), ("contact", C.c_int32), ("drive_kind", C.c_int32),
        ("drive_vec0", _f64p), ("drive_vec1", _f64p),
        ("drive_table_len", C.c_int32), ("drive_table_t", _f64p), ("drive_table_v", _f64p),
        ("load_kind", C.c_int32), ("n_loaded", C.c_int32),
        ("loaded_dofs", _i32p), ("load_vec", _f64p),
        ("load_consts", C.c_double * DFX_MAX_LOAD_CONSTS),
        ("n_damped", C.c_int32), ("damped_blocks", _i32p),
    ]


class DfxLeaf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("bstride", C.c_int64)]


class DfxParams(C.Structure):
    _fields_ = [
        ("centroid_node_vectors", DfxLeaf), ("reference_vector", DfxLeaf),
        ("k_stretch", DfxLeaf), ("k_shear", DfxLeaf), ("k_rot", DfxLeaf),
        ("k_per_bond", C.c_int32 * 3),
        ("damping", DfxLeaf), ("damping_per_dof", C.c_int32),
        ("inertia", DfxLeaf), ("contact", DfxLeaf), ("drive", DfxLeaf), ("block_centroids", DfxLeaf),
    ]


class DfxParamGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "centroid_node_vectors", "reference_vector", "k_stretch", "k_shear", "k_rot",
        "damping", "inertia", "contact", "drive", "block_centroids")]


class DfxOptions(C.Structure):
    _fields_ = [("init_step_variant", C.c_int32), ("threads", C.c_int32), ("max_steps", C.c_int64),
                ("design_order", C.c_void_p)]


class DfxGeometryDesc(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("n_npb", C.c_int32), ("n_design", C.c_int32),
                ("base_nodes", C.c_void_p), ("node_design", C.c_void_p)]


class DfxConstraintDesc(C.Structure):
    _fields_ = [("n_bonds", C.c_int32), ("bonds", C.c_void_p), ("n_boundary", C.c_int32), ("boundary_nodes", C.c_void_p),
                ("angles", C.c_int32), ("edges", C.c_int32),
                ("min_void_angle", C.c_double), ("min_block_angle", C.c_double), ("min_edge_length", C.c_double)]


DFX_OBJ_KINETIC, DFX_OBJ_ANGULAR = 0, 1
DFX_CONTACT_NONE, DFX_CONTACT_ANGLE, DFX_CONTACT_DISTANCE = 0, 1, 2


class DfxObjective(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_target", C.c_int32), ("target_free_ids", C.c_void_p), ("weights", C.c_void_p),
                ("arm", C.c_void_p), ("arm_bstride", C.c_int64)]


class DfxStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("accepted", C.c_int64), ("rhs_evals", C.c_int64),
                ("status", C.c_int32), ("reserved", C.c_int32), ("last_dt", C.c_double)]


STATS_DTYPE = np.dtype([("steps", "<i8"), ("accepted", "<i8"), ("rhs_evals", "<i8"),
                        ("status", "<i4"), ("reserved", "<i4"), ("last_dt", "<f8")])
assert STATS_DTYPE.itemsize == C.sizeof(DfxStats)


class TopologySpec:
    """Host-side, numpy description of what `setup_dynamic_solver` closes over
    (reference `dynamics.py:60-136`).  `to_desc()` yields the C struct; the numpy arrays are
    kept alive on the instance."""

    def __init__(self, n_blocks, n_npb, bond_nodes, constrained_dofs=(), bond_energy=DFX_BOND_LIGAMENT,
                 contact=False, drive_kind=DFX_DRIVE_ZERO, drive_vec0=None, drive_vec1=None,
                 load_kind=DFX_LOAD_NONE, loaded_dofs=(), load_vec=None, load_consts=(), damped_blocks=(),
                 drive_table=None):
        self.n_blocks, self.n_npb = int(n_blocks), int(n_npb)
        self.bond_nodes = np.ascontiguousarray(np.asarray(bond_nodes, dtype=np.int32).reshape(-1, 2))
        self.constrained_dofs = np.ascontiguousarray(np.asarray(constrained_dofs, dtype=np.int32).reshape(-1))
        # contact: False / True (angle-based) or DFX_CONTACT_* (2 = distance-based between the void edges)
        self.bond_energy, self.contact, self.drive_kind = int(bond_energy), int(contact), int(drive_kind)
        if self.contact not in (DFX_CONTACT_NONE, DFX_CONTACT_ANGLE, DFX_CONTACT_DISTANCE):
            raise ValueError(f"unknown contact kind {contact}")
        nc = len(self.constrained_dofs)

        def vec(v):
            if v is None:
                return None
            v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
            if len(v) != nc:
                raise ValueError(f"drive vector has {len(v)} entries, expected n_constrained={nc}")
            return v
        self.drive_vec0, self.drive_vec1 = vec(drive_vec0), vec(drive_vec1)
        self.drive_table = None
        if drive_table is not None:
            tt, tv = (np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1)) for x in drive_table)
            if len(tt) != len(tv) or len(tt) < 1 or np.any(np.diff(tt) < 0):
                raise ValueError("tabulated drive needs equally long, increasing times and values")
            self.drive_table = (tt, tv)
        if (self.drive_kind == DFX_DRIVE_TABLE) != (self.drive_table is not None):
            raise ValueError("drive_table must be given exactly for DFX_DRIVE_TABLE")
        self.load_kind = int(load_kind)
        self.loaded_dofs = np.ascontiguousarray(np.asarray(loaded_dofs, dtype=np.int32).reshape(-1))
        self.load_vec = None if load_vec is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(load_vec, dtype=np.float64), self.loaded_dofs.shape))
        self.load_consts = tuple(float(x) for x in load_consts)
        self.damped_blocks = np.ascontiguousarray(np.asarray(damped_blocks, dtype=np.int32).reshape(-1))
        n_dof = 3 * self.n_blocks
        if self.bond_nodes.size and (self.bond_nodes.min() < 0 or self.bond_nodes.max() >= self.n_blocks * self.n_npb):
            raise ValueError("bond_connectivity refers to a node outside the geometry")
        if nc and (self.constrained_dofs.min() < 0 or self.constrained_dofs.max() >= n_dof):
            raise ValueError("constrained_block_DOF_pairs refers to a DOF outside the geometry")
        if len(np.unique(self.constrained_dofs)) != nc:
            raise ValueError("constrained_block_DOF_pairs contains a repeated [block, DOF] pair")
        mask = np.ones(n_dof, dtype=bool)
        mask[self.constrained_dofs] = False
        self.free_dofs = np.nonzero(mask)[0].astype(np.int64)
        self.n_free = len(self.free_dofs)
        self.n_bonds = len(self.bond_nodes)
        self.n_drive_params = len(DRIVE_PARAM_NAMES[self.drive_kind])

    def to_desc(self):
        d = DfxTopologyDesc()
        d.n_blocks, d.n_npb, d.n_bonds = self.n_blocks, self.n_npb, self.n_bonds
        d.bond_nodes = self.bond_nodes.ctypes.data_as(_i32p)
        d.n_constrained = len(self.constrained_dofs)
        d.constrained_dofs = self.constrained_dofs.ctypes.data_as(_i32p)
        d.bond_energy, d.contact, d.drive_kind = self.bond_energy, int(self.contact), self.drive_kind
        d.drive_vec0 = self.drive_vec0.ctypes.data_as(_f64p) if self.drive_vec0 is not None else None
        d.drive_vec1 = self.drive_vec1.ctypes.data_as(_f64p) if self.drive_vec1 is not None else None
        if self.drive_table is not None:
            d.drive_table_len = len(self.drive_table[0])
            d.drive_table_t = self.drive_table[0].ctypes.data_as(_f64p)
            d.drive_table_v = self.drive_table[1].ctypes.data_as(_f64p)
        d.load_kind, d.n_loaded = self.load_kind, len(self.loaded_dofs)
        d.loaded_dofs = self.loaded_dofs.ctypes.data_as(_i32p)
        d.load_vec = self.load_vec.ctypes.data_as(_f64p) if self.load_vec is not None else None
        for i, c in enumerate(self.load_consts[:DFX_MAX_LOAD_CONSTS]):
            d.load_consts[i] = c
        d.n_damped = len(self.damped_blocks)
        d.damped_blocks = self.damped_blocks.ctypes.data_as(_i32p)
        return d


LEAF_NAMES = ("centroid_node_vectors", "reference_vector", "k_stretch", "k_shear", "k_rot",
              "damping", "inertia", "contact", "drive", "block_centroids")


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class ParamSet:
    """The runtime leaves of one call, for `batch` designs.

    Every leaf is a float64 array (numpy on the host for the oracle, torch on the device for
    libdfx) of its base shape (shared by all designs) or `(batch,) + base shape`:
      centroid_node_vectors (n_blocks, n_npb, 2)   reference_vector (n_bonds, 2)
      k_stretch / k_shear / k_rot  () or (n_bonds,) if named in `per_bond`
      damping () or (n_damped, 3) if `damping_per_dof`      inertia (n_free,)
      contact (3,) = (min_angle, cutoff_angle, k_contact)   drive (n_drive_params,)
    """

    def __init__(self, spec: TopologySpec, batch: int, leaves: dict, per_bond=(), damping_per_dof=False):
        self.spec, self.batch = spec, int(batch)
        self.per_bond = tuple(per_bond)
        self.damping_per_dof = bool(damping_per_dof)
        self.base_shapes = {
            "centroid_node_vectors": (spec.n_blocks, spec.n_npb, 2),
            "reference_vector": (spec.n_bonds, 2),
            "k_stretch": (spec.n_bonds,) if "k_stretch" in self.per_bond else (),
            "k_shear": (spec.n_bonds,) if "k_shear" in self.per_bond else (),
            "k_rot": (spec.n_bonds,) if "k_rot" in self.per_bond else (),
            "damping": (len(spec.damped_blocks), 3) if self.damping_per_dof else (),
            "inertia": (spec.n_free,),
            "contact": (3,),
            "drive": (spec.n_drive_params,),
            "block_centroids": (spec.n_blocks, 2),
        }
        self.leaves, self.batched = {}, {}
        for name in LEAF_NAMES:
            a = leaves.get(name)
            needed = not ((name == "contact" and not spec.contact) or
                          (name == "drive" and spec.n_drive_params == 0) or
                          (name == "damping" and len(spec.damped_blocks) == 0) or
                          (name == "block_centroids" and spec.contact != DFX_CONTACT_DISTANCE))
            if a is None:
                if needed:
                    raise ValueError(f"missing parameter leaf '{name}'")
                continue
            base = self.base_shapes[name]
            shape = tuple(a.shape)
            if shape == base:
                self.batched[name] = False
            elif shape == (self.batch,) + base:
                self.batched[name] = True
            else:
                raise ValueError(f"leaf '{name}' has shape {shape}; expected {base} or {(self.batch,) + base}")
            self.leaves[name] = a

    def numel(self, name):
        n = 1
        for s in self.base_shapes[name]:
            n *= s
        return n

    def to_struct(self):
        p = DfxParams()
        for name in LEAF_NAMES:
            a = self.leaves.get(name)
            leaf = DfxLeaf(_ptr(a), self.numel(name) if (a is not None and self.batched[name]) else 0)
            setattr(p, name, leaf)
        for i, n in enumerate(("k_stretch", "k_shear", "k_rot")):
            p.k_per_bond[i] = int(n in self.per_bond)
        p.damping_per_dof = int(self.damping_per_dof)
        return p

    def aug_size_of_listed_leaves(self):
        """2*(2 n_free) + 1 + sizes of the leaves listed here (what libdfx counts when aug_size=0)."""
        return 4 * self.spec.n_free + 1 + sum(self.numel(n) for n in self.leaves)
