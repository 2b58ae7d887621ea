"""Shared builders for the test problems (inputs at the solver boundary).

`load_golden(name)` rebuilds a problem from a fixture written by oracle/make_golden.py;
`quads_problem(...)` / `kagome_problem(...)` build the reference configurations (cfg1 / cfg2 recipes
of SURVEY section 8d) at any lattice size."""

import os

import numpy as np

from difflexmm_b200 import _abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEAVES = _abi.LEAF_NAMES


class Case:
    def __init__(self, spec, leaves, per_bond, damping_per_dof, y0, ts, rtol, atol, aug_size, g=None, ref=None):
        self.spec, self.leaves, self.per_bond, self.damping_per_dof = spec, leaves, per_bond, damping_per_dof
        self.y0, self.ts, self.rtol, self.atol, self.aug_size, self.g, self.ref = y0, ts, rtol, atol, aug_size, g, ref


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else []


def load_golden(name) -> Case:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    spec = _abi.TopologySpec(
        int(z["n_blocks"]), int(z["n_npb"]), z["bond_nodes"], z["constrained_dofs"], int(z["bond_energy"]),
        int(z["contact"]), int(z["drive_kind"]),
        z["drive_vec0"] if z["drive_vec0"].size else None, z["drive_vec1"] if z["drive_vec1"].size else None,
        int(z["load_kind"]), z["loaded_dofs"], None, tuple(z["load_consts"]), z["damped_blocks"])
    leaves = {n: z["leaf_" + n] for n in LEAVES if "leaf_" + n in z.files}
    if spec.n_drive_params:
        leaves["drive"] = z["drive"]
    per_bond = tuple(n for n in ("k_stretch", "k_shear", "k_rot") if leaves[n].ndim == 1)
    dpd = "damping" in leaves and leaves["damping"].ndim == 2
    ref = {k: z[k] for k in z.files if k.startswith("grad_") or k in ("ys", "y0_bar", "ts_bar", "fwd_steps", "fwd_accepted",
                                                                         "bwd_steps", "bwd_accepted")}
    return Case(spec, leaves, per_bond, dpd, z["y0"], z["ts"], float(z["rtol"]), float(z["atol"]), int(z["aug_size"]),
                z["g"], ref)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den > 0 else np.linalg.norm(a)
