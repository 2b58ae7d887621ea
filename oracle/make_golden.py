"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz with oracle/ref_literal.py.

Each fixture stores the complete inputs at the solver boundary (topology arrays, parameter
leaves, y0, ts, rtol/atol, the cotangent g) and the outputs of the literal torch-autograd
restatement of the reference (forward trajectory `ys`; adjoint cotangents of y0, ts and every
parameter leaf; attempted/accepted step counts).  tests/ compare both the fast C++ oracle and
the CUDA path against these files.

Run from the repo root:  python oracle/make_golden.py [case ...]      (minutes per case on CPU)

The literal restatement follows jax 0.4.8's odeint from knowledge of its source (jax is not
installable here): fixtures pin the *restatement*, not a real JAX run -- "parity unpinned".
"""

import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from difflexmm_b200 import _abi  # noqa: E402
from difflexmm_b200.geometry import (DOFsInfo, KagomeGeometry, QuadGeometry, RotatedSquareGeometry,  # noqa: E402
                                     compute_inertia)
from oracle import ref_literal as L  # noqa: E402

F64 = torch.float64
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def T(x):
    return torch.tensor(x, dtype=F64)


def _pulse(t, amplitude, loading_rate, windowed=True):
    on = (t > 0.) & (t < loading_rate ** -1) if windowed else (t > 0.)
    return amplitude * torch.where(on, (1 - torch.cos(2 * math.pi * loading_rate * t)) / 2, torch.zeros((), dtype=F64))


def run_case(name, *, n_blocks, n_npb, bonds, cons, bond_energy, use_contact, drive_kind, drive_vec0, drive_vec1,
             drive_params, load, damped_blocks, leaves, block_centroids, ts, rtol, atol, y0=None, density_leaf=None,
             extra_leaf_entries=0, g_mode="kinetic"):
    """leaves: dict of torch tensors with the libdfx leaf names; drive_params: dict name->value."""
    free, cons_ids, _ = DOFsInfo(n_blocks, np.stack([cons // 3, cons % 3], -1))
    nf = len(free)
    names = _abi.DRIVE_PARAM_NAMES[drive_kind]
    v0 = None if drive_vec0 is None else torch.as_tensor(drive_vec0, dtype=F64)
    v1 = None if drive_vec1 is None else torch.as_tensor(drive_vec1, dtype=F64)

    def cfn(t, **p):
        if drive_kind == _abi.DFX_DRIVE_ZERO:
            return torch.zeros((), dtype=F64)
        if drive_kind in (_abi.DFX_DRIVE_PULSE, _abi.DFX_DRIVE_HARMONIC):
            return _pulse(t - p["input_delay"], p["amplitude"], p["loading_rate"], drive_kind == _abi.DFX_DRIVE_PULSE) * v0
        if drive_kind == _abi.DFX_DRIVE_RAMP:
            return p["amplitude"] * torch.where(t < p["loading_rate"] ** -1, t * p["loading_rate"], torch.ones((), dtype=F64)) * v0
        if drive_kind == _abi.DFX_DRIVE_STATIC_PULSE:
            cs, csr = p["compressive_strain"], p["compressive_strain_rate"]
            static = torch.where(t < cs * csr ** -1, t * csr, cs) * v1
            return static + _pulse(t - cs * csr ** -1 - p["input_delay"], p["amplitude"], p["loading_rate"]) * v0
        raise ValueError(drive_kind)

    loading_fn, loaded = None, None
    if load is not None:
        loaded = load["dofs"]
        if load["kind"] == _abi.DFX_LOAD_RAMP:
            c0, c1 = load["consts"]
            loading_fn = lambda state, t: c0 * torch.where(t < 1.0 / c1, t * c1, torch.ones((), dtype=F64))  # noqa: E731
        else:
            c0, c1 = load["consts"]
            loading_fn = lambda state, t: 2 * c0 / c1 ** 2 * torch.cosh(t / c1 - 3) ** (-2) * torch.tanh(3 - t / c1)  # noqa: E731

    prob = L.Problem(n_blocks, n_npb, bonds, cons, bond_energy={0: "ligament", 1: "linearized", 2: "spring"}[bond_energy],
                     use_contact=use_contact, constrained_DOFs_fn=cfn, loaded_DOF_ids=loaded, loading_fn=loading_fn,
                     damped_blocks=damped_blocks)
    P = dict(block_centroids=block_centroids, centroid_node_vectors=leaves["centroid_node_vectors"],
             k_stretch=leaves["k_stretch"], k_shear=leaves["k_shear"], k_rot=leaves["k_rot"],
             reference_vector=leaves["reference_vector"], inertia=leaves["inertia"],
             constraint_params={n: T(drive_params[n]) for n in names}, loading_params={})
    if bond_energy == _abi.DFX_BOND_SPRING:  # StretchingTorsionalSpringParams has no k_shear / reference_vector leaves
        del P["k_shear"], P["reference_vector"]
    if damped_blocks is not None:
        P["damping"] = leaves["damping"]
    if use_contact:
        P["min_angle"], P["cutoff_angle"], P["k_contact"] = leaves["contact"][0], leaves["contact"][1], leaves["contact"][2]
    if density_leaf is not None:
        P["density"] = density_leaf  # present in the reference pytree, unused inside odeint
    y0 = torch.zeros(2 * nf, dtype=F64) if y0 is None else torch.as_tensor(y0, dtype=F64)
    ts = torch.as_tensor(ts, dtype=F64)
    st_f, st_b = {}, {}
    t0 = time.time()
    ys = L.solve_forward(prob, y0, ts, P, rtol, atol, stats=st_f)
    t1 = time.time()
    if g_mode == "kinetic":  # d/dys of sum_t sum_dof 1/2 m v^2
        g = torch.zeros_like(ys)
        g[:, nf:] = ys[:, nf:] * leaves["inertia"]
    else:  # generic smooth functional touching displacements and velocities
        w = torch.linspace(0.5, 1.5, ys.numel(), dtype=F64).reshape(ys.shape)
        g = w * torch.cos(ys) + 0.1 * w
    y0_bar, ts_bar, gb = L.solve_adjoint(prob, ys, ts, P, g, rtol, atol, stats=st_b)
    t2 = time.time()
    _, flat = L.flatten_leaves(P)
    aug_size = 4 * nf + 1 + sum(int(torch.as_tensor(x).numel()) for x in flat)  # = length of the literal run's augmented vector
    out = dict(
        n_blocks=n_blocks, n_npb=n_npb, bond_nodes=np.asarray(bonds, dtype=np.int32), constrained_dofs=np.asarray(cons, dtype=np.int32),
        bond_energy=bond_energy, contact=2 if use_contact == "distance" else int(bool(use_contact)), drive_kind=drive_kind,
        drive_vec0=np.zeros(0) if drive_vec0 is None else np.asarray(drive_vec0, dtype=np.float64),
        drive_vec1=np.zeros(0) if drive_vec1 is None else np.asarray(drive_vec1, dtype=np.float64),
        load_kind=0 if load is None else load["kind"], loaded_dofs=np.zeros(0, dtype=np.int32) if load is None else np.asarray(load["dofs"], dtype=np.int32),
        load_consts=np.zeros(0) if load is None else np.asarray(load["consts"], dtype=np.float64),
        damped_blocks=np.zeros(0, dtype=np.int32) if damped_blocks is None else np.asarray(damped_blocks, dtype=np.int32),
        y0=y0.numpy(), ts=ts.numpy(), rtol=rtol, atol=atol, g=g.numpy(), aug_size=aug_size,
        ys=ys.numpy(), y0_bar=y0_bar.numpy(), ts_bar=ts_bar.numpy(),
        fwd_steps=st_f["steps"], fwd_accepted=st_f["accepted"], bwd_steps=st_b["steps"], bwd_accepted=st_b["accepted"],
        drive=np.asarray([drive_params[n] for n in names], dtype=np.float64),
    )
    for k, v in leaves.items():
        out["leaf_" + k] = np.asarray(v.detach().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float64)
    gmap = {"centroid_node_vectors": "centroid_node_vectors", "reference_vector": "reference_vector",
            "k_stretch": "k_stretch", "k_shear": "k_shear", "k_rot": "k_rot", "damping": "damping", "inertia": "inertia"}
    for k, src in gmap.items():
        if src in gb:
            out["grad_" + k] = gb[src].numpy()
    if use_contact:
        out["grad_contact"] = np.array([gb["min_angle"].item(), gb["cutoff_angle"].item(), gb["k_contact"].item()])
    if use_contact == "distance":  # the only energy that reads block positions (energy.py:395-405)
        out["leaf_block_centroids"] = np.asarray(block_centroids.detach().numpy(), dtype=np.float64)
        out["grad_block_centroids"] = gb["block_centroids"].numpy()
    out["grad_drive"] = np.array([gb["constraint_params." + n].item() for n in names])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: n_free={nf} fwd {st_f} {t1 - t0:.0f}s  bwd {st_b} {t2 - t1:.0f}s  aug_size={aug_size}", flush=True)


def quad_case(name, n1, n2, *, contact_window, noise, n_t, t_end, rtol, atol, seed=0, g_mode="kinetic", use_contact=True, angle=25.,
              initial_rotation=0.0):
    torch.manual_seed(seed)
    geo = QuadGeometry(n1, n2, spacing=15., bond_length=2.25)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    hs, vs = geo.get_design_from_rotated_square(angle * math.pi / 180)
    hs = hs + noise * torch.randn_like(hs)
    vs = vs + noise * torch.randn_like(vs)
    cnv, cent = cnvf(hs, vs), bc(hs, vs)
    mid = (n2 // 2) * n1  # driven block on the left edge
    pairs = np.array([[mid, 0], [mid, 1], [mid, 2], [0, 0], [0, 1], [0, 2], [n1 - 1, 0], [n1 - 1, 1], [n1 - 1, 2]])
    cons = pairs[:, 0] * 3 + pairs[:, 1]
    free, _, _ = DOFsInfo(geo.n_blocks, pairs)
    lv = np.zeros(len(cons))
    lv[0] = 1
    rho = 6.18e-9
    inertia = compute_inertia(cnv, rho).reshape(-1)[free]
    damp = T(0.0186) * torch.ones(geo.n_blocks, 3, dtype=F64) * T(
        [2 * (0.36125 * rho * 15 ** 2 * 1.19) ** .5] * 2 + [2 * (0.02175026 * rho * 15 ** 4 * 1.5) ** .5])
    leaves = dict(centroid_node_vectors=cnv, reference_vector=refv(), k_stretch=T(120.), k_shear=T(1.19), k_rot=T(1.5),
                  damping=damp, inertia=inertia, contact=T([contact_window[0], contact_window[1], 1.5]))
    run_case(name, n_blocks=geo.n_blocks, n_npb=4, bonds=bonds(), cons=cons, bond_energy=0, use_contact=use_contact,
             drive_kind=_abi.DFX_DRIVE_PULSE, drive_vec0=lv, drive_vec1=None,
             drive_params=dict(amplitude=7.5, loading_rate=30., input_delay=0.1 / 30), load=None,
             damped_blocks=np.arange(geo.n_blocks), leaves=leaves, block_centroids=cent,
             ts=np.linspace(0, t_end, n_t), rtol=rtol, atol=atol, density_leaf=T(rho), g_mode=g_mode,
             y0=_counter_rotated_state(geo, free, initial_rotation) if initial_rotation else None)


def _counter_rotated_state(geo, free, angle):
    """initial state with neighbouring blocks rotated by +-angle (the rotating-squares mechanism): voids close on one side,
    so that the distance-based contact takes its point-on-edge branch, not only the hinge-to-hinge distance"""
    u = np.zeros((geo.n_blocks, 3))
    ids = np.arange(geo.n_blocks)
    u[:, 2] = angle * (-1.0) ** (ids % geo.n1_blocks + ids // geo.n1_blocks)
    return np.concatenate([u.reshape(-1)[free], np.zeros(len(free))])


def spring_case(name, n1, n2, *, n_t, t_end, rtol, atol, seed=4):
    """quad lattice whose bonds are zero-length stretching / torsional springs (energy.py:49-66,
    StretchingTorsionalSpringParams utils.py:80-91): no k_shear / reference_vector leaves in the reference pytree;
    per-bond k_rot leaf, harmonic drive, scalar damping."""
    torch.manual_seed(seed)
    geo = QuadGeometry(n1, n2, spacing=15., bond_length=2.25)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    hs, vs = geo.get_design_from_rotated_square(20 * math.pi / 180)
    hs = hs + 0.3 * torch.randn_like(hs)
    vs = vs + 0.3 * torch.randn_like(vs)
    cnv, cent = cnvf(hs, vs), bc(hs, vs)
    mid = (n2 // 2) * n1
    pairs = np.array([[mid, 0], [mid, 1], [mid, 2], [0, 0], [0, 1], [0, 2], [n1 - 1, 0], [n1 - 1, 1], [n1 - 1, 2]])
    cons = pairs[:, 0] * 3 + pairs[:, 1]
    free, _, _ = DOFsInfo(geo.n_blocks, pairs)
    lv = np.zeros(len(cons))
    lv[0] = 1
    rho = 6.18e-9
    inertia = compute_inertia(cnv, rho).reshape(-1)[free]
    nb = len(bonds())
    # k_shear / reference_vector: placeholders required by the libdfx ABI, ignored by DFX_BOND_SPRING
    leaves = dict(centroid_node_vectors=cnv, reference_vector=torch.tensor([[1., 0.]], dtype=F64).repeat(nb, 1),
                  k_stretch=T(2.5), k_shear=T(0.), k_rot=1.5 * (1 + 0.2 * torch.rand(nb, dtype=F64)),
                  damping=T(2.0e-5), inertia=inertia)
    run_case(name, n_blocks=geo.n_blocks, n_npb=4, bonds=bonds(), cons=cons, bond_energy=_abi.DFX_BOND_SPRING, use_contact=False,
             drive_kind=_abi.DFX_DRIVE_HARMONIC, drive_vec0=lv, drive_vec1=None,
             drive_params=dict(amplitude=3., loading_rate=40., input_delay=0.002), load=None,
             damped_blocks=np.arange(geo.n_blocks), leaves=leaves, block_centroids=cent,
             ts=np.linspace(0, t_end, n_t), rtol=rtol, atol=atol, density_leaf=T(rho), g_mode="generic")


def kagome_case(name, n1, n2, *, n_t, t_end, rtol, atol, seed=1):
    torch.manual_seed(seed)
    geo = KagomeGeometry(n1, n2, direct_basis=20. * np.array([[1., 0.], [math.cos(math.pi / 3), math.sin(math.pi / 3)]]),
                         bond_length=2.25)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    s1 = 0.4 * torch.randn(n1 + 1, n2, 2, dtype=F64)
    s2 = 0.4 * torch.randn(n1, n2 + 1, 2, dtype=F64)
    s3 = 0.4 * torch.randn(n1, n2, 2, dtype=F64)
    cnv, cent = cnvf(s1, s2, s3), bc(s1, s2, s3)
    drv = 2 * n1 * (n2 // 2)
    pairs = np.array([[drv, 0], [drv, 1], [drv, 2], [0, 0], [0, 1], [0, 2], [2 * n1 - 1, 0], [2 * n1 - 1, 1], [2 * n1 - 1, 2]])
    cons = pairs[:, 0] * 3 + pairs[:, 1]
    free, _, _ = DOFsInfo(geo.n_blocks, pairs)
    lv = np.zeros(len(cons))
    lv[0] = 1
    rho = 6.18e-9
    inertia = compute_inertia(cnv, rho).reshape(-1)[free]
    leaves = dict(centroid_node_vectors=cnv, reference_vector=refv(),
                  k_stretch=120. * (1 + 0.1 * torch.rand(len(bonds()), dtype=F64)),  # per-bond leaf
                  k_shear=T(1.19), k_rot=T(1.5), damping=T(2.0e-5),                  # scalar damping leaf
                  inertia=inertia, contact=T([-15 * math.pi / 180, -10 * math.pi / 180, 1.5]))
    run_case(name, n_blocks=geo.n_blocks, n_npb=3, bonds=bonds(), cons=cons, bond_energy=0, use_contact=True,
             drive_kind=_abi.DFX_DRIVE_HARMONIC, drive_vec0=lv, drive_vec1=None,
             drive_params=dict(amplitude=6., loading_rate=40., input_delay=0.002), load=None,
             damped_blocks=np.arange(geo.n_blocks), leaves=leaves, block_centroids=cent,
             ts=np.linspace(0, t_end, n_t), rtol=rtol, atol=atol, density_leaf=T(rho), g_mode="generic")


def static_pulse_case(name, n1, n2, *, n_t, rtol, atol):
    """cfg4 recipe in miniature: top/bottom rows follow a compression ramp, a left block gets the
    delayed pulse (problems/quads_kinetic_energy_static_tuning.py:124-196,275-281)."""
    geo = QuadGeometry(n1, n2, spacing=15., bond_length=2.25)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    hs, vs = geo.get_design_from_rotated_square(25 * math.pi / 180)
    cnv, cent = cnvf(hs, vs), bc(hs, vs)
    drv = (n2 // 2) * n1
    bottom = np.arange(n1)
    top = np.arange(geo.n_blocks - n1, geo.n_blocks)
    pairs = np.concatenate([
        np.array([[drv, 0], [drv, 1], [drv, 2]]),
        np.stack([np.tile(bottom, 3), np.repeat([1, 0, 2], n1)], -1),
        np.stack([np.tile(top, 3), np.repeat([1, 0, 2], n1)], -1)])
    cons = pairs[:, 0] * 3 + pairs[:, 1]
    free, _, _ = DOFsInfo(geo.n_blocks, pairs)
    dyn = np.zeros(len(cons))
    dyn[0] = 1
    sta = np.zeros(len(cons))
    sta[3:3 + n1] = 0.5
    sta[3 + 3 * n1:3 + 4 * n1] = -0.5
    sta *= (n2 - 1) * 15.
    rho = 6.18e-9
    inertia = compute_inertia(cnv, rho).reshape(-1)[free]
    damp = T(0.0186) * torch.ones(geo.n_blocks, 3, dtype=F64) * T(
        [2 * (0.36125 * rho * 15 ** 2 * 1.19) ** .5] * 2 + [2 * (0.02175026 * rho * 15 ** 4 * 1.5) ** .5])
    leaves = dict(centroid_node_vectors=cnv, reference_vector=refv(), k_stretch=T(120.), k_shear=T(1.19), k_rot=T(1.5),
                  damping=damp, inertia=inertia, contact=T([-10 * math.pi / 180, -5 * math.pi / 180, 1.5]))
    cs, csr, f = 0.01, 25.0, 30.
    t_static = cs / csr
    delay = 0.1 / f
    ts = np.concatenate([[0.], np.linspace(t_static + delay, t_static + delay + 0.2 / f, n_t)])
    run_case(name, n_blocks=geo.n_blocks, n_npb=4, bonds=bonds(), cons=cons, bond_energy=0, use_contact=True,
             drive_kind=_abi.DFX_DRIVE_STATIC_PULSE, drive_vec0=dyn, drive_vec1=sta,
             drive_params=dict(amplitude=7.5, loading_rate=f, compressive_strain=cs, compressive_strain_rate=csr,
                               input_delay=delay), load=None,
             damped_blocks=np.arange(geo.n_blocks), leaves=leaves, block_centroids=cent,
             ts=ts, rtol=rtol, atol=atol, density_leaf=T(rho))


def ramp_sech2_case(name, n1, n2, *, n_t, t_end, rtol, atol, seed=6):
    """the two signal kinds no other fixture holds: constrained DOFs on a linear ramp to a plateau
    (problems/hinge_characterization.py:134-139) and an external load with the sech^2 * tanh pulse of
    scripts/pulse_RS.py:49-50 on the right-hand column; per-DOF damping, contact present but not entered."""
    torch.manual_seed(seed)
    geo = QuadGeometry(n1, n2, spacing=15., bond_length=2.25)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    hs, vs = geo.get_design_from_rotated_square(25 * math.pi / 180)
    hs = hs + 0.2 * torch.randn_like(hs)
    vs = vs + 0.2 * torch.randn_like(vs)
    cnv, cent = cnvf(hs, vs), bc(hs, vs)
    mid = (n2 // 2) * n1
    pairs = np.array([[mid, 0], [mid, 1], [mid, 2], [0, 0], [0, 1], [0, 2]])
    cons = pairs[:, 0] * 3 + pairs[:, 1]
    free, _, _ = DOFsInfo(geo.n_blocks, pairs)
    lv = np.zeros(len(cons))
    lv[0] = 1
    rho = 6.18e-9
    inertia = compute_inertia(cnv, rho).reshape(-1)[free]
    damp = T(0.0186) * torch.ones(geo.n_blocks, 3, dtype=F64) * T(
        [2 * (0.36125 * rho * 15 ** 2 * 1.19) ** .5] * 2 + [2 * (0.02175026 * rho * 15 ** 4 * 1.5) ** .5])
    leaves = dict(centroid_node_vectors=cnv, reference_vector=refv(), k_stretch=T(120.), k_shear=T(1.19), k_rot=T(1.5),
                  damping=damp, inertia=inertia, contact=T([-15 * math.pi / 180, -10 * math.pi / 180, 1.5]))
    loaded = np.array([(j * n1 + n1 - 1) * 3 for j in range(n2)])  # x DOF of the right-hand column
    run_case(name, n_blocks=geo.n_blocks, n_npb=4, bonds=bonds(), cons=cons, bond_energy=0, use_contact=True,
             drive_kind=_abi.DFX_DRIVE_RAMP, drive_vec0=lv, drive_vec1=None,
             drive_params=dict(amplitude=2.5, loading_rate=250.), load=dict(kind=_abi.DFX_LOAD_SECH2, dofs=loaded, consts=(2.0e-7, 0.0012)),
             damped_blocks=np.arange(geo.n_blocks), leaves=leaves, block_centroids=cent,
             ts=np.linspace(0, t_end, n_t), rtol=rtol, atol=atol, density_leaf=T(rho), g_mode="generic")


def tensile_case(name, n1_cells, bond_energy, *, n_t, t_end):
    """tests/test_difflexmm.py:35-146 in miniature (loading_fn, explicit inertia, clamped x DOFs)."""
    geo = RotatedSquareGeometry(n1_cells=n1_cells, n2_cells=1, spacing=1.0)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    cnv, cent = cnvf(0.), bc(0.)
    k_stretch = 1.0
    mass = 1.0
    Jrot = 1.815 ** -2 / 4 * mass
    inertia_full = torch.ones(geo.n_blocks, 3, dtype=F64) * T([mass, mass, Jrot])
    damping = 0.05 * torch.ones(geo.n_blocks, 3, dtype=F64) * T([1., 1., 0.25])
    pairs = np.array([[0, 0], [geo.n1_blocks, 0]])
    cons = pairs[:, 0] * 3 + pairs[:, 1]
    free, _, _ = DOFsInfo(geo.n_blocks, pairs)
    rate = 0.001
    leaves = dict(centroid_node_vectors=cnv, reference_vector=refv(), k_stretch=T(k_stretch), k_shear=T(1.851e-2),
                  k_rot=T(1.534e-4 / 4), damping=damping, inertia=inertia_full.reshape(-1)[free])
    run_case(name, n_blocks=geo.n_blocks, n_npb=4, bonds=bonds(), cons=cons, bond_energy=bond_energy, use_contact=False,
             drive_kind=_abi.DFX_DRIVE_ZERO, drive_vec0=None, drive_vec1=None, drive_params={},
             load=dict(kind=_abi.DFX_LOAD_RAMP, dofs=np.array([(geo.n1_blocks - 1) * 3, (geo.n_blocks - 1) * 3]), consts=(0.2, rate)),
             damped_blocks=np.arange(geo.n_blocks), leaves=leaves, block_centroids=cent,
             ts=np.linspace(0, t_end, n_t), rtol=1e-8, atol=1e-8, extra_leaf_entries=geo.n_blocks * 3,  # the inertia leaf itself
             g_mode="generic")


CASES = {
    "quads_4x3_contact_active": lambda: quad_case("quads_4x3_contact_active", 4, 3, contact_window=(20 * math.pi / 180, 60 * math.pi / 180),
                                                  noise=0.3, n_t=5, t_end=0.008, rtol=1e-8, atol=1e-4),
    "quads_5x4_tight": lambda: quad_case("quads_5x4_tight", 5, 4, contact_window=(-15 * math.pi / 180, -10 * math.pi / 180),
                                         noise=0.2, n_t=4, t_end=0.006, rtol=1e-9, atol=1e-9, seed=3, g_mode="generic"),
    "kagome_3x2_perbond": lambda: kagome_case("kagome_3x2_perbond", 3, 2, n_t=4, t_end=0.008, rtol=1e-8, atol=1e-4),
    # tight tolerance: with top and bottom rows clamped this small lattice is stiff and, at atol=1e-4, borderline
    # accept/reject decisions make the trajectory sensitive to round-off at the 1e-6 level
    "static_pulse_4x4": lambda: static_pulse_case("static_pulse_4x4", 4, 4, n_t=3, rtol=1e-9, atol=1e-8),
    "tensile_linearized": lambda: tensile_case("tensile_linearized", 2, 1, n_t=4, t_end=60.),
    "tensile_ligament": lambda: tensile_case("tensile_ligament", 2, 0, n_t=4, t_end=60.),
    "ramp_sech2_4x3": lambda: ramp_sech2_case("ramp_sech2_4x3", 4, 3, n_t=5, t_end=0.008, rtol=1e-9, atol=1e-8),
    # distance-based contact (energy.py:222-330, angle_based=False): window in length units around the hinge length 2.25
    "quads_4x3_distance_contact": lambda: quad_case("quads_4x3_distance_contact", 4, 3, contact_window=(0.5, 3.0), noise=0.3, n_t=3, t_end=0.0004,
                                                    rtol=1e-9, atol=1e-7, seed=5, g_mode="generic", use_contact="distance", initial_rotation=0.12),
    "springs_4x3": lambda: spring_case("springs_4x3", 4, 3, n_t=4, t_end=0.02, rtol=1e-8, atol=1e-6),
}

if __name__ == "__main__":
    torch.set_num_threads(1)
    for c in (sys.argv[1:] or list(CASES)):
        CASES[c]()
