"""CPU tests of the oracle (test infrastructure): it must reproduce
  * the golden vectors written by the literal torch-autograd restatement of the reference
    (oracle/ref_literal.py via oracle/make_golden.py),
  * the reference's own tests for this path (tests/test_difflexmm.py:35-146 tensile known answer,
    :149-176 frame invariance of the ligament energy),
  * finite differences (force = -dE/du; adjoint gradient = d objective / d parameter).
PARITY STATUS: unpinned against a real JAX run (jax is not installable in this image)."""

import math
import os

import numpy as np
import pytest

from cases import golden_names, load_golden, rel_l2
from difflexmm_b200 import _abi
from oracle import Oracle


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_literal_golden(name):
    c = load_golden(name)
    orc = Oracle(c.spec)
    ps = orc.params(1, c.leaves, c.per_bond, c.damping_per_dof)
    ys, st = orc.forward(ps, c.y0, c.ts, c.rtol, c.atol)
    assert st["status"][0] == 0
    # distance-based contact: the force jumps where the closest vertex / edge pair of a void changes (the gap is a min of
    # four distances, energy.py:255-275), so the step-size controller rejects steps there and the rejections depend on
    # round-off: same accepted steps, a few more or fewer rejected ones, results equal at the integration tolerance
    # (the two RHS themselves agree to 1e-15: test_distance_contact_rhs_matches_the_literal_restatement)
    nonsmooth = c.spec.contact == _abi.DFX_CONTACT_DISTANCE
    traj_tol, grad_tol = (1e-6, 1e-6) if nonsmooth else (1e-9, 1e-7)
    if nonsmooth:
        assert abs(int(st["steps"][0]) - int(c.ref["fwd_steps"])) <= 0.05 * c.ref["fwd_steps"]
        assert abs(int(st["accepted"][0]) - int(c.ref["fwd_accepted"])) <= 0.02 * c.ref["fwd_accepted"]
    else:
        assert int(st["steps"][0]) == int(c.ref["fwd_steps"]) and int(st["accepted"][0]) == int(c.ref["fwd_accepted"])
    assert rel_l2(ys[0], c.ref["ys"]) < traj_tol
    y0b, tsb, gr, sb = orc.adjoint(ps, c.ref["ys"][None], c.ts, c.g[None], c.rtol, c.atol, aug_size=c.aug_size)
    assert sb["status"][0] == 0
    # borderline accept / reject decisions flip on round-off: the count is a diagnostic, the cotangents are the check
    assert abs(int(sb["steps"][0]) - int(c.ref["bwd_steps"])) <= max(3, int((0.05 if nonsmooth else 0.005) * c.ref["bwd_steps"]))
    assert rel_l2(y0b[0], c.ref["y0_bar"]) < grad_tol
    assert rel_l2(tsb[0], c.ref["ts_bar"]) < 1e-6
    for k, v in gr.items():
        key = "grad_" + k
        if key in c.ref and np.abs(c.ref[key]).max() > 1e-9:
            assert rel_l2(v[0], c.ref[key]) < grad_tol, k


def test_distance_contact_rhs_matches_the_literal_restatement():
    """distance-based contact (reference energy.py:222-330, build_contact_energy(angle_based=False)): the closed-form
    force of the C++ oracle against torch autograd of the literal energy, at the fixture's counter-rotated start and at
    random states (point-on-edge and end-point branches both occur), and the augmented RHS (H w and every parameter
    cotangent, block_centroids included) against the literal vector-Jacobian product"""
    import torch
    from oracle import ref_literal as L
    c = load_golden("quads_4x3_distance_contact")
    orc = Oracle(c.spec)
    ps = orc.params(1, c.leaves, c.per_bond, c.damping_per_dof)
    T = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)  # noqa: E731
    lv = c.leaves
    prob = L.Problem(c.spec.n_blocks, c.spec.n_npb, c.spec.bond_nodes, c.spec.constrained_dofs, bond_energy="ligament",
                     use_contact="distance", constrained_DOFs_fn=None, damped_blocks=c.spec.damped_blocks)
    P = dict(block_centroids=T(lv["block_centroids"]), centroid_node_vectors=T(lv["centroid_node_vectors"]), k_stretch=T(lv["k_stretch"]),
             k_shear=T(lv["k_shear"]), k_rot=T(lv["k_rot"]), reference_vector=T(lv["reference_vector"]), inertia=T(lv["inertia"]),
             damping=T(lv["damping"]), min_angle=T(lv["contact"][0]), cutoff_angle=T(lv["contact"][1]), k_contact=T(lv["contact"][2]),
             constraint_params={}, loading_params={})
    rng = np.random.default_rng(0)
    nf = c.spec.n_free
    branches = set()
    for trial in range(4):
        y = c.y0.copy() if trial == 0 else c.y0 + np.concatenate([0.3 * rng.standard_normal(nf), rng.standard_normal(nf)])
        a = orc.rhs(ps, y, 0.0)
        with torch.enable_grad():
            b = prob.rhs(T(y), torch.tensor(0.0, dtype=torch.float64), P, create_graph=False).detach().numpy()
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
        # which branch the gaps take: below the hinge-to-hinge distance = a point-on-edge (or far-end) candidate won
        U = np.zeros(3 * c.spec.n_blocks)
        U[c.spec.free_dofs] = y[:nf]
        disp = T(U.reshape(-1, 3))
        cnv = P["centroid_node_vectors"]
        cur = P["block_centroids"][:, None] + cnv + L.block_to_node_kinematics(disp, cnv)[:, :, :2]
        bonds = torch.as_tensor(np.asarray(c.spec.bond_nodes), dtype=torch.int64)
        gaps = L.void_edge_distance(cur, bonds)
        flat = cur.reshape(-1, 2)
        hinge = (flat[bonds[:, 0]] - flat[bonds[:, 1]]).norm(dim=1).repeat(2)
        branches |= {"hinge"} if bool((gaps >= hinge - 1e-12).any()) else set()
        branches |= {"edge"} if bool((gaps < hinge - 1e-9).any()) else set()
        # augmented RHS: z = [y | y_bar | t0_bar | leaves], literal vjp through torch autograd
        names, leaves = L.flatten_leaves(P)
        ybar = rng.standard_normal(2 * nf)
        with torch.enable_grad():
            yy = T(y).requires_grad_(True)
            ll = [T(x).clone().requires_grad_(True) for x in leaves]
            ydot = prob.rhs(yy, torch.tensor(0.0, dtype=torch.float64), L._rebuild(P, names, ll), create_graph=True)
            grads = torch.autograd.grad(ydot, [yy] + ll, grad_outputs=T(ybar), allow_unused=True)
        lit = {n: (np.zeros(tuple(x.shape)) if gq is None else gq.numpy()) for n, gq, x in zip(names, grads[1:], ll)}
        z = np.zeros(orc.aug_size(ps))
        z[:2 * nf], z[2 * nf:4 * nf] = y, ybar
        out = orc.aug_rhs(ps, z, 0.0)
        assert np.abs(out[2 * nf:4 * nf] - grads[0].numpy()).max() <= 1e-12 * np.abs(grads[0].numpy()).max()
        # layout of the oracle's quadratures (oracle/dfx_oracle.cpp aug_layout): cnv, ref, ks, ksh, kr, damping, inertia, contact, drive, centroids
        o = 4 * nf + 1
        nn, nb = c.spec.n_blocks * c.spec.n_npb, c.spec.n_bonds
        for name, size in (("centroid_node_vectors", 2 * nn), ("reference_vector", 2 * nb), ("k_stretch", 1), ("k_shear", 1), ("k_rot", 1),
                           ("damping", 3 * len(c.spec.damped_blocks)), ("inertia", nf)):
            ref = lit[name].reshape(-1)
            if name == "damping":  # entries on constrained DOFs have no equation
                ref = ref.copy()
            assert np.abs(out[o:o + size] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), name
            o += size
        ref = np.array([lit["min_angle"], lit["cutoff_angle"], lit["k_contact"]]).reshape(-1)
        assert np.abs(out[o:o + 3] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())
        o += 3 + c.spec.n_drive_params
        ref = lit["block_centroids"].reshape(-1)
        assert np.abs(ref).max() > 0 and np.abs(out[o:o + ref.size] - ref).max() <= 1e-11 * np.abs(ref).max()
    assert branches == {"hinge", "edge"}


@pytest.mark.parametrize("lattice", ["cfg1_quads_24x16", "cfg2_kagome_20x12"])
def test_full_size_rhs_and_augmented_rhs_match_the_literal_autograd(lattice):
    """the independent derivation at FULL size: on the cfg1 / cfg3 lattice (quads 24 x 16, 712 bonds) and the cfg2 lattice
    (kagome 20 x 12), contact window moved so that contact is active, pulse drive running: the closed-form RHS and augmented
    RHS of the C++ oracle (force, H w, t_bar and every parameter cotangent) against torch autograd of the literal energy
    (oracle/ref_literal.py), at a random state"""
    import torch
    from difflexmm_b200.problems import KagomeFocusing, QuadsFocusing
    from oracle import ref_literal as L
    # contact windows that contain the rest void angles (quads 40 / 140 degrees, kagome 120 degrees)
    P = QuadsFocusing(min_angle=20 * math.pi / 180, cutoff_angle=60 * math.pi / 180) if lattice.startswith("cfg1") else \
        KagomeFocusing(min_angle=90 * math.pi / 180, cutoff_angle=150 * math.pi / 180)
    spec, drive = P.lower()
    if lattice.startswith("cfg1"):
        design = [d[0] for d in P.random_ensemble(1, noise=0.15, seed0=3)]
    else:  # the regular kagome lattice with its vertices moved a little
        g = torch.Generator().manual_seed(3)
        design = [d + 0.3 * torch.randn(d.shape, generator=g, dtype=d.dtype) for d in P.initial_design()]
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(design)
    lv = {k: v.numpy() for k, v in leaves.items()}
    orc = Oracle(spec)
    ps = orc.params(1, lv, pb, dpd)
    T = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)  # noqa: E731
    names = ("amplitude", "loading_rate", "input_delay")
    v0 = T(spec.drive_vec0)

    def cfn(t, **p):  # loading.pulse_drive written out: amplitude (1 - cos(2 pi rate tau)) / 2 inside one period
        tau = t - p["input_delay"]
        on = (tau > 0.) & (tau < p["loading_rate"] ** -1)
        return p["amplitude"] * torch.where(on, (1 - torch.cos(2 * math.pi * p["loading_rate"] * tau)) / 2, torch.zeros((), dtype=torch.float64)) * v0

    prob = L.Problem(spec.n_blocks, spec.n_npb, spec.bond_nodes, spec.constrained_dofs, bond_energy="ligament", use_contact=True,
                     constrained_DOFs_fn=cfn, damped_blocks=spec.damped_blocks)
    cp = P.control_params(design, "cpu")
    Pd = dict(block_centroids=T(cp.geometrical_params.block_centroids), centroid_node_vectors=T(lv["centroid_node_vectors"]),
              k_stretch=T(lv["k_stretch"]), k_shear=T(lv["k_shear"]), k_rot=T(lv["k_rot"]), reference_vector=T(lv["reference_vector"]),
              inertia=T(lv["inertia"]), damping=T(lv["damping"]), min_angle=T(lv["contact"][0]), cutoff_angle=T(lv["contact"][1]),
              k_contact=T(lv["contact"][2]), constraint_params={n: T(lv["drive"][i]) for i, n in enumerate(names)}, loading_params={})
    assert spec.drive_kind == _abi.DFX_DRIVE_PULSE and spec.contact == _abi.DFX_CONTACT_ANGLE
    rng = np.random.default_rng(0)
    nf = spec.n_free
    y = np.concatenate([0.05 * rng.standard_normal(nf), 5 * rng.standard_normal(nf)])
    tq = float(lv["drive"][2]) + 0.3 / float(lv["drive"][1])  # inside the pulse
    a = orc.rhs(ps, y, tq)
    lnames, lleaves = L.flatten_leaves(Pd)
    ybar = rng.standard_normal(2 * nf)
    with torch.enable_grad():
        yy, tt = T(y).requires_grad_(True), torch.tensor(tq, dtype=torch.float64, requires_grad=True)
        ll = [T(x).clone().requires_grad_(True) for x in lleaves]
        ydot = prob.rhs(yy, tt, L._rebuild(Pd, lnames, ll), create_graph=True)
        grads = torch.autograd.grad(ydot, [yy, tt] + ll, grad_outputs=T(ybar), allow_unused=True)
    b = ydot.detach().numpy()
    # the 1/x contact energy amplifies round-off near its asymptote: agreement relative to the largest entry
    assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()
    lit = {n: (np.zeros(tuple(x.shape)) if gq is None else gq.numpy()) for n, gq, x in zip(lnames, grads[2:], ll)}
    z = np.zeros(orc.aug_size(ps))
    z[:2 * nf], z[2 * nf:4 * nf] = y, ybar
    out = orc.aug_rhs(ps, z, -tq)
    close = lambda got, ref, what: (np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), what)  # noqa: E731
    checks = [close(out[2 * nf:4 * nf], grads[0].numpy(), "y_bar"), close(out[4 * nf], grads[1].item(), "t_bar")]
    o = 4 * nf + 1
    nn, nb = spec.n_blocks * spec.n_npb, spec.n_bonds
    for name, size in (("centroid_node_vectors", 2 * nn), ("reference_vector", 2 * nb), ("k_stretch", 1), ("k_shear", 1), ("k_rot", 1),
                       ("damping", 3 * len(spec.damped_blocks) if dpd else 1), ("inertia", nf)):
        checks.append(close(out[o:o + size], lit[name].reshape(-1), name))
        o += size
    checks.append(close(out[o:o + 3], np.array([lit["min_angle"], lit["cutoff_angle"], lit["k_contact"]]).reshape(-1), "contact"))
    o += 3
    checks.append(close(out[o:o + 3], np.array([lit["constraint_params." + n] for n in names]).reshape(-1), "drive"))
    assert all(ok for ok, _ in checks), [w for ok, w in checks if not ok]
    assert np.abs(lit["k_contact"]).max() > 0 and np.abs(out[o:o + 3]).max() > 0  # contact active, drive running


def _rotated_square_chain(n1_cells):
    from difflexmm_b200.geometry import RotatedSquareGeometry
    geo = RotatedSquareGeometry(n1_cells=n1_cells, n2_cells=1, spacing=1.0)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    return geo, cnvf(0.).numpy(), bonds(), refv().numpy()


@pytest.mark.parametrize("n1_cells", [5, 10, 20])
@pytest.mark.parametrize("bond_energy", [_abi.DFX_BOND_LINEARIZED, _abi.DFX_BOND_LIGAMENT])
def test_reference_tensile_known_answer(n1_cells, bond_energy):
    """reference tests/test_difflexmm.py:35-146: a chain of rotated squares clamped on the left and pulled by a
    ramped end load reaches the applied strain (rel 1e-4), both ligament energies, default 1e-8 tolerances."""
    geo, cnv, bonds, ref = _rotated_square_chain(n1_cells)
    k_stretch, mass = 1.0, 1.0
    Jrot = 1.815 ** -2 / 4 * mass * geo.spacing ** 2
    inertia = np.tile([mass, mass, Jrot], (geo.n_blocks, 1))
    damping = 0.05 * np.tile([(k_stretch * mass) ** 0.5, (k_stretch * mass) ** 0.5,
                              (k_stretch * mass) ** 0.5 * geo.spacing ** 2 / 4], (geo.n_blocks, 1))
    cons = np.array([0 * 3 + 0, geo.n1_blocks * 3 + 0])
    loaded = np.array([(geo.n1_blocks - 1) * 3, (geo.n_blocks - 1) * 3])
    rate = 0.001 * (k_stretch / mass) ** 0.5
    for strain in (0.2, 0.4, 0.6):
        final_load = strain * geo.spacing * k_stretch
        spec = _abi.TopologySpec(geo.n_blocks, 4, bonds, cons, bond_energy=bond_energy, load_kind=_abi.DFX_LOAD_RAMP,
                                 loaded_dofs=loaded, load_consts=(final_load, rate), damped_blocks=np.arange(geo.n_blocks))
        orc = Oracle(spec)
        leaves = dict(centroid_node_vectors=cnv, reference_vector=ref, k_stretch=k_stretch, k_shear=1.851e-2 * k_stretch,
                      k_rot=1.534e-4 / 4 * k_stretch * geo.spacing ** 2, damping=damping,
                      inertia=inertia.reshape(-1)[spec.free_dofs])
        ps = orc.params(1, leaves, damping_per_dof=True)
        ts = np.linspace(0, 3 / rate, 100)
        ys, st = orc.forward(ps, np.zeros(2 * spec.n_free), ts, 1e-8, 1e-8)
        assert st["status"][0] == 0
        fields = orc.expand_fields(ps, ys, ts)
        got = fields[0, -1, 0, geo.n1_blocks - 1, 0] / (geo.spacing * (geo.n1_blocks - 1))
        assert abs((got - strain) / strain) < 1e-4


def test_reference_frame_invariance_of_ligament_energy():
    """reference tests/test_difflexmm.py:149-176: two bonds between three nodes carried by one rigid rotation have
    zero strain energy (< 1e-30) for rotations in [-pi, pi]."""
    nodes = np.array([[[0., 0.], [1., 0.], [1., 1.]]])  # one rigid unit with three nodes
    spec = _abi.TopologySpec(1, 3, [[0, 1], [1, 2]], [])
    orc = Oracle(spec)
    leaves = dict(centroid_node_vectors=nodes, reference_vector=np.array([[1., 0.], [0., 1.]]), k_stretch=1., k_shear=1.,
                  k_rot=1., inertia=np.ones(3))
    ps = orc.params(1, leaves)
    for t in np.linspace(-math.pi, math.pi, 50):
        assert orc.energy(ps, np.array([0., 0., t])) < 1e-30


def _small_quads(contact_window):
    from difflexmm_b200.problems import QuadsFocusing
    P = QuadsFocusing(n1_blocks=8, n2_blocks=7, min_angle=contact_window[0], cutoff_angle=contact_window[1],
                      simulation_time=0.01, n_timepoints=6)
    spec, drive = P.lower()
    hs, vs = P.random_ensemble(1, noise=0.05, seed0=7)
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs((hs[0], vs[0]))
    return P, spec, {k: v.numpy() for k, v in leaves.items()}, pb, dpd, aug, y0.numpy(), ts.numpy()


@pytest.mark.parametrize("window", [(-15 * math.pi / 180, -10 * math.pi / 180), (20 * math.pi / 180, 60 * math.pi / 180)],
                         ids=["contact_inactive", "contact_active"])
def test_force_is_minus_energy_gradient(window):
    """finite differences of the energy (written from the energy definitions) against the analytic force in the RHS"""
    P, spec, lv, pb, dpd, aug, y0, ts = _small_quads(window)
    orc = Oracle(spec)
    ps = orc.params(1, lv, pb, dpd)
    rng = np.random.default_rng(0)
    nf = spec.n_free
    u = 0.3 * rng.standard_normal(nf)
    t = 0.0  # drive at rest: constrained DOFs are zero
    f = orc.rhs(ps, np.concatenate([u, np.zeros(nf)]), t)[nf:] * lv["inertia"]

    def E(uu):
        U = np.zeros(3 * spec.n_blocks)
        U[spec.free_dofs] = uu
        return orc.energy(ps, U)
    idx = rng.choice(nf, 25, replace=False)
    fd = np.array([-(E(u + 1e-6 * np.eye(nf)[i]) - E(u - 1e-6 * np.eye(nf)[i])) / 2e-6 for i in idx])
    assert rel_l2(f[idx], fd) < 1e-7


def test_adjoint_gradient_matches_finite_differences():
    """d(objective)/d(parameter) from the adjoint against central differences of the forward solve, tight tolerance"""
    P, spec, lv, pb, dpd, aug, y0, ts = _small_quads((-15 * math.pi / 180, -10 * math.pi / 180))
    orc = Oracle(spec)
    nf = spec.n_free
    w = np.linspace(0.5, 1.5, len(ts) * 2 * nf).reshape(len(ts), 2 * nf)
    rtol = atol = 1e-11

    def objective(lvx):
        ys, st = orc.forward(orc.params(1, lvx, pb, dpd), y0, ts, rtol, atol)
        return float((w * np.sin(ys[0])).sum()), ys
    J, ys = objective(lv)
    g = (w * np.cos(ys[0]))[None]
    y0b, tsb, gr, sb = orc.adjoint(orc.params(1, lv, pb, dpd), ys, ts, g, rtol, atol)
    rng = np.random.default_rng(1)
    for name, eps in (("centroid_node_vectors", 1e-6), ("reference_vector", 1e-6), ("k_shear", 1e-6), ("inertia", 1e-12),
                      ("damping", 1e-9), ("drive", 2e-7)):
        base = np.asarray(lv[name], dtype=np.float64)
        d = rng.standard_normal(base.shape)
        if name == "drive":  # (amplitude, rate, delay) span four decades: perturb each relative to its size
            d = d * base
        lp, lm = dict(lv), dict(lv)
        lp[name], lm[name] = base + eps * d, base - eps * d
        fd = (objective(lp)[0] - objective(lm)[0]) / (2 * eps)
        an = float((gr[name][0] * d).sum())
        tol = 2e-4 if name == "drive" else 2e-5  # the pulse delay has a large second derivative: FD truncation
        assert abs(fd - an) <= tol * max(abs(fd), abs(an)), name


def test_forward_converges_to_an_independent_scipy_integration():
    """the restated Dormand-Prince stepper (controller, FSAL, dense output) against scipy's DOP853
    with its own step control at rtol = atol = 1e-12 -- an independent integrator on the same RHS; trajectories must
    agree far below the production tolerance.  Also the augmented (adjoint) system: its cotangents against scipy
    integrating the same augmented RHS backwards over every output interval."""
    from scipy.integrate import solve_ivp
    P, spec, lv, pb, dpd, aug, y0, ts = _small_quads((-15 * math.pi / 180, -10 * math.pi / 180))
    orc = Oracle(spec)
    ps = orc.params(1, lv, pb, dpd)
    ys, st = orc.forward(ps, y0, ts, 1e-11, 1e-12)
    sol = solve_ivp(lambda t, y: orc.rhs(ps, y, t), (ts[0], ts[-1]), y0, method="DOP853", t_eval=ts, rtol=1e-12, atol=1e-12)
    assert sol.success
    assert rel_l2(ys[0], sol.y.T) < 1e-8
    # adjoint: z = (y, y_bar, t0_bar, args_bar) in negated time, restarted at every output (jax _odeint_rev)
    nf = spec.n_free
    w = np.linspace(0.5, 1.5, len(ts) * 2 * nf).reshape(len(ts), 2 * nf)
    g = w * np.cos(ys[0])
    y0b, tsb, gr, sb = orc.adjoint(ps, ys, ts, g[None], 1e-11, 1e-12)
    n_aug = orc.aug_size(ps)
    y_bar = g[-1].copy()
    tail = np.zeros(n_aug - 4 * nf)  # t0_bar followed by the parameter cotangents
    for i in range(len(ts) - 1, 0, -1):
        tail[0] -= float(orc.rhs(ps, ys[0, i], ts[i]) @ g[i])
        z0 = np.concatenate([ys[0, i], y_bar, tail])
        s = solve_ivp(lambda s_, z: orc.aug_rhs(ps, z, s_), (-ts[i], -ts[i - 1]), z0, method="DOP853", rtol=1e-12, atol=1e-14)
        assert s.success
        z1 = s.y[:, -1]
        y_bar = z1[2 * nf:4 * nf] + g[i - 1]
        tail = z1[4 * nf:]
    assert rel_l2(y0b[0], y_bar) < 1e-7
    flat = np.concatenate([np.asarray(gr[k][0]).reshape(-1) for k in gr])
    assert np.isfinite(tail).all() and abs(np.linalg.norm(tail[1:]) - np.linalg.norm(flat)) < 1e-6 * np.linalg.norm(flat)


def test_pin_against_jax_script_skips_cleanly_without_jax():
    """oracle/pin_against_jax.py regenerates and diffs every fixture with the unmodified reference on real JAX as soon as
    `import jax` works; in an image without JAX it must say so and exit 0 (it sits next to the fixtures it would pin)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "pin_against_jax.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    try:
        import jax  # noqa: F401
        assert "PINNED" in out.stdout, out.stdout
    except ImportError:
        assert "SKIP" in out.stdout
