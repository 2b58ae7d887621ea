"""The reference-facing API on the GPU: setup_dynamic_solver / solve_dynamics / ControlParams, differentiation through
the solver with torch.autograd (the analogue of jax.grad through odeint), batches of designs."""

import math

import numpy as np
import pytest
import torch

from cases import rel_l2

pytestmark = pytest.mark.gpu


def _problem(**kw):
    from difflexmm_b200.problems import QuadsFocusing
    args = dict(n1_blocks=8, n2_blocks=7, simulation_time=0.012, n_timepoints=9, target_shift=(1, 1))
    args.update(kw)
    return QuadsFocusing(**args)


def test_setup_dynamic_solver_matches_oracle_fields():
    """same call sequence as the reference's ForwardProblem.setup / forward (problems/quads_focusing.py:239-310)"""
    from difflexmm_b200.dynamics import setup_dynamic_solver
    from difflexmm_b200.energy import build_contact_energy, build_strain_energy, combine_block_energies, ligament_energy
    from difflexmm_b200.geometry import QuadGeometry
    from oracle import Oracle
    P = _problem()
    spec, drive = P.lower()
    geo = QuadGeometry(P.n1_blocks, P.n2_blocks, spacing=P.spacing, bond_length=P.bond_length)
    block_centroids, centroid_node_vectors, bond_connectivity, reference_bond_vectors = geo.get_parametrization()
    bonds = bond_connectivity()
    energy = combine_block_energies(build_strain_energy(bonds, ligament_energy), build_contact_energy(bonds))
    solve_dynamics = setup_dynamic_solver(geo, energy, constrained_block_DOF_pairs=P.constrained_block_DOF_pairs,
                                          constrained_DOFs_fn=drive, damped_blocks=np.arange(geo.n_blocks),
                                          rtol=P.rtol, atol=P.atol)
    design = P.initial_design()
    cp = P.control_params(design, "cuda")
    fields = solve_dynamics(torch.zeros(2, geo.n_blocks, 3, dtype=torch.float64), P.timepoints(), cp)
    assert fields.shape == (P.n_timepoints, 2, geo.n_blocks, 3)
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(design)
    orc = Oracle(spec)
    ps = orc.params(1, {k: v.numpy() for k, v in leaves.items()}, pb, dpd)
    ys, _ = orc.forward(ps, y0.numpy(), ts.numpy(), P.rtol, P.atol)
    ref = orc.expand_fields(ps, ys, ts.numpy())[0]
    assert rel_l2(fields.cpu().numpy(), ref) <= 1e-6


def test_distance_based_contact_through_the_api_matches_the_oracle():
    """build_contact_energy(bonds, angle_based=False) (reference energy.py:222-330, 364-407) through setup_dynamic_solver:
    fields and the gradient of a functional w.r.t. block_centroids / centroid_node_vectors / contact parameters against
    the C++ oracle (itself pinned to torch autograd of the literal energy in tests/test_oracle.py)"""
    from difflexmm_b200 import _abi
    from difflexmm_b200.dynamics import setup_dynamic_solver
    from difflexmm_b200.energy import build_contact_energy, build_strain_energy, combine_block_energies, ligament_energy
    from difflexmm_b200.geometry import QuadGeometry
    from difflexmm_b200.utils import ContactParams
    from oracle import Oracle
    P = _problem(simulation_time=0.003, n_timepoints=4, rtol=1e-9, atol=1e-8)
    spec0, drive = P.lower()
    geo = QuadGeometry(P.n1_blocks, P.n2_blocks, spacing=P.spacing, bond_length=P.bond_length)
    bonds = geo.get_parametrization()[2]()
    energy = combine_block_energies(build_strain_energy(bonds, ligament_energy), build_contact_energy(bonds, angle_based=False))
    assert energy.contact == _abi.DFX_CONTACT_DISTANCE
    solve_dynamics = setup_dynamic_solver(geo, energy, constrained_block_DOF_pairs=P.constrained_block_DOF_pairs,
                                          constrained_DOFs_fn=drive, damped_blocks=np.arange(geo.n_blocks), rtol=P.rtol, atol=P.atol)
    design = P.initial_design()
    cp = P.control_params(design, "cuda")
    # contact window in length units around the hinge length (the gap never exceeds it): contact always active
    window = ContactParams(min_angle=torch.tensor(0.25 * P.bond_length, dtype=torch.float64),
                           cutoff_angle=torch.tensor(1.5 * P.bond_length, dtype=torch.float64),
                           k_contact=cp.mechanical_params.contact_params.k_contact)
    cen = cp.geometrical_params.block_centroids.clone().requires_grad_(True)
    cnv = cp.geometrical_params.centroid_node_vectors.clone().requires_grad_(True)
    kc = torch.tensor(float(window.k_contact), dtype=torch.float64, device="cuda", requires_grad=True)
    cp = cp._replace(geometrical_params=cp.geometrical_params._replace(block_centroids=cen, centroid_node_vectors=cnv),
                     mechanical_params=cp.mechanical_params._replace(contact_params=window._replace(k_contact=kc)))
    fields = solve_dynamics(torch.zeros(2, geo.n_blocks, 3, dtype=torch.float64), P.timepoints(), cp)
    w = torch.linspace(0.5, 1.5, fields.numel(), dtype=torch.float64, device="cuda").reshape(fields.shape)
    (w * torch.sin(fields)).sum().backward()
    assert cen.grad is not None and float(cen.grad.abs().max()) > 0
    # the same solve on the oracle, at the solver boundary
    solver = solve_dynamics.solver if hasattr(solve_dynamics, "solver") else solve_dynamics
    spec = solver.spec
    assert spec.contact == _abi.DFX_CONTACT_DISTANCE
    from difflexmm_b200.dynamics import lower_params
    leaves, pb, dpd, aug = lower_params(spec, drive, cp, None, "cuda", inertia_full=None)
    orc = Oracle(spec)
    lv = {k: v.detach().cpu().numpy() for k, v in leaves.items()}
    ps = orc.params(1, lv, pb, dpd)
    ts = P.timepoints().numpy()
    ys, st = orc.forward(ps, np.zeros(2 * spec.n_free), ts, P.rtol, P.atol)
    ref = orc.expand_fields(ps, ys, ts)[0]
    assert rel_l2(fields.detach().cpu().numpy(), ref) <= 1e-6
    # cotangent of the free-DOF trajectory: constrained DOFs follow the drive, whose parameters are not differentiated here
    gf = (w * torch.cos(fields)).detach().cpu().numpy()
    free = np.asarray(spec.free_dofs)
    g = np.concatenate([gf[:, 0].reshape(len(ts), -1)[:, free], gf[:, 1].reshape(len(ts), -1)[:, free]], axis=1)
    _, _, gr, sb = orc.adjoint(ps, ys, ts, g[None], P.rtol, P.atol, aug)
    assert rel_l2(cen.grad.cpu().numpy(), gr["block_centroids"][0]) <= 1e-5
    assert abs(kc.grad.item() - gr["contact"][0][2]) <= 1e-5 * abs(gr["contact"][0][2])


def test_design_gradient_through_the_solver_matches_finite_differences():
    """value_and_grad(target kinetic energy)(design) -- the hot call of the reference's optimisation loop
    (problems/quads_focusing.py:565-569) -- against central differences of the forward solve"""
    P = _problem(rtol=1e-10, atol=1e-10)
    P.setup()
    hs, vs = P.initial_design()
    hs = hs.clone().requires_grad_(True)
    vs = vs.clone().requires_grad_(True)
    J = P.target_kinetic_energy((hs, vs))
    J.backward()
    assert torch.isfinite(hs.grad).all() and torch.isfinite(vs.grad).all()
    rng = np.random.default_rng(0)
    d_hs, d_vs = torch.from_numpy(rng.standard_normal(hs.shape)), torch.from_numpy(rng.standard_normal(vs.shape))
    eps = 1e-5
    with torch.no_grad():
        Jp = P.target_kinetic_energy((hs + eps * d_hs, vs + eps * d_vs))
        Jm = P.target_kinetic_energy((hs - eps * d_hs, vs - eps * d_vs))
    fd = ((Jp - Jm) / (2 * eps)).item()
    an = ((hs.grad * d_hs).sum() + (vs.grad * d_vs).sum()).item()
    assert abs(fd - an) <= 1e-5 * max(abs(fd), abs(an))


def test_batch_of_designs_equals_one_by_one():
    P = _problem()
    P.setup()
    hs, vs = P.random_ensemble(4, noise=0.05)
    Jb = P.target_kinetic_energy((hs, vs), batch=4)
    J1 = torch.stack([P.target_kinetic_energy((hs[i], vs[i])) for i in range(4)])
    assert torch.allclose(Jb, J1, rtol=1e-12)
    st = P.solver.last_forward_stats.numpy()
    assert (st["status"] == 0).all() and (st["steps"] > 0).all()


def test_launch_order_history_does_not_change_results():
    """the solver orders the forward / adjoint launches of a batch it has evaluated before by that evaluation's step counts
    (DynamicSolver.adjoint_launch_options, lib_forward): a scheduling hint only -- values and gradients of the second
    evaluation are bitwise those of the first, and the hint is really taken (more designs than SMs / forward slots)"""
    P = _problem(simulation_time=0.004, n_timepoints=4)
    s = P.setup()
    B = 300
    hs, vs = P.random_ensemble(B, noise=0.05)
    out = []
    for rep in range(2):
        d = [hs.clone().cuda().requires_grad_(True), vs.clone().cuda().requires_grad_(True)]
        if rep == 1:
            opt_a, order_a = s.adjoint_launch_options(B)
            assert order_a is not None and order_a.numel() == B and opt_a.design_order
            prev = s.last_adjoint_stats.steps_device()
            assert bool((prev[order_a.long()][:-1] >= prev[order_a.long()][1:]).all())  # longest first by the previous adjoint
        J = P.target_kinetic_energy(d, batch=B, fused=True)
        J.sum().backward()
        out.append((J.detach().clone(), d[0].grad.clone(), d[1].grad.clone()))
    assert (s.last_adjoint_stats.numpy()["status"] == 0).all()
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)
    # another batch size: no history to use
    assert s.adjoint_launch_options(B + 1)[1] is None


def test_expand_fields_kernel_matches_torch_postprocessing():
    from difflexmm_b200 import _abi
    P = _problem()
    s = P.setup()
    design = P.initial_design()
    cp = P.control_params(design, "cuda")
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(design, device="cuda")
    ps = _abi.ParamSet(P.spec, 1, {k: v.contiguous() for k, v in leaves.items()}, pb, dpd)
    ys, _ = s.lib_forward(ps, y0, ts)
    a = s._lib.expand_fields(s.handle, ps, ys, ts)
    b = s.expand_fields(ys, ts, cp)
    assert torch.allclose(a, b, rtol=1e-13, atol=1e-13)


def test_lattice_beyond_the_fast_kernels_uses_the_generic_path():
    """26 x 22 quads (572 units > 512 threads): generic forward / adjoint kernels with state spilled to the
    L2-resident scratch (the cfg5 route), against the C++ oracle"""
    from difflexmm_b200 import _abi
    from oracle import Oracle
    P = _problem(n1_blocks=26, n2_blocks=22, simulation_time=0.004, n_timepoints=4, target_shift=(2, 2))
    s = P.setup()
    design = P.initial_design()
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(design, device="cuda")
    ps = _abi.ParamSet(P.spec, 1, {k: v.contiguous() for k, v in leaves.items()}, pb, dpd)
    ys, st = s.lib_forward(ps, y0, ts)
    assert st.numpy()["status"][0] == 0
    orc = Oracle(P.spec)
    ph = orc.params(1, {k: v.cpu().numpy() for k, v in leaves.items()}, pb, dpd)
    ys_h, st_h = orc.forward(ph, y0.cpu().numpy(), ts.cpu().numpy(), P.rtol, P.atol)
    assert rel_l2(ys[0].cpu().numpy(), ys_h[0]) <= 1e-6
    g = np.cos(ys_h) + 0.2
    y0b_h, tsb_h, gr_h, _ = orc.adjoint(ph, ys_h, ts.cpu().numpy(), g, P.rtol, P.atol, aug)
    y0b, tsb, gr, sb = s.lib_adjoint(ps, torch.as_tensor(ys_h, device="cuda"), ts, torch.as_tensor(g, device="cuda"), aug)
    assert sb.numpy()["status"][0] == 0
    for k in gr_h:
        if np.abs(gr_h[k]).max() > 1e-9:
            assert rel_l2(gr[k][0].cpu().numpy(), gr_h[k][0]) <= 1e-5, k


def test_static_tuning_multitask_objective():
    """cfg4 recipe (static pre-compression ramp + delayed pulse, two tasks with weights 0.75 / -0.25, summed design
    gradient; reference problems/quads_kinetic_energy_static_tuning.py:431-484) on a small lattice: the weighted
    multi-task gradient equals the gradient of the weighted sum, and one task matches the C++ oracle"""
    from difflexmm_b200 import _abi
    from difflexmm_b200.parallel import multitask_value_and_grad
    from difflexmm_b200.problems import QuadsStaticTuning
    from oracle import Oracle
    base = dict(n1_blocks=8, n2_blocks=8, simulation_time_dynamic=0.008, n_timepoints=6, target_shift=(1, 1),
                compressive_strain_rate=25.0)
    tasks = [dict(compressive_strain=0.01), dict(compressive_strain=0.03)]
    weights = [0.75, -0.25]
    probs = [QuadsStaticTuning(**base, **t) for t in tasks]
    for p in probs:
        p.setup()
    hs, vs = probs[0].initial_design()

    def task_vg(design, p, weight):
        h = design[0].clone().requires_grad_(True)
        v = design[1].clone().requires_grad_(True)
        J = weight * p.target_kinetic_energy((h, v))
        J.backward()
        return J.detach(), [h.grad, v.grad]

    design = [hs.cuda(), vs.cuda()]
    J, grads = multitask_value_and_grad(task_vg, design, probs, weights)
    h = design[0].clone().requires_grad_(True)
    v = design[1].clone().requires_grad_(True)
    Jsum = sum(w * p.target_kinetic_energy((h, v)) for w, p in zip(weights, probs))
    Jsum.backward()
    assert torch.allclose(J, Jsum.detach(), rtol=1e-12)
    for a, b in zip(grads, (h.grad, v.grad)):
        assert (a - b).abs().max() <= 1e-10 * b.abs().max()
    # one task against the oracle at the boundary
    p = probs[1]
    leaves, pb, dpd, aug, y0, ts = p.boundary_inputs((hs, vs))
    orc = Oracle(p.spec)
    ph = orc.params(1, {k: x.numpy() for k, x in leaves.items()}, pb, dpd)
    ys_h, _ = orc.forward(ph, y0.numpy(), ts.numpy(), p.rtol, p.atol)
    dl = {k: x.cuda().contiguous() for k, x in leaves.items()}
    ys_d, st = p.solver.lib_forward(_abi.ParamSet(p.spec, 1, dl, pb, dpd), y0.cuda(), ts.cuda())
    assert st.numpy()["status"][0] == 0
    assert rel_l2(ys_d[0].cpu().numpy(), ys_h[0]) <= 1e-6


def test_cfg5_lattice_100x100_with_active_contact():
    """cfg5 of BASELINE.json: quads 100 x 100 (10^4 units, 19 800 bonds, ~30k free DOFs) with the contact window
    moved to [+15, +25] degrees so that contact is active (rest void angles are 40 / 140 degrees);
    short horizon so that the C++ oracle finishes in seconds.  Generic kernels (state in the L2-resident scratch)."""
    from difflexmm_b200 import _abi
    from oracle import Oracle
    P = _problem(n1_blocks=100, n2_blocks=100, simulation_time=0.006, n_timepoints=3, target_shift=(2, 2),
                 min_angle=15 * math.pi / 180, cutoff_angle=45 * math.pi / 180)
    s = P.setup()
    design = P.initial_design()
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(design, device="cuda")
    ps = _abi.ParamSet(P.spec, 1, {k: v.contiguous() for k, v in leaves.items()}, pb, dpd)
    ys, st = s.lib_forward(ps, y0, ts)
    assert st.numpy()["status"][0] == 0
    orc = Oracle(P.spec)
    ph = orc.params(1, {k: v.cpu().numpy() for k, v in leaves.items()}, pb, dpd)
    ys_h, st_h = orc.forward(ph, y0.cpu().numpy(), ts.cpu().numpy(), P.rtol, P.atol)
    assert int(st.numpy()["steps"][0]) == int(st_h["steps"][0])
    assert rel_l2(ys[0].cpu().numpy(), ys_h[0]) <= 1e-6
    nf = P.spec.n_free
    g = np.zeros_like(ys_h)
    g[:, :, nf:] = ys_h[:, :, nf:] * leaves["inertia"].cpu().numpy()
    y0b_h, tsb_h, gr_h, sb_h = orc.adjoint(ph, ys_h, ts.cpu().numpy(), g, P.rtol, P.atol, aug)
    y0b, tsb, gr, sb = s.lib_adjoint(ps, torch.as_tensor(ys_h, device="cuda"), ts, torch.as_tensor(g, device="cuda"), aug)
    assert sb.numpy()["status"][0] == 0
    assert np.abs(gr_h["contact"]).max() > 0  # contact really is active
    for k in gr_h:
        if np.abs(gr_h[k]).max() > 1e-9:
            assert rel_l2(gr[k][0].cpu().numpy(), gr_h[k][0]) <= 1e-5, k


def test_tabulated_drive_matches_oracle():
    """measured-input style drive (jnp.interp table) through libdfx vs the C++ oracle, forward and adjoint"""
    from difflexmm_b200 import _abi, _lib
    from difflexmm_b200.dynamics import lower_topology
    from difflexmm_b200.loading import tabulated_drive
    from oracle import Oracle
    P = _problem()
    spec0, drive0 = P.lower()
    times = np.linspace(0.0, 0.012, 25)
    values = 7.5 * (1 - np.cos(2 * np.pi * 30 * np.clip(times - 0.001, 0, None))) / 2
    drive = tabulated_drive(times, values, drive0.vec0)
    spec, _ = lower_topology(P.geometry, P.energy(P.geometry.bond_connectivity()), None, None,
                             P.constrained_block_DOF_pairs, drive, np.arange(P.geometry.n_blocks))
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(P.initial_design())
    leaves = {k: v for k, v in leaves.items() if k != "drive"}
    orc = Oracle(spec)
    ph = orc.params(1, {k: v.numpy() for k, v in leaves.items()}, pb, dpd)
    ys_h, st_h = orc.forward(ph, y0.numpy(), ts.numpy(), P.rtol, P.atol)
    g = np.sin(ys_h) + 0.1
    y0b_h, tsb_h, gr_h, _ = orc.adjoint(ph, ys_h, ts.numpy(), g, P.rtol, P.atol)
    topo = _lib.Topology(spec, torch.cuda.current_device())
    ps = _abi.ParamSet(spec, 1, {k: v.cuda().contiguous() for k, v in leaves.items()}, pb, dpd)
    opt = _abi.DfxOptions(0, 0, 0)
    ys, st = _lib.forward(topo, ps, y0.cuda(), ts.cuda(), P.rtol, P.atol, opt)
    assert st.numpy()["status"][0] == 0 and np.abs(ys_h).max() > 0
    assert rel_l2(ys[0].cpu().numpy(), ys_h[0]) <= 1e-6
    y0b, tsb, gr, sb = _lib.adjoint(topo, ps, torch.as_tensor(ys_h, device="cuda"), ts.cuda(), torch.as_tensor(g, device="cuda"),
                                    P.rtol, P.atol, 0, opt)
    assert sb.numpy()["status"][0] == 0
    # ts_bar[1:] are the t_bar terms of the interval starts; ts_bar[0] = t0_bar is the quadrature of an integrand that
    # jumps at every knot of the table (d drive / dt is piecewise constant), so it is only first-order accurate in the
    # step size and moves with the accept / reject sequence: it is compared on the scale of the vector
    tsb_d = tsb[0].cpu().numpy()
    assert rel_l2(tsb_d[1:], tsb_h[0][1:]) <= 1e-5
    assert abs(tsb_d[0] - tsb_h[0][0]) <= 1e-4 * np.abs(tsb_h[0]).max()
    for k in gr_h:
        if np.abs(gr_h[k]).max() > 1e-9:
            assert rel_l2(gr[k][0].cpu().numpy(), gr_h[k][0]) <= 1e-5, k


def test_status_flags_and_per_design_inputs():
    """per-design y0 / ts (batch strides), max_steps -> DFX_STATUS_MAX_STEPS with NaN outputs, init_step_variant switch"""
    from difflexmm_b200 import _abi, _lib
    from oracle import Oracle
    P = _problem()
    P.lower()
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(P.initial_design())
    spec = P.spec
    topo = _lib.Topology(spec, torch.cuda.current_device())
    B = 3
    dl = {k: v.cuda().contiguous() for k, v in leaves.items()}
    ps = _abi.ParamSet(spec, B, dl, pb, dpd)
    rng = np.random.default_rng(0)
    y0b = 1e-3 * rng.standard_normal((B, 2 * spec.n_free))
    tsb = np.stack([ts.numpy() * s for s in (1.0, 0.8, 0.5)])
    ys, st = _lib.forward(topo, ps, torch.as_tensor(y0b, device="cuda"), torch.as_tensor(tsb, device="cuda"), P.rtol, P.atol,
                          _abi.DfxOptions(0, 0, 0))
    assert (st.numpy()["status"] == 0).all()
    orc = Oracle(spec)
    ph = orc.params(B, {k: v.numpy() for k, v in leaves.items()}, pb, dpd)
    ys_h, _ = orc.forward(ph, y0b, tsb, P.rtol, P.atol)
    for b in range(B):
        assert rel_l2(ys[b].cpu().numpy(), ys_h[b]) <= 1e-6
    # later jax releases use max(d1, d2) in initial_step_size: a different, equally valid step sequence
    ys1, st1 = _lib.forward(topo, ps, torch.as_tensor(y0b, device="cuda"), torch.as_tensor(tsb, device="cuda"), P.rtol, P.atol,
                            _abi.DfxOptions(1, 0, 0))
    ys1_h, _ = orc.forward(ph, y0b, tsb, P.rtol, P.atol, variant=1)
    assert rel_l2(ys1.cpu().numpy(), ys1_h) <= 1e-6
    # a step budget that cannot reach the first output: status flag + NaN outputs, never garbage
    ys2, st2 = _lib.forward(topo, ps, torch.as_tensor(y0b, device="cuda"), torch.as_tensor(tsb, device="cuda"), P.rtol, P.atol,
                            _abi.DfxOptions(0, 0, 3))
    s2 = st2.numpy()
    assert (s2["status"] & _abi.DFX_STATUS_MAX_STEPS).all() and (s2["steps"] == 3).all()
    assert torch.isnan(ys2[:, 1:]).all() and torch.isfinite(ys2[:, 0]).all()


@pytest.mark.parametrize("kernel", ["fast", "generic"])
def test_fused_kinetic_objective_equals_the_unfused_path(kernel, monkeypatch):
    """dfx_forward + dfx_objective + dfx_adjoint_objective (cotangent formed inside the adjoint kernel)
    against the torch objective on the expanded fields + dfx_adjoint with a materialised cotangent; batch of designs
    with non-uniform upstream weights; the generic adjoint kernel takes the cotangent-materialising route."""
    if kernel == "generic":
        monkeypatch.setenv("DFX_ADJOINT_KERNEL", "generic")
        monkeypatch.setenv("DFX_FORWARD_KERNEL", "generic")
    P = _problem()
    P.setup()
    B = 3
    hs0, vs0 = P.random_ensemble(B, noise=0.05)
    w = torch.tensor([1.0, -0.5, 2.0], dtype=torch.float64, device="cuda")
    res = []
    for fused in (False, True):
        hs, vs = hs0.clone().requires_grad_(True), vs0.clone().requires_grad_(True)
        J = P.target_kinetic_energy((hs, vs), batch=B, fused=fused)
        assert J.shape == (B,)
        (J * w.to(J.device)).sum().backward()
        assert (P.solver.last_adjoint_stats.numpy()["status"] == 0).all()
        res.append((J.detach().cpu(), hs.grad.cpu(), vs.grad.cpu()))
    (J0, gh0, gv0), (J1, gh1, gv1) = res
    assert torch.allclose(J0, J1, rtol=1e-12, atol=0)
    assert rel_l2(gh1.numpy(), gh0.numpy()) <= 1e-10 and rel_l2(gv1.numpy(), gv0.numpy()) <= 1e-10
    # unbatched call returns a scalar
    J_s = P.target_kinetic_energy((hs0[0], vs0[0]), fused=True)
    assert J_s.dim() == 0 and abs(J_s.item() - J0[0].item()) <= 1e-12 * abs(J0[0].item())


def test_fused_objective_rejects_constrained_targets():
    P = _problem()
    s = P.setup()
    with pytest.raises(ValueError):
        s.target_free_ids(np.array([int(np.asarray(P.spec.constrained_dofs)[0]) // 3]))


@pytest.mark.parametrize("lattice", ["quads", "kagome"])
def test_device_geometry_matches_the_torch_design_maps(lattice):
    """dfx_geometry_forward / dfx_geometry_vjp against the differentiable torch restatement of the
    reference's design maps (geometry.py:607-952) and compute_inertia (geometry.py:144-160): values and VJP."""
    from difflexmm_b200.geometry import KagomeGeometry, QuadGeometry, compute_inertia
    from difflexmm_b200.geometry_device import DeviceGeometry
    rng = np.random.default_rng(3)
    if lattice == "quads":
        geo = QuadGeometry(7, 5, spacing=15.0, bond_length=2.25)
        geo.compute_geometry()
        base = geo.get_design_from_rotated_square(25 * math.pi / 180)
    else:
        geo = KagomeGeometry(5, 4, direct_basis=20.0 * np.array([[1, 0], [math.cos(math.pi / 3), math.sin(math.pi / 3)]]),
                             bond_length=2.25)
        geo.compute_geometry()
        base = [torch.zeros(s, dtype=torch.float64) for s in geo.design_shapes]
    B, rho = 3, 6.18e-9
    designs = [torch.stack([b + torch.from_numpy(rng.uniform(-1, 1, b.shape)) for _ in range(B)]) for b in base]
    dg = DeviceGeometry(geo, "cuda")
    # reference: torch maps, one design at a time, on the CPU
    d_ref = [d.clone().requires_grad_(True) for d in designs]
    cnv_r = torch.stack([geo.centroid_node_vectors(*[d[i] for d in d_ref]) for i in range(B)])
    cen_r = torch.stack([geo.block_centroids(*[d[i] for d in d_ref]) for i in range(B)])
    ine_r = compute_inertia(cnv_r, torch.tensor(rho, dtype=torch.float64))
    d_dev = [d.clone().cuda().requires_grad_(True) for d in designs]
    cnv, cen, ine = dg(d_dev, rho)
    assert rel_l2(cnv.detach().cpu().numpy(), cnv_r.detach().numpy()) <= 1e-14
    assert rel_l2(cen.detach().cpu().numpy(), cen_r.detach().numpy()) <= 1e-14
    assert rel_l2(ine.detach().cpu().numpy(), ine_r.detach().numpy()) <= 1e-13
    w_cnv, w_cen = torch.from_numpy(rng.standard_normal(cnv_r.shape)), torch.from_numpy(rng.standard_normal(cen_r.shape))
    w_ine = torch.from_numpy(rng.standard_normal(ine_r.shape)) / ine_r.detach().abs()
    ((cnv_r * w_cnv).sum() + (cen_r * w_cen).sum() + (ine_r * w_ine).sum()).backward()
    ((cnv * w_cnv.cuda()).sum() + (cen * w_cen.cuda()).sum() + (ine * w_ine.cuda()).sum()).backward()
    for a, b in zip(d_dev, d_ref):
        assert rel_l2(a.grad.cpu().numpy(), b.grad.numpy()) <= 1e-12
    # unbatched call
    cnv1, cen1, ine1 = dg([d[0] for d in designs], rho)
    assert cnv1.shape == cnv_r.shape[1:] and torch.equal(cnv1, cnv[0].detach())


@pytest.mark.parametrize("lattice", ["quads", "quads_boundary", "kagome"])
def test_device_constraints_match_the_torch_constraints_and_their_jacobian(lattice):
    """dfx_constraints_eval (values + fixed-width sparse Jacobian, one launch per batch) against the torch
    restatement of the reference's angle / edge-length constraints (problems/quads_focusing.py:473-544) and its autograd
    Jacobian (the reference: jit(jacobian(...)), :585-588, :613-616), to 1e-12"""
    from difflexmm_b200.geometry import KagomeGeometry, QuadGeometry
    from difflexmm_b200.geometry_device import DeviceConstraints, DeviceGeometry
    from difflexmm_b200.optimization import angle_constraints, edge_length_constraints
    rng = np.random.default_rng(11)
    if lattice.startswith("quads"):
        geo = QuadGeometry(6, 4, spacing=15.0, bond_length=2.25)
        geo.compute_geometry()
        base = geo.get_design_from_rotated_square(25 * math.pi / 180)
    else:
        geo = KagomeGeometry(4, 3, direct_basis=20.0 * np.array([[1, 0], [math.cos(math.pi / 3), math.sin(math.pi / 3)]]),
                             bond_length=2.25)
        geo.compute_geometry()
        base = [torch.zeros(s, dtype=torch.float64) for s in geo.design_shapes]
    boundary = lattice == "quads_boundary"
    min_void, min_block, min_edge = 5 * math.pi / 180, 30 * math.pi / 180, 3.0
    B = 3
    designs = [torch.stack([b + torch.from_numpy(rng.uniform(-0.8, 0.8, b.shape)) for _ in range(B)]) for b in base]
    dg = DeviceGeometry(geo, "cuda")
    dc = DeviceConstraints(dg, min_void, min_block, min_edge, boundary_angle_constraint=boundary)
    flat, _ = dg.flatten(designs)
    c, jac = dc(flat)
    dense = dc.dense_jacobian(jac).cpu()
    ang, edg = angle_constraints(geo, min_void, min_block, boundary), edge_length_constraints(geo, min_edge)
    n_bonds = len(np.asarray(geo.bond_connectivity()))
    assert dc.n_angle_rows == 4 * n_bonds + (2 * (geo.n1_blocks + geo.n2_blocks) if boundary else 0)
    assert dc.n_rows == dc.n_angle_rows + geo.n_blocks * geo.n_npb
    sizes = [int(np.prod(s)) for s in geo.design_shapes]
    for b in range(B):
        def both(x):
            parts, o = [], 0
            for s, n in zip(geo.design_shapes, sizes):
                parts.append(x[o:o + n].reshape(tuple(s)))
                o += n
            return torch.cat([ang(parts), edg(parts)])
        x = flat[b].reshape(-1).cpu()
        ref = both(x)
        J_ref = torch.autograd.functional.jacobian(both, x)
        assert (c[b].cpu() - ref).abs().max() <= 1e-12 * max(1.0, float(ref.abs().max()))
        assert (dense[b] - J_ref).abs().max() <= 1e-12 * max(1.0, float(J_ref.abs().max()))
    # values only
    c2, none = dc(flat, want_jacobian=False)
    assert none is None and torch.equal(c2, c)


def test_constrained_batched_mma_gives_a_feasible_ascent():
    """run_optimization_mma with min_void_angle / min_block_angle / min_edge_length set (the switches of the
    reference's run_optimization_nlopt, problems/quads_focusing.py:546-652): every best design is feasible to the
    reference's tolerance (1e-8), no worse than its feasible start, and the constraints are active for some instance"""
    from difflexmm_b200.optimization import OptimizationProblem
    P = _problem()
    P.setup()
    B = 3
    guesses = P.random_ensemble(B, noise=0.03)
    opt = OptimizationProblem(P)
    x0 = opt.flatten(guesses).cuda()
    dc0 = opt.device_constraints(0.0, 0.0, 0.0)
    bounds = dict(lower_bound=float(x0.min()) - 1.0, upper_bound=float(x0.max()) + 1.0)

    def family_minima(x):  # smallest void angle, block angle, edge length over the batch
        c, _ = dc0(x)
        na = dc0.n_angle_rows
        return float((-c[:, :na // 2]).min()), float((-c[:, na // 2:na]).min()), float((-c[:, na:]).min())

    # the unconstrained run tells which quantities the objective wants to shrink; thresholds halfway between the start
    # and that design make the start feasible and the unconstrained optimum infeasible
    opt_free = OptimizationProblem(P)
    best_free, best_f_free = opt_free.run_optimization_mma([g.cuda() for g in guesses], n_iterations=8, **bounds)
    start, free = family_minima(x0), family_minima(opt_free.flatten(best_free).cuda())
    assert any(f < s * (1 - 1e-6) for s, f in zip(start, free))
    void_t, block_t, edge_t = [0.5 * (s + f) if f < s * (1 - 1e-6) else 0.97 * s for s, f in zip(start, free)]
    mins = dict(min_void_angle=void_t, min_block_angle=block_t, min_edge_length=edge_t)
    dc = opt.device_constraints(**mins)
    assert float(dc(x0)[0].max()) < 0 < float(dc(opt_free.flatten(best_free).cuda())[0].max())
    J0, _ = opt.objective_and_grad(x0)
    best, best_f = opt.run_optimization_mma([g.cuda() for g in guesses], n_iterations=8, **bounds, **mins)
    assert len(opt.objective_values) == 8 and len(opt.constraints_violation["angles"]) >= 8
    assert len(opt.constraints_violation["edge_lengths"]) >= 8
    c_best, _ = dc(opt.flatten(best).cuda())
    assert float(c_best.max()) <= 1e-8
    assert (opt.optimizer.best_violation <= 1e-8).all()
    assert (best_f >= J0 * (1 - 1e-12)).all() and (best_f > J0 * 1.001).any()


def test_device_rotated_square_map_matches_the_torch_design_map():
    """rest of SURVEY 8 f1: RotatedSquareGeometry (reference geometry.py:354-443), design = one angle per lattice, on the
    device (dfx_rotated_square_forward / _vjp) against the torch map + compute_inertia: values 1e-14, VJP 1e-12"""
    from difflexmm_b200.geometry import RotatedSquareGeometry, compute_inertia
    from difflexmm_b200.geometry_device import DeviceRotatedSquare
    geo = RotatedSquareGeometry(3, 2, spacing=15.0, bond_length=2.25)
    geo.compute_geometry()
    rho = 6.18e-9
    angles = torch.tensor([0.35, -0.2, 0.05], dtype=torch.float64)
    a_ref = angles.clone().requires_grad_(True)
    cnv_r = torch.stack([geo.centroid_node_vectors(a_ref[i]) for i in range(3)])
    ine_r = compute_inertia(cnv_r, torch.tensor(rho, dtype=torch.float64))
    drs = DeviceRotatedSquare(geo, "cuda")
    a_dev = angles.clone().cuda().requires_grad_(True)
    rho_dev = torch.tensor(rho, dtype=torch.float64, device="cuda", requires_grad=True)
    cnv, cen, ine = drs(a_dev, rho_dev)
    assert cnv.shape == cnv_r.shape and rel_l2(cnv.detach().cpu().numpy(), cnv_r.detach().numpy()) <= 1e-14
    assert rel_l2(ine.detach().cpu().numpy(), ine_r.detach().numpy()) <= 1e-13
    assert torch.equal(cen.cpu(), geo.block_centroids())
    rng = np.random.default_rng(2)
    w_cnv = torch.from_numpy(rng.standard_normal(cnv_r.shape))
    w_ine = torch.from_numpy(rng.standard_normal(ine_r.shape)) / ine_r.detach().abs()
    ((cnv_r * w_cnv).sum() + (ine_r * w_ine).sum()).backward()
    ((cnv * w_cnv.cuda()).sum() + (ine * w_ine.cuda()).sum()).backward()
    assert rel_l2(a_dev.grad.cpu().numpy(), a_ref.grad.numpy()) <= 1e-12
    assert abs(rho_dev.grad.item() - float((ine_r.detach() * w_ine).sum() / rho)) <= 1e-10 * abs(rho_dev.grad.item())
    cnv1, _, ine1 = drs(angles[0], rho)  # unbatched call
    assert cnv1.shape == cnv_r.shape[1:] and torch.equal(cnv1, cnv[0].detach())


def test_batched_mma_improves_an_ensemble_of_designs():
    """several MMA instances advanced in lock-step, one batched forward + adjoint per iteration"""
    from difflexmm_b200.optimization import OptimizationProblem
    P = _problem()
    P.setup()
    B = 3
    guesses = P.random_ensemble(B, noise=0.03)
    opt = OptimizationProblem(P)
    x0 = opt.flatten(guesses)
    J0, g0 = opt.objective_and_grad(x0.cuda())
    assert J0.shape == (B,) and g0.shape == x0.shape
    best, best_f = opt.run_optimization_mma([g.cuda() for g in guesses], n_iterations=6,
                                            lower_bound=float(x0.min()) - 1.0, upper_bound=float(x0.max()) + 1.0)
    assert len(opt.objective_values) == 6
    assert (best_f >= J0 * (1 - 1e-12)).all() and (best_f > J0 * 1.001).any()
    sol = opt.compute_best_forward()
    assert sol.fields.shape == (P.n_timepoints, 2, P.geometry.n_blocks, 3)


def test_angular_momentum_objective_gradient_matches_finite_differences():
    """the spin objective of the reference (problems/quads_spin.py:395-428, energy.py:502-519): its cotangent has
    displacement AND velocity parts, differentiated through dfx_adjoint"""
    P = _problem(rtol=1e-10, atol=1e-10)
    P.setup()
    hs, vs = P.initial_design()
    tb = P.target_blocks()
    center = P.geometry.block_centroids(hs, vs)[tb].mean(0)
    hs = hs.clone().requires_grad_(True)
    vs = vs.clone().requires_grad_(True)
    L = P.target_angular_momentum((hs, vs), spin_center=center)
    L.backward()
    rng = np.random.default_rng(5)
    d_hs, d_vs = torch.from_numpy(rng.standard_normal(hs.shape)), torch.from_numpy(rng.standard_normal(vs.shape))
    eps = 1e-5
    with torch.no_grad():
        Lp = P.target_angular_momentum((hs + eps * d_hs, vs + eps * d_vs), spin_center=center)
        Lm = P.target_angular_momentum((hs - eps * d_hs, vs - eps * d_vs), spin_center=center)
    fd = ((Lp - Lm) / (2 * eps)).item()
    an = ((hs.grad * d_hs).sum() + (vs.grad * d_vs).sum()).item()
    assert abs(fd - an) <= 2e-5 * max(abs(fd), abs(an)), (fd, an)


def test_zero_length_spring_energy_through_the_api():
    """stretching_torsional_spring_energy + StretchingTorsionalSpringParams (reference energy.py:49-66,
    utils.py:80-91) through setup_dynamic_solver, against the C++ oracle; gradient w.r.t. both stiffnesses"""
    from difflexmm_b200 import _abi
    from difflexmm_b200.dynamics import lower_params, setup_dynamic_solver
    from difflexmm_b200.energy import build_strain_energy, stretching_torsional_spring_energy
    from difflexmm_b200.geometry import QuadGeometry
    from difflexmm_b200.loading import pulse_drive
    from difflexmm_b200.utils import (ControlParams, GeometricalParams, LigamentParams, MechanicalParams,
                                      StretchingTorsionalSpringParams)
    from oracle import Oracle
    geo = QuadGeometry(5, 4, spacing=15.0, bond_length=2.25)
    bc, cnvf, bonds, refv = geo.get_parametrization()
    hs, vs = geo.get_design_from_rotated_square(20 * math.pi / 180)
    pairs = np.array([[10, 0], [10, 1], [10, 2], [0, 0], [0, 1], [0, 2], [4, 0], [4, 1], [4, 2]])
    vec = np.zeros(len(pairs))
    vec[0] = 1.0
    solve = setup_dynamic_solver(geo, build_strain_energy(bonds(), stretching_torsional_spring_energy),
                                 constrained_block_DOF_pairs=pairs, constrained_DOFs_fn=pulse_drive(vec),
                                 damped_blocks=np.arange(geo.n_blocks), rtol=1e-8, atol=1e-6)
    ks = torch.tensor(2.5, dtype=torch.float64, device="cuda", requires_grad=True)
    kr = torch.tensor(1.5, dtype=torch.float64, device="cuda", requires_grad=True)

    def params(bond_params):
        return ControlParams(
            geometrical_params=GeometricalParams(block_centroids=bc(hs, vs), centroid_node_vectors=cnvf(hs, vs)),
            mechanical_params=MechanicalParams(bond_params=bond_params, density=6.18e-9, damping=2.0e-5),
            constraint_params=dict(amplitude=3.0, loading_rate=40.0, input_delay=0.002))

    cp = params(StretchingTorsionalSpringParams(k_stretch=ks, k_rot=kr))
    ts = torch.linspace(0, 0.02, 5, dtype=torch.float64)
    fields = solve(torch.zeros(2, geo.n_blocks, 3, dtype=torch.float64), ts, cp)
    (fields[:, 1] ** 2).sum().backward()
    s = solve.solver
    assert s.spec.bond_energy == _abi.DFX_BOND_SPRING
    leaves, pb, dpd, aug = lower_params(s.spec, s.drive, cp, None, "cpu")
    assert aug == 4 * s.spec.n_free + 1 + (2 * geo.n_blocks + 8 * geo.n_blocks) + 2 + 1 + 1 + s.spec.n_free + 3  # no k_shear / reference_vector
    orc = Oracle(s.spec)
    lv = {k: v.detach().cpu().numpy() for k, v in leaves.items()}
    ph = orc.params(1, lv, pb, dpd)
    y0 = np.zeros(2 * s.spec.n_free)
    ys_h, _ = orc.forward(ph, y0, ts.numpy(), 1e-8, 1e-6)
    fields_h = orc.expand_fields(ph, ys_h, ts.numpy())
    assert rel_l2(fields.detach().cpu().numpy(), fields_h[0]) <= 1e-6
    g = np.zeros_like(ys_h)
    g[:, :, s.spec.n_free:] = 2 * ys_h[:, :, s.spec.n_free:]
    _, _, gr_h, _ = orc.adjoint(ph, ys_h, ts.numpy(), g, 1e-8, 1e-6, aug)
    assert abs(ks.grad.item() - gr_h["k_stretch"][0]) <= 1e-5 * abs(gr_h["k_stretch"][0])
    assert abs(kr.grad.item() - gr_h["k_rot"][0]) <= 1e-5 * abs(gr_h["k_rot"][0])
    with pytest.raises(TypeError):  # ligament parameters with the spring energy
        solve(torch.zeros(2, geo.n_blocks, 3, dtype=torch.float64), ts,
              params(LigamentParams(k_stretch=1.0, k_shear=1.0, k_rot=1.0, reference_vector=refv())))


def test_c_abi_error_behaviour_on_the_device():
    """return codes + dfx_last_error instead of crashes or silent fallbacks: undersized workspace, NULL pointers,
    bad sizes, inputs of the wrong length at the Python mirror"""
    import ctypes as C
    from difflexmm_b200 import _abi, _lib
    P = _problem()
    s = P.setup()
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(P.initial_design(), device="cuda")
    ps = _abi.ParamSet(P.spec, 1, {k: v.contiguous() for k, v in leaves.items()}, pb, dpd)
    p = ps.to_struct()
    N, n_t = 2 * P.spec.n_free, ts.shape[0]
    ys = torch.empty((1, n_t, N), dtype=torch.float64, device="cuda")
    stats = torch.zeros((1, _abi.STATS_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    lib, h = _lib.lib, s.handle._h
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    need = lib.dfx_forward_workspace_bytes(h, 1)
    assert need > 8 and lib.dfx_adjoint_workspace_bytes(h, 1) > need
    ws = torch.empty((need,), dtype=torch.uint8, device="cuda")

    def fwd(y0_ptr, n_t_, ws_bytes, batch=1):
        return lib.dfx_forward(h, C.byref(p), batch, C.c_void_p(y0_ptr), C.c_int64(0), C.c_void_p(ts.data_ptr()), C.c_int64(0), n_t_,
                               C.c_double(1e-8), C.c_double(1e-4), None, C.c_void_p(ys.data_ptr()), C.c_void_p(stats.data_ptr()),
                               C.c_void_p(ws.data_ptr()), C.c_size_t(ws_bytes), stream)

    assert fwd(y0.data_ptr(), n_t, need) == 0
    assert fwd(y0.data_ptr(), n_t, 8) != 0 and b"workspace too small" in lib.dfx_last_error()
    assert fwd(None, n_t, need) != 0 and b"NULL" in lib.dfx_last_error()
    assert fwd(y0.data_ptr(), 0, need) != 0
    assert fwd(y0.data_ptr(), n_t, need, batch=0) != 0
    assert lib.dfx_adjoint_objective(h, C.byref(p), 1, C.c_void_p(ys.data_ptr()), C.c_void_p(ts.data_ptr()), C.c_int64(0), n_t, None,
                                   C.c_double(1e-8), C.c_double(1e-4), C.c_int64(0), None, None, None, None, None, None,
                                   C.c_size_t(0), stream) != 0
    torch.cuda.synchronize()
    with pytest.raises(ValueError):  # state0 of the wrong size
        s.lib_forward(ps, y0[:-2], ts)
    with pytest.raises(RuntimeError):  # CUDA-only: there is no CPU path
        from difflexmm_b200.dynamics import DynamicSolver
        DynamicSolver(P.spec, P.drive, device="cpu")


@pytest.mark.parametrize("kernel", ["fast", "generic"])
def test_fused_angular_momentum_objective_equals_the_unfused_path(kernel, monkeypatch):
    """DFX_OBJ_ANGULAR: value, explicit inertia / arm derivatives and the in-kernel cotangent (displacement AND
    velocity parts) against the torch objective on the expanded fields + dfx_adjoint"""
    if kernel == "generic":
        monkeypatch.setenv("DFX_ADJOINT_KERNEL", "generic")
        monkeypatch.setenv("DFX_FORWARD_KERNEL", "generic")
    P = _problem()
    P.setup()
    hs0, vs0 = P.initial_design()
    center = P.geometry.block_centroids(hs0, vs0)[P.target_blocks()].mean(0) + torch.tensor([3.0, -2.0], dtype=torch.float64)
    res = []
    for fused in (False, True):
        hs, vs = hs0.clone().requires_grad_(True), vs0.clone().requires_grad_(True)
        L = P.target_angular_momentum((hs, vs), spin_center=center, fused=fused)
        assert L.dim() == 0
        L.backward()
        res.append((L.item(), hs.grad.clone(), vs.grad.clone()))
    (L0, gh0, gv0), (L1, gh1, gv1) = res
    assert abs(L0 - L1) <= 1e-11 * abs(L0)
    assert rel_l2(gh1.numpy(), gh0.numpy()) <= 1e-9 and rel_l2(gv1.numpy(), gv0.numpy()) <= 1e-9
    # a batch with per-design weights
    B = 2
    hsb, vsb = P.random_ensemble(B, noise=0.03)
    hsb, vsb = hsb.cuda().requires_grad_(True), vsb.cuda().requires_grad_(True)
    Lb = P.target_angular_momentum((hsb, vsb), spin_center=center, batch=B)
    (Lb * torch.tensor([1.0, -2.0], dtype=torch.float64, device="cuda")).sum().backward()
    hs1, vs1 = hsb[1].detach().cpu().requires_grad_(True), vsb[1].detach().cpu().requires_grad_(True)
    L1s = P.target_angular_momentum((hs1, vs1), spin_center=center)
    (-2.0 * L1s).backward()
    assert abs(Lb[1].item() - L1s.item()) <= 1e-10 * abs(L1s.item())
    assert rel_l2(hsb.grad[1].cpu().numpy(), hs1.grad.numpy()) <= 1e-8


def test_fused_objective_on_the_kagome_lattice():
    """the fused design-to-gradient path (device geometry with three shift arrays, triangles) on a kagome lattice
    against the unfused torch path"""
    from difflexmm_b200.problems import KagomeFocusing
    P = KagomeFocusing(n1_cells=8, n2_cells=6, simulation_time=0.01, n_timepoints=6, target_shift=(1, 1))
    P.setup()
    rng = np.random.default_rng(2)
    base = [d + 0.2 * torch.from_numpy(rng.standard_normal(d.shape)) for d in P.initial_design()]
    res = []
    for fused in (False, True):
        d = [x.clone().requires_grad_(True) for x in base]
        J = P.target_kinetic_energy(d, fused=fused)
        J.backward()
        res.append((J.item(), [x.grad.clone() for x in d]))
    (J0, g0), (J1, g1) = res
    assert abs(J0 - J1) <= 1e-11 * abs(J0)
    for a, b in zip(g1, g0):
        assert rel_l2(a.numpy(), b.numpy()) <= 1e-8


def test_bench_prints_the_contract_line():
    """bench.py end to end on a tiny ensemble (shortened horizon: not a measurement): every key the driver reads"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--designs", "8", "--steps", "2", "--warmup", "1",
                          "--cpu-designs", "2", "--horizon-scale", "0.05"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["gpu_launches"] == 6 and d["failed_designs"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["achieved"] > 0 and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "TFLOP/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert "PROFILING_ONLY_horizon_scale" in d["config"] and "workload" in d["config"]


def test_cfg4_lattice_size_uses_the_512_thread_kernels():
    """quads 24 x 18 (432 units, 822 bonds: the 512-thread instances of both fast kernels, static ramp + delayed pulse
    drive with 150 constrained DOFs) against the C++ oracle, short dynamic window"""
    from difflexmm_b200 import _abi
    from difflexmm_b200.problems import QuadsStaticTuning
    from oracle import Oracle
    P = QuadsStaticTuning(simulation_time_dynamic=0.004, n_timepoints=4, compressive_strain=0.02, compressive_strain_rate=25.0)
    s = P.setup()
    assert P.spec.n_blocks == 432 and P.spec.n_bonds == 822
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(P.initial_design(), device="cuda")
    ps = _abi.ParamSet(P.spec, 1, {k: v.contiguous() for k, v in leaves.items()}, pb, dpd)
    ys, st = s.lib_forward(ps, y0, ts)
    assert st.numpy()["status"][0] == 0
    orc = Oracle(P.spec)
    ph = orc.params(1, {k: v.cpu().numpy() for k, v in leaves.items()}, pb, dpd)
    ys_h, st_h = orc.forward(ph, y0.cpu().numpy(), ts.cpu().numpy(), P.rtol, P.atol)
    assert abs(int(st.numpy()["steps"][0]) - int(st_h["steps"][0])) <= 2
    assert rel_l2(ys[0].cpu().numpy(), ys_h[0]) <= 1e-6
    nf = P.spec.n_free
    g = np.zeros_like(ys_h)
    g[:, :, nf:] = ys_h[:, :, nf:] * leaves["inertia"].cpu().numpy()
    y0b_h, tsb_h, gr_h, sb_h = orc.adjoint(ph, ys_h, ts.cpu().numpy(), g, P.rtol, P.atol, aug)
    y0b, tsb, gr, sb = s.lib_adjoint(ps, torch.as_tensor(ys_h, device="cuda"), ts, torch.as_tensor(g, device="cuda"), aug)
    assert sb.numpy()["status"][0] == 0
    for k in gr_h:
        if np.abs(gr_h[k]).max() > 1e-9:
            assert rel_l2(gr[k][0].cpu().numpy(), gr_h[k][0]) <= 1e-5, k
