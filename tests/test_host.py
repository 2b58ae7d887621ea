"""Host-side mirror of the reference interface: geometry, DOF bookkeeping, lowering of
setup_dynamic_solver's arguments and of ControlParams (no GPU needed)."""

import math
import os

import numpy as np
import pytest
import torch

from difflexmm_b200 import _abi
from difflexmm_b200.dynamics import lower_params, lower_topology
from difflexmm_b200.energy import (build_contact_energy, build_strain_energy, combine_block_energies, ligament_energy,
                                   ligament_energy_linearized)
from difflexmm_b200.geometry import (DOFsInfo, KagomeGeometry, QuadGeometry, RotatedSquareGeometry, compute_inertia,
                                     polygon_area, polygon_centroid, polygon_polar_moment)
from difflexmm_b200.loading import pulse_drive, ramp_load
from difflexmm_b200.problems import KagomeFocusing, QuadsFocusing, QuadsStaticTuning


def test_polygon_properties_of_a_rectangle():
    v = torch.tensor([[[0., 0.], [2., 0.], [2., 1.], [0., 1.]]], dtype=torch.float64)
    assert torch.allclose(polygon_area(v), torch.tensor([2.], dtype=torch.float64))
    assert torch.allclose(polygon_centroid(v), torch.tensor([[1., 0.5]], dtype=torch.float64))
    # polar moment of a b x h rectangle about its centroid: b h (b^2 + h^2) / 12
    assert torch.allclose(polygon_polar_moment(v), torch.tensor([2 * (4 + 1) / 12.], dtype=torch.float64))
    I = compute_inertia(v, 3.0)
    assert torch.allclose(I, torch.tensor([[6., 6., 2.5]], dtype=torch.float64))


def test_dofs_info_order():
    free, cons, all_ids = DOFsInfo(3, np.array([[2, 1], [0, 0]]))
    assert cons.tolist() == [7, 0]                       # pair order (reference geometry.py:174)
    assert free.tolist() == [1, 2, 3, 4, 5, 6, 8]        # ascending
    assert all_ids.tolist() == list(range(9))
    free, cons, _ = DOFsInfo(2, np.array([]))
    assert len(cons) == 0 and free.tolist() == list(range(6))


def test_quad_geometry_counts_and_rotated_square_design():
    g = QuadGeometry(24, 16, spacing=15., bond_length=2.25)
    bc, cnv, bonds, ref = g.get_parametrization()
    assert g.n_blocks == 384 and bonds().shape == (728, 2) and ref().shape == (728, 2)
    hs, vs = g.get_design_from_rotated_square(25 * math.pi / 180)
    assert hs.shape == (25, 16, 2) and vs.shape == (24, 17, 2)
    c = cnv(hs, vs)
    assert c.shape == (384, 4, 2)
    # centroid_node_vectors are measured from the centroid
    assert torch.allclose(polygon_centroid(c), torch.zeros(384, 2, dtype=torch.float64), atol=1e-12)
    # every node belongs to at most one bond
    assert len(np.unique(bonds())) == bonds().size
    # the quad design "from rotated square(angle)" is the rotated-square lattice of the opposite angle: node 0 of
    # block (n1, n2) takes horizontal_shift[n1 + 1, n2], whose parity is flipped (reference geometry.py:866-873, 936-941)
    rs = RotatedSquareGeometry(12, 8, spacing=15., bond_length=2.25)
    rs.compute_geometry()
    c_rs = rs.centroid_node_vectors(-25 * math.pi / 180)
    assert torch.allclose(c_rs, c, atol=1e-12)


def test_kagome_geometry_counts():
    g = KagomeGeometry(20, 12, 20. * np.array([[1., 0.], [math.cos(math.pi / 3), math.sin(math.pi / 3)]]), 2.25)
    bc, cnv, bonds, ref = g.get_parametrization()
    assert g.n_blocks == 480 and g.n_npb == 3
    assert bonds().shape == (240 + 220 + 228, 2)
    assert len(np.unique(bonds())) == bonds().size
    c = cnv()
    assert c.shape == (480, 3, 2)
    assert torch.allclose(polygon_centroid(c), torch.zeros(480, 2, dtype=torch.float64), atol=1e-12)
    # bond reference vectors close the gap between the two bonded nodes of the regular lattice
    nodes = (c + bc()[:, None, :]).reshape(-1, 2)
    b = bonds()
    assert torch.allclose(nodes[b[:, 1]] - nodes[b[:, 0]], ref(), atol=1e-12)


@pytest.mark.parametrize("P,sizes", [
    (QuadsFocusing, dict(n_blocks=384, n_bonds=728, n_cons=42, n_free=1110, aug=12009)),
    (KagomeFocusing, dict(n_blocks=480, n_bonds=688, n_cons=48, n_free=1392, aug=None)),
    (QuadsStaticTuning, dict(n_blocks=432, n_bonds=822, n_cons=150, n_free=1146, aug=None)),
])
def test_reference_configurations_lower_to_the_survey_sizes(P, sizes):
    p = P()
    spec, drive = p.lower()
    assert spec.n_blocks == sizes["n_blocks"] and spec.n_bonds == sizes["n_bonds"]
    assert len(spec.constrained_dofs) == sizes["n_cons"] and spec.n_free == sizes["n_free"]
    leaves, pb, dpd, aug, y0, ts = p.boundary_inputs(p.initial_design())
    assert leaves["inertia"].shape == (spec.n_free,) and dpd and pb == ()
    assert leaves["drive"].shape == (spec.n_drive_params,)
    if sizes["aug"]:
        assert aug == sizes["aug"]
    # what libdfx counts by itself when aug_size = 0 lacks block_centroids and density
    ps = _abi.ParamSet(spec, 1, {k: v.numpy() for k, v in leaves.items()}, pb, dpd)
    assert aug == ps.aug_size_of_listed_leaves() + 2 * spec.n_blocks + 1


def test_batched_lowering_keeps_per_design_counts():
    p = QuadsFocusing(n1_blocks=8, n2_blocks=7)
    spec, drive = p.lower()
    hs, vs = p.random_ensemble(3, noise=0.05)
    l1, _, _, aug1, _, _ = p.boundary_inputs((hs[0], vs[0]))
    lb, pb, dpd, augb, _, _ = p.boundary_inputs((hs, vs), batch=3)
    assert augb == aug1
    assert lb["centroid_node_vectors"].shape == (3,) + l1["centroid_node_vectors"].shape
    assert lb["inertia"].shape == (3, spec.n_free)
    assert torch.allclose(lb["inertia"][0], l1["inertia"])
    assert lb["damping"].shape == l1["damping"].shape  # shared leaf stays shared
    ps = _abi.ParamSet(spec, 3, {k: v.numpy() for k, v in lb.items()}, pb, dpd)
    assert ps.batched["centroid_node_vectors"] and not ps.batched["reference_vector"]


def test_gradients_flow_from_the_design_to_the_leaves():
    p = QuadsFocusing(n1_blocks=8, n2_blocks=7)
    p.lower()
    hs, vs = p.initial_design()
    hs = hs.clone().requires_grad_(True)
    leaves, *_ = p.boundary_inputs((hs, vs))
    (leaves["inertia"].sum() + leaves["centroid_node_vectors"].pow(2).sum()).backward()
    assert hs.grad is not None and torch.isfinite(hs.grad).all() and hs.grad.abs().sum() > 0


def test_arbitrary_scalar_drive_can_be_tabulated():
    """a Python closure as the driving signal (what the reference accepts) enters through loading.tabulate_drive"""
    import math
    import torch
    from difflexmm_b200.loading import tabulate_drive, tabulated_drive
    f = lambda t: 2.0 * math.sin(40.0 * t) ** 2 if t > 0.001 else 0.0  # noqa: E731
    times = np.linspace(0.0, 0.05, 2001)
    d = tabulate_drive(f, times, [1.0, 0.0])
    assert isinstance(d, tabulated_drive) and d.kind == _abi.DFX_DRIVE_TABLE and len(d.values) == 2001
    t = torch.tensor([0.0, 0.0123, 0.03, 0.07], dtype=torch.float64)
    s0, s1 = d.channels(t)
    ref = torch.tensor([f(0.0), f(0.0123), f(0.03), f(0.05)], dtype=torch.float64)  # constant beyond the last sample
    assert torch.allclose(s0, ref, atol=2e-6) and float(s1.abs().max()) == 0.0


def test_vocabulary_is_closed_and_loud():
    g = QuadGeometry(4, 3, spacing=15., bond_length=2.25)
    g.compute_geometry()
    bonds = g.bond_connectivity()
    energy = combine_block_energies(build_strain_energy(bonds, ligament_energy), build_contact_energy(bonds))
    with pytest.raises(TypeError):
        lower_topology(g, lambda u, cp: 0.0)                                  # arbitrary Python energy
    with pytest.raises(TypeError):
        lower_topology(g, energy, constrained_block_DOF_pairs=[[0, 0]], constrained_DOFs_fn=lambda t: 0.0)
    with pytest.raises(TypeError):
        build_strain_energy(bonds, lambda *a, **k: 0.0)
    dist = combine_block_energies(build_strain_energy(bonds, ligament_energy), build_contact_energy(bonds, angle_based=False))
    assert dist.contact == _abi.DFX_CONTACT_DISTANCE and energy.contact == _abi.DFX_CONTACT_ANGLE  # both contact models lower
    assert lower_topology(g, dist)[0].contact == _abi.DFX_CONTACT_DISTANCE
    with pytest.raises(ValueError):
        lower_topology(g, energy, constrained_block_DOF_pairs=[[0, 0], [0, 0]], constrained_DOFs_fn=pulse_drive([1, 0]))
    spec, drive = lower_topology(g, build_strain_energy(bonds, ligament_energy_linearized), [[11, 0]], ramp_load(0.2, 1e-3),
                                 [[0, 0], [0, 1]], pulse_drive([1., 0.]), np.arange(g.n_blocks))
    assert spec.bond_energy == _abi.DFX_BOND_LINEARIZED and not spec.contact
    assert spec.load_kind == _abi.DFX_LOAD_RAMP and spec.loaded_dofs.tolist() == [33]
    assert spec.drive_kind == _abi.DFX_DRIVE_PULSE and spec.n_drive_params == 3


def test_solver_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from difflexmm_b200.dynamics import setup_dynamic_solver
    g = QuadGeometry(4, 3, spacing=15., bond_length=2.25)
    g.compute_geometry()
    with pytest.raises(RuntimeError):
        setup_dynamic_solver(g, build_strain_energy(g.bond_connectivity(), ligament_energy), device="cpu")


def test_tabulated_drive_matches_numpy_interp():
    """`excited_blocks_fn = jnp.interp(t, times, values)` (reference problems/quads_focusing.py:223-227): oracle,
    torch post-processing and numpy agree, including the clamped ends and the slope used as velocity"""
    from difflexmm_b200.dynamics import drive_values
    from difflexmm_b200.loading import tabulated_drive
    from oracle import Oracle
    times = np.linspace(0.002, 0.03, 15)
    values = 3.0 * np.sin(40 * times) ** 2
    g = QuadGeometry(4, 3, spacing=15., bond_length=2.25)
    g.compute_geometry()
    drive = tabulated_drive(times, values, [1., 0., -0.5])
    spec, drive = lower_topology(g, build_strain_energy(g.bond_connectivity(), ligament_energy), None, None,
                                 [[0, 0], [0, 1], [5, 2]], drive, np.arange(g.n_blocks))
    assert spec.drive_kind == _abi.DFX_DRIVE_TABLE and spec.n_drive_params == 0
    ts = np.linspace(0.0, 0.035, 23)
    u, v = drive_values(drive, torch.as_tensor(ts)[None], {}, "cpu")
    ref = np.interp(ts, times, values)
    assert np.allclose(u[0, :, 0].numpy(), ref, rtol=1e-14, atol=1e-15)
    assert np.allclose(u[0, :, 2].numpy(), -0.5 * ref, rtol=1e-14, atol=1e-15)
    orc = Oracle(spec)
    cnv = g.centroid_node_vectors(*g.get_design_from_rotated_square(0.3)).numpy()
    leaves = dict(centroid_node_vectors=cnv, reference_vector=g.reference_bond_vectors().numpy(), k_stretch=1., k_shear=1.,
                  k_rot=1., damping=0.1, inertia=np.ones(spec.n_free))
    ps = orc.params(1, leaves)
    f = orc.expand_fields(ps, np.zeros((1, len(ts), 2 * spec.n_free)), ts).reshape(1, len(ts), 2, -1)
    assert np.allclose(f[0, :, 0, 0], ref, rtol=1e-14, atol=1e-15)
    assert np.allclose(f[0, :, 1, 0], v[0, :, 0].numpy(), rtol=1e-12, atol=1e-12)
    inside = (ts > times[0]) & (ts < times[-1])
    assert np.all(f[0, ~inside, 1, 0] == 0.0) and np.any(f[0, inside, 1, 0] != 0.0)


# ---- optimisation layer -------------------------------------------------------------------------
def test_batched_mma_converges_and_instances_are_independent():
    import torch
    from difflexmm_b200.optimization import BatchedMMA
    torch.manual_seed(0)
    B, n = 3, 6
    A = torch.rand(B, n, dtype=torch.float64) + 0.5
    c = torch.randn(B, n, dtype=torch.float64) * 2

    def make(rows):
        return lambda x: ((0.5 * A[rows] * (x - c[rows]) ** 2).sum(1), A[rows] * (x - c[rows]))

    opt = BatchedMMA(make(slice(None)), torch.zeros(B, n, dtype=torch.float64), -1.0, 1.0, maximize=False)
    bx, bf = opt.run(40)
    assert opt.n_evals == 40 and len(opt.history) == 40
    assert (bx - c.clamp(-1, 1)).abs().max() < 1e-5          # box-constrained minimiser
    for i in range(B):                                        # lock-step batching does not couple the instances
        o1 = BatchedMMA(make(slice(i, i + 1)), torch.zeros(1, n, dtype=torch.float64), -1.0, 1.0, maximize=False)
        x1, f1 = o1.run(40)
        assert torch.equal(x1[0], bx[i]) and torch.equal(f1[0], bf[i])
    # maximisation, no bounds
    opt = BatchedMMA(lambda x: (-(x ** 2).sum(1), -2 * x), torch.full((2, 3), 3.0, dtype=torch.float64), maximize=True)
    bx, bf = opt.run(60)
    assert bx.abs().max() < 1e-4 and (bf <= 0).all()


def test_batched_mma_with_inequality_constraints_reaches_the_constrained_minimum():
    """inequality constraints inside the batched MMA (reference: nlopt LD_MMA + add_inequality_mconstraint,
    problems/quads_focusing.py:578-627): linear pair constraints and one dense quadratic row through the fixed-width sparse
    Jacobian interface, against scipy SLSQP on every instance; every recorded best design is feasible to the tolerance"""
    import numpy as np
    import torch
    from scipy.optimize import minimize
    from difflexmm_b200.optimization import BatchedMMA
    torch.manual_seed(0)
    B, n = 4, 6
    t = torch.randn(B, n, dtype=torch.float64) * 2
    cols = torch.zeros(n, n, dtype=torch.int64)
    for j in range(n - 1):
        cols[j, 0], cols[j, 1] = j, j + 1
    cols[n - 1] = torch.arange(n)

    def con(x):
        c = torch.zeros(x.shape[0], n, dtype=torch.float64)
        J = torch.zeros(x.shape[0], n, n, dtype=torch.float64)
        c[:, :n - 1] = x[:, :-1] + x[:, 1:] - 0.5
        J[:, :n - 1, 0] = 1.0
        J[:, :n - 1, 1] = 1.0
        c[:, n - 1] = (x ** 2).sum(1) - 4.0
        J[:, n - 1, :] = 2 * x
        return c, J, cols

    opt = BatchedMMA(lambda x: (((x - t) ** 2).sum(1), 2 * (x - t)), torch.zeros(B, n, dtype=torch.float64), -3.0, 3.0,
                     maximize=False, constraints=con)
    bx, bf = opt.run(60)
    assert opt.n_evals == 60 and len(opt.history) == 60 and len(opt.violation_history) == 60
    assert (opt.best_violation <= 1e-8).all() and (con(bx)[0] <= 1e-8).all()
    for b in range(B):
        tb = t[b].numpy()
        ref = minimize(lambda x: ((x - tb) ** 2).sum(), np.zeros(n), jac=lambda x: 2 * (x - tb), method="SLSQP", bounds=[(-3, 3)] * n,
                       constraints=[{"type": "ineq", "fun": lambda x, j=j: 0.5 - x[j] - x[j + 1]} for j in range(n - 1)]
                       + [{"type": "ineq", "fun": lambda x: 4 - (x ** 2).sum()}], options={"ftol": 1e-14, "maxiter": 500})
        assert abs(float(bf[b]) - ref.fun) <= 1e-5 * abs(ref.fun) and np.abs(bx[b].numpy() - ref.x).max() <= 1e-3
    # an infeasible start is driven into the feasible set
    opt = BatchedMMA(lambda x: (((x - t) ** 2).sum(1), 2 * (x - t)), torch.full((B, n), 2.5, dtype=torch.float64), -3.0, 3.0,
                     maximize=False, constraints=con)
    assert (con(opt.x)[0] > 0).any()
    opt.run(40)
    assert (opt.best_violation <= 1e-8).all()


def test_constraints_match_a_per_bond_restatement():
    """angle / edge-length constraints (reference problems/quads_focusing.py:473-544) against a literal per-bond loop
    over the reference's compute_edge_unit_vectors / angle_between_unit_vectors (geometry.py:181-253)"""
    import math
    import torch
    from difflexmm_b200.geometry import QuadGeometry
    from difflexmm_b200.optimization import angle_constraints, edge_length_constraints, quad_boundary_node_ids
    geo = QuadGeometry(5, 4, spacing=15.0, bond_length=2.25)
    geo.compute_geometry()
    rng = np.random.default_rng(1)
    design = [d + torch.from_numpy(rng.uniform(-1, 1, d.shape)) for d in geo.get_design_from_rotated_square(25 * math.pi / 180)]
    nodes = geo.centroid_node_vectors(*design).numpy()

    def unit(v):
        return v / np.linalg.norm(v)

    def edges(n):
        blk, l = n // 4, n % 4
        return unit(nodes[blk, (l + 1) % 4] - nodes[blk, l]), unit(nodes[blk, (l - 1) % 4] - nodes[blk, l])

    def ang(u1, u2):
        return math.atan2(u1[0] * u2[1] - u1[1] * u2[0], u1[0] * u2[0] + u1[1] * u2[1])

    rows = []
    for n1, n2 in geo.bond_connectivity():
        a1, a2 = edges(n1)
        b1, b2 = edges(n2)
        rows.append([ang(b2, a1) % (2 * math.pi), ang(a2, b1) % (2 * math.pi), ang(a1, a2) % (2 * math.pi), ang(b1, b2) % (2 * math.pi)])
    rows = np.array(rows)
    bnd = np.array([ang(*edges(n)) % (2 * math.pi) for n in quad_boundary_node_ids(5, 4)])
    mv, mb = 0.05, 0.3
    want = np.concatenate([-(rows[:, 0] - mv), -(rows[:, 1] - mv), -(rows[:, 2] - mb), -(rows[:, 3] - mb), -(bnd - mb)])
    got = angle_constraints(geo, mv, mb, boundary_angle_constraint=True)(design).numpy()
    assert got.shape == (4 * len(rows) + 2 * (5 + 4),) and np.abs(got - want).max() < 1e-13
    assert angle_constraints(geo, mv, mb)(design).shape == (4 * len(rows),)
    el = edge_length_constraints(geo, 1.0)(design).numpy()
    want_el = -(np.linalg.norm(np.roll(nodes, 1, axis=1) - nodes, axis=2).reshape(-1) - 1.0)
    assert np.abs(el - want_el).max() < 1e-13
    # Jacobian by autograd (the reference takes jax.jacobian)
    d0 = [d.clone().requires_grad_(True) for d in design]
    J = torch.autograd.functional.jacobian(lambda a, b: edge_length_constraints(geo, 1.0)([a, b]), tuple(d0))
    assert J[0].shape == (20 * 4,) + tuple(design[0].shape) and torch.isfinite(J[0]).all()


def test_save_load_data_is_pickle_compatible_with_the_reference(tmp_path):
    """files written here load under the reference's class path `difflexmm.utils.SolutionData` (utils.py:9-25,166-201);
    files from the reference side load here"""
    import pickle
    import sys
    import types
    from typing import Any, NamedTuple
    import torch
    from difflexmm_b200.utils import SolutionData, load_data, save_data
    sd = SolutionData(torch.zeros(3, 2), torch.ones(3, 4, 2), np.array([[0, 5]]), torch.linspace(0, 1, 4), torch.zeros(4, 2, 3, 3))
    path = save_data(tmp_path / "out" / "solution.pkl", {"solution": sd, "objective_values": [1.0, 2.0]})
    assert "difflexmm" not in sys.modules
    back = load_data(path)
    assert isinstance(back["solution"], SolutionData) and isinstance(back["solution"].fields, np.ndarray)
    assert back["solution"].fields.shape == (4, 2, 3, 3) and back["objective_values"] == [1.0, 2.0]

    class RefSolutionData(NamedTuple):  # what the reference's unpickler resolves the global to
        block_centroids: Any
        centroid_node_vectors: Any
        bond_connectivity: Any
        timepoints: Any
        fields: Any

    mod = types.ModuleType("difflexmm.utils")
    mod.SolutionData = RefSolutionData
    RefSolutionData.__module__, RefSolutionData.__qualname__ = "difflexmm.utils", "SolutionData"
    sys.modules["difflexmm"], sys.modules["difflexmm.utils"] = types.ModuleType("difflexmm"), mod
    try:
        with open(path, "rb") as f:
            ref_side = pickle.load(f)
        assert isinstance(ref_side["solution"], RefSolutionData)
        ref_file = tmp_path / "ref.pkl"
        with open(ref_file, "wb") as f:
            pickle.dump(RefSolutionData(*[np.asarray(v) for v in back["solution"]]), f)
    finally:
        del sys.modules["difflexmm"], sys.modules["difflexmm.utils"]
    assert isinstance(load_data(ref_file), SolutionData)


def test_design_vertex_table_of_the_device_geometry():
    """the vertex -> design-variable table handed to dfx_geometry_create, against the reference's indexing
    (geometry.py:866-883 quads: vertex 0..3 <- hs[a+1,b], vs[a,b+1], hs[a,b], vs[a,b]; kagome :690-720)"""
    import torch
    from difflexmm_b200.geometry import KagomeGeometry, QuadGeometry, RotatedSquareGeometry
    from difflexmm_b200.geometry_device import design_vertex_table
    n1, n2 = 5, 4
    geo = QuadGeometry(n1, n2, spacing=15.0, bond_length=2.25)
    geo.compute_geometry()
    base, nd, shapes, sizes = design_vertex_table(geo)
    assert shapes == [(n1 + 1, n2, 2), (n1, n2 + 1, 2)] and sizes == [(n1 + 1) * n2, n1 * (n2 + 1)]
    n_hs = sizes[0]
    for blk in range(n1 * n2):
        a, b = blk % n1, blk // n1  # row-major over n2 then n1 (geometry._grid_row_major)
        want = [(a + 1) * n2 + b, n_hs + a * (n2 + 1) + b + 1, a * n2 + b, n_hs + a * (n2 + 1) + b]
        assert nd[blk].tolist() == want
    assert torch.allclose(base, geo.reference_node_vectors(*[torch.zeros(s, dtype=torch.float64) for s in shapes]))
    kag = KagomeGeometry(4, 3)
    kag.compute_geometry()
    base, nd, shapes, sizes = design_vertex_table(kag)
    assert nd.shape == (24, 3) and nd.min() >= 0 and nd.max() == sum(sizes) - 1
    counts = np.bincount(nd.numpy().ravel(), minlength=sum(sizes))
    assert counts.min() == 1 and counts.max() == 2  # boundary shifts feed one vertex, interior ones two
    with pytest.raises(TypeError):
        design_vertex_table(RotatedSquareGeometry(2, 2))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port timed through the bench harness) prints one JSON line with the keys the
    driver reads; shortened horizon and two designs so that it takes a second"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-designs", "2", "--horizon-scale", "0.05"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "designs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["dtype"] == "f64" and line["scaling"] == "strong"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "designs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_every_python_file_compiles():
    """bench.py, __graft_entry__.py, tools/ and the package byte-compile (no GPU needed to catch a syntax error)"""
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, f) for f in ("bench.py", "__graft_entry__.py")]
    for sub in ("tools", "difflexmm_b200", "oracle", "tests"):
        for dp, _, fs in os.walk(os.path.join(root, sub)):
            files += [os.path.join(dp, f) for f in fs if f.endswith(".py")]
    assert len(files) > 20
    for f in files:
        py_compile.compile(f, doraise=True)


def test_optimization_problem_flatten_unflatten_roundtrip():
    """design tuple <-> flat vector (the reference uses jax.flatten_util.ravel_pytree, quads_focusing.py:562-563),
    single designs and batches"""
    import torch
    from difflexmm_b200.optimization import OptimizationProblem
    from difflexmm_b200.problems import KagomeFocusing, QuadsFocusing
    for P in (QuadsFocusing(n1_blocks=8, n2_blocks=7), KagomeFocusing(n1_cells=8, n2_cells=6)):
        P.lower()
        opt = OptimizationProblem(P)
        rng = np.random.default_rng(0)
        shapes = [tuple(s) for s in P.geometry.design_shapes]
        single = [torch.from_numpy(rng.standard_normal(s)) for s in shapes]
        x = opt.flatten(single)
        assert x.shape == (1, sum(int(np.prod(s)) for s in shapes))
        back = opt.unflatten(x)
        assert all(torch.equal(b[0], a) for a, b in zip(single, back))
        batch = [torch.from_numpy(rng.standard_normal((3,) + s)) for s in shapes]
        xb = opt.flatten(batch)
        assert xb.shape[0] == 3 and all(torch.equal(b, a) for a, b in zip(batch, opt.unflatten(xb)))
        # ravel order: first design array first, C order (as ravel_pytree)
        assert torch.equal(xb[1, :int(np.prod(shapes[0]))], batch[0][1].reshape(-1))
    with pytest.raises(ImportError):
        OptimizationProblem(P).run_optimization_nlopt(None, 1)


def test_damping_leaf_broadcasts_like_the_reference():
    """the reference multiplies the damping leaf by ones((n_damped, 3)) (loading.py:93-101), so (3,), (1, 3) and
    (n_damped, 1) leaves are valid; the lowering hands the kernel the broadcast array, keeps the leaf's own element count
    in the augmented-state size, and autograd sums the cotangent back to the leaf's shape"""
    from difflexmm_b200.utils import ContactParams, ControlParams, GeometricalParams, LigamentParams, MechanicalParams
    P = QuadsFocusing(n1_blocks=8, n2_blocks=7)
    spec, drive = P.lower()
    nb = spec.n_blocks
    bc, cnvf, bonds, refv = P.geometry.get_parametrization()
    hs, vs = P.initial_design()
    cnv = cnvf(hs, vs)

    def lower(damping):
        cp = ControlParams(
            geometrical_params=GeometricalParams(block_centroids=bc(hs, vs), centroid_node_vectors=cnv),
            mechanical_params=MechanicalParams(bond_params=LigamentParams(k_stretch=120., k_shear=1.19, k_rot=1.5, reference_vector=refv()),
                                               density=6.18e-9, damping=damping,
                                               contact_params=ContactParams(min_angle=-0.26, cutoff_angle=-0.17, k_contact=1.5)),
            constraint_params=dict(amplitude=7.5, loading_rate=30., input_delay=0.1 / 30))
        return lower_params(spec, drive, cp, None, "cpu")

    full = torch.rand(nb, 3, dtype=torch.float64)
    leaves_full, _, dpd_full, aug_full = lower(full)
    row = torch.tensor([1e-5, 2e-5, 3e-5], dtype=torch.float64, requires_grad=True)
    leaves_row, _, dpd_row, aug_row = lower(row)
    assert dpd_full and dpd_row and leaves_row["damping"].shape == (nb, 3)
    assert torch.equal(leaves_row["damping"], row.detach().expand(nb, 3))
    assert aug_full - aug_row == 3 * nb - 3  # the leaf counts with its own size
    leaves_row["damping"].sum().backward()
    assert torch.allclose(row.grad, torch.full((3,), float(nb), dtype=torch.float64))
    col = torch.rand(nb, 1, dtype=torch.float64)
    assert torch.equal(lower(col)[0]["damping"], col.expand(nb, 3))
    with pytest.raises(ValueError):
        lower(torch.rand(nb, 2, dtype=torch.float64))


def test_run_optimization_mma_lowers_the_problem_first():
    """a fresh OptimizationProblem (geometry not lowered yet) must reach the point where it needs the CUDA library, not
    fail on `geometry is None` while flattening the initial guess"""
    from difflexmm_b200.optimization import OptimizationProblem
    P = QuadsFocusing(n1_blocks=8, n2_blocks=7)
    assert P.geometry is None
    opt = OptimizationProblem(P)
    x = opt.flatten if False else None  # noqa: F841
    try:
        opt.run_optimization_mma(None, 1)
    except AttributeError as e:  # 'NoneType' object has no attribute 'design_shapes' was the bug
        assert "design_shapes" not in str(e), e
    except Exception:
        pass
    assert P.geometry is not None
