"""TEST INFRASTRUCTURE ONLY -- literal CPU restatement of the reference hot path.

This file restates, formula by formula, what the reference computes on the path
`setup_dynamic_solver(...) -> solve_dynamics` and what `jax.grad` derives for it, using
torch float64 + torch.autograd in place of jax.numpy + jax autodiff.  It is slow (one
autograd graph per RHS call) and exists to pin the closed forms used by the fast C
oracle (`oracle/dfx_oracle.c`) and by the CUDA kernels.  Nothing in the product package
imports it.

PARITY STATUS: "parity unpinned" against real JAX -- jax / jax-md are not installable in
this image (no wheel, no network), so the integrator below follows the published
algorithm of `jax.experimental.ode` (jax 0.4.8, pinned at `/root/reference/poetry.lock:614`)
from knowledge of that source; the physics follows the in-tree reference files cited per
function.  What *is* pinned: the reference's own two tests for this path
(`tests/test_difflexmm.py:35-146` tensile known answer, `:149-176` frame invariance) are
re-run against this file in `tests/test_oracle.py`.

Reference files followed:
  kinematics.py:13-37,40-81      energy.py:120-176,99-117,204-219,333-361,364-407,410-491
  geometry.py:181-253            loading.py:12-47,71-106       dynamics.py:20-57,138-184
Third-party (absent): jax.experimental.ode (odeint, _odeint_rev), jax_md.quantity.force,
jax_md.smap.bond.
"""

import math

import torch

F64 = torch.float64


# ------------------------------------------------------------------ kinematics.py:13-37
def block_to_node_kinematics(block_displacement, centroid_node_vectors):
    """u_node = u_c + (R(theta) - I) r, third component = theta. (n_blocks, npb, 3)"""
    theta = block_displacement[:, 2]
    c, s = torch.cos(theta), torch.sin(theta)
    rx, ry = centroid_node_vectors[..., 0], centroid_node_vectors[..., 1]
    ux = block_displacement[:, None, 0] + (c[:, None] - 1) * rx - s[:, None] * ry
    uy = block_displacement[:, None, 1] + s[:, None] * rx + (c[:, None] - 1) * ry
    return torch.stack([ux, uy, theta[:, None].expand_as(ux)], -1)


# ------------------------------------------------------------------ energy.py:120-176
def ligament_strains(DOFs1, DOFs2, reference_vector):
    dU = DOFs2[:, :2] - DOFs1[:, :2]
    dRot = DOFs2[:, 2] - DOFs1[:, 2]
    mean_rot = (DOFs2[:, 2] + DOFs1[:, 2]) / 2
    current = dU + reference_vector
    current_angle = torch.atan2(current[:, 1], current[:, 0])
    c, s = torch.cos(mean_rot), torch.sin(mean_rot)
    ref = torch.ones((len(DOFs1), 2), dtype=F64) * reference_vector
    pushed_x = c * ref[:, 0] - s * ref[:, 1]
    pushed_y = s * ref[:, 0] + c * ref[:, 1]
    pushed_angle = torch.atan2(pushed_y, pushed_x)
    axial = ((current * current).sum(-1) / (reference_vector * reference_vector).sum(-1)) ** 0.5 - 1
    shear = torch.remainder(current_angle - pushed_angle + math.pi, 2 * math.pi) - math.pi
    return axial, shear, dRot


def ligament_energy(nodal_DOFs, reference_vector, k_stretch=1., k_shear=1., k_rot=1.):
    axial, shear, dRot = ligament_strains(*nodal_DOFs, reference_vector=reference_vector)
    l0 = torch.linalg.norm(reference_vector, dim=-1)
    return k_stretch * (axial * l0) ** 2 / 2 + k_shear * (shear * l0) ** 2 / 2 + k_rot * dRot ** 2 / 2


# ------------------------------------------------------------------ energy.py:70-117
def ligament_strains_linearized(DOFs1, DOFs2, reference_vector):
    dU = DOFs2[:, :2] - DOFs1[:, :2]
    dRot = DOFs2[:, 2] - DOFs1[:, 2]
    n2 = torch.linalg.norm(reference_vector, dim=-1) ** 2
    axial = (dU * reference_vector).sum(-1) / n2
    cross = reference_vector[..., 0] * dU[:, 1] - reference_vector[..., 1] * dU[:, 0]
    shear = cross / n2 - (DOFs2[:, 2] + DOFs1[:, 2]) / 2
    return axial, shear, dRot


def ligament_energy_linearized(nodal_DOFs, reference_vector, k_stretch=1., k_shear=1., k_rot=1.):
    axial, shear, dRot = ligament_strains_linearized(*nodal_DOFs, reference_vector=reference_vector)
    l0 = torch.linalg.norm(reference_vector, dim=-1)
    return k_stretch * (axial * l0) ** 2 / 2 + k_shear * (shear * l0) ** 2 / 2 + k_rot * dRot ** 2 / 2


# ------------------------------------------------------------------ energy.py:49-66
def stretching_torsional_spring_energy(nodal_DOFs, k_stretch=1., k_rot=1.):
    DOFs1, DOFs2 = nodal_DOFs
    dU = DOFs2[:, :2] - DOFs1[:, :2]
    dRot = DOFs2[:, 2] - DOFs1[:, 2]
    return k_stretch * (dU * dU).sum(-1) / 2 + k_rot * dRot ** 2 / 2


# ------------------------------------------------------------------ energy.py:179-197 (+ jax-md smap.bond)
def strain_energy_bonds(node_displacements, bond_connectivity, bond_energy_fn, **bond_params):
    Ua = node_displacements[bond_connectivity[:, 0]]
    Ub = node_displacements[bond_connectivity[:, 1]]
    return bond_energy_fn((Ua, Ub), **bond_params).sum()


# ------------------------------------------------------------------ geometry.py:181-253
def _edge_unit_vectors(current_block_nodes, node_ids):
    n_sides = current_block_nodes.shape[1]
    blk, loc = node_ids // n_sides, node_ids % n_sides
    node = current_block_nodes[blk, loc]
    u1 = current_block_nodes[blk, (loc + 1) % n_sides] - node
    u1 = u1 / torch.linalg.norm(u1, dim=-1, keepdim=True)
    u2 = current_block_nodes[blk, (loc - 1) % n_sides] - node
    u2 = u2 / torch.linalg.norm(u2, dim=-1, keepdim=True)
    return u1, u2


def _angle_between(u1, u2):
    return torch.atan2(u1[:, 0] * u2[:, 1] - u1[:, 1] * u2[:, 0], u1[:, 0] * u2[:, 0] + u1[:, 1] * u2[:, 1])


def void_angles(current_block_nodes, bond_connectivity):
    """energy.py:204-219: the two void angles of every bond, concatenated (2*n_bonds,)."""
    b1n1, b1n2 = _edge_unit_vectors(current_block_nodes, bond_connectivity[:, 0])
    b2n1, b2n2 = _edge_unit_vectors(current_block_nodes, bond_connectivity[:, 1])
    va1 = _angle_between(b2n2, b1n1)
    va2 = _angle_between(b1n2, b2n1)
    return torch.cat([va1, va2])


# ------------------------------------------------------------------ energy.py:333-361
# energy.py:222-251, vectorised over a leading axis like the reference's vmap
def point_to_edge_distance(point, edge):
    x0, x1 = edge[..., 0, :], edge[..., 1, :]
    t = ((point - x0) * (x1 - x0)).sum(-1) / ((x1 - x0) * (x1 - x0)).sum(-1)
    inside = torch.sum((point - x0) ** 2 - (t[..., None] * (x1 - x0)) ** 2, -1) ** 0.5
    return torch.where((t >= 0) & (t <= 1), inside,
                       torch.where(t < 0, torch.sum((point - x0) ** 2, -1) ** 0.5, torch.sum((point - x1) ** 2, -1) ** 0.5))


# energy.py:255-275 (jnp.min: the cotangent is shared equally among tied minima, as torch.amin does)
def edges_distance(edge_1, edge_2):
    d = [point_to_edge_distance(edge_2[..., 0, :], edge_1), point_to_edge_distance(edge_2[..., 1, :], edge_1),
         point_to_edge_distance(edge_1[..., 0, :], edge_2), point_to_edge_distance(edge_1[..., 1, :], edge_2)]
    return torch.amin(torch.stack(d, -1), -1)


# energy.py:282-328
def void_edge_distance(current_block_nodes, bond_connectivity):
    npb = current_block_nodes.shape[1]
    n1, n2 = bond_connectivity[:, 0], bond_connectivity[:, 1]
    at = lambda n, shift: current_block_nodes[n // npb, (n + shift) % npb]  # noqa: E731
    pts1, pts1_prev, pts1_next = at(n1, 0), at(n1, -1), at(n1, 1)
    pts2, pts2_prev, pts2_next = at(n2, 0), at(n2, -1), at(n2, 1)
    d1 = edges_distance(torch.stack((pts1, pts1_next), 1), torch.stack((pts2, pts2_prev), 1))
    d2 = edges_distance(torch.stack((pts1, pts1_prev), 1), torch.stack((pts2, pts2_next), 1))
    return torch.cat((d1, d2))


def contact_energy(current_void_angles, min_angle, cutoff_angle, k_contact):
    x = (current_void_angles - cutoff_angle) / (cutoff_angle - min_angle)
    inner = k_contact / 4 * (cutoff_angle - min_angle) ** 2 * ((x + 1) ** -1 - (x - 1) ** -1 - 2)
    zero = torch.zeros_like(inner)
    return torch.where(current_void_angles < min_angle, zero,
                       torch.where(current_void_angles < cutoff_angle, inner, zero))


# ------------------------------------------------------------------ problem description
class Problem:
    """Static description of one solver instance (what `setup_dynamic_solver` closes over).

    bond_connectivity (n_bonds,2) int64; free/constrained ids per geometry.py:163-178;
    `bond_energy`: 'ligament' | 'linearized' | 'spring'; `use_contact`; `constrained_DOFs_fn(t, **cp)`
    torch callable returning scalar or (n_constrained,); `loading_fn(state, t, **lp)`;
    loaded ids / damped ids global DOF numbers.
    """

    def __init__(self, n_blocks, n_npb, bond_connectivity, constrained_DOF_ids, bond_energy="ligament",
                 use_contact=False, constrained_DOFs_fn=None, loaded_DOF_ids=None, loading_fn=None,
                 damped_blocks=None):
        self.n_blocks, self.n_npb = n_blocks, n_npb
        self.bonds = torch.as_tensor(bond_connectivity, dtype=torch.int64)
        self.constrained = torch.as_tensor(constrained_DOF_ids, dtype=torch.int64)
        mask = torch.ones(3 * n_blocks, dtype=torch.bool)
        mask[self.constrained] = False
        self.free = torch.nonzero(mask)[:, 0]
        self.bond_energy_fn = {"ligament": ligament_energy, "linearized": ligament_energy_linearized,
                               "spring": stretching_torsional_spring_energy}[bond_energy]
        self.spring = bond_energy == "spring"
        self.use_contact = use_contact
        self.constrained_DOFs_fn = constrained_DOFs_fn or (lambda t, **kw: torch.zeros((), dtype=F64))
        self.loaded = None if loaded_DOF_ids is None else torch.as_tensor(loaded_DOF_ids, dtype=torch.int64)
        self.loading_fn = loading_fn
        self.damped_DOF_ids = None
        if damped_blocks is not None:
            db = torch.as_tensor(damped_blocks, dtype=torch.int64)
            self.damped_DOF_ids = (db[:, None] * 3 + torch.arange(3)[None]).reshape(-1)
            self.n_damped = len(db)

    # kinematics.py:56-79
    def kinematics(self, free_DOFs, t, constraint_params):
        all_DOFs = torch.zeros(3 * self.n_blocks, dtype=F64)
        if len(self.constrained):
            val = self.constrained_DOFs_fn(t, **constraint_params)
            all_DOFs = all_DOFs.index_put((self.constrained,),
                                          (val * torch.ones(len(self.constrained), dtype=F64)))
        all_DOFs = all_DOFs.index_put((self.free,), free_DOFs)
        return all_DOFs.reshape(self.n_blocks, 3)

    # energy.py:425-447, 381-405, 462-468
    def energy(self, block_displacement, P):
        cnv = P["centroid_node_vectors"]
        node_disp = block_to_node_kinematics(block_displacement, cnv)
        if self.spring:  # StretchingTorsionalSpringParams: k_stretch, k_rot only (utils.py:80-91)
            E = strain_energy_bonds(node_disp.reshape(-1, 3), self.bonds, self.bond_energy_fn,
                                    k_stretch=P["k_stretch"], k_rot=P["k_rot"])
        else:
            E = strain_energy_bonds(node_disp.reshape(-1, 3), self.bonds, self.bond_energy_fn,
                                    k_stretch=P["k_stretch"], k_shear=P["k_shear"], k_rot=P["k_rot"],
                                    reference_vector=P["reference_vector"])
        if self.use_contact:
            current = P["block_centroids"][:, None] + cnv + node_disp[:, :, :2]
            gaps = void_edge_distance(current, self.bonds) if self.use_contact == "distance" else void_angles(current, self.bonds)
            E = E + contact_energy(gaps, P["min_angle"], P["cutoff_angle"], P["k_contact"]).sum()
        return E

    # dynamics.py:33-55 ; loading.py:36-45, 96-104
    def rhs(self, y, t, P, create_graph):
        nf = len(self.free)
        u, v = y[:nf], y[nf:]
        if not u.requires_grad:
            u = u.detach().requires_grad_(True)
        E = self.energy(self.kinematics(u, t, P["constraint_params"]), P)
        force = -torch.autograd.grad(E, u, create_graph=create_graph)[0]
        load = torch.zeros(nf, dtype=F64)
        if self.loaded is not None and self.loading_fn is not None:
            full = torch.zeros(3 * self.n_blocks, dtype=F64)
            val = self.loading_fn(y, t, **P["loading_params"])
            full = full.index_put((self.loaded,), val * torch.ones(len(self.loaded), dtype=F64))
            load = load + full[self.free]
        if self.damped_DOF_ids is not None:
            full = torch.zeros(3 * self.n_blocks, dtype=F64)
            dvals = (P["damping"] * torch.ones((self.n_damped, 3), dtype=F64)).reshape(-1)
            full = full.index_put((self.damped_DOF_ids,), dvals)
            load = load - full[self.free] * v
        return torch.cat([v, (force + load) / P["inertia"]])


def flatten_leaves(P):
    """Differentiable leaves of the `args` pytree, in a fixed order.  Returns (names, tensors)."""
    names, leaves = [], []
    for k in sorted(P):
        if k in ("constraint_params", "loading_params"):
            for kk in sorted(P[k]):
                names.append(f"{k}.{kk}")
                leaves.append(P[k][kk])
        else:
            names.append(k)
            leaves.append(P[k])
    return names, leaves


def _rebuild(P, names, leaves):
    out = {"constraint_params": {}, "loading_params": {}}
    for n, l in zip(names, leaves):
        if "." in n:
            a, b = n.split(".")
            out[a][b] = l
        else:
            out[n] = l
    return out


# ------------------------------------------------------------------ jax.experimental.ode
ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
         -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def initial_step_size(fun, t0, y0, order, rtol, atol, f0, variant="sum"):
    scale = atol + y0.abs() * rtol
    d0 = torch.linalg.norm(y0 / scale).item()
    d1 = torch.linalg.norm(f0 / scale).item()
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = fun(y1, t0 + h0)
    d2 = torch.linalg.norm((f1 - f0) / scale).item() / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        # jax 0.4.8: (0.01 / jnp.max(d1 + d2)) ** (1/(order+1)); later jax: max(d1, d2)
        D = (d1 + d2) if variant == "sum" else max(d1, d2)
        h1 = (0.01 / D) ** (1. / (order + 1.))
    return min(100. * h0, h1)


def runge_kutta_step(func, y0, f0, t0, dt):
    k = [f0]
    for i in range(6):
        yi = y0 + dt * sum(b * kj for b, kj in zip(BETA[i], k) if b != 0)
        k.append(func(yi, t0 + dt * ALPHA[i]))
    y1 = dt * sum(c * kj for c, kj in zip(C_SOL, k) if c != 0) + y0
    y1_err = dt * sum(c * kj for c, kj in zip(C_ERR, k) if c != 0)
    return y1, k[-1], y1_err, k


def odeint_flat(func, y0, ts, rtol, atol, variant="sum", stats=None):
    """`_odeint` of jax.experimental.ode on a flat float64 vector; `func(y, t)` detached."""
    t = float(ts[0])
    f = func(y0, t)
    dt = initial_step_size(func, t, y0, 4, rtol, atol, f, variant)
    y, last_t = y0, t
    coeff = [y0] * 5
    outs = [y0]
    nsteps = naccept = 0
    for target in [float(x) for x in ts[1:]]:
        while t < target and dt > 0:
            y1, f1, err, k = runge_kutta_step(func, y, f, t, dt)
            tol = atol + rtol * torch.maximum(y.abs(), y1.abs())
            ratio = math.sqrt(torch.mean((err / tol) ** 2).item())
            y_mid = y + dt * sum(c * kj for c, kj in zip(C_MID, k) if c != 0)
            dy0, dy1 = k[0], k[-1]
            new_coeff = [
                -2. * dt * dy0 + 2. * dt * dy1 - 8. * y - 8. * y1 + 16. * y_mid,
                5. * dt * dy0 - 3. * dt * dy1 + 18. * y + 14. * y1 - 32. * y_mid,
                -4. * dt * dy0 + dt * dy1 - 11. * y - 5. * y1 + 16. * y_mid,
                dt * dy0, y]
            dfactor = 1.0 if ratio < 1 else 0.2
            if ratio == 0:
                new_dt = dt * 10.0
            else:
                new_dt = dt * min(10.0, max(ratio ** (-1.0 / 5.0) * 0.9, dfactor))
            nsteps += 1
            if ratio <= 1.:
                y, f, last_t, t, coeff = y1, f1, t, t + dt, new_coeff
                naccept += 1
            dt = max(new_dt, 0.)
        x = (target - last_t) / (t - last_t)
        out = coeff[0]
        for c in coeff[1:]:
            out = out * x + c
        outs.append(out)
    if stats is not None:
        stats["steps"] = stats.get("steps", 0) + nsteps
        stats["accepted"] = stats.get("accepted", 0) + naccept
    return torch.stack(outs)


def solve_forward(prob: Problem, y0, ts, P, rtol, atol, variant="sum", stats=None):
    """free-DOF solution (n_t, 2*n_free) -- `odeint(rhs, _state0, ...)` at dynamics.py:166."""
    def func(y, t):
        with torch.enable_grad():
            return prob.rhs(y.detach(), torch.tensor(t, dtype=F64), P, create_graph=False).detach()
    return odeint_flat(func, y0, ts, rtol, atol, variant, stats)


def solve_adjoint(prob: Problem, ys, ts, P, g, rtol, atol, variant="sum", stats=None):
    """`_odeint_rev`: returns (y0_bar, ts_bar, dict of leaf cotangents)."""
    names, leaves = flatten_leaves(P)
    leaves = [torch.as_tensor(l, dtype=F64) for l in leaves]
    sizes = [l.numel() for l in leaves]
    N = ys.shape[1]

    def vjp_all(y, t, ybar):
        with torch.enable_grad():
            yy = y.detach().requires_grad_(True)
            tt = torch.tensor(t, dtype=F64, requires_grad=True)
            ll = [l.detach().requires_grad_(True) for l in leaves]
            ydot = prob.rhs(yy, tt, _rebuild(P, names, ll), create_graph=True)
            grads = torch.autograd.grad(ydot, [yy, tt] + ll, grad_outputs=ybar, allow_unused=True)
        grads = [torch.zeros_like(x) if gx is None else gx for gx, x in zip(grads, [yy, tt] + ll)]
        return ydot.detach(), grads

    def aug(z, s):
        y, ybar = z[:N], z[N:2 * N]
        ydot, grads = vjp_all(y, -s, ybar)
        return torch.cat([-ydot, grads[0].reshape(-1), grads[1].reshape(1)] + [gq.reshape(-1) for gq in grads[2:]])

    def plain(y, t):
        with torch.enable_grad():
            return prob.rhs(y.detach(), torch.tensor(t, dtype=F64), P, create_graph=False).detach()

    n_t = len(ts)
    y_bar = g[-1].clone()
    t0_bar = torch.zeros(1, dtype=F64)
    args_bar = torch.zeros(sum(sizes), dtype=F64)
    ts_bar = [None] * n_t
    for i in range(n_t - 1, 0, -1):
        t_bar = torch.dot(plain(ys[i], float(ts[i])), g[i])
        ts_bar[i] = t_bar
        t0_bar = t0_bar - t_bar
        z0 = torch.cat([ys[i], y_bar, t0_bar, args_bar])
        z = odeint_flat(aug, z0, [-float(ts[i]), -float(ts[i - 1])], rtol, atol, variant, stats)[1]
        y_bar, t0_bar, args_bar = z[N:2 * N], z[2 * N:2 * N + 1], z[2 * N + 1:]
        y_bar = y_bar + g[i - 1]
    ts_bar[0] = t0_bar[0]
    out, off = {}, 0
    for n, l, sz in zip(names, leaves, sizes):
        out[n] = args_bar[off:off + sz].reshape(l.shape)
        off += sz
    return y_bar, torch.stack(ts_bar), out
