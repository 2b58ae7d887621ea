"""Design optimisation on top of the solver.

* `angle_constraints`, `edge_length_constraints`: the inequality constraints of the reference's OptimizationProblems
  (`problems/quads_focusing.py:473-544`), as differentiable torch functions of the design (Jacobians by autograd).
* `BatchedMMA`: many independent instances of the method of moving asymptotes advanced in lock-step, so that every
  optimiser iteration is ONE batched value-and-gradient call of the solver (the reference runs one nlopt `LD_MMA`
  instance per process, `problems/quads_focusing.py:546-652`).  Box constraints, and -- with `constraints=` -- the
  inequality constraints the reference adds through `add_inequality_mconstraint` (`min_void_angle`, `min_block_angle`,
  `min_edge_length`): every instance solves the dual of its separable sub-problem, all instances at once, on the
  device, from the fixed-width sparse Jacobian of `geometry_device.DeviceConstraints`.  Written from Svanberg's
  conservative convex separable approximation scheme (SIAM J. Optim. 12, 2002) with MMA-type approximations; it is
  not nlopt's code, and its iterates are not claimed to match nlopt's ("parity unpinned": nlopt is not installed here).
* `OptimizationProblem`: the reference's driver object (objective values / design history, best forward solution).
"""
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from .geometry import compute_edge_angles, compute_edge_lengths, compute_edge_unit_vectors, angle_between_unit_vectors

_F64 = torch.float64
_TWO_PI = 2 * np.pi


# ------------------------------------------------------------------------------------------------
# constraints (<= 0 when satisfied)
# ------------------------------------------------------------------------------------------------
def quad_boundary_node_ids(n1_blocks: int, n2_blocks: int):
    """vertices on the outer boundary of a quad lattice (reference `quads_focusing.py:477-489`)."""
    n_blocks = n1_blocks * n2_blocks
    return np.concatenate([
        np.arange(n1_blocks) * 4 + 3,                                       # bottom edge
        np.arange(n1_blocks - 1, n_blocks, n1_blocks) * 4 + 0,              # right edge
        np.arange(n_blocks - 1, n_blocks - n1_blocks - 1, -1) * 4 + 1,      # top edge
        np.arange(0, n_blocks, n1_blocks) * 4 + 2,                          # left edge
    ]).astype(np.int64)


def angle_constraints(geometry, min_void_angle=0.0, min_block_angle=0.0, boundary_angle_constraint=False) -> Callable:
    """design -> concat(-(void_1 - min_void), -(void_2 - min_void), -(block_1 - min_block), -(block_2 - min_block)
    [, -(boundary block angles - min_block)]), angles taken mod 2 pi (reference `quads_focusing.py:473-531`)."""
    bonds = geometry.bond_connectivity()
    boundary = quad_boundary_node_ids(geometry.n1_blocks, geometry.n2_blocks) if boundary_angle_constraint else None

    def fn(design):
        nodes = geometry.centroid_node_vectors(*design)
        v1, v2, b1, b2 = [torch.remainder(a, _TWO_PI) for a in compute_edge_angles(nodes, bonds)]
        parts = [-(v1 - min_void_angle), -(v2 - min_void_angle), -(b1 - min_block_angle), -(b2 - min_block_angle)]
        if boundary is not None:
            e1, e2 = compute_edge_unit_vectors(nodes, boundary)
            parts.append(-(torch.remainder(angle_between_unit_vectors(e1, e2), _TWO_PI) - min_block_angle))
        return torch.cat(parts)

    return fn


def edge_length_constraints(geometry, min_edge_length) -> Callable:
    """design -> -(edge lengths - min_edge_length), flattened (reference `quads_focusing.py:533-544`)."""

    def fn(design):
        return -(compute_edge_lengths(geometry.centroid_node_vectors(*design)).reshape(-1) - min_edge_length)

    return fn


# ------------------------------------------------------------------------------------------------
# batched MMA (box constraints)
# ------------------------------------------------------------------------------------------------
class BatchedMMA:
    """B independent MMA instances over n variables each, advanced together.

    `evaluate(x: (B, n)) -> (f: (B,), grad: (B, n))` is called once per iteration for all instances.  Each instance is
    a small state machine: its candidate is either accepted (the separable approximation was conservative at the
    candidate; asymptote distances sigma adapt, rho relaxes) or rejected (rho grows, a more conservative candidate is
    built from the same expansion point).  One evaluation per instance per iteration either way.
    """

    def __init__(self, evaluate: Callable, x0, lower=None, upper=None, maximize=True, constraints: Optional[Callable] = None,
                 constraint_tolerance: float = 1e-8, dual_iterations: int = 200, initial_move: float = 0.5):
        """`constraints(x: (B, n)) -> (c: (B, m), jac: (B, m, K), columns: (m, K) int64)`: inequality constraints
        `c <= 0` with a fixed-width sparse Jacobian, `jac[b, i, k] = d c_i / d x[columns[i, k]]` (padding entries carry
        zeros).  `constraint_tolerance` is the feasibility tolerance (the reference passes 1e-8 to nlopt).
        `initial_move`: initial asymptote distance as a fraction of the bound span (nlopt's MMA uses 0.5; with wide bounds
        and steep objectives the first candidates then sit at the move limit, 0.45 x span away from the start)."""
        self.evaluate = evaluate
        self.constraints, self.ctol, self.dual_iterations = constraints, float(constraint_tolerance), int(dual_iterations)
        self.sign = -1.0 if maximize else 1.0
        self.x = torch.as_tensor(x0, dtype=_F64).clone()
        B, n = self.x.shape
        dev = self.x.device
        full = lambda v, d: torch.full((B, n), d, dtype=_F64, device=dev) if v is None else \
            torch.as_tensor(v, dtype=_F64, device=dev).expand(B, n).clone()
        self.lb, self.ub = full(lower, -np.inf), full(upper, np.inf)
        span = self.ub - self.lb
        self.sigma = torch.where(torch.isfinite(span), float(initial_move) * span, torch.ones_like(span))
        self.sigma_min = torch.where(torch.isfinite(span), 1e-8 * span, torch.full_like(span, 1e-12))
        self.sigma_max = torch.where(torch.isfinite(span), 10.0 * span, torch.full_like(span, np.inf))
        self.rho = torch.ones(B, dtype=_F64, device=dev)
        self.x_prev, self.x_prevprev = self.x.clone(), self.x.clone()
        self.n_accepted = torch.zeros(B, dtype=torch.int64, device=dev)
        self.n_evals = 0
        self.f = self.g = None           # value / gradient (of the minimised function) at the expansion point
        self.best_f = self.best_x = None  # in the caller's sense (maximised or minimised)
        self.history = []                 # per evaluation: (B,) objective values in the caller's sense

    # separable convex approximation around x: sum_j g_j dx_j + (|g_j| sigma_j + rho / 2) dx_j^2 / (sigma_j^2 - dx_j^2)
    def _candidate(self):
        x, g, sig = self.x, self.g, self.sigma
        u = g.abs() * sig + 0.5 * self.rho[:, None]
        lo = torch.maximum(self.lb - x, -0.9 * sig)
        hi = torch.minimum(self.ub - x, 0.9 * sig)
        s2 = sig * sig

        def slope(dx):
            return g + u * 2.0 * dx * s2 / (s2 - dx * dx) ** 2

        # the slope is increasing on (-sigma, sigma): bisection on its root, clipped to the move limits
        a, b = lo.clone(), hi.clone()
        sa, sb = slope(a), slope(b)
        for _ in range(60):
            m = 0.5 * (a + b)
            pos = slope(m) > 0
            b = torch.where(pos, m, b)
            a = torch.where(pos, a, m)
        dx = 0.5 * (a + b)
        dx = torch.where(sa >= 0, lo, dx)   # minimum at / below the lower move limit
        dx = torch.where(sb <= 0, hi, dx)   # minimum at / above the upper move limit
        frac = dx * dx / (s2 - dx * dx)
        approx = self.f + (g * dx + u * frac).sum(1)
        w = 0.5 * frac.sum(1)
        return x + dx, approx, w

    def _record(self, x, f_min):
        val = self.sign * f_min
        better = val > self.best_f if self.sign < 0 else val < self.best_f
        self.best_f = torch.where(better, val, self.best_f)
        self.best_x = torch.where(better[:, None], x, self.best_x)
        self.history.append(val.detach().cpu().clone())

    def _eval(self, x):
        f, g = self.evaluate(x)
        self.n_evals += 1
        return self.sign * torch.as_tensor(f, dtype=_F64, device=x.device), self.sign * torch.as_tensor(g, dtype=_F64, device=x.device)

    # ---- inequality constraints: dual of the separable sub-problem ------------------------------------------------
    # Every function is approximated around x by  F(x) + sum_j (G_j s_j^2 dx_j + (|G_j| s_j + rho_F / 2) dx_j^2) / (s_j^2 - dx_j^2)
    # (value and gradient exact at dx = 0, convex on |dx_j| < s_j).  For multipliers y >= 0 the Lagrangian is separable
    # and has the same form with U_j = s_j^2 (g_j + sum_i y_i J_ij), V_j = s_j (|g_j| + sum_i y_i |J_ij|) + (rho + y . rho_c) / 2;
    # its minimiser is dx_j = -s_j a_j / (1 + sqrt(1 - a_j^2)), a_j = U_j / (V_j s_j), clipped to the move limits.  The
    # dual function W(y) is concave with gradient c~_i(dx(y)) (the approximated constraints at the minimiser).
    def _sub_primal(self, y):
        B, n = self.x.shape
        yk = (y[:, :, None] * self._ones_k).reshape(B, -1)
        lin = torch.zeros((B, n), dtype=_F64, device=y.device).scatter_add_(1, self._cidx, yk * self._J)
        mag = torch.zeros((B, n), dtype=_F64, device=y.device).scatter_add_(1, self._cidx, yk * self._absJ)
        sig = self.sigma
        V = sig * (self.g.abs() + mag) + 0.5 * (self.rho + (y * self.rho_c).sum(1))[:, None]
        a = (sig * (self.g + lin) / V).clamp(-1.0, 1.0)
        dx = -sig * a / (1.0 + torch.sqrt(1.0 - a * a))
        return torch.minimum(torch.maximum(dx, self._lo), self._hi)

    def _approximations(self, dx):
        """-> objective approximation (B,), constraint approximations (B, m), w = 1/2 sum_j dx_j^2 / (s_j^2 - dx_j^2) (B,)"""
        B = dx.shape[0]
        sig, s2 = self.sigma, self.sigma * self.sigma
        den = 1.0 / (s2 - dx * dx)
        t1, t2 = s2 * dx * den, dx * dx * den           # linear / curvature shape functions per variable
        w = 0.5 * t2.sum(1)
        approx_f = self.f + (self.g * t1 + self.g.abs() * sig * t2).sum(1) + self.rho * w
        rows = (self._J * t1.gather(1, self._cidx) + self._absJ * (sig * t2).gather(1, self._cidx)).reshape(B, -1, self._K).sum(2)
        approx_c = self.c + rows + self.rho_c * w[:, None]
        return approx_f, approx_c, w

    def _candidate_constrained(self):
        x, sig = self.x, self.sigma
        B = x.shape[0]
        self._lo = torch.maximum(self.lb - x, -0.9 * sig)
        self._hi = torch.minimum(self.ub - x, 0.9 * sig)
        # projected gradient ascent on W(y), y >= 0, Barzilai-Borwein steps per instance; the best iterate is kept
        y = self.y.clone()
        dx = self._sub_primal(y)
        _, grad, _ = self._approximations(dx)
        scale = self._absJ.reshape(B, -1, self._K).sum(2).square().sum(1).clamp_min(1e-300)
        alpha = (1.0 / (scale * (sig * sig).mean(1))).clamp(1e-12, 1e12)  # ~ 1 / |d grad / d y|
        best_y, best_viol = y.clone(), torch.full((B,), np.inf, dtype=_F64, device=x.device)
        for _ in range(self.dual_iterations):
            # merit for choosing the iterate: complementarity / feasibility residual |y - max(0, y + c~)|
            res = (y - torch.clamp_min(y + grad, 0.0)).abs().amax(1)
            better = res < best_viol
            best_viol = torch.where(better, res, best_viol)
            best_y = torch.where(better[:, None], y, best_y)
            y_new = torch.clamp_min(y + alpha[:, None] * grad, 0.0)
            dx = self._sub_primal(y_new)
            _, grad_new, _ = self._approximations(dx)
            sy, sg = y_new - y, grad_new - grad
            bb = (sy * sy).sum(1) / (-(sy * sg).sum(1)).clamp_min(1e-300)
            alpha = torch.where(torch.isfinite(bb) & (bb > 0), bb, alpha).clamp(1e-12, 1e12)
            y, grad = y_new, grad_new
        res = (y - torch.clamp_min(y + grad, 0.0)).abs().amax(1)
        best_y = torch.where((res < best_viol)[:, None], y, best_y)
        self.y = best_y
        dx = self._sub_primal(best_y)
        approx_f, approx_c, w = self._approximations(dx)
        # safeguard for an inexact dual solution: where the expansion point is feasible, shrink the step along the ray
        # until every approximated constraint holds (each is convex along the ray and <= 0 at its start)
        start_ok = (self.c <= self.ctol).all(1)
        bad = start_ok & (approx_c > self.ctol).any(1)
        if bool(bad.any()):
            lo_t = torch.zeros(B, dtype=_F64, device=x.device)
            hi_t = torch.ones(B, dtype=_F64, device=x.device)
            for _ in range(40):
                mid = 0.5 * (lo_t + hi_t)
                ok = (self._approximations(dx * mid[:, None])[1] <= self.ctol).all(1)
                lo_t = torch.where(ok, mid, lo_t)
                hi_t = torch.where(ok, hi_t, mid)
            dx = torch.where(bad[:, None], dx * lo_t[:, None], dx)
            approx_f, approx_c, w = self._approximations(dx)
        return x + dx, approx_f, approx_c, w

    def _eval_constraints(self, x):
        c, jac, cols = self.constraints(x)
        B = x.shape[0]
        c = torch.as_tensor(c, dtype=_F64, device=x.device)
        jac = torch.as_tensor(jac, dtype=_F64, device=x.device).reshape(B, c.shape[1], -1)
        if getattr(self, "_K", None) is None:
            self._K = jac.shape[2]
            cols = torch.as_tensor(cols, dtype=torch.int64, device=x.device).reshape(c.shape[1], self._K)
            self._cidx = cols.reshape(1, -1).expand(B, -1)
            self._ones_k = torch.ones(self._K, dtype=_F64, device=x.device)
        return c, jac.reshape(B, -1)

    def _run_constrained(self, n_evaluations: int):
        dev = self.x.device
        B = self.x.shape[0]
        if self.f is None:
            self.x = torch.minimum(torch.maximum(self.x, self.lb), self.ub)
            self.f, self.g = self._eval(self.x)
            self.c, self._J = self._eval_constraints(self.x)
            self._absJ = self._J.abs()
            m = self.c.shape[1]
            self.rho_c = torch.ones((B, m), dtype=_F64, device=dev)
            self.y = torch.zeros((B, m), dtype=_F64, device=dev)
            viol = self.c.amax(1)
            self.best_f, self.best_x, self.best_violation = (self.sign * self.f).clone(), self.x.clone(), viol.clone()
            self.history.append((self.sign * self.f).detach().cpu().clone())
            self.violation_history = [viol.detach().cpu().clone()]
        while self.n_evals < n_evaluations:
            cand, approx_f, approx_c, w = self._candidate_constrained()
            fc, gc = self._eval(cand)
            cc, Jc = self._eval_constraints(cand)
            viol = cc.amax(1)
            val = self.sign * fc
            self.history.append(val.detach().cpu().clone())
            self.violation_history.append(viol.detach().cpu().clone())
            # best design: feasible beats infeasible, then the objective; among infeasible ones the smaller violation
            feas, best_feas = viol <= self.ctol, self.best_violation <= self.ctol
            improves = val > self.best_f if self.sign < 0 else val < self.best_f
            better = torch.isfinite(fc) & ((feas & ~best_feas) | (feas & best_feas & improves) | (~feas & ~best_feas & (viol < self.best_violation)))
            self.best_f = torch.where(better, val, self.best_f)
            self.best_x = torch.where(better[:, None], cand, self.best_x)
            self.best_violation = torch.where(better, viol, self.best_violation)
            # conservative at the candidate?  objective and every constraint
            tol_f = 1e-12 * fc.abs()
            ok_f = (approx_f >= fc - tol_f) | ~torch.isfinite(approx_f)
            ok_c = approx_c >= cc - 1e-12 * cc.abs() - 1e-14
            ok = ok_f & ok_c.all(1) & torch.isfinite(fc)
            wc = w.clamp_min(1e-300)
            grow = torch.minimum(10.0 * self.rho, 1.1 * (self.rho + (fc - approx_f) / wc))
            grow = torch.where(torch.isfinite(grow), grow, 10.0 * self.rho)
            self.rho = torch.where(ok, torch.clamp(0.1 * self.rho, min=1e-5), torch.where(ok_f, self.rho, grow))
            grow_c = torch.minimum(10.0 * self.rho_c, 1.1 * (self.rho_c + (cc - approx_c) / wc[:, None]))
            grow_c = torch.where(torch.isfinite(grow_c), grow_c, 10.0 * self.rho_c)
            self.rho_c = torch.where(ok[:, None], torch.clamp(0.1 * self.rho_c, min=1e-5), torch.where(ok_c, self.rho_c, grow_c))
            okc = ok[:, None]
            osc = (cand - self.x) * (self.x - self.x_prev)
            gamma = torch.where(osc < 0, 0.7, torch.where(osc > 0, 1.2, 1.0))
            adapt = okc & (self.n_accepted >= 1)[:, None]
            self.sigma = torch.where(adapt, torch.minimum(torch.maximum(self.sigma * gamma, self.sigma_min), self.sigma_max), self.sigma)
            self.x_prevprev = torch.where(okc, self.x_prev, self.x_prevprev)
            self.x_prev = torch.where(okc, self.x, self.x_prev)
            self.x = torch.where(okc, cand, self.x)
            self.f = torch.where(ok, fc, self.f)
            self.g = torch.where(okc, gc, self.g)
            self.c = torch.where(okc, cc, self.c)
            self._J = torch.where(okc, Jc, self._J)
            self._absJ = self._J.abs()
            self.n_accepted = self.n_accepted + ok.to(torch.int64)
        return self.best_x, self.best_f

    def run(self, n_evaluations: int):
        """total number of objective evaluations per instance, like nlopt's `maxeval`. -> (best_x, best_f)"""
        if self.constraints is not None:
            return self._run_constrained(n_evaluations)
        if self.f is None:
            self.x = torch.minimum(torch.maximum(self.x, self.lb), self.ub)
            self.f, self.g = self._eval(self.x)
            self.best_f, self.best_x = (self.sign * self.f).clone(), self.x.clone()
            self.history.append((self.sign * self.f).detach().cpu().clone())
        while self.n_evals < n_evaluations:
            cand, approx, w = self._candidate()
            fc, gc = self._eval(cand)
            self._record(cand, fc)
            ok = (approx >= fc - 1e-12 * fc.abs()) | ~torch.isfinite(approx)
            ok = ok & torch.isfinite(fc)
            # rejected: more conservative approximation around the same point
            grow = torch.minimum(10.0 * self.rho, 1.1 * (self.rho + (fc - approx) / w.clamp_min(1e-300)))
            grow = torch.where(torch.isfinite(grow), grow, 10.0 * self.rho)
            self.rho = torch.where(ok, torch.clamp(0.1 * self.rho, min=1e-5), grow)
            # accepted: move, adapt the asymptote distances from the sign pattern of the last two moves
            okc = ok[:, None]
            osc = (cand - self.x) * (self.x - self.x_prev)
            gamma = torch.where(osc < 0, 0.7, torch.where(osc > 0, 1.2, 1.0))
            adapt = okc & (self.n_accepted >= 1)[:, None]
            self.sigma = torch.where(adapt, torch.minimum(torch.maximum(self.sigma * gamma, self.sigma_min), self.sigma_max), self.sigma)
            self.x_prevprev = torch.where(okc, self.x_prev, self.x_prevprev)
            self.x_prev = torch.where(okc, self.x, self.x_prev)
            self.x = torch.where(okc, cand, self.x)
            self.f = torch.where(ok, fc, self.f)
            self.g = torch.where(okc, gc, self.g)
            self.n_accepted = self.n_accepted + ok.to(torch.int64)
        return self.best_x, self.best_f


# ------------------------------------------------------------------------------------------------
# driver object
# ------------------------------------------------------------------------------------------------
class OptimizationProblem:
    """Counterpart of the reference's OptimizationProblem classes (`problems/quads_focusing.py:376-680`) for the
    problems of `difflexmm_b200.problems`: objective = target kinetic energy (maximised), history of objective values
    and designs, best forward solution.  `run_optimization_mma` optimises a whole batch of initial guesses at once."""

    def __init__(self, forward_problem, name="optimization"):
        self.forward_problem = forward_problem
        self.name = name
        self.objective_values, self.design_values = [], []
        self.best_designs = self.best_objectives = self.forward_solution = None

    # design tuple <-> flat vector (reference: jax.flatten_util.ravel_pytree)
    def _shapes(self):
        return [tuple(s) for s in self.forward_problem.geometry.design_shapes]

    def flatten(self, design: Sequence[torch.Tensor]):
        shapes = self._shapes()
        parts = [torch.as_tensor(d, dtype=_F64) for d in design]
        batched = parts[0].dim() == len(shapes[0]) + 1
        B = parts[0].shape[0] if batched else 1
        return torch.cat([p.reshape(B, -1) for p in parts], dim=1)

    def unflatten(self, x):
        shapes, out, o = self._shapes(), [], 0
        for s in shapes:
            n = int(np.prod(s))
            out.append(x[:, o:o + n].reshape(x.shape[0], *s))
            o += n
        return out

    def objective_and_grad(self, x):
        """(B, n) flat designs -> objective (B,), gradient (B, n): one batched forward + adjoint launch of libdfx."""
        P = self.forward_problem
        dev = P.solver.device
        design = [d.to(dev).requires_grad_(True) for d in self.unflatten(x)]
        J = P.target_kinetic_energy(design, batch=x.shape[0], fused=True)
        J.sum().backward()
        grad = torch.cat([d.grad.reshape(x.shape[0], -1) for d in design], dim=1)
        return J.detach().to(x.device), grad.to(x.device)

    def device_constraints(self, min_void_angle=None, min_block_angle=None, min_edge_length=None, boundary_angle_constraint=False):
        """`geometry_device.DeviceConstraints` of this problem's lattice (values + sparse Jacobian in one launch)."""
        from .geometry_device import DeviceConstraints
        if self.forward_problem.geometry is None:
            self.forward_problem.lower()
        return DeviceConstraints(self.forward_problem.device_geometry(), min_void_angle, min_block_angle, min_edge_length,
                                 boundary_angle_constraint)

    def run_optimization_mma(self, initial_guesses, n_iterations: int, lower_bound: Optional[float] = None,
                             upper_bound: Optional[float] = None, min_void_angle: Optional[float] = None,
                             min_block_angle: Optional[float] = None, min_edge_length: Optional[float] = None,
                             boundary_angle_constraint=False, initial_move: float = 0.5):
        """Batched counterpart of `run_optimization_nlopt` (`problems/quads_focusing.py:546-652`): MMA, objective
        maximised, `n_iterations` evaluations per instance (`opt.set_maxeval`), scalar box bounds, and the reference's
        inequality constraints under the same switches (angle constraints when both `min_void_angle` and
        `min_block_angle` are given, edge-length constraints when `min_edge_length` is given; tolerance 1e-8)."""
        if self.forward_problem.geometry is None:  # flatten() reads the design shapes of the lowered geometry
            self.forward_problem.lower()
        x0 = self.flatten(initial_guesses)
        constraints = None
        self.constraints_violation = {"angles": [], "edge_lengths": []}
        if (min_void_angle is not None and min_block_angle is not None) or min_edge_length is not None:
            dc = self.device_constraints(min_void_angle, min_block_angle, min_edge_length, boundary_angle_constraint)
            x0 = x0.to(dc.dg.device)

            def constraints(x):
                c, jac = dc(x)
                if dc.n_angle_rows:
                    self.constraints_violation["angles"].append(c[:, :dc.n_angle_rows].amax(1).cpu().numpy())
                if dc.n_rows > dc.n_angle_rows:
                    self.constraints_violation["edge_lengths"].append(c[:, dc.n_angle_rows:].amax(1).cpu().numpy())
                return c, jac, dc.scalar_columns

        opt = BatchedMMA(self.objective_and_grad, x0, lower_bound, upper_bound, maximize=True, constraints=constraints,
                         initial_move=initial_move)
        best_x, best_f = opt.run(n_iterations)
        self.objective_values = [h.numpy() for h in opt.history]
        self.best_designs, self.best_objectives = self.unflatten(best_x), best_f
        self.optimizer = opt
        return self.best_designs, best_f

    def run_optimization_nlopt(self, *args, **kwargs):
        """The reference's single-instance nlopt loop needs the `nlopt` package, which this image does not ship."""
        try:
            import nlopt  # noqa: F401
        except ImportError as e:
            raise ImportError("nlopt is not installed; use run_optimization_mma (batched, box + angle / edge-length constraints) or drive "
                              "objective_and_grad / angle_constraints / edge_length_constraints from your own optimiser") from e
        raise NotImplementedError("with nlopt available, pass objective_and_grad to nlopt.opt(nlopt.LD_MMA, n) as the reference does")

    def compute_best_forward(self, index: Optional[int] = None):
        """forward solution of the best design (of instance `index`, default: the best instance)"""
        if self.best_designs is None:
            raise ValueError("No design has been optimized yet.")
        i = int(torch.argmax(self.best_objectives)) if index is None else index
        self.forward_solution = self.forward_problem.solve([d[i] for d in self.best_designs])
        return self.forward_solution
