"""Design optimisation on top of the solver.

* `angle_constraints`, `edge_length_constraints`: the inequality constraints of the reference's OptimizationProblems
  (`problems/quads_focusing.py:473-544`), as differentiable torch functions of the design (Jacobians by autograd).
* `BatchedMMA`: many independent instances of the method of moving asymptotes advanced in lock-step, so that every
  optimiser iteration is ONE batched value-and-gradient call of the solver (the reference runs one nlopt `LD_MMA`
  instance per process, `problems/quads_focusing.py:546-652`).  Box constraints only -- the form the reference uses
  when `min_void_angle` / `min_edge_length` are left at None.  Written from Svanberg's conservative convex separable
  approximation scheme (SIAM J. Optim. 12, 2002) with MMA-type approximations; it is not nlopt's code, and its
  iterates are not claimed to match nlopt's ("parity unpinned": nlopt is not installed here).
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

    def __init__(self, evaluate: Callable, x0, lower=None, upper=None, maximize=True):
        self.evaluate = evaluate
        self.sign = -1.0 if maximize else 1.0
        self.x = torch.as_tensor(x0, dtype=_F64).clone()
        B, n = self.x.shape
        dev = self.x.device
        full = lambda v, d: torch.full((B, n), d, dtype=_F64, device=dev) if v is None else \
            torch.as_tensor(v, dtype=_F64, device=dev).expand(B, n).clone()
        self.lb, self.ub = full(lower, -np.inf), full(upper, np.inf)
        span = self.ub - self.lb
        self.sigma = torch.where(torch.isfinite(span), 0.5 * span, torch.ones_like(span))
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

    def run(self, n_evaluations: int):
        """total number of objective evaluations per instance, like nlopt's `maxeval`. -> (best_x, best_f)"""
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

    def run_optimization_mma(self, initial_guesses, n_iterations: int, lower_bound: Optional[float] = None,
                             upper_bound: Optional[float] = None):
        """Batched counterpart of `run_optimization_nlopt` without the angle / edge-length constraints: MMA, objective
        maximised, `n_iterations` evaluations per instance (`opt.set_maxeval`), scalar box bounds."""
        if self.forward_problem.geometry is None:  # flatten() reads the design shapes of the lowered geometry
            self.forward_problem.lower()
        x0 = self.flatten(initial_guesses)
        opt = BatchedMMA(self.objective_and_grad, x0, lower_bound, upper_bound, maximize=True)
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
            raise ImportError("nlopt is not installed; use run_optimization_mma (batched, box constraints) or drive "
                              "objective_and_grad / angle_constraints / edge_length_constraints from your own optimiser") from e
        raise NotImplementedError("with nlopt available, pass objective_and_grad to nlopt.opt(nlopt.LD_MMA, n) as the reference does")

    def compute_best_forward(self, index: Optional[int] = None):
        """forward solution of the best design (of instance `index`, default: the best instance)"""
        if self.best_designs is None:
            raise ValueError("No design has been optimized yet.")
        i = int(torch.argmax(self.best_objectives)) if index is None else index
        self.forward_solution = self.forward_problem.solve([d[i] for d in self.best_designs])
        return self.forward_solution
