"""`setup_dynamic_solver` -- the reference's solver factory, backed by libdfx (CUDA, sm_100a).

Same signature and return protocol as the reference (`difflexmm/dynamics.py:60-186`):

    solve_dynamics = setup_dynamic_solver(geometry, energy_fn, loaded_block_DOF_pairs, loading_fn,
                                          constrained_block_DOF_pairs, constrained_DOFs_fn,
                                          damped_blocks, rtol, atol)
    fields = solve_dynamics(state0, timepoints, control_params)   # (n_t, 2, n_blocks, 3)

The call `odeint(rhs, _state0, timepoints, control_params, _inertia, rtol, atol)` at
`dynamics.py:166` is replaced by `dfx_forward`; differentiating through the solver (the
reference relies on odeint's `custom_vjp`) is `dfx_adjoint`, wired in with a
`torch.autograd.Function` (the torch analogue of `jax.custom_vjp`; the `jax.ffi` binding of the
same C ABI is described in INTEGRATION.md).  Everything before (geometry maps, `compute_inertia`)
and after (field reconstruction, objectives) the boundary stays ordinary differentiable torch.

Extension over the reference: `solve_dynamics.batch(state0, timepoints, control_params, batch=B)`
integrates B designs in one launch (leaves may carry a leading batch axis) -- the design batch is
the GPU dimension of this path.  There is no CPU path: without the CUDA
library the solver raises.
"""

from typing import Optional

import numpy as np
import torch

from . import _abi
from .energy import BlockEnergy
from .geometry import Geometry, compute_inertia
from .loading import DriveSignal, LoadSignal, zero_drive
from .utils import ControlParams

_F64 = torch.float64


def _np_int(a):
    if a is None:
        return np.zeros((0,), dtype=np.int64)
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a).astype(np.int64)


def lower_topology(geometry: Geometry, energy_fn, loaded_block_DOF_pairs=None, loading_fn=None,
                   constrained_block_DOF_pairs=(), constrained_DOFs_fn=None, damped_blocks=None):
    """Everything `setup_dynamic_solver` closes over, lowered to the static topology of libdfx
    (reference `dynamics.py:91-126`, `kinematics.py:52-54`, `loading.py:30-34,88-92`).
    Returns (TopologySpec, DriveSignal)."""
    if not isinstance(energy_fn, BlockEnergy):
        raise TypeError(
            "energy_fn must be built with difflexmm_b200.energy.build_strain_energy / build_contact_energy / "
            "combine_block_energies: the CUDA solver has no path for arbitrary Python energies")
    pairs = _np_int(constrained_block_DOF_pairs).reshape(-1, 2)
    constrained = pairs[:, 0] * 3 + pairs[:, 1]
    if constrained_DOFs_fn is None:
        drive = zero_drive()
    elif isinstance(constrained_DOFs_fn, DriveSignal):
        drive = constrained_DOFs_fn
    else:
        raise TypeError(
            "constrained_DOFs_fn must be a difflexmm_b200.loading drive signal (pulse_drive, harmonic_drive, "
            "ramp_drive, static_pulse_drive, tabulated_drive) or None; arbitrary Python callables cannot run inside the CUDA solver")
    load_kind, loaded, load_vec, load_consts = _abi.DFX_LOAD_NONE, (), None, ()
    if loaded_block_DOF_pairs is not None and loading_fn is not None:
        if not isinstance(loading_fn, LoadSignal):
            raise TypeError("loading_fn must be a difflexmm_b200.loading load signal (ramp_load, sech2_load)")
        lp = _np_int(loaded_block_DOF_pairs).reshape(-1, 2)
        loaded = lp[:, 0] * 3 + lp[:, 1]
        load_kind, load_vec, load_consts = loading_fn.kind, loading_fn.load_vector, loading_fn.consts
    spec = _abi.TopologySpec(
        n_blocks=geometry.n_blocks, n_npb=geometry.n_npb, bond_nodes=energy_fn.bond_connectivity,
        constrained_dofs=constrained, bond_energy=energy_fn.bond_kind, contact=energy_fn.contact,
        drive_kind=drive.kind, drive_vec0=drive.vec0, drive_vec1=drive.vec1,
        drive_table=(drive.times, drive.values) if drive.kind == _abi.DFX_DRIVE_TABLE else None,
        load_kind=load_kind, loaded_dofs=loaded, load_vec=load_vec, load_consts=load_consts,
        damped_blocks=_np_int(damped_blocks) if damped_blocks is not None else ())
    return spec, drive


def _as_t(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=_F64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=device)


def lower_params(spec, drive, control_params: ControlParams, batch: Optional[int], device, per_bond=(), inertia_full=None):
    """ControlParams -> leaves of libdfx (differentiable torch) + size of the reference's
    augmented adjoint state for one design.

    Batch rule (batch is not None): a leaf carries the batch axis iff its ndim is one more than
    its base ndim.  Base ndims: centroid_node_vectors 3, block_centroids 2, reference_vector 2,
    inertia 2, k_* 0 (1 if named in `per_bond`), density 0, contact/drive/loading scalars 0,
    damping 0 or 2.  Without a batch, k_* of ndim 1 are per-bond leaves as in the reference
    (`utils.py:68-71`).
    """
    gp, mp = control_params.geometrical_params, control_params.mechanical_params
    bp = mp.bond_params
    B = batch
    nb = spec.n_blocks
    n_entries = 0  # per-design entries of (control_params, _inertia)

    def split(x, base_ndim):
        """-> (tensor, is_batched, per-design numel)"""
        t = _as_t(x, device)
        batched = B is not None and t.dim() == base_ndim + 1
        if batched and t.shape[0] != B:
            raise ValueError(f"leading axis {t.shape[0]} of a batched leaf does not match batch={B}")
        return t, batched, (t.numel() // B if batched else t.numel())

    leaves = {}
    cnv, cnv_b, n = split(gp.centroid_node_vectors, 3)
    leaves["centroid_node_vectors"] = cnv
    n_entries += n
    if gp.block_centroids is not None:
        cen, _, n = split(gp.block_centroids, 2)
        n_entries += n
        if spec.contact == _abi.DFX_CONTACT_DISTANCE:  # the only energy that reads block positions (energy.py:395-405)
            leaves["block_centroids"] = cen
    elif spec.contact == _abi.DFX_CONTACT_DISTANCE:
        raise ValueError("distance-based contact needs geometrical_params.block_centroids")
    spring = spec.bond_energy == _abi.DFX_BOND_SPRING
    if spring != (not hasattr(bp, "reference_vector")):
        raise TypeError("stretching_torsional_spring_energy takes StretchingTorsionalSpringParams, the ligament energies "
                        "take LigamentParams (reference utils.py:62-94)")
    if spring:
        # the reference pytree has neither leaf; libdfx wants valid pointers and ignores them (include/dfx.h)
        leaves["reference_vector"] = torch.tensor([[1.0, 0.0]], dtype=_F64, device=device).repeat(spec.n_bonds, 1)
    else:
        leaves["reference_vector"], _, n = split(bp.reference_vector, 2)
        n_entries += n
    pb = []
    for name in ("k_stretch", "k_shear", "k_rot"):
        if spring and name == "k_shear":
            leaves[name] = torch.zeros((), dtype=_F64, device=device)
            continue
        k = _as_t(getattr(bp, name), device)
        is_pb = (name in per_bond) if B is not None else (k.dim() == 1)
        if is_pb:
            pb.append(name)
        leaves[name], _, n = split(k, 1 if is_pb else 0)
        n_entries += n
    damping_per_dof = False
    n_damped = len(spec.damped_blocks)
    if mp.damping is not None:
        d = _as_t(mp.damping, device)
        if d.dim() - (1 if (B is not None and d.dim() in (1, 3)) else 0) == 0:
            d, _, n = split(d, 0)
        else:
            # the reference multiplies the leaf by ones((n_damped, 3)) (loading.py:93-101): anything that broadcasts against
            # that shape is valid -- (3,), (1, 3), (n_damped, 1).  The leaf keeps its own size in the augmented state
            # (aug_size counts its elements); the kernel receives the broadcast (n_damped, 3) array and autograd sums the
            # cotangent back to the leaf's shape.
            if B is None and n_damped and tuple(d.shape) != (n_damped, 3):
                try:
                    n = d.numel()
                    d = d * torch.ones((n_damped, 3), dtype=_F64, device=d.device)
                except RuntimeError:
                    raise ValueError(f"damping must be a scalar or broadcast against (n_damped_blocks, 3) = ({n_damped}, 3); "
                                     f"got {tuple(_as_t(mp.damping, device).shape)}") from None
                if tuple(d.shape) != (n_damped, 3):
                    raise ValueError(f"damping must be a scalar or broadcast against (n_damped_blocks, 3) = ({n_damped}, 3)")
            else:
                d, db, n = split(d, 2)
                if n_damped and tuple(d.shape[-2:]) != (n_damped, 3):
                    raise ValueError(f"damping must be a scalar or of shape (n_damped_blocks, 3) = ({n_damped}, 3)")
            damping_per_dof = True
        n_entries += n
        if n_damped:
            leaves["damping"] = d
    density = None
    if mp.density is not None:
        density, dens_b, n = split(mp.density, 0) if _as_t(mp.density, device).dim() - (
            1 if (B is not None and _as_t(mp.density, device).dim() == 1 and cnv_b and
                  _as_t(mp.density, device).shape[0] == B and B != nb) else 0) == 0 else split(mp.density, 1)
        n_entries += n
    free = torch.as_tensor(spec.free_dofs, device=device)
    if mp.inertia is None:
        if density is None:
            raise ValueError("mechanical_params needs either `inertia` or `density`")
        dens = density
        if B is not None and cnv_b and dens.dim() >= 1 and dens.shape[0] == B:
            dens = dens.reshape(B, *([1] * (1 if dens.dim() == 1 else 0)), *dens.shape[1:])
        # `inertia_full`: the same quantity already evaluated by the caller (libdfx geometry kernel)
        if inertia_full is None:
            inertia_full = compute_inertia(cnv, dens)
    else:
        inertia_full, _, n = split(mp.inertia, 2)
        n_entries += n
    leaves["inertia"] = inertia_full.reshape(*inertia_full.shape[:-2], nb * 3)[..., free]
    n_entries += spec.n_free
    if mp.contact_params is not None:
        n_entries += sum(split(v, 0)[2] for v in mp.contact_params)
    if spec.contact:
        if mp.contact_params is None:
            raise ValueError("the energy includes contact but mechanical_params.contact_params is None")
        cp = mp.contact_params
        leaves["contact"] = torch.stack(torch.broadcast_tensors(
            _as_t(cp.min_angle, device), _as_t(cp.cutoff_angle, device), _as_t(cp.k_contact, device)), dim=-1)
    n_entries += sum(split(v, 0)[2] for v in control_params.loading_params.values())
    n_entries += sum(split(v, 0)[2] for v in control_params.constraint_params.values())
    if spec.n_drive_params:
        missing = [n for n in drive.param_names if n not in control_params.constraint_params]
        if missing:
            raise KeyError(f"constraint_params lacks {missing} required by {type(drive).__name__}")
        leaves["drive"] = torch.stack(torch.broadcast_tensors(
            *[_as_t(control_params.constraint_params[n], device) for n in drive.param_names]), dim=-1)
    if control_params.magnetic_params is not None:
        n_entries += sum(_as_t(v, device).numel() for v in control_params.magnetic_params if v is not None)
    leaves = {k: v.contiguous() for k, v in leaves.items()}
    aug_size = 4 * spec.n_free + 1 + n_entries
    return leaves, tuple(pb), damping_per_dof, aug_size


class _OdeSolve(torch.autograd.Function):
    """forward = dfx_forward, backward = dfx_adjoint (the reference's odeint custom_vjp)."""

    @staticmethod
    def forward(ctx, solver, meta, y0, ts, *leaf_tensors):
        names = meta["names"]
        leaves = dict(zip(names, leaf_tensors))
        ps = _abi.ParamSet(solver.spec, meta["batch"], leaves, meta["per_bond"], meta["damping_per_dof"])
        ys, stats = solver.lib_forward(ps, y0, ts)
        ctx.solver, ctx.meta, ctx.ps = solver, meta, ps
        ctx.save_for_backward(ys, ts, y0)
        solver.last_forward_stats = stats
        return ys

    @staticmethod
    def backward(ctx, g):
        solver, meta, ps = ctx.solver, ctx.meta, ctx.ps
        ys, ts, y0 = ctx.saved_tensors
        y0_bar, ts_bar, grads, stats = solver.lib_adjoint(ps, ys, ts, g.contiguous(), meta["aug_size"])
        solver.last_adjoint_stats = stats
        out = []
        for n in meta["names"]:
            gl = grads[n]
            out.append(gl if ps.batched[n] else gl.sum(0))
        if y0.dim() == 1:
            y0_bar = y0_bar.sum(0)
        if ts.dim() == 1:
            ts_bar = ts_bar.sum(0)
        return (None, None, y0_bar, ts_bar, *out)


class _DeviceObjective(torch.autograd.Function):
    """J[b] evaluated on the device (dfx_forward + dfx_objective): the kinetic energy of the target DOFs or their
    angular momentum about a spin centre, summed over the output times; backward = dfx_adjoint_objective, whose
    kernel forms the cotangent dJ/dys itself, plus the explicit dJ/d(inertia) and dJ/d(arm).  The trajectory never
    leaves libdfx buffers."""

    @staticmethod
    def forward(ctx, solver, meta, target_ids, arm, y0, ts, *leaf_tensors):
        names = meta["names"]
        leaves = dict(zip(names, leaf_tensors))
        ps = _abi.ParamSet(solver.spec, meta["batch"], leaves, meta["per_bond"], meta["damping_per_dof"])
        ys, stats = solver.lib_forward(ps, y0, ts)
        solver.last_forward_stats = stats
        arm_c = None if arm is None else arm.contiguous()
        value, ibar, abar = solver._lib.objective_value(solver.handle, ps, ys, target_ids, meta["kind"], arm_c)
        ctx.solver, ctx.meta, ctx.ps, ctx.target_ids, ctx.arm = solver, meta, ps, target_ids, arm_c
        ctx.arm_batched = arm is not None and arm.dim() == 3
        ctx.save_for_backward(ys, ts, y0, ibar, abar if abar is not None else ibar)
        return value

    @staticmethod
    def backward(ctx, gJ):
        solver, meta, ps = ctx.solver, ctx.meta, ctx.ps
        ys, ts, y0, ibar, abar = ctx.saved_tensors
        opt, order = solver.adjoint_launch_options(ps.batch)  # noqa: F841 (order kept alive until the launch is queued)
        y0_bar, ts_bar, grads, stats = solver._lib.adjoint_objective(
            solver.handle, ps, ys, ts, ctx.target_ids, gJ, solver.rtol, solver.atol, meta["aug_size"], opt,
            meta["kind"], ctx.arm)
        solver.last_adjoint_stats = stats
        grads["inertia"] = grads["inertia"] + gJ[:, None] * ibar  # explicit dependence of J on the masses
        arm_bar = None
        if ctx.arm is not None:
            arm_bar = gJ[:, None, None] * abar
            if not ctx.arm_batched:
                arm_bar = arm_bar.sum(0)
        out = []
        for n in meta["names"]:
            gl = grads[n]
            out.append(gl if ps.batched[n] else gl.sum(0))
        if y0.dim() == 1:
            y0_bar = y0_bar.sum(0)
        if ts.dim() == 1:
            ts_bar = ts_bar.sum(0)
        return (None, None, None, arm_bar, y0_bar, ts_bar, *out)


class DynamicSolver:
    """Device-side solver for one topology: owns the libdfx handle, launches forward / adjoint."""

    def __init__(self, spec: _abi.TopologySpec, drive: DriveSignal, rtol=1e-8, atol=1e-8, device=None,
                 init_step_variant=0, max_steps=0, threads=0):
        from . import _lib  # raises if the CUDA library is missing: there is no fallback
        self._lib = _lib
        self.spec, self.drive = spec, drive
        self.rtol, self.atol = float(rtol), float(atol)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("difflexmm_b200 runs on CUDA devices only (no CPU path)")
        self.options = _abi.DfxOptions(int(init_step_variant), int(threads), int(max_steps))
        self.handle = _lib.Topology(spec, self.device.index if self.device.index is not None
                                    else torch.cuda.current_device())
        self.last_forward_stats = None
        self.last_adjoint_stats = None

    # -- raw launches (device tensors) -------------------------------------------------------
    def lib_forward(self, ps, y0, ts):
        # a batch with more designs than forward slots (2 CTAs per SM) that was solved before -- the next iteration of an
        # optimisation loop -- is launched longest-first by the step counts of that previous solve (results do not depend
        # on the launch order; it only shortens the tail of the launch)
        opt, order, prev = self.options, None, self.last_forward_stats  # noqa: F841 (order kept alive until the launch is queued)
        if prev is not None and prev.steps_device().numel() == ps.batch:
            opt, order = self._lib.longest_first(prev, self.options, min_batch=297)
        return self._lib.forward(self.handle, ps, y0, ts, self.rtol, self.atol, opt)

    def adjoint_launch_options(self, batch):
        """Launch order of the adjoint of `batch` designs: longest first, so that a launch with more designs than SMs does
        not end with a few long designs on an otherwise idle GPU (one CTA per SM; attempted steps vary by a factor of two
        over an ensemble).  Predictor of a design's length: the adjoint step count of the PREVIOUS evaluation of the same
        batch when there is one -- the next iteration of an optimisation loop, where designs move little -- else the step
        count of the forward solve just done (a weak predictor: correlation 0.15 on the cfg3 ensemble, the tail then costs
        10 % of the launch, `tools/tail_analysis.py`).  The results do not depend on the order.  -> (options, order tensor)"""
        for prev in (self.last_adjoint_stats, self.last_forward_stats):
            if prev is not None and prev.steps_device().numel() == batch:
                return self._lib.longest_first(prev, self.options)
        return self.options, None

    def lib_adjoint(self, ps, ys, ts, g, aug_size):
        opt, order = self.adjoint_launch_options(ps.batch)  # noqa: F841 (order kept alive until the launch is queued)
        return self._lib.adjoint(self.handle, ps, ys, ts, g, self.rtol, self.atol, aug_size, opt)

    # -- differentiable odeint ------------------------------------------------------------
    def odeint(self, y0, ts, leaves, batch, per_bond=(), damping_per_dof=False, aug_size=0):
        """ys[B, n_t, 2*n_free] for `batch` designs; differentiable w.r.t. y0, ts and the leaves."""
        names = [n for n in _abi.LEAF_NAMES if n in leaves]
        meta = dict(names=names, batch=batch, per_bond=tuple(per_bond), damping_per_dof=damping_per_dof,
                    aug_size=int(aug_size))
        return _OdeSolve.apply(self, meta, y0.contiguous(), ts.contiguous(), *[leaves[n].contiguous() for n in names])

    def odeint_objective(self, y0, ts, leaves, target_free_ids, batch, per_bond=(), damping_per_dof=False, aug_size=0,
                         kind=_abi.DFX_OBJ_KINETIC, arm=None):
        """J[B] of the solution of `odeint` (forward + objective + adjoint inside libdfx, see `_DeviceObjective`);
        differentiable.  kind DFX_OBJ_KINETIC: sum_t sum_{f in target_free_ids} 1/2 m_f v_f(t)^2;
        DFX_OBJ_ANGULAR: angular momentum of the target blocks with `arm` = reference centroid - spin centre."""
        names = [n for n in _abi.LEAF_NAMES if n in leaves]
        meta = dict(names=names, batch=batch, per_bond=tuple(per_bond), damping_per_dof=damping_per_dof,
                    aug_size=int(aug_size), kind=int(kind))
        return _DeviceObjective.apply(self, meta, target_free_ids, arm, y0.contiguous(), ts.contiguous(),
                                      *[leaves[n].contiguous() for n in names])

    def odeint_kinetic(self, y0, ts, leaves, target_free_ids, batch, per_bond=(), damping_per_dof=False, aug_size=0):
        return self.odeint_objective(y0, ts, leaves, target_free_ids, batch, per_bond, damping_per_dof, aug_size)

    # -- reference-level solve ---------------------------------------------------------------
    def solve(self, state0, timepoints, control_params: ControlParams, batch: Optional[int] = None, per_bond=()):
        spec, dev = self.spec, self.device
        B = batch
        leaves, pb, dpd, aug_size = lower_params(spec, self.drive, control_params, B, dev, per_bond)
        free = torch.as_tensor(spec.free_dofs, device=dev)
        state0 = _as_t(state0, dev)
        ts = _as_t(timepoints, dev)
        y0 = state0.reshape(*state0.shape[:-3], 2, spec.n_blocks * 3)[..., free]  # dynamics.py:156
        y0 = y0.reshape(*y0.shape[:-2], 2 * spec.n_free)
        ys = self.odeint(y0, ts, leaves, 1 if B is None else B, pb, dpd, aug_size)
        fields = self.expand_fields(ys, ts, control_params)
        return fields[0] if B is None else fields

    def target_free_ids(self, target_blocks):
        """positions, in the free-DOF vector, of the three DOFs of every target block (they must be free)."""
        dofs = (np.asarray(target_blocks, dtype=np.int64)[:, None] * 3 + np.arange(3)[None]).reshape(-1)
        pos = np.searchsorted(self.spec.free_dofs, dofs)
        if np.any(pos >= self.spec.n_free) or np.any(np.asarray(self.spec.free_dofs)[np.minimum(pos, self.spec.n_free - 1)] != dofs):
            raise ValueError("target blocks of the on-device kinetic objective must not have constrained DOFs")
        return torch.as_tensor(pos.astype(np.int32), device=self.device)

    def angular_momentum_objective(self, state0, timepoints, control_params: ControlParams, target_blocks, spin_center,
                                   batch: Optional[int] = None, per_bond=(), inertia_full=None):
        """Fused objective of the reference's spin problem (`problems/quads_spin.py:395-428`, `energy.py:502-519`):
        angular momentum of `target_blocks` about `spin_center` ((2,) or (B, 2)), summed over blocks and output times."""
        return self._fused_objective(_abi.DFX_OBJ_ANGULAR, state0, timepoints, control_params, target_blocks, batch,
                                     per_bond, inertia_full, spin_center)

    def kinetic_objective(self, state0, timepoints, control_params: ControlParams, target_blocks,
                          batch: Optional[int] = None, per_bond=(), inertia_full=None):
        """Fused objective of the reference's focusing problems (`problems/quads_focusing.py:453-467`):
        sum over output times of the kinetic energy of `target_blocks`.  -> (B,) tensor (scalar without batch),
        differentiable w.r.t. control_params; forward, objective and adjoint all run inside libdfx."""
        return self._fused_objective(_abi.DFX_OBJ_KINETIC, state0, timepoints, control_params, target_blocks, batch,
                                     per_bond, inertia_full, None)

    def _fused_objective(self, kind, state0, timepoints, control_params, target_blocks, batch, per_bond, inertia_full,
                         spin_center):
        spec, dev = self.spec, self.device
        leaves, pb, dpd, aug_size = lower_params(spec, self.drive, control_params, batch, dev, per_bond, inertia_full)
        free = torch.as_tensor(spec.free_dofs, device=dev)
        state0 = _as_t(state0, dev)
        y0 = state0.reshape(*state0.shape[:-3], 2, spec.n_blocks * 3)[..., free]
        y0 = y0.reshape(*y0.shape[:-2], 2 * spec.n_free)
        arm = None
        if kind == _abi.DFX_OBJ_ANGULAR:
            # arm = reference centroid of the target blocks - spin centre (differentiable w.r.t. both)
            tb = torch.as_tensor(np.asarray(target_blocks, dtype=np.int64), device=dev)
            cen = _as_t(control_params.geometrical_params.block_centroids, dev)
            center = _as_t(spin_center, dev)
            arm = cen.index_select(-2, tb) - (center[..., None, :] if center.dim() == 2 else center)
            if batch is not None and arm.dim() == 2:
                arm = arm[None].expand(batch, -1, -1)
        J = self.odeint_objective(y0, _as_t(timepoints, dev), leaves, self.target_free_ids(target_blocks),
                                  1 if batch is None else batch, pb, dpd, aug_size, kind, arm)
        return J[0] if batch is None else J

    def expand_fields(self, ys, ts, control_params):
        """(B, n_t, 2, n_blocks, 3) from the free-DOF solution (reference `dynamics.py:129-136,
        169-182`, without the dense Jacobian): free DOFs are copied; constrained DOFs follow the
        drive signal (displacement) and its time derivative (velocity).  Differentiable torch: it is
        cheap post-processing outside the time loop (libdfx offers the same as `dfx_expand_fields`)."""
        spec, dev = self.spec, self.device
        B, n_t = ys.shape[0], ys.shape[1]
        nf, nd = spec.n_free, spec.n_blocks * 3
        free = torch.as_tensor(spec.free_dofs, device=dev)
        out = torch.zeros((B, n_t, 2, nd), dtype=_F64, device=dev)
        out = out.index_copy(3, free, ys.reshape(B, n_t, 2, nf))
        if len(spec.constrained_dofs) and spec.drive_kind != _abi.DFX_DRIVE_ZERO:
            cons = torch.as_tensor(spec.constrained_dofs.astype(np.int64), device=dev)
            u_c, v_c = drive_values(self.drive, ts if ts.dim() == 2 else ts[None].expand(B, n_t),
                                    control_params.constraint_params, dev)
            out = out.index_copy(3, cons, torch.stack([u_c.expand(B, n_t, -1), v_c.expand(B, n_t, -1)], dim=2))
        return out.reshape(B, n_t, 2, spec.n_blocks, 3)


def drive_values(drive: DriveSignal, tb, constraint_params, device):
    """displacement and velocity of every constrained DOF at times `tb` (B, n_t): (u_c, du_c/dt), each
    (B, n_t, n_constrained).  The time derivative is taken per channel with autograd on the `where`-selected
    branch, which is what `jacobian(kinematics, argnums=1)` yields in the reference (`dynamics.py:130-134`)."""
    params = {n: _as_t(constraint_params[n], device) for n in drive.param_names}
    params = {n: (p.reshape(-1, 1) if p.dim() == 1 else p) for n, p in params.items()}
    need_graph = any(p.requires_grad for p in params.values())
    with torch.enable_grad():
        t = tb.detach().clone().requires_grad_(True)
        s0, s1 = drive.channels(t, **params)
        rates = []
        for sk in (s0, s1):
            if sk.requires_grad:
                rates.append(torch.autograd.grad(sk.sum(), t, create_graph=need_graph, allow_unused=True)[0])
            else:
                rates.append(None)
    s0, s1 = drive.channels(tb, **params)
    n_c = len(drive.vec0) if drive.vec0 is not None else len(drive.vec1)
    u = torch.zeros((*tb.shape, n_c), dtype=_F64, device=device)
    v = torch.zeros((*tb.shape, n_c), dtype=_F64, device=device)
    for vec, sk, rk in ((drive.vec0, s0, rates[0]), (drive.vec1, s1, rates[1])):
        if vec is None:
            continue
        vt = torch.as_tensor(vec, device=device)
        u = u + sk[..., None] * vt
        if rk is not None:
            v = v + rk[..., None] * vt
    return u, v


def setup_dynamic_solver(
        geometry: Geometry,
        energy_fn,
        loaded_block_DOF_pairs=None,
        loading_fn=None,
        constrained_block_DOF_pairs=(),
        constrained_DOFs_fn=None,
        damped_blocks=None,
        rtol: float = 1e-8,
        atol: float = 1e-8,
        *, device=None, init_step_variant: int = 0, max_steps: int = 0, threads: int = 0):
    """Set up the dynamic solver (reference `dynamics.py:60-186`, same positional arguments).

    Returns `solve_dynamics(state0, timepoints, control_params) -> (n_t, 2, n_blocks, 3)`.
    `solve_dynamics.batch(...)` and `solve_dynamics.solver` expose the batched device solver.
    """
    spec, drive = lower_topology(geometry, energy_fn, loaded_block_DOF_pairs, loading_fn,
                                 constrained_block_DOF_pairs, constrained_DOFs_fn, damped_blocks)
    solver = DynamicSolver(spec, drive, rtol, atol, device, init_step_variant, max_steps, threads)

    def solve_dynamics(state0, timepoints, control_params: ControlParams):
        return solver.solve(state0, timepoints, control_params)

    def batch(state0, timepoints, control_params: ControlParams, batch: int, per_bond=()):
        return solver.solve(state0, timepoints, control_params, batch=batch, per_bond=per_bond)

    solve_dynamics.batch = batch
    solve_dynamics.solver = solver
    return solve_dynamics
