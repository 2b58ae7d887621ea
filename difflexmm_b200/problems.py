"""Forward / objective problems on top of the solver -- the callers of the hot path.

Mirrors the boundary conditions, drive signals, control parameters and objectives of the
reference problem files (they define the configurations the solver is benchmarked on):
  QuadsFocusing   <- problems/quads_focusing.py:25-317, 408-471        (cfg1, cfg3, cfg5)
  KagomeFocusing  <- problems/kagome_focusing.py:25-270, 380-420       (cfg2)
  QuadsStaticTuning <- problems/quads_kinetic_energy_static_tuning.py:25-330, 431-484  (cfg4)
Default parameter sets are the ones of the reference notebooks.  Every
problem can be lowered on the host (`lower()`: topology + drive, no GPU needed) and set up on
a device (`setup()`: builds the CUDA solver).
"""

import math
from dataclasses import dataclass, field
from typing import Any, Optional, Tuple

import numpy as np
import torch

from .dynamics import DynamicSolver, lower_params, lower_topology
from .energy import (angular_momentum, build_contact_energy, build_strain_energy, combine_block_energies, kinetic_energy,
                     ligament_energy, ligament_energy_linearized)
from .geometry import KagomeGeometry, QuadGeometry, compute_inertia
from .loading import pulse_drive, static_pulse_drive
from .utils import (ContactParams, ControlParams, GeometricalParams, LigamentParams, MechanicalParams, SolutionData)

_F64 = torch.float64


def _tile3(blocks, first=(0, 1, 2)):
    """[[b, d] for d in first for b in blocks] -- the reference's tile / repeat pattern."""
    blocks = np.asarray(blocks, dtype=np.int64)
    return np.stack([np.tile(blocks, 3), np.repeat(np.asarray(first), len(blocks))], -1)


class _ProblemBase:
    geometry = None
    _solver: Optional[DynamicSolver] = None

    def energy(self, bonds):
        strain = build_strain_energy(bonds, ligament_energy_linearized if self.linearized_strains else ligament_energy)
        return combine_block_energies(strain, build_contact_energy(bonds)) if self.use_contact else strain

    def lower(self):
        """-> (TopologySpec, DriveSignal): everything static, on the host."""
        geo = self.make_geometry()
        geo.compute_geometry()
        self.geometry = geo  # published only when complete: lower() may be called again from another thread
        pairs, drive = self.constraints(geo)
        self.constrained_block_DOF_pairs = pairs
        spec, drive = lower_topology(geo, self.energy(geo.bond_connectivity()), None, None, pairs, drive,
                                     np.arange(geo.n_blocks))
        self.spec, self.drive = spec, drive
        return spec, drive

    def setup(self, device=None, **solver_kwargs):
        spec, drive = self.lower()
        self._solver = DynamicSolver(spec, drive, self.rtol, self.atol, device, **solver_kwargs)
        return self._solver

    @property
    def solver(self) -> DynamicSolver:
        if self._solver is None:
            self.setup()
        return self._solver

    def _torch_geometry(self, design, device):
        """design arrays (optionally with a leading batch axis) -> (centroid_node_vectors, block_centroids) with the
        differentiable torch maps of `geometry.py` (one design at a time)."""
        parts = [torch.as_tensor(d, dtype=_F64, device=device) for d in design]
        geo = self.geometry
        if parts[0].dim() == 4:  # batch of designs
            cnv = torch.stack([geo.centroid_node_vectors(*d) for d in zip(*parts)])
            cen = torch.stack([geo.block_centroids(*d) for d in zip(*parts)])
            return cnv, cen
        return geo.centroid_node_vectors(*parts), geo.block_centroids(*parts)

    def device_geometry(self):
        """libdfx design map of this problem's lattice (`geometry_device.DeviceGeometry`), created on first use."""
        if getattr(self, "_device_geometry", None) is None:
            from .geometry_device import DeviceGeometry
            self._device_geometry = DeviceGeometry(self.geometry, self.solver.device)
        return self._device_geometry

    def state0(self, device=None):
        return torch.zeros((2, self.geometry.n_blocks, 3), dtype=_F64, device=device)

    def boundary_inputs(self, design, batch=None, device="cpu"):
        """Leaves at the solver boundary (what crosses into libdfx) for a design (or a batch of designs):
        (leaves, per_bond, damping_per_dof, aug_size, y0, ts)."""
        if self.geometry is None:
            self.lower()
        cp = self.control_params(design, device)
        leaves, pb, dpd, aug = lower_params(self.spec, self.drive, cp, batch, device)
        y0 = torch.zeros(2 * self.spec.n_free, dtype=_F64, device=device)
        return leaves, pb, dpd, aug, y0, self.timepoints(device)

    def solve(self, design, batch=None) -> SolutionData:
        s = self.solver
        cp = self.control_params(design, s.device)
        fields = s.solve(self.state0(s.device), self.timepoints(s.device), cp, batch=batch)
        return SolutionData(cp.geometrical_params.block_centroids, cp.geometrical_params.centroid_node_vectors,
                            self.geometry.bond_connectivity(), self.timepoints(s.device), fields)

    def target_kinetic_energy(self, design, batch=None, fused=False):
        """objective of the reference OptimizationProblem (`quads_focusing.py:453-467`): kinetic energy of the
        target blocks summed over the output times.  `fused=True` evaluates it inside libdfx
        (`DynamicSolver.kinetic_objective`): no fields, no cotangent tensor."""
        if fused:
            # design -> parameters, forward solve, objective and both backward passes inside libdfx
            s = self.solver
            cnv, cen, inertia = self.device_geometry()(design, self.density)
            cp = self.control_params(design, s.device, geom=(cnv, cen))
            return s.kinetic_objective(self.state0(s.device), self.timepoints(s.device), cp, self.target_blocks(),
                                       batch=batch, inertia_full=inertia)
        sol = self.solve(design, batch)
        tb = torch.as_tensor(self.target_blocks(), device=sol.fields.device)
        inertia = compute_inertia(sol.centroid_node_vectors, torch.as_tensor(self.density, dtype=_F64,
                                                                            device=sol.fields.device))
        v = sol.fields[..., 1, :, :].index_select(-2, tb)
        m = inertia.index_select(-2, tb)
        if batch is None:
            return kinetic_energy(v, m)
        return (m[:, None] * v ** 2 / 2).sum(dim=(1, 2, 3))

    def target_angular_momentum(self, design, spin_center="center", batch=None, fused=False):
        """objective of the reference's spin problem (`problems/quads_spin.py:395-428`): angular momentum of the target
        blocks about `spin_center`, summed over blocks and output times.  "center" = mean reference centroid of the
        target blocks for THIS design, held fixed (not differentiated), as in the reference where it is evaluated
        once from the forward input.  `fused=True` (and batches) evaluate it inside libdfx
        (`DynamicSolver.angular_momentum_objective`)."""
        if fused or batch is not None:
            s = self.solver
            cnv, cen, inertia = self.device_geometry()(design, self.density)
            tb = torch.as_tensor(self.target_blocks(), device=s.device)
            center = cen.index_select(-2, tb).mean(-2).detach() if isinstance(spin_center, str) else spin_center
            cp = self.control_params(design, s.device, geom=(cnv, cen))
            return s.angular_momentum_objective(self.state0(s.device), self.timepoints(s.device), cp, self.target_blocks(),
                                                center, batch=batch, inertia_full=inertia)
        sol = self.solve(design)
        dev = sol.fields.device
        tb = torch.as_tensor(self.target_blocks(), device=dev)
        cen = sol.block_centroids.index_select(0, tb)
        center = cen.mean(0).detach() if isinstance(spin_center, str) else torch.as_tensor(spin_center, dtype=_F64, device=dev)
        inertia = compute_inertia(sol.centroid_node_vectors.index_select(0, tb),
                                  torch.as_tensor(self.density, dtype=_F64, device=dev))
        u = sol.fields[:, 0].index_select(1, tb)[..., :2]
        v = sol.fields[:, 1].index_select(1, tb)
        return angular_momentum(cen[None] + u, v, inertia[None], center).sum()


@dataclass
class QuadsFocusing(_ProblemBase):
    n1_blocks: int = 24
    n2_blocks: int = 16
    spacing: float = 15.
    bond_length: float = 2.25
    k_stretch: Any = 120.
    k_shear: Any = 1.19
    k_rot: Any = 1.5
    density: Any = 6.18e-9
    damping: Any = None
    amplitude: Any = 7.5
    loading_rate: Any = 30.
    input_delay: Any = 0.1 / 30.
    n_excited_blocks: int = 2
    loaded_side: str = "left"
    input_shift: int = 0
    simulation_time: Any = 2. / 30.
    n_timepoints: int = 200
    linearized_strains: bool = False
    use_contact: bool = True
    k_contact: Any = 1.5
    min_angle: Any = -15. * math.pi / 180
    cutoff_angle: Any = -10. * math.pi / 180
    n_blocks_clamped_corner: int = 2
    atol: float = 1e-4
    rtol: float = 1e-8
    target_size: Tuple[int, int] = (2, 2)
    target_shift: Tuple[int, int] = (4, 5)
    initial_angle: float = 25. * math.pi / 180

    def __post_init__(self):
        if self.damping is None:  # notebooks/quads_focusing_3dp_pla_shims.ipynb:487-491
            rho, s = self.density, self.spacing
            self.damping = 0.0186 * np.array([2 * (0.36125 * rho * s ** 2 * self.k_shear) ** 0.5] * 2 +
                                             [2 * (0.02175026 * rho * s ** 4 * self.k_rot) ** 0.5]) * np.ones(
                (self.n1_blocks * self.n2_blocks, 3))

    def make_geometry(self):
        return QuadGeometry(self.n1_blocks, self.n2_blocks, self.spacing, self.bond_length)

    def constraints(self, geo):
        n1, n2, ne, sh, nc = geo.n1_blocks, geo.n2_blocks, self.n_excited_blocks, self.input_shift, self.n_blocks_clamped_corner
        if self.loaded_side == "left":
            driven = _tile3(np.arange((n2 - ne) // 2 + sh, (n2 + ne) // 2 + sh) * n1)
        elif self.loaded_side == "right":
            driven = _tile3(np.arange((n2 - ne) // 2 + sh, (n2 + ne) // 2 + sh) * n1 + (n1 - 1))
        elif self.loaded_side == "bottom":
            driven = _tile3(np.arange((n1 - ne) // 2 + sh, (n1 + ne) // 2 + sh), (1, 0, 2))
        elif self.loaded_side == "top":
            driven = _tile3(np.arange((n1 - ne) // 2 + sh, (n1 + ne) // 2 + sh) + n1 * (n2 - 1), (1, 0, 2))
        else:
            raise ValueError(f"Unknown loaded_side: {self.loaded_side}. Should be either 'left', 'right', 'bottom' or 'top'.")
        nb = geo.n_blocks
        bl = np.concatenate([np.arange(0, nc), [i * n1 for i in range(1, nc)]])
        br = np.concatenate([np.arange(n1 - nc, n1), [(i + 1) * n1 - 1 for i in range(1, nc)]])
        tr = np.concatenate([np.arange(nb - nc, nb), [nb - i * n1 - 1 for i in range(1, nc)]])
        tl = np.concatenate([np.arange(nb - n1, nb - n1 + nc), [nb - n1 - i * n1 for i in range(1, nc)]])
        pairs = np.concatenate([driven] + [_tile3(c) for c in (bl, br, tr, tl)]).astype(np.int64)
        vec = np.zeros(len(pairs))
        vec[:ne] = 1
        return pairs, pulse_drive(vec)

    def timepoints(self, device=None):
        return torch.linspace(0, self.simulation_time, self.n_timepoints, dtype=_F64, device=device)

    def initial_design(self):
        return self.make_geometry().get_design_from_rotated_square(self.initial_angle)

    def control_params(self, design, device=None, geom=None):
        geo = self.geometry
        cnv, cen = geom if geom is not None else self._torch_geometry(design, device)
        amplitude = self.amplitude if self.loaded_side in ("left", "bottom") else -self.amplitude
        return ControlParams(
            geometrical_params=GeometricalParams(block_centroids=cen, centroid_node_vectors=cnv),
            mechanical_params=MechanicalParams(
                bond_params=LigamentParams(k_stretch=self.k_stretch, k_shear=self.k_shear, k_rot=self.k_rot,
                                           reference_vector=geo.reference_bond_vectors().to(device)),
                density=self.density, damping=self.damping,
                contact_params=ContactParams(k_contact=self.k_contact, min_angle=self.min_angle,
                                             cutoff_angle=self.cutoff_angle)),
            constraint_params=dict(amplitude=amplitude, loading_rate=self.loading_rate, input_delay=self.input_delay))

    def target_blocks(self):
        n1, n2, ts, sh = self.n1_blocks, self.n2_blocks, self.target_size, self.target_shift
        return np.array([j * n1 + i
                         for i in range((n1 - ts[0]) // 2 + sh[0], (n1 + ts[0]) // 2 + sh[0])
                         for j in range((n2 - ts[1]) // 2 + sh[1], (n2 + ts[1]) // 2 + sh[1])], dtype=np.int64)

    def random_ensemble(self, n_designs, noise=0.15, seed0=0):
        """cfg3: the initial design plus U(-1,1)*noise*spacing on both shift arrays, one numpy PRNG stream per
        design (the reference draws with jax.random, `...random_initial_guess.ipynb:480-484`)."""
        hs0, vs0 = self.initial_design()
        hs, vs = [], []
        for k in range(n_designs):
            rng = np.random.default_rng(seed0 + k)
            hs.append(hs0 + torch.from_numpy(rng.uniform(-1, 1, hs0.shape) * noise * self.spacing))
            vs.append(vs0 + torch.from_numpy(rng.uniform(-1, 1, vs0.shape) * noise * self.spacing))
        return torch.stack(hs), torch.stack(vs)


@dataclass
class KagomeFocusing(_ProblemBase):
    n1_cells: int = 20
    n2_cells: int = 12
    cell_size: float = 20.
    cell_angle: float = math.pi / 3
    bond_length: float = 2.25
    k_stretch: Any = 120.
    k_shear: Any = 1.19
    k_rot: Any = 1.5
    density: Any = 6.18e-9
    damping: Any = None
    amplitude: Any = 10.
    loading_rate: Any = 30.
    input_delay: Any = 0.1 / 30.
    n_excited_blocks: int = 2
    input_shift: int = 0
    simulation_time: Any = 3. / 30.
    n_timepoints: int = 200
    linearized_strains: bool = False
    use_contact: bool = True
    k_contact: Any = 1.5
    min_angle: Any = -15. * math.pi / 180
    cutoff_angle: Any = -10. * math.pi / 180
    n_blocks_clamped_corner: int = 2
    atol: float = 1e-4
    rtol: float = 1e-8
    target_size: Tuple[int, int] = (2, 2)
    target_shift: Tuple[int, int] = (3, 2)

    def __post_init__(self):
        if self.damping is None:  # notebooks/kagome_focusing_3dp_pla_shims.ipynb:243-247
            rho, s = self.density, self.cell_size
            self.damping = 0.0186 * np.array([2 * (0.070175913225 * rho * s ** 2 * self.k_shear) ** 0.5] * 2 +
                                             [2 * (0.0009477510275 * rho * s ** 4 * self.k_rot) ** 0.5]) * np.ones(
                (2 * self.n1_cells * self.n2_cells, 3))

    def make_geometry(self):
        basis = self.cell_size * np.array([[1., 0.], [math.cos(self.cell_angle), math.sin(self.cell_angle)]])
        return KagomeGeometry(self.n1_cells, self.n2_cells, basis, self.bond_length)

    def constraints(self, geo):
        n1, n2, ne, nc, ncell = geo.n1_cells, geo.n2_cells, self.n_excited_blocks, self.n_blocks_clamped_corner, geo.n_cells
        driven = _tile3(np.arange(2 * n1 * ((n2 - ne) // 2), 2 * n1 * ((n2 + ne) // 2), 2 * n1))
        bl = np.concatenate([np.arange(0, nc), [i * n1 for i in range(1, nc)]]) * 2
        br = np.concatenate([np.arange(n1 - nc, n1) * 2, [(i + 1) * 2 * n1 - 1 for i in range(0, nc)]])
        tr = np.concatenate([np.arange(ncell - nc, ncell), [ncell - i * n1 - 1 for i in range(1, nc)]]) * 2 + 1
        tl = np.concatenate([np.arange(ncell - n1, ncell - n1 + nc) * 2 + 1,
                             np.array([ncell - n1 - i * n1 for i in range(0, nc)]) * 2])
        pairs = np.concatenate([driven] + [_tile3(c) for c in (bl, br, tr, tl)]).astype(np.int64)
        vec = np.zeros(len(pairs))
        vec[:ne] = 1
        return pairs, pulse_drive(vec)

    def timepoints(self, device=None):
        return torch.linspace(0, self.simulation_time, self.n_timepoints, dtype=_F64, device=device)

    def initial_design(self):
        n1, n2 = self.n1_cells, self.n2_cells
        return (torch.zeros(n1 + 1, n2, 2, dtype=_F64), torch.zeros(n1, n2 + 1, 2, dtype=_F64),
                torch.zeros(n1, n2, 2, dtype=_F64))

    def control_params(self, design, device=None, geom=None):
        geo = self.geometry
        cnv, cen = geom if geom is not None else self._torch_geometry(design, device)
        return ControlParams(
            geometrical_params=GeometricalParams(block_centroids=cen, centroid_node_vectors=cnv),
            mechanical_params=MechanicalParams(
                bond_params=LigamentParams(k_stretch=self.k_stretch, k_shear=self.k_shear, k_rot=self.k_rot,
                                           reference_vector=geo.reference_bond_vectors().to(device)),
                density=self.density, damping=self.damping,
                contact_params=ContactParams(k_contact=self.k_contact, min_angle=self.min_angle,
                                             cutoff_angle=self.cutoff_angle)),
            constraint_params=dict(amplitude=self.amplitude, loading_rate=self.loading_rate, input_delay=self.input_delay))

    def target_blocks(self):
        n1, n2, ts, sh = self.n1_cells, self.n2_cells, self.target_size, self.target_shift
        return np.array([[2 * (j * n1 + i), 2 * (j * n1 + i) + 1]
                         for i in range((n1 - ts[0]) // 2 + sh[0], (n1 + ts[0]) // 2 + sh[0])
                         for j in range((n2 - ts[1]) // 2 + sh[1], (n2 + ts[1]) // 2 + sh[1])], dtype=np.int64).reshape(-1)


@dataclass
class QuadsStaticTuning(_ProblemBase):
    """cfg4: static pre-compression of the top/bottom rows, then a delayed pulse on the left edge; one task =
    one (amplitude, loading_rate, compressive_strain, compressive_strain_rate) tuple."""
    n1_blocks: int = 24
    n2_blocks: int = 18
    spacing: float = 15.
    bond_length: float = 2.25
    k_stretch: Any = 120.
    k_shear: Any = 1.19
    k_rot: Any = 1.5
    density: Any = 6.18e-9
    damping: Any = None
    n_excited_blocks: int = 2
    input_shift: int = 0
    simulation_time_dynamic: Any = 2. / 30.
    n_timepoints: int = 200
    linearized_strains: bool = False
    use_contact: bool = True
    k_contact: Any = 1.5
    min_angle: Any = -10. * math.pi / 180
    cutoff_angle: Any = -5. * math.pi / 180
    atol: float = 1e-4
    rtol: float = 1e-8
    # one task
    amplitude: Any = 7.5
    loading_rate: Any = 30.
    compressive_strain: Any = 0.01
    compressive_strain_rate: Any = 0.25
    target_size: Tuple[int, int] = (2, 2)
    target_shift: Tuple[int, int] = (2, 2)
    initial_angle: float = 25. * math.pi / 180

    __post_init__ = QuadsFocusing.__post_init__
    make_geometry = QuadsFocusing.make_geometry
    initial_design = QuadsFocusing.initial_design
    target_blocks = QuadsFocusing.target_blocks

    def constraints(self, geo):
        n1, n2, ne, sh = geo.n1_blocks, geo.n2_blocks, self.n_excited_blocks, self.input_shift
        driven = _tile3(np.arange((n2 - ne) // 2 + sh, (n2 + ne) // 2 + sh) * n1)
        bottom = _tile3(np.arange(0, n1), (1, 0, 2))
        top = _tile3(np.arange(geo.n_blocks - n1, geo.n_blocks), (1, 0, 2))
        pairs = np.concatenate([driven, bottom, top]).astype(np.int64)
        dyn = np.zeros(len(pairs))
        dyn[:ne] = 1
        sta = np.zeros(len(pairs))
        sta[3 * ne:3 * ne + n1] = 0.5
        sta[3 * ne + 3 * n1:3 * ne + 4 * n1] = -0.5
        return pairs, static_pulse_drive(dyn, sta * (n2 - 1) * self.spacing)

    def timepoints(self, device=None):
        t_static = self.compressive_strain / self.compressive_strain_rate
        delay = 0.1 / self.loading_rate
        dyn = torch.linspace(t_static + delay, t_static + delay + self.simulation_time_dynamic, self.n_timepoints,
                             dtype=_F64, device=device)
        return torch.cat([torch.zeros(1, dtype=_F64, device=device), dyn])

    def control_params(self, design, device=None, geom=None):
        geo = self.geometry
        cnv, cen = geom if geom is not None else self._torch_geometry(design, device)
        return ControlParams(
            geometrical_params=GeometricalParams(block_centroids=cen, centroid_node_vectors=cnv),
            mechanical_params=MechanicalParams(
                bond_params=LigamentParams(k_stretch=self.k_stretch, k_shear=self.k_shear, k_rot=self.k_rot,
                                           reference_vector=geo.reference_bond_vectors().to(device)),
                density=self.density, damping=self.damping,
                contact_params=ContactParams(k_contact=self.k_contact, min_angle=self.min_angle,
                                             cutoff_angle=self.cutoff_angle)),
            constraint_params=dict(amplitude=self.amplitude, loading_rate=self.loading_rate,
                                   compressive_strain=self.compressive_strain,
                                   compressive_strain_rate=self.compressive_strain_rate,
                                   input_delay=0.1 / self.loading_rate))
