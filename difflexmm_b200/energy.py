"""Energy vocabulary of the solver front-end.

In the reference `energy_fn` is an opaque JAX closure that `jax.grad` differentiates every RHS
call (`difflexmm/energy.py:410-491`, `dynamics.py:31`).  A CUDA kernel cannot run a Python
closure, so the builders below keep the reference's names and signatures but return small
*descriptors*; `setup_dynamic_solver` lowers them to the static topology of libdfx, whose kernels
contain the analytic gradient / Hessian-vector products of exactly these energies.  Anything that
is not one of these descriptors is rejected loudly (there is no fallback path).
"""

import numpy as np
import torch

from . import _abi


class _BondEnergy:
    def __init__(self, name, kind):
        self.__name__, self.kind = name, kind

    def __repr__(self):
        return f"<bond energy {self.__name__}>"


#: nonlinear ligament (reference `energy.py:158-176`)
ligament_energy = _BondEnergy("ligament_energy", _abi.DFX_BOND_LIGAMENT)
#: linearised ligament (reference `energy.py:99-117`)
ligament_energy_linearized = _BondEnergy("ligament_energy_linearized", _abi.DFX_BOND_LINEARIZED)
#: zero-length stretching + torsional spring between coincident nodes (reference `energy.py:49-66`); bond_params must
#: be `utils.StretchingTorsionalSpringParams`.  Runs on the generic kernels.
stretching_torsional_spring_energy = _BondEnergy("stretching_torsional_spring_energy", _abi.DFX_BOND_SPRING)


class BlockEnergy:
    """Base of the energy descriptors accepted by `setup_dynamic_solver`."""

    bond_connectivity = None
    bond_kind = None
    contact = False


class StrainEnergy(BlockEnergy):
    def __init__(self, bond_connectivity, bond_energy_fn):
        if not isinstance(bond_energy_fn, _BondEnergy):
            raise TypeError(
                "bond_energy_fn must be difflexmm_b200.energy.ligament_energy, ligament_energy_linearized or "
                "stretching_torsional_spring_energy; arbitrary Python energies cannot run inside the CUDA solver")
        self.bond_connectivity = np.asarray(bond_connectivity, dtype=np.int64).reshape(-1, 2)
        self.bond_kind = bond_energy_fn.kind


class ContactEnergy(BlockEnergy):
    """angle-based (reference `energy.py:204-219`) or distance-based between the two void edges of every bond
    (`energy.py:222-330`; `contact_params.min_angle` / `cutoff_angle` are lengths then, and the block centroids become a
    differentiable leaf; runs on the generic kernels)"""

    def __init__(self, bond_connectivity, angle_based=True):
        self.bond_connectivity = np.asarray(bond_connectivity, dtype=np.int64).reshape(-1, 2)
        self.contact = _abi.DFX_CONTACT_ANGLE if angle_based else _abi.DFX_CONTACT_DISTANCE


class CombinedEnergy(BlockEnergy):
    def __init__(self, parts):
        strain = [p for p in parts if isinstance(p, StrainEnergy)]
        contact = [p for p in parts if isinstance(p, ContactEnergy)]
        if len(strain) != 1 or len(contact) > 1 or len(strain) + len(contact) != len(parts):
            raise TypeError("combine_block_energies expects one strain energy and at most one contact energy")
        self.bond_connectivity, self.bond_kind = strain[0].bond_connectivity, strain[0].bond_kind
        if contact:
            if not np.array_equal(contact[0].bond_connectivity, self.bond_connectivity):
                raise ValueError("strain and contact energies must share the bond connectivity")
            self.contact = contact[0].contact


def build_strain_energy(bond_connectivity, bond_energy_fn=ligament_energy_linearized):
    """reference `energy.py:410-449` (same default bond energy)."""
    return StrainEnergy(bond_connectivity, bond_energy_fn)


def build_contact_energy(bond_connectivity, angle_based=True):
    """reference `energy.py:364-407`"""
    return ContactEnergy(bond_connectivity, angle_based)


def combine_block_energies(*energy_fns):
    """reference `energy.py:452-470`"""
    return CombinedEnergy(list(energy_fns))


def kinetic_energy(block_velocity, inertia):
    """sum(inertia * v^2 / 2) (reference `energy.py:494-499`)."""
    return torch.sum(inertia * block_velocity ** 2 / 2)


def angular_momentum(block_position, block_velocity, inertia, reference_point=(0.0, 0.0)):
    """(position - reference) x (m v) + I omega per block, shape (..., n_blocks) (reference `energy.py:502-519`)."""
    ref = torch.as_tensor(reference_point, dtype=block_position.dtype, device=block_position.device)
    r = block_position[..., :2] - ref
    p = block_velocity[..., :2] * inertia[..., :2]
    return r[..., 0] * p[..., 1] - r[..., 1] * p[..., 0] + block_velocity[..., 2] * inertia[..., 2]
