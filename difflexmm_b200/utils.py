"""Parameter containers of the dynamic solver (the API surface the reference exposes).

Mirrors the NamedTuple pytrees of the reference (`difflexmm/utils.py:48-163`): same
class names, same field names, same defaults, so code written against the reference
builds a `ControlParams` for this solver unchanged.  Leaves are torch float64 tensors
(or Python floats); `None` leaves are structural, exactly as in the reference.
"""

from typing import Any, Dict, NamedTuple, Optional, Union


class SolutionData(NamedTuple):
    """Solution fields lumped with the geometry (reference `utils.py:9-25`)."""

    block_centroids: Any
    centroid_node_vectors: Any
    bond_connectivity: Any
    timepoints: Any
    fields: Any


class GeometricalParams(NamedTuple):
    """reference `utils.py:48-59`"""

    block_centroids: Any  # (n_blocks, 2)
    centroid_node_vectors: Any  # (n_blocks, n_nodes_per_block, 2)


class LigamentParams(NamedTuple):
    """reference `utils.py:62-77`; each stiffness is a scalar or an (n_bonds,) array."""

    k_stretch: Any
    k_shear: Any
    k_rot: Any
    reference_vector: Any  # (n_bonds, 2)


BondParams = Union[LigamentParams]


class ContactParams(NamedTuple):
    """reference `utils.py:97-111`"""

    min_angle: Any
    cutoff_angle: Any
    k_contact: Any


class MechanicalParams(NamedTuple):
    """reference `utils.py:128-142`"""

    bond_params: BondParams
    density: Any
    inertia: Optional[Any] = None
    damping: Any = 0.
    contact_params: Optional[ContactParams] = None


class ControlParams(NamedTuple):
    """reference `utils.py:145-163`"""

    geometrical_params: GeometricalParams
    mechanical_params: MechanicalParams
    magnetic_params: Optional[Any] = None
    loading_params: Dict = dict()
    constraint_params: Dict = dict()
