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


class StretchingTorsionalSpringParams(NamedTuple):
    """reference `utils.py:80-91`: zero-length springs (`energy.stretching_torsional_spring_energy`)."""

    k_stretch: Any
    k_rot: Any


BondParams = Union[LigamentParams, StretchingTorsionalSpringParams]


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


# ------------------------------------------------------------------------------------------------
# save_data / load_data (reference `utils.py:166-201`), pickle-compatible with the reference
# ------------------------------------------------------------------------------------------------
_REFERENCE_MODULE = "difflexmm.utils"
_TUPLE_TYPES = {"SolutionData": SolutionData, "GeometricalParams": GeometricalParams, "LigamentParams": LigamentParams,
                "StretchingTorsionalSpringParams": StretchingTorsionalSpringParams,
                "ContactParams": ContactParams, "MechanicalParams": MechanicalParams, "ControlParams": ControlParams}


def _to_numpy(obj):
    """torch tensors -> numpy, recursively through (named) tuples, lists and dicts."""
    import numpy as np
    try:
        import torch
        if isinstance(obj, torch.Tensor):
            return obj.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):
        return type(obj)(*(_to_numpy(v) for v in obj))
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_numpy(v) for v in obj)
    if isinstance(obj, dict):
        return {k: _to_numpy(v) for k, v in obj.items()}
    if isinstance(obj, np.generic):
        return obj.item()
    return obj


def save_data(path_or_filename, data: object, reference_compatible: bool = True):
    """Saves data via `pickle` (reference `utils.py:166-181`).  Tensors are stored as numpy arrays.  With
    `reference_compatible` the parameter / solution NamedTuples are pickled under the reference's class path
    (`difflexmm.utils.SolutionData`, ...), so the file loads with the reference's own `load_data` and plotting."""
    import io
    import pickle
    from pathlib import Path

    path = Path(path_or_filename)
    path.parent.mkdir(parents=True, exist_ok=True)
    data = _to_numpy(data)
    if not reference_compatible:
        with open(path, "wb") as file:
            pickle.dump(data, file)
        return path

    import sys
    import types

    refs = {name: _RefClass(name) for name in _TUPLE_TYPES}

    class _Pickler(pickle.Pickler):
        def reducer_override(self, obj):
            cls = type(obj)
            if cls in _TUPLE_TYPES.values():
                return refs[cls.__name__], tuple(obj)
            return NotImplemented

    # pickle checks that the global `difflexmm.utils.<name>` resolves to the object being pickled: provide stub
    # modules for the duration of the dump (the reference package itself is not a dependency of this one)
    saved = {k: sys.modules.get(k) for k in ("difflexmm", _REFERENCE_MODULE)}
    pkg, mod = types.ModuleType("difflexmm"), types.ModuleType(_REFERENCE_MODULE)
    pkg.utils = mod
    for name, ref in refs.items():
        setattr(mod, name, ref)
    sys.modules["difflexmm"], sys.modules[_REFERENCE_MODULE] = pkg, mod
    try:
        buf = io.BytesIO()
        _Pickler(buf, protocol=4).dump(data)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    with open(path, "wb") as file:
        file.write(buf.getvalue())
    return path


class _RefClass:
    """Callable stand-in that pickles as the global `difflexmm.utils.<name>`."""

    def __init__(self, name):
        self.name = name

    def __reduce__(self):
        return self.name  # a string means "pickle as the global <__module__>.<name>"

    __module__ = _REFERENCE_MODULE

    def __call__(self, *args):
        return _TUPLE_TYPES[self.name](*args)


def _reconstruct_jax_array(fun, args, arr_state, aval_state=None):
    """numpy stand-in for `jax._src.array._reconstruct_array` (files written by the reference hold jax arrays)."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


def load_data(path_or_filename):
    """Loads a pickle written by `save_data` here or by the reference (reference `utils.py:184-201`): classes of
    `difflexmm.utils` map to the NamedTuples of this module and pickled jax arrays come back as numpy arrays."""
    import pickle

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module in (_REFERENCE_MODULE, __name__) and name in _TUPLE_TYPES:
                return _TUPLE_TYPES[name]
            if module.startswith("jax") and name == "_reconstruct_array":
                return _reconstruct_jax_array
            return super().find_class(module, name)

    with open(path_or_filename, "rb") as file:
        return _Unpickler(file).load()
