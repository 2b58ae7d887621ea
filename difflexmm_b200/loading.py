"""Drive and load signal vocabulary (reference: opaque closures `constrained_DOFs_fn(t, **params)`
and `loading_fn(state, t, **params)`, `dynamics.py:62-66`).

Every form used by the reference problems is available as a descriptor; the
CUDA kernels evaluate the signal, its time derivative and its parameter derivatives analytically.
Each descriptor is also callable on torch tensors (same formula, differentiable) -- that is used
only for the cheap post-processing of the constrained DOFs (`dynamics.py:129-136`), never inside
the time loop.
"""

import math

import numpy as np
import torch

from . import _abi


def _where_pulse(tau, amplitude, loading_rate, windowed):
    on = (tau > 0.) & ((tau < 1.0 / loading_rate) if windowed else torch.ones_like(tau, dtype=torch.bool))
    shape = (1 - torch.cos(2 * math.pi * loading_rate * tau)) / 2
    return amplitude * torch.where(on, shape, torch.zeros_like(shape))


class DriveSignal:
    """u_c(t) = vec0[c]*s0(t; params) + vec1[c]*s1(t; params) for the constrained DOFs."""

    kind = _abi.DFX_DRIVE_ZERO

    def __init__(self, vec0=None, vec1=None):
        self.vec0 = None if vec0 is None else np.asarray(vec0, dtype=np.float64).reshape(-1)
        self.vec1 = None if vec1 is None else np.asarray(vec1, dtype=np.float64).reshape(-1)

    @property
    def param_names(self):
        return _abi.DRIVE_PARAM_NAMES[self.kind]

    def channels(self, t, **p):  # -> (s0, s1) torch, broadcast over t
        z = torch.zeros_like(t)
        return z, z

    def __call__(self, t, **params):
        s0, s1 = self.channels(t, **params)
        out = 0.
        if self.vec0 is not None:
            out = out + s0[..., None] * torch.as_tensor(self.vec0, device=t.device)
        if self.vec1 is not None:
            out = out + s1[..., None] * torch.as_tensor(self.vec1, device=t.device)
        return out


class zero_drive(DriveSignal):
    """`lambda t: 0` (reference default, `dynamics.py:66`)."""


class pulse_drive(DriveSignal):
    """amplitude*(1-cos(2 pi f (t-delay)))/2 on 0 < t-delay < 1/f, times `loading_vector`
    (reference `problems/quads_focusing.py:211-222`)."""
    kind = _abi.DFX_DRIVE_PULSE

    def __init__(self, loading_vector):
        super().__init__(loading_vector)

    def channels(self, t, amplitude, loading_rate, input_delay):
        return _where_pulse(t - input_delay, amplitude, loading_rate, True), torch.zeros_like(t)


class harmonic_drive(DriveSignal):
    """same as the pulse but on t-delay > 0 (reference `problems/quads_spin.py:210-221`)."""
    kind = _abi.DFX_DRIVE_HARMONIC

    def __init__(self, loading_vector):
        super().__init__(loading_vector)

    def channels(self, t, amplitude, loading_rate, input_delay):
        return _where_pulse(t - input_delay, amplitude, loading_rate, False), torch.zeros_like(t)


class ramp_drive(DriveSignal):
    """amplitude*(t < 1/f ? t f : 1) (reference `problems/hinge_characterization.py:134-139`)."""
    kind = _abi.DFX_DRIVE_RAMP

    def __init__(self, loading_vector):
        super().__init__(loading_vector)

    def channels(self, t, amplitude, loading_rate):
        return amplitude * torch.where(t < 1.0 / loading_rate, t * loading_rate, torch.ones_like(t)), torch.zeros_like(t)


class static_pulse_drive(DriveSignal):
    """static compression ramp + delayed pulse
    (reference `problems/quads_kinetic_energy_static_tuning.py:176-196`).  `static_vector` carries
    the geometric factor `(n2_blocks-1)*spacing` of the reference."""
    kind = _abi.DFX_DRIVE_STATIC_PULSE

    def __init__(self, dynamic_vector, static_vector):
        super().__init__(dynamic_vector, static_vector)

    def channels(self, t, amplitude, loading_rate, compressive_strain, compressive_strain_rate, input_delay):
        t_static = compressive_strain / compressive_strain_rate
        s0 = _where_pulse(t - t_static - input_delay, amplitude, loading_rate, True)
        s1 = torch.where(t < t_static, t * compressive_strain_rate, compressive_strain * torch.ones_like(t))
        return s0, s1


class tabulated_drive(DriveSignal):
    """`excited_blocks_fn(t) * loading_vector` with `excited_blocks_fn = lambda t: jnp.interp(t, times, values)`: the
    measured input signal of the experiment notebooks (reference `problems/quads_focusing.py:223-227`,
    exp/.../experiment_vs_simulation.ipynb cell 12).  No differentiable parameters."""
    kind = _abi.DFX_DRIVE_TABLE

    def __init__(self, times, values, loading_vector):
        super().__init__(loading_vector)
        self.times = np.asarray(times, dtype=np.float64).reshape(-1)
        self.values = np.asarray(values, dtype=np.float64).reshape(-1)

    def channels(self, t):
        xp = torch.as_tensor(self.times, device=t.device)
        fp = torch.as_tensor(self.values, device=t.device)
        i = torch.clamp(torch.searchsorted(xp, t.detach().contiguous(), right=True), 1, len(xp) - 1)
        dx, df = xp[i] - xp[i - 1], fp[i] - fp[i - 1]
        f = torch.where(dx == 0, fp[i], fp[i - 1] + (t - xp[i - 1]) / torch.where(dx == 0, torch.ones_like(dx), dx) * df)
        f = torch.where(t < xp[0], fp[0].expand_as(f), torch.where(t > xp[-1], fp[-1].expand_as(f), f))
        return f, torch.zeros_like(t)


def tabulate_drive(excited_blocks_fn, times, loading_vector):
    """An arbitrary scalar signal `excited_blocks_fn(t)` (a Python closure with captured constants, as the reference's
    problems pass to `build_constrained_kinematics`) as a `tabulated_drive`: the closure is sampled at `times` on the host and
    the kernels interpolate linearly between the samples (`jnp.interp` semantics, constant beyond the ends).  This is an
    approximation whose error is the interpolation error of the chosen grid -- choose it as fine as the signal demands; its
    parameters are not differentiable.  Closures that cannot be tabulated (state dependent) stay out of scope."""
    times = np.asarray(times, dtype=np.float64).reshape(-1)
    values = np.array([float(excited_blocks_fn(float(t))) for t in times], dtype=np.float64)
    return tabulated_drive(times, values, loading_vector)


class LoadSignal:
    """external force on the loaded DOFs, load_vec[l]*s(t) with captured constants."""
    kind = _abi.DFX_LOAD_NONE
    consts = ()

    def __init__(self, load_vector=None):
        self.load_vector = load_vector


class ramp_load(LoadSignal):
    """final_load*(t < 1/rate ? t*rate : 1) (reference `tests/test_difflexmm.py:85-86`)."""
    kind = _abi.DFX_LOAD_RAMP

    def __init__(self, final_load, loading_rate, load_vector=None):
        super().__init__(load_vector)
        self.consts = (float(final_load), float(loading_rate))


class sech2_load(LoadSignal):
    """2A/s^2 * cosh(t/s-3)^-2 * tanh(3-t/s) (reference `scripts/pulse_RS.py:49-50`)."""
    kind = _abi.DFX_LOAD_SECH2

    def __init__(self, amplitude, sharpness, load_vector=None):
        super().__init__(load_vector)
        self.consts = (float(amplitude), float(sharpness))
