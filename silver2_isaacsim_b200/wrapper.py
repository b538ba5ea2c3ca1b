"""Drop-in twins of the reference force-engine wrappers.

``WarpHydrodynamicsWrapper`` keeps the constructor signature, attribute names and
``calculate_hydrodynamic_forces`` contract of the reference's GPU wrapper
(/root/reference/src/scripts/physics/warp_hydrodynamics_wrapper.py:10-12, :79-132):
CUDA torch tensors of shape (N,3)/(N,4) in, eight (N,3) CUDA tensors out -- but for
any N in one launch instead of ``dim=1``.

``NumbaHydrodynamicsWrapper`` keeps the CPU flavour's contract
(numba_hydrodynamics_wrapper.py:9-10, :34-53): (3,)/(4,) NumPy arrays in, the
nine-tuple ``(buoyancy_force, drag_force, lift_force, drag_torque, added_mass_force,
added_mass_torque, center_of_buoyancy, center_of_pressure, sub_ratio)`` of float64
NumPy arrays + a Python float out.  It evaluates on the GPU in fp64 mode (there is no
CPU path in this package) and also accepts (N,3) batches.

Where the unmodified reference raises ``TypeError`` (wet body with speed <= 1e-6,
numba_hydrodynamics.py:118,143) these wrappers return the evident intent
(cop = cob, projected area 0); pass ``strict_reference_errors=True`` to raise instead.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import HydroEngine


class _WrapperBase:
    def __init__(self, width, depth, height, linear_drag_coefficient, angular_drag_coefficient,
                 linear_damping, angular_damping, water_density, gravity, linear_mass_coeff,
                 angular_mass_coeff, lift_coefficient, device="cuda:0", dtype=torch.float32,
                 mass: float = 1.0, strict_reference_errors: bool = False, warp_compat: bool = False):
        # attribute names of numba_hydrodynamics_wrapper.py:12-24
        self.width, self.depth, self.height = width, depth, height
        self.total_volume = width * depth * height
        self.water_density, self.gravity = water_density, gravity
        self.linear_drag_coefficient = linear_drag_coefficient
        self.angular_drag_coefficient = angular_drag_coefficient
        self.linear_damping, self.angular_damping = linear_damping, angular_damping
        self.lift_coefficient = lift_coefficient
        self.linear_mass_coeff, self.angular_mass_coeff = linear_mass_coeff, angular_mass_coeff
        self.device = device
        self.mass = mass
        self.strict_reference_errors = strict_reference_errors
        self.warp_compat = warp_compat  # reproduce the Warp twin's deviations
        self._dtype = dtype
        self._ctor = [width, depth, height, linear_drag_coefficient, angular_drag_coefficient,
                      linear_damping, angular_damping, water_density, gravity, linear_mass_coeff,
                      angular_mass_coeff, lift_coefficient]
        self._engines = {}

    def _engine(self, n: int) -> HydroEngine:
        e = self._engines.get(n)
        if e is None:
            e = HydroEngine(n, dtype=self._dtype, device=self.device)
            e.set_params_uniform(self._ctor, self.mass)
            e.set_warp_compat(self.warp_compat)
            self._engines[n] = e
        return e


class WarpHydrodynamicsWrapper(_WrapperBase):
    """GPU flavour: torch CUDA tensors in, eight torch CUDA tensors out."""

    def calculate_hydrodynamic_forces(self, t_position, t_orientation, t_linear_velocity,
                                      t_angular_velocity, t_linear_acceleration, t_angular_acceleration):
        n = t_position.shape[0]
        e = self._engine(n)
        cast = lambda t: t.to(device=e.device, dtype=e.dtype).contiguous()
        out = e.components(cast(t_position), cast(t_orientation), cast(t_linear_velocity),
                           cast(t_angular_velocity), cast(t_linear_acceleration),
                           cast(t_angular_acceleration), return_flags=self.strict_reference_errors)
        if self.strict_reference_errors and bool(out[9].any()):
            raise TypeError("reference defect: wet body with speed <= 1e-6 "
                            "(calculate_pressure_and_area returns None)")
        return tuple(out[:8])  # warp_hydrodynamics_wrapper.py:123-132 returns 8 tensors (no sub_ratio)


class NumbaHydrodynamicsWrapper(_WrapperBase):
    """CPU-flavour contract (NumPy in / NumPy out), evaluated in fp64 on the GPU."""

    def __init__(self, *args, **kw):
        kw.setdefault("dtype", torch.float64)
        super().__init__(*args, **kw)

    def calculate_hydrodynamic_forces(self, position, orientation_quat, linear_vel, angular_vel,
                                      linear_accel, angular_accel):
        single = np.ndim(position) == 1
        prep = lambda a, c: np.asarray(a, dtype=np.float64).reshape(-1, c)
        arrs = [prep(position, 3), prep(orientation_quat, 4), prep(linear_vel, 3), prep(angular_vel, 3),
                prep(linear_accel, 3), prep(angular_accel, 3)]
        e = self._engine(arrs[0].shape[0])
        dev = [torch.as_tensor(a).to(device=e.device, dtype=e.dtype) for a in arrs]
        out = e.components(*dev, return_flags=True)
        flags = out[9].cpu().numpy()
        if self.strict_reference_errors and flags.any():
            raise TypeError("reference defect: wet body with speed <= 1e-6 "
                            "(calculate_pressure_and_area returns None)")
        host = [o.cpu().numpy().astype(np.float64) for o in out[:9]]
        if single:
            return tuple(h[0] for h in host[:8]) + (float(host[8][0]),)
        return tuple(host)
