"""ctypes/NumPy front-end of the CPU parity oracle (``hydro_oracle.c``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``silver2_isaacsim_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or the
timed CPU baseline.

The arithmetic follows /root/reference/src/scripts/physics/numba_hydrodynamics.py
(seven ``@njit`` functions), numba_hydrodynamics_wrapper.py:9-112 (geometry /
added-mass precompute) and hydrodynamics_behavior.py:194-238 (numeric tail) --
see the per-function citations in ``hydro_oracle.c``.

Parity pin: ``tests/golden/reference_numba_golden.npz`` was produced by running
the untouched reference Numba code in the build container
(``oracle/make_golden.py``); ``tests/test_oracle_golden.py`` checks this oracle
against it and against the SURVEY.md Appendix B known-answer vectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhydro_oracle.so")

# order of the reference wrapper ctor, numba_hydrodynamics_wrapper.py:9-10
CTOR_FIELDS = (
    "width", "depth", "height", "linear_drag_coefficient", "angular_drag_coefficient",
    "linear_damping", "angular_damping", "water_density", "gravity",
    "linear_mass_coeff", "angular_mass_coeff", "lift_coefficient",
)

OUT_DTYPE = np.dtype(
    [
        ("buoyancy_force", "f8", 3),
        ("drag_force", "f8", 3),
        ("lift_force", "f8", 3),
        ("drag_torque", "f8", 3),
        ("added_mass_force", "f8", 3),
        ("added_mass_torque", "f8", 3),
        ("center_of_buoyancy", "f8", 3),
        ("center_of_pressure", "f8", 3),
        ("sub_ratio", "f8"),
        ("reference_raises", "i4"),
        ("_pad", "i4"),
    ],
    align=True,
)
COMPONENT_NAMES = OUT_DTYPE.names[:8]


def build(force: bool = False) -> str:
    """Compile ``hydro_oracle.c`` into ``oracle/_build`` (gcc, strict IEEE)."""
    src = os.path.join(_HERE, "hydro_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        assert L.oracle_sizeof_out() == OUT_DTYPE.itemsize, "oracle_out_t layout mismatch"
        _lib = L
    return _lib


def set_warp_compat(enable: bool) -> None:
    """Switch the oracle to the Warp twin's deviations (C1, C3)."""
    lib().oracle_set_warp_compat(ctypes.c_int(1 if enable else 0))


def set_added_mass_dense(matrices=None, slot_type=None) -> None:
    """Dense 6x6 added-mass matrices for the batched drivers: body i uses
    ``matrices[slot_type[i % len(slot_type)]]``; ``None`` restores the wrapper's diagonal."""
    if matrices is None:
        lib().oracle_set_added_mass_dense(0, None, 0, None)
        return
    m = np.ascontiguousarray(np.asarray(matrices, dtype=np.float64)).reshape(-1, 6, 6)
    st = np.ascontiguousarray(np.zeros(1) if slot_type is None else slot_type, dtype=np.int32)
    rc = lib().oracle_set_added_mass_dense(ctypes.c_int(m.shape[0]), _ptr(m), ctypes.c_int(st.size), _ptr(st))
    if rc != 0:
        raise ValueError("bad dense added-mass table")


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _f64(a, shape_tail):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    assert a.shape[1:] == shape_tail, (a.shape, shape_tail)
    return a


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ctor_array(ctor, n):
    ctor = np.ascontiguousarray(np.asarray(ctor, dtype=np.float64))
    if ctor.ndim == 1:
        assert ctor.shape == (12,)
        return ctor, 0
    assert ctor.shape == (n, 12), ctor.shape
    return ctor, 12


def components(ctor, pos, quat_xyzw, lin_vel, ang_vel, lin_acc, ang_acc, n_threads: int = 0):
    """Batched ``solve_hydrodynamics`` (numba_hydrodynamics.py:255-314).

    ``ctor``: (12,) or (n,12) in ``CTOR_FIELDS`` order.  Returns a structured
    array of ``OUT_DTYPE`` (one record per body).
    """
    pos = _f64(pos, (3,))
    n = pos.shape[0]
    quat = _f64(quat_xyzw, (4,))
    v, w = _f64(lin_vel, (3,)), _f64(ang_vel, (3,))
    a, al = _f64(lin_acc, (3,)), _f64(ang_acc, (3,))
    ctor, stride = _ctor_array(ctor, n)
    out = np.zeros(n, dtype=OUT_DTYPE)
    lib().oracle_components_batch(
        ctypes.c_int64(n), _ptr(ctor), ctypes.c_int64(stride), _ptr(pos), _ptr(quat), _ptr(v),
        _ptr(w), _ptr(a), _ptr(al), _ptr(out), ctypes.c_int(n_threads))
    return out


@dataclass
class StepResult:
    force: np.ndarray        # (n,3) net force after clamp
    torque: np.ndarray       # (n,3) net torque after clamp
    components: np.ndarray   # OUT_DTYPE records
    flags: np.ndarray        # bit0 = reference raises TypeError (A.8), bit1 = clamp active
    prev_lin: np.ndarray     # updated previous-step velocities
    prev_ang: np.ndarray


def step(ctor, mass, pos, quat, lin_vel, ang_vel, prev_lin, prev_ang, dt,
         quat_order: str = "xyzw", n_threads: int = 0) -> StepResult:
    """One behaviour step (hydrodynamics_behavior.py:194-238) for n bodies in float64."""
    pos = _f64(pos, (3,))
    n = pos.shape[0]
    quat = _f64(quat, (4,))
    v, w = _f64(lin_vel, (3,)), _f64(ang_vel, (3,))
    pl = np.array(prev_lin, dtype=np.float64, order="C", copy=True)
    pa = np.array(prev_ang, dtype=np.float64, order="C", copy=True)
    assert pl.shape == (n, 3) and pa.shape == (n, 3)
    ctor, stride = _ctor_array(ctor, n)
    mass = np.ascontiguousarray(np.asarray(mass, dtype=np.float64).reshape(-1))
    mstride = 0 if mass.shape[0] == 1 and n != 1 else 1
    if mstride:
        assert mass.shape[0] == n
    F = np.zeros((n, 3))
    T = np.zeros((n, 3))
    comp = np.zeros(n, dtype=OUT_DTYPE)
    flags = np.zeros(n, dtype=np.int32)
    lib().oracle_step_batch(
        ctypes.c_int64(n), _ptr(ctor), ctypes.c_int64(stride), _ptr(mass), ctypes.c_int64(mstride),
        _ptr(pos), _ptr(quat), ctypes.c_int(1 if quat_order == "wxyz" else 0), _ptr(v), _ptr(w),
        _ptr(pl), _ptr(pa), ctypes.c_double(dt), _ptr(F), _ptr(T), _ptr(comp), _ptr(flags),
        ctypes.c_int(n_threads))
    return StepResult(F, T, comp, flags, pl, pa)


def robot_wrench(pos, force, torque, bodies_per_robot: int) -> np.ndarray:
    """(n_robots,6) wrench about each robot's slot-0 body."""
    pos = _f64(pos, (3,))
    F, T = _f64(force, (3,)), _f64(torque, (3,))
    n = pos.shape[0]
    assert n % bodies_per_robot == 0
    r = n // bodies_per_robot
    out = np.zeros((r, 6))
    lib().oracle_robot_wrench(ctypes.c_int64(r), ctypes.c_int64(bodies_per_robot), _ptr(pos),
                              _ptr(F), _ptr(T), _ptr(out))
    return out


# ---------------------------------------------------------------------------
# Independent NumPy restatement of the numeric tail, used to cross-check the C
# tail above (hydrodynamics_behavior.py:194-226, line by line).
# ---------------------------------------------------------------------------
def numpy_tail(comp, pos, mass):
    pos = np.asarray(pos, dtype=np.float64)
    t_b = np.cross(comp["center_of_buoyancy"] - pos, comp["buoyancy_force"])      # :212
    t_d = np.cross(comp["center_of_pressure"] - pos, comp["drag_force"])          # :213
    t_l = np.cross(comp["center_of_pressure"] - pos, comp["lift_force"])          # :214
    F = comp["buoyancy_force"] + comp["drag_force"] + comp["lift_force"] + comp["added_mass_force"]  # :217
    T = t_b + t_d + t_l + comp["drag_torque"] + comp["added_mass_torque"]         # :218
    max_force = np.asarray(mass, dtype=np.float64).reshape(-1, 1) * 500.0         # :221-222
    mag = np.linalg.norm(F, axis=-1, keepdims=True)                               # :223
    scale = np.minimum(max_force / (mag + 1e-6), 1.0)                             # :224
    return F * scale, T * scale, scale[:, 0]
