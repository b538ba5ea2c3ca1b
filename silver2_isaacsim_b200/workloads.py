"""Seeded synthetic inputs for the five BASELINE.json configurations.

All generators draw in ``dtype`` (float32 by default) so that the very same
numbers, up-cast exactly, feed the float64 oracle.  Nothing here touches a GPU.

C1  single README-default buoy (1 m cube), 10 000-step free-body rollout
C2  SILVER2 hexapod (19 bodies) x 4096 envs, part-type table parameters
C3  2^20 heterogeneous boxes per GPU, per-body coefficient records
C4  16.8 M bodies (8 x 110 592 robots x 19), robot-contiguous shards, +-20 % per-robot jitter
C5  1024 bodies, uniform README parameters (small-batch latency)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import params as P

SEED_BASE = 20261018


@dataclass
class Workload:
    name: str
    dt: float
    rho: float
    g: float
    pos: np.ndarray            # (n,3)
    quat_xyzw: np.ndarray      # (n,4)
    lin_vel: np.ndarray        # (n,3)
    ang_vel: np.ndarray        # (n,3)
    prev_lin: np.ndarray       # (n,3) previous-step velocity (added-mass derivative)
    prev_ang: np.ndarray       # (n,3)
    coeff: Optional[np.ndarray] = None       # (n,11) per-body records (COEFF_FIELDS) ...
    table: Optional[np.ndarray] = None       # ... or (n_types,11) part-type table
    slot_type: Optional[np.ndarray] = None   # (bodies_per_robot,) int32 slot -> type
    bodies_per_robot: int = 0                # 0 = no articulation structure
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    def coeff_per_body(self) -> np.ndarray:
        """(n,11) float64 records whichever parameter mode the workload uses."""
        if self.coeff is not None:
            return np.asarray(self.coeff, dtype=np.float64)
        idx = np.tile(self.slot_type, self.n // len(self.slot_type))
        return np.asarray(self.table, dtype=np.float64)[idx]

    def ctor_rows(self) -> np.ndarray:
        return P.coeff_to_ctor_rows(self.coeff_per_body(), self.rho, self.g)

    def masses(self) -> np.ndarray:
        return self.coeff_per_body()[:, 10].copy()

    def transforms(self) -> np.ndarray:
        """PhysX tensor-API layout (n,7) = [p, q_xyzw]."""
        return np.concatenate([self.pos, self.quat_xyzw], axis=1)

    def velocities(self) -> np.ndarray:
        """PhysX tensor-API layout (n,6) = [v, omega]."""
        return np.concatenate([self.lin_vel, self.ang_vel], axis=1)


def _unit_quats(rng, n, dtype):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(dtype)


def _velocities(rng, n, dtype):
    sig = np.array([0.05, 0.5, 2.0])
    v = rng.standard_normal((n, 3)) * sig[rng.integers(0, 3, size=n)][:, None]
    w = rng.standard_normal((n, 3)) * sig[rng.integers(0, 3, size=n)][:, None]
    return v.astype(dtype), w.astype(dtype)


def _prev_from_accel(rng, v, w, dt, dtype):
    a = rng.standard_normal(v.shape) * 5.0
    al = rng.standard_normal(w.shape) * 5.0
    return (v.astype(np.float64) - dt * a).astype(dtype), (w.astype(np.float64) - dt * al).astype(dtype)


def hexapod_envs(n_envs: int = 4096, seed: int = SEED_BASE + 2, dtype=np.float32,
                 jitter: float = 0.0, name: str = "C2") -> Workload:
    """C2 (and, with ``jitter``, C4): robots of 19 bodies in HEXAPOD_SLOTS order."""
    rng = np.random.default_rng(seed)
    table, slot_type, rho, g = P.hexapod_table()
    B = len(slot_type)
    n = n_envs * B
    dt = 1.0 / 120.0  # locomotion scene, SURVEY.md Appendix D
    base = np.empty((n_envs, 3))
    base[:, 0:2] = rng.uniform(-50.0, 50.0, size=(n_envs, 2))
    base[:, 2] = rng.uniform(-20.0, 0.6, size=n_envs)
    off = rng.uniform(-0.4, 0.4, size=(n_envs, B, 3))
    off[:, 0, :] = 0.0  # slot 0 is the robot base
    pos = (base[:, None, :] + off).reshape(n, 3).astype(dtype)
    quat = _unit_quats(rng, n, dtype)
    v, w = _velocities(rng, n, dtype)
    pl, pa = _prev_from_accel(rng, v, w, dt, dtype)
    wl = Workload(name=name, dt=dt, rho=rho, g=g, pos=pos, quat_xyzw=quat, lin_vel=v, ang_vel=w,
                  prev_lin=pl, prev_ang=pa, bodies_per_robot=B,
                  meta={"n_envs": n_envs, "seed": seed})
    if jitter > 0.0:
        # heterogeneous SoA: per-robot multiplicative jitter on every coefficient
        rec = table[np.tile(slot_type, n_envs)]
        jit = rng.uniform(1.0 - jitter, 1.0 + jitter, size=(n_envs, 1, P.N_COEFF))
        wl.coeff = (rec.reshape(n_envs, B, P.N_COEFF) * jit).reshape(n, P.N_COEFF).astype(dtype)
    else:
        wl.table = table.astype(dtype)
        wl.slot_type = slot_type
    return wl


def heterogeneous_boxes(n: int = 1 << 20, seed: int = SEED_BASE + 3, dtype=np.float32,
                        xy_range: float = 50.0, name: str = "C3") -> Workload:
    """C3: randomised dimensions/coefficients, ~60 % fully wet / 30 % partial / 10 % dry."""
    rng = np.random.default_rng(seed)
    rho, g = 1025.0, 9.81
    dt = 1.0 / 120.0
    dims = rng.uniform(0.05, 2.0, size=(n, 3))
    coeff = np.empty((n, P.N_COEFF))
    coeff[:, 0:3] = dims
    coeff[:, 3] = rng.uniform(0.5, 1.5, size=n)     # linearDragCoefficient
    coeff[:, 4] = rng.uniform(0.05, 1.0, size=n)    # angularDragCoefficient
    coeff[:, 5] = rng.uniform(5.0, 400.0, size=n)   # linearDamping
    coeff[:, 6] = rng.uniform(1.0, 200.0, size=n)   # angularDamping
    coeff[:, 7] = rng.uniform(0.0, 0.3, size=n)     # linearAddedMassCoefficient
    coeff[:, 8] = rng.uniform(0.0, 0.2, size=n)     # angularAddedMassCoefficient
    coeff[:, 9] = rng.uniform(0.0, 1.0, size=n)     # liftCoefficient
    coeff[:, 10] = rng.uniform(0.3, 1.5, size=n) * rho * dims.prod(axis=1)  # mass
    pos = np.empty((n, 3))
    pos[:, 0:2] = rng.uniform(-xy_range, xy_range, size=(n, 2))
    pos[:, 2] = rng.uniform(-1.5, 1.0, size=n) * dims.max(axis=1)
    quat = _unit_quats(rng, n, dtype)
    v, w = _velocities(rng, n, dtype)
    pl, pa = _prev_from_accel(rng, v, w, dt, dtype)
    return Workload(name=name, dt=dt, rho=rho, g=g, pos=pos.astype(dtype), quat_xyzw=quat,
                    lin_vel=v, ang_vel=w, prev_lin=pl, prev_ang=pa, coeff=coeff.astype(dtype),
                    meta={"seed": seed})


def sharded_robots(n_robots: int, seed: int = SEED_BASE + 4, dtype=np.float32) -> Workload:
    """C4 shard: whole robots, C2 parameters with +-20 % per-robot jitter (per-body records)."""
    return hexapod_envs(n_robots, seed=seed, dtype=dtype, jitter=0.2, name="C4")


def uniform_small_batch(n: int = 1024, seed: int = SEED_BASE + 5, dtype=np.float32) -> Workload:
    """C5: README-default parameters for every body, C3 state distribution."""
    wl = heterogeneous_boxes(n, seed=seed, dtype=dtype, name="C5")
    p = P.HydroParams()
    mass = 0.5 * p.waterDensity * p.xDimension * p.yDimension * p.zDimension
    wl.coeff = None
    wl.table = np.asarray([p.coeff_record(mass)], dtype=dtype)
    wl.slot_type = np.zeros(1, dtype=np.int32)
    # re-draw the depth for the 1 m cube so the wet/partial/dry mix is kept
    rng = np.random.default_rng(seed + 1000)
    wl.pos[:, 2] = rng.uniform(-1.5, 1.0, size=n).astype(dtype)
    return wl


def readme_buoy(dtype=np.float64) -> Workload:
    """C1 initial state: README-default 1 m cube, m = rho V / 2, released from z = 1 m."""
    p = P.HydroParams()
    mass = 0.5 * p.waterDensity
    q = np.array([[0.05, 0.02, 0.0, 1.0]])
    q /= np.linalg.norm(q)
    z3 = np.zeros((1, 3), dtype=dtype)
    return Workload(name="C1", dt=1.0 / 60.0, rho=p.waterDensity, g=p.gravity,
                    pos=np.array([[0.0, 0.0, 1.0]], dtype=dtype), quat_xyzw=q.astype(dtype),
                    lin_vel=z3.copy(), ang_vel=z3.copy(), prev_lin=z3.copy(), prev_ang=z3.copy(),
                    table=np.asarray([p.coeff_record(mass)], dtype=dtype),
                    slot_type=np.zeros(1, dtype=np.int32),
                    meta={"steps": 10000, "inertia_diag": [mass / 6.0] * 3})
