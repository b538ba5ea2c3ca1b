"""Parity far outside the benchmark distribution: dimensions over five decades, speeds from creeping to
violent, flat plates and needles, strongly non-unit quaternions, depths from the waterline to the abyss.
Same header as the CUDA kernels (host instantiation) against the float64 oracle, on the CPU; the GPU twin
of this test is tests/test_gpu_parity.py::test_stress_distribution_gpu."""
import numpy as np
import pytest

from silver2_isaacsim_b200 import params as P
from silver2_isaacsim_b200 import workloads as W
from tests import emul, scoring


def stress_workload(n=40_000, seed=77, dtype=np.float32):
    rng = np.random.default_rng(seed)
    wl = W.heterogeneous_boxes(n, seed=seed, dtype=dtype)
    lu = lambda lo, hi, size: np.exp(rng.uniform(np.log(lo), np.log(hi), size=size))
    coeff = np.asarray(wl.coeff, dtype=np.float64)
    dims = lu(1e-2, 50.0, (n, 3))                      # 1 cm pebbles to 50 m hulls, plates and needles
    coeff[:, 0:3] = dims
    coeff[:, 5] = lu(1e-2, 5e3, n)                     # linearDamping
    coeff[:, 6] = lu(1e-2, 5e3, n)                     # angularDamping
    coeff[:, 10] = lu(0.05, 5.0, n) * wl.rho * dims.prod(axis=1)
    pos = np.asarray(wl.pos, dtype=np.float64)
    pos[:, 0:2] = rng.uniform(-2000, 2000, (n, 2))
    kind = rng.integers(0, 4, n)
    ext = dims.max(axis=1)
    z = np.where(kind == 0, rng.uniform(-1.2, 1.2, n) * ext,          # around the waterline
        np.where(kind == 1, rng.uniform(-0.02, 0.02, n) * ext,        # grazing it
        np.where(kind == 2, -lu(1.0, 6000.0, n), lu(0.1, 100.0, n)))) # deep / airborne
    pos[:, 2] = z
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q *= 1.0 + rng.uniform(-2e-3, 2e-3, (n, 1))        # the reference never normalises
    speed = lu(1e-5, 30.0, (n, 1))
    v = rng.normal(size=(n, 3)); v *= speed / np.linalg.norm(v, axis=1, keepdims=True)
    spin = lu(1e-5, 30.0, (n, 1))
    w = rng.normal(size=(n, 3)); w *= spin / np.linalg.norm(w, axis=1, keepdims=True)
    acc = rng.normal(size=(n, 3)) * lu(1e-3, 50.0, (n, 1))
    aacc = rng.normal(size=(n, 3)) * lu(1e-3, 50.0, (n, 1))
    c = lambda a: a.astype(dtype)
    v, w = c(v), c(w)
    return W.Workload(name="stress", dt=wl.dt, rho=wl.rho, g=wl.g, pos=c(pos), quat_xyzw=c(q), lin_vel=v, ang_vel=w,
                      prev_lin=c(v.astype(np.float64) - wl.dt * acc), prev_ang=c(w.astype(np.float64) - wl.dt * aacc),
                      coeff=c(coeff), meta={"seed": seed})


def _ref(oracle, wl):
    return oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                       wl.prev_lin, wl.prev_ang, wl.dt)


def test_stress_fp64_mode(oracle):
    wl = stress_workload(dtype=np.float64)
    ref = _ref(oracle, wl)
    F, T, _, _ = emul.step(wl, emul.MODE_FP64)
    scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
    assert np.isfinite(F).all() and np.isfinite(T).all()
    assert scoring.fp64_ok(F, ref.force, scale).all()
    pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
    # arms of a 50 m hull: the reference's own rounding of R k + p scales with the half-extent too
    pn += np.asarray(wl.coeff, dtype=np.float64)[:, :3].max(axis=1) * np.abs(ref.force).max(axis=1)
    assert scoring.fp64_ok(T, ref.torque, scale, extra=pn).all()


def test_stress_fp32_mode(oracle):
    wl = stress_workload()
    ref = _ref(oracle, wl)
    F, T, _, _ = emul.step(wl, emul.MODE_FP32_FAST)
    assert np.isfinite(F).all() and np.isfinite(T).all()
    okF = scoring.fp32_ok(F, ref.force)
    okT = scoring.fp32_ok(T, ref.torque)
    errF, denF = scoring.vec_err(F, ref.force)
    errT, denT = scoring.vec_err(T, ref.torque)
    print(f"stress fp32: force pass {okF.mean():.6f} (worst {np.max(errF / np.maximum(1e-5 * denF, 1e-6)):.2f}x), "
          f"torque pass {okT.mean():.6f} (worst {np.max(errT / np.maximum(1e-5 * denT, 1e-6)):.2f}x)")
    assert okF.all() and okT.all()   # non-unit quaternions: every body is re-evaluated in float64
