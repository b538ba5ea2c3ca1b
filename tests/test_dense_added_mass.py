"""Dense 6x6 added mass: ``calculate_added_mass`` (numba_hydrodynamics.py:219-253)
takes a full body-frame matrix; the wrapper only ever builds a diagonal one.

Chain of evidence: unmodified reference with a full matrix (golden, tests/golden/reference_numba_dense_am.npz,
made by oracle/make_golden_dense.py) -> C oracle -> the model header on the host (CPU) -> the CUDA engine (GPU).
"""
import os

import numpy as np
import pytest

from silver2_isaacsim_b200 import workloads as W
from tests import emul, scoring

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_numba_dense_am.npz")
SLOTS = np.array([0, 1, 2, 3, 1, 0, 2], dtype=np.int32)


@pytest.fixture(scope="module")
def dense_golden():
    return np.load(GOLDEN)


@pytest.fixture()
def dense_oracle(oracle, dense_golden):
    oracle.set_added_mass_dense(dense_golden["matrices"], dense_golden["slot_type"])
    yield oracle
    oracle.set_added_mass_dense(None)


def _workload(g):
    wl = W.heterogeneous_boxes(768, seed=W.SEED_BASE + 404)
    assert np.array_equal(wl.pos.astype(np.float64), g["pos"])  # the generator's inputs, rebuilt from the seed
    return wl


def test_oracle_matches_reference_with_full_matrix(dense_oracle, dense_golden):
    g = dense_golden
    out = dense_oracle.components(g["ctor"], g["pos"], g["quat"], g["v"], g["w"], g["a"], g["al"])
    assert not g["raised"].any()
    for name in ("added_mass_force", "added_mass_torque", "drag_force", "lift_force", "buoyancy_force"):
        ref = g[name]
        err = np.abs(out[name] - ref).max(axis=1)
        assert (err <= 1e-12 * np.maximum(np.abs(ref).max(axis=1), 1.0)).all(), name
    # the matrices really are exercised: off-diagonal terms change the answer
    dense_oracle.set_added_mass_dense(None)
    diag = dense_oracle.components(g["ctor"], g["pos"], g["quat"], g["v"], g["w"], g["a"], g["al"])
    wet = g["sub_ratio"] > 0
    assert np.abs(diag["added_mass_force"][wet] - g["added_mass_force"][wet]).max() > 1.0


def test_model_header_dense(dense_oracle, dense_golden):
    g = dense_golden
    wl = _workload(g)
    ref = dense_oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                            wl.prev_lin, wl.prev_ang, wl.dt)
    F, T = emul.step_dense(wl, emul.MODE_FP64, g["matrices"], g["slot_type"])
    scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
    pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
    assert scoring.fp64_ok(F, ref.force, scale).all()
    assert scoring.fp64_ok(T, ref.torque, scale, extra=pn).all()
    F, T = emul.step_dense(wl, emul.MODE_FP32_FAST, g["matrices"], g["slot_type"])
    scoring.assert_fp32(F, ref.force, "dense force")
    scoring.assert_fp32(T, ref.torque, "dense torque")


def test_diagonal_matrix_reproduces_the_wrapper(oracle):
    """A dense table holding each body's own wrapper diagonal gives the diagonal path's answer."""
    wl = W.uniform_small_batch(256)
    c = wl.coeff_per_body()[0].astype(np.float64)
    vol = c[0] * c[1] * c[2]
    d = np.array([vol * c[7] * wl.rho] * 3 + [vol * (c[1] ** 2 + c[2] ** 2) * c[8] * wl.rho,
                                              vol * (c[0] ** 2 + c[2] ** 2) * c[8] * wl.rho,
                                              vol * (c[0] ** 2 + c[1] ** 2) * c[8] * wl.rho])
    F0, T0, _, _ = emul.step(wl, emul.MODE_FP64)
    F1, T1 = emul.step_dense(wl, emul.MODE_FP64, np.diag(d)[None], [0])
    assert np.allclose(F0, F1, rtol=1e-13, atol=1e-12) and np.allclose(T0, T1, rtol=1e-13, atol=1e-12)


# --------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_engine_dense_added_mass(dense_oracle, dense_golden, built_lib, dtype):
    import torch

    from silver2_isaacsim_b200 import HydroEngine

    g = dense_golden
    wl = _workload(g)
    ref = dense_oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                            wl.prev_lin, wl.prev_ang, wl.dt)
    td = getattr(torch, dtype)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev).to(td).contiguous()
    e = HydroEngine(wl.n, dtype=td, device=dev)
    e.set_workload_params(wl)
    e.set_added_mass_dense(g["matrices"], g["slot_type"])
    e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
    F, T = e.step(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), wl.dt)
    assert e.last_kernel == "direct"
    F, T = F.double().cpu().numpy(), T.double().cpu().numpy()
    if dtype == "float64":
        scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
        pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
        assert scoring.fp64_ok(F, ref.force, scale).all()
        assert scoring.fp64_ok(T, ref.torque, scale, extra=pn).all()
    else:
        scoring.assert_fp32(F, ref.force, "dense force")
        scoring.assert_fp32(T, ref.torque, "dense torque")
    # the carried state is the same as on the diagonal path
    assert np.array_equal(e.prev_velocities().double().cpu().numpy()[:, :3], wl.lin_vel.astype(np.float64))

    # switching back restores the wrapper's diagonal
    e.set_added_mass_dense(None)
    dense_oracle.set_added_mass_dense(None)
    ref0 = dense_oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                             wl.prev_lin, wl.prev_ang, wl.dt)
    e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
    F0, _ = e.step(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), wl.dt)
    ok = scoring.fp32_ok(F0.double().cpu().numpy(), ref0.force)
    assert ok.all()


@pytest.mark.gpu
def test_engine_dense_large_batch_and_robots(oracle, dense_golden, built_lib):
    """Past the tile threshold the dense path still routes to the per-body kernel, robot wrenches included."""
    import torch

    from silver2_isaacsim_b200 import HydroEngine

    M = dense_golden["matrices"]
    wl = W.hexapod_envs(4096)  # 77824 bodies > 148 * 256
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(a, device=dev)
    oracle.set_added_mass_dense(M, SLOTS)
    try:
        ref = oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                          wl.prev_lin, wl.prev_ang, wl.dt)
    finally:
        oracle.set_added_mass_dense(None)
    refw = oracle.robot_wrench(wl.pos, ref.force, ref.torque, 19)
    e = HydroEngine(wl.n, device=dev)
    e.set_workload_params(wl)
    e.set_articulation(19)
    e.set_added_mass_dense(M, SLOTS)
    e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
    Wr = torch.empty(wl.n // 19, 6, device=dev)
    F, T, _ = e.step(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), wl.dt, out_robot_wrench=Wr)
    assert e.last_kernel == "direct"
    scoring.assert_fp32(F.double().cpu().numpy(), ref.force, "dense force")
    scoring.assert_fp32(T.double().cpu().numpy(), ref.torque, "dense torque")
    w = Wr.double().cpu().numpy()
    den = np.abs(refw).max(axis=1, keepdims=True)
    assert (np.abs(w - refw) <= 1e-4 * den + 1e-4).mean() > 0.999
