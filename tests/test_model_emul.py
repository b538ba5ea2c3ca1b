"""Per-body arithmetic of the CUDA kernels (csrc/h2o_model.cuh, host-instantiated) vs the
float64 oracle, on the CPU.  Same header, same precision policies as the device code."""
import numpy as np
import pytest

from silver2_isaacsim_b200 import workloads as W
from tests import emul, scoring

N = 60000


def _ref(oracle, wl):
    return oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                       wl.prev_lin, wl.prev_ang, wl.dt)


@pytest.mark.parametrize("make", [lambda: W.heterogeneous_boxes(N), lambda: W.hexapod_envs(N // 19),
                                  lambda: W.sharded_robots(N // 19), lambda: W.uniform_small_batch(4096)],
                         ids=["C3", "C2", "C4", "C5"])
def test_fp32_mode_policy(oracle, make):
    wl = make()
    ref = _ref(oracle, wl)
    F, T, _, _ = emul.step(wl, emul.MODE_FP32_FAST)  # what the fused fp32 step kernels run
    scoring.assert_fp32(F, ref.force, f"{wl.name} force")
    scoring.assert_fp32(T, ref.torque, f"{wl.name} torque")


def test_fp64_mode(oracle):
    for wl, near in ((W.heterogeneous_boxes(N, dtype=np.float64, xy_range=1.0), True),
                     (W.heterogeneous_boxes(N), False), (W.hexapod_envs(N // 19), False)):
        ref = _ref(oracle, wl)
        F, T, _, _ = emul.step(wl, emul.MODE_FP64)
        scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
        assert scoring.fp64_ok(F, ref.force, scale).all()
        # strict torque criterion near the origin; the reference's own world-space lever arms
        # lose |p| * 1e-16 / arm elsewhere (SURVEY.md 8(d) torque caveat)
        if near:
            assert scoring.fp64_ok(T, ref.torque, scale).all()
        pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
        assert scoring.fp64_ok(T, ref.torque, scale, extra=pn).all()


def test_keypoint_mask_matches_explicit_keypoints():
    """The 13-sum / 27-compare waterline equals testing the 27 reference keypoints one by one."""
    wl = W.heterogeneous_boxes(20000, dtype=np.float64)
    _, _, comp, masks = emul.step(wl, emul.MODE_FP64)
    q, p, h = wl.quat_xyzw, wl.pos, wl.coeff[:, :3] / 2
    x, y, z, w = q.T
    r2 = np.stack([x * (z + z) - w * (y + y), y * (z + z) + w * (x + x), 1 - (x * (x + x) + y * (y + y))], 1)
    expect = np.zeros(wl.n, np.uint32)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                zz = ((r2[:, 0] * (i * h[:, 0]) + r2[:, 1] * (j * h[:, 1])) + r2[:, 2] * (k * h[:, 2])) + p[:, 2]
                expect |= (zz < 0).astype(np.uint32) << np.uint32((i + 1) + 3 * (j + 1) + 9 * (k + 1))
    assert (expect == masks).all()


def test_all_fp32_arithmetic_is_not_enough(oracle):
    """Documents WHY the waterline is carried in fp64: pure-fp32 arithmetic misses the bound."""
    wl = W.heterogeneous_boxes(N)
    ref = _ref(oracle, wl)
    F, _, _, _ = emul.step(wl, emul.MODE_ALL_FP32)
    assert (~scoring.fp32_ok(F, ref.force)).sum() > 0
