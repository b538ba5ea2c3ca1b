"""Parity of the CUDA path (through the C ABI) against the float64 oracle and the
reference's golden vectors.  Runs on the B200 box: ``pytest -m gpu``."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from silver2_isaacsim_b200 import params as P
from silver2_isaacsim_b200 import workloads as W
from tests import scoring

pytestmark = pytest.mark.gpu

NAMES = ("buoyancy_force", "drag_force", "lift_force", "drag_torque", "added_mass_force",
         "added_mass_torque", "center_of_buoyancy", "center_of_pressure")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _engine(wl, dtype, dev, kernel="auto", stats=False):
    from silver2_isaacsim_b200 import HydroEngine

    e = HydroEngine(wl.n, dtype=dtype, device=dev)
    e.set_workload_params(wl)
    e.set_kernel(kernel)
    e.enable_stats(stats)
    return e


def _t(a, dtype, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dtype).contiguous()


def _ref(oracle, wl, quat=None, order="xyzw"):
    return oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw if quat is None else quat, wl.lin_vel,
                       wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt, quat_order=order)


def _run_step(e, wl, dtype, dev, layout="split", robot=False):
    e.set_prev(_t(wl.prev_lin, dtype, dev), _t(wl.prev_ang, dtype, dev))
    if layout == "split":
        out = e.step(_t(wl.pos, dtype, dev), _t(wl.quat_xyzw, dtype, dev), _t(wl.lin_vel, dtype, dev),
                     _t(wl.ang_vel, dtype, dev), wl.dt, robot_wrench=robot)
    elif layout == "view":
        out = e.step_view(_t(wl.pos, dtype, dev), _t(wl.quat_xyzw, dtype, dev), _t(wl.velocities(), dtype, dev),
                          wl.dt, robot_wrench=robot)
    else:
        out = e.step_physx(_t(wl.transforms(), dtype, dev), _t(wl.velocities(), dtype, dev), wl.dt,
                           robot_wrench=robot)
    torch.cuda.synchronize()
    return [o.double().cpu().numpy() for o in out]


def _check(wl, dtype, ref, F, T, what):
    if dtype == torch.float32:
        scoring.assert_fp32(F, ref.force, what + " force")
        scoring.assert_fp32(T, ref.torque, what + " torque")
    else:
        scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
        assert scoring.fp64_ok(F, ref.force, scale).all(), what
        pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
        assert scoring.fp64_ok(T, ref.torque, scale, extra=pn).all(), what


# --------------------------------------------------------------------------- fused step
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("kernel", ["tile", "direct"])
@pytest.mark.parametrize("layout", ["split", "physx", "view"])
def test_step_heterogeneous_boxes(oracle, dev, dtype, kernel, layout):
    """C3 distribution, per-body records; n is deliberately not a multiple of the tile size."""
    wl = W.heterogeneous_boxes(100_003, seed=W.SEED_BASE + 33)
    ref = _ref(oracle, wl)
    e = _engine(wl, dtype, dev, kernel)
    F, T = _run_step(e, wl, dtype, dev, layout)
    assert e.last_kernel == kernel
    _check(wl, dtype, ref, F, T, f"C3 {kernel} {layout}")
    # v_prev <- v (hydrodynamics_behavior.py:237-238)
    prev = e.prev_velocities().double().cpu().numpy()
    assert (prev[:, :3] == wl.lin_vel.astype(np.float64)).all() and (prev[:, 3:] == wl.ang_vel.astype(np.float64)).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("kernel", ["tile", "direct"])
def test_step_hexapod_table_and_robot_wrench(oracle, dev, dtype, kernel):
    """C2: part-type table staged in shared memory + per-robot wrench (segmented shuffle)."""
    wl = W.hexapod_envs(2048 + 3)
    ref = _ref(oracle, wl)
    e = _engine(wl, dtype, dev, kernel)
    F, T, Wr = _run_step(e, wl, dtype, dev, "view" if kernel == "tile" else "split", robot=True)
    _check(wl, dtype, ref, F, T, f"C2 {kernel}")
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, wl.bodies_per_robot)
    err, den = scoring.vec_err(Wr, want)
    # the wrench sums 19 bodies: compare against the sum of magnitudes, not the cancelled net
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), wl.bodies_per_robot)
    tol = (1e-5 if dtype == torch.float32 else 1e-11) * np.abs(mag).max(axis=1) * 20
    assert (err <= tol + 1e-6).all(), float((err / (tol + 1e-6)).max())


@pytest.mark.parametrize("every", [1, 5, 97], ids=["all_flagged", "every_5th", "every_97th"])
@pytest.mark.parametrize("params", ["table", "per_body"])
def test_flagged_bodies_with_robot_wrench(oracle, dev, every, params):
    """fp32 mode, tile kernel, per-robot wrench: bodies the fast path flags (here: quaternions far from unit) are
    re-evaluated in float64 at the end of their CTA and the robot sums are patched with the difference.
    every = 1 overflows every CTA's deferred list (bitmap + end-of-CTA sweep); 5 fills it with ~30 bodies per tile
    (list AND sweep); 97 stays inside the list (1 - 2 bodies per tile)."""
    wl = W.hexapod_envs(1024 + 5) if params == "table" else W.sharded_robots(1024 + 5)
    q = wl.quat_xyzw.astype(np.float64).copy()
    q[::every] *= 1.0 + 3e-4                      # |q|^2 - 1 = 6e-4 >> FAST_PATH_MAX_DQ
    wl.quat_xyzw = q.astype(np.float32)
    ref = _ref(oracle, wl)
    e = _engine(wl, torch.float32, dev, "tile", stats=True)
    F, T, Wr = _run_step(e, wl, torch.float32, dev, "split", robot=True)
    assert e.last_kernel == "tile"
    _check(wl, torch.float32, ref, F, T, f"flagged every {every}")
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, wl.bodies_per_robot)
    err, _ = scoring.vec_err(Wr, want)
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), wl.bodies_per_robot)
    tol = 1e-5 * np.abs(mag).max(axis=1) * 20
    assert (err <= tol + 1e-6).all(), float((err / (tol + 1e-6)).max())
    st = e.stats()
    assert st["reevaluated_bodies"] >= (wl.n + every - 1) // every   # every marked body took the float64 path
    # the engine-wide bitmap is clean again: an unflagged second step must not re-evaluate anything extra
    wl2 = W.hexapod_envs(1024 + 5) if params == "table" else W.sharded_robots(1024 + 5)
    e.stats(reset=True)
    F2, T2, _ = _run_step(e, wl2, torch.float32, dev, "split", robot=True)
    _check(wl2, torch.float32, _ref(oracle, wl2), F2, T2, "step after a flagged step")
    assert e.stats()["reevaluated_bodies"] < 0.01 * wl2.n


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_robot_tiles_with_uniform_and_mixed_warps(oracle, dev, dtype):
    """Robot-mode tile kernels skip the keypoint compares (and what hangs on them) for warps in which no lane is cut
    by the surface.  Tiles of 152 bodies built so that every case occurs: whole tiles deep under water / dry, one cut
    lane in an otherwise deep warp, deep and dry lanes mixed (no cut lane: skipped, masks all-or-nothing per lane),
    and bodies whose TOP face lies exactly on z = 0 (identity quaternion, p_z = -h_z: fully submerged by z_max <= 0,
    but the top keypoints are not wet -- such a lane must take the compares)."""
    wl = W.sharded_robots(8 * 12)                       # 12 tiles of 8 robots
    z = wl.pos[:, 2].astype(np.float64)
    tile = np.arange(wl.n) // 152
    lane = np.arange(wl.n) % 152
    z[tile == 0] = -5.0 - 0.01 * lane[tile == 0]          # deep
    z[tile == 1] = 5.0 + 0.01 * lane[tile == 1]           # dry
    z[tile == 2] = -5.0
    z[(tile == 2) & (lane == 40)] = 0.01                  # one cut lane in warp 1
    z[tile == 3] = np.where(lane[tile == 3] % 2 == 0, -5.0, 5.0)   # deep / dry alternating: no lane cut
    z[tile == 4] = -5.0
    top = (tile == 4) & (lane % 9 == 0)                   # top face exactly on the surface
    wl.quat_xyzw[top] = np.array([0, 0, 0, 1], dtype=np.float32)
    z[top] = -0.5 * wl.coeff[top, 2].astype(np.float64)   # -h_z, exact in fp32 (a halved fp32 number)
    z[tile == 5] = -5.0
    wl.quat_xyzw[(tile == 5) & (lane % 5 == 0)] *= np.float32(1.001)  # flagged lanes inside skipped warps
    wl.pos[:, 2] = z.astype(np.float32)
    ref = _ref(oracle, wl)
    e = _engine(wl, dtype, dev, "tile", stats=True)
    F, T, Wr = _run_step(e, wl, dtype, dev, "split", robot=True)
    assert e.last_kernel == "tile"
    _check(wl, dtype, ref, F, T, "uniform / mixed warps")
    assert (F[tile == 1] == 0).all() and (T[tile == 1] == 0).all()
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, wl.bodies_per_robot)
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), wl.bodies_per_robot)
    err, _ = scoring.vec_err(Wr, want)
    assert (err <= (1e-5 if dtype == torch.float32 else 1e-11) * np.abs(mag).max(axis=1) * 20 + 1e-6).all()
    st = e.stats()
    assert st["wet_bodies"] == float((ref.components["sub_ratio"] > 0).sum())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("bpr", [1, 2, 4, 5, 7, 12, 32, 45])
def test_robot_wrench_any_robot_size(oracle, dev, bpr, dtype):
    """The 19-body hexapod has a compiled-in robot size; every other size takes the run-time path
    (whole-robot tiles sized per bodies_per_robot; robots of < 6 bodies have more sums than threads)
    and must give the same per-robot sums."""
    n_robots = 48_000 // bpr + 3
    wl = W.heterogeneous_boxes(n_robots * bpr, seed=500 + bpr, xy_range=2.0)
    e = _engine(wl, dtype, dev, "tile")
    e.set_articulation(bpr)
    F, T, Wr = _run_step(e, wl, dtype, dev, "split", robot=True)
    assert e.last_kernel == "tile"
    ref = _ref(oracle, wl)
    if dtype == torch.float32:
        scoring.assert_fp32(F, ref.force, f"robots of {bpr} force")
        scoring.assert_fp32(T, ref.torque, f"robots of {bpr} torque")
    else:
        _check(wl, dtype, ref, F, T, f"robots of {bpr}")
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, bpr)
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), bpr)
    err, _ = scoring.vec_err(Wr, want)
    tol = (1e-5 if dtype == torch.float32 else 1e-11) * np.abs(mag).max(axis=1) * 20
    assert (err <= tol + 1e-6).all(), float((err / (tol + 1e-6)).max())


@pytest.mark.parametrize("kernel", ["tile", "direct"])
@pytest.mark.parametrize("n_slots,n_types", [(1, 1), (3, 2), (7, 4), (64, 16), (200, 5)])
def test_part_table_any_slot_map(oracle, dev, kernel, n_slots, n_types):
    """Part-type table mode with arbitrary slot maps (body i -> table[slot_type[i % n_slots]]): the
    slot phase must survive tile boundaries and the direct-kernel tail for any n_slots, not just 19."""
    from silver2_isaacsim_b200 import HydroEngine, params as P

    n = 50_003
    wl = W.heterogeneous_boxes(n, seed=900 + n_slots)
    rng = np.random.default_rng(n_slots)
    table = np.asarray(wl.coeff, dtype=np.float64)[rng.choice(n, n_types, replace=False)]
    slots = rng.integers(0, n_types, n_slots).astype(np.int32)
    slots[:min(n_slots, n_types)] = np.arange(min(n_slots, n_types))
    coeff = table[slots[np.arange(n) % n_slots]]
    e = HydroEngine(n, device=dev)
    e.set_globals(wl.rho, wl.g)
    e.set_part_table(table, slots)
    e.set_kernel(kernel)
    F, T = _run_step(e, wl, torch.float32, dev)
    assert e.last_kernel == kernel
    ref = oracle.step(P.coeff_to_ctor_rows(coeff, wl.rho, wl.g), coeff[:, 10].copy(), wl.pos, wl.quat_xyzw,
                      wl.lin_vel, wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt)
    scoring.assert_fp32(F, ref.force, "table force")
    scoring.assert_fp32(T, ref.torque, "table torque")


@pytest.mark.parametrize("kernel", ["tile", "direct"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_stress_distribution_gpu(oracle, dev, dtype, kernel):
    """GPU twin of tests/test_stress_distribution.py: five decades of sizes and speeds, plates and
    needles, deep / airborne / grazing bodies, and quaternions up to 2e-3 away from unit (the reference
    never normalises; fp32 mode takes the world-frame evaluation for those)."""
    from tests.test_stress_distribution import stress_workload

    wl = stress_workload(n=120_001, dtype=np.float32 if dtype == torch.float32 else np.float64)
    ref = _ref(oracle, wl)
    e = _engine(wl, dtype, dev, kernel)
    F, T = _run_step(e, wl, dtype, dev)
    assert e.last_kernel == kernel
    assert np.isfinite(F).all() and np.isfinite(T).all()
    if dtype == torch.float32:
        # every quaternion of this set is far from unit: the fast path flags every body and the float64
        # re-evaluation answers, so the strict bound holds even with |p| = 2 km
        scoring.assert_fp32(F, ref.force, "stress force")
        scoring.assert_fp32(T, ref.torque, "stress torque")
    else:
        scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
        assert scoring.fp64_ok(F, ref.force, scale).all()
        pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
        pn += np.asarray(wl.coeff, dtype=np.float64)[:, :3].max(axis=1) * np.abs(ref.force).max(axis=1)
        assert scoring.fp64_ok(T, ref.torque, scale, extra=pn).all()


@pytest.mark.parametrize("src", [torch.float32, torch.float64], ids=["cols32", "cols64"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_params_struct_of_arrays(oracle, dev, dtype, src):
    """Eleven per-attribute columns give the same engine as the (N,11) record upload."""
    from silver2_isaacsim_b200 import HydroEngine

    wl = W.heterogeneous_boxes(50_001, seed=61)
    coeff = np.asarray(wl.coeff, dtype=np.float64)
    cols = [torch.as_tensor(np.ascontiguousarray(coeff[:, k]), device=dev).to(src).contiguous() for k in range(11)]
    e = HydroEngine(wl.n, dtype=dtype, device=dev)
    e.set_globals(wl.rho, wl.g)
    e.set_params_soa(cols)
    F, T = _run_step(e, wl, dtype, dev)
    e2 = _engine(wl, dtype, dev)
    F2, T2 = _run_step(e2, wl, dtype, dev)
    assert np.array_equal(F, F2) and np.array_equal(T, T2)
    _check(wl, dtype, _ref(oracle, wl), F, T, "soa params")
    with pytest.raises(ValueError):
        e.set_params_soa(cols[:10])


@pytest.mark.parametrize("kernel", ["tile", "direct"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_unequal_robots_by_offsets(oracle, dev, dtype, kernel):
    """Articulation given as offsets: robots of 19 links mixed with single-body buoys (the reference's
    main scene is one SILVER2 + the Obsea buoy) and a few odd sizes; wrench about each robot's first body."""
    rng = np.random.default_rng(3)
    sizes = rng.choice([19, 1, 19, 1, 7, 40, 19], size=4000)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    wl = W.heterogeneous_boxes(int(offsets[-1]), seed=71, xy_range=2.0)
    e = _engine(wl, dtype, dev, kernel)
    e.set_articulation_offsets(offsets)
    assert e.n_robots == len(sizes)
    F, T, Wr = _run_step(e, wl, dtype, dev, robot=True)
    assert e.last_kernel == kernel and Wr.shape == (len(sizes), 6)
    ref = _ref(oracle, wl)
    if dtype == torch.float32:
        scoring.assert_fp32(F, ref.force, "offset robots force")
        scoring.assert_fp32(T, ref.torque, "offset robots torque")
    else:
        _check(wl, dtype, ref, F, T, "offset robots")
    pos, Fr, Tr = wl.pos.astype(np.float64), ref.force, ref.torque
    want = np.zeros((len(sizes), 6)); mag = np.zeros((len(sizes), 6))
    for r in range(len(sizes)):
        sl = slice(offsets[r], offsets[r + 1])
        arm = pos[sl] - pos[offsets[r]]
        want[r, :3] = Fr[sl].sum(0); want[r, 3:] = (Tr[sl] + np.cross(arm, Fr[sl])).sum(0)
        mag[r, :3] = np.abs(Fr[sl]).sum(0); mag[r, 3:] = (np.abs(Tr[sl]) + np.abs(np.cross(arm, Fr[sl]))).sum(0)
    err, _ = scoring.vec_err(Wr, want)
    tol = (1e-5 if dtype == torch.float32 else 1e-11) * mag.max(axis=1) * 20
    assert (err <= tol + 1e-6).all(), float((err / (tol + 1e-6)).max())
    # equal runs again
    e2 = _engine(wl, dtype, dev, kernel)
    with pytest.raises(Exception):
        e2.set_articulation_offsets([0, 5, 5, wl.n])      # empty robot
    with pytest.raises(Exception):
        e2.set_articulation_offsets([0, 5, wl.n - 1])     # does not end at n_bodies


def test_random_configurations(oracle, dev):
    """Seeded sweep over the configuration space: batch sizes around every granule (1 body ... a few
    tiles), ingest layout, parameter mode, kernel request, quaternion order, robot size, precision."""
    from silver2_isaacsim_b200 import HydroEngine, params as P

    rng = np.random.default_rng(2026)
    sizes = [1, 2, 3, 4, 5, 31, 32, 33, 127, 128, 129, 255, 257, 1000, 4097, 37_887, 37_888, 37_889, 75_777, 131_073]
    for case in range(72):
        bpr = int(rng.choice([0, 0, 1, 3, 19, 20]))
        n = int(rng.choice(sizes))
        if bpr:
            n = max(bpr, n // bpr * bpr)
        dtype = torch.float32 if rng.random() < 0.6 else torch.float64
        layout = str(rng.choice(["split", "physx", "view"]))
        kernel = str(rng.choice(["auto", "tile", "direct"]))
        order = str(rng.choice(["xyzw", "wxyz"]))
        table = rng.random() < 0.4
        wl = W.heterogeneous_boxes(n, seed=3000 + case, xy_range=3.0,
                                   dtype=np.float32 if dtype == torch.float32 else np.float64)
        coeff = np.asarray(wl.coeff, dtype=np.float64)
        e = HydroEngine(n, dtype=dtype, device=dev, quat_order=order)
        e.set_globals(wl.rho, wl.g)
        if table:
            n_types, n_slots = int(rng.integers(1, 6)), int(rng.integers(1, 24))
            tab = coeff[rng.choice(n, n_types)]
            slots = rng.integers(0, n_types, n_slots).astype(np.int32)
            coeff = tab[slots[np.arange(n) % n_slots]]
            e.set_part_table(tab, slots)
        else:
            e.set_params_per_body(coeff)
        e.set_articulation(bpr)
        e.set_kernel(kernel)
        what = f"case {case}: n={n} bpr={bpr} {dtype} {layout} {kernel} {order} table={table}"
        quat = wl.quat_xyzw if order == "xyzw" else wl.quat_xyzw[:, [3, 0, 1, 2]]
        wl_in = W.Workload(**{**wl.__dict__, "quat_xyzw": np.ascontiguousarray(quat)})   # what the caller hands over
        out = _run_step(e, wl_in, dtype, dev, layout, robot=bpr > 0)
        F, T = out[0], out[1]
        ref = oracle.step(P.coeff_to_ctor_rows(coeff, wl.rho, wl.g), coeff[:, 10].copy(), wl.pos, wl.quat_xyzw,
                          wl.lin_vel, wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt)
        if dtype == torch.float32:
            scoring.assert_fp32(F, ref.force, what + " force")
            scoring.assert_fp32(T, ref.torque, what + " torque")
        else:
            scale = scoring.force_scale(coeff, wl.rho, wl.g)
            pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
            assert scoring.fp64_ok(F, ref.force, scale).all() and scoring.fp64_ok(T, ref.torque, scale, extra=pn).all(), what
        if bpr:
            want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, bpr)
            mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), bpr)
            err, _ = scoring.vec_err(out[2], want)
            tol = (1e-5 if dtype == torch.float32 else 1e-11) * np.abs(mag).max(axis=1) * 20
            assert (err <= tol + 1e-6).all(), what
        prev = e.prev_velocities().double().cpu().numpy()
        assert (prev[:, :3] == wl.lin_vel.astype(np.float64)).all() and (prev[:, 3:] == wl.ang_vel.astype(np.float64)).all(), what
        e.close()


@pytest.mark.parametrize("kernel", ["tile", "direct"])
def test_carried_velocities_over_several_steps(oracle, dev, kernel):
    """The engine-owned v_prev / w_prev (hydrodynamics_behavior.py:196-202, :237-238) across steps with
    changing velocities and step sizes, a skipped step (dt <= 1e-6 leaves the state alone) included."""
    wl = W.sharded_robots(3000)
    e = _engine(wl, torch.float32, dev, kernel)
    rng = np.random.default_rng(5)
    prev_l, prev_a = np.zeros_like(wl.lin_vel, dtype=np.float64), np.zeros_like(wl.ang_vel, dtype=np.float64)
    pos, quat = _t(wl.pos, torch.float32, dev), _t(wl.quat_xyzw, torch.float32, dev)
    v, w = wl.lin_vel.copy(), wl.ang_vel.copy()
    for k, dt in enumerate([1 / 120, 1 / 60, 1e-7, 1 / 240, 1 / 120]):
        F, T, Wr = e.step(pos, quat, _t(v, torch.float32, dev), _t(w, torch.float32, dev), dt, robot_wrench=True)
        if dt > 1e-6:
            ref = oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, v, w, prev_l, prev_a, dt)
            prev_l, prev_a = ref.prev_lin, ref.prev_ang
            scoring.assert_fp32(F.cpu().numpy(), ref.force, f"step {k} force")
            scoring.assert_fp32(T.cpu().numpy(), ref.torque, f"step {k} torque")
        got = e.prev_velocities().double().cpu().numpy()
        assert (got[:, :3] == prev_l).all() and (got[:, 3:] == prev_a).all(), k
        v = (v + rng.normal(size=v.shape).astype(np.float32) * 0.05).astype(np.float32)
        w = (w + rng.normal(size=w.shape).astype(np.float32) * 0.05).astype(np.float32)


def test_step_sharded_robots_per_body_records(oracle, dev):
    """C4 shard: heterogeneous per-robot jitter (per-body records) + robot wrench, PhysX layout."""
    wl = W.sharded_robots(4099)
    ref = _ref(oracle, wl)
    e = _engine(wl, torch.float32, dev, "tile")
    e.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
    F, T, Wr = e.step_physx(_t(wl.transforms(), torch.float32, dev), _t(wl.velocities(), torch.float32, dev), wl.dt,
                            robot_wrench=True)
    torch.cuda.synchronize()
    _check(wl, torch.float32, ref, F.double().cpu().numpy(), T.double().cpu().numpy(), "C4")
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, 19)
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), 19)
    err, _ = scoring.vec_err(Wr.double().cpu().numpy(), want)
    assert (err <= 2e-4 * np.abs(mag).max(axis=1) + 1e-6).all()


def test_quaternion_order_flag(oracle, dev):
    """Isaac core hands wxyz; the reference permutes to xyzw (hydrodynamics_behavior.py:194)."""
    wl = W.heterogeneous_boxes(50_000, seed=5)
    ref = _ref(oracle, wl)
    e = _engine(wl, torch.float32, dev, "tile")
    e.quat_order = "wxyz"
    e.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
    F, T = e.step(_t(wl.pos, torch.float32, dev), _t(wl.quat_xyzw[:, [3, 0, 1, 2]], torch.float32, dev),
                  _t(wl.lin_vel, torch.float32, dev), _t(wl.ang_vel, torch.float32, dev), wl.dt)
    _check(wl, torch.float32, ref, F.double().cpu().numpy(), T.double().cpu().numpy(), "wxyz")


def test_first_step_reset_and_dt_guard(oracle, dev):
    """v_prev starts at zero (hydrodynamics_behavior.py:196-198), reset() restores that,
    dt <= 1e-6 leaves outputs and state untouched (:139)."""
    wl = W.hexapod_envs(512)
    e = _engine(wl, torch.float64, dev, "direct")
    args = [_t(a, torch.float64, dev) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
    zero = np.zeros_like(wl.prev_lin)
    ref0 = oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, zero, zero, wl.dt)
    F0, _ = e.step(*args, wl.dt)
    F1, _ = e.step(*args, wl.dt)  # second step: acceleration is now exactly zero
    ref1 = oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.lin_vel,
                       wl.ang_vel, wl.dt)
    scale = scoring.force_scale(wl.coeff_per_body(), wl.rho, wl.g)
    assert scoring.fp64_ok(F0.cpu().numpy(), ref0.force, scale).all()
    assert scoring.fp64_ok(F1.cpu().numpy(), ref1.force, scale).all()
    e.reset()
    F2, _ = e.step(*args, wl.dt)
    assert torch.equal(F0, F2)
    sentinel = torch.full_like(F0, 7.0)
    out = sentinel.clone()
    before = e.prev_velocities().clone()
    e.step(*args, 1e-7, out_force=out, out_torque=sentinel.clone())
    torch.cuda.synchronize()
    assert torch.equal(out, sentinel) and torch.equal(before, e.prev_velocities())


# --------------------------------------------------------------------------- components vs golden
@pytest.mark.parametrize("group", ["c3", "c3f64", "c2", "near", "edge"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_components_against_reference_golden(golden, dev, group, dtype):
    """h2o_components vs outputs of the UNMODIFIED reference Numba code (tests/golden)."""
    from silver2_isaacsim_b200 import HydroEngine

    d = golden[group]
    if dtype == torch.float32 and group in ("c3f64", "near"):
        pytest.skip("inputs are not fp32-representable")
    n = len(d["pos"])
    ctor = d["ctor"]
    coeff = np.concatenate([ctor[:, 0:7], ctor[:, 9:12], d["mass"][:, None]], axis=1)
    e = HydroEngine(n, dtype=dtype, device=dev, water_density=float(ctor[0, 7]), gravity=float(ctor[0, 8]))
    e.set_params_per_body(coeff)
    out = e.components(*[_t(d[k], dtype, dev) for k in ("pos", "quat", "v", "w", "a", "al")], return_flags=True)
    torch.cuda.synchronize()
    flags = out[9].cpu().numpy().astype(bool)
    assert (flags == d["raised"]).all()  # exactly the bodies for which the reference raises (A.8)
    ok = ~d["raised"]
    if group == "edge" and dtype == torch.float32:
        ok &= np.abs(d["pos"][:, 2] - np.float32(d["pos"][:, 2])) == 0  # fp32-representable heights only
        ok &= d["ctor"][:, 0] > 1e-3
    scale = scoring.force_scale(coeff, 1025.0, 9.81)
    for k, name in enumerate(NAMES + ("sub_ratio",)):
        x, y = out[k].double().cpu().numpy()[ok], d[name][ok]
        if dtype == torch.float32:
            good = scoring.fp32_ok(x, y, rel=2e-5 if "center" in name else 1e-5)
            assert good.mean() >= 0.999, (group, name, float(good.mean()))
        else:
            extra = np.abs(d["pos"][ok]).max(axis=1) if "center" in name else 0.0
            sc = scale[ok] if "center" not in name else np.ones(ok.sum())
            good = scoring.fp64_ok(x, y, sc, extra=extra, rel=2e-12)
            assert good.all(), (group, name, int((~good).sum()))


def test_wrapper_twins_keep_the_reference_contract(golden, dev):
    """WarpHydrodynamicsWrapper / NumbaHydrodynamicsWrapper: ctor signature + return contract."""
    from silver2_isaacsim_b200 import NumbaHydrodynamicsWrapper, WarpHydrodynamicsWrapper

    d = golden["edge"]
    readme = [1, 1, 1, 1.2, 0.8, 300, 150, 1025, 9.81, 0.05, 0.02, 1.0]
    w = NumbaHydrodynamicsWrapper(*readme)
    assert w.total_volume == 1.0 and w.water_density == 1025 and w.lift_coefficient == 1.0
    r = w.calculate_hydrodynamic_forces(d["pos"][0], d["quat"][0], d["v"][0], d["w"][0], d["a"][0], d["al"][0])
    assert len(r) == 9 and isinstance(r[8], float) and r[0].shape == (3,) and r[0].dtype == np.float64
    for k, name in enumerate(NAMES):
        np.testing.assert_allclose(r[k], d[name][0], rtol=1e-12, atol=1e-12)
    assert r[8] == pytest.approx(0.7, rel=1e-13)
    strict = NumbaHydrodynamicsWrapper(*readme, strict_reference_errors=True)
    with pytest.raises(TypeError):  # GV5: the reference raises for a wet body at rest
        strict.calculate_hydrodynamic_forces(d["pos"][4], d["quat"][4], d["v"][4], d["w"][4], d["a"][4], d["al"][4])
    g = WarpHydrodynamicsWrapper(width=1, depth=1, height=1, linear_drag_coefficient=1.2, angular_drag_coefficient=0.8,
                                 linear_damping=300, angular_damping=150, water_density=1025, gravity=9.81,
                                 linear_mass_coeff=0.05, angular_mass_coeff=0.02, lift_coefficient=1.0, device="cuda:0")
    t = lambda k: torch.as_tensor(d[k][:1], dtype=torch.float32, device=dev)
    out = g.calculate_hydrodynamic_forces(t("pos"), t("quat"), t("v"), t("w"), t("a"), t("al"))
    assert len(out) == 8 and all(o.shape == (1, 3) and o.is_cuda for o in out)
    np.testing.assert_allclose(out[0].cpu().numpy()[0], d["buoyancy_force"][0], rtol=1e-6)


# --------------------------------------------------------------------------- launch modes
def test_bound_step_and_graph_rollout_equal_eager(dev):
    """C5: a captured CUDA-graph rollout reproduces the same steps launched one by one."""
    wl = W.uniform_small_batch(1024)
    outs = []
    for mode in ("eager", "bound", "graph"):
        e = _engine(wl, torch.float32, dev)
        ten = [_t(a, torch.float32, dev) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
        if mode == "eager":
            for _ in range(4):
                F, T = e.step(*ten, wl.dt)
        else:
            F, T = e.bind(*ten)
            if mode == "bound":
                for _ in range(4):
                    e.step_bound(wl.dt)
            else:
                e.capture_rollout(4, wl.dt)  # capturing launches nothing and leaves the state alone
                assert e.launch_count == 0
                e.launch_rollout()
        torch.cuda.synchronize()
        outs.append((F.clone(), T.clone(), e.prev_velocities().clone(), e.launch_count))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    assert outs[0][3] == outs[1][3] == outs[2][3] == 4


def test_setters_drop_a_captured_rollout(dev):
    """A captured graph bakes in device pointers and constants; every setter that changes one of them must
    drop it, and launching then fails loudly instead of replaying over freed / stale memory."""
    from silver2_isaacsim_b200 import _lib as L

    wl = W.hexapod_envs(64)
    e = _engine(wl, torch.float32, dev)
    ten = [_t(a, torch.float32, dev) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
    e.bind(*ten, robot_wrench=True)
    table2 = np.asarray(wl.table, dtype=np.float64) * 1.5
    setters = [
        lambda: e.set_part_table(table2, wl.slot_type),
        lambda: e.set_params_per_body(wl.coeff_per_body()),
        lambda: e.set_globals(1000.0, 9.8),
        lambda: e.set_environment((0.1, 0.0, 0.0), 0.2),
        lambda: e.set_added_mass_dense(np.eye(6)),
        lambda: e.set_added_mass_dense(None),
        lambda: e.set_articulation(19),
        lambda: e.set_articulation_offsets([0, 19, wl.n]),
        lambda: e.set_kernel("direct"),
        lambda: e.set_tile_config(0),
        lambda: e.enable_stats(True),
        lambda: setattr(e, "quat_order", "wxyz"),
        lambda: e.set_rollout_mode(False),
        lambda: e.set_surface_heights(torch.zeros(wl.n, device=dev)),
    ]
    for k, setter in enumerate(setters):
        if k == 8:  # unequal robots changed the wrench shape: bind again with equal runs
            e.set_articulation(19)
            e.bind(*ten, robot_wrench=True)
        e.capture_rollout(2, wl.dt)
        e.launch_rollout()
        setter()
        with pytest.raises(L.H2OError, match="NOT_CONFIGURED"):
            e.launch_rollout()
    torch.cuda.synchronize()


def test_capture_leaves_state_alone(dev):
    """Capturing a free-body rollout launches nothing: pose, velocities and carried velocities are untouched."""
    wl = W.uniform_small_batch(256)
    e = _engine(wl, torch.float64, dev)
    ten = [_t(a, torch.float64, dev) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
    before = [t.clone() for t in ten]
    e.set_prev(_t(wl.prev_lin, torch.float64, dev), _t(wl.prev_ang, torch.float64, dev))
    prev0 = e.prev_velocities().clone()
    e.bind(*ten)
    e.set_rollout_mode(free_bodies=True, gravity=wl.g)
    e.capture_rollout(5, wl.dt)
    torch.cuda.synchronize()
    assert e.launch_count == 0
    assert all(torch.equal(a, b) for a, b in zip(ten, before)) and torch.equal(e.prev_velocities(), prev0)
    e.launch_rollout()
    torch.cuda.synchronize()
    assert e.launch_count == 10 and not torch.equal(ten[0], before[0])


def test_step_host_pipeline(oracle, dev):
    """Host-buffer entry point (NumPy in/out, pinned tensors in/out) == device entry point."""
    wl = W.sharded_robots(9001)
    e = _engine(wl, torch.float32, dev)
    e.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
    F, T, Wr = e.step_host(wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.dt, robot_wrench=True)
    e2 = _engine(wl, torch.float32, dev)
    F2, T2, W2 = _run_step(e2, wl, torch.float32, dev, "split", robot=True)
    # chunk tails go through the direct kernel: same arithmetic, possibly different FMA contraction
    for x, y in ((F, F2), (T, T2)):
        assert scoring.fp32_ok(x, y, rel=2e-6).mean() > 0.999 and scoring.fp32_ok(x, y, rel=1e-4).all()
    ref = _ref(oracle, wl)
    _check(wl, torch.float32, ref, F, T, "step_host")
    np.testing.assert_allclose(Wr, W2, rtol=1e-5, atol=1e-4)
    pin = [torch.as_tensor(a).pin_memory() for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
    oF, oT = torch.empty(wl.n, 3).pin_memory(), torch.empty(wl.n, 3).pin_memory()
    e3 = _engine(wl, torch.float32, dev)
    e3.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
    e3.step_host(*pin, wl.dt, out_force=oF, out_torque=oT)
    assert e.last_host_path == "staged" and e3.last_host_path == "zero-copy"  # NumPy arrays vs pinned tensors
    for x, y in ((oF.numpy(), F), (oT.numpy(), T)):  # different chunking (no robot tiles) -> rounding-level
        assert scoring.fp32_ok(x, y, rel=2e-6).mean() > 0.999 and scoring.fp32_ok(x, y, rel=1e-4).all()
    _check(wl, torch.float32, ref, oF.numpy().astype(np.float64), oT.numpy().astype(np.float64), "step_host zero-copy")
    # PhysX layout on the host (two input arrays), pinned and pageable, with the robot wrench
    for pinned in (True, False):
        e4 = _engine(wl, torch.float32, dev)
        e4.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
        tr, ve = wl.transforms(), wl.velocities()
        outs = {}
        if pinned:  # zero-copy needs every buffer page-locked, outputs included
            tr, ve = torch.as_tensor(tr).pin_memory(), torch.as_tensor(ve).pin_memory()
            outs = dict(out_force=torch.empty(wl.n, 3).pin_memory(), out_torque=torch.empty(wl.n, 3).pin_memory(),
                        out_robot_wrench=torch.empty(wl.n // 19, 6).pin_memory())
        F4, T4, W4 = e4.step_host_physx(tr, ve, wl.dt, robot_wrench=True, **outs)
        assert e4.last_host_path == ("zero-copy" if pinned else "staged")
        _check(wl, torch.float32, ref, np.asarray(F4, dtype=np.float64), np.asarray(T4, dtype=np.float64), "step_host_physx")
        np.testing.assert_allclose(np.asarray(W4), W2, rtol=1e-5, atol=1e-4)


def test_step_host_table_slots_and_height_field(oracle, dev):
    """Host pipeline chunking keeps the slot phase of a part table (n_slots does not divide the chunk
    size) and offsets a borrowed per-body height field per chunk."""
    from silver2_isaacsim_b200 import HydroEngine, params as P

    n, n_slots, n_types = 300_007, 7, 3
    wl = W.heterogeneous_boxes(n, seed=4242)
    rng = np.random.default_rng(7)
    table = np.asarray(wl.coeff, dtype=np.float64)[rng.choice(n, n_types, replace=False)]
    slots = np.array([0, 1, 2, 1, 0, 2, 2], dtype=np.int32)
    coeff = table[slots[np.arange(n) % n_slots]]
    eta = (0.2 * np.sin(0.5 * wl.pos[:, 0].astype(np.float64))).astype(np.float32)
    ctor, mass = P.coeff_to_ctor_rows(coeff, wl.rho, wl.g), coeff[:, 10].copy()
    for use_eta in (False, True):
        e = HydroEngine(n, device=dev)
        e.set_globals(wl.rho, wl.g)
        e.set_part_table(table, slots)
        e.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
        pos = wl.pos.astype(np.float64).copy()
        if use_eta:
            e.set_surface_heights(torch.as_tensor(eta, device=dev))
            pos[:, 2] -= eta.astype(np.float64)
        F, T = e.step_host(wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.dt)
        ref = oracle.step(ctor, mass, pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt)
        scoring.assert_fp32(F, ref.force, f"step_host table eta={use_eta} force")
        scoring.assert_fp32(T, ref.torque, f"step_host table eta={use_eta} torque")


def test_unaligned_views_are_refused(dev):
    """The boundary requires 16-byte aligned tensors (include/h2o.h): contiguous views with a 12-byte
    offset are refused with H2O_ERR_ALIGNMENT by the DLPack and the raw-pointer entry points alike."""
    import ctypes
    from silver2_isaacsim_b200 import _lib as L

    wl = W.heterogeneous_boxes(4096, seed=31)
    e = _engine(wl, torch.float32, dev)
    pad = lambda a: torch.cat([torch.zeros(1, a.shape[1], device=dev), _t(a, torch.float32, dev)])[1:]
    pos, quat, lin, ang = pad(wl.pos), _t(wl.quat_xyzw, torch.float32, dev), pad(wl.lin_vel), pad(wl.ang_vel)
    assert pos.is_contiguous() and pos.data_ptr() % 16 != 0
    with pytest.raises(L.H2OError, match="ALIGNMENT"):
        e.step(pos, quat, lin, ang, wl.dt)
    F, T = torch.empty(wl.n, 3, device=dev), torch.empty(wl.n, 3, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = e._lib.h2o_step(e._h, p(pos), p(quat), p(lin), p(ang), ctypes.c_double(wl.dt), p(F), p(T), None,
                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 7  # H2O_ERR_ALIGNMENT (include/h2o.h)


def test_graph_rollout_with_robot_wrench(oracle, dev):
    """A captured rollout over bound tensors carries the per-robot wrench output too."""
    wl = W.hexapod_envs(4096)
    e = _engine(wl, torch.float32, dev)
    ten = [_t(a, torch.float32, dev) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
    F, T, Wr = e.bind(*ten, robot_wrench=True)
    e.capture_rollout(4, wl.dt)
    Wr.fill_(777.0)
    e.launch_rollout()
    torch.cuda.synchronize()
    # static state: from the second step on v_prev == v, so the steady answer has zero acceleration
    ref = oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                      wl.lin_vel.copy(), wl.ang_vel.copy(), wl.dt)
    _check(wl, torch.float32, ref, F.double().cpu().numpy(), T.double().cpu().numpy(), "graph rollout robots")
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, 19)
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), 19)
    err, _ = scoring.vec_err(Wr.double().cpu().numpy(), want)
    assert (err <= 2e-4 * np.abs(mag).max(axis=1) + 1e-6).all()


def test_stats_vector(oracle, dev):
    wl = W.heterogeneous_boxes(70_001, seed=11)
    ref = _ref(oracle, wl)
    for kernel in ("tile", "direct"):
        e = _engine(wl, torch.float32, dev, kernel, stats=True)
        _run_step(e, wl, torch.float32, dev)
        s = e.stats(reset=True)
        assert s["bodies"] == wl.n and s["nonfinite_bodies"] == 0
        assert s["wet_bodies"] == int((ref.components["sub_ratio"] > 0).sum())
        assert abs(s["clamped_bodies"] - int(((ref.flags & 2) != 0).sum())) <= 1
        norms = np.linalg.norm(ref.force, axis=1)
        assert s["sum_force_norm"] == pytest.approx(norms.sum(), rel=1e-6)
        assert s["max_force_norm"] == pytest.approx(norms.max(), rel=1e-6)
        assert e.stats()["bodies"] == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_stats_vector_with_robot_tiles(oracle, dev, dtype):
    """Statistics through the whole-robot tile kernels (hexapod-specialised and run-time robot size),
    table parameters; robots whose tail goes through the per-body kernel are counted once."""
    for wl, bpr in ((W.hexapod_envs(4099), 19), (W.heterogeneous_boxes(7 * 9001, seed=12, xy_range=2.0), 7)):
        ref = _ref(oracle, wl)
        e = _engine(wl, dtype, dev, "tile", stats=True)
        e.set_articulation(bpr)
        _run_step(e, wl, dtype, dev, robot=True)
        assert e.last_kernel == "tile"
        s = e.stats(reset=True)
        assert s["bodies"] == wl.n and s["nonfinite_bodies"] == 0
        assert s["wet_bodies"] == int((ref.components["sub_ratio"] > 0).sum())
        norms = np.linalg.norm(ref.force, axis=1)
        assert s["sum_force_norm"] == pytest.approx(norms.sum(), rel=1e-6)
        assert s["max_force_norm"] == pytest.approx(norms.max(), rel=1e-6)


# --------------------------------------------------------------------------- error behaviour
def test_error_codes(dev):
    from silver2_isaacsim_b200 import H2OError, HydroEngine

    e = HydroEngine(64, dtype=torch.float32, device=dev)
    z3, z4 = torch.zeros(64, 3, device=dev), torch.zeros(64, 4, device=dev)
    with pytest.raises(H2OError, match="NOT_CONFIGURED"):
        e.step(z3, z4, z3, z3, 0.01)
    e.set_params_uniform(P.HydroParams().ctor_row(), 512.5)
    with pytest.raises(H2OError, match="BAD_SHAPE"):
        e.step(z3[:32], z4, z3, z3, 0.01)
    with pytest.raises(H2OError, match="BAD_DTYPE"):
        e.step(z3.double(), z4, z3, z3, 0.01)
    with pytest.raises(H2OError, match="BAD_DEVICE"):
        e.step(z3.cpu(), z4, z3, z3, 0.01)
    with pytest.raises(H2OError, match="NOT_CONTIGUOUS"):
        e.step(torch.zeros(3, 64, device=dev).t(), z4, z3, z3, 0.01)
    with pytest.raises(H2OError, match="NOT_CONFIGURED"):
        e.step(z3, z4, z3, z3, 0.01, robot_wrench=False, out_robot_wrench=torch.zeros(4, 6, device=dev))
    with pytest.raises(H2OError, match="BAD_SHAPE"):
        e.set_articulation(19)
    with pytest.raises(H2OError, match="NOT_CONFIGURED"):
        e.step_bound(0.01)
    F, T = e.step(z3, torch.tensor([[0, 0, 0, 1.0]], device=dev).repeat(64, 1), z3, z3, 0.01)
    assert torch.isfinite(F).all() and torch.isfinite(T).all()


# --------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_full_size_oracle_parity(oracle, dev, dtype):
    """BASELINE config 3 at full size: every one of the 2^20 bodies against the float64 oracle."""
    wl = W.heterogeneous_boxes(1 << 20)
    ref = _ref(oracle, wl)
    e = _engine(wl, dtype, dev, "tile")
    F, T = _run_step(e, wl, dtype, dev)
    assert e.last_kernel == "tile"
    _check(wl, dtype, ref, F, T, "C3 2^20")
    wet = ref.components["sub_ratio"] > 0
    partial = wet & (ref.components["sub_ratio"] < 1)
    assert 0.3 < wet.mean() < 0.95 and partial.mean() > 0.2      # the wet / partial / dry mix of SURVEY 8(d)
    assert (F[~wet] == 0).all() and (T[~wet] == 0).all()          # dry bodies: exact zeros


def test_full_size_properties(dev):
    """BASELINE size (2^20 bodies): size-independent properties of the fused step."""
    wl = W.heterogeneous_boxes(1 << 20)
    e = _engine(wl, torch.float32, dev, "tile")
    ten = [_t(a, torch.float32, dev) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
    pl, pa = _t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev)
    e.set_prev(pl, pa)
    F, T = e.step(*ten, wl.dt)
    F, T = F.clone(), T.clone()
    assert torch.isfinite(F).all() and torch.isfinite(T).all()
    # (1) the tile kernel and the direct kernel are the same function of the inputs
    e.set_kernel("direct")
    e.set_prev(pl, pa)
    Fd, Td = e.step(*ten, wl.dt)
    for a, b in ((F, Fd), (T, Td)):  # same arithmetic, FMA contraction may differ between kernels
        err, den = (a - b).abs().max(dim=1).values, b.abs().max(dim=1).values
        assert (err <= 2e-6 * den + 1e-6).float().mean() > 0.999 and (err <= 1e-4 * den + 1e-6).all()
    # (2) horizontal translation invariance (the x,y position never enters the body-relative arms)
    e.set_kernel("tile")
    shifted = ten[0].clone()
    shifted[:, :2] += torch.tensor([123.0, -77.0], device=dev)
    e.set_prev(pl, pa)
    Fs, Ts = e.step(shifted, *ten[1:], wl.dt)
    assert torch.equal(F, Fs) and torch.equal(T, Ts)
    # (3) bodies entirely above the surface feel nothing; the clamp bounds |F| by 500 m
    h = torch.as_tensor(wl.coeff[:, :3], device=dev)
    high = ten[0][:, 2] > 0.5 * torch.linalg.norm(h, dim=1) + 1e-3
    assert high.any() and (F[high] == 0).all() and (T[high] == 0).all()
    mass = torch.as_tensor(wl.coeff[:, 10], device=dev)
    assert (torch.linalg.norm(F, dim=1) <= 500.0 * mass * (1 + 1e-5) + 1e-5).all()
    # (4) carried state: v_prev <- v, and with v_prev == v the added-mass terms vanish, so a
    #     second identical step equals a step whose previous velocity was set explicitly
    prev = e.prev_velocities()
    assert torch.equal(prev[:, :3], ten[2]) and torch.equal(prev[:, 3:], ten[3])
    F2, T2 = e.step(*ten, wl.dt)
    e.set_prev(ten[2], ten[3])
    F3, T3 = e.step(*ten, wl.dt)
    assert torch.equal(F2, F3) and torch.equal(T2, T3)


def test_degenerate_inputs_stay_finite(dev):
    """Zero quaternion, zero velocity, zero-size box, huge depth, NaN state: no crash, NaN is counted."""
    from silver2_isaacsim_b200 import H2OError, HydroEngine

    with pytest.raises(H2OError, match="BAD_ARGUMENT"):
        HydroEngine(0, device=dev)  # empty engine is rejected at the boundary
    n = 8
    e = HydroEngine(n, device=dev)
    coeff = np.tile(np.array(P.HydroParams().coeff_record(512.5)), (n, 1))
    coeff[1, :3] = 0.0            # zero-size box
    coeff[2, 10] = 0.0            # zero mass -> clamp forces everything to ~0
    e.set_params_per_body(coeff)
    e.enable_stats(True)
    pos = torch.zeros(n, 3, device=dev); pos[:, 2] = -0.2; pos[3, 2] = -1e6
    quat = torch.zeros(n, 4, device=dev); quat[:, 3] = 1.0; quat[4] = 0.0   # body 4: zero quaternion
    v = torch.zeros(n, 3, device=dev); v[5] = torch.tensor([1e-7, 0, 0])   # body 5: below the speed threshold
    w = torch.zeros(n, 3, device=dev)
    pos[6, 0] = float("nan"); v[7, 2] = float("nan")
    F, T = e.step(pos, quat, v, w, 1 / 60)
    assert torch.isfinite(F[:6]).all() and torch.isfinite(T[:6]).all()
    assert torch.isfinite(F[6]).all()          # x,y position never enters the wrench
    assert not torch.isfinite(F[7]).all()      # a NaN velocity propagates ...
    s = e.stats()
    assert s["bodies"] == n and s["nonfinite_bodies"] == 1   # ... and is counted
    assert s["still_wet_bodies"] >= 5          # wet and at rest: where the reference raises TypeError
    assert (F[2].abs() <= 1e-5).all()          # zero mass: clamp scale = 0 * 500 / |F|


@pytest.mark.parametrize("group", ["c2", "c3", "edge", "batched"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_warp_compat_against_the_reference_warp_kernel(golden_warp, dev, group, dtype):
    """f3 / a11: h2o_components in Warp-compat mode against vectors produced by the reference's own Warp kernel
    source (tests/golden/reference_warp_golden.npz, oracle/make_golden_warp.py) -- every group incl. the edge
    cases (un-normalised quaternions, faces exactly on the waterline), flags == the bodies for which that source
    reads an unassigned variable."""
    from silver2_isaacsim_b200 import HydroEngine
    from tests.test_oracle_golden import warp_tolerance

    d = golden_warp[group]
    n = len(d["pos"])
    ctor = np.broadcast_to(d["ctor"], (n, 12)) if d["ctor"].ndim == 1 else d["ctor"]
    coeff = np.concatenate([ctor[:, 0:7], ctor[:, 9:12], np.ones((n, 1))], axis=1)
    e = HydroEngine(n, dtype=dtype, device=dev, water_density=float(ctor[0, 7]), gravity=float(ctor[0, 8]))
    e.set_params_per_body(coeff)
    e.set_warp_compat(True)
    out = e.components(*[_t(d[k], dtype, dev) for k in ("pos", "quat", "v", "w", "a", "al")], return_flags=True)
    torch.cuda.synchronize()
    assert (out[9].cpu().numpy().astype(bool) == d["raised"]).all()
    ok = ~d["raised"]
    for k, name in enumerate(NAMES):
        x, y = out[k].double().cpu().numpy()[ok], d[name][ok].astype(np.float64)
        err = np.abs(x - y).max(axis=1)
        assert (err <= warp_tolerance(name, y)).all(), (group, name, float((err / warp_tolerance(name, y)).max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_warp_compat_fused_step(oracle, dev, dtype):
    """The fused step in Warp-compat mode (what the production behaviour script computes: Warp wrapper + the
    behaviour tail) against the oracle with its Warp switch -- itself pinned to the reference's Warp kernel source
    (tests/test_oracle_golden.py) -- incl. the per-robot wrench; switching back restores the Numba semantics."""
    wl = W.sharded_robots(1500)
    q = wl.quat_xyzw.astype(np.float64).copy()
    q[::7] *= 1.0 + 2e-3                      # some un-normalised quaternions: quat_rotate != R(q) there
    wl.quat_xyzw = q.astype(np.float32)
    try:
        oracle.set_warp_compat(True)
        ref = _ref(oracle, wl)
    finally:
        oracle.set_warp_compat(False)
    num = _ref(oracle, wl)
    assert (np.abs(ref.force - num.force).max(axis=1) > 1e-3 * np.abs(num.force).max(axis=1)).mean() > 0.03
    e = _engine(wl, dtype, dev)
    e.set_warp_compat(True)
    F, T, Wr = _run_step(e, wl, dtype, dev, "split", robot=True)
    assert e.last_kernel == "direct"
    _check(wl, dtype, ref, F, T, "warp-compat step")
    want = oracle.robot_wrench(wl.pos, ref.force, ref.torque, wl.bodies_per_robot)
    mag = oracle.robot_wrench(wl.pos, np.abs(ref.force), np.abs(ref.torque), wl.bodies_per_robot)
    err, _ = scoring.vec_err(Wr, want)
    assert (err <= (1e-5 if dtype == torch.float32 else 1e-11) * np.abs(mag).max(axis=1) * 20 + 1e-6).all()
    e.set_warp_compat(False)
    F, T, _ = _run_step(e, wl, dtype, dev, "split", robot=True)
    _check(wl, dtype, num, F, T, "back to Numba semantics")


def test_warp_compat_components(golden, oracle, dev):
    """f3: the components entry point can reproduce the Warp twin's deviations (forward rotation of
    the accelerations, cob = cop = p when dry); the default stays Numba semantics."""
    from silver2_isaacsim_b200 import WarpHydrodynamicsWrapper

    d = golden["c2"]
    n = 512
    ctor = d["ctor"][0]
    sel = np.arange(n)
    state = {k: d[k][sel] for k in ("pos", "quat", "v", "w", "a", "al")}
    state["pos"][:8, 2] = 5.0  # a few dry bodies
    try:
        oracle.set_warp_compat(True)
        ref = oracle.components(ctor, *[state[k] for k in ("pos", "quat", "v", "w", "a", "al")])
    finally:
        oracle.set_warp_compat(False)
    num = oracle.components(ctor, *[state[k] for k in ("pos", "quat", "v", "w", "a", "al")])
    assert np.abs(ref["added_mass_force"] - num["added_mass_force"]).max() > 1e-3   # the modes do differ
    w = WarpHydrodynamicsWrapper(*ctor, device="cuda:0", warp_compat=True)
    t = lambda k: torch.as_tensor(state[k], dtype=torch.float32, device=dev)
    out = w.calculate_hydrodynamic_forces(t("pos"), t("quat"), t("v"), t("w"), t("a"), t("al"))
    for k, name in enumerate(NAMES):
        assert scoring.fp32_ok(out[k].cpu().numpy(), ref[name], rel=2e-5).all(), name
    assert np.allclose(out[6].cpu().numpy()[:8], state["pos"][:8], rtol=1e-6)       # dry: cob = p
    w0 = WarpHydrodynamicsWrapper(*ctor, device="cuda:0")                            # default: Numba semantics
    out0 = w0.calculate_hydrodynamic_forces(t("pos"), t("quat"), t("v"), t("w"), t("a"), t("al"))
    assert scoring.fp32_ok(out0[4].cpu().numpy(), num["added_mass_force"], rel=2e-5).all()
    assert (out0[6].cpu().numpy()[:8] == 0).all()


@pytest.mark.parametrize("n_robots", [1, 7, 8, 9, 2000, 2003])
def test_no_out_of_bounds_writes(oracle, dev, n_robots):
    """Outputs live inside larger buffers with sentinel guards on both sides; tile tails, partial
    robots-per-tile and the direct-kernel remainder must not write a byte outside their rows
    (compute-sanitizer is not available on this pool, so the guards are the bounds check)."""
    wl = W.hexapod_envs(n_robots, seed=123 + n_robots)
    n, R, G = wl.n, n_robots, 64
    ref = _ref(oracle, wl)
    for kernel in ("tile", "direct"):
        e = _engine(wl, torch.float32, dev, kernel)
        e.set_prev(_t(wl.prev_lin, torch.float32, dev), _t(wl.prev_ang, torch.float32, dev))
        bufF = torch.full((n + 2 * G, 3), 777.0, device=dev)
        bufT = torch.full((n + 2 * G, 3), 777.0, device=dev)
        bufW = torch.full((R + 2 * G, 6), 777.0, device=dev)
        F, T, Wr = bufF[G:G + n], bufT[G:G + n], bufW[G:G + R]
        e.step(_t(wl.pos, torch.float32, dev), _t(wl.quat_xyzw, torch.float32, dev), _t(wl.lin_vel, torch.float32, dev),
               _t(wl.ang_vel, torch.float32, dev), wl.dt, out_force=F, out_torque=T, out_robot_wrench=Wr)
        torch.cuda.synchronize()
        for buf, rows in ((bufF, n), (bufT, n), (bufW, R)):
            assert (buf[:G] == 777.0).all() and (buf[G + rows:] == 777.0).all(), (kernel, n_robots)
        assert not (F == 777.0).all(dim=1).any() and not (Wr == 777.0).all(dim=1).any()  # every row written
        scoring.assert_fp32(F.cpu().numpy(), ref.force, "guarded force")
        scoring.assert_fp32(T.cpu().numpy(), ref.torque, "guarded torque")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_environment_current_and_surface_height(oracle, dev, dtype):
    """f4: a uniform current and a raised water surface equal the reference model evaluated with the
    flow-relative velocity and the surface-relative height (accelerations and v_prev unchanged)."""
    wl = W.heterogeneous_boxes(60_000, seed=99)
    cur, eta = np.array([0.4, -0.25, 0.1]), 0.35
    cur = cur.astype(np.float32).astype(np.float64)
    e = _engine(wl, dtype, dev, "tile")
    e.set_environment(current=cur, surface_z=eta)
    F, T = _run_step(e, wl, dtype, dev)
    pos = wl.pos.astype(np.float64) - np.array([0, 0, eta])
    if dtype == torch.float32:  # what the kernel sees: fp32 velocity difference, fp64 height difference
        vrel = (wl.lin_vel - cur.astype(np.float32)).astype(np.float64)
        prel = (wl.prev_lin.astype(np.float64) - wl.lin_vel.astype(np.float64)) + vrel
    else:
        vrel = wl.lin_vel.astype(np.float64) - cur
        prel = wl.prev_lin.astype(np.float64) - cur
    ref = oracle.step(wl.ctor_rows(), wl.masses(), pos, wl.quat_xyzw, vrel, wl.ang_vel, prel, wl.prev_ang, wl.dt)
    wl2 = W.Workload(**{**wl.__dict__, "pos": pos})
    _check(wl2, dtype, ref, F, T, "environment")
    prev = e.prev_velocities().double().cpu().numpy()
    assert (prev[:, :3] == wl.lin_vel.astype(np.float64)).all()   # the carried state stays the BODY velocity
    e.set_environment()                                           # back to still water at z = 0
    F0, T0 = _run_step(e, wl, dtype, dev)
    _check(wl, dtype, _ref(oracle, wl), F0, T0, "environment reset")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_surface_height_field(oracle, dev, dtype):
    """f4: a non-flat surface (elevation sampled at every body) equals the reference model evaluated at
    the surface-relative height p_z - eta_i; the array is borrowed, so in-place updates are seen."""
    wl = W.heterogeneous_boxes(200_000, seed=77)     # past the tile threshold: must route to the per-body kernel
    x, y = wl.pos[:, 0].astype(np.float64), wl.pos[:, 1].astype(np.float64)
    eta = (0.3 * np.cos(0.7 * x + 0.2) + 0.15 * np.sin(1.3 * y - 0.4)).astype(np.float32)
    e = _engine(wl, dtype, dev, "auto")
    eta_t = torch.as_tensor(eta, device=dev).to(dtype)
    e.set_surface_heights(eta_t)
    F, T = _run_step(e, wl, dtype, dev)
    assert e.last_kernel == "direct"
    pos = wl.pos.astype(np.float64).copy()
    pos[:, 2] -= eta.astype(np.float64)
    ref = oracle.step(wl.ctor_rows(), wl.masses(), pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel,
                      wl.prev_lin, wl.prev_ang, wl.dt)
    _check(W.Workload(**{**wl.__dict__, "pos": pos}), dtype, ref, F, T, "surface height field")
    eta_t.zero_()                                     # borrowed: the next step sees the flat surface again
    F0, T0 = _run_step(e, wl, dtype, dev)
    _check(wl, dtype, _ref(oracle, wl), F0, T0, "surface height field zeroed")
    with pytest.raises(ValueError):
        e.set_surface_heights(eta_t[:-1])
    e.set_surface_heights(None)
    e.set_kernel("tile")
    _run_step(e, wl, dtype, dev)
    assert e.last_kernel == "tile"


def test_integration_md_raw_ctypes_snippet(dev):
    """The raw C-ABI binding shown in INTEGRATION.md section 3 (no helper layer) works as written."""
    import ctypes
    from silver2_isaacsim_b200 import _lib

    n = 4096
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.h2o_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    assert lib.h2o_create(ctypes.byref(h), ctypes.c_int64(n), 0, 0) == 0
    ctor = (ctypes.c_double * 12)(1, 1, 1, 1.2, 0.8, 300, 150, 1025, 9.81, 0.05, 0.02, 1.0)
    assert lib.h2o_set_params_uniform(h, ctor, ctypes.c_double(512.5)) == 0
    pos = torch.zeros(n, 3, device=dev); pos[:, 2] = -0.2
    quat_xyzw = torch.zeros(n, 4, device=dev); quat_xyzw[:, 3] = 1
    lin_vel = torch.zeros(n, 3, device=dev); lin_vel[:, 0] = 0.1; lin_vel[:, 2] = -0.3
    ang_vel = torch.zeros(n, 3, device=dev)
    out_force, out_torque = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.h2o_step(h, p(pos), p(quat_xyzw), p(lin_vel), p(ang_vel), ctypes.c_double(1 / 60),
                      p(out_force), p(out_torque), None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.h2o_last_error()
    torch.cuda.synchronize()
    # SURVEY.md Appendix B GV1 buoyancy + drag (first step: v_prev = 0 -> added mass acts as well)
    assert abs(float(out_force[0, 2]) - (7038.675 + 75.915 + 2.1525 + 0.7 * 51.25 * 0.3 * 60)) < 0.05
    assert lib.h2o_destroy(h) == 0
