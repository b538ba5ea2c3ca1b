"""BASELINE config 1 (single README buoy, 10 000 steps) and the behaviour adapter, on the GPU."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from silver2_isaacsim_b200 import params as P
from silver2_isaacsim_b200 import workloads as W
from tests import scoring

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def buoy_record(oracle):
    from oracle import free_body

    wl = W.readme_buoy()
    ctor = P.HydroParams().ctor_row()
    mass = float(wl.table[0, 10])
    rec = free_body.rollout(ctor, mass, wl.pos[0], wl.quat_xyzw[0], wl.lin_vel[0], wl.ang_vel[0], wl.dt, 10000)
    return wl, ctor, mass, rec


def test_c1_buoy_settles(buoy_record):
    """Config 1 feasibility: the 1 m^3 half-density buoy dropped from 1 m settles at its waterline."""
    wl, ctor, mass, rec = buoy_record
    assert not rec["raised"].any()  # started dry -> never at rest while wet
    z = rec["pos"][:, 2]
    assert z[0] == 1.0 and abs(z[-1]) < 5e-3 and np.abs(rec["v"][-1]).max() < 2e-2
    assert np.abs(rec["F"]).max() <= 500 * mass * (1 + 1e-12)  # clamp bound


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
def test_c1_teacher_forced_10k_steps(buoy_record, dev, dtype):
    """Replay the oracle's 10 000-state sequence through the CUDA step (one body per time step,
    previous velocities = the behaviour's carried state) and compare F, tau step by step."""
    from silver2_isaacsim_b200 import HydroEngine

    wl, ctor, mass, rec = buoy_record
    n = len(rec["pos"])
    e = HydroEngine(n, dtype=dtype, device=dev)
    e.set_params_uniform(ctor, mass)
    if dtype == torch.float32:
        # feed fp32-representable states to both sides
        cast = lambda a: a.astype(np.float32).astype(np.float64)
        from oracle import hydro_oracle as O
        ref = O.step(np.array(ctor, float), [mass], cast(rec["pos"]), cast(rec["quat"]), cast(rec["v"]), cast(rec["w"]),
                     cast(rec["prev_v"]), cast(rec["prev_w"]), wl.dt)
        Fr, Tr = ref.force, ref.torque
    else:
        cast = lambda a: a
        Fr, Tr = rec["F"], rec["T"]
    t = lambda a: torch.as_tensor(cast(a), device=dev).to(dtype).contiguous()
    e.set_prev(t(rec["prev_v"]), t(rec["prev_w"]))
    F, T = e.step(t(rec["pos"]), t(rec["quat"]), t(rec["v"]), t(rec["w"]), wl.dt)
    F, T = F.double().cpu().numpy(), T.double().cpu().numpy()
    if dtype == torch.float32:
        scoring.assert_fp32(F, Fr, "C1 force")
        scoring.assert_fp32(T, Tr, "C1 torque")
    else:
        scale = np.full(n, 1025.0 * 9.81)
        assert scoring.fp64_ok(F, Fr, scale).all()
        assert scoring.fp64_ok(T, Tr, scale).all()  # |p_xy| ~ 0: the strict torque criterion applies


def test_c1_free_running_graph_rollout(buoy_record, dev):
    """10 000 steps on the device (fused step + free-body stepper, 100 replays of a captured
    100-step CUDA graph) against the float64 NumPy rollout.

    The model is discontinuous (a keypoint crossing the surface moves the centre of buoyancy by a
    finite amount), so the rollout is chaotic: perturbing the NumPy rollout's initial height by
    1e-15 moves it by 5e-9 after 500 steps and 3e-4 after 1000.  Hence: tight agreement while
    rounding noise has not been amplified yet, same settled state at the end, drift reported."""
    from silver2_isaacsim_b200 import HydroEngine

    wl, ctor, mass, rec = buoy_record
    e = HydroEngine(1, dtype=torch.float64, device=dev)
    e.set_params_uniform(ctor, mass)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev).contiguous()
    pos, quat, v, w = t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel)
    F, T = e.bind(pos, quat, v, w)
    e.set_rollout_mode(free_bodies=True, gravity=wl.g)
    e.capture_rollout(100, wl.dt)       # capturing does not advance the state
    drift = {}
    for k in range(100):
        e.launch_rollout()
        n_done = 100 * (k + 1)
        if n_done in (100, 200, 500, 1000, 5000):
            torch.cuda.synchronize()
            drift[n_done] = float(np.abs(pos.cpu().numpy()[0] - rec["pos"][n_done]).max())
    torch.cuda.synchronize()
    pf, qf, vf, wf = rec["final"]
    drift[10000] = float(np.abs(pos.cpu().numpy()[0] - pf).max())
    print("free-running drift |p_gpu - p_numpy|:", {k: "%.2e" % d for k, d in drift.items()})
    assert e.launch_count == 20000      # (fused step + stepper) x 10 000, all of them graph nodes
    assert drift[100] < 1e-11 and drift[200] < 1e-9, drift
    assert max(drift.values()) < 2e-2, drift
    assert np.abs(v.cpu().numpy()[0] - vf).max() < 2e-2
    assert abs(pos.cpu().numpy()[0][2] - pf[2]) < 5e-3 and abs(pf[2]) < 5e-3  # same waterline


def test_c1_persistent_rollout_kernel(buoy_record, dev):
    """The same 10 000 free-running steps in ONE kernel launch (state in registers across the steps),
    validated exactly like the graph rollout above, plus the trace the kernel writes every 100 steps."""
    from silver2_isaacsim_b200 import HydroEngine

    wl, ctor, mass, rec = buoy_record
    e = HydroEngine(1, dtype=torch.float64, device=dev)
    e.set_params_uniform(ctor, mass)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev).contiguous()
    pos, quat, v, w = t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel)
    F, T = e.bind(pos, quat, v, w)
    trace = e.rollout_persistent(10000, wl.dt, gravity=wl.g, trace_every=100)
    torch.cuda.synchronize()
    assert e.launch_count == 1 and tuple(trace.shape) == (100, 1, 9)
    tr = trace.cpu().numpy()[:, 0]
    drift = {n: float(np.abs(tr[n // 100 - 1, :3] - rec["pos"][n]).max()) for n in (100, 200, 500, 1000, 5000)}
    pf, qf, vf, wf = rec["final"]
    drift[10000] = float(np.abs(pos.cpu().numpy()[0] - pf).max())
    print("persistent rollout drift |p_gpu - p_numpy|:", {k: "%.2e" % d for k, d in drift.items()})
    assert drift[100] < 1e-11 and drift[200] < 1e-9, drift
    assert max(drift.values()) < 2e-2, drift
    assert np.abs(v.cpu().numpy()[0] - vf).max() < 2e-2
    assert abs(pos.cpu().numpy()[0][2] - pf[2]) < 5e-3 and abs(pf[2]) < 5e-3  # same waterline
    assert np.array_equal(tr[-1, :3], pos.cpu().numpy()[0]) and np.array_equal(tr[-1, 3:6], v.cpu().numpy()[0])
    # carried velocities = the velocities before the last stepper call; last wrench in the bound outputs
    assert torch.isfinite(F).all() and torch.isfinite(T).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("params", ["table", "per_body"])
def test_persistent_rollout_equals_stepwise(dev, dtype, params):
    """C5-style batch: K steps in one persistent launch == K x (fused step + free-body stepper) launched
    one by one (same arithmetic, same roundings of the carried state; only FMA contraction may differ)."""
    wl = W.uniform_small_batch(1024 if params == "table" else 3000)
    if params == "per_body":
        # per-body records: the README cube of every body with +-20 % jitter on every coefficient (the explicit
        # harness stepper is stable for these; light bodies with strong dampers would blow up)
        rng = np.random.default_rng(88)
        rec = np.repeat(np.asarray(wl.table, dtype=np.float64), wl.n, axis=0)
        wl.coeff = (rec * rng.uniform(0.8, 1.2, rec.shape)).astype(np.float32)
        wl.table = wl.slot_type = None
    K = 25
    res = []
    for mode in ("stepwise", "persistent"):
        from silver2_isaacsim_b200 import HydroEngine
        e = HydroEngine(wl.n, dtype=dtype, device=dev)
        e.set_workload_params(wl)
        e.enable_stats(True)
        tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dtype).contiguous()
        ten = [tt(a) for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
        e.set_prev(tt(wl.prev_lin), tt(wl.prev_ang))
        F, T = e.bind(*ten)
        if mode == "stepwise":
            for _ in range(K):
                e.step_bound(wl.dt)
                e.integrate_free_bodies(*ten, F, T, wl.dt, wl.g)
        else:
            e.rollout_persistent(K, wl.dt, gravity=wl.g)
        torch.cuda.synchronize()
        st = e.stats(reset=True)
        assert st["bodies"] == K * wl.n and st["nonfinite_bodies"] == 0
        res.append([x.double().cpu().numpy() for x in ten] + [F.double().cpu().numpy(), T.double().cpu().numpy(),
                                                              e.prev_velocities().double().cpu().numpy()])
        assert e.launch_count == (2 * K if mode == "stepwise" else 1)
    tol = 2e-4 if dtype == torch.float32 else 1e-9   # 25 steps of a discontinuous model amplify rounding
    names = ("pos", "quat", "lin_vel", "ang_vel", "force", "torque", "prev")
    for name, a, b in zip(names, *res):
        scale = np.abs(a).max(axis=1, keepdims=True) + 1e-3
        frac = float((np.abs(a - b) <= tol * scale).all(axis=1).mean())
        assert frac >= 0.995, (name, frac)   # the rest: a keypoint crossed the surface one step apart


class FakeRigidPrimView:
    """Stand-in for omni.isaac.core.prims.RigidPrimView (the three methods the reference uses)."""

    def __init__(self, pos, quat_wxyz, vel, masses):
        self.pos, self.quat, self.vel, self.masses = pos, quat_wxyz, vel, masses
        self.applied = None
        self.valid = True

    def is_valid(self):
        return self.valid

    def get_world_poses(self, clone=False):
        return self.pos, self.quat

    def get_velocities(self, clone=False):
        return self.vel

    def get_masses(self, clone=False):
        return self.masses

    def apply_forces_and_torques_at_pos(self, forces, torques, positions, is_global):
        assert is_global
        self.applied = (forces.clone(), torques.clone(), positions.clone())


def test_batched_behavior_adapter(oracle, dev):
    """f1: one behaviour for the 19 SILVER2 prims + the buoy, JSON name matching, wxyz input,
    two physics steps (v_prev = 0 then v_prev = v), reset on stop."""
    from silver2_isaacsim_b200.behavior import BatchedHydrodynamicsBehavior

    names = list(P.HEXAPOD_SLOTS) + ["Obsea_Buoy"]
    rng = np.random.default_rng(7)
    n = len(names)
    pos = np.concatenate([rng.uniform(-0.4, 0.4, (19, 3)) + [2, 10.7, -0.2], [[5.0, 5.0, -0.3]]]).astype(np.float32)
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True); q = q.astype(np.float32)
    vel = rng.normal(size=(n, 6)).astype(np.float32) * 0.3
    cfg = P.load_config()
    masses = np.array([cfg["masses"].get(P.match_part(nm, cfg) or "", 512.5) for nm in names], dtype=np.float32)
    tt = lambda a: torch.as_tensor(a, device=dev)
    view = FakeRigidPrimView(tt(pos), tt(q[:, [3, 0, 1, 2]].copy()), tt(vel), tt(masses))
    b = BatchedHydrodynamicsBehavior(names, view, device="cuda:0", robot_offsets=[0, 19, 20])  # robot + buoy
    assert b.part_of[0] == "body" and b.part_of[1] == "coxa" and b.part_of[-1] is None
    assert b.exposed["Obsea_Buoy"].xDimension == 1.0 and b.exposed["Tibia_5"].linearDamping == 20.0
    b._on_physics_step(1 / 60)          # not playing yet: nothing happens (engine is None)
    assert view.applied is None
    b.on_play()
    b._on_physics_step(1e-7)            # dt guard (:139)
    assert view.applied is None
    dt = 1 / 60
    ctor = np.array([b.exposed[nm].ctor_row() for nm in names])
    zero = np.zeros((n, 3))
    for prev_v, prev_w in ((zero, zero), (vel[:, :3], vel[:, 3:])):
        b._on_physics_step(dt)
        ref = oracle.step(ctor, masses, pos, q, vel[:, :3], vel[:, 3:], prev_v, prev_w, dt)
        F, T, Ppos = view.applied
        scoring.assert_fp32(F.cpu().numpy(), ref.force, "behaviour force")
        scoring.assert_fp32(T.cpu().numpy(), ref.torque, "behaviour torque")
        assert torch.equal(Ppos, view.pos)
        # per-robot net wrench: the 19-link robot about its Body prim, the buoy about itself
        wr = b.robot_wrench.double().cpu().numpy()
        assert wr.shape == (2, 6)
        want = oracle.robot_wrench(pos[:19], ref.force[:19], ref.torque[:19], 19)[0]
        assert np.abs(wr[0] - want).max() <= 2e-4 * (np.abs(ref.force[:19]).sum() + np.abs(ref.torque[:19]).sum())
        assert np.allclose(wr[1, :3], ref.force[19], rtol=2e-5, atol=1e-5)
        assert np.allclose(wr[1, 3:], ref.torque[19], rtol=2e-5, atol=1e-5)
    view.valid = False
    view.applied = None
    b._on_physics_step(dt)              # invalid view: skipped (:139)
    assert view.applied is None
    b.on_stop()
    assert b._engine is None
    view.valid = True
    b.on_play()                         # play again: carried velocities start from zero (:240-245)
    b._on_physics_step(dt)
    ref = oracle.step(ctor, masses, pos, q, vel[:, :3], vel[:, 3:], zero, zero, dt)
    scoring.assert_fp32(view.applied[0].cpu().numpy(), ref.force, "after reset")
