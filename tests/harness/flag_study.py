"""Conditioning-flag study on the CPU (host instantiation of the model header vs the float64 oracle).

    python tests/harness/flag_study.py [n_bodies] [seeds] [workloads]

For each workload: how many bodies the fp32 fast path flags for float64 re-evaluation, how many UNFLAGGED
bodies miss the fp32-mode bound (must be 0) and how close the worst unflagged one comes, then the same with
the fallback applied (what the kernels return).  This is how FLAG_KAPPA_T / FLAG_KAPPA_F / FLAG_AREA_COND
in csrc/h2o_model.cuh were chosen.
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import hydro_oracle as O  # noqa: E402
from silver2_isaacsim_b200 import workloads as W  # noqa: E402
from tests import emul, scoring  # noqa: E402
from tests.test_stress_distribution import stress_workload  # noqa: E402

MAKERS = {"C3": lambda n, s: W.heterogeneous_boxes(n, seed=s), "C2": lambda n, s: W.hexapod_envs(n // 19, seed=s),
          "C4": lambda n, s: W.sharded_robots(n // 19, seed=s), "C5": lambda n, s: W.uniform_small_batch(n, seed=s),
          "stress": lambda n, s: stress_workload(n, seed=s)}

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    seeds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else list(MAKERS)
    for name in names:
        tot = flagged = 0
        worst = {"F": 0.0, "T": 0.0}
        worst_fb = {"F": 0.0, "T": 0.0}
        bad = {"F": 0, "T": 0}
        for seed in seeds:
            wl = MAKERS[name](n, seed)
            ref = O.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin,
                         wl.prev_ang, wl.dt)
            emul.lib().emul_set_no_fallback(ctypes.c_int(1))
            F0, T0, comp, _ = emul.step(wl, emul.MODE_FP32_FAST)
            emul.lib().emul_set_no_fallback(ctypes.c_int(0))
            F1, T1, _, _ = emul.step(wl, emul.MODE_FP32_FAST)
            flag = comp[:, 27] == 2
            tot += wl.n
            flagged += int(flag.sum())
            for nm, x0, x1, y in (("F", F0, F1, ref.force), ("T", T0, T1, ref.torque)):
                err, den = scoring.vec_err(x0, y)
                tol = np.maximum(1e-5 * den, 1e-6)
                r = err / tol
                bad[nm] += int(((r > 1) & ~flag).sum())
                worst[nm] = max(worst[nm], float(r[~flag].max()))
                err1, _ = scoring.vec_err(x1, y)
                worst_fb[nm] = max(worst_fb[nm], float((err1 / tol).max()))
        print(f"{name:6s} {tot:9d} bodies: flagged {flagged} ({flagged / tot:.2e}); unflagged misses F {bad['F']} T {bad['T']}; "
              f"worst unflagged F {worst['F']:.2f}x T {worst['T']:.2f}x tol; with fallback worst F {worst_fb['F']:.2f}x "
              f"T {worst_fb['T']:.2f}x", flush=True)
