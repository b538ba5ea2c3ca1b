"""Precision study on the CPU: the model header (host-instantiated, tests/host_emul) vs the f64 oracle.

    python tests/harness/precision_study.py [n_bodies]

Prints, per workload and precision policy, how many force / torque vectors miss the fp32
criterion of SURVEY.md 8(d) and the error quantiles.  This is how the mixed-precision policy of
DESIGN.md section 4 was chosen (fp32 storage; waterline, ratio, buoyancy and buoyancy arm in fp64).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import hydro_oracle as O  # noqa: E402
from silver2_isaacsim_b200 import workloads as W  # noqa: E402
from tests import emul, scoring  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    for name, wl in (("C3", W.heterogeneous_boxes(n)), ("C2", W.hexapod_envs(n // 19)),
                     ("C4", W.sharded_robots(n // 38)), ("C5", W.uniform_small_batch(min(n, 100000)))):
        ref = O.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin,
                     wl.prev_ang, wl.dt)
        for mode, label in ((emul.MODE_FP32_FAST, "fp32 fused step (kernel path)"), (emul.MODE_FP32, "fp32 generic"),
                            (emul.MODE_ALL_FP32, "all-fp32 arithmetic"), (emul.MODE_FP32_STORE_FP64_MATH, "fp32 storage, fp64 math")):
            F, T, _, _ = emul.step(wl, mode)
            for nm, x, y in (("F", F, ref.force), ("T", T, ref.torque)):
                err, den = scoring.vec_err(x, y)
                tol = np.maximum(1e-5 * den, 1e-6)
                rel = err[den > 0] / den[den > 0]
                print(f"{name} {label:30s} {nm}: miss {int((err > tol).sum()):5d}/{wl.n}  median {np.median(rel):.2e} "
                      f"p99.9 {np.quantile(rel, 0.999):.2e}  worst {float((err / tol).max()):.2f}x tol")
