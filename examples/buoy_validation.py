"""The reference README's buoy validation, stand-alone on the GPU.

A 1 m cube of half the water's density is released from z = 1 m with a small tilt and left to bob
until it floats half submerged.  Forces come from the fused step kernel, the motion from the
device-side free-body stepper (semi-implicit Euler + gravity; inside Isaac Sim PhysX does this part),
both replayed from one captured CUDA graph of 100 steps.  After every replay one CSV row is written in the
column order of the reference's LogVelocity behaviour (log_velocity.py:17-20), and the real-time
factor is reported like its BenchmarkRtf (benchmark_rtf.py:37-71).

    python examples/buoy_validation.py [--steps 6000] [--n 1] [--csv /tmp/buoy.csv]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W      # noqa: E402
from silver2_isaacsim_b200.trace import RtfMeter, VelocityTrace    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6000)
ap.add_argument("--n", type=int, default=1, help="identical buoys stepped side by side")
ap.add_argument("--csv", default="/tmp/buoy_velocity_log.csv")
args = ap.parse_args()

dev = torch.device("cuda:0")
wl = W.readme_buoy()
eng = HydroEngine(args.n, dtype=torch.float64, device=dev)
eng.set_globals(wl.rho, wl.g)
eng.set_part_table(wl.table, wl.slot_type)
rep = lambda a: torch.as_tensor(np.repeat(np.asarray(a, dtype=np.float64), args.n, axis=0), device=dev).contiguous()
pos, quat, v, w = rep(wl.pos), rep(wl.quat_xyzw), rep(wl.lin_vel), rep(wl.ang_vel)
F, T = eng.bind(pos, quat, v, w)
eng.set_rollout_mode(free_bodies=True, gravity=wl.g)   # every captured step: forces, then the stepper

GRAPH = 100
eng.capture_rollout(GRAPH, wl.dt)
trace, rtf = VelocityTrace(args.csv), RtfMeter(report_every=1000)
for k in range(args.steps // GRAPH):
    eng.launch_rollout()
    torch.cuda.synchronize()
    rtf.on_physics_step(wl.dt, GRAPH)
    trace.sample(pos.cpu(), v.cpu(), w.cpu())           # one row per replayed graph
r = rtf.report()
z = float(pos[0, 2])
print(f"{args.n} buoy(s), {r['steps']} steps of {wl.dt * 1e3:.2f} ms: RTF {r['rtf']:.1f}x, {r['steps_per_s']:.0f} steps/s; "
      f"final height of the centre {z:+.4f} m (half submerged = 0), |v| {float(v[0].norm()):.2e} m/s; trace: {args.csv}")
