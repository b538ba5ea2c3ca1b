"""Minimal driver for ncu: a few fused steps of the C3 workload (2^20 bodies, fp32 or fp64)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W

n = int(os.environ.get("N", 1 << 20)); steps = int(os.environ.get("STEPS", 8))
dtype = torch.float64 if os.environ.get("DTYPE", "f32") == "f64" else torch.float32
kind = os.environ.get("WL", "c3")
dev = torch.device("cuda:0")
wl = W.heterogeneous_boxes(n) if kind == "c3" else (W.sharded_robots(n // 19) if kind == "c4" else W.hexapod_envs(n // 19))
eng = HydroEngine(wl.n, dtype=dtype, device=dev); eng.set_workload_params(wl)
eng.set_kernel(os.environ.get("KERNEL", "tile"))
npdt = np.float32 if dtype == torch.float32 else np.float64
t = lambda a: torch.as_tensor(np.ascontiguousarray(a.astype(npdt)), device=dev)
eng.set_prev(t(wl.prev_lin), t(wl.prev_ang))
eng.bind(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), robot_wrench=(kind != "c3"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(steps):
    flush.zero_()          # evict the inputs from L2 between steps
    eng.step_bound(wl.dt)
torch.cuda.synchronize()
print("done", eng.last_kernel, eng.launch_count)
