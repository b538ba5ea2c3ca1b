"""Fused step (C3, fp32, per-body records, 2^20 bodies) for the three ingest layouts: split (pos, quat, v, w),
RigidPrimView (pos, quat, velocities (N,6)) and PhysX tensor API (transforms (N,7), velocities (N,6)).
Graph replay over 6 resident batches (L2-cold), median of 31 replays of 24 steps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W

dev = torch.device("cuda:0")
n = int(os.environ.get("N", 1 << 20))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
for layout in ("split", "view", "physx"):
    es = []
    for b in range(6):
        wl = W.heterogeneous_boxes(n, seed=100 + b)
        e = HydroEngine(n, device=dev); e.set_workload_params(wl); e.set_kernel("tile")
        e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
        if layout == "split":
            e.bind(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel))
        elif layout == "view":
            e.bind(t(wl.pos), t(wl.quat_xyzw), velocities=t(wl.velocities()))
        else:
            e.bind(transforms=t(wl.transforms()), velocities=t(wl.velocities()))
        es.append((e, wl))
    dt = es[0][1].dt
    for e, _ in es: e.step_bound(dt)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(24): es[i % 6][0].step_bound(dt)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(31):
        ev0.record(); g.replay(); ev1.record(); torch.cuda.synchronize(); ts.append(ev0.elapsed_time(ev1) * 1e3 / 24)
    us = float(np.median(ts))
    print(f"{layout:6s}: {us:6.2f} us/step  {168 * n / us / 1e3:7.1f} GB/s  kernel {es[0][0].last_kernel}", flush=True)
    del es, g
