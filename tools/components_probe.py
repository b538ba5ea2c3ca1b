"""Throughput of the full-signature components entry point (h2o_components: the wrapper twins' path)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W
dev = torch.device("cuda:0")
for n in (1 << 20, 1 << 22):
    for dtype in (torch.float32, torch.float64):
        batches = []
        for b in range(3 if n == 1 << 20 else 2):
            wl = W.heterogeneous_boxes(n, seed=300 + b)
            e = HydroEngine(n, dtype=dtype, device=dev); e.set_workload_params(wl)
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dtype)
            a = (wl.lin_vel - wl.prev_lin) / wl.dt; al = (wl.ang_vel - wl.prev_ang) / wl.dt
            batches.append((e, [t(x) for x in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, a, al)]))
        for e, ten in batches: e.components(*ten)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 60
        ev0.record()
        for i in range(reps):
            e, ten = batches[i % len(batches)]; e.components(*ten)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / reps
        esz = 4 if dtype == torch.float32 else 8
        bytes_per_body = (19 + 11 + 25) * esz
        print(f"n={n} {str(dtype)[6:]}: {us:.1f} us/call  {n/us/1e3:.2f} G bodies/s  {n*bytes_per_body/us/1e3:.0f} GB/s ({bytes_per_body} B/body; includes 9 output allocations per call)", flush=True)
        del batches
