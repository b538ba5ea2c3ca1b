"""Time the fused step (C3, fp32) for every tile-kernel tuning variant; L2-cold via batch cycling."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W

n = int(os.environ.get("N", 1 << 20)); nb = 6; reps = int(os.environ.get("REPS", 3000))
dev = torch.device("cuda:0")
batches = []
for b in range(nb):
    wl = W.heterogeneous_boxes(n, seed=100 + b)
    e = HydroEngine(n, device=dev); e.set_workload_params(wl); e.set_kernel("tile")
    t = lambda a: torch.as_tensor(a, device=dev)
    e.set_prev(t(wl.prev_lin), t(wl.prev_ang)); e.bind(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel))
    batches.append((e, wl))
import pynvml as nv
nv.nvmlInit(); H = nv.nvmlDeviceGetHandleByIndex(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cfgs = [int(c) for c in os.environ.get("CFGS", "0").split(",")]
use_graph = int(os.environ.get("GRAPH", "0"))
for rnd in range(2):
    for cfg in cfgs:
        for e, _ in batches: e.set_tile_config(cfg)
        for i in range(30): batches[i % nb][0].step_bound(batches[0][1].dt)
        torch.cuda.synchronize()
        if use_graph:
            g = torch.cuda.CUDAGraph(); per = 10 * nb
            with torch.cuda.graph(g):
                for i in range(per): batches[i % nb][0].step_bound(batches[0][1].dt)
            g.replay(); torch.cuda.synchronize(); ev0.record()
            for _ in range(reps // per): g.replay()
            ev1.record(); reps_done = (reps // per) * per
        else:
            ev0.record()
            for i in range(reps): batches[i % nb][0].step_bound(batches[0][1].dt)
            ev1.record(); reps_done = reps
        clk = []
        while not ev1.query():
            clk.append(nv.nvmlDeviceGetClockInfo(H, nv.NVML_CLOCK_SM))
        pw = nv.nvmlDeviceGetPowerUsage(H) / 1000.0
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / reps_done
        print(f"round {rnd} cfg {cfg}: {us:7.2f} us/step  {168*n/us/1e3:7.1f} GB/s  {n/us/1e3:6.2f} G bodies/s  ctas/SM {batches[0][0].ctas_per_sm} clk {int(np.median(clk)) if clk else -1} MHz {pw:.0f} W", flush=True)
