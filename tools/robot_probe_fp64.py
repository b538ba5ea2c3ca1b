import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W
dev = torch.device("cuda:0")
wls = [W.sharded_robots(110592 // 2, seed=s, dtype=np.float64) for s in (1, 2)]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for robot in (False, True):
    es = []
    for wl in wls:
        e = HydroEngine(wl.n, dtype=torch.float64, device=dev); e.set_workload_params(wl); e.set_kernel("tile")
        t = lambda a: torch.as_tensor(a, device=dev)
        e.set_prev(t(wl.prev_lin), t(wl.prev_ang)); e.bind(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), robot_wrench=robot)
        es.append(e)
    for i in range(20): es[i % 2].step_bound(wls[0].dt)
    torch.cuda.synchronize(); ev0.record()
    for i in range(400): es[i % 2].step_bound(wls[0].dt)
    ev1.record(); torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / 400
    print(f"fp64 robot_wrench={robot}: {us:.1f} us/step {wls[0].n/us/1e3:.2f} G bodies/s {336*wls[0].n/us/1e3:.0f} GB/s ctas/SM {es[0].ctas_per_sm}")
    del es
