"""Host-side cost per call of the public entry points (small batch, so the GPU is never the limit)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W
dev = torch.device("cuda:0"); wl = W.uniform_small_batch(1024)
e = HydroEngine(wl.n, device=dev); e.set_workload_params(wl)
t = lambda a: torch.as_tensor(a, device=dev)
pos, quat, v, w = t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel)
F, T = torch.empty_like(pos), torch.empty_like(pos)
def timeit(fn, n=3000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    el = time.perf_counter() - t0; torch.cuda.synchronize(); return el / n * 1e6
print("step() with caller-allocated outputs : %.1f us/call" % timeit(lambda: e.step(pos, quat, v, w, wl.dt, out_force=F, out_torque=T)))
print("step() allocating outputs            : %.1f us/call" % timeit(lambda: e.step(pos, quat, v, w, wl.dt)))
e.bind(pos, quat, v, w, out_force=F, out_torque=T)
print("step_bound()                         : %.1f us/call" % timeit(lambda: e.step_bound(wl.dt)))
