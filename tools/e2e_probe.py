"""PCIe probe: raw pinned H2D / D2H bandwidth and step_host time for several chunk counts."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W
dev = torch.device("cuda:0"); n = 1 << 20
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device=dev)
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); el = time.perf_counter() - t
    print(f"{name}: {20 * h.numel() / el / 1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(); h2 = torch.empty_like(h).pin_memory(); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); el = time.perf_counter() - t
print(f"H2D+D2H concurrent: {20 * h.numel() / el / 1e9:.1f} GB/s each direction")
wl = W.heterogeneous_boxes(n)
pin = [torch.as_tensor(a).pin_memory() for a in (wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel)]
oF, oT = torch.empty(n, 3).pin_memory(), torch.empty(n, 3).pin_memory()
for cps in os.environ.get("CPS", "3,4,5,6,8").split(","):
    os.environ["H2O_HOST_CHUNKS"] = cps
    e = HydroEngine(n, device=dev); e.set_workload_params(wl)
    for _ in range(3): e.step_host(*pin, wl.dt, out_force=oF, out_torque=oT)
    t = time.perf_counter()
    for _ in range(20): e.step_host(*pin, wl.dt, out_force=oF, out_torque=oT)
    el = (time.perf_counter() - t) / 20
    print(f"chunks {cps}: {el*1e3:.3f} ms/step  {n/el/1e9:.3f} G bodies/s  (H2D {n*52/el/1e9:.1f} GB/s)")
