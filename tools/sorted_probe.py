"""Upper bound of an in-tile regime sort for C3: the SAME 2^20 bodies, reordered on the host so that inside every
128-body tile the bodies the surface cuts come first, then the fully submerged, then the dry ones (warps become
regime-uniform), run through a library built with -DH2O_SKIP_ALL=true (warp-uniform skip in the default kernel).
    H2O_LIB_PATH=.../libh2o_b200_skipall.so python tools/sorted_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W

dev = torch.device("cuda:0")
n = 1 << 20
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)


def regime(wl):
    q = wl.quat_xyzw.astype(np.float64); x, y, z, w = q.T
    r2 = np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=1)
    ext = (np.abs(r2) * 0.5 * wl.coeff[:, 0:3].astype(np.float64)).sum(axis=1)
    pz = wl.pos[:, 2].astype(np.float64)
    return np.where(pz - ext >= 0, 2, np.where(pz + ext < 0, 1, 0))   # 0 cut, 1 fully submerged, 2 dry


def sort_tiles(wl, tile):
    key = regime(wl) + 4 * (np.arange(wl.n) // tile)
    order = np.argsort(key, kind="stable")
    for name in ("pos", "quat_xyzw", "lin_vel", "ang_vel", "prev_lin", "prev_ang", "coeff"):
        setattr(wl, name, getattr(wl, name)[order])
    return wl


for label, prep in (("as drawn", lambda wl: wl), ("sorted per 128-body tile", lambda wl: sort_tiles(wl, 128)),
                    ("sorted per 1024 bodies", lambda wl: sort_tiles(wl, 1024))):
    es = []
    for b in range(6):
        wl = prep(W.heterogeneous_boxes(n, seed=100 + b))
        e = HydroEngine(n, device=dev); e.set_workload_params(wl); e.set_kernel("tile")
        e.set_prev(t(wl.prev_lin), t(wl.prev_ang)); e.bind(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel))
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
    ev0.record()
    for _ in range(1200): g.replay()
    ev1.record(); torch.cuda.synchronize()
    print(f"{label:26s}: burst {np.median(ts):6.2f} us/step   sustained {ev0.elapsed_time(ev1) * 1e3 / (1200 * 24):6.2f} us/step", flush=True)
    del es, g
