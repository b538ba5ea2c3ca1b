"""Compare the outputs of tile-kernel tuning variants with the default configuration (same inputs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W

cfgs = [int(c) for c in os.environ.get("CFGS", "1,4,7").split(",")]
dev = torch.device("cuda:0")
ok = True
for n in (1 << 20, (1 << 18) + 77, 148 * 256 + 5):
    wl = W.heterogeneous_boxes(n, seed=7)
    t = lambda a: torch.as_tensor(a, device=dev)
    ref = None
    for cfg in [0] + cfgs:
        e = HydroEngine(n, device=dev); e.set_workload_params(wl); e.set_kernel("tile"); e.set_tile_config(cfg)
        e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
        F, T = e.step(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), wl.dt)
        pv = e.prev_velocities().clone()
        out = [x.double().cpu().numpy() for x in (F, T, pv)]
        if ref is None:
            ref = out; continue
        for name, a, b in zip("F T prev".split(), ref, out):
            err = np.abs(a - b) / (1e-6 + 1e-5 * np.abs(a))
            bad = float((err > 2).mean())
            print(f"n={n} cfg {cfg} {name}: max scaled diff {err.max():.3g}, frac>2 {bad:.2e}")
            if name.startswith("prev") and err.max() != 0: ok = False
            if bad > 1e-3 or err.max() > 100: ok = False
print("CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
