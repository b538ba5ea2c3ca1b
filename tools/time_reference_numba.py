"""Time the UNMODIFIED reference Numba path in the build container (needs /root/reference; it cannot
travel to the GPU box, which is why bench.py's CPU arm is the C port).  Prints body-updates/s for
(a) the wrapper called once per body from Python, as the reference is meant to be used, and
(b) the C oracle port on the same inputs, for scale."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_numba, hydro_oracle as O
from silver2_isaacsim_b200 import workloads as W

wl = W.heterogeneous_boxes(20000)
ctor = wl.ctor_rows()
a = (wl.lin_vel.astype(np.float64) - wl.prev_lin) / wl.dt
al = (wl.ang_vel.astype(np.float64) - wl.prev_ang) / wl.dt
Wrapper, _ = ref_numba.load()
wrappers = [Wrapper(*row) for row in ctor[:2000]]
args = [(wl.pos[i].astype(np.float64), wl.quat_xyzw[i].astype(np.float64), wl.lin_vel[i].astype(np.float64),
         wl.ang_vel[i].astype(np.float64), a[i], al[i]) for i in range(2000)]
wrappers[0].calculate_hydrodynamic_forces(*args[0])  # JIT
best = 1e9
for _ in range(3):
    t = time.perf_counter()
    for w, x in zip(wrappers, args):
        try:
            w.calculate_hydrodynamic_forces(*x)
        except TypeError:
            pass
    best = min(best, time.perf_counter() - t)
print(f"reference NumbaHydrodynamicsWrapper, one Python call per body: {best / 2000 * 1e6:.1f} us/body = {2000 / best:.3g} updates/s (1 core)")
O.components(ctor, wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, a, al)
for thr in (1, O.max_threads()):
    t = time.perf_counter()
    for _ in range(20):
        O.components(ctor, wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, a, al, n_threads=thr)
    el = time.perf_counter() - t
    print(f"C oracle port, {thr} thread(s): {20 * wl.n / el:.3g} updates/s")
