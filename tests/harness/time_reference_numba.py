"""Time the UNMODIFIED reference Numba path in the build container (needs /root/reference; it cannot
travel to the GPU box, which is why bench.py's CPU arm is the C port).  Prints body-updates/s for
(a) the wrapper called once per body from Python, as the reference is meant to be used, and
(b) the C oracle port on the same inputs, for scale."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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

# (c) SURVEY.md 8(d) "CPU baseline (2)": a harness-side @njit(parallel=True) prange driver calling the
# UNTOUCHED reference solve_hydrodynamics once per body (uniform README parameters so that one set of
# geometry arrays serves every body; moving bodies only -- a wet body at rest raises).
import numba
from numba import njit, prange
_, solve = ref_numba.load()
wref = Wrapper(1, 1, 1, 1.2, 0.8, 300, 150, 1025, 9.81, 0.05, 0.02, 1.0)
geom = (wref._local_keypoints, wref._local_face_centers, wref._face_areas, wref._local_face_normals, wref._added_mass_matrix)

@njit(parallel=True)
def drive(pos, quat, v, w, a, al, kp, fc, fa, fn, am, out):
    for i in prange(pos.shape[0]):
        r = solve(pos[i], quat[i], v[i], w[i], a[i], al[i], 1.0, 1025.0, 9.81, 1.2, 0.8, 300.0, 150.0, 1.0, kp, fc, fa, fn, am)
        out[i, 0:3] = r[0] + r[1] + r[2] + r[4]
        out[i, 3:6] = r[3] + r[5]

n = 200000
big = W.heterogeneous_boxes(n)
f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
args2 = [f8(big.pos), f8(big.quat_xyzw), f8(big.lin_vel), f8(big.ang_vel),
         f8((big.lin_vel.astype(np.float64) - big.prev_lin) / big.dt), f8((big.ang_vel.astype(np.float64) - big.prev_ang) / big.dt)]
out = np.zeros((n, 6))
drive(*[x[:64] for x in args2], *geom, out[:64])  # JIT
for thr in (1, numba.get_num_threads()):
    numba.set_num_threads(thr)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); drive(*args2, *geom, out); best = min(best, time.perf_counter() - t)
    print(f"reference solve_hydrodynamics in a prange driver, {thr} thread(s): {n / best:.3g} updates/s ({best / n * 1e6:.2f} us/body)")
