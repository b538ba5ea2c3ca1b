"""Full-size parity report on the GPU: CUDA path (through the C ABI) vs the float64 oracle.

Scores every body of C2 (4096 hexapods), C3 (2^20 boxes), a C4 shard (2^16 robots) and C5 with the
criteria of SURVEY.md 8(d), and separately the classes the survey asks to single out:
threshold-adjacent bodies (a keypoint / face centre within 1e-5*dim of the surface, speed within 10 %
of 1e-6, ratio within 10 % of 1e-9 -- detected with the oracle) and bodies for which the
unmodified reference raises (wet, at rest).  Writes a markdown table to stdout.
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import hydro_oracle as O
from silver2_isaacsim_b200 import HydroEngine, workloads as W

dev = torch.device("cuda:0")


def threshold_adjacent(wl, ref):
    q = wl.quat_xyzw.astype(np.float64); x, y, z, w = q.T
    r2 = np.stack([x * (z + z) - w * (y + y), y * (z + z) + w * (x + x), 1 - (x * (x + x) + y * (y + y))], 1)
    h = wl.coeff_per_body()[:, :3] / 2
    pz = wl.pos[:, 2].astype(np.float64)
    near = np.zeros(wl.n, bool)
    band = 1e-5 * (2 * h).max(axis=1)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                zz = r2[:, 0] * i * h[:, 0] + r2[:, 1] * j * h[:, 1] + r2[:, 2] * k * h[:, 2] + pz
                near |= np.abs(zz) < band
    speed = np.linalg.norm(wl.lin_vel.astype(np.float64), axis=1)
    near |= np.abs(speed - 1e-6) < 1e-7
    ratio = ref.components["sub_ratio"]
    near |= (ratio > 0) & (np.abs(ratio - 1e-9) < 1e-10)
    return near


def score(x, y, rel, absol):
    err = np.abs(x - y).max(axis=1); den = np.abs(y).max(axis=1)
    return err <= np.maximum(rel * den, absol), err / np.maximum(np.maximum(rel * den, absol), 1e-300)


rows = []
for name, wl in (("C2 4096 hexapods", W.hexapod_envs(4096)), ("C3 2^20 boxes", W.heterogeneous_boxes(1 << 20)),
                 ("C4 shard 2^16 robots", W.sharded_robots(1 << 16)), ("C5 1024 uniform", W.uniform_small_batch(1024))):
    ref = O.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt)
    adj = threshold_adjacent(wl, ref); raises = (ref.flags & 1) != 0
    scale = wl.rho * wl.g * wl.coeff_per_body()[:, :3].prod(axis=1)
    pn = np.abs(wl.pos).max(axis=1).astype(float) * np.abs(ref.force).max(axis=1)
    for dtype, label in ((torch.float32, "fp32"), (torch.float64, "fp64")):
        npdt = np.float32 if dtype == torch.float32 else np.float64
        e = HydroEngine(wl.n, dtype=dtype, device=dev); e.set_workload_params(wl)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a.astype(npdt)), device=dev)
        e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
        F, T = e.step(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), wl.dt)
        F, T = F.double().cpu().numpy(), T.double().cpu().numpy()
        if dtype == torch.float32:
            okF, wF = score(F, ref.force, 1e-5, 1e-6); okT, wT = score(T, ref.torque, 1e-5, 1e-6)
            crit = "1e-5 rel / 1e-6 abs"
        else:
            okF, wF = score(F, ref.force, 1e-12, 1e-15 * scale)
            err = np.abs(T - ref.torque).max(axis=1); tol = np.maximum(1e-12 * (np.abs(ref.torque).max(axis=1) + pn), 1e-15 * scale)
            okT, wT = err <= tol, err / tol
            crit = "1e-12 rel (torque: + |p||F| term)"
        for cls, sel in (("all", np.ones(wl.n, bool)), ("threshold-adjacent", adj), ("reference raises", raises)):
            if sel.sum() == 0:
                rows.append((name, label, cls, 0, "-", "-", "-", "-", crit)); continue
            rows.append((name, label, cls, int(sel.sum()), f"{okF[sel].mean():.6f}", f"{wF[sel].max():.2f}",
                         f"{okT[sel].mean():.6f}", f"{wT[sel].max():.2f}", crit))
        del e
print("| workload | mode | class | bodies | force pass | worst (x tol) | torque pass | worst (x tol) | criterion |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(str(c) for c in r) + " |")
