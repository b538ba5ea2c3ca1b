"""Golden vectors for the dense 6x6 added-mass generalisation.

TEST INFRASTRUCTURE (build container only: imports the unmodified reference).  Every case runs the
reference's ``solve_hydrodynamics`` through its own wrapper, with the wrapper's ``_added_mass_matrix``
replaced by a full matrix -- the argument ``calculate_added_mass`` (numba_hydrodynamics.py:219-253)
already accepts.  Output: tests/golden/reference_numba_dense_am.npz

    python -m oracle.make_golden_dense
"""
import os

import numpy as np

from silver2_isaacsim_b200 import workloads as W

from . import ref_numba
from .hydro_oracle import COMPONENT_NAMES

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_numba_dense_am.npz")
SLOTS = np.array([0, 1, 2, 3, 1, 0, 2], dtype=np.int32)


def dense_matrices(seed=W.SEED_BASE + 403):
    """Four body-frame 6x6 matrices, fp32-representable: symmetric with surge-pitch / sway-roll
    coupling, a full symmetric positive-definite one, a NON-symmetric one (pins the row/column
    convention), and a plain diagonal (must reproduce the wrapper's structure)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((4, 6, 6))
    d = np.array([6.0, 9.0, 14.0, 0.4, 0.7, 0.5])
    m[0] = np.diag(d)
    m[0][0, 4] = m[0][4, 0] = 0.8
    m[0][1, 3] = m[0][3, 1] = -0.6
    a = rng.normal(size=(6, 6))
    m[1] = a @ a.T * 0.5 + np.diag(d)
    m[2] = np.diag(d) + rng.normal(size=(6, 6)) * 0.7
    m[3] = np.diag(d * 1.5)
    return m.astype(np.float32).astype(np.float64)


def main():
    wl = W.heterogeneous_boxes(768, seed=W.SEED_BASE + 404)
    a = (wl.lin_vel.astype(np.float64) - wl.prev_lin.astype(np.float64)) / wl.dt
    al = (wl.ang_vel.astype(np.float64) - wl.prev_ang.astype(np.float64)) / wl.dt
    M = dense_matrices()
    d = dict(ctor=wl.ctor_rows(), mass=wl.masses(), pos=wl.pos.astype(np.float64),
             quat=wl.quat_xyzw.astype(np.float64), v=wl.lin_vel.astype(np.float64),
             w=wl.ang_vel.astype(np.float64), prev_lin=wl.prev_lin.astype(np.float64),
             prev_ang=wl.prev_ang.astype(np.float64), a=a, al=al, dt=np.float64(wl.dt),
             coeff=wl.coeff_per_body().astype(np.float64), rho=np.float64(wl.rho), g=np.float64(wl.g),
             matrices=M, slot_type=SLOTS)
    out, raised = ref_numba.components_via_wrapper(d["ctor"], d["pos"], d["quat"], d["v"], d["w"], a, al,
                                                   dense=M, dense_slot=SLOTS)
    blob = dict(d)
    for name in COMPONENT_NAMES:
        blob[name] = out[name]
    blob["sub_ratio"] = out["sub_ratio"]
    blob["raised"] = raised
    wet = out["sub_ratio"] > 0
    print(f"dense: n={len(raised)} raised={int(raised.sum())} wet={int(wet.sum())} "
          f"|F_am| median {np.median(np.linalg.norm(out['added_mass_force'][wet], axis=1)):.3g}")
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
