"""Generate tests/golden/reference_warp_golden.npz from the UNMODIFIED reference Warp twin.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_warp

For every case it calls the reference's own ``WarpHydrodynamicsWrapper.calculate_hydrodynamic_forces``
(/root/reference/src/scripts/physics/warp_hydrodynamics_wrapper.py:79) once per body -- exactly how
hydrodynamics_behavior.py:205-209 uses it -- with the kernel source of warp_hydrodynamics.py executed through
``oracle/warp_shim`` (NumPy float32 stand-in for the ``warp`` package, which this image does not have; a real
``warp`` is used instead when importable).  Stored: float32-representable inputs + the eight float32 outputs
+ ``raised`` (the kernel source reads an unassigned variable: SURVEY.md Appendix C4 / A.8).

Groups:  c2  hexapod envs (part-table parameters)      c3  heterogeneous boxes
         edge  the hand-written edge cases of oracle/make_golden.py (rounded to float32)
         batched  512 uniform-parameter bodies through ONE ``dim = N`` launch of the reference kernel
                  (the wrapper itself only launches dim = 1)
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_warp  # noqa: E402
from oracle.make_golden import README_CTOR, _from_workload, edge_cases  # noqa: E402
from silver2_isaacsim_b200 import workloads as W  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_warp_golden.npz")


def _f32(d):
    return {k: (np.asarray(v, dtype=np.float32).astype(np.float64) if k not in ("ctor", "mass") else np.asarray(v))
            for k, v in d.items()}


def main():
    groups = {
        "c2": _f32(_from_workload(W.hexapod_envs(64, seed=W.SEED_BASE + 112))),
        "c3": _f32(_from_workload(W.heterogeneous_boxes(2048, seed=W.SEED_BASE + 113))),
        "edge": _f32(edge_cases()),
    }
    blob = {}
    for g, d in groups.items():
        out, raised = ref_warp.components_via_wrapper(d["ctor"], d["pos"], d["quat"], d["v"], d["w"], d["a"], d["al"])
        for k, v in d.items():
            blob[f"{g}/{k}"] = v
        for name in ref_warp.NAMES:
            blob[f"{g}/{name}"] = out[name]
        blob[f"{g}/raised"] = raised
        print(f"{g}: n={len(raised)} raised={int(raised.sum())}")
    # one dim = N launch of the reference kernel (uniform README parameters, C3 state distribution)
    d = _f32(_from_workload(W.uniform_small_batch(512, seed=W.SEED_BASE + 115)))
    per_body, raised = ref_warp.components_via_wrapper(README_CTOR, d["pos"], d["quat"], d["v"], d["w"], d["a"], d["al"])
    keep = ~raised
    for k in ("pos", "quat", "v", "w", "a", "al"):
        d[k] = d[k][keep]
    batched = ref_warp.components_batched(README_CTOR, d["pos"], d["quat"], d["v"], d["w"], d["a"], d["al"])
    for name in ref_warp.NAMES:  # N launches of dim 1 == 1 launch of dim N, bit for bit
        assert batched[name].tobytes() == per_body[name][keep].tobytes(), name
        blob[f"batched/{name}"] = batched[name]
    for k in ("pos", "quat", "v", "w", "a", "al"):
        blob[f"batched/{k}"] = d[k]
    blob["batched/ctor"] = np.asarray(README_CTOR, dtype=np.float64)
    blob["batched/raised"] = np.zeros(int(keep.sum()), bool)
    print(f"batched: n={int(keep.sum())} (dim=N launch == per-body launches)")
    blob["meta/backend"] = np.array([ref_warp.load()[3], f"numpy {np.__version__}"])
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
