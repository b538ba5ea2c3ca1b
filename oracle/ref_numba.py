"""Live import of the UNMODIFIED reference Numba code (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may import this module; it exists to (1) validate the C
restatement in ``hydro_oracle.c`` against the real thing and (2) generate the
golden vectors committed under ``tests/golden/`` (``oracle/make_golden.py``).

The reference is imported in place, namespace-package style:
    /root/reference/src/scripts/physics/numba_hydrodynamics.py
    /root/reference/src/scripts/physics/numba_hydrodynamics_wrapper.py
``cache=True`` in its ``@njit`` decorators would write ``__pycache__/*.nbi``
into the read-only tree, so NUMBA_CACHE_DIR is pointed at /tmp first.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("H2O_REFERENCE_ROOT", "/root/reference")
_SCRIPTS = os.path.join(REFERENCE_ROOT, "src", "scripts")


def available() -> bool:
    return os.path.isfile(os.path.join(_SCRIPTS, "physics", "numba_hydrodynamics.py"))


def load():
    """Return (NumbaHydrodynamicsWrapper, solve_hydrodynamics) from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/h2o_numba_cache")
    sys.dont_write_bytecode = True
    if _SCRIPTS not in sys.path:
        sys.path.insert(0, _SCRIPTS)
    from physics.numba_hydrodynamics import solve_hydrodynamics  # type: ignore
    from physics.numba_hydrodynamics_wrapper import NumbaHydrodynamicsWrapper  # type: ignore

    return NumbaHydrodynamicsWrapper, solve_hydrodynamics


def components_via_wrapper(ctor_rows, pos, quat_xyzw, lin_vel, ang_vel, lin_acc, ang_acc,
                           dense=None, dense_slot=None):
    """Call the reference wrapper once per body (how it is meant to be used).

    ``dense`` (n_types,6,6) + ``dense_slot``: body i's wrapper instance gets
    ``_added_mass_matrix = dense[dense_slot[i % len(dense_slot)]]`` after construction, i.e. the
    unmodified ``solve_hydrodynamics`` / ``calculate_added_mass`` run with a full matrix.

    Returns (records, raised) where ``records`` uses hydro_oracle.OUT_DTYPE and
    ``raised[i]`` is True when the reference threw TypeError.
    """
    import numpy as np

    from .hydro_oracle import COMPONENT_NAMES, OUT_DTYPE

    Wrapper, _ = load()
    n = len(pos)
    ctor_rows = np.asarray(ctor_rows, dtype=np.float64)
    out = np.zeros(n, dtype=OUT_DTYPE)
    raised = np.zeros(n, dtype=bool)
    cache = {}
    for i in range(n):
        row = tuple(ctor_rows[i] if ctor_rows.ndim == 2 else ctor_rows)
        ty = -1 if dense is None else int(dense_slot[i % len(dense_slot)])
        w = cache.get((row, ty))
        if w is None:
            w = cache[(row, ty)] = Wrapper(*row)
            if ty >= 0:
                w._added_mass_matrix = np.ascontiguousarray(dense[ty], dtype=np.float64)
        try:
            r = w.calculate_hydrodynamic_forces(pos[i], quat_xyzw[i], lin_vel[i], ang_vel[i],
                                                lin_acc[i], ang_acc[i])
        except TypeError:
            raised[i] = True
            continue
        for k, name in enumerate(COMPONENT_NAMES):
            out[name][i] = r[k]
        out["sub_ratio"][i] = r[8]
    return out, raised
