"""The UNMODIFIED reference Warp twin of the path, executed on the CPU.

TEST INFRASTRUCTURE ONLY.  Nothing under ``silver2_isaacsim_b200/`` may import this.

``/root/reference/src/scripts/physics/warp_hydrodynamics.py`` (the kernel) and
``warp_hydrodynamics_wrapper.py`` (WarpHydrodynamicsWrapper) are imported in place.  They need the ``warp``
package, which this image does not have; ``oracle/warp_shim`` supplies the dozen
primitives they use in NumPy float32, so the reference's own source decides every branch, threshold and
formula of the Warp twin.  With a real ``warp`` installed that one is used instead
and the same harness becomes the A/B of SURVEY.md 8(f3).

Uses: ``oracle/make_golden_warp.py`` (fixture for the Warp-compat mode), ``tests/test_reference_live.py``.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("H2O_REFERENCE_ROOT", "/root/reference")
_SCRIPTS = os.path.join(REFERENCE_ROOT, "src", "scripts")
_SHIM = os.path.join(_HERE, "warp_shim")
NAMES = ("buoyancy_force", "drag_force", "lift_force", "drag_torque", "added_mass_force", "added_mass_torque",
         "center_of_buoyancy", "center_of_pressure")

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(_SCRIPTS, "physics", "warp_hydrodynamics.py"))


def load():
    """Return (WarpHydrodynamicsWrapper, solve_hydrodynamics_kernel, wp, backend) with backend 'warp' | 'shim'."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not present: {REFERENCE_ROOT}")
    try:
        wp = importlib.import_module("warp")
        backend = "shim" if getattr(wp, "__shim__", False) else "warp"
    except ImportError:
        sys.path.insert(0, _SHIM)
        try:
            wp = importlib.import_module("warp")
        finally:
            sys.path.remove(_SHIM)
        backend = "shim"
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # no __pycache__ in the read-only reference tree
    for name in ("physics.warp_hydrodynamics", "physics.warp_hydrodynamics_wrapper"):
        sys.modules.pop(name, None)
    had_pkg = "physics" in sys.modules
    sys.path.insert(0, _SCRIPTS)
    try:
        from physics.warp_hydrodynamics import solve_hydrodynamics_kernel  # type: ignore
        from physics.warp_hydrodynamics_wrapper import WarpHydrodynamicsWrapper  # type: ignore
    finally:
        sys.path.remove(_SCRIPTS)
        sys.dont_write_bytecode = old_flag
        if not had_pkg:
            pass  # the namespace package 'physics' stays importable for ref_numba.load()
    _loaded = (WarpHydrodynamicsWrapper, solve_hydrodynamics_kernel, wp, backend)
    return _loaded


def components_via_wrapper(ctor, pos, quat, v, w, a, al):
    """One ``WarpHydrodynamicsWrapper.calculate_hydrodynamic_forces`` call per body -- exactly how
    hydrodynamics_behavior.py:205-209 uses it ((1,3)/(1,4) float32 tensors in, eight (1,3) tensors out).
    Returns ({name: (n,3) float32}, raised (n,) bool); ``raised`` marks bodies for which the kernel source
    reads a variable it never assigned (undefined behaviour under real Warp)."""
    import torch

    Wrapper, _, _, backend = load()
    device = "cpu"
    ctor = np.asarray(ctor, dtype=np.float64)
    n = len(pos)
    out = {k: np.zeros((n, 3), np.float32) for k in NAMES}
    raised = np.zeros(n, bool)
    cache = {}
    t = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(1, -1), device=device)
    for i in range(n):
        row = tuple(ctor if ctor.ndim == 1 else ctor[i])
        wrp = cache.get(row)
        if wrp is None:
            wrp = cache[row] = Wrapper(*[float(c) for c in row], device=device)
        try:
            res = wrp.calculate_hydrodynamic_forces(t(pos[i]), t(quat[i]), t(v[i]), t(w[i]), t(a[i]), t(al[i]))
        except (UnboundLocalError, TypeError):
            raised[i] = True
            continue
        for k, r in zip(NAMES, res):
            out[k][i] = r.numpy().reshape(3)
    return out, raised


def components_batched(ctor_row, pos, quat, v, w, a, al):
    """The A/B of SURVEY.md 8(f3): the reference's kernel launched ONCE with dim = N over batched buffers
    (the wrapper only ever launches dim = 1), uniform parameters.  Bodies that would raise are not caught
    here: use inputs without them."""
    Wrapper, kern, wp, backend = load()
    n = len(pos)
    wrp = Wrapper(*[float(c) for c in ctor_row], device="cpu")
    arr = lambda x, dt: wp.array(np.asarray(x, dtype=np.float32), dtype=dt, device="cpu")
    ins = [arr(pos, wp.vec3), arr(quat, wp.quat), arr(v, wp.vec3), arr(w, wp.vec3), arr(a, wp.vec3), arr(al, wp.vec3),
           wrp.wp_keypoints, wrp.wp_normals, wrp.wp_centers, wrp.wp_areas, wrp.wp_params, wrp.wp_added_mass]
    outs = [wp.zeros(n, dtype=wp.vec3, device="cpu") for _ in NAMES]
    wp.launch(kernel=kern, dim=n, inputs=ins, outputs=outs, device="cpu")
    if backend == "warp":
        wp.synchronize()
        return {k: o.numpy().reshape(n, 3) for k, o in zip(NAMES, outs)}
    return {k: o.data.copy() for k, o in zip(NAMES, outs)}
