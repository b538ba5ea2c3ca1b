"""The UNMODIFIED reference Numba implementation of the path, imported live.

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing under ``silver2_isaacsim_b200/`` may import this.

Two places the reference can come from:

  live    /root/reference/src/scripts/physics/{numba_hydrodynamics,numba_hydrodynamics_wrapper}.py,
          imported in place (build container only -- the tree does not exist on the GPU box);
  staged  oracle/_ref/physics/*.bc, the same two modules byte-compiled from where they lie by
          ``oracle/stage_reference.py`` (git-ignored build artefact that travels to the GPU box).

Uses: (1) validate the C restatement ``hydro_oracle.c`` and generate ``tests/golden`` (here),
(2) score the CUDA path against the real reference ON the GPU box, (3) bench.py's
``kind: "reference"`` CPU legs: the wrapper called once per body from Python
(as shipped, 1 core) and an ``@njit(parallel=True)`` ``prange`` driver over the untouched
``solve_hydrodynamics`` (all cores), plus the NumPy float64 behaviour tail
(hydrodynamics_behavior.py:196-226).

The reference decorates with ``@njit(cache=True, ...)``; Numba's on-disk cache needs the SOURCE file
next to the code (``no locator available``), which neither a read-only tree nor bytecode offers
reliably, so ``cache`` is dropped while the modules are imported (JIT results are identical; only
the ~10 s compile is repeated per process).  Nothing else about the code is touched.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("H2O_REFERENCE_ROOT", "/root/reference")
_SCRIPTS = os.path.join(REFERENCE_ROOT, "src", "scripts")
_STAGED = os.path.join(_HERE, "_ref")

_loaded = None


def live_available() -> bool:
    return os.path.isfile(os.path.join(_SCRIPTS, "physics", "numba_hydrodynamics.py"))


def staged_available() -> bool:
    return os.path.isfile(os.path.join(_STAGED, "physics", "numba_hydrodynamics.bc")) and \
        os.path.isfile(os.path.join(_STAGED, "physics", "numba_hydrodynamics_wrapper.bc"))


def available() -> bool:
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return live_available() or staged_available()


def source() -> str:
    """'live' (reference tree imported in place), 'staged' (oracle/_ref bytecode) or 'none'."""
    return "live" if live_available() else ("staged" if staged_available() else "none")


def load():
    """Return (NumbaHydrodynamicsWrapper, solve_hydrodynamics) of the unmodified reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = source()
    if src == "none":
        raise RuntimeError(f"reference not present: neither {REFERENCE_ROOT} nor {_STAGED} (oracle/stage_reference.py)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/h2o_numba_cache")  # never inside either tree
    import numba

    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # no __pycache__ in the read-only reference tree
    real_njit = numba.njit

    def njit_without_disk_cache(*args, **kwargs):
        kwargs.pop("cache", None)
        return real_njit(*args, **kwargs)

    numba.njit = njit_without_disk_cache
    for name in ("physics", "physics.numba_hydrodynamics", "physics.numba_hydrodynamics_wrapper"):
        sys.modules.pop(name, None)
    try:
        if src == "live":
            sys.path.insert(0, _SCRIPTS)
            try:
                from physics.numba_hydrodynamics import solve_hydrodynamics  # type: ignore
                from physics.numba_hydrodynamics_wrapper import NumbaHydrodynamicsWrapper  # type: ignore
            finally:
                sys.path.remove(_SCRIPTS)
        else:
            mods = _import_staged()
            solve_hydrodynamics = mods["numba_hydrodynamics"].solve_hydrodynamics
            NumbaHydrodynamicsWrapper = mods["numba_hydrodynamics_wrapper"].NumbaHydrodynamicsWrapper
    finally:
        numba.njit = real_njit
        sys.dont_write_bytecode = old_flag
    _loaded = (NumbaHydrodynamicsWrapper, solve_hydrodynamics)
    return _loaded


def _import_staged():
    """What CPython's sourceless import does with a .pyc, for the two staged modules: unmarshal the code
    object behind the 16-byte header and execute it in a module of the package ``physics`` (the wrapper
    imports ``.numba_hydrodynamics`` relatively)."""
    import marshal
    import types

    pkg = types.ModuleType("physics")
    pkg.__path__ = []  # a package, with nothing importable from disk
    sys.modules["physics"] = pkg
    out = {}
    for m in ("numba_hydrodynamics", "numba_hydrodynamics_wrapper"):
        with open(os.path.join(_STAGED, "physics", m + ".bc"), "rb") as f:
            blob = f.read()
        code = marshal.loads(blob[16:])
        mod = types.ModuleType("physics." + m)
        mod.__package__ = "physics"
        mod.__file__ = code.co_filename
        sys.modules["physics." + m] = mod
        setattr(pkg, m, mod)
        exec(code, mod.__dict__)
        out[m] = mod
    return out


def components_via_wrapper(ctor_rows, pos, quat_xyzw, lin_vel, ang_vel, lin_acc, ang_acc,
                           dense=None, dense_slot=None):
    """Call the reference wrapper once per body (how it is meant to be used).

    ``dense`` (n_types,6,6) + ``dense_slot``: body i's wrapper instance gets
    ``_added_mass_matrix = dense[dense_slot[i % len(dense_slot)]]`` after construction, i.e. the
    unmodified ``solve_hydrodynamics`` / ``calculate_added_mass`` run with a full matrix.

    Returns (records, raised) where ``records`` uses hydro_oracle.OUT_DTYPE and
    ``raised[i]`` is True when the reference threw TypeError.
    """
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


# ---------------------------------------------------------------------------------------------
# Batched driver: the untouched solve_hydrodynamics once per body inside a prange loop
# (SURVEY.md 8(d) "CPU baseline (2)").  The harness owns only the loop and the per-body geometry
# arrays the wrapper's constructor would have built (numba_hydrodynamics_wrapper.py:55-112).
# ---------------------------------------------------------------------------------------------
def box_geometry(ctor_rows):
    """Per-body constant arrays exactly as ``NumbaHydrodynamicsWrapper.__init__`` builds them
    (numba_hydrodynamics_wrapper.py:55-112), vectorised over bodies:
    keypoints (n,27,3), face_centers (n,6,3), face_areas (n,6), face_normals (6,3), added_mass (n,6,6).
    1128 B/body -- callers chunk.  ``check_geometry`` pins this against real wrapper instances."""
    c = np.asarray(ctor_rows, dtype=np.float64).reshape(-1, 12)
    n = c.shape[0]
    w, d, h = c[:, 0], c[:, 1], c[:, 2]
    x, y, z = w / 2.0, d / 2.0, h / 2.0
    zero = np.zeros(n)
    kp = np.empty((n, 27, 3))
    k = 0
    for zz in (z, zero, -z):            # top, middle, bottom layer
        for yy in (y, zero, -y):
            for xx in (-x, zero, x):
                kp[:, k, 0], kp[:, k, 1], kp[:, k, 2] = xx, yy, zz
                k += 1
    fc = np.zeros((n, 6, 3))
    fc[:, 0, 0], fc[:, 1, 0] = x, -x
    fc[:, 2, 1], fc[:, 3, 1] = y, -y
    fc[:, 4, 2], fc[:, 5, 2] = z, -z
    fa = np.stack([d * h, d * h, w * h, w * h, w * d, w * d], axis=1)
    fn = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    vol = w * d * h
    rho, cam, cama = c[:, 7], c[:, 9], c[:, 10]
    am = np.zeros((n, 6, 6))
    lin = vol * cam * rho
    am[:, 0, 0] = am[:, 1, 1] = am[:, 2, 2] = lin
    am[:, 3, 3] = vol * (d ** 2 + h ** 2) * cama * rho
    am[:, 4, 4] = vol * (w ** 2 + h ** 2) * cama * rho
    am[:, 5, 5] = vol * (w ** 2 + d ** 2) * cama * rho
    return np.ascontiguousarray(kp), fc, np.ascontiguousarray(fa), fn, am


def check_geometry(ctor_rows, sample: int = 32, seed: int = 0) -> None:
    """The vectorised arrays above equal, bit for bit, what the reference constructor builds."""
    Wrapper, _ = load()
    c = np.asarray(ctor_rows, dtype=np.float64).reshape(-1, 12)
    idx = np.unique(np.random.default_rng(seed).integers(0, len(c), size=min(sample, len(c))))
    kp, fc, fa, fn, am = box_geometry(c[idx])
    for j, i in enumerate(idx):
        w = Wrapper(*c[i])
        assert np.array_equal(w._local_keypoints, kp[j]) and np.array_equal(w._local_face_centers, fc[j])
        assert np.array_equal(w._face_areas, fa[j]) and np.array_equal(w._local_face_normals, fn)
        assert np.array_equal(w._added_mass_matrix, am[j]) and w.total_volume == c[i, 0] * c[i, 1] * c[i, 2]


_driver = None


def prange_driver():
    """``drive(pos, quat, v, w, a, al, ctor, kp, fc, fa, fn, am, out)``: for every body one call of the
    reference ``solve_hydrodynamics`` (numba_hydrodynamics.py:255-314); ``out`` is (n,25) =
    eight 3-vectors in the reference's return order + sub_ratio.  Bodies at rest (speed <= 1e-6) must
    not be passed: a wet one makes the reference raise, which a prange body cannot."""
    global _driver
    if _driver is None:
        from numba import njit, prange

        _, solve = load()

        @njit(parallel=True)
        def drive(pos, quat, v, w, a, al, ctor, kp, fc, fa, fn, am, out):
            for i in prange(pos.shape[0]):
                c = ctor[i]
                r = solve(pos[i], quat[i], v[i], w[i], a[i], al[i], c[0] * c[1] * c[2], c[7], c[8], c[3], c[4],
                          c[5], c[6], c[11], kp[i], fc[i], fa[i], fn, am[i])
                out[i, 0:3] = r[0]
                out[i, 3:6] = r[1]
                out[i, 6:9] = r[2]
                out[i, 9:12] = r[3]
                out[i, 12:15] = r[4]
                out[i, 15:18] = r[5]
                out[i, 18:21] = r[6]
                out[i, 21:24] = r[7]
                out[i, 24] = r[8]

        _driver = drive
    return _driver


def set_threads(n: int) -> int:
    import numba

    n = max(1, min(int(n), numba.config.NUMBA_NUM_THREADS))
    numba.set_num_threads(n)
    return n


def max_threads() -> int:
    import numba

    return int(numba.config.NUMBA_NUM_THREADS)


class ReferenceStepper:
    """One behaviour step (hydrodynamics_behavior.py:194-238) for n bodies through the real reference:
    finite-difference acceleration (NumPy), ``solve_hydrodynamics`` per body (prange driver), lever
    arms / sum / clamp (NumPy, ``hydro_oracle.numpy_tail``).  Geometry is built once (constructor work
    in the reference).  At-rest bodies go through the wrapper one by one, with the reference's
    ``TypeError`` caught and reported in ``raised``."""

    def __init__(self, ctor_rows, masses, chunk: int = 1 << 18):
        self.ctor = np.ascontiguousarray(np.asarray(ctor_rows, dtype=np.float64).reshape(-1, 12))
        self.mass = np.asarray(masses, dtype=np.float64).reshape(-1)
        self.n = self.ctor.shape[0]
        self.chunk = int(chunk)
        self.geom = [box_geometry(self.ctor[b:b + self.chunk]) for b in range(0, self.n, self.chunk)]
        self.drive = prange_driver()

    def components(self, pos, quat, v, w, a, al):
        from .hydro_oracle import COMPONENT_NAMES, OUT_DTYPE

        f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
        pos, quat, v, w, a, al = map(f8, (pos, quat, v, w, a, al))
        raw = np.zeros((self.n, 25))
        still = np.linalg.norm(v, axis=1) <= 1e-6
        raised = np.zeros(self.n, dtype=bool)
        for ci, b in enumerate(range(0, self.n, self.chunk)):
            e = min(self.n, b + self.chunk)
            kp, fc, fa, fn, am = self.geom[ci]
            if still[b:e].any():
                keep = np.nonzero(~still[b:e])[0]
                sub = np.zeros((len(keep), 25))
                self.drive(pos[b:e][keep], quat[b:e][keep], v[b:e][keep], w[b:e][keep], a[b:e][keep], al[b:e][keep],
                           self.ctor[b:e][keep], kp[keep], fc[keep], fa[keep], fn, am[keep], sub)
                raw[b:e][keep] = sub
            else:
                self.drive(pos[b:e], quat[b:e], v[b:e], w[b:e], a[b:e], al[b:e], self.ctor[b:e], kp, fc, fa, fn, am,
                           raw[b:e])
        out = np.zeros(self.n, dtype=OUT_DTYPE)
        for k, name in enumerate(COMPONENT_NAMES):
            out[name] = raw[:, 3 * k:3 * k + 3]
        out["sub_ratio"] = raw[:, 24]
        idx = np.nonzero(still)[0]
        if len(idx):
            rec, r = components_via_wrapper(self.ctor[idx], pos[idx], quat[idx], v[idx], w[idx], a[idx], al[idx])
            out[idx] = rec
            raised[idx] = r
        return out, raised

    def step(self, pos, quat, v, w, prev_v, prev_w, dt):
        """Returns (force, torque, components, raised); raised bodies carry zeros."""
        from .hydro_oracle import numpy_tail

        v64, w64 = np.asarray(v, dtype=np.float64), np.asarray(w, dtype=np.float64)
        a = (v64 - np.asarray(prev_v, dtype=np.float64)) / dt      # hydrodynamics_behavior.py:200-202
        al = (w64 - np.asarray(prev_w, dtype=np.float64)) / dt
        comp, raised = self.components(pos, quat, v64, w64, a, al)
        F, T, _ = numpy_tail(comp, np.asarray(pos, dtype=np.float64), self.mass)
        return F, T, comp, raised
