"""Stage the reference's own Numba implementation of the path for the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

``/root/reference`` exists only in the build container.  Like a C reference that is compiled
from its sources where they lie into ``oracle/_ref/*.so``, the two Python modules of this path

    /root/reference/src/scripts/physics/numba_hydrodynamics.py          (the seven @njit functions)
    /root/reference/src/scripts/physics/numba_hydrodynamics_wrapper.py  (NumbaHydrodynamicsWrapper)

are byte-compiled from where they lie into ``oracle/_ref/physics/*.bc`` (CPython bytecode in the
``.pyc`` container format, no source text; the extension is not ``.pyc`` because repository snapshots
commonly drop ``*.pyc`` / ``__pycache__`` -- the gpurun snapshot does).  ``oracle/_ref/`` is git-ignored (kept out of history) but not gpurun-ignored, so the
compiled modules travel to the B200 box exactly like the in-tree ``.so`` files.  Nothing is
modified: the code objects are what CPython would build from the untouched files, and Numba JITs
from bytecode.  ``oracle/ref_numba.py`` loads them (marshal + exec, exactly what CPython's sourceless import does) when
``/root/reference`` is absent.

    python -m oracle.stage_reference          # (re)stage; no-op without /root/reference
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
MODULES = ("numba_hydrodynamics", "numba_hydrodynamics_wrapper")


def reference_physics_dir(root: str | None = None) -> str:
    root = root or os.environ.get("H2O_REFERENCE_ROOT", "/root/reference")
    return os.path.join(root, "src", "scripts", "physics")


def stage(root: str | None = None, quiet: bool = False) -> bool:
    """Byte-compile the reference modules into oracle/_ref/physics/.  Returns False (and leaves
    whatever is staged alone) when the reference tree is not present."""
    src_dir = reference_physics_dir(root)
    if not all(os.path.isfile(os.path.join(src_dir, m + ".py")) for m in MODULES):
        return False
    out_dir = os.path.join(REF_DIR, "physics")
    os.makedirs(out_dir, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "magic": py_compile.importlib.util.MAGIC_NUMBER.hex(),
                "modules": {}}
    for m in MODULES:
        src = os.path.join(src_dir, m + ".py")
        dst = os.path.join(out_dir, m + ".bc")
        # dfile: the path shown in tracebacks; keep the reference's own location
        py_compile.compile(src, cfile=dst, dfile=src, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        with open(src, "rb") as f:
            manifest["modules"][m] = {"source": src, "sha256": hashlib.sha256(f.read()).hexdigest()}
    try:
        import numba

        manifest["numba"] = numba.__version__
    except Exception:
        manifest["numba"] = None
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if not quiet:
        print(f"staged {len(MODULES)} reference modules (bytecode only) into {out_dir}")
    return True


def staged() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, "physics", m + ".bc")) for m in MODULES)


if __name__ == "__main__":
    if not stage():
        print("reference tree not present; nothing staged", file=sys.stderr)
        sys.exit(0 if staged() else 1)
