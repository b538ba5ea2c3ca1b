"""ctypes front-end of tests/host_emul/emul.cpp: the model header (the per-body arithmetic
every CUDA kernel inlines) instantiated for the host.  TEST TOOL ONLY."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_emul", "libh2o_emul.so")
SRC = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
HDR = os.path.join(ROOT, "silver2_isaacsim_b200", "csrc", "h2o_model.cuh")
MODE_FP64, MODE_FP32, MODE_ALL_FP32, MODE_FP32_STORE_FP64_MATH, MODE_FP32_FAST = 0, 1, 2, 3, 4
_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))
        if stale:
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            subprocess.run(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-mfma", "-ffp-contract=fast",
                            "-fvisibility=hidden", "-o", SO, SRC], check=True)
        _lib = ctypes.CDLL(SO)
    return _lib


def step(wl, mode, exact_trig=0):
    """Run one fused step of workload ``wl`` through the host-instantiated model."""
    n = wl.n
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F, T = np.zeros((n, 3)), np.zeros((n, 3))
    comp, masks = np.zeros((n, 28)), np.zeros(n, np.uint32)
    arrs = [d(wl.pos), d(wl.quat_xyzw), d(wl.lin_vel), d(wl.ang_vel), d(wl.prev_lin), d(wl.prev_ang),
            d(wl.coeff_per_body())]
    rc = lib().emul_step(mode, exact_trig, ctypes.c_int64(n), *[p(a) for a in arrs], ctypes.c_double(wl.rho),
                         ctypes.c_double(wl.g), ctypes.c_double(wl.dt), p(F), p(T), p(comp), p(masks))
    assert rc == 0
    return F, T, comp, masks


def step_dense(wl, mode, matrices, slot_type):
    """Fused step with dense added-mass matrices (n_types,6,6); body i uses matrices[slot_type[i % n_slots]]."""
    n = wl.n
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F, T = np.zeros((n, 3)), np.zeros((n, 3))
    m = d(matrices).reshape(-1, 6, 6)
    st = np.ascontiguousarray(slot_type, dtype=np.int32)
    arrs = [d(wl.pos), d(wl.quat_xyzw), d(wl.lin_vel), d(wl.ang_vel), d(wl.prev_lin), d(wl.prev_ang),
            d(wl.coeff_per_body())]
    rc = lib().emul_step_dense(mode, ctypes.c_int64(n), *[p(a) for a in arrs], ctypes.c_double(wl.rho),
                               ctypes.c_double(wl.g), ctypes.c_double(wl.dt), p(F), p(T), p(m), p(st),
                               ctypes.c_int(st.size))
    assert rc == 0
    return F, T
