"""A NumPy stand-in for the dozen `warp` (NVIDIA Warp) primitives the reference's Warp twin uses.

TEST INFRASTRUCTURE ONLY (never imported by the product; only oracle/ref_warp.py puts this directory on
sys.path, and only when the real `warp` package is not installed -- it is not in this image).

Purpose: execute the reference's UNMODIFIED kernel source
    /root/reference/src/scripts/physics/warp_hydrodynamics.py          (five @wp.func + one @wp.kernel)
    /root/reference/src/scripts/physics/warp_hydrodynamics_wrapper.py  (WarpHydrodynamicsWrapper)
on the CPU, so that the Warp twin's behaviour is pinned by the reference's own code
rather than by a description of it.  `@wp.func` / `@wp.kernel` bodies are plain Python; what has to be
supplied is the type and builtin layer below.  Arithmetic is float32 like Warp's `float` / `vec3` / `quat`
(NumPy keeps float32 when Python scalars are mixed in).  `quat_rotate` restates Warp's published formula
(warp/native/quat.h):  x (2 w^2 - 1) + 2 q_v (q_v . x) + 2 w (q_v x x).
Launches made between capture_begin / capture_end are recorded, not run, and replayed by capture_launch,
as CUDA-graph capture does.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
__shim__ = True


class vec3:
    __slots__ = ("v",)

    def __init__(self, *a):
        if len(a) == 0:
            self.v = np.zeros(3, f32)
        elif len(a) == 1:
            x = a[0]
            if isinstance(x, vec3):
                self.v = x.v.copy()
            elif np.ndim(x) == 0:
                self.v = np.full(3, f32(x), f32)
            else:
                self.v = np.asarray(x, dtype=f32).reshape(3).copy()
        else:
            self.v = np.array([f32(c) for c in a], dtype=f32)

    def __getitem__(self, i):
        return self.v[i]

    def __add__(self, o):
        return vec3(self.v + o.v)

    def __sub__(self, o):
        return vec3(self.v - o.v)

    def __neg__(self):
        return vec3(-self.v)

    def __mul__(self, s):
        return vec3(self.v * f32(s))

    __rmul__ = __mul__

    def __truediv__(self, s):
        return vec3(self.v / f32(s))

    def __repr__(self):
        return f"vec3{tuple(float(c) for c in self.v)}"


class quat:
    __slots__ = ("v",)

    def __init__(self, *a):
        if len(a) == 1:
            self.v = np.asarray(a[0].v if isinstance(a[0], quat) else a[0], dtype=f32).reshape(4).copy()
        else:
            self.v = np.array([f32(c) for c in a], dtype=f32) if a else np.array([0, 0, 0, 1], f32)

    def __getitem__(self, i):
        return self.v[i]


_WIDTH = {vec3: 3, quat: 4, float: 1, f32: 1}


class array:
    """wp.array: as an annotation (`wp.array(dtype=wp.vec3)`) it is a placeholder; with data it is a typed
    1-D array whose elements read back as vec3 / quat / float32."""

    def __init__(self, data=None, dtype=float, device=None, shape=None):
        self.dtype = dtype
        self.device = device
        w = _WIDTH[dtype]
        if data is None:
            self.data = np.zeros((0, w), f32)
        else:
            self.data = np.asarray(data, dtype=f32).reshape(-1, w).copy()

    @property
    def shape(self):
        return (self.data.shape[0],)

    def __getitem__(self, i):
        row = self.data[i]
        if self.dtype is vec3:
            return vec3(row)
        if self.dtype is quat:
            return quat(row)
        return row[0]

    def __setitem__(self, i, val):
        self.data[i] = val.v if isinstance(val, (vec3, quat)) else f32(val)

    def assign(self, other):
        self.data[...] = other.data.reshape(self.data.shape)


def zeros(n, dtype=float, device=None):
    a = array(None, dtype=dtype, device=device)
    a.data = np.zeros((int(n), _WIDTH[dtype]), f32)
    return a


def from_torch(t, dtype=float):
    a = array(None, dtype=dtype)
    a.data = t.detach().cpu().numpy().astype(f32).reshape(-1, _WIDTH[dtype]).copy()
    return a


def to_torch(a):
    import torch

    d = a.data if a.data.shape[1] > 1 else a.data[:, 0]
    return torch.from_numpy(d.copy())


def func(f):
    return f


class _Kernel:
    def __init__(self, f):
        self.f = f


def kernel(f):
    return _Kernel(f)


_tid = 0
_capturing = None


def tid():
    return _tid


def launch(kernel, dim, inputs=(), outputs=(), device=None):
    def run():
        global _tid
        for i in range(int(dim)):
            _tid = i
            kernel.f(*inputs, *outputs)

    if _capturing is not None:
        _capturing.append(run)
    else:
        run()


def capture_begin(device=None):
    global _capturing
    _capturing = []


def capture_end(device=None):
    global _capturing
    g, _capturing = _capturing, None
    return g


def capture_launch(graph):
    for run in graph:
        run()


def length(a):
    return np.sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2])


def dot(a, b):
    return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]


def cross(a, b):
    x, y = a.v, b.v
    return vec3(x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0])


def cw_mul(a, b):
    return vec3(a.v * b.v)


def quat_rotate(q, x):
    qx, qy, qz, qw = q.v
    c = f32(2.0) * qw * qw - f32(1.0)
    d = f32(2.0) * (qx * x.v[0] + qy * x.v[1] + qz * x.v[2])
    two_w = qw * f32(2.0)
    return vec3(x.v[0] * c + qx * d + (qy * x.v[2] - qz * x.v[1]) * two_w,
                x.v[1] * c + qy * d + (qz * x.v[0] - qx * x.v[2]) * two_w,
                x.v[2] * c + qz * d + (qx * x.v[1] - qy * x.v[0]) * two_w)


def sin(x):
    return np.sin(f32(x))


def asin(x):
    return np.arcsin(f32(x))
