"""Precision study (CPU only): model header (host-instantiated) vs the f64 oracle."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hydro_oracle as O
from silver2_isaacsim_b200 import workloads as W

E = ctypes.CDLL(os.path.join(os.path.dirname(__file__), '..', 'tests', '_emul', 'libh2o_emul.so'))
def P(a): return a.ctypes.data_as(ctypes.c_void_p)

def emul(wl, mode, exact):
    n = wl.n
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    F = np.zeros((n,3)); T = np.zeros((n,3)); comp = np.zeros((n,28)); masks = np.zeros(n, np.uint32)
    arrs = [d(wl.pos), d(wl.quat_xyzw), d(wl.lin_vel), d(wl.ang_vel), d(wl.prev_lin), d(wl.prev_ang), d(wl.coeff_per_body())]
    E.emul_step(mode, exact, ctypes.c_int64(n), *[P(a) for a in arrs], ctypes.c_double(wl.rho), ctypes.c_double(wl.g),
                ctypes.c_double(wl.dt), P(F), P(T), P(comp), P(masks))
    return F, T, comp, masks

def score(x, y, rel, abs_):
    err = np.abs(x-y).max(axis=1); den = np.abs(y).max(axis=1)
    tol = np.maximum(rel*den, abs_)
    return err, den, err > tol

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    for name, wl in [('C3', W.heterogeneous_boxes(n)), ('C2', W.hexapod_envs(n//19))]:
        t=time.time()
        ref = O.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt)
        print(name, 'n', wl.n, 'oracle %.2fs'%(time.time()-t), 'raises', int((ref.flags&1).sum()), 'clamped', int((ref.flags&2).sum()>>1))
        for mode, exact, label, rel, ab in [(1,0,'fp32-mixed',1e-5,1e-6),(2,0,'fp32-all',1e-5,1e-6),(3,0,'fp32store-f64arith',1e-5,1e-6)]:
            F,T,comp,masks = emul(wl, mode, exact)
            for nm,x,y in [('F',F,ref.force),('T',T,ref.torque)]:
                err,den,bad = score(x,y,rel,ab)
                relerr = err/np.maximum(den,1e-30)
                wet = den>0
                print(f'  {label:20s} {nm}: fail {int(bad.sum()):6d}/{wl.n}  median rel {np.median(relerr[wet]):.2e} p99.9 {np.quantile(relerr[wet],0.999):.2e} max rel {relerr[wet].max():.2e} max abs {err.max():.2e}')
