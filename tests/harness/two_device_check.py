"""One process, two devices: engines on cuda:0 and cuda:1 stepped alternately while the current device stays 0
(every entry point must switch to its handle's device and back)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from silver2_isaacsim_b200 import HydroEngine, workloads as W
from oracle import hydro_oracle as O
from tests import scoring
assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
res = []
for d in (0, 1):
    dev = torch.device("cuda", d)
    wl = W.hexapod_envs(4096, seed=10 + d)
    e = HydroEngine(wl.n, device=dev); e.set_workload_params(wl); e.set_kernel("tile")
    t = lambda a: torch.as_tensor(a, device=dev)
    e.set_prev(t(wl.prev_lin), t(wl.prev_ang))
    res.append((e, wl, [t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel)]))
outs = [e.step(*ten, wl.dt, robot_wrench=True) for e, wl, ten in res]
assert torch.cuda.current_device() == 0
for d in (0, 1): torch.cuda.synchronize(d)
for (e, wl, _), (F, T, Wr) in zip(res, outs):
    ref = O.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin, wl.prev_ang, wl.dt)
    scoring.assert_fp32(F.cpu().numpy(), ref.force, f"{F.device} force", min_pass=0.9999)
    scoring.assert_fp32(T.cpu().numpy(), ref.torque, f"{F.device} torque", min_pass=0.9999)
    assert F.device == e.device and Wr.device == e.device
    Fh, Th = e.step_host(wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.dt)[:2]
    assert np.isfinite(Fh).all()
# a tensor on the wrong device is refused
try:
    res[1][0].step(*res[0][2], res[0][1].dt)
    raise SystemExit("tensor on cuda:0 accepted by the cuda:1 engine")
except Exception as ex:
    assert "DEVICE" in str(ex), ex
print("two-device check ok")
