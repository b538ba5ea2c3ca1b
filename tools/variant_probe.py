"""A/B of library variants (csrc/Makefile TAG=...) on the three workloads that matter: C3 2^20 (per-body records),
C4 shard (19-body robots, per-body records, wrench), C2 (part table, wrench); graph replay, medians.
    VARIANTS=,mb5,hex3 python tools/variant_probe.py      ('' = the default library)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
from silver2_isaacsim_b200 import HydroEngine, workloads as W
dev = torch.device("cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def engines(make, nb, robot):
    out = []
    for b in range(nb):
        wl = make(b)
        e = HydroEngine(wl.n, device=dev); e.set_workload_params(wl)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
        e.set_prev(t(wl.prev_lin), t(wl.prev_ang)); e.bind(t(wl.pos), t(wl.quat_xyzw), t(wl.lin_vel), t(wl.ang_vel), robot_wrench=robot)
        out.append((e, wl))
    return out
def run(es, per, rounds=31):
    dt = es[0][1].dt
    for e, _ in es: e.step_bound(dt)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(per): es[i %% len(es)][0].step_bound(dt)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(rounds):
        ev0.record(); g.replay(); ev1.record(); torch.cuda.synchronize(); ts.append(ev0.elapsed_time(ev1) * 1e3 / per)
    return float(np.median(ts)), float(np.percentile(ts, 10))
res = {}
es = engines(lambda b: W.heterogeneous_boxes(1 << 20, seed=100 + b), 6, False); res["c3_1M"] = run(es, 24); del es
es = engines(lambda b: W.sharded_robots(110592, seed=200 + b), 2, True); res["c4_shard"] = run(es, 10); del es
def deep(wl):
    wl.pos[:, 2] -= 2.0   # every body below the surface: a fleet on the sea bed (the reference scene sits at -18.4 m)
    return wl
es = engines(lambda b: deep(W.sharded_robots(110592, seed=200 + b)), 2, True); res["c4_deep"] = run(es, 10); del es
es = engines(lambda b: W.hexapod_envs(4096, seed=300 + b), 1, True); res["c2"] = run(es, 200); del es
es = engines(lambda b: deep(W.hexapod_envs(4096, seed=300 + b)), 1, True); res["c2_deep"] = run(es, 200); del es
es = engines(lambda b: W.heterogeneous_boxes(1 << 18, seed=400 + b), 24, False); res["c3_256k"] = run(es, 48); del es
print(json.dumps(res))
''' % ROOT
for tag in os.environ.get("VARIANTS", "").split(","):
    env = dict(os.environ)
    if tag:
        env["H2O_LIB_PATH"] = os.path.join(ROOT, "silver2_isaacsim_b200", "lib", f"libh2o_b200_{tag}.so")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{tag or 'default':>10s}: " + "  ".join(f"{k} {v[0]:.2f} (p10 {v[1]:.2f})" for k, v in d.items()), flush=True)
    except Exception:
        print(tag, "FAILED", r.stderr[-800:], flush=True)
