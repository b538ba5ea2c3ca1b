"""C4 of SURVEY.md 8(d): STRONG scaling of a fixed fleet -- 8 x 110 592 hexapods x 19 bodies = 16 809 984
bodies, +-20 % per-robot parameter jitter (per-body records), per-robot wrench output -- sharded in
robot-contiguous blocks over the ranks of one torchrun job (1, 2, 4 or 8 GPUs).  No data-path collective.

The fleet is generated in 8 blocks of 110 592 robots (seed = SEED_BASE + 40 + block), so the global data
set is the same for every world size; rank r owns blocks [r*8/G, (r+1)*8/G).  Each rank also scores a
2^16-robot sample of its shard against the float64 oracle (fp32 criteria of tests/scoring.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tests/harness/c4_scaling.py
Prints one JSON line on rank 0.
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from silver2_isaacsim_b200 import HydroEngine, sharding, workloads as W

BLOCKS, ROBOTS_PER_BLOCK, BPR = 8, 110592, 19
STEPS, WARMUP = int(os.environ.get("STEPS", 200)), 10
SAMPLE_ROBOTS = 1 << 16

rank, world, local = sharding.init_distributed()
assert BLOCKS % world == 0, "world size must divide 8"
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
mine = range(rank * BLOCKS // world, (rank + 1) * BLOCKS // world)
parts = [W.sharded_robots(ROBOTS_PER_BLOCK, seed=W.SEED_BASE + 40 + b) for b in mine]
cat = lambda f: np.concatenate([f(p) for p in parts])
n = sum(p.n for p in parts)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
eng = HydroEngine(n, device=dev)
eng.set_globals(parts[0].rho, parts[0].g)
eng.set_params_per_body(cat(lambda p: p.coeff_per_body()))
eng.set_articulation(BPR)
eng.set_prev(t(cat(lambda p: p.prev_lin)), t(cat(lambda p: p.prev_ang)))
pos, quat, lin, ang = (t(cat(lambda p: p.pos)), t(cat(lambda p: p.quat_xyzw)), t(cat(lambda p: p.lin_vel)),
                       t(cat(lambda p: p.ang_vel)))
F, T, Wr = eng.bind(pos, quat, lin, ang, robot_wrench=True)
dt = parts[0].dt

# ---- parity sample (first step, before the timed loop overwrites v_prev)
eng.step_bound(dt)
torch.cuda.synchronize()
sample = parts[0]
ns = min(SAMPLE_ROBOTS, ROBOTS_PER_BLOCK) * BPR
sys.path.insert(0, ROOT)
from oracle import hydro_oracle as O
from tests import scoring
ref = O.step(sample.ctor_rows()[:ns], sample.masses()[:ns], sample.pos[:ns], sample.quat_xyzw[:ns], sample.lin_vel[:ns],
             sample.ang_vel[:ns], sample.prev_lin[:ns].copy(), sample.prev_ang[:ns].copy(), dt, n_threads=max(1, 16 // world))
okF = scoring.fp32_ok(F[:ns].double().cpu().numpy(), ref.force)
okT = scoring.fp32_ok(T[:ns].double().cpu().numpy(), ref.torque)
refw = O.robot_wrench(sample.pos[:ns].astype(np.float64), ref.force, ref.torque, BPR)
w = Wr[:ns // BPR].double().cpu().numpy()
okW = np.abs(w - refw).max(axis=1) <= np.maximum(1e-5 * np.abs(refw).max(axis=1), 1e-4)

# ---- timed loop: one shard is >= 370 MB of traffic per step, far beyond the 126 MB L2
for _ in range(WARMUP):
    eng.step_bound(dt)
torch.cuda.synchronize()
sharding.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    eng.step_bound(dt)
e1.record()
torch.cuda.synchronize()
sharding.barrier()
ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
total = sharding.sum_over_ranks(float(n), dev)
passF = sharding.sum_over_ranks(float(okF.sum()), dev) / sharding.sum_over_ranks(float(okF.size), dev)
passT = sharding.sum_over_ranks(float(okT.sum()), dev) / sharding.sum_over_ranks(float(okT.size), dev)
passW = sharding.sum_over_ranks(float(okW.sum()), dev) / sharding.sum_over_ranks(float(okW.size), dev)
if rank == 0:
    us = ms * 1e3 / STEPS
    bytes_per_body = 168 + 24.0 / BPR
    print(json.dumps({"config": "C4 strong scaling: 884 736 hexapods (16 809 984 bodies) sharded over the ranks",
                      "n_gpus": world, "bodies_total": int(total), "bodies_per_gpu": n, "us_per_step": us,
                      "updates_per_s": total / (us * 1e-6), "gbs_per_gpu": n * bytes_per_body / (us * 1e-6) / 1e9,
                      "kernel": eng.last_kernel, "steps": STEPS,
                      "parity_sample_robots_per_gpu": ns // BPR, "fp32_pass_force": passF, "fp32_pass_torque": passT,
                      "pass_robot_wrench": passW}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
