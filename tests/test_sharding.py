"""Multi-GPU host logic on the CPU: env-sharding arithmetic and the gloo world_size-2 path."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_robots_partitions_whole_robots():
    from silver2_isaacsim_b200.sharding import shard_robots

    for total, world in ((110592 * 8, 8), (4099, 4), (7, 8), (1, 1), (100, 3)):
        shards = [shard_robots(total, 19, world, r) for r in range(world)]
        assert shards[0].robot_start == 0
        for a, b in zip(shards, shards[1:]):
            assert a.robot_start + a.n_robots == b.robot_start      # contiguous, no overlap
        assert shards[-1].robot_start + shards[-1].n_robots == total
        counts = [s.n_robots for s in shards]
        assert max(counts) - min(counts) <= 1                        # balanced
        assert all(s.n_bodies == s.n_robots * 19 and s.body_start == s.robot_start * 19 for s in shards)
    sl = shard_robots(10, 19, 2, 1).body_slice()
    assert (sl.start, sl.stop) == (95, 190)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %(root)r)
    import numpy as np, torch
    import torch.distributed as dist
    from silver2_isaacsim_b200 import sharding, workloads as W
    from oracle import hydro_oracle as O

    rank, world, local = sharding.init_distributed("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    # every rank generates the same global workload and keeps its own block of whole robots
    wl = W.hexapod_envs(33)
    sh = sharding.shard_robots(33, wl.bodies_per_robot, world, rank)
    sl = sh.body_slice()
    # the per-rank "engine" on a CPU-only box is the oracle (the GPU engine has no CPU path);
    # what is under test is the partition, the statistics all-reduce and the timing reduction
    r = O.step(wl.ctor_rows()[sl], wl.masses()[sl], wl.pos[sl], wl.quat_xyzw[sl], wl.lin_vel[sl], wl.ang_vel[sl],
               wl.prev_lin[sl], wl.prev_ang[sl], wl.dt, n_threads=1)
    wr = O.robot_wrench(wl.pos[sl], r.force, r.torque, wl.bodies_per_robot)
    norms = np.linalg.norm(r.force, axis=1)
    stats = torch.tensor([norms.sum(), norms.max(), float((r.components["sub_ratio"] > 0).sum()),
                          float(((r.flags & 2) != 0).sum()), 0.0, 0.0, float(sh.n_bodies), 0.0], dtype=torch.float64)
    g = sharding.allreduce_stats(stats)
    tmax = sharding.max_over_ranks(1.0 + rank)
    tsum = sharding.sum_over_ranks(float(sh.n_bodies))
    tmin = sharding.min_over_ranks(1.0 + rank)
    # per-region timing vectors of bench.py: element-wise max over ranks
    tvec = sharding.max_over_ranks_vec([1.0 + rank, 5.0 - 3 * rank, 2.0]).tolist()
    sharding.barrier()
    out = dict(rank=rank, start=sh.robot_start, count=sh.n_robots, g=g, tmax=tmax, tsum=tsum, tmin=tmin, tvec=tvec,
               wrench_sum=wr.sum(axis=0).tolist())
    json.dump(out, open(os.path.join(%(tmp)r, f"rank{rank}.json"), "w"))
    dist.destroy_process_group()
""")


def test_gloo_world_size_2(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "tmp": str(tmp_path)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    from silver2_isaacsim_b200 import workloads as W
    outs = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    assert (outs[0]["start"], outs[0]["count"], outs[1]["start"], outs[1]["count"]) == (0, 17, 17, 16)
    # global statistics == single-process statistics of the whole workload
    wl = W.hexapod_envs(33)
    r = oracle.step(wl.ctor_rows(), wl.masses(), wl.pos, wl.quat_xyzw, wl.lin_vel, wl.ang_vel, wl.prev_lin,
                    wl.prev_ang, wl.dt)
    norms = np.linalg.norm(r.force, axis=1)
    for o in outs:
        g = o["g"]
        assert g["bodies"] == wl.n and o["tsum"] == wl.n and o["tmax"] == 2.0
        assert o["tmin"] == 1.0 and o["tvec"] == [2.0, 5.0, 2.0]
        assert abs(g["sum_force_norm"] - norms.sum()) <= 1e-9 * norms.sum()
        assert g["max_force_norm"] == norms.max()
        assert g["wet_bodies"] == int((r.components["sub_ratio"] > 0).sum())
    # shards reproduce the unsharded per-robot wrenches
    wr = oracle.robot_wrench(wl.pos, r.force, r.torque, 19)
    np.testing.assert_allclose(np.add(outs[0]["wrench_sum"], outs[1]["wrench_sum"]), wr.sum(axis=0), rtol=1e-10)
