"""Environment sharding across the GPUs of one box (one process per GPU).

The hot path has no exchange step: a body reads only its own state and coefficients
(solve_hydrodynamics, numba_hydrodynamics.py:255-314) and the per-robot wrench couples
only the bodies of one robot.  So the partition is contiguous blocks of WHOLE robots per
rank and there is NO data-path collective.  ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests) is used for rendezvous, barriers, the max-over-ranks timing reduction and
the optional all-reduce of the 8-scalar statistics vector.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import STATS_FIELDS

# how each statistics field combines across ranks
_MAX_FIELDS = ("max_force_norm",)


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    robot_start: int
    n_robots: int
    bodies_per_robot: int

    @property
    def body_start(self) -> int:
        return self.robot_start * self.bodies_per_robot

    @property
    def n_bodies(self) -> int:
        return self.n_robots * self.bodies_per_robot

    def body_slice(self) -> slice:
        return slice(self.body_start, self.body_start + self.n_bodies)


def shard_robots(n_robots_total: int, bodies_per_robot: int, world_size: int, rank: int) -> Shard:
    """Contiguous, balanced blocks of whole robots: the first ``rem`` ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0,{world_size})")
    base, rem = divmod(int(n_robots_total), int(world_size))
    start = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return Shard(rank, world_size, start, count, max(1, int(bodies_per_robot)))


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group when launched under torchrun."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Timing reduction: every multi-GPU number is the slowest rank's."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or _default_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def min_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or _default_device())
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def max_over_ranks_vec(values, device: Optional[torch.device] = None):
    """Element-wise max over ranks of a vector of timings (one entry per timed region)."""
    import numpy as np

    v = np.asarray(values, dtype=np.float64)
    if not dist.is_initialized():
        return v
    t = torch.as_tensor(v, dtype=torch.float64).to(device or _default_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or _default_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_stats(stats: torch.Tensor) -> dict:
    """Global statistics from each rank's (8,) float64 vector (``HydroEngine.stats_tensor``):
    sums for the counters, max for ``max_force_norm``.  The only collective on the path, and
    an optional one (<= 64 bytes, latency-bound)."""
    s = stats.detach().clone().to(torch.float64)
    if dist.is_initialized():
        mx = s.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        for i, name in enumerate(STATS_FIELDS):
            if name in _MAX_FIELDS:
                s[i] = mx[i]
    return dict(zip(STATS_FIELDS, s.cpu().tolist()))


def _default_device() -> torch.device:
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")
