"""Multi-GPU plumbing: replicas only.

Self-play games are independent and the weights are read-only (training/self-play/src/self_play.rs:94-141,179-246),
so N GPUs = N independent evaluators, each with its own worker pool, queue and positions; nothing crosses GPUs on the
data path.  The only collective use is measurement: a barrier around the timed region and the max over ranks of the
device time (bench.py).  Backend "nccl" on the GPU box, "gloo" in the CPU tests.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional


@dataclass
class Replica:
    rank: int
    world: int
    local_rank: int
    dist: Optional[object] = None  # torch.distributed when world > 1
    device: Optional[object] = None

    def barrier(self) -> None:
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def aggregate_throughput(self, units_local: float, seconds_local: float) -> float:
        """Whole-job throughput: all ranks' units over the slowest rank's time."""
        return self.sum_over_ranks(units_local) / self.max_over_ranks(seconds_local)

    def close(self) -> None:
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None


def init_from_env(backend: str = "nccl") -> Replica:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  world == 1 never touches torch.distributed."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return Replica(rank, world, local_rank)
    import torch
    import torch.distributed as dist

    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", device_id=device)
    else:
        device = torch.device("cpu")
        dist.init_process_group(backend)
    return Replica(rank, world, local_rank, dist, device)


def rank_seed(base: int, rank: int) -> int:
    """Per-replica position stream (seed = 0xCA7705 + gpu index)."""
    return base + rank


def partition_workers(n_workers: int, n_gpus: int) -> List[List[int]]:
    """Self-play worker w feeds the evaluator of GPU w % n_gpus."""
    assert n_gpus >= 1
    return [[w for w in range(n_workers) if w % n_gpus == g] for g in range(n_gpus)]
