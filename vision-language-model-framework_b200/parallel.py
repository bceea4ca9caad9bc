"""Data parallelism for the hot path: one process per GPU, images sharded by contiguous slices, weights replicated.

Inference has no collective (every image is independent — BatchNorm uses running statistics in eval). Probe training
exchanges two tiny buffers per step (BatchNorm batch sums, then the flat gradient bucket) with ``all_reduce(SUM)`` —
NCCL over NVLink on GPUs, gloo in the CPU tests. See DESIGN.md §multi-GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class DistEnv:
    rank: int = 0
    world: int = 1
    local_rank: int = 0

    @property
    def is_distributed(self) -> bool:
        return self.world > 1


def env_from_os() -> DistEnv:
    """torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK); single process when absent."""
    return DistEnv(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                   int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: Optional[str] = None) -> DistEnv:
    env = env_from_os()
    if env.is_distributed and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(env.local_rank)
            dist.init_process_group(backend, rank=env.rank, world_size=env.world,
                                    device_id=torch.device("cuda", env.local_rank))
        else:
            dist.init_process_group(backend, rank=env.rank, world_size=env.world)
    return env


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ``n`` items owned by ``rank``: the first ``n % world`` ranks get one extra item,
    so concatenating the ranks' outputs in rank order restores the original order."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_rows(t: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gather per-rank row blocks (sizes from ``shard_range``) into the full [n_total, ...] tensor on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
