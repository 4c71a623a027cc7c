"""Read-sharded multi-GPU run: one process per GPU, database replicated, each rank probes
its own contiguous share of the reads, then ONE exchange -- a uint8 sum all-reduce (NCCL over NVLink) of
the per-database-k-mer occurrence counters, each rank's counters clamped to ci_min first -- and the
per-genome table is derived from the summed counters.  Per-genome tables of shards are never summed:
the >= ci_min threshold is global and hits have set semantics.

The reference has nothing to compare with (it is a single process around subprocesses,
scripts/select_db.py:50-76); this module is the new host-side plumbing, on torch.distributed.
"""
from __future__ import annotations

import os
from typing import Tuple


class _DevMem:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def device_view_u8(ptr: int, n: int, device: int = 0):
    """torch uint8 tensor aliasing n bytes of device memory at ptr (no copy)."""
    import torch
    return torch.as_tensor(_DevMem(ptr, n), device="cuda:%d" % device)


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) share of n_items for this rank."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_summed_count(ci_min: int, world: int) -> int:
    return ci_min * world


def check_reducible(ci_min: int, world: int) -> None:
    if max_summed_count(ci_min, world) > 255:
        raise ValueError("ci_min * world_size must fit in uint8 for the counter all-reduce")


def allreduce_counts(counts, group=None):
    """In-place sum over ranks of a uint8 tensor of clamped counters (NCCL on GPU, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def reduce_query(query, device: int = 0, group=None) -> None:
    """The exchange step for a live GPU Query: export -> all-reduce in place -> import."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    check_reducible(query_ci_min(query), dist.get_world_size(group))
    ptr, n = query.counts_export()          # joins the library's streams
    t = device_view_u8(ptr, n, device)
    allreduce_counts(t, group)
    torch.cuda.synchronize(device)          # NCCL ran on torch's stream; the library uses its own
    query.counts_import()


def query_ci_min(query) -> int:
    return getattr(query, "ci_min", 2)


def init_from_env(backend: str = "nccl"):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local
