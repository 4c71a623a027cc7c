"""Read-sharded multi-GPU run: one process per GPU, database replicated, each rank probes
its own contiguous share of the reads, then ONE exchange step -- the sum over ranks of the per-database-k-mer
occurrence counters, each rank's counters clamped to ci_min first -- and the per-genome table is derived from
the summed counters.  Per-genome tables of shards are never summed: the >= ci_min threshold is global and
hits have set semantics.

The sum comes in two equivalent forms (NCCL over NVLink on GPUs, gloo on CPU):
  "sparse" (default)  the counter table is >99.9 % zeros, so every rank all-gathers its non-zero counters
                      (64-bit entries: index | count << 32, a few MB) and adds the other ranks' entries into
                      its own table;
  "dense"             one uint8 sum all-reduce of the whole table (|D| bytes, 0.16 GB for the default-scale
                      database) -- MLG_EXCHANGE=dense.

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


def device_view_i64(ptr: int, n: int, device: int = 0):
    """torch int64 tensor aliasing n 8-byte words of device memory at ptr (no copy)."""
    import torch

    class _M:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(_M(), device="cuda:%d" % device)


def allgather_sparse(entries, group=None):
    """entries: 1-D int64 tensor (CPU for gloo, CUDA for NCCL) of this rank's sparse counters.  Returns the list
    of every rank's entries (zero-padded to the longest; padding has count 0 and is ignored by the merge) and the
    true lengths."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_loc = torch.tensor([entries.numel()], dtype=torch.int64, device=entries.device)
    flat = entries.is_cuda          # NCCL: one flat output buffer per collective; gloo: the list form
    if flat:
        sz = torch.empty(world, dtype=torch.int64, device=entries.device)
        dist.all_gather_into_tensor(sz, n_loc, group=group)
        sizes = sz.tolist()
    else:
        szl = [torch.zeros_like(n_loc) for _ in range(world)]
        dist.all_gather(szl, n_loc, group=group)
        sizes = [int(x.item()) for x in szl]
    m = max(max(sizes), 1)
    send = torch.zeros(m, dtype=torch.int64, device=entries.device)
    send[:entries.numel()] = entries
    if flat:
        out = torch.empty(world * m, dtype=torch.int64, device=entries.device)
        dist.all_gather_into_tensor(out, send, group=group)
        recv = [out[r * m:(r + 1) * m] for r in range(world)]
    else:
        recv = [torch.empty(m, dtype=torch.int64, device=entries.device) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
    return recv, sizes


def exchange_mode() -> str:
    m = os.environ.get("MLG_EXCHANGE", "sparse")
    if m not in ("sparse", "dense"):
        raise ValueError("MLG_EXCHANGE must be 'sparse' or 'dense'")
    return m


def reduce_query(query, device: int = 0, group=None, mode: str = None) -> None:
    """The exchange step for a live GPU Query (see the module docstring for the two forms)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    check_reducible(query_ci_min(query), world)
    mode = mode or exchange_mode()
    if mode == "dense":
        ptr, n = query.counts_export()          # joins the library's streams
        t = device_view_u8(ptr, n, device)
        allreduce_counts(t, group)
        torch.cuda.synchronize(device)          # NCCL ran on torch's stream; the library uses its own
        query.counts_import()
        return
    ptr, n = query.counts_export_sparse()       # joins the library's streams
    dev = "cuda:%d" % device
    # ONE all-gather of fixed-capacity buffers: word 0 = this rank's entry count, then its entries, zero padding (count 0:
    # ignored by the merge).  The capacity adapts: if some rank had more entries than fit, every rank sees that in
    # the gathered counts and the gather is repeated with room for the largest.
    global _SPARSE_CAP
    while True:
        cap = max(_SPARSE_CAP, 1024)
        send = torch.zeros(cap + 1, dtype=torch.int64, device=dev)
        send[0] = n
        if 0 < n <= cap:
            send[1:n + 1] = device_view_i64(ptr, n, device)
        out = torch.empty(world * (cap + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, send, group=group)
        counts = out[::cap + 1].tolist()        # synchronises torch's stream: the gather has landed
        if max(counts) <= cap:
            break
        _SPARSE_CAP = 1 << (int(max(counts)) - 1).bit_length()
    _SPARSE_CAP = max(1024, 1 << (2 * int(max(counts))).bit_length())     # room for twice what this job needed
    # (the count words need no masking: read as entries their count field, bits 32.., is zero, and the merge skips those)
    torch.cuda.synchronize(device)
    lo, hi = rank * (cap + 1), (rank + 1) * (cap + 1)
    if lo:
        query.counts_merge_sparse(out.data_ptr(), lo)                      # every rank before this one ...
    if hi < out.numel():
        query.counts_merge_sparse(out.data_ptr() + 8 * hi, out.numel() - hi)   # ... and every rank after it
    query.sync()                                # `out` may be released after this


_SPARSE_CAP = 1 << 18


def query_ci_min(query) -> int:
    return getattr(query, "ci_min", 2)


def bind_to_gpu_numa(local_gpu: int) -> dict:
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, so that the pinned host buffers it
    allocates afterwards (first touch) and its copy-issuing threads are local to that GPU's PCIe root.  With 8
    ranks streaming reads from host memory, remote-node buffers halve the aggregate host->device rate.
    Best effort: returns what it did, never raises."""
    info = {"gpu": local_gpu, "numa_node": None, "cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = local_gpu
        if vis:
            try:
                phys = int(vis.split(",")[local_gpu])
            except (ValueError, IndexError):
                phys = local_gpu
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info.update(numa_node=node, cpus=len(use))
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)
    return info


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
