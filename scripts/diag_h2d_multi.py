#!/usr/bin/env python
"""Isolated host->device bandwidth of N ranks copying at once, no kernels: what bounds the end-to-end leg of bench.py at
N > 1.  Every rank pins one buffer the size of a 10 M-read batch (387 MB) and copies it to its
GPU `reps` times, all ranks between the same two barriers; reported per rank and in aggregate, for one copy per step and
for the 16 MiB chunks the library's push path uses.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/diag_h2d_multi.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from metalign_b200 import dist as mdist
    numa = mdist.bind_to_gpu_numa(local) if not os.environ.get("MLG_NO_NUMA_BIND") else {"disabled": True}
    nbytes = int(os.environ.get("DIAG_BYTES", str(387_072_656)))
    reps = int(os.environ.get("DIAG_REPS", "10"))
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.fill_(rank + 1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {"rank": rank, "world": world, "bytes": nbytes, "numa": numa, "cpus": len(os.sched_getaffinity(0))}
    for name, chunk in (("whole", nbytes), ("chunks_16MiB", 16 << 20)):
        for timed in (False, True):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps if timed else 2):
                for o in range(0, nbytes, chunk):
                    d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        out[name] = {"gbs_rank": nbytes * reps / ms / 1e6, "ms_per_copy": ms / reps, "wall_ms_per_copy_incl_barrier": wall * 1e3 / reps}
    if world > 1:
        t = torch.tensor([out["whole"]["ms_per_copy"], out["chunks_16MiB"]["ms_per_copy"]], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        if rank == 0:
            agg = {"world": world, "aggregate_gbs_whole": nbytes * world / float(tmax[0]) / 1e6,
                   "aggregate_gbs_chunks": nbytes * world / float(tmax[1]) / 1e6,
                   "per_rank_gbs_whole": [round(g["whole"]["gbs_rank"], 1) for g in gathered],
                   "per_rank_gbs_chunks": [round(g["chunks_16MiB"]["gbs_rank"], 1) for g in gathered],
                   "numa": [g["numa"] for g in gathered], "cpus_visible": out["cpus"]}
            print(json.dumps(agg))
        dist.destroy_process_group()
    else:
        print(json.dumps({"world": 1, "aggregate_gbs_whole": out["whole"]["gbs_rank"], "aggregate_gbs_chunks": out["chunks_16MiB"]["gbs_rank"],
                          "numa": numa, "cpus_visible": out["cpus"]}))


if __name__ == "__main__":
    main()
