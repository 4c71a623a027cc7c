#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: read k-mers/s (and Gbases/s) queried against the
CMash-style sketch database, plus the probe kernel's fraction of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one whole pass of the hot path over one batch of synthetic reads: canonical 60-mer counting
against the database (K1), intersection, multi-k prefix expansion (K2), per-genome tables (K3), and -- with
more than one GPU -- the single all-reduce of the counter table.  Workload at N=1 = BASELINE.json configs[1]:
10M x 150 bp reads simulated from database genomes, full default-scale database (2e5 genomes x 1000 slots),
k range 30-60-10.  N>1: weak scaling, 10M reads per GPU, database replicated.

Environment overrides (for quick runs): MLG_BENCH_G, MLG_BENCH_READS (per GPU), MLG_BENCH_CPU_READS,
MLG_BENCH_SKIP_CPU=1.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

KS = (30, 40, 50, 60)
K = 60
READ_LEN = 150
SEED = 20200529


def env_int(name, default):
    v = os.environ.get(name)
    return int(float(v)) if v else default


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled through NVML every few ms by a thread, DURING the
    timed region only (start() / stop() bracket it)."""
    BAD = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index, period_s=0.004):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.stop_flag, self.thread, self.err = [], set(), False, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, attr in self.BAD:
                    if r & getattr(nv, attr):
                        self.reasons.add(name)
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.nv or not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples: %s" % self.err]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "sm_mhz_min": float(min(self.sm))}


def measured_traffic():
    """DRAM bytes per k-mer of the probe kernel from the committed `ncu --set full` capture (profiles/k1_traffic.json,
    written by scripts/ncu_summary.py): dram__bytes_read.sum + dram__bytes_write.sum over the k-mers of that launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def synth_params(G, paired):
    import synth
    return synth.params(G=G, n=1000, K=K, seed=SEED, n_present=500, read_len=READ_LEN, paired=paired)


# ------------------------------------------------------------------------------------------ reference arm
def cpu_baseline_run(p, keys, r0, nreads, threads=0, repeats=1):
    """The CPU restatement of the reference path (oracle/oracle.c) timed on the host cores: R1 counting of a
    bounded read sample + R3-R5.  Database build (the analogue of CMash loading its HDF5/trie) is not timed."""
    import synth
    from oracle.oracle_c import OracleDB, OracleQuery, lib as olib
    if threads:
        olib().orc_set_threads(threads)
    cores = olib().orc_max_threads()
    t0 = time.perf_counter()
    db = OracleDB(keys, p.G, p.n, K, KS)
    t_build = time.perf_counter() - t0
    bases, nmask = synth.reads_packed(p, r0, nreads)
    times, n_kmers, res = [], 0, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        q = OracleQuery(db)
        q.push_packed(bases, nmask, None, nreads, p.read_len)
        res = q.finish()
        times.append(time.perf_counter() - t0)
        n_kmers = res["n_kmers"]
        q.close()
    db.close()
    return dict(times=times, n_kmers=n_kmers, cores=cores, build_s=t_build, n_intersect=res["n_intersect"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    G = env_int("MLG_BENCH_G", 200_000)
    sample = env_int("MLG_BENCH_CPU_READS", 2_000_000)
    p = synth_params(G, 0)
    keys = synth.sketch_keys(p)
    r = cpu_baseline_run(p, keys, 0, sample, repeats=args.warmup + args.steps)
    times = r["times"][args.warmup:]
    sec = float(np.mean(times))
    val = r["n_kmers"] / sec
    line = {
        "impl": "reference", "metric": "read k-mers/sec queried vs CMash DB", "value": val, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u128", "data": "synthetic",
        "gbases_per_s": sample * READ_LEN / sec / 1e9,
        "config": {"workload": "configs[1]: synthetic 150bp reads simulated from database genomes vs %d genomes x 1000 slots, k=30-60-10" % G,
                   "genomes": G, "reads_per_step": sample, "read_len": READ_LEN},
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": r["cores"], "kind": "port",
                         "sample": "%d reads of the workload per step; CPU restatement of the reference path (oracle/oracle.c), "
                                   "not KMC/CMash binaries (absent); database build %.1f s not timed" % (sample, r["build_s"])},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import synth
    from metalign_b200.api import Context, Database
    from metalign_b200 import dist as mdist

    os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    rank, world, local = mdist.init_from_env("nccl")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    import torch.distributed as tdist
    numa = mdist.bind_to_gpu_numa(local) if not os.environ.get("MLG_BENCH_NO_NUMA_BIND") else {"disabled": True}

    G = env_int("MLG_BENCH_G", 200_000)
    reads_per_gpu = env_int("MLG_BENCH_READS", 10_000_000)
    paired = 1 if world > 1 else 0       # configs[2]: paired reads when sharded over several GPUs
    p = synth_params(G, paired)
    ctx = Context(local)
    comp_stream = torch.cuda.ExternalStream(ctx.streams()[0], device=local)

    # database: generated on the device, built on the device
    t0 = time.perf_counter()
    d_keys = torch.empty(G * p.n * 2, dtype=torch.int64, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_keys.data_ptr(), None) == 0
    db = Database.from_device_keys(ctx, d_keys.data_ptr(), G, p.n, K, KS)
    torch.cuda.synchronize()
    t_db = time.perf_counter() - t0
    keys_host = None
    if rank == 0 and world == 1 and not os.environ.get("MLG_BENCH_SKIP_CPU"):
        keys_host = d_keys.cpu().numpy().view(np.uint64).reshape(-1, 2)
    del d_keys
    torch.cuda.empty_cache()

    # this rank's reads, generated on the device; pinned host copy for the end-to-end leg
    r0 = rank * reads_per_gpu
    nbb, nmb = synth.packed_sizes(reads_per_gpu, READ_LEN)
    d_bases = torch.empty(nbb, dtype=torch.uint8, device="cuda")
    d_nmask = torch.empty(nmb, dtype=torch.uint8, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), r0, reads_per_gpu, d_bases.data_ptr(), d_nmask.data_ptr(), None) == 0
    h_bases = torch.empty(nbb, dtype=torch.uint8, pin_memory=True)
    h_nmask = torch.empty(nmb, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases); h_nmask.copy_(d_nmask)
    # the end-to-end leg hands N over as (start, length) runs, the compact form of the same information
    # (mlg_query_push_packed_nruns): 0.1 % of the bases are N, so the mask would be a third of the PCIe bytes
    from metalign_b200 import codec
    runs_np = codec.nmask_to_runs(h_nmask.numpy(), reads_per_gpu * READ_LEN)
    h_runs = torch.empty(max(1, runs_np.size), dtype=torch.int32, pin_memory=True)
    h_runs[:runs_np.size].copy_(torch.from_numpy(runs_np.reshape(-1).view(np.int32)))
    n_runs = runs_np.shape[0]
    nk = len(KS)
    h_num = torch.empty(G * nk, dtype=torch.int64, pin_memory=True)
    h_den = torch.empty(G * nk, dtype=torch.int64, pin_memory=True)
    h_ci = torch.empty(G * nk, dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()

    exch = []      # wall time of the cross-rank exchange per step (joins the probe first, so it includes waiting for it)

    def step(host: bool):
        q = db.query(2, "exact", True)
        if host:
            if os.environ.get("MLG_BENCH_E2E_MASK"):
                q.push_packed_ptr(h_bases.data_ptr(), h_nmask.data_ptr(), None, reads_per_gpu, READ_LEN, device=False)
            else:
                q.push_packed_nruns_ptr(h_bases.data_ptr(), h_runs.data_ptr() if n_runs else None, n_runs, None,
                                        reads_per_gpu, READ_LEN)
        else:
            q.push_packed_ptr(d_bases.data_ptr(), d_nmask.data_ptr(), None, reads_per_gpu, READ_LEN, device=True)
        if world > 1:
            tx = time.perf_counter()
            mdist.reduce_query(q, local)
            exch.append(time.perf_counter() - tx)
        # the step's result = the containment table (what select_db.py consumes); the integer numerators and the
        # (static) denominators stay on the device unless asked for (MLG_BENCH_READBACK_ALL=1)
        if os.environ.get("MLG_BENCH_READBACK_ALL"):
            ni = q.finish_into(h_num.data_ptr(), h_den.data_ptr(), h_ci.data_ptr())
        else:
            ni = q.finish_into(None, None, h_ci.data_ptr())
        st = q.stats()
        q.close()
        return ni, st

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(host: bool, steps: int, warmup: int, sampler=None):
        for _ in range(warmup):
            step(host)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(comp_stream)
        t0 = time.perf_counter()
        stats = []
        for _ in range(steps):
            ni, st = step(host)
            stats.append(st)
        e1.record(comp_stream)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, stats, ni, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, wall_dev, st_dev, ni, clocks = timed(False, args.steps, args.warmup, sampler)
    ms_e2e, wall_e2e, st_e2e, ni2, _ = timed(True, args.steps, args.warmup)
    assert ni == ni2

    kmers_step = st_dev[-1]["n_kmers"]
    if world > 1:
        t = torch.tensor([kmers_step], dtype=torch.int64, device="cuda")
        tdist.all_reduce(t)
        kmers_total = int(t.item())
    else:
        kmers_total = kmers_step
    bases_total = reads_per_gpu * READ_LEN * world
    sec_step = ms_dev / 1e3 / args.steps
    sec_step_e2e = ms_e2e / 1e3 / args.steps
    value = kmers_total / sec_step
    e2e_value = kmers_total / sec_step_e2e

    # roofline of the dominant kernel (K1 probe), from the library's CUDA events on its compute stream
    peak, peak_src = measured_peak()
    probe_ms = float(np.mean([s["ms_probe"] for s in st_dev]))
    nbases = reads_per_gpu * READ_LEN
    bucket_bytes = st_dev[-1]["bucket_bytes"]
    # one 32-byte sector per level-1 bucket fetch + the packed bases and N mask read once.  Layout 0
    # fetches one bucket per k-mer; the super-k-mer layout (1) one per run of windows sharing a minimizer, and
    # the reduced fetch count is reported next to it.
    layout = int(st_dev[-1]["layout"])
    fetches = int(st_dev[-1]["n_bucket_fetches"])
    alg_bytes = fetches * bucket_bytes + nbases // 4 + nbases // 8
    achieved = alg_bytes / (probe_ms / 1e3) / 1e9
    query_ms = float(np.mean([s["ms_query"] for s in st_dev]))
    tr = measured_traffic()
    traffic = tr["dram_bytes_per_kmer"] * kmers_step if tr and G == tr.get("genomes") and tr.get("layout", 0) == int(st_dev[-1]["layout"]) else None
    traffic_src = ("%s: %.2f DRAM bytes per k-mer (ncu --set full, %s) x k-mers per step" % (tr["source"], tr["dram_bytes_per_kmer"], tr["kernel"])
                   if traffic is not None else "no ncu capture for this database size")
    launches = int(sum(s["gpu_launches"] for s in st_dev))

    if layout == 2 and tr and tr.get("layout") == 2:
        limiter = ("instruction issue and load latency at 16 resident warps per SM, not HBM bandwidth: one level-1 access per ~%.0f "
                   "k-mers (DRAM moves %.0f GB/s under ncu, 128 bytes per random access), %.0f warp instructions per 32 k-mers at "
                   "%.0f %% issue-slot utilisation (ncu, %s); see DESIGN.md section 4"
                   % (kmers_step / max(1, fetches), tr.get("dram_gbs_under_ncu", 0), tr.get("warp_instructions_per_32_kmers", 0),
                      tr.get("issue_active_pct", 0), tr.get("source", "")))
    elif layout == 1 and tr and tr.get("layout") == 1:
        limiter = ("instruction issue, not HBM: the super-k-mer kernel fetches one bucket pair per ~17 k-mers, so DRAM runs at "
                   "%.0f GB/s while the SMs issue %.0f warp instructions per 32 k-mers at %.0f %% issue-slot utilisation with %.0f %% "
                   "of the warp slots occupied (ncu, %s); frac is low by construction, see DESIGN.md section 4"
                   % (tr.get("dram_gbs_under_ncu", 0), tr.get("warp_instructions_per_32_kmers", 0), tr.get("issue_active_pct", 0),
                      tr.get("warps_active_pct", 0), tr.get("source", "")))
    else:
        limiter = "random 32-byte DRAM sectors (one per k-mer behind an L2 prefilter)" if layout == 0 else "see profiles/"
    if rank == 0:
        line = {
            "metric": "read k-mers/sec queried vs CMash DB", "value": value, "unit": "k-mers/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u128", "data": "synthetic",
            "gbases_per_s": bases_total / sec_step / 1e9,
            "config": {
                "workload": "configs[%d]: synthetic %dM x %dbp %sreads per GPU simulated from database genomes vs %d genomes x 1000 sketch slots (full default-scale DB), k=30-60-10, ci_min=2, gate=exact"
                            % (2 if world > 1 else 1, reads_per_gpu // 1_000_000, READ_LEN, "paired " if paired else "", G),
                "genomes": G, "sketch_slots": 1000, "reads_per_gpu": reads_per_gpu, "read_len": READ_LEN,
                "db_distinct_kmers": st_dev[-1]["n_db_distinct"], "intersect": ni,
                "parallelism": "reads sharded x%d, DB replicated, 1 uint8 all-reduce of the counter table" % world if world > 1 else "single GPU",
                "l2_policy": "inputs (%.2f GB packed reads) and level-1 table (%.2f GB) both exceed the 126 MB L2; no flush needed"
                             % ((nbb + nmb) / 1e9, (st_dev[-1]["filter_words"] * 4 if layout == 2 else st_dev[-1]["n_buckets"] * (32 if layout == 1 else bucket_bytes)) / 1e9),
                "db_build_s": round(t_db, 3),
            },
            "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": int(st_e2e[-1]["h2d_bytes"]) * world,
                    "d2h_bytes_per_step": int(st_e2e[-1]["d2h_bytes"]) * world, "ms_per_step": sec_step_e2e * 1e3,
                    "gbases_per_s": bases_total / sec_step_e2e / 1e9},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": {2: "k1_minimizer_probe", 1: "k1_superkmer_probe"}.get(layout, "k1_decode_canon_probe"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         # the same fraction with the bytes the DRAM really moves (ncu): this part fetches 128 bytes per random access
                         "dram_frac_from_ncu_traffic": (traffic / (probe_ms / 1e3) / 1e9 / peak) if traffic else None,
                         "algorithmic_bytes_per_launch_set": int(alg_bytes), "kernel_ms_per_step": probe_ms,
                         "finish_stage_ms_per_step": query_ms, "kmers_per_s_kernel_only": kmers_step / (probe_ms / 1e3),
                         "layout": layout, "bucket_fetches_per_step": fetches, "kmers_per_bucket_fetch": kmers_step / max(1, fetches),
                         "algorithmic_bytes_rule": "level-1 fetches x %d B + packed bases + N mask (SURVEY.md 8d, minimizer bucketing: layout 2 = one 32-byte sector of the minimizer-identity bit array per super-k-mer, layout 1 = one 64-byte fingerprint-bucket pair per super-k-mer, layout 0 = one sector per k-mer)" % bucket_bytes,
                         "sector_per_kmer_equivalent_gbs": (kmers_step * 32 + nbases // 4 + nbases // 8) / (probe_ms / 1e3) / 1e9,
                         "limiter": limiter},
            "clocks": clocks,
            "host_numa_binding_rank0": numa,
            "wall_ms_per_step": wall_dev * 1e3 / args.steps,
            "exchange_wall_ms_per_step_incl_probe_join": (float(np.mean(exch[args.warmup:args.warmup + args.steps])) * 1e3 if exch else None),
        }
        if keys_host is not None:
            sample = env_int("MLG_BENCH_CPU_READS", reads_per_gpu)      # the whole workload: ~10 s on 16 cores
            sample = min(sample, reads_per_gpu)
            r = cpu_baseline_run(p, keys_host, 0, sample, repeats=2)
            line["cpu_baseline"] = {
                "value": r["n_kmers"] / float(np.mean(r["times"])), "unit": "k-mers/s", "cores": r["cores"], "kind": "port",
                "sample": "%d reads of the same workload, twice (%.1f s of CPU work in all); CPU restatement of the reference path "
                          "(oracle/oracle.c), not KMC/CMash binaries (absent here); its database build (%.1f s) is not timed"
                          % (sample, float(np.sum(r["times"])), r["build_s"])}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    db.close()
    ctx.close()
    if world > 1:
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
