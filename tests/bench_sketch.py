#! /usr/bin/env python
"""Sketch-builder throughput: random genomes -> bottom-1000 MinHash sketches (k = 60) on one GPU, next to
the CPU restatement of CMash's CountEstimator (oracle/sketch_oracle.c, all host threads) on a bounded sample.
    python tests/bench_sketch.py [--genomes 64] [--mbp 3] [--cpu_genomes 4]
Prints one JSON line.  The kernel is instruction-bound (MurmurHash3 = two 64-bit multiplies per 8 bytes), so the figure
is windows hashed per second; each base is read from HBM once."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import numpy as np  # noqa: E402

from metalign_b200.api import Context  # noqa: E402
from metalign_b200.sketch import sketch_genomes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=64)
    ap.add_argument("--mbp", type=float, default=3.0)
    ap.add_argument("--cpu_genomes", type=int, default=4)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    L = int(a.mbp * 1e6)
    genomes = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L, dtype=np.uint8)].tobytes() for _ in range(a.genomes)]
    with Context(0) as ctx:
        sketch_genomes(ctx, genomes[:2], 1000, 60)                       # warm-up
        t0 = time.perf_counter()
        mins, counts, kmers, st = sketch_genomes(ctx, genomes, 1000, 60)
        wall = time.perf_counter() - t0
    from oracle import sketch_oracle as so
    t0 = time.perf_counter()
    om, oc, ok = so.sketch_genomes(genomes[:a.cpu_genomes], 1000, 60)
    cpu = time.perf_counter() - t0
    same = bool(np.array_equal(mins[:a.cpu_genomes], om) and np.array_equal(kmers[:a.cpu_genomes], ok))
    cpu_windows = a.cpu_genomes * (L - 59)
    print(json.dumps({"metric": "genome k-mers sketched / s (bottom-1000 MinHash, k=60, MurmurHash3 mod prime)",
                      "genomes": a.genomes, "mbp_per_genome": a.mbp, "windows": st["n_windows"], "candidates": st["n_candidates"],
                      "kernel_ms": st["ms_kernels"], "gpu_windows_per_s_kernels": st["n_windows"] / (st["ms_kernels"] / 1e3),
                      "gpu_windows_per_s_call": st["n_windows"] / wall, "call_s": wall,
                      "cpu_windows_per_s": cpu_windows / cpu, "cpu_cores": os.cpu_count(), "cpu_sample_genomes": a.cpu_genomes,
                      "matches_cpu_restatement": same}))


if __name__ == "__main__":
    main()
