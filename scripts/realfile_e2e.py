"""Experiment: the drop-in select_db path on a real FASTQ file of the bench workload's shape (10 M x 150 bp), timed
stage by stage: native ingest -> pinned batches -> GPU -> containment table.  Shows that once the hot path takes
milliseconds, parsing the file is what a run waits for."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth
from metalign_b200 import ingest
from metalign_b200.api import Context, Database, pinned_array

KS = (30, 40, 50, 60)
G = int(float(os.environ.get("RF_G", "2e5")))
nreads = int(float(os.environ.get("RF_READS", "1e7")))
path = os.environ.get("RF_PATH", "/tmp/rf_reads.fq")
L = 150
p = synth.params(G=G, n=1000, n_present=min(500, G), read_len=L)
t0 = time.perf_counter()
if not os.path.exists(path):
    with open(path, "wb") as f:
        step = 500_000
        for a in range(0, nreads, step):
            m = min(step, nreads - a)
            arr = synth.reads_ascii(p, a, m)
            rec = np.empty((m, 2 * L + 10), dtype=np.uint8)
            rec[:, 0] = ord("@"); rec[:, 1:5] = np.frombuffer(b"read", dtype=np.uint8); rec[:, 5] = 10
            rec[:, 6:6 + L] = arr; rec[:, 6 + L] = 10; rec[:, 7 + L] = ord("+"); rec[:, 8 + L] = 10
            rec[:, 9 + L:9 + 2 * L] = ord("I"); rec[:, 9 + 2 * L] = 10
            f.write(rec.tobytes())
t_gen = time.perf_counter() - t0
ctx = Context(0)
d_k = torch.empty(G * 1000 * 2, dtype=torch.int64, device="cuda")
synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None)
t0 = time.perf_counter()
db = Database.from_device_keys(ctx, d_k.data_ptr(), G, 1000, 60, KS)
t_db = time.perf_counter() - t0
del d_k
for threads in (int(os.environ.get("RF_THREADS", "0")),) * int(os.environ.get("RF_REPEATS", "2")):
    t0 = time.perf_counter()
    rd = ingest.PackedBatches(path, "fastq", reads_per_batch=2_000_000, threads=threads, alloc=pinned_array)
    t_open = time.perf_counter() - t0
    q = db.query()
    t_ing = t_push = 0.0
    n = 0
    while True:
        a = time.perf_counter()
        try:
            bases, runs, off, m = next(rd)
        except StopIteration:
            break
        b = time.perf_counter()
        q.push_packed_nruns(bases, runs if len(runs) else None, off, m)
        c = time.perf_counter()
        t_ing += b - a; t_push += c - b; n += m
    a = time.perf_counter()
    res = q.finish()
    t_fin = time.perf_counter() - a
    total = time.perf_counter() - t0
    st = res["stats"]; q.close(); rd.close()
    print(json.dumps({"reads": n, "file_GB": os.path.getsize(path) / 1e9, "total_s": total, "ingest_wait_s": t_ing, "push_call_s": t_push,
                      "finish_s": t_fin, "alloc_open_s": t_open, "kmers": st["n_kmers"], "Gkmers_s_file_to_table": st["n_kmers"] / total / 1e9,
                      "file_MBps": os.path.getsize(path) / total / 1e6, "probe_ms": st["ms_probe"], "I": st["n_intersect"],
                      "db_build_s": t_db, "gen_s": t_gen, "host_threads": os.cpu_count()}), flush=True)
db.close()
