"""Experiment: probe-kernel throughput vs database size / bucket format / CTAs per SM (device-resident reads)."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import synth
from metalign_b200.api import Context, Database

KS = (30, 40, 50, 60)
nreads = int(float(os.environ.get("SWEEP_READS", "4e6")))
Gs = [int(float(x)) for x in os.environ.get("SWEEP_G", "2e3,4e3,6e3,8e3,1e4,2e4,5e4,1e5,2e5").split(",")]
ctx = Context(0)
nbb, nmb = synth.packed_sizes(nreads, 150)
d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda"); d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
for G in Gs:
    p = synth.params(G=G, n=1000, n_present=min(500, G))
    d_k = torch.empty(G * 1000 * 2, dtype=torch.int64, device="cuda")
    synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None)
    db = Database.from_device_keys(ctx, d_k.data_ptr(), G, 1000, 60, KS)
    del d_k
    synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, nreads, d_b.data_ptr(), d_m.data_ptr(), None)
    ms = []
    for rep in range(4):
        q = db.query()
        q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, 150, device=True)
        r = q.finish(); st = q.stats(); q.close()
        ms.append(st["ms_probe"])
    t = min(ms[1:])
    print(json.dumps({"G": G, "layout": st["layout"], "fetch_bytes": st["bucket_bytes"], "table_MB": st["n_buckets"] * 32 / 1e6,
                      "probe_ms": t, "Gkmers_s": st["n_kmers"] / t / 1e6, "GBps": st["n_kmers"] * 32 / t / 1e6,
                      "finish_ms": st["ms_query"], "I": st["n_intersect"]}), flush=True)
    db.close()
