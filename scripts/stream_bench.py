"""BASELINE.json configs[3] in miniature: a long read set streamed from pinned host memory in batches through
ONE query (mlg_query_push_packed_nruns, two alternating host buffer sets), K1 overlapping the copies.  Reports the
sustained end-to-end rate; the only per-job costs (finish stage, result copy) are paid once at the end.

  python scripts/stream_bench.py            # 8 batches x 10 M reads (80 M reads, 12 Gbases) vs 2e5 genomes
  STREAM_BATCHES=50 ...                     # the full 500 M reads of configs[3] on one GPU
"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth
from metalign_b200 import codec
from metalign_b200.api import Context, Database

KS = (30, 40, 50, 60)
G = int(float(os.environ.get("STREAM_G", "2e5")))
per = int(float(os.environ.get("STREAM_READS", "1e7")))
nb = int(os.environ.get("STREAM_BATCHES", "8"))
L = 150
ctx = Context(0)
p = synth.params(G=G, n=1000, n_present=min(500, G), read_len=L)
d_k = torch.empty(G * 1000 * 2, dtype=torch.int64, device="cuda")
synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None)
db = Database.from_device_keys(ctx, d_k.data_ptr(), G, 1000, 60, KS)
del d_k
nbb, nmb = synth.packed_sizes(per, L)
sets = []
for s in range(2):                      # two different batches of reads, pinned
    d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda"); d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
    synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), s * per, per, d_b.data_ptr(), d_m.data_ptr(), None)
    hb = torch.empty(nbb, dtype=torch.uint8, pin_memory=True); hb.copy_(d_b)
    runs = codec.nmask_to_runs(d_m.cpu().numpy(), per * L)
    hr = torch.empty(max(1, runs.size), dtype=torch.int32, pin_memory=True)
    hr[:runs.size].copy_(torch.from_numpy(runs.reshape(-1).view(np.int32)))
    sets.append((hb, hr, runs.shape[0]))
    del d_b, d_m
h_ci = torch.empty(G * 4, dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    q = db.query()
    for b in range(nb):
        hb, hr, nr = sets[b & 1]
        q.push_packed_nruns_ptr(hb.data_ptr(), hr.data_ptr(), nr, None, per, L)
    ni = q.finish_into(None, None, h_ci.data_ptr())
    dt = time.perf_counter() - t0
    st = q.stats(); q.close()
    print(json.dumps({"reads": per * nb, "batches": nb, "seconds": dt, "Gkmers_s_e2e": st["n_kmers"] / dt / 1e9,
                      "Gbases_s_e2e": per * nb * L / dt / 1e9, "h2d_GB": st["h2d_bytes"] / 1e9, "h2d_GBps": st["h2d_bytes"] / dt / 1e9,
                      "probe_ms_total": st["ms_probe"], "finish_ms": st["ms_query"], "I": ni}), flush=True)
db.close()
