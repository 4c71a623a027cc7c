"""Experiment: cost of the multi-GPU seam on ONE GPU (export = clamp touched counters, import = rebuild the
present list from the counter table), without the all-reduce itself."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import synth
from metalign_b200.api import Context, Database

KS = (30, 40, 50, 60)
nreads = int(float(os.environ.get("SWEEP_READS", "1e7")))
G = int(float(os.environ.get("SWEEP_G", "2e5")))
ctx = Context(0)
nbb, nmb = synth.packed_sizes(nreads, 150)
d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda"); d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
p = synth.params(G=G, n=1000, n_present=min(500, G))
d_k = torch.empty(G * 1000 * 2, dtype=torch.int64, device="cuda")
synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None)
db = Database.from_device_keys(ctx, d_k.data_ptr(), G, 1000, 60, KS)
del d_k
synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, nreads, d_b.data_ptr(), d_m.data_ptr(), None)
for mode in ("plain", "seam", "plain", "seam"):
    q = db.query()
    q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, 150, device=True)
    q.sync()
    t0 = time.perf_counter()
    if mode == "seam":
        ptr, n = q.counts_export()
        t1 = time.perf_counter()
        q.counts_import()
    else:
        t1 = t0
    r = q.finish(); st = q.stats(); q.close()
    t2 = time.perf_counter()
    print(json.dumps({"mode": mode, "export_ms": (t1 - t0) * 1e3, "finish_wall_ms": (t2 - t1) * 1e3, "ms_query": st["ms_query"],
                      "ms_probe": st["ms_probe"], "I": st["n_intersect"]}), flush=True)
db.close()
