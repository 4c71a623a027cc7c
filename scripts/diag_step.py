"""Diagnostic: where the host-side time of one device-resident bench step goes (per API call, wall clock)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import synth
from metalign_b200.api import Context, Database

KS = (30, 40, 50, 60)
G = int(float(os.environ.get("DIAG_G", "2e5"))); nreads = int(float(os.environ.get("DIAG_READS", "1e7")))
ctx = Context(0)
p = synth.params(G=G, n=1000, n_present=500)
d_k = torch.empty(G * 1000 * 2, dtype=torch.int64, device="cuda")
synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None)
db = Database.from_device_keys(ctx, d_k.data_ptr(), G, 1000, 60, KS); del d_k
nbb, nmb = synth.packed_sizes(nreads, 150)
d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda"); d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, nreads, d_b.data_ptr(), d_m.data_ptr(), None)
h_g = torch.empty(1 << 16, dtype=torch.int32, pin_memory=True); h_ci = torch.empty((1 << 16) * 4, dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize()
acc = {}
def tick(name, t0):
    t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
N = 30
for it in range(N + 5):
    if it == 5: acc.clear(); T0 = time.perf_counter()
    t = time.perf_counter()
    q = db.query(2, "exact", True); t = tick("begin", t)
    q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, 150, device=True); t = tick("push (async)", t)
    ni, rows = q.finish_sparse_into(h_g.data_ptr(), None, None, h_ci.data_ptr(), 1 << 16); t = tick("finish (waits for the GPU)", t)
    st = q.stats(); t = tick("stats", t)
    q.close(); t = tick("close", t)
total = (time.perf_counter() - T0) / N * 1e3
print("per step %.3f ms wall; probe %.3f ms, finish stage %.3f ms" % (total, st["ms_probe"], st["ms_query"]))
for k, v in acc.items(): print("  %-28s %.1f us" % (k, v / N * 1e6))
