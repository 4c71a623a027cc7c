"""Diagnostic: host->device bandwidth seen by torch and by the library's push path."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import synth
from metalign_b200.api import Context, Database

n = 256 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("torch pinned H2D %.1f GB/s" % (n / dt / 1e9))
hp = torch.empty(n, dtype=torch.uint8)
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize()
print("torch pageable H2D %.1f GB/s" % (n / (time.perf_counter() - t) / 1e9))
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t = time.perf_counter(); h2.copy_(d); torch.cuda.synchronize()
print("torch pinned D2H %.1f GB/s" % (n / (time.perf_counter() - t) / 1e9))

ctx = Context(0)
p = synth.params(G=2000, n=1000, n_present=100)
db = Database.from_keys(ctx, synth.sketch_keys(p), p.G, p.n)
nreads = 2_000_000
nbb, nmb = synth.packed_sizes(nreads, 150)
b, m = synth.reads_packed(p, 0, nreads)
hb = torch.empty(nbb, dtype=torch.uint8, pin_memory=True); hm = torch.empty(nmb, dtype=torch.uint8, pin_memory=True)
hb.copy_(torch.from_numpy(b)); hm.copy_(torch.from_numpy(m))
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    q = db.query(); t1 = time.perf_counter()
    q.push_packed_ptr(hb.data_ptr(), hm.data_ptr(), None, nreads, 150); t2 = time.perf_counter()
    q.sync(); t3 = time.perf_counter()
    r = q.finish(); t4 = time.perf_counter()
    q.close(); t5 = time.perf_counter()
    print("begin %.2f push %.2f sync %.2f finish %.2f close %.2f ms; probe %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, r["stats"]["ms_probe"]))
