"""Wall-clock of the drop-in as a user runs it: a FRESH process `python scripts/select_db.py reads.fq data/`, from process
start to the three output files, database load included (the built .mlgdb form, scripts/build_db.py).  Set-up (untimed):
a 10 M x 150 bp FASTQ of the bench workload, the 2e5-genome database saved in built form, db_info.txt, and organism files
for the genomes the reads hit.

  python scripts/cold_select_main.py            # RF_G=2e5 RF_READS=1e7 RF_DIR=/tmp/mlg_cold RF_THREADS=16
"""
import ctypes as C, gzip, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import synth
from metalign_b200 import select_db
from metalign_b200.api import Context, Database

KS = (30, 40, 50, 60)
G = int(float(os.environ.get("RF_G", "2e5")))
nreads = int(float(os.environ.get("RF_READS", "1e7")))
base = os.environ.get("RF_DIR", "/tmp/mlg_cold")
threads = int(os.environ.get("RF_THREADS", str(os.cpu_count())))
L = 150
data = os.path.join(base, "data")
os.makedirs(os.path.join(data, "organism_files"), exist_ok=True)
reads_path = os.path.join(base, "reads.fq")
p = synth.params(G=G, n=1000, n_present=min(500, G), read_len=L)
if not os.path.exists(reads_path):
    with open(reads_path, "wb") as f:
        step = 500_000
        for a in range(0, nreads, step):
            m = min(step, nreads - a)
            arr = synth.reads_ascii(p, a, m)
            rec = np.empty((m, 2 * L + 10), dtype=np.uint8)
            rec[:, 0] = ord("@"); rec[:, 1:5] = np.frombuffer(b"read", dtype=np.uint8); rec[:, 5] = 10
            rec[:, 6:6 + L] = arr; rec[:, 6 + L] = 10; rec[:, 7 + L] = ord("+"); rec[:, 8 + L] = 10
            rec[:, 9 + L:9 + 2 * L] = ord("I"); rec[:, 9 + 2 * L] = 10
            f.write(rec.tobytes())
names = ["taxid_%d_%d_genomic.fna.gz" % (100000 + g // 5, 1 + g % 5) for g in range(G)]
assert names == sorted(names)
db_path = os.path.join(data, select_db.DB_BASENAME)
t_build = t_save = 0.0
with Context(0) as ctx:
    d_k = torch.empty(G * 1000 * 2, dtype=torch.int64, device="cuda")
    synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None)
    t0 = time.perf_counter()
    db = Database.from_device_keys(ctx, d_k.data_ptr(), G, 1000, 60, KS, names=names)
    t_build = time.perf_counter() - t0
    del d_k
    t0 = time.perf_counter()
    db.save(db_path)
    t_save = time.perf_counter() - t0
    # which genomes will be selected: only their organism files are opened
    nbb, nmb = synth.packed_sizes(nreads, L)
    d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda"); d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
    synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, nreads, d_b.data_ptr(), d_m.data_ptr(), None)
    q = db.query()
    q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, L, device=True)
    res = q.finish(); q.close(); db.close()
    hit = np.flatnonzero(res["ci"][:, -1] > 0)
    del d_b, d_m
with open(os.path.join(data, "db_info.txt"), "w") as f:
    f.write("Accession\tLength\tTaxID\tLineage\tTaxID_Lineage\n")
    for g, nm in enumerate(names):
        taxid = select_db.taxid_of(nm)
        f.write("ACC%07d.1\t3000000\t%s\tn|n|n|n|n|n|n|n\t2|1224|1236|91347|543|561|%d|%s\n" % (g, taxid, 100000 + g // 5, taxid))
for g in hit:
    fp = os.path.join(data, "organism_files", names[g])
    if not os.path.exists(fp):
        with gzip.open(fp, "wt") as gz:
            gz.write(">ACC%07d.1 genome %d\n%s\n" % (g, g, "ACGT" * 500))
torch.cuda.synchronize()
for rep in range(int(os.environ.get("RF_REPEATS", "3"))):
    out = os.path.join(base, "out%d" % rep)
    t0 = time.perf_counter()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "select_db.py"), reads_path, data, "--temp_dir", out,
                           "--threads", str(threads)], env=dict(os.environ, MLG_TIMING="1"))
    wall = time.perf_counter() - t0
    sel = sum(1 for _ in open(os.path.join(out, "subset_db_info.txt"))) - 2
    print(json.dumps({"wall_s_process_start_to_files": wall, "reads": nreads, "reads_file_GB": os.path.getsize(reads_path) / 1e9,
                      "db_file_GB": os.path.getsize(db_path) / 1e9, "genomes": G, "selected_accessions": sel, "threads": threads,
                      "db_build_s_once": t_build, "db_save_s_once": t_save, "genomes_with_k60_hits": int(hit.size)}), flush=True)
