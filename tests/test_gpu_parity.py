"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.
Bit-exact for counts / denominators / the intersection set; containment compared with == as well
(both sides do one IEEE double division of identical integers; the north-star tolerance is 1e-12)."""
import ctypes as C
import os
import random

import numpy as np
import pytest

import synth
from metalign_b200 import codec, dbformat
from metalign_b200.api import Database, MlgError
from oracle import oracle_py

from helpers import adversarial_case, oracle_c_run

pytestmark = pytest.mark.gpu
KS = (30, 40, 50, 60)


def _check(res, I_gpu, ref, I_ref, tag=""):
    assert res["n_kmers"] == ref["n_kmers"], tag
    assert res["n_intersect"] == ref["n_intersect"], tag
    assert np.array_equal(I_gpu, I_ref), tag
    assert np.array_equal(res["den"], ref["den"]), tag
    assert np.array_equal(res["num"], ref["num"]), tag
    assert np.array_equal(res["ci"], ref["ci"]), tag
    assert np.max(np.abs(res["ci"] - ref["ci"]), initial=0.0) <= 1e-12, tag


def test_golden_cases(ctx, golden_cases):
    for c in golden_cases:
        db = Database.from_sketches(ctx, c["sketches"], c["K"], c["ks"])
        for e in c["expect"]:
            q = db.query(e["ci_min"], e["gate"], e["count_empty_in_den"])
            q.push_reads(c["reads"])
            r = q.finish()
            I = [codec.key_to_kmer(a, b, c["K"]) for a, b in q.intersection()]
            q.close()
            assert r["num"].tolist() == e["num"], (c["name"], e["gate"])
            assert r["den"].tolist() == e["den"], (c["name"], e["gate"])
            assert I == e["I"], (c["name"], e["gate"])
        db.close()


@pytest.mark.parametrize("seed", range(60))
def test_adversarial_vs_oracle(ctx, seed):
    rng = random.Random(1000 + seed)
    c = adversarial_case(rng)
    G, n = len(c["sketches"]), len(c["sketches"][0])
    keys = codec.sketches_to_keys(c["sketches"], c["K"])
    db = Database.from_keys(ctx, keys, G, n, c["K"], c["ks"])
    for gate in ("exact", "none"):
        ci_min = rng.choice([1, 2, 3])
        ce = rng.random() < 0.5
        ref, I_ref = oracle_c_run(keys, G, n, c["K"], c["ks"], lambda q: q.push_reads(c["reads"]), ci_min, gate, ce)
        q = db.query(ci_min, gate, ce)
        q.push_reads(c["reads"])
        res = q.finish()
        _check(res, q.intersection(), ref, I_ref, (seed, gate))
        q.close()
        py = oracle_py.run(c["reads"], c["sketches"], K=c["K"], ks=c["ks"], ci_min=ci_min, gate=gate, count_empty_in_den=ce)
        assert res["num"].tolist() == py["num"]
    db.close()


@pytest.fixture(scope="module")
def workload(ctx):
    p = synth.params(G=300, n=200, seed=11, len_min=20000, len_max=60000, n_present=40)
    keys = synth.sketch_keys(p)
    nreads = 60000
    bases, nmask = synth.reads_packed(p, 0, nreads)
    refs = {}
    for gate in ("exact", "none"):
        refs[gate] = oracle_c_run(keys, p.G, p.n, 60, KS, lambda q: q.push_packed(bases, nmask, None, nreads, p.read_len), 2, gate, True)
    db = Database.from_keys(ctx, keys, p.G, p.n, 60, KS)
    yield dict(p=p, keys=keys, nreads=nreads, bases=bases, nmask=nmask, refs=refs, db=db)
    db.close()


def test_synthetic_packed_host_fixed_len(workload):
    w = workload
    assert w["refs"]["exact"][0]["n_intersect"] > 100
    for gate in ("exact", "none"):
        q = w["db"].query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"][gate], tag=gate)
        assert res["stats"]["gpu_launches"] >= 3 and res["stats"]["ms_probe"] > 0
        q.close()


def test_synthetic_packed_host_offsets_and_split_batches(workload):
    w = workload
    L = w["p"].read_len
    reads = codec.unpack_reads(w["bases"], w["nmask"], None, w["nreads"], L)
    q = w["db"].query()
    cut = [0, 1, 7000, 7001, 31000, w["nreads"]]
    for a, b in zip(cut[:-1], cut[1:]):
        bs, nm, off = codec.pack_reads(reads[a:b])
        q.push_packed(bs, nm, off, b - a)
        q.sync()
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()


def test_synthetic_ascii_push(workload):
    w = workload
    text = synth.reads_ascii(w["p"], 0, w["nreads"]).reshape(-1)
    off = np.arange(w["nreads"] + 1, dtype=np.uint64) * np.uint64(w["p"].read_len)
    q = w["db"].query()
    q.push_ascii(text, off)
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()


def test_no_nmask_means_no_n(workload):
    w = workload
    ref, I_ref = oracle_c_run(w["keys"], w["p"].G, w["p"].n, 60, KS,
                              lambda q: q.push_packed(w["bases"], None, None, w["nreads"], w["p"].read_len))
    q = w["db"].query()
    q.push_packed(w["bases"], None, None, w["nreads"], w["p"].read_len)
    res = q.finish()
    _check(res, q.intersection(), ref, I_ref)
    q.close()


def test_device_generator_and_device_push(ctx, workload):
    import torch
    w = workload
    p = w["p"]
    nbb, nmb = synth.packed_sizes(w["nreads"], p.read_len)
    d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda")
    d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, w["nreads"], d_b.data_ptr(), d_m.data_ptr(), None) == 0
    assert np.array_equal(d_b.cpu().numpy(), w["bases"]) and np.array_equal(d_m.cpu().numpy(), w["nmask"])
    d_k = torch.empty(p.G * p.n * 2, dtype=torch.int64, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_k.data_ptr(), None) == 0
    assert np.array_equal(d_k.cpu().numpy().view(np.uint64).reshape(-1, 2), w["keys"])
    db2 = Database.from_device_keys(ctx, d_k.data_ptr(), p.G, p.n, 60, KS)
    assert db2.n_distinct == w["db"].n_distinct and db2.n_entries == w["db"].n_entries
    q = db2.query()
    q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, w["nreads"], p.read_len, device=True)
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()
    # device offsets
    d_off = torch.arange(w["nreads"] + 1, dtype=torch.int64, device="cuda") * p.read_len
    q = db2.query()
    q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), d_off.data_ptr(), w["nreads"], 0, device=True)
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()
    db2.close()


def test_counter_export_import_equals_single_run(workload):
    """the multi-GPU seam on one GPU: two half queries, counters summed, presence derived from the sum"""
    import torch
    w = workload
    L = w["p"].read_len
    reads = codec.unpack_reads(w["bases"], w["nmask"], None, w["nreads"], L)
    halves = []
    for part in (reads[0::2], reads[1::2]):
        q = w["db"].query()
        bs, nm, off = codec.pack_reads(part)
        q.push_packed(bs, nm, off, len(part))
        halves.append(q)
    from metalign_b200.dist import device_view_u8
    views = [device_view_u8(*q.counts_export()) for q in halves]
    assert int((views[0].to(torch.int32) + views[1].to(torch.int32)).max()) <= 4
    views[0] += views[1]            # what the all-reduce does, on one GPU
    torch.cuda.synchronize()
    halves[0].counts_import()
    res = halves[0].finish()
    ref, I_ref = w["refs"]["exact"]
    assert np.array_equal(res["num"], ref["num"]) and res["n_intersect"] == ref["n_intersect"]
    assert np.array_equal(halves[0].intersection(), I_ref)
    with pytest.raises(MlgError):
        halves[0].push_reads(["ACGT"])
    for q in halves:
        q.close()


def test_mlgdb_roundtrip(ctx, workload, tmp_path):
    w = workload
    p = w["p"]
    names = ["taxid_%d_genomic.fna.gz" % g for g in range(p.G)]
    path = str(tmp_path / "db.mlgdb")
    dbformat.write(path, w["keys"], names, p.G, p.n, 60, KS)
    assert np.array_equal(dbformat.read_keys(path), w["keys"])
    db = Database.load(ctx, path)
    assert db.names == names and (db.G, db.n, db.K, db.ks) == (p.G, p.n, 60, KS)
    q = db.query()
    q.push_packed(w["bases"], w["nmask"], None, w["nreads"], p.read_len)
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    assert np.array_equal(db.denominators(True), w["refs"]["exact"][0]["den"])
    q.close()
    db.close()


def test_sparse_result_and_table_reuse(ctx, workload):
    """finish_sparse: the rows of the genomes with a hit at any k, equal to the dense table's non-zero rows -- also with
    room for fewer rows than there are (the call is repeated), and for several queries in a row on one database (the
    counter table is handed on, cleared entry by entry, instead of being zeroed as a whole)"""
    w = workload
    p = w["p"]
    db = Database.from_keys(ctx, w["keys"], p.G, p.n, 60, KS)
    for rep, (gate, cap) in enumerate((("exact", 1 << 16), ("none", 3), ("exact", 1 << 16))):
        ref, I_ref = w["refs"][gate]
        q = db.query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], p.read_len)
        res = q.finish_sparse(cap)
        rows = np.flatnonzero(ref["num"].sum(axis=1) > 0)
        assert rows.size > 3 and np.array_equal(res["genomes"], rows.astype(np.uint32)), (rep, gate)
        assert np.array_equal(res["num"], ref["num"][rows]) and np.array_equal(res["den"], ref["den"][rows])
        assert np.array_equal(res["ci"], ref["ci"][rows]) and res["n_intersect"] == ref["n_intersect"]
        assert np.array_equal(q.intersection(), I_ref)
        q.close()
    db.close()


def test_built_database_file(ctx, workload, tmp_path, monkeypatch):
    """the built form of a database (mlg_db_save: the device structures themselves) loads to the same answers for both
    gates, keeps the names, refuses a file of another build tag, and cannot be written from a database that kept P"""
    w = workload
    p = w["p"]
    names = ["taxid_%d_genomic.fna.gz" % g for g in range(p.G)]
    src = str(tmp_path / "src.mlgdb")
    dbformat.write(src, w["keys"], names, p.G, p.n, 60, KS)
    db = Database.load(ctx, src)
    built = str(tmp_path / "built.mlgdb")
    db.save(built)
    db.close()
    assert dbformat.read_header(built)["built"] and dbformat.read_names(built) == names
    with pytest.raises(ValueError):
        dbformat.read_keys(built)
    db2 = Database.load(ctx, built)
    assert db2.names == names and (db2.G, db2.n, db2.K, db2.ks) == (p.G, p.n, 60, KS)
    assert np.array_equal(db2.denominators(True), w["refs"]["exact"][0]["den"])
    for gate in ("exact", "none"):
        q = db2.query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], p.read_len)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"][gate], tag=("built", gate))
        q.close()
    db2.close()
    raw = bytearray(open(built, "rb").read())
    head = (68 + len("\n".join(names).encode()) + 15) // 16 * 16
    raw[head] ^= 0xFF                                   # the build tag
    bad = str(tmp_path / "stale.mlgdb")
    open(bad, "wb").write(raw)
    with pytest.raises(MlgError):
        Database.load(ctx, bad)
    open(bad, "wb").write(raw[: len(raw) // 2])          # truncated
    with pytest.raises(MlgError):
        Database.load(ctx, bad)
    monkeypatch.setenv("MLG_KEEP_P", "1")
    db3 = Database.from_keys(ctx, w["keys"], p.G, p.n, 60, KS, names=names)
    with pytest.raises(MlgError):
        db3.save(str(tmp_path / "no.mlgdb"))
    db3.close()


def test_sixteen_byte_buckets(ctx, workload, monkeypatch):
    w = workload
    monkeypatch.setenv("MLG_BUCKET_SLOTS", "4")
    monkeypatch.setenv("MLG_LAYOUT", "0")          # 16-byte buckets exist in the whole-k-mer hash layout only
    monkeypatch.setenv("MLG_BUCKET_LOAD", "3.0")   # forces many overflowing buckets through the exact path
    db = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    q = db.query()
    q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
    res = q.finish()
    assert res["stats"]["bucket_bytes"] == 16
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()
    db.close()


def test_overfull_32_byte_buckets(ctx, workload, monkeypatch):
    w = workload
    monkeypatch.setenv("MLG_BUCKET_LOAD", "7.5")
    db = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    q = db.query(2, "none", False)
    q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
    res = q.finish()
    ref, I_ref = oracle_c_run(w["keys"], w["p"].G, w["p"].n, 60, KS,
                              lambda oq: oq.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len), 2, "none", False)
    _check(res, q.intersection(), ref, I_ref)
    q.close()
    db.close()


def test_other_k_and_ranges(ctx):
    """K=31 (key fits one word), K=45 with a 3-value k range, K=63 (largest)."""
    for K, ks, seed in ((31, (15, 21, 31), 3), (45, (20, 33, 45), 4), (63, (31, 47, 63), 5), (60, (60,), 6)):
        p = synth.params(G=40, n=50, K=K, seed=seed, len_min=5000, len_max=9000, n_present=10, read_len=100)
        keys = synth.sketch_keys(p)
        nreads = 20000
        bases, nmask = synth.reads_packed(p, 0, nreads)
        db = Database.from_keys(ctx, keys, p.G, p.n, K, ks)
        for gate in ("exact", "none"):
            ref, I_ref = oracle_c_run(keys, p.G, p.n, K, ks, lambda q: q.push_packed(bases, nmask, None, nreads, 100), 2, gate, True)
            q = db.query(2, gate, True)
            q.push_packed(bases, nmask, None, nreads, 100)
            res = q.finish()
            _check(res, q.intersection(), ref, I_ref, (K, gate))
            assert ref["n_intersect"] > 0
            q.close()
        db.close()


def test_empty_and_degenerate_inputs(ctx):
    p = synth.params(G=5, n=8, seed=2, len_min=3000, len_max=4000, n_present=3)
    keys = synth.sketch_keys(p)
    db = Database.from_keys(ctx, keys, p.G, p.n, 60, KS)
    q = db.query()
    r = q.finish()          # no reads at all
    assert r["num"].sum() == 0 and r["n_kmers"] == 0 and r["n_intersect"] == 0 and q.intersection().shape == (0, 2)
    q.close()
    q = db.query()
    q.push_reads(["", "ACGT", "N" * 200, "acgtn" * 30])   # nothing long enough / valid
    q.push_reads([])
    r = q.finish()
    assert r["n_intersect"] == 0 and r["n_kmers"] == oracle_py.run(["", "ACGT", "N" * 200, "acgtn" * 30], [[]], K=60)["n_kmers"]
    q.close()
    db.close()
    # a database whose slots are all empty
    empty = np.full((6, 2), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    db = Database.from_keys(ctx, empty, 2, 3, 60, KS)
    assert db.n_distinct == 0 and db.n_entries == 0
    q = db.query()
    q.push_reads(["ACGT" * 40])
    r = q.finish()
    assert r["num"].sum() == 0 and r["den"].tolist() == [[1] * 4, [1] * 4] and r["n_kmers"] == 101
    q.close()
    db.close()
    with pytest.raises(MlgError):
        Database.from_keys(ctx, keys, p.G, p.n, 60, (40, 30))   # not ascending
    with pytest.raises(MlgError):
        Database.from_sketches(ctx, [["ACGN" * 15]], 60, KS)     # non-ACGT sketch k-mer


def test_midsize_against_oracle(ctx):
    """1M reads x 150 against 2000 genomes x 1000 slots: full table compared with the C oracle."""
    p = synth.params(G=2000, n=1000, seed=20200529, n_present=60)
    keys = synth.sketch_keys(p)
    nreads = 1_000_000
    bases, nmask = synth.reads_packed(p, 0, nreads)
    ref, I_ref = oracle_c_run(keys, p.G, p.n, 60, KS, lambda q: q.push_packed(bases, nmask, None, nreads, p.read_len))
    db = Database.from_keys(ctx, keys, p.G, p.n, 60, KS)
    q = db.query()
    q.push_packed(bases, nmask, None, nreads, p.read_len)
    res = q.finish()
    _check(res, q.intersection(), ref, I_ref)
    assert res["n_intersect"] > 1000
    q.close()
    db.close()


@pytest.mark.parametrize("filter_mb", ["0", "0.016", "64"])
def test_prefilter_variants(ctx, workload, monkeypatch, filter_mb):
    """no prefilter, a starved prefilter (~2 bits per key, most probes pass) and the default one"""
    w = workload
    monkeypatch.setenv("MLG_FILTER_MB", filter_mb)
    monkeypatch.setenv("MLG_LAYOUT", "0")       # the prefilter belongs to the whole-k-mer-hash layout
    db = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    q = db.query()
    q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
    res = q.finish()
    assert (res["stats"]["filter_words"] == 0) == (filter_mb == "0") and res["stats"]["layout"] == 0
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()
    db.close()


def test_deep_coverage_many_hits(ctx):
    """every read comes from two small genomes: most windows near sketch positions hit, the exact-path queue
    is drained many times and counters saturate"""
    p = synth.params(G=6, n=400, seed=9, len_min=3000, len_max=3500, n_present=2, strain_period=0, tiny_pct=0,
                     sub_per_64k=0, n_per_64k=0)
    keys = synth.sketch_keys(p)
    nreads = 120000
    bases, nmask = synth.reads_packed(p, 0, nreads)
    ref, I_ref = oracle_c_run(keys, p.G, p.n, 60, KS, lambda q: q.push_packed(bases, nmask, None, nreads, p.read_len))
    db = Database.from_keys(ctx, keys, p.G, p.n, 60, KS)
    q = db.query()
    q.push_packed(bases, nmask, None, nreads, p.read_len)
    res = q.finish()
    _check(res, q.intersection(), ref, I_ref)
    assert res["n_intersect"] >= 300
    q.close()
    db.close()


def test_nruns_push_equals_mask_push(ctx, workload):
    """N given as (start, length) runs (mlg_query_push_packed_nruns) == N given as a bit mask, in several
    batches and with offsets; a batch without any N takes the mask-free kernel."""
    w = workload
    p, nreads = w["p"], w["nreads"]
    nb = nreads * p.read_len
    runs = codec.nmask_to_runs(w["nmask"], nb)
    assert runs.shape[0] > 10 and np.array_equal(codec.runs_to_nmask(runs, nb)[: (nb + 7) // 8], w["nmask"][: (nb + 7) // 8])
    q = w["db"].query()
    q.push_packed_nruns(w["bases"], runs, None, nreads, p.read_len)
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()
    # small copy chunks: runs are sliced per chunk, some straddle chunk boundaries
    import os
    os.environ["MLG_CHUNK_MB"] = "0.25"
    try:
        q = w["db"].query()
        off = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(p.read_len)
        q.push_packed_nruns(w["bases"], runs, off, nreads, 0)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"]["exact"])
        q.close()
    finally:
        del os.environ["MLG_CHUNK_MB"]


def test_nruns_long_runs_and_no_n(ctx):
    """runs longer than a read, runs touching batch start / end, and nruns=None"""
    rng = random.Random(77)
    K = 60
    genome = "".join(rng.choice("ACGT") for _ in range(3000))
    sketches = [[genome[i:i + K] for i in range(0, 2400, 100)]]
    reads = []
    for i in range(400):
        a = rng.randint(0, len(genome) - 200)
        r = list(genome[a:a + 200])
        if i % 3 == 0:
            s = rng.randint(0, 199); l = rng.randint(1, 260)
            for j in range(s, min(200, s + l)):
                r[j] = "N"
        reads.append("".join(r))
    reads[0] = "N" * 200
    reads[-1] = reads[-1][:150] + "N" * 50
    keys = codec.sketches_to_keys(sketches, K)
    ref, I_ref = oracle_c_run(keys, 1, len(sketches[0]), K, KS, lambda q: q.push_reads(reads + reads))
    bases, nmask, off = codec.pack_reads(reads)
    runs = codec.nmask_to_runs(nmask, int(off[-1]))
    assert runs[:, 1].max() > 150
    db = Database.from_keys(ctx, keys, 1, len(sketches[0]), K, KS)
    q = db.query()
    q.push_packed_nruns(bases, runs, off, len(reads))
    q.push_packed(bases, nmask, off, len(reads))
    res = q.finish()
    _check(res, q.intersection(), ref, I_ref)
    q.close()
    clean = [r.replace("N", "A") for r in reads]
    ref2, I2 = oracle_c_run(keys, 1, len(sketches[0]), K, KS, lambda q: q.push_reads(clean))
    b2, m2, off2 = codec.pack_reads(clean)
    q = db.query(1)
    ref2, I2 = oracle_c_run(keys, 1, len(sketches[0]), K, KS, lambda q: q.push_reads(clean), 1)
    q.push_packed_nruns(b2, None, off2, len(clean))
    res = q.finish()
    _check(res, q.intersection(), ref2, I2)
    q.close()
    db.close()


def test_all_layouts_k60(ctx, workload, monkeypatch):
    """K = 60 defaults to the minimizer-bitmap layout (one bitmap sector per super-k-mer, ~5 per 150-base read);
    MLG_LAYOUT=1 keeps the fingerprint-pair super-k-mer layout, MLG_LAYOUT=0 the whole-k-mer hash layout.  Same
    results, and the fetch counts tell them apart."""
    w = workload
    q = w["db"].query()
    q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
    res = q.finish()
    _check(res, q.intersection(), *w["refs"]["exact"])
    st = res["stats"]
    q.close()
    assert st["layout"] == 2 and st["filter_words"] >= 1 << 15 and st["bucket_bytes"] == 32
    assert 0.05 * st["n_kmers"] < st["n_bucket_fetches"] < 0.14 * st["n_kmers"]      # ~8 runs of 32-base minimizers per 91 windows
    monkeypatch.setenv("MLG_LAYOUT", "1")
    db1 = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    for gate in ("exact", "none"):
        q = db1.query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"][gate])
        st = res["stats"]
        assert st["layout"] == 1 and st["filter_words"] == 0 and st["bucket_bytes"] == 64
        assert 0.03 * st["n_kmers"] < st["n_bucket_fetches"] < 0.09 * st["n_kmers"]
        q.close()
    db1.close()
    monkeypatch.setenv("MLG_LAYOUT", "0")
    db0 = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    for gate in ("exact", "none"):
        q = db0.query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"][gate])
        assert res["stats"]["layout"] == 0 and res["stats"]["n_bucket_fetches"] == res["stats"]["n_kmers"]
        q.close()
    db0.close()


@pytest.mark.parametrize("layout", [2, 1])
def test_superkmer_adversarial_minimizers(ctx, monkeypatch, layout):
    """both minimizer layouts; reads built to stress the super-k-mer path: low-complexity sequence (every window shares one minimizer),
    tandem repeats (the minimizer recurs), strictly decreasing minimizers are approximated by random sequence with
    many N (segments restart), reads of 60..400 bases (several 96-window segments per lane), and a database whose
    sketch k-mers overlap heavily (many keys per minimizer -> overflowing buckets -> exact path)."""
    rng = random.Random(2024)
    K = 60
    genome = "".join(rng.choice("ACGT") for _ in range(4000))
    lowc = ("A" * 70 + "ACACACACAC" * 8 + "T" * 70 + genome[:300]) * 2
    sketches = [[genome[i:i + K] for i in range(0, 600)],                         # 600 overlapping k-mers: shared minimizers
                [lowc[i:i + K] for i in range(0, 600)],
                [oracle_py.rc(genome[i:i + K]) for i in range(1000, 1600)]]
    reads = []
    for src in (genome, lowc):
        for _ in range(300):
            L = rng.choice([60, 61, 75, 150, 155, 156, 157, 251, 400])
            a = rng.randint(0, len(src) - L)
            r = src[a:a + L]
            if rng.random() < 0.5:
                r = oracle_py.rc(r)
            r = list(r)
            if rng.random() < 0.3:
                for _ in range(rng.randint(1, 4)):
                    r[rng.randint(0, L - 1)] = "N"
            reads.append("".join(r))
    reads += ["A" * 200, "ACGT" * 60, "T" * 59, "", "N" * 100, genome[:60]]
    keys = codec.sketches_to_keys(sketches, K)
    monkeypatch.setenv("MLG_LAYOUT", str(layout))
    db = Database.from_keys(ctx, keys, 3, 600, K, KS)
    for ci_min, gate in ((1, "exact"), (2, "none"), (3, "exact")):
        ref, I_ref = oracle_c_run(keys, 3, 600, K, KS, lambda q: q.push_reads(reads), ci_min, gate, True)
        q = db.query(ci_min, gate, True)
        q.push_reads(reads)
        res = q.finish()
        _check(res, q.intersection(), ref, I_ref, (ci_min, gate))
        assert res["stats"]["layout"] == layout
        q.close()
    assert ref["n_intersect"] > 500
    db.close()


@pytest.mark.parametrize("layout,load", [(2, None), (2, "64"), (1, None)])
def test_every_window_is_a_database_kmer(ctx, monkeypatch, layout, load):
    """the sketches hold EVERY 60-mer of a 20 kb sequence and the reads come from it, so every N-free window must be
    found exactly once: any window lost or counted twice by the run bookkeeping of the minimizer kernels (runs
    spanning blocks, more than four runs in a block of 16 windows, segments of long reads, crowded buckets with
    load = 64 keys per bucket) changes the counters"""
    rng = random.Random(77)
    K, G, n = 60, 20, 1000
    genome = "".join(rng.choice("ACGT") for _ in range(G * n + K - 1))
    sketches = [[genome[g * n + j:g * n + j + K] for j in range(n)] for g in range(G)]
    reads = []
    for _ in range(4000):
        L = rng.choice([60, 100, 150, 150, 150, 151, 250, 330])
        a = rng.randint(0, len(genome) - L)
        r = genome[a:a + L]
        if rng.random() < 0.5:
            r = oracle_py.rc(r)
        if rng.random() < 0.2:
            r = list(r)
            r[rng.randint(0, L - 1)] = "N"
            r = "".join(r)
        reads.append(r)
    keys = codec.sketches_to_keys(sketches, K)
    monkeypatch.setenv("MLG_LAYOUT", str(layout))
    if load:
        monkeypatch.setenv("MLG_MZ_LOAD", load)
    db = Database.from_keys(ctx, keys, G, n, K, KS)
    for ci_min in (1, 30):
        ref, I_ref = oracle_c_run(keys, G, n, K, KS, lambda q: q.push_reads(reads), ci_min, "exact", True)
        q = db.query(ci_min, "exact", True)
        q.push_reads(reads)
        res = q.finish()
        _check(res, q.intersection(), ref, I_ref, (layout, load, ci_min))
        assert res["stats"]["layout"] == layout
        q.close()
        if ci_min == 1:
            assert ref["n_intersect"] > 0.95 * G * n
    db.close()


def _mz_order(s32):
    """kmer.cuh mz_order of a 32-mer given as a string"""
    a = int("".join(str("ACGT".index(ch)) for ch in s32[:16]), 4)
    b = int("".join(str("ACGT".index(ch)) for ch in oracle_py.rc(s32[16:])), 4)
    return ((a + b) * 0x9E3779B1) & 0x03FFFFFF


def test_minimizer_ties_use_the_alias_table(ctx, monkeypatch):
    """database 60-mers whose smallest 32-mer order is reached at two positions with DIFFERENT content: a read carrying
    such a k-mer forward finds its leftmost minimum, a read carrying the reverse complement its rightmost one; the second
    identity is served by the alias table (kmer.cuh).  The order only looks at the 26 bases in the middle of a 32-mer,
    so planting the same low-order core twice, 28 bases apart, makes the tie."""
    rng = random.Random(31)
    K = 60
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    tied = []
    while len(tied) < 40:
        core = rnd(26)
        s = rnd(3) + core + rnd(2) + core + rnd(3)
        assert len(s) == K
        o = [_mz_order(s[p:p + 32]) for p in range(K - 31)]
        if o[0] == o[28] == min(o) and s[0:32] != s[28:60] and o.count(min(o)) == 2:
            tied.append(s)
    plain = [rnd(K) for _ in range(60)]
    sketches = [tied[:20] + plain[:30], [oracle_py.rc(x) for x in tied[20:]] + plain[30:]]
    reads = []
    for x in tied + plain:
        for strand in (0, 1):
            r = rnd(rng.randint(0, 40)) + (x if strand == 0 else oracle_py.rc(x)) + rnd(rng.randint(0, 40))
            reads += [r] * rng.randint(1, 3)
    keys = codec.sketches_to_keys(sketches, K)
    monkeypatch.setenv("MLG_LAYOUT", "2")
    db = Database.from_keys(ctx, keys, 2, 50, K, KS)
    for ci_min in (1, 2, 4):
        ref, I_ref = oracle_c_run(keys, 2, 50, K, KS, lambda q: q.push_reads(reads), ci_min, "exact", True)
        q = db.query(ci_min, "exact", True)
        q.push_reads(reads)
        res = q.finish()
        _check(res, q.intersection(), ref, I_ref, ci_min)
        assert res["stats"]["layout"] == 2
        q.close()
        if ci_min == 1:
            assert ref["n_intersect"] == 100
    db.close()


@pytest.mark.parametrize("mode", ["off", "tiny_buffer", "no_room"])
def test_hit_list_fallbacks(ctx, workload, monkeypatch, mode):
    """the per-k-mer hit lists are precomputed at database build; without them (MLG_PRECOMPUTE_HITS=0), for the k-mers
    that did not fit the buffer, or when the records do not fit the device beside P (what happens at 2e9 slots; here the
    free-memory figure is overridden), the same expansion runs on the fly at query time"""
    w = workload
    if mode == "off":
        monkeypatch.setenv("MLG_PRECOMPUTE_HITS", "0")
    elif mode == "no_room":
        monkeypatch.setenv("MLG_HIT_BUDGET_BYTES", "4096")
    else:
        monkeypatch.setenv("MLG_HIT_CAP_WORDS", "4000")      # room for a few hundred of the ~60000 k-mers
    db = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    for gate in ("exact", "none"):
        q = db.query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"][gate], tag=(mode, gate))
        q.close()
    db.close()


def test_database_built_from_chunks(ctx, workload):
    """mlg_db_builder_*: the database handed over a few genomes at a time (device buffers, reused) answers like the one
    built from all keys at once; genomes out of order, a gap, or a missing tail are refused"""
    import torch
    w = workload
    p = w["p"]
    keys = torch.from_numpy(w["keys"].view(np.int64).reshape(-1)).cuda()
    step = max(1, p.G // 7)
    buf = torch.empty(step * p.n * 2, dtype=torch.int64, device="cuda")

    def chunks(stop=p.G, first=0):
        for g0 in range(first, stop, step):
            c = min(step, stop - g0)
            buf[: c * p.n * 2].copy_(keys[g0 * p.n * 2:(g0 + c) * p.n * 2])
            torch.cuda.synchronize()
            yield buf.data_ptr(), g0, c
    db = Database.from_device_chunks(ctx, chunks(), p.G, p.n, 60, KS)
    for gate in ("exact", "none"):
        q = db.query(2, gate, True)
        q.push_packed(w["bases"], w["nmask"], None, w["nreads"], p.read_len)
        res = q.finish()
        _check(res, q.intersection(), *w["refs"][gate], tag=("chunks", gate))
        q.close()
    db.close()
    with pytest.raises(MlgError):
        Database.from_device_chunks(ctx, chunks(stop=p.G - 1), p.G, p.n, 60, KS)         # a genome short
    with pytest.raises(MlgError):
        Database.from_device_chunks(ctx, chunks(first=step), p.G, p.n, 60, KS)           # does not start at genome 0


def test_full_size_properties(ctx):
    """BASELINE.json configs[1] at full size (2e5 genomes x 1000 slots, 10 M x 150 bp reads), where the oracle is too
    slow to be the checker: size-independent properties of the path instead.
      (a) batching invariance: one push == the same reads in three pushes (host N-runs path and device path);
      (b) counting: the k-mers seen >= 2 times when every read is pushed twice == the k-mers seen >= 1 time;
      (c) strand symmetry: reverse-complementing every read changes nothing (canonical counting);
      (d) the per-genome numerators never exceed the denominators, and only genomes the reads were simulated
          from (or strains sharing their sketches) collect more than a handful of k=60 hits."""
    import torch
    G, n, nreads, L = 200_000, 1000, 10_000_000, 150
    p = synth.params(G=G, n=n, seed=20200529, n_present=500, read_len=L)
    d_keys = torch.empty(G * n * 2, dtype=torch.int64, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_keys.data_ptr(), None) == 0
    db = Database.from_device_keys(ctx, d_keys.data_ptr(), G, n, 60, KS)
    del d_keys
    nbb, nmb = synth.packed_sizes(nreads, L)
    d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda")
    d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, nreads, d_b.data_ptr(), d_m.data_ptr(), None) == 0
    torch.cuda.synchronize()

    def run(pushes, ci_min=2):
        q = db.query(ci_min, "exact", True)
        for f in pushes:
            f(q)
        r = q.finish()
        I = q.intersection()
        q.close()
        return r, I

    whole, I_whole = run([lambda q: q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, L, device=True)])
    assert whole["n_kmers"] > 0.9 * nreads * (L - 59) and whole["n_intersect"] > 10000
    # (a) three host pushes with N as runs; cut at read boundaries that are multiples of 64 bases (16-byte units)
    hb, hm = d_b.cpu().numpy(), d_m.cpu().numpy()
    cuts = [0, 3_200_000, 3_200_000 + 4_000_064, nreads]
    assert all((c * L) % 64 == 0 for c in cuts[:-1])

    def host_push(a, b):
        nb = (b - a) * L
        sb = np.ascontiguousarray(hb[a * L // 4: a * L // 4 + ((nb + 63) // 64) * 16 + 16])
        sm = np.ascontiguousarray(hm[a * L // 8: a * L // 8 + ((nb + 63) // 64) * 8 + 16])
        runs = codec.nmask_to_runs(sm, nb)
        return lambda q: q.push_packed_nruns(sb, runs, None, b - a, L)

    parts, I_parts = run([host_push(cuts[i], cuts[i + 1]) for i in range(3)])
    for k in ("num", "den", "ci"):
        assert np.array_equal(parts[k], whole[k]), k
    assert parts["n_kmers"] == whole["n_kmers"] and np.array_equal(I_parts, I_whole)
    # (b) every read twice, threshold 2 == every read once, threshold 1
    dev = lambda q: q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, L, device=True)
    twice, I_twice = run([dev, dev], ci_min=2)
    once1, I_once1 = run([dev], ci_min=1)
    assert np.array_equal(twice["num"], once1["num"]) and np.array_equal(I_twice, I_once1)
    assert twice["n_kmers"] == 2 * whole["n_kmers"]
    # (c) strand symmetry on the first 1 M reads
    m = 1_000_000
    reads = codec.unpack_reads(hb[: m * L // 4 + 16], hm[: m * L // 8 + 16], None, m, L)
    comp = str.maketrans("ACGTN", "TGCAN")
    rc_reads = [r.translate(comp)[::-1] for r in reads]
    b1, m1, o1 = codec.pack_reads(reads)
    b2, m2, o2 = codec.pack_reads(rc_reads)
    fwd, I_f = run([lambda q: q.push_packed(b1, m1, o1, m)])
    rev, I_r = run([lambda q: q.push_packed(b2, m2, o2, m)])
    assert np.array_equal(fwd["num"], rev["num"]) and np.array_equal(I_f, I_r) and fwd["n_kmers"] == rev["n_kmers"]
    # (d) sanity of the table
    assert (whole["num"] <= whole["den"]).all() and (whole["den"][:, -1] >= 1).all()
    high = np.flatnonzero(whole["num"][:, -1] >= 10)         # genomes with real support at k = 60
    assert 0 < high.size <= 500 * 3
    db.close()


def test_ten_x_database_on_one_gpu(ctx):
    """BASELINE.json configs[4]'s database at FULL size -- 2e6 genomes x 1000 slots (2e9 slots, 1.6e9 distinct k-mers) --
    built through the chunk-fed builder on one GPU.  The oracle cannot hold this database, so the check is a round trip
    that exercises the high end of every index: the sketches of 48 genomes spread over the whole range (the last genome
    included) become reads -- every sketch k-mer twice, in a random orientation, with random flanks -- and
      (a) the intersection is exactly the canonical forms of those k-mers,
      (b) for the sampled genomes num / den / ci at every k equal what the C oracle computes from the database that
          holds those 48 genomes alone (other genomes cannot take hits away: forward and reverse-complement prefixes of
          unrelated random sequence do not coincide),
      (c) a genome that is not sampled, and no strain of a sampled one, collects nothing at k = 60."""
    import torch
    G, n = 2_000_000, 1000
    free_b, total_b = torch.cuda.mem_get_info()
    if total_b < 150 * (1 << 30):
        pytest.skip("needs a 180 GB device")
    p = synth.params(G=G, n=n, seed=20200529, n_present=500)
    step = 100_000
    d_chunk = torch.empty(step * n * 2, dtype=torch.int64, device="cuda")

    def chunks():
        for g0 in range(0, G, step):
            c = min(step, G - g0)
            assert synth.cuda_lib().syn_cuda_gen_sketch_keys_range(C.byref(p), g0, c, d_chunk.data_ptr(), None) == 0
            yield d_chunk.data_ptr(), g0, c
    db = Database.from_device_chunks(ctx, chunks(), G, n, 60, KS)
    rng = random.Random(4)
    sample = sorted({0, 1, G // 2, G - 2, G - 1} | {rng.randrange(G) for _ in range(43)})
    sub = np.empty((len(sample) * n, 2), dtype=np.uint64)
    for i, g in enumerate(sample):
        assert synth.cuda_lib().syn_cuda_gen_sketch_keys_range(C.byref(p), g, 1, d_chunk.data_ptr(), None) == 0
        sub[i * n:(i + 1) * n] = d_chunk[: n * 2].cpu().numpy().view(np.uint64).reshape(-1, 2)
    del d_chunk
    reads = []
    for hi, lo in sub:
        if hi == codec.EMPTY:
            continue
        x = codec.key_to_kmer(int(hi), int(lo), 60)
        for _ in range(2):
            y = oracle_py.rc(x) if rng.random() < 0.5 else x
            reads.append("".join(rng.choice("ACGT") for _ in range(rng.randint(0, 40))) + y + "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 40))))
    rng.shuffle(reads)
    q = db.query()
    q.push_reads(reads)
    res = q.finish()
    I_gpu = q.intersection()
    q.close()
    db.close()
    ref, I_ref = oracle_c_run(sub, len(sample), n, 60, KS, lambda oq: oq.push_reads(reads))
    assert np.array_equal(I_gpu, I_ref) and len(I_ref) > 40_000, (len(I_gpu), len(I_ref))
    idx = np.asarray(sample)
    assert np.array_equal(res["den"][idx], ref["den"])
    assert np.array_equal(res["num"][idx], ref["num"]) and np.array_equal(res["ci"][idx], ref["ci"])
    assert (ref["num"][:, -1] > 0).all()
    hit60 = set(np.nonzero(res["num"][:, -1])[0].tolist())
    # strains share sketch positions with their parent genome (synth_core.h): a hit outside the sample is a relative of a sampled genome
    related = set()
    for g in sample:
        related.update(range(max(0, g - p.strain_period), min(G, g + p.strain_period + 1)))
    assert set(sample) <= hit60 and hit60 <= related, sorted(hit60 - related)[:10]


def test_headline_config_against_oracle(ctx):
    """BASELINE.json configs[1] at FULL size -- 2e5 genomes x 1000 slots, 10 M x 150 bp reads, the workload bench.py
    quotes its headline on -- with the complete per-genome tables and the intersection compared with the C oracle
    (about a minute of host time: the oracle's database build dominates)."""
    import torch
    from oracle.oracle_c import OracleDB, OracleQuery, lib as olib
    olib().orc_set_threads(len(os.sched_getaffinity(0)))
    G, n, nreads, L = 200_000, 1000, 10_000_000, 150
    p = synth.params(G=G, n=n, seed=20200529, n_present=500, read_len=L)
    d_keys = torch.empty(G * n * 2, dtype=torch.int64, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_sketch_keys(C.byref(p), d_keys.data_ptr(), None) == 0
    keys = d_keys.cpu().numpy().view(np.uint64).reshape(-1, 2)
    db = Database.from_device_keys(ctx, d_keys.data_ptr(), G, n, 60, KS)
    del d_keys
    nbb, nmb = synth.packed_sizes(nreads, L)
    d_b = torch.empty(nbb, dtype=torch.uint8, device="cuda")
    d_m = torch.empty(nmb, dtype=torch.uint8, device="cuda")
    assert synth.cuda_lib().syn_cuda_gen_reads_packed(C.byref(p), 0, nreads, d_b.data_ptr(), d_m.data_ptr(), None) == 0
    torch.cuda.synchronize()
    q = db.query()
    q.push_packed_ptr(d_b.data_ptr(), d_m.data_ptr(), None, nreads, L, device=True)
    res = q.finish()
    I_gpu = q.intersection()
    q.close()
    db.close()
    bases, nmask = d_b.cpu().numpy(), d_m.cpu().numpy()
    del d_b, d_m
    torch.cuda.empty_cache()
    odb = OracleDB(keys, G, n, 60, KS)
    oq = OracleQuery(odb)
    oq.push_packed(bases, nmask, None, nreads, L)
    ref = oq.finish()
    I_ref = oq.intersection()
    oq.close()
    odb.close()
    assert ref["n_intersect"] > 10000 and (ref["num"][:, -1] > 0).sum() > 100
    _check(res, I_gpu, ref, I_ref, "configs[1] full size")


def test_superkmer_overfull_pairs(ctx, workload, monkeypatch):
    """mean load 7.5 per 8-slot half: most halves are full or overflowed, so most windows go through the queue and the
    exact compare (a full half is treated as overflowed; there is no flag bit in this layout)"""
    w = workload
    monkeypatch.setenv("MLG_LAYOUT", "1")
    monkeypatch.setenv("MLG_SK_LOAD", "7.5")
    db = Database.from_keys(ctx, w["keys"], w["p"].G, w["p"].n, 60, KS)
    q = db.query()
    q.push_packed(w["bases"], w["nmask"], None, w["nreads"], w["p"].read_len)
    res = q.finish()
    assert res["stats"]["layout"] == 1
    _check(res, q.intersection(), *w["refs"]["exact"])
    q.close()
    db.close()
