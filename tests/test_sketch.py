"""Sketch builder: CMash MakeStreamingDNADatabase.py semantics (local_tests/retrain_and_test_metalign.sh:49).
CPU: the C oracle against MurmurHash3's published vectors and against a literal Python transcription of
MinHash.CountEstimator.add(); GPU: mlg_sketch_genomes bit-exact against the oracle."""
import bisect
import gzip
import os
import random

import numpy as np
import pytest

from oracle import sketch_oracle as so

PRIME = so.PRIME
M64 = (1 << 64) - 1


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _fmix(k):
    k ^= k >> 33
    k = (k * 0xff51afd7ed558ccd) & M64
    k ^= k >> 33
    k = (k * 0xc4ceb9fe1a85ec53) & M64
    return k ^ (k >> 33)


def py_murmur3_x64_128(data: bytes, seed: int = 0):
    """MurmurHash3_x64_128, straight from the published algorithm"""
    c1, c2 = 0x87c37b91114253d5, 0x4cf5ad432745937f
    h1 = h2 = seed
    n = len(data)
    for i in range(n // 16):
        k1 = int.from_bytes(data[16 * i:16 * i + 8], "little")
        k2 = int.from_bytes(data[16 * i + 8:16 * i + 16], "little")
        k1 = (k1 * c1) & M64; k1 = _rotl(k1, 31); k1 = (k1 * c2) & M64; h1 ^= k1
        h1 = _rotl(h1, 27); h1 = (h1 + h2) & M64; h1 = (h1 * 5 + 0x52dce729) & M64
        k2 = (k2 * c2) & M64; k2 = _rotl(k2, 33); k2 = (k2 * c1) & M64; h2 ^= k2
        h2 = _rotl(h2, 31); h2 = (h2 + h1) & M64; h2 = (h2 * 5 + 0x38495ab5) & M64
    tail = data[16 * (n // 16):]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & M64; k2 = _rotl(k2, 33); k2 = (k2 * c1) & M64; h2 ^= k2
    if tail:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & M64; k1 = _rotl(k1, 31); k1 = (k1 * c2) & M64; h1 ^= k1
    h1 ^= n; h2 ^= n
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    h1 = _fmix(h1); h2 = _fmix(h2)
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    return h1, h2


def py_count_estimator(seq: str, n: int, K: int, prime: int = PRIME):
    """CMash MinHash.CountEstimator(n, ksize=K, save_kmers='y', rev_comp=False): add_sequence + add, transcribed"""
    import re
    mins, counts, kmers = [prime] * n, [0] * n, [""] * n
    for piece in re.compile("[^ACTG]").split(seq.upper()):
        for i in range(len(piece) - K + 1):
            kmer = piece[i:i + K]
            h = py_murmur3_x64_128(kmer.encode())[0] % prime
            if h >= mins[-1]:
                continue
            j = bisect.bisect_left(mins, h)
            if mins[j] == h:
                counts[j] += 1
            else:
                mins.insert(j, h); mins.pop()
                counts.insert(j, 1); counts.pop()
                kmers.insert(j, kmer); kmers.pop()
    return mins, counts, kmers


def test_murmur3_published_vectors():
    """the three vectors quoted wherever MurmurHash3_x64_128 is documented (seed 0)"""
    vec = {b"": "00000000000000000000000000000000", b"hello": "cbd8a7b341bd9b025b1e906a48ae1d19",
           b"The quick brown fox jumps over the lazy dog": "e34bbc7bbc071b6c7a433ca9c49a9347"}
    rng = random.Random(1)
    for data, want in vec.items():
        assert "%016x%016x" % so.murmur3_x64_128(data) == want == "%016x%016x" % py_murmur3_x64_128(data)
    for _ in range(300):
        data = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 80)))
        seed = rng.choice([0, 1, 0xDEADBEEF])
        assert so.murmur3_x64_128(data, seed) == py_murmur3_x64_128(data, seed)


def _genomes(rng):
    rnd = lambda L: "".join(rng.choice("ACGT") for _ in range(L))
    g = [rnd(3000), rnd(59), rnd(60), "", rnd(700) + "N" + rnd(61) + "nn" + rnd(30) + "R" + rnd(400),
         rnd(500).lower() + rnd(500), ("ACGTTGCAAGGCT" * 60), rnd(200) * 6, "A" * 300]
    u = rnd(800)
    g.append(u + "N" + u[100:500] + ">" + u[::-1])          # repeated k-mers across records
    return g


@pytest.mark.parametrize("n,K", [(50, 60), (8, 21), (200, 32)])
def test_oracle_matches_python_transcription(n, K):
    rng = random.Random(100 + n)
    genomes = _genomes(rng)
    mins, counts, kmers = so.sketch_genomes(genomes, n, K)
    for g, seq in enumerate(genomes):
        pm, pc, pk = py_count_estimator(seq, n, K)
        assert mins[g].tolist() == pm and counts[g].tolist() == pc
        assert [bytes(x).rstrip(b"\0").decode() for x in kmers[g]] == pk
    assert (mins[3] == PRIME).all() and (counts[8] <= 300).all()


def test_last_element_is_not_counted_while_it_is_the_maximum():
    """CountEstimator.add() returns on `h >= mins[-1]`, so repeats of the sketch's largest element are not counted once the
    sketch is full; the oracle (and the GPU path) keep that quirk"""
    rng = random.Random(12)
    u = "".join(rng.choice("ACGT") for _ in range(400))
    seq = u + "N" + u + "N" + u                       # every 60-mer occurs three times
    n = 40
    mins, counts, _ = so.sketch_genomes([seq], n, 60)
    pm, pc, _ = py_count_estimator(seq, n, 60)
    assert counts[0].tolist() == pc and mins[0].tolist() == pm
    assert set(counts[0][:-1].tolist()) == {3} and counts[0][-1] < 3


def test_read_fasta(tmp_path):
    from metalign_b200.sketch import read_fasta
    p = tmp_path / "g.fna"
    p.write_text(">c1 desc\nACGT\nacgt\n>c2\nTTTT\n\n>empty\n>c3\nGG\n")
    assert read_fasta(str(p)) == b"ACGTacgtNTTTTNGG"
    with gzip.open(str(p) + ".gz", "wb") as f:
        f.write(p.read_bytes())
    assert read_fasta(str(p) + ".gz") == b"ACGTacgtNTTTTNGG"


def test_build_database_host_logic(tmp_path, monkeypatch):
    """build_database's host side (sorted-basename order, batching, empty slots, the .mlgdb it writes) with the device call
    replaced by the oracle -- the GPU test below runs the real thing"""
    from metalign_b200 import codec, dbformat, sketch
    rng = random.Random(8)
    paths = []
    for i, L in enumerate((5000, 40, 9000, 700, 3000)):            # 40 bases: a genome without a single 60-mer
        p = tmp_path / ("g%d_%s.fna%s" % (9 - i, "x" * i, ".gz" if i % 2 else ""))
        seq = "".join(rng.choice("ACGT") for _ in range(L))
        with (gzip.open if i % 2 else open)(str(p), "wb") as f:
            f.write((">a\n%s\n>b\n%s\n" % (seq[:L // 2], seq[L // 2:])).encode())
        paths.append(str(p))
    calls = []

    def fake(ctx, genomes, n, K, prime=0):
        calls.append(len(genomes))
        m, c, k = so.sketch_genomes(genomes, n, K)
        return m, c, k, {"n_windows": 0, "n_candidates": 0, "ms_kernels": 0.0, "passes": 1}

    monkeypatch.setattr(sketch, "sketch_genomes", fake)
    out = str(tmp_path / "db.mlgdb")
    sketch.build_database(None, paths, out, n=300, K=60, batch_bytes=6000)
    order = sorted(paths, key=os.path.basename)
    assert len(calls) >= 2 and sum(calls) == len(paths)
    assert dbformat.read_names(out) == [os.path.basename(p) for p in order]
    h = dbformat.read_header(out)
    assert (h["G"], h["n"], h["K"], list(h["ks"])) == (5, 300, 60, [30, 40, 50, 60])
    _, _, ok = so.sketch_genomes([sketch.read_fasta(p) for p in order], 300, 60)
    keys = dbformat.read_keys(out).reshape(-1, 2)
    assert np.array_equal(keys, codec.ascii_slots_to_keys(ok.reshape(-1, 60), 60))
    empty = (keys[:, 0] == np.uint64(0xFFFFFFFFFFFFFFFF)).reshape(5, 300).sum(axis=1)
    assert empty.max() == 300 and empty.min() == 0                # the 40-base genome has no k-mer at all


# ------------------------------------------------------------------------------------------------ GPU
def _check(ctx, genomes, n, K, prime=0):
    from metalign_b200.sketch import sketch_genomes
    mins, counts, kmers, st = sketch_genomes(ctx, genomes, n, K, prime)
    om, oc, ok = so.sketch_genomes(genomes, n, K, prime or PRIME)
    assert np.array_equal(mins, om) and np.array_equal(counts, oc) and np.array_equal(kmers, ok)
    return st


@pytest.mark.gpu
@pytest.mark.parametrize("n,K", [(50, 60), (8, 21), (200, 32), (64, 64), (30, 16), (5, 1)])
def test_gpu_sketch_matches_oracle_small(ctx, n, K):
    rng = random.Random(7 + K)
    st = _check(ctx, _genomes(rng), n, K)
    assert st["n_windows"] > 0 and st["passes"] >= 1


@pytest.mark.gpu
def test_gpu_sketch_matches_oracle_large_and_thresholds(ctx):
    """genomes of 0.2-2 Mbp with n = 1000 (the threshold keeps ~4n of up to 2e6 windows), a tandem-repeat genome whose
    distinct k-mers are far fewer than its windows (the threshold has to be widened: extra passes), 0.5 % non-ACGT
    symbols, and a custom prime"""
    rng = np.random.default_rng(5)
    genomes = []
    for L in (200_000, 2_000_000, 1_000, 50_000):
        a = rng.integers(0, 4, L, dtype=np.uint8)
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[a].copy()
        s[rng.random(L) < 0.005] = ord("N")
        s[rng.random(L) < 0.01] += 32                      # some lower case (N -> n stays invalid)
        genomes.append(s.tobytes())
    unit = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 1500, dtype=np.uint8)].tobytes()
    genomes.append(unit * 200)                              # 300 kb, 1500 distinct 60-mers
    st = _check(ctx, genomes, 1000, 60)
    assert st["passes"] >= 2 and st["n_candidates"] < 0.2 * st["n_windows"]
    _check(ctx, genomes[:2], 100, 60, prime=1000003)


@pytest.mark.gpu
def test_gpu_sketch_to_database_and_query(ctx, tmp_path):
    """scripts/make_sketch_db.py end to end: FASTA files -> .mlgdb -> reads drawn from one genome light up that genome"""
    from metalign_b200 import codec, dbformat
    from metalign_b200.api import Database
    from metalign_b200.sketch import build_database, read_fasta
    rng = random.Random(3)
    paths = []
    for i in range(4):
        p = tmp_path / ("genome_%d.fna%s" % (3 - i, ".gz" if i % 2 else ""))
        recs = ["".join(rng.choice("ACGT") for _ in range(rng.randint(3000, 9000))) for _ in range(3)]
        data = "".join(">c%d\n%s\n" % (j, "\n".join(r[k:k + 70] for k in range(0, len(r), 70))) for j, r in enumerate(recs))
        with (gzip.open if i % 2 else open)(str(p), "wb") as f:
            f.write(data.encode())
        paths.append(str(p))
    out = str(tmp_path / "db.mlgdb")
    build_database(ctx, paths, out, n=200, K=60)
    order = sorted(paths, key=os.path.basename)
    assert dbformat.read_names(out) == [os.path.basename(p) for p in order]
    _, _, ok = so.sketch_genomes([read_fasta(p) for p in order], 200, 60)
    assert np.array_equal(dbformat.read_keys(out).reshape(-1, 2), codec.ascii_slots_to_keys(ok.reshape(-1, 60), 60))
    db = Database.load(ctx, out)
    src = read_fasta(order[2]).decode()
    reads = [src[a:a + 150] for a in range(0, len(src) - 150, 7)]
    from helpers import oracle_c_run
    for gate in ("exact", "none"):
        ref, _ = oracle_c_run(dbformat.read_keys(out), 4, 200, 60, (30, 40, 50, 60), lambda q: q.push_reads(reads), 1, gate, True)
        q = db.query(1, gate, True)
        q.push_reads(reads)
        res = q.finish()
        q.close()
        assert np.array_equal(res["num"], ref["num"]) and np.array_equal(res["den"], ref["den"]), (gate, res["num"], ref["num"])
    db.close()
    # sketch k-mers are stored as they occur on the forward strand: without the prefilter gate every one of them is found,
    # with it only the canonical half survives at k = 60
    ci = res["ci"][:, -1]
    assert ci[2] > 0.9 and ci[[0, 1, 3]].max() < 0.05, (res["num"], res["den"], res["ci"], ref["ci"])
