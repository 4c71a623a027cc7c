"""Host codec round trips (no GPU): packed stream, N mask, N runs."""
import random

import numpy as np

from metalign_b200 import codec


def test_pack_unpack_round_trip():
    rng = random.Random(3)
    reads = ["".join(rng.choice("ACGTN") for _ in range(rng.randint(0, 90))) for _ in range(50)]
    bases, nmask, off = codec.pack_reads(reads)
    assert bases.size % 16 == 0 and nmask.size % 16 == 0
    assert codec.unpack_reads(bases, nmask, off) == reads


def test_nmask_runs_round_trip():
    rng = random.Random(4)
    for _ in range(20):
        n = rng.randint(1, 700)
        isn = np.zeros(n, dtype=np.uint8)
        for _ in range(rng.randint(0, 12)):
            a = rng.randint(0, n - 1)
            isn[a:a + rng.randint(1, 70)] = 1
        mask = np.packbits(isn)
        runs = codec.nmask_to_runs(mask, n)
        assert runs.dtype == np.uint32 and runs.shape[1] == 2
        # sorted, non-overlapping, non-adjacent, inside the stream
        ends = runs[:, 0].astype(np.int64) + runs[:, 1]
        assert (runs[:, 1] > 0).all() and (ends <= n).all()
        assert (runs[1:, 0].astype(np.int64) > ends[:-1]).all()
        assert int(runs[:, 1].sum()) == int(isn.sum())
        back = codec.runs_to_nmask(runs, n)
        assert np.array_equal(back[: mask.size], mask)


def test_empty_runs():
    runs = codec.nmask_to_runs(np.zeros(16, dtype=np.uint8), 100)
    assert runs.shape == (0, 2)


def test_canonical_keys_matches_the_string_form():
    """codec.canonical_keys (vectorised, what scripts/check_db_against_kmc.py uses) against oracle_py.canon on strings, for
    every K class: one word, the 64-bit boundary, two words, and palindromes"""
    import random
    from oracle import oracle_py
    rng = random.Random(8)
    for K in (1, 5, 16, 31, 32, 33, 47, 60, 63):
        kmers = ["".join(rng.choice("ACGT") for _ in range(K)) for _ in range(300)]
        if K % 2 == 0:
            half = "".join(rng.choice("ACGT") for _ in range(K // 2))
            kmers.append(half + oracle_py.rc(half))                         # its own reverse complement
        keys = np.array([codec.kmer_to_key(s) for s in kmers], dtype=np.uint64)
        got = codec.canonical_keys(keys, K)
        assert [codec.key_to_kmer(int(a), int(b), K) for a, b in got] == [oracle_py.canon(s) for s in kmers], K


def test_check_db_against_kmc_script(tmp_path):
    """scripts/check_db_against_kmc.py end to end: a native source-form database against the KMC database of its sketches
    (written by the independent encoder), and a mismatch is reported with a non-zero exit"""
    import os
    import random
    import subprocess
    import sys
    import kmcdb
    from conftest import ROOT
    from metalign_b200 import dbformat
    from oracle import oracle_py
    rng = random.Random(4)
    K, G, n = 60, 6, 20
    sketches = [["".join(rng.choice("ACGT") for _ in range(K)) if rng.random() < 0.9 else "" for _ in range(n)] for _ in range(G)]
    sketches[3][0] = oracle_py.rc(sketches[1][0] or "A" * K)                # the same k-mer from the other strand: one KMC entry
    keys = codec.sketches_to_keys(sketches, K)
    db = str(tmp_path / "db.mlgdb")
    dbformat.write(db, keys.reshape(-1), ["g%d" % i for i in range(G)], G, n, K, [30, 40, 50, 60])
    cnt = {}
    for sk in sketches:
        for s in sk:
            if s:
                c = oracle_py.canon(s)
                cnt[c] = min(3, cnt.get(c, 0) + 1)
    kmcdb.write(str(tmp_path / "dump"), cnt, K, counter_size=1, version=0x200)
    script = os.path.join(ROOT, "scripts", "check_db_against_kmc.py")
    r = subprocess.run([sys.executable, script, str(tmp_path / "dump"), db], capture_output=True, text=True)
    assert r.returncode == 0 and "identical k-mer sets" in r.stdout, r.stdout + r.stderr
    cnt.pop(next(iter(cnt)))
    kmcdb.write(str(tmp_path / "dump2"), cnt, K, counter_size=1, version=0)
    r = subprocess.run([sys.executable, script, str(tmp_path / "dump2"), db], capture_output=True, text=True)
    assert r.returncode != 0 and "MISMATCH" in (r.stdout + r.stderr)
