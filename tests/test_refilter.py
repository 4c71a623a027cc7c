"""CMash's post-processing when --sensitive is absent (re-filter to k-mers unique to one organism).
Host side (metalign_b200/cmash_tail.py: refilter_unique) against the string/set restatement in oracle/oracle_py.py on
small random databases; on the GPU, Query.hit_flags against the oracle's hit set and the whole CSV."""
import random

import numpy as np
import pytest

from metalign_b200 import cmash_tail, codec, dbformat
from oracle import oracle_py


def _case(rng, G, n, K, ks, share=0.4):
    """sketches with shared k-mers, shared prefixes of different length, duplicates inside a sketch, '' slots; reads = I directly"""
    pool = ["".join(rng.choice("ACGT") for _ in range(K)) for _ in range(G * n // 2 + 4)]
    sketches = []
    for g in range(G):
        sk = []
        for j in range(n):
            x = rng.random()
            if x < 0.1:
                sk.append("")
            elif x < 0.1 + share:
                s = rng.choice(pool)
                if rng.random() < 0.3:                       # same prefix, different tail
                    cut = rng.choice(ks)
                    s = s[:cut] + "".join(rng.choice("ACGT") for _ in range(K - cut))
                sk.append(s)
            else:
                sk.append("".join(rng.choice("ACGT") for _ in range(K)))
        sketches.append(sk)
    present = [s for sk in sketches for s in sk if s and rng.random() < 0.5]
    I = {oracle_py.canon(s) for s in present}
    return sketches, I


def _flags_from_H(H, sketches, ks, cand):
    """what Query.hit_flags returns: 1 at the representative slot of each hit class (here: the first slot of the class)"""
    n = len(sketches[0])
    out = np.zeros((len(cand), len(ks), n), dtype=np.uint8)
    for row, g in enumerate(cand):
        for ki, k in enumerate(ks):
            hit = {sketches[g][j][:k] for (gg, kk, j) in H if gg == g and kk == k}
            seen = set()
            for j, s in enumerate(sketches[g]):
                if s and s[:k] in hit and s[:k] not in seen:
                    seen.add(s[:k])
                    out[row, ki, j] = 1
    return out


@pytest.mark.parametrize("seed", range(12))
def test_host_refilter_matches_oracle(seed):
    rng = random.Random(seed)
    K = rng.choice([12, 20, 33, 60])
    ks = sorted(set(rng.sample(range(max(4, K - 30), K), 2)) | {K})
    G, n = rng.randint(2, 9), rng.randint(3, 12)
    sketches, I = _case(rng, G, n, K, ks)
    H = oracle_py.query_hits(sorted(I), sketches, ks, gate="none")
    _, _, ci = oracle_py.containment_table(H, sketches, ks)
    cand, num_o, den_o, ci_o = oracle_py.refilter_unique(H, sketches, ks, ci, 0.0)
    keys = codec.sketches_to_keys(sketches, K).reshape(G, n, 2)
    num, den, ci2 = cmash_tail.refilter_unique(keys[cand], _flags_from_H(H, sketches, ks, cand), K, ks)
    assert num.tolist() == num_o and den.tolist() == den_o
    assert ci2.tolist() == ci_o
    if len(cand) > 1:
        assert (np.asarray(den_o) <= np.asarray([[len({s[:k] for s in sketches[g] if s}) for k in ks] for g in cand])).all()


def test_read_keys_rows(tmp_path):
    rng = random.Random(1)
    K, G, n = 60, 7, 5
    sketches = [["".join(rng.choice("ACGT") for _ in range(K)) if rng.random() < 0.8 else "" for _ in range(n)] for _ in range(G)]
    keys = codec.sketches_to_keys(sketches, K)
    p = str(tmp_path / "db.mlgdb")
    dbformat.write(p, keys.reshape(-1), ["g%d" % i for i in range(G)], G, n, K, [30, 40, 50, 60])
    got = dbformat.read_keys_rows(p, [6, 0, 3])
    assert np.array_equal(got, keys.reshape(G, n, 2)[[6, 0, 3]])
    with pytest.raises(ValueError):
        dbformat.read_keys_rows(p, [7])


@pytest.mark.gpu
@pytest.mark.parametrize("gate", ["exact", "none"])
def test_gpu_hit_flags_and_specific_csv(ctx, tmp_path, gate):
    """Query.hit_flags == the oracle's hit set, class by class, and the CSV of the non-sensitive mode == the one built from
    the oracle's tables through the same pandas tail"""
    from metalign_b200.api import Database
    rng = random.Random(3)
    K, ks = 60, [30, 40, 50, 60]
    G, n = 40, 30
    sketches, I = _case(rng, G, n, K, ks, share=0.5)
    reads = []
    for x in sorted(I):
        for _ in range(2):
            reads.append(x if rng.random() < 0.5 else oracle_py.rc(x))
    names = ["taxid_%d_genomic.fna.gz" % i for i in range(G)]
    db = Database.from_sketches(ctx, sketches, K, ks, names=names)
    q = db.query(2, gate, True)
    q.push_reads(reads)
    res = q.finish_sparse()
    H = oracle_py.query_hits(sorted(I), sketches, ks, gate=gate)
    _, _, ci = oracle_py.containment_table(H, sketches, ks)
    cand, num_o, den_o, ci_o = oracle_py.refilter_unique(H, sketches, ks, ci, 0.0)
    assert len(cand) > 3
    flags = q.hit_flags(cand)
    for row, g in enumerate(cand):
        for ki, k in enumerate(ks):
            got = {sketches[g][j][:k] for j in np.nonzero(flags[row, ki])[0]}
            assert got == {sketches[g][j][:k] for (gg, kk, j) in H if gg == g and kk == k}, (g, k)
            assert flags[row, ki].sum() == len(got)                       # one representative slot per class
    keys = codec.sketches_to_keys(sketches, K).reshape(G, n, 2)
    out = cmash_tail.write_results_csv_specific(str(tmp_path / "specific.csv"), names, ks, K, res["genomes"], res["ci"],
                                                lambda c: keys[np.asarray(c, dtype=np.int64)], q.hit_flags, 0.0)
    q.close()
    db.close()
    want = cmash_tail.filter_and_sort(cmash_tail.containment_frame([names[g] for g in cand], ks, np.asarray(ci_o)), 0.0)
    want.to_csv(str(tmp_path / "want.csv"), index=True, encoding="utf-8")
    assert open(tmp_path / "specific.csv", "rb").read() == open(tmp_path / "want.csv", "rb").read()
    assert len(out) <= len(cand)
