"""CPU tests of the oracle itself: golden cases, Python-vs-C cross-check, packed-vs-ASCII inputs."""
import random

import numpy as np
import pytest

import synth
from metalign_b200 import codec
from oracle import oracle_py
from oracle.oracle_c import OracleDB, OracleQuery

from helpers import adversarial_case, oracle_c_run


def _run_c(case, e):
    G, n = len(case["sketches"]), len(case["sketches"][0])
    keys = codec.sketches_to_keys(case["sketches"], case["K"])
    return oracle_c_run(keys, G, n, case["K"], case["ks"], lambda q: q.push_reads(case["reads"]),
                        e["ci_min"], e["gate"], e["count_empty_in_den"])


def test_golden_cases_python_oracle(golden_cases):
    for c in golden_cases:
        for e in c["expect"]:
            r = oracle_py.run(c["reads"], c["sketches"], K=c["K"], ks=c["ks"], ci_min=e["ci_min"], gate=e["gate"],
                              count_empty_in_den=e["count_empty_in_den"])
            assert r["num"] == e["num"], c["name"]
            assert r["den"] == e["den"], c["name"]
            assert r["I"] == e["I"], c["name"]


def test_golden_cases_c_oracle(golden_cases):
    for c in golden_cases:
        for e in c["expect"]:
            r, I = _run_c(c, e)
            assert r["num"].tolist() == e["num"], c["name"]
            assert r["den"].tolist() == e["den"], c["name"]
            assert [codec.key_to_kmer(a, b, c["K"]) for a, b in I] == e["I"], c["name"]
            exp_ci = [[(nu / de if nu > 0 else 0.0) for nu, de in zip(rn, rd)] for rn, rd in zip(e["num"], e["den"])]
            assert np.array_equal(r["ci"], np.array(exp_ci))   # same IEEE division


@pytest.mark.parametrize("seed", range(40))
def test_python_vs_c_oracle_adversarial(seed):
    rng = random.Random(seed)
    c = adversarial_case(rng)
    G, n = len(c["sketches"]), len(c["sketches"][0])
    for gate in ("exact", "none"):
        ci_min = rng.choice([1, 2, 3])
        ce = rng.random() < 0.5
        py = oracle_py.run(c["reads"], c["sketches"], K=c["K"], ks=c["ks"], ci_min=ci_min, gate=gate, count_empty_in_den=ce)
        keys = codec.sketches_to_keys(c["sketches"], c["K"])
        r, I = oracle_c_run(keys, G, n, c["K"], c["ks"], lambda q: q.push_reads(c["reads"]), ci_min, gate, ce)
        assert r["num"].tolist() == py["num"]
        assert r["den"].tolist() == py["den"]
        assert r["n_kmers"] == py["n_kmers"]
        assert [codec.key_to_kmer(a, b, c["K"]) for a, b in I] == py["I"]
        assert np.array_equal(r["ci"], np.array(py["ci"]).reshape(G, len(c["ks"])))


def test_python_vs_c_oracle_synthetic():
    p = synth.params(G=20, n=50, len_min=5000, len_max=20000, n_present=8)
    keys = synth.sketch_keys(p)
    sk = codec.keys_to_sketches(keys, 20, 50, 60)
    reads = [bytes(r).decode() for r in synth.reads_ascii(p, 0, 6000)]
    for gate in ("exact", "none"):
        py = oracle_py.run(reads, sk, gate=gate)
        r, I = oracle_c_run(keys, 20, 50, 60, (30, 40, 50, 60), lambda q: q.push_reads(reads), 2, gate, True)
        assert r["num"].tolist() == py["num"] and r["den"].tolist() == py["den"]
        assert r["n_intersect"] == len(py["I"]) > 0
    # the rc-canonical half of the sketch loses its k>30 hits under the exact gate
    ex = oracle_py.run(reads, sk, gate="exact")["num"]
    no = oracle_py.run(reads, sk, gate="none")["num"]
    tot_ex, tot_no = np.sum(ex, axis=0), np.sum(no, axis=0)
    assert tot_ex[0] == tot_no[0] and tot_ex[3] < tot_no[3]


def test_packed_equals_ascii_in_c_oracle():
    p = synth.params(G=12, n=30, len_min=4000, len_max=9000, n_present=5)
    keys = synth.sketch_keys(p)
    reads = [bytes(r).decode() for r in synth.reads_ascii(p, 0, 3000)]
    bases, nmask = synth.reads_packed(p, 0, 3000)
    assert codec.unpack_reads(bases, nmask, None, 3000, p.read_len) == reads
    b2, m2, off = codec.pack_reads(reads)
    assert np.array_equal(b2, bases) and np.array_equal(m2, nmask)
    ra, Ia = oracle_c_run(keys, 12, 30, 60, (30, 40, 50, 60), lambda q: q.push_reads(reads))
    rp, Ip = oracle_c_run(keys, 12, 30, 60, (30, 40, 50, 60), lambda q: q.push_packed(bases, nmask, None, 3000, p.read_len))
    ro, Io = oracle_c_run(keys, 12, 30, 60, (30, 40, 50, 60), lambda q: q.push_packed(bases, nmask, off, 3000))
    for r, I in ((rp, Ip), (ro, Io)):
        assert np.array_equal(r["num"], ra["num"]) and r["n_kmers"] == ra["n_kmers"] and np.array_equal(I, Ia)


def test_counts_are_not_additive_across_shards_but_counter_table_is():
    """per-genome tables of shards must not be summed; the clamped per-k-mer counters can."""
    p = synth.params(G=12, n=30, len_min=4000, len_max=9000, n_present=5)
    keys = synth.sketch_keys(p)
    reads = [bytes(r).decode() for r in synth.reads_ascii(p, 0, 3000)]
    full, I_full = oracle_c_run(keys, 12, 30, 60, (30, 40, 50, 60), lambda q: q.push_reads(reads))
    db = OracleDB(keys, 12, 30, 60, (30, 40, 50, 60))
    qs = [OracleQuery(db), OracleQuery(db)]
    qs[0].push_reads(reads[0::2])
    qs[1].push_reads(reads[1::2])
    summed = qs[0].export_counts().astype(np.uint16) + qs[1].export_counts().astype(np.uint16)
    assert summed.max() <= 4
    merged = OracleQuery(db)
    merged.import_counts(summed.astype(np.uint8))
    m = merged.finish()
    assert np.array_equal(m["num"], full["num"]) and m["n_intersect"] == full["n_intersect"]
    shard_sum = qs[0].finish()["num"] + qs[1].finish()["num"]
    assert not np.array_equal(shard_sum, full["num"])


def test_select_organisms_matches_reference_logic():
    info = {"562": [["a"], "1", "n", "1|2|3|4|5|6|561|562"], "562.1": [["b"], "1", "n", "1|2|3|4|5|6|561|562.1"],
            "9": [["c"], "1", "n", "1|2|3|4|5|6||9"], "9.1": [["d"], "1", "n", "1|2|3|4|5|6||9.1"]}
    rows = [("taxid_562_genomic.fna.gz", 0.5), ("taxid_562_1_genomic.fna.gz", 0.4), ("taxid_9_genomic.fna.gz", 0.01),
            ("taxid_9_1_genomic.fna.gz", 0.02), ("taxid_562_genomic.fna.gz", 0.001)]
    assert oracle_py.select_organisms(rows, info, cutoff=0.01) == [
        "taxid_562_genomic.fna.gz", "taxid_9_genomic.fna.gz", "taxid_9_1_genomic.fna.gz"]
    assert len(oracle_py.select_organisms(rows, info, cutoff=0.01, strain_level=True)) == 4
    with pytest.raises(KeyError):
        oracle_py.select_organisms([("taxid_77_genomic.fna.gz", 1.0)], info)
