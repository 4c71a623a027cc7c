"""The reference-driven fixture (tests/golden/ref_e2e_case, made by tests/golden/make_ref_e2e_fixture.py): files the
UNMODIFIED reference `scripts/select_db.py` wrote when its four subprocesses (select_db.py:50-76) were the
oracle-backed stubs of tests/golden/stub_tools/.

CPU tests: the C oracle (an independent restatement) reproduces the committed files from the same inputs; the KMC
database files of the fixture decode to the expected k-mer sets.
GPU test: the drop-in `select_main`, through the CUDA library, writes byte-identical files.
"""
import argparse
import filecmp
import gzip
import os

import numpy as np
import pytest

import kmcdb
from conftest import ROOT
from helpers import oracle_c_run
from metalign_b200 import cmash_tail, codec, dbformat, select_db
from oracle import oracle_py

CASE = os.path.join(ROOT, "tests", "golden", "ref_e2e_case")
DATA = os.path.join(CASE, "data")
RUNS = [("default", "reads.fq", {}), ("strain_level", "reads.fq", {"strain_level": True}),
        ("cutoff_0.5", "reads.fq", {"cutoff": 0.5}), ("fasta_gz", "reads.fa.gz", {"cutoff": 0.0})]


def _reads():
    with open(os.path.join(CASE, "reads.fq")) as f:
        lines = f.read().split("\n")
    return [lines[i + 1] for i in range(0, len(lines) - 1, 4)]


def _db():
    path = os.path.join(DATA, select_db.DB_BASENAME)
    h = dbformat.read_header(path)
    return dbformat.read_keys(path), dbformat.read_names(path), h


def test_c_oracle_reproduces_reference_run(tmp_path):
    keys, names, h = _db()
    reads = _reads()
    res, I = oracle_c_run(keys, h["G"], h["n"], h["K"], h["ks"], lambda q: q.push_reads(reads), 2, "exact", True)
    out = tmp_path / "cmash_query_results.csv"
    cmash_tail.write_results_csv(str(out), names, h["ks"], res["ci"], 0.0)
    exp = os.path.join(CASE, "expected_default")
    assert open(out, "rb").read() == open(os.path.join(exp, "cmash_query_results.csv"), "rb").read()
    dumped = [ln.split()[0] for ln in open(os.path.join(exp, "60mers_intersection_dump"))]
    assert dumped == [codec.key_to_kmer(a, b, h["K"]) for a, b in I] and len(dumped) > 50


def test_fixture_kmc_databases_decode():
    keys, names, h = _db()
    sk = [[codec.key_to_kmer(a, b, h["K"]) if a != codec.EMPTY else "" for a, b in keys[g * h["n"]:(g + 1) * h["n"]]]
          for g in range(h["G"])]
    D = oracle_py.db_kmer_set(sk)
    hdr, recs = kmcdb.read(os.path.join(DATA, "cmash_db_n1000_k60_dump"))            # kmc -k60 -fa -ci0 -cs3
    assert hdr["version"] == 0x200 and hdr["k"] == 60 and hdr["canonical"] and {x for x, _ in recs} == D
    cnt = oracle_py.count_read_kmers(_reads(), 60)
    hdr, recs = kmcdb.read(os.path.join(CASE, "expected_default", "reads_60mers"))   # kmc -k60 -fq -ci2 -cs3
    assert dict(recs) == {x: min(c, 3) for x, c in cnt.items() if c >= 2} and hdr["min_count"] == 2
    hdr, recs = kmcdb.read(os.path.join(CASE, "expected_default", "60mers_intersection"))
    assert hdr["version"] == 0 and {x for x, _ in recs} == oracle_py.intersect(cnt, D, 2)


def test_kmcdb_roundtrip_both_layouts(tmp_path):
    import random
    rng = random.Random(3)
    for version in (0, 0x200):
        for k, csz in ((60, 1), (21, 2), (32, 4)):
            kmers = {"".join(rng.choice("ACGT") for _ in range(k)): rng.randint(1, 200) for _ in range(300)}
            p = str(tmp_path / ("db_%d_%d" % (version, k)))
            kmcdb.write(p, kmers, k, counter_size=csz, version=version)
            hdr, recs = kmcdb.read(p)
            assert dict(recs) == kmers and hdr["k"] == k and hdr["total"] == len(kmers)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,rfile,over", RUNS)
def test_dropin_writes_the_files_the_reference_wrote(tmp_path, tag, rfile, over):
    d = dict(reads=os.path.join(CASE, rfile), data=DATA, cmash_results="NONE", cutoff=0.01, db="AUTO", db_dir="AUTO",
             dbinfo_in="AUTO", dbinfo_out="AUTO", input_type="AUTO", keep_temp_files=True, strain_level=False,
             temp_dir=str(tmp_path / "out"), threads=4)          # exactly the Namespace the reference's parser builds
    d.update(over)
    select_db.select_main(argparse.Namespace(**d))
    exp = os.path.join(CASE, "expected_" + tag)
    for name in ("cmash_query_results.csv", "cmashed_db.fna", "subset_db_info.txt"):
        assert filecmp.cmp(str(tmp_path / "out" / name), os.path.join(exp, name), shallow=False), (tag, name)
    # kmc_dump's file ("<k-mer>\t<count>", count = min of the two -cs3 counters) and the FASTA rewrite of select_db.py:61-65
    assert filecmp.cmp(str(tmp_path / "out" / "60mers_intersection_dump"), os.path.join(exp, "60mers_intersection_dump"), shallow=False), tag
    kmers = [ln.split()[0] for ln in open(os.path.join(exp, "60mers_intersection_dump"))]
    assert open(tmp_path / "out" / "60mers_intersection_dump.fa").read() == "".join(">seq\n%s\n" % x for x in kmers)


def test_native_kmc_reader_on_the_fixture_databases(tmp_path):
    """the product's C++ reader of KMC databases (csrc/kmcdb.h) against the independent Python encoder / decoder: the three
    databases of the reference-driven fixture (layout 0x200 written by the stub `kmc`, layout 0 by the stub `kmc_tools`),
    random databases of other k / counter sizes, and damaged files"""
    import random
    from metalign_b200 import ingest
    for prefix in (os.path.join(DATA, "cmash_db_n1000_k60_dump"), os.path.join(CASE, "expected_default", "reads_60mers"),
                   os.path.join(CASE, "expected_default", "60mers_intersection")):
        hdr, recs = kmcdb.read(prefix)
        keys, info, counts = ingest.read_kmc_database(prefix, with_counts=True)
        assert info["k"] == 60 and info["total"] == len(recs) == hdr["total"] and info["version"] == hdr["version"] and info["canonical"]
        assert [codec.key_to_kmer(a, b, 60) for a, b in keys] == [x for x, _ in recs] and counts.tolist() == [c for _, c in recs]
    # the shipped-database check a user would run: the k-mers of the KMC dump == the canonical k-mers of the native file
    keys, names, h = _db()
    dkeys, _ = ingest.read_kmc_database(os.path.join(DATA, "cmash_db_n1000_k60_dump"))
    have = {codec.key_to_kmer(a, b, 60) for a, b in dkeys}
    want = {oracle_py.canon(codec.key_to_kmer(a, b, 60)) for a, b in keys if a != codec.EMPTY}
    assert have == want
    rng = random.Random(9)
    for version in (0, 0x200):
        for k, csz in ((21, 2), (32, 4), (63, 1), (13, 1)):
            kmers = {"".join(rng.choice("ACGT") for _ in range(k)): rng.randint(1, 250) for _ in range(500)}
            p = str(tmp_path / ("db_%d_%d" % (version, k)))
            kmcdb.write(p, kmers, k, counter_size=csz, version=version)
            got, info, counts = ingest.read_kmc_database(p, with_counts=True)
            assert {codec.key_to_kmer(a, b, k): int(c) for (a, b), c in zip(got, counts)} == kmers
    raw = open(p + ".kmc_suf", "rb").read()
    open(p + ".kmc_suf", "wb").write(raw[:-9])
    with pytest.raises(IOError):
        ingest.read_kmc_database(p)
    with pytest.raises(IOError):
        ingest.read_kmc_database(str(tmp_path / "missing"))


def test_native_kmc_reader_survives_damaged_files(tmp_path):
    """byte damage and truncation of both files of a KMC database: the C++ reader answers or raises IOError, it never reads
    out of bounds (run in a child process so that a crash would fail this test instead of ending the test run)"""
    import random
    import subprocess
    import sys
    rng = random.Random(2)
    kmers = {"".join(rng.choice("ACGT") for _ in range(60)): rng.randint(1, 3) for _ in range(400)}
    for version in (0, 0x200):
        kmcdb.write(str(tmp_path / ("good%d" % version)), kmers, 60, counter_size=1, version=version)
    code = r'''
import random, sys
sys.path.insert(0, %r)
from metalign_b200 import ingest
rng = random.Random(7)
ok = bad = 0
for version in (0, 0x200):
    pre = bytearray(open(%r + "/good%%d.kmc_pre" %% version, "rb").read())
    suf = bytearray(open(%r + "/good%%d.kmc_suf" %% version, "rb").read())
    for trial in range(400):
        a, b = bytearray(pre), bytearray(suf)
        which = a if trial %% 3 else b
        if trial %% 7 == 0:
            del which[rng.randrange(4, len(which)):]
        else:
            for _ in range(rng.randint(1, 4)):
                i = rng.randrange(len(which))
                which[i] = rng.randrange(256)
            if trial %% 11 == 0 and len(a) > 60:          # damage inside the header block at the end of .kmc_pre
                for _ in range(3):
                    a[len(a) - 1 - rng.randrange(56)] = rng.randrange(256)
        open(%r + "/bad.kmc_pre", "wb").write(bytes(a)); open(%r + "/bad.kmc_suf", "wb").write(bytes(b))
        try:
            ingest.read_kmc_database(%r + "/bad", with_counts=True); ok += 1
        except (IOError, MemoryError, ValueError):
            bad += 1
print("ok", ok, "refused", bad)
''' % (ROOT, str(tmp_path), str(tmp_path), str(tmp_path), str(tmp_path), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "refused" in r.stdout, (r.returncode, r.stdout[-300:], r.stderr[-600:])
