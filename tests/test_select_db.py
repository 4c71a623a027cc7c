"""Host side of the drop-in (select_db.py / metalign.py / cmash_tail / ingest), on CPU.
The expected files under tests/golden/select_case/expected_* were produced by the UNMODIFIED reference
(tests/golden/make_select_fixtures.py runs /root/reference/scripts/select_db.py)."""
import argparse
import filecmp
import gzip
import os
import random
import shutil

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT
from metalign_b200 import cmash_tail, ingest, select_db

CASE = os.path.join(ROOT, "tests", "golden", "select_case")


def _args(tmp, **over):
    d = dict(reads=os.path.join(CASE, "reads.fq"), data=os.path.join(CASE, "data"),
             cmash_results=os.path.join(CASE, "cmash_query_results.csv"), cutoff=0.01, db=str(tmp / "cmashed_db.fna"),
             db_dir="AUTO", dbinfo_in="AUTO", dbinfo_out=str(tmp / "subset_db_info.txt"), input_type="AUTO",
             keep_temp_files=False, strain_level=False, temp_dir=str(tmp), threads=4)
    d.update(over)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("tag,over", [("default", {}), ("strain_level", {"strain_level": True}),
                                      ("cutoff_0.3", {"cutoff": 0.3}), ("cutoff_0", {"cutoff": 0.0})])
def test_selection_outputs_match_reference(tmp_path, tag, over):
    select_db.select_main(_args(tmp_path, **over))
    exp = os.path.join(CASE, "expected_" + tag)
    for name in ("cmashed_db.fna", "subset_db_info.txt"):
        assert filecmp.cmp(str(tmp_path / name), os.path.join(exp, name), shallow=False), (tag, name)


def test_cli_flags_match_reference_surface():
    ns = select_db.select_parseargs(["r.fq", "data"])
    assert (ns.cmash_results, ns.cutoff, ns.db, ns.db_dir, ns.dbinfo_in, ns.dbinfo_out, ns.input_type,
            ns.keep_temp_files, ns.strain_level, ns.temp_dir, ns.threads) == \
           ("NONE", 0.01, "AUTO", "AUTO", "AUTO", "AUTO", "AUTO", False, False, "AUTO/", 4)
    from metalign_b200 import metalign
    m = metalign.metalign_parseargs(["r.fq", "data", "--sensitive"])
    assert m.sensitive and m.min_abundance == 10 ** -4 and m.output == "abundances.tsv" and m.pct_id == 0.5
    assert m.read_cutoff == 1 and m.sampleID == "NONE" and m.threads == 4


def test_bad_cutoff_and_unknown_extension(tmp_path):
    with pytest.raises(SystemExit):
        select_db.select_main(_args(tmp_path, cutoff=1.5))
    with pytest.raises(SystemExit):
        select_db.select_main(_args(tmp_path, reads="reads.txt"))


def test_missing_taxid_raises_keyerror(tmp_path):
    csv = tmp_path / "r.csv"
    csv.write_text(",k=30,k=40,k=50,k=60\ntaxid_77_genomic.fna.gz,1.0,1.0,1.0,1.0\n")
    with pytest.raises(KeyError):
        select_db.select_main(_args(tmp_path, cmash_results=str(csv)))


def test_cmash_tail_csv_format(tmp_path):
    names = ["taxid_%d_genomic.fna.gz" % i for i in range(5)]
    ci = np.array([[0.5, 0.4, 0.3, 0.25], [1.0, 1.0, 1.0, 1.0], [0.2, 0.0, 0.0, 0.0], [0.1, 0.1, 0.1, 1 / 3], [0, 0, 0, 0.0]])
    path = str(tmp_path / "out.csv")
    out = cmash_tail.write_results_csv(path, names, (30, 40, 50, 60), ci)
    lines = open(path).read().splitlines()
    assert lines[0] == ",k=30,k=40,k=50,k=60"
    assert [ln.split(",")[0] for ln in lines[1:]] == [names[1], names[3], names[0]]       # k=60 > 0, descending
    assert lines[2].split(",")[-1] == repr(1 / 3)                                           # shortest round-trip repr
    back = pd.read_csv(path, index_col=0)
    assert np.array_equal(back.values, out.values)


def test_cmash_tail_sparse_rows_give_the_same_bytes(tmp_path):
    """the tail fed with rows of the genomes that have a hit (Query.finish_sparse) == the tail fed with the dense table,
    byte for byte, ties in the sort column included (CMash filters before it sorts, so the sort input is identical)"""
    rng = np.random.default_rng(11)
    G = 5000
    names = ["taxid_%d_genomic.fna.gz" % i for i in range(G)]
    ci = np.zeros((G, 4))
    hit = np.sort(rng.choice(G, 900, replace=False))
    ci[hit, 0] = rng.integers(1, 40, hit.size) / 1000.0                      # k=30 hits ...
    top = hit[rng.random(hit.size) < 0.6]
    ci[top, 3] = rng.integers(1, 12, top.size) / 1000.0                      # ... some with k=60 hits, full of ties
    ci[top, 1] = ci[top, 2] = ci[top, 3]
    dense, sparse = str(tmp_path / "dense.csv"), str(tmp_path / "sparse.csv")
    cmash_tail.write_results_csv(dense, names, (30, 40, 50, 60), ci)
    cmash_tail.write_results_csv_sparse(sparse, names, (30, 40, 50, 60), hit.astype(np.uint32), ci[hit])
    assert open(dense, "rb").read() == open(sparse, "rb").read()
    assert len(open(dense).read().splitlines()) == 1 + top.size


def _write(path, text, gz=False):
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(text)
    else:
        with open(path, "w") as f:
            f.write(text)


@pytest.mark.parametrize("gz", [False, True])
def test_ingest_fastq_and_fasta(tmp_path, gz):
    fq = "@r1\nACGTN\n+\nIIIII\n@r2 x\nacgtacgt\n+r2\nACGTACGT\n@r3\n\n+\n\n@r4\nGG\n+\nII"
    p = str(tmp_path / ("a.fq.gz" if gz else "a.fq"))
    _write(p, fq, gz)
    assert ingest.detect_input_type(p) == "fastq"
    got = list(ingest.batches(p, "fastq", reads_per_batch=3))
    reads = []
    for text, off in got:
        reads += [bytes(text[int(off[i]):int(off[i + 1])]).decode() for i in range(off.size - 1)]
    assert reads == ["ACGTN", "acgtacgt", "", "GG"] and len(got) == 2
    fa = ">s1 d\r\nACGT\r\n>s2\nGGCC\nTTAA\n;comment\n\n>s3\nN\n"
    p = str(tmp_path / ("b.fasta.gz" if gz else "b.fasta"))
    _write(p, fa, gz)
    assert ingest.detect_input_type(p) == "fasta"
    reads = []
    for text, off in ingest.batches(p, "fasta"):
        reads += [bytes(text[int(off[i]):int(off[i + 1])]).decode() for i in range(off.size - 1)]
    assert reads == ["ACGT", "GGCC", "TTAA", "N"]


def test_make_db_from_fasta_dump(tmp_path):
    """scripts/make_db.py: the FASTA dump local_tests/dump_kmers.py writes (empty line for an unused slot) -> .mlgdb"""
    import random
    import subprocess
    import sys
    import numpy as np
    from metalign_b200 import codec, dbformat
    rng = random.Random(1)
    G, n, K = 7, 13, 60
    sk = [["" if rng.random() < 0.2 else "".join(rng.choice("ACGT") for _ in range(K)) for _ in range(n)] for _ in range(G)]
    names = ["taxid_%d_genomic.fna.gz" % (100 + g) for g in range(G)]
    dump = tmp_path / "dump.fa"
    with open(dump, "w") as f:
        i = 0
        for g in range(G):
            for s in sk[g]:
                f.write(">seq%d\n%s\n" % (i, s))
                i += 1
    (tmp_path / "names.txt").write_text("\n".join(names) + "\n")
    out = tmp_path / "db.mlgdb"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, os.path.join(root, "scripts", "make_db.py"), str(dump), str(tmp_path / "names.txt"), str(out),
                           "-n", str(n), "-k", str(K)])
    assert dbformat.read_names(str(out)) == names
    h = dbformat.read_header(str(out))
    assert (h["G"], h["n"], h["K"], h["ks"]) == (G, n, K, [30, 40, 50, 60])
    assert np.array_equal(dbformat.read_keys(str(out)), codec.sketches_to_keys(sk, K))
    # chunk boundaries anywhere, CRLF, last record empty and unterminated
    with open(dump, "rb") as f:
        raw = f.read().replace(b"\n", b"\r\n")
    raw = raw + b">last\r\n"
    (tmp_path / "crlf.fa").write_bytes(raw)
    want = np.concatenate([codec.sketches_to_keys(sk, K), np.full((1, 2), codec.EMPTY, dtype=np.uint64)])
    for chunk in (7, 64, 1000, 1 << 20):
        assert np.array_equal(dbformat.keys_from_dump_fasta(str(tmp_path / "crlf.fa"), K, chunk_bytes=chunk), want)


def test_make_db_from_h5(tmp_path):
    """scripts/make_db_from_h5.py against a tiny file in CMash's HDF5 layout (needs h5py: skipped where it is absent, as in
    this repo's build image)"""
    h5py = pytest.importorskip("h5py")
    import importlib.util
    from metalign_b200 import codec, dbformat
    spec = importlib.util.spec_from_file_location("make_db_from_h5", os.path.join(ROOT, "scripts", "make_db_from_h5.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = random.Random(4)
    K, n = 60, 5
    sk = {"b_genome.fna.gz": ["".join(rng.choice("ACGT") for _ in range(K)) for _ in range(n)],
          "a_genome.fna.gz": ["".join(rng.choice("ACGT") for _ in range(K)) for _ in range(3)] + ["", ""]}
    p = str(tmp_path / "train.h5")
    with h5py.File(p, "w") as f:
        grp = f.create_group("CountEstimators")
        for name, kmers in sk.items():
            g = grp.create_group(name)
            g.attrs["ksize"] = K
            g.create_dataset("kmers", data=np.array([k.encode() for k in kmers], dtype="S%d" % K))
    out = str(tmp_path / "db.mlgdb")
    mod.main([p, out])
    assert dbformat.read_names(out) == sorted(sk)
    want = codec.sketches_to_keys([sk[k] for k in sorted(sk)], K)
    assert np.array_equal(dbformat.read_keys(out).reshape(-1), np.asarray(want).reshape(-1))
