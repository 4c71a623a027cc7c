"""End to end through the drop-in on the GPU: reads file + data/ directory in, the three files of the
reference's contract out (cmash_query_results.csv, cmashed_db.fna, subset_db_info.txt), checked against the
CPU oracle pushed through the same pandas tail and the same selection rule."""
import argparse
import gzip
import os

import numpy as np
import pytest

import synth
from metalign_b200 import cmash_tail, codec, dbformat, select_db
from oracle import oracle_py

from helpers import oracle_c_run

pytestmark = pytest.mark.gpu
KS = (30, 40, 50, 60)


def _make_data_dir(tmp_path, p, keys, names):
    data = tmp_path / "data"
    (data / "organism_files").mkdir(parents=True)
    dbformat.write(str(data / select_db.DB_BASENAME), keys, names, p.G, p.n, 60, KS)
    with open(data / "db_info.txt", "w") as f:
        f.write("Accession\tLength\tTaxID\tLineage\tTaxID_Lineage\n")
        for g, name in enumerate(names):
            taxid = select_db.taxid_of(name)
            species = "" if g % 7 == 3 else str(5000 + g // 2)          # pairs of strains share a species; some have none
            lin = "2|1224|1236|91347|543|561|%s|%s" % (species, taxid)
            f.write("\t".join(["ACC%05d.1" % g, str(4000 + g), taxid, "n|n|n|n|n|n|n|n", lin]) + "\n")
            with gzip.open(data / "organism_files" / name, "wt") as gz:
                gz.write(">ACC%05d.1 genome %d\n%s\n" % (g, g, "ACGT" * (10 + g % 5)))
    return str(data)


@pytest.mark.parametrize("gate,fmt", [("exact", "fastq"), ("none", "fasta.gz")])
def test_select_main_end_to_end(tmp_path, gate, fmt):
    p = synth.params(G=60, n=120, seed=21, len_min=8000, len_max=20000, n_present=12)
    keys = synth.sketch_keys(p)
    names = ["taxid_%d_%d_genomic.fna.gz" % (1000 + g // 2, 1 + g % 2) for g in range(p.G)]
    assert names == sorted(names)
    data = _make_data_dir(tmp_path, p, keys, names)
    nreads = 30000
    reads = [bytes(r).decode() for r in synth.reads_ascii(p, 0, nreads)]
    if fmt == "fastq":
        rp = tmp_path / "reads.fq"
        with open(rp, "w") as f:
            for i, r in enumerate(reads):
                f.write("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    else:
        rp = tmp_path / "reads.fasta.gz"
        with gzip.open(rp, "wt") as f:
            for i, r in enumerate(reads):
                f.write(">r%d\n%s\n" % (i, r))
    tmp = tmp_path / "tmp"
    args = argparse.Namespace(reads=str(rp), data=data, cmash_results="NONE", cutoff=0.01, db="AUTO", db_dir="AUTO",
                              dbinfo_in="AUTO", dbinfo_out="AUTO", input_type="AUTO", keep_temp_files=True,
                              strain_level=False, temp_dir=str(tmp), threads=4, gate=gate, device=0, db_file="AUTO")
    chosen = select_db.select_main(args)

    # oracle through the same tail
    ref, I_ref = oracle_c_run(keys, p.G, p.n, 60, KS, lambda q: q.push_reads(reads), 2, gate, True)
    exp_csv = tmp_path / "expected.csv"
    frame = cmash_tail.write_results_csv(str(exp_csv), names, KS, ref["ci"], 0.0)
    assert open(tmp / "cmash_query_results.csv", "rb").read() == open(exp_csv, "rb").read()
    rows = [(name, float(v)) for name, v in zip(frame.index, frame["k=60"])]
    taxid2info = select_db.read_dbinfo(args)
    assert chosen == oracle_py.select_organisms(rows, taxid2info, 0.01, False) and len(chosen) > 0
    # the debug artefact: sorted canonical 60-mers of the intersection
    dump = [ln.split("\t") for ln in open(tmp / "60mers_intersection_dump").read().splitlines()]
    assert [d[0] for d in dump] == [codec.key_to_kmer(a, b, 60) for a, b in I_ref]
    assert all(len(d) == 2 and 1 <= int(d[1]) <= 3 for d in dump)
    # subset files are consistent with the selection
    info = open(tmp / "subset_db_info.txt").read().splitlines()
    assert info[0].startswith("Accesion\t") and info[1].startswith("Unmapped\t0")
    assert [ln.split("\t")[2] for ln in info[2:]] == [select_db.taxid_of(c) for c in chosen]
    fna = open(tmp / "cmashed_db.fna").read()
    assert fna.count(">") == len(chosen)
