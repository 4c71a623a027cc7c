#!/usr/bin/env python
"""Generates tests/golden/ref_e2e_case/: a small but complete Metalign `data/` directory, a reads file, and the
files the UNMODIFIED reference (`/root/reference/scripts/select_db.py`, whole `select_main`, no `--cmash_results`
seam) writes from them when the four tools it starts -- `kmc`, `kmc_tools`, `kmc_dump`,
`StreamingQueryDNADatabase.py` (select_db.py:50-76) -- are the oracle-backed stubs of tests/golden/stub_tools/
put first on PATH.  What this pins that nothing else does: the argv the reference builds for each tool, its
dump -> FASTA rewrite (select_db.py:61-65), the way it reads the CSV back (:80-96) and make_db_and_dbinfo
(:99-117), all executed by the reference's own code on the outputs of the CPU oracle.  The tools themselves remain
models (KMC / CMash are absent: "parity unpinned").

The database artefacts are made the way local_tests/retrain_and_test_metalign.sh:49-66 makes them:
sketches -> FASTA dump in dump_kmers.py's format -> `kmc -k60 -fa -ci0 -cs3` -> cmash_db_n1000_k60_dump.kmc_*.
The training "HDF5" is a JSON stand-in (h5py is not installed); the native cmash_db_n1000_k60.mlgdb the drop-in
reads is converted from the same FASTA dump by scripts/make_db.py.

Run (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_ref_e2e_fixture.py
"""
import gzip
import json
import os
import random
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASE = os.path.join(HERE, "ref_e2e_case")
STUBS = os.path.join(HERE, "stub_tools")
REF = "/root/reference/scripts/select_db.py"
K, N_SLOTS, READ_LEN = 60, 32, 150
COMP = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s.translate(COMP)[::-1]


def mutate(rng, s, rate):
    return "".join(rng.choice([c for c in "ACGT" if c != ch]) if rng.random() < rate else ch for ch in s)


def world(rng):
    """genomes: name -> (taxid, name lineage, taxid lineage, [(accession, sequence)...]); sketches: name -> n slots"""
    def rand_seq(n):
        return "".join(rng.choice("ACGT") for _ in range(n))
    spec = [
        # taxid, species taxid, parent taxid (strain of), length, n accessions
        ("1001.1", "1001", None, 4200, 1), ("1001.2", "1001", "1001.1", 0, 1), ("1001.3", "1001", "1001.1", 0, 2),
        ("1002.1", "1002", None, 3600, 1), ("1002.2", "1002", "1002.1", 0, 1),
        ("1003.1", "1003", None, 5000, 2), ("1004.1", "1004", None, 3000, 1), ("1005.1", "1005", None, 3300, 1),
        ("2001.1", "", None, 2500, 1), ("2001.2", "", "2001.1", 0, 1),          # viruses: species field empty
        ("1006.1", "1006", None, 75, 1),                                         # tiny genome: 16 windows < n slots
        ("1007.1", "1007", None, 3100, 1), ("1008.1", "1008", None, 2900, 1), ("1009.1", "1009", None, 3400, 1),
    ]
    seqs, genomes, sketches, positions = {}, {}, {}, {}
    for taxid, species, parent, length, nacc in spec:
        if parent is None:
            s = rand_seq(length)
            if taxid == "1003.1":        # a 45-base repeat: two sketch slots that share their 30- and 40-prefix
                s = s[:1000] + s[200:245] + s[1045:]
        else:
            s = mutate(rng, seqs[parent], 0.004)
        seqs[taxid] = s
        name = "taxid_" + taxid.replace(".", "_") + "_genomic.fna.gz"
        cut = [0] + sorted(rng.sample(range(200, len(s) - 200), nacc - 1)) + [len(s)] if nacc > 1 else [0, len(s)]
        recs = [("NC_%s%02d.1" % (taxid.replace(".", ""), i), s[cut[i]:cut[i + 1]]) for i in range(nacc)]
        if species:
            names = "Bacteria|Phylum%s|Class|Order|Family|Genus%s|Species %s|Strain %s" % (species[-1], species[-2:], species, taxid)
            taxids = "2|12|123|1234|12345|9%s|%s|%s" % (species, species, taxid)
        else:
            names = "Viruses|||||||Phage %s" % taxid
            taxids = "10239|||||||%s" % taxid
        genomes[name] = (taxid, names, taxids, recs)
        # sketch: N_SLOTS window starts of the FIRST record's strand (a strain reuses its parent's positions), '' padding
        first = recs[0][1] if nacc == 1 else s
        nwin = len(first) - K + 1
        if parent is None:
            pos = sorted(rng.sample(range(nwin), min(N_SLOTS, nwin)))
            if taxid == "1003.1":
                pos = sorted(set(pos[:-2]) | {200, 1000})
        else:
            pos = positions[parent]
        positions[taxid] = pos
        sk = [first[p:p + K] for p in pos]
        rng.shuffle(sk)                   # slot order = hash rank in CMash: unrelated to position
        sketches[name] = sk + [""] * (N_SLOTS - len(sk))
    return genomes, sketches, seqs


def reads_for(rng, seqs):
    cov = {"1001.1": 9.0, "1002.2": 5.0, "1003.1": 6.0, "1004.1": 1.0, "2001.1": 7.0, "1006.1": 30.0, "1008.1": 3.0}
    reads = []
    for taxid, c in cov.items():
        s = seqs[taxid]
        L = min(READ_LEN, len(s))
        for _ in range(max(2, int(c * len(s) / L))):
            a = rng.randint(0, len(s) - L)
            r = s[a:a + L]
            if rng.random() < 0.5:
                r = rc(r)
            r = list(mutate(rng, r, 0.004))
            for i in range(len(r)):
                x = rng.random()
                if x < 0.002:
                    r[i] = "N"
                elif x < 0.01:
                    r[i] = r[i].lower()
            reads.append("".join(r))
    reads += ["ACGTACGTAC", "N" * 70, ""]            # shorter than K, all N, empty
    rng.shuffle(reads)
    return reads


def write_data(data, genomes, sketches):
    os.makedirs(os.path.join(data, "organism_files"))
    with open(os.path.join(data, "db_info.txt"), "w") as f:
        f.write("Accession\tLength\tTaxID\tLineage\tTaxID_Lineage\n")
        for name, (taxid, names, taxids, recs) in genomes.items():
            for acc, s in recs:
                f.write("\t".join([acc, str(len(s)), taxid, names, taxids]) + "\n")
    for name, (taxid, _, _, recs) in genomes.items():
        with open(os.path.join(data, "organism_files", name), "wb") as raw:
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
                for acc, s in recs:
                    f.write((">%s synthetic %s\n" % (acc, taxid)).encode())
                    for i in range(0, len(s), 70):
                        f.write((s[i:i + 70] + "\n").encode())
    # training database stand-in (CMash: group CountEstimators/<basename>/kmers) + empty prefilter
    order = list(genomes)
    random.Random(5).shuffle(order)                  # file order is NOT the import order: CMash sorts by basename
    with open(os.path.join(data, "cmash_db_n1000_k60.h5"), "w") as f:
        json.dump({"stand_in_for": "CMash training HDF5", "ksize": K, "names": order,
                   "sketches": [sketches[nm] for nm in order]}, f)
    open(os.path.join(data, "cmash_db_n1000_k60_30-60-10.bf"), "w").write("stand-in for the hydra Bloom prefilter\n")
    # dump_kmers.py:7-14: every slot of every sketch, import (= sorted basename) order, '' slots included
    dump = os.path.join(data, "cmash_db_n1000_k60_dump.fa")
    with open(dump, "w") as f:
        i = 0
        for nm in sorted(genomes):
            for kmer in sketches[nm]:
                f.write(">seq%d\n%s\n" % (i, kmer))
                i += 1
    return dump


def main():
    if not os.path.exists(REF):
        sys.exit("the reference is not mounted here; fixtures can only be regenerated in the build container")
    rng = random.Random(20200529)
    shutil.rmtree(CASE, ignore_errors=True)
    data = os.path.join(CASE, "data")
    genomes, sketches, seqs = world(rng)
    dump = write_data(data, genomes, sketches)
    env = dict(os.environ, PATH=STUBS + os.pathsep + os.environ["PATH"], MLG_STUB_GATE="exact")
    # retrain_and_test_metalign.sh:66
    subprocess.check_call(["kmc", "-v", "-k60", "-fa", "-ci0", "-cs3", "-t8", "-jlogsample", dump,
                           os.path.join(data, "cmash_db_n1000_k60_dump"), "."], env=env, cwd=data)
    os.remove(os.path.join(data, "logsample"))
    names_txt = os.path.join(CASE, "sketch_names.txt")
    with open(names_txt, "w") as f:
        f.write("\n".join(sorted(genomes)) + "\n")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "make_db.py"), dump, names_txt,
                           os.path.join(data, "cmash_db_n1000_k60.mlgdb"), "-n", str(N_SLOTS)])
    reads = reads_for(rng, seqs)
    with open(os.path.join(CASE, "reads.fq"), "w") as f:
        for i, r in enumerate(reads):
            f.write("@read%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    with gzip.GzipFile(os.path.join(CASE, "reads.fa.gz"), "wb", mtime=0) as f:
        for i, r in enumerate(reads):
            f.write((">read%d\n%s\n" % (i, r)).encode())
    runs = (("default", "reads.fq", []), ("strain_level", "reads.fq", ["--strain_level"]),
            ("cutoff_0.5", "reads.fq", ["--cutoff", "0.5"]), ("fasta_gz", "reads.fa.gz", ["--cutoff", "0.0"]))
    for tag, rfile, extra in runs:
        out = os.path.join(CASE, "expected_" + tag)
        os.makedirs(out)
        subprocess.check_call([sys.executable, REF, os.path.join(CASE, rfile), data, "--temp_dir", out,
                               "--keep_temp_files"] + extra, env=env, cwd=out)
        junk_files = ["log_sample", "60mers_intersection_dump.fa"]
        if tag != "default":              # the KMC databases of one run are kept: fixtures for the product's KMC reader
            junk_files += ["reads_60mers.kmc_pre", "reads_60mers.kmc_suf", "60mers_intersection.kmc_pre", "60mers_intersection.kmc_suf"]
        for junk in junk_files:
            p = os.path.join(out, junk)
            if os.path.exists(p):
                os.remove(p)
    os.remove(dump)
    print("fixtures written under", CASE)


if __name__ == "__main__":
    main()
