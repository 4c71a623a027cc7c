#!/usr/bin/env python
"""Generates tests/golden/select_case/: a toy data/ directory, a CMash-style results CSV, and the outputs the
UNMODIFIED reference produces from them (scripts/select_db.py of /root/reference, run here through its
`--cmash_results` seam, which skips KMC/CMash but exercises read_dbinfo, the cutoff / one-strain-per-species
selection and make_db_and_dbinfo).  The fixtures travel to the GPU box; /root/reference does not.

Run (in the build container only):  python tests/golden/make_select_fixtures.py
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "select_case")
REF = "/root/reference/scripts/select_db.py"

LINEAGE = {
    # taxid: (name lineage, taxid lineage)  -- 7 pipes each (data/spec_db_info.txt:49-55)
    "562.1": ("Bacteria|Proteobacteria|Gammaproteobacteria|Enterobacterales|Enterobacteriaceae|Escherichia|Escherichia coli|E. coli K-12",
              "2|1224|1236|91347|543|561|562|562.1"),
    "562.2": ("Bacteria|Proteobacteria|Gammaproteobacteria|Enterobacterales|Enterobacteriaceae|Escherichia|Escherichia coli|E. coli O157",
              "2|1224|1236|91347|543|561|562|562.2"),
    "1280.1": ("Bacteria|Firmicutes|Bacilli|Bacillales|Staphylococcaceae|Staphylococcus|Staphylococcus aureus|S. aureus NCTC",
               "2|1239|91061|1385|90964|1279|1280|1280.1"),
    "10239.1": ("Viruses|||||||Some phage", "10239|||||||10239.1"),
    "10239.2": ("Viruses|||||||Other phage", "10239|||||||10239.2"),
    "4932.1": ("Eukaryota|Ascomycota|Saccharomycetes|Saccharomycetales|Saccharomycetaceae|Saccharomyces|Saccharomyces cerevisiae|S288C",
               "2759|4890|4891|4892|4893|4930|4932|4932.1"),
}
ACCESSIONS = {"562.1": ["NC_000913.3"], "562.2": ["NC_002695.2", "NC_002128.1"], "1280.1": ["NZ_LS483365.1"],
              "10239.1": ["NC_001416.1"], "10239.2": ["NC_001604.1"], "4932.1": ["NC_001133.9", "NC_001134.8"]}
CSV_ROWS = [  # name, k=30, k=40, k=50, k=60  (already in CMash's output order: k=60 descending)
    ("taxid_562_2_genomic.fna.gz", 0.93, 0.91, 0.9, 0.88),
    ("taxid_562_1_genomic.fna.gz", 0.9, 0.85, 0.8, 0.75),
    ("taxid_10239_1_genomic.fna.gz", 0.5, 0.45, 0.41, 0.4),
    ("taxid_10239_2_genomic.fna.gz", 0.3, 0.25, 0.21, 0.2),
    ("taxid_1280_1_genomic.fna.gz", 0.05, 0.03, 0.02, 0.01),
    ("taxid_4932_1_genomic.fna.gz", 0.02, 0.01, 0.006, 0.004),
]


def main():
    if not os.path.exists(REF):
        sys.exit("the reference is not mounted here; fixtures can only be regenerated in the build container")
    shutil.rmtree(CASE, ignore_errors=True)
    data = os.path.join(CASE, "data")
    os.makedirs(os.path.join(data, "organism_files"))
    with open(os.path.join(data, "db_info.txt"), "w") as f:
        f.write("Accession\tLength\tTaxID\tLineage\tTaxID_Lineage\n")
        for taxid, accs in ACCESSIONS.items():
            for i, acc in enumerate(accs):
                f.write("\t".join([acc, str(1000 + 37 * i + len(taxid)), taxid, LINEAGE[taxid][0], LINEAGE[taxid][1]]) + "\n")
    for taxid, accs in ACCESSIONS.items():
        name = "taxid_" + taxid.replace(".", "_") + "_genomic.fna.gz"
        with gzip.open(os.path.join(data, "organism_files", name), "wt") as f:
            for i, acc in enumerate(accs):
                f.write(">%s synthetic record %d of %s\n" % (acc, i, taxid))
                f.write(("ACGT" * 20 + "\n") * (2 + i) + "GATTACA" * (len(taxid)) + "\n")
    with open(os.path.join(CASE, "cmash_query_results.csv"), "w") as f:
        f.write(",k=30,k=40,k=50,k=60\n")
        for row in CSV_ROWS:
            f.write(",".join([row[0]] + [repr(x) for x in row[1:]]) + "\n")
    open(os.path.join(CASE, "reads.fq"), "w").write("@r\nACGT\n+\nIIII\n")
    for tag, extra in (("default", []), ("strain_level", ["--strain_level"]), ("cutoff_0.3", ["--cutoff", "0.3"]),
                       ("cutoff_0", ["--cutoff", "0.0"])):
        out = os.path.join(CASE, "expected_" + tag)
        os.makedirs(out)
        subprocess.check_call([sys.executable, REF, os.path.join(CASE, "reads.fq"), data,
                               "--cmash_results", os.path.join(CASE, "cmash_query_results.csv"),
                               "--temp_dir", out, "--db", os.path.join(out, "cmashed_db.fna"),
                               "--dbinfo_out", os.path.join(out, "subset_db_info.txt")] + extra)
    print("fixtures written under", CASE)


if __name__ == "__main__":
    main()
