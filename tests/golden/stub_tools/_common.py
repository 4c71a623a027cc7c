"""TEST INFRASTRUCTURE shared by the stub executables in this directory.

The reference's hot path (scripts/select_db.py:43-76) is four subprocesses: `kmc`, `kmc_tools`, `kmc_dump`
and CMash's `StreamingQueryDNADatabase.py`.  None of them is vendored under /root/reference or installed here.
These stubs take the argv the UNMODIFIED reference builds, check it, and answer with the CPU oracle
(oracle/oracle_py.py), so that the reference's own glue -- argv construction, the dump -> FASTA rewrite at
select_db.py:61-65, the CSV consumption at :80-96, make_db_and_dbinfo -- runs end to end and its output
files can be committed as fixtures (tests/golden/make_ref_e2e_fixture.py).  They model the tools'
semantics as SURVEY.md section 3.3 states them; they are NOT the tools.
"""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def open_text(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")


def read_sequences(path, fmt):
    """fmt 'fq': 4-line FASTQ records; 'fa': header line + ONE sequence line per record (KMC -fa; multi-line
    FASTA would be -fm, which the reference never passes)."""
    seqs = []
    with open_text(path) as fh:
        lines = fh.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    if fmt == "fq":
        if len(lines) % 4:
            sys.exit("stub kmc: FASTQ with a truncated record")
        for i in range(0, len(lines), 4):
            if not lines[i].startswith("@") or not lines[i + 2].startswith("+"):
                sys.exit("stub kmc: malformed FASTQ record at line %d" % (i + 1))
            seqs.append(lines[i + 1])
    else:
        i = 0
        while i < len(lines):
            if not lines[i].startswith(">"):
                sys.exit("stub kmc: expected a FASTA header at line %d" % (i + 1))
            seqs.append(lines[i + 1] if i + 1 < len(lines) and not lines[i + 1].startswith(">") else "")
            i += 2 if i + 1 < len(lines) and not lines[i + 1].startswith(">") else 1
    return seqs
