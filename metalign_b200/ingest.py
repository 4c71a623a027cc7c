"""Read-file ingest: FASTA / FASTQ (optionally gzip) -> batches of concatenated sequence bytes + offsets for
Query.push_ascii.  In the reference this job belongs to KMC (`kmc -fq|-fa`, scripts/select_db.py:46-52):
  -fq  4-line FASTQ records, the sequence is line 2 of each record
  -fa  FASTA with one sequence line per record; here every non-header line is taken as its own record, which
       is the same thing for single-line FASTA
Symbols outside ACGTacgt are left in place: the device treats them as N (they break the k-mer run).
Vectorised with numpy; a multi-threaded native parser is the next step.
"""
from __future__ import annotations

import gzip
from typing import Iterator, Tuple

import numpy as np


def _read_all(path: str) -> np.ndarray:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    return np.frombuffer(data, dtype=np.uint8)


def _line_table(buf: np.ndarray):
    """start / end (exclusive, newline and trailing '\\r' stripped) of every line"""
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate([[0], nl + 1]).astype(np.int64)
    ends = np.concatenate([nl, [buf.size]]).astype(np.int64)
    if starts.size and starts[-1] >= buf.size:          # file ends with a newline
        starts, ends = starts[:-1], ends[:-1]
    cr = (ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13)
    ends = ends - cr.astype(np.int64)
    return starts, ends


def sequence_lines(buf: np.ndarray, input_type: str):
    starts, ends = _line_table(buf)
    if input_type == "fastq":
        sel = np.arange(1, starts.size, 4)
    elif input_type == "fasta":
        nonempty = ends > starts
        first = buf[np.minimum(starts, max(buf.size - 1, 0))] if buf.size else np.zeros(0, np.uint8)
        sel = np.flatnonzero(nonempty & (first != ord(">")) & (first != ord(";")))
    else:
        raise ValueError("input_type must be 'fastq' or 'fasta'")
    return starts[sel], ends[sel]


def batches(path: str, input_type: str, reads_per_batch: int = 2_000_000) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    """yields (text uint8[total], off uint64[n+1]) per batch of reads"""
    buf = _read_all(path)
    s, e = sequence_lines(buf, input_type)
    for a in range(0, s.size, reads_per_batch):
        sb, eb = s[a:a + reads_per_batch], e[a:a + reads_per_batch]
        lens = eb - sb
        off = np.zeros(lens.size + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        total = int(off[-1])
        # gather: output position p of read i comes from input position sb[i] + (p - off[i])
        shift = np.repeat(sb - off[:-1].astype(np.int64), lens)
        text = buf[np.arange(total, dtype=np.int64) + shift]
        yield np.ascontiguousarray(text), off
    if s.size == 0:
        return


def detect_input_type(reads_path: str) -> str:
    """same rule as scripts/select_db.py:144-153: by extension, '.gz' ignored"""
    parts = reads_path.split(".")
    if parts[-1] == "gz":
        parts = parts[:-1]
    if parts[-1] in ("fq", "fastq"):
        return "fastq"
    if parts[-1] in ("fa", "fna", "fasta"):
        return "fasta"
    raise SystemExit("Could not auto-determine file type. Use --input_type.")
