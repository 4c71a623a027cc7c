"""metalign_b200 -- B200-native database-selection hot path of Metalign (select_db.py's KMC + CMash stage).

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/metalign_b200.h)
  _lib.py          ctypes loader (no CPU fallback)
  api.py           Context / Database / Query objects over the C ABI
  codec.py         host encodings (k-mer keys, 2-bit packed reads)
  dbformat.py      native .mlgdb database file
  ingest.py        FASTA / FASTQ(.gz) reader producing batches for Query.push_*
  cmash_tail.py    the pandas filter / sort / to_csv tail CMash ends with (kept on the host verbatim)
  select_db.py     drop-in for scripts/select_db.py of the reference (same flags, same files out)
  metalign.py      drop-in for scripts/metalign.py (calls select_db, then the reference's mapper if present)
  dist.py          read-sharded multi-GPU run: one process per GPU, one all-reduce of the counter table
"""
__version__ = "0.1.0"
