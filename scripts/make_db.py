#! /usr/bin/env python
"""Build the native database file (.mlgdb) from a FASTA dump of the CMash sketch k-mers.

Input is what local_tests/dump_kmers.py of the reference writes from the training HDF5 (one record per sketch
slot, in CountEstimator order, empty sequence for an unused slot) plus the list of sketch names in the same
(sorted-basename) order:
    python scripts/make_db.py dump.fa names.txt out.mlgdb [-n 1000] [-k 60] [--k_range 30-60-10]
Reading the HDF5 directly needs h5py, which this image does not have.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metalign_b200 import codec, dbformat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump_fasta")
    ap.add_argument("names")
    ap.add_argument("out")
    ap.add_argument("-n", type=int, default=1000)
    ap.add_argument("-k", type=int, default=60)
    ap.add_argument("--k_range", default="30-60-10")
    a = ap.parse_args()
    lo, hi, step = (int(x) for x in a.k_range.split("-"))
    ks = [k for k in range(lo, hi + 1, step) if k <= a.k]
    names = [ln.strip() for ln in open(a.names) if ln.strip()]
    G = len(names)
    try:
        keys = dbformat.keys_from_dump_fasta(a.dump_fasta, a.k, expect_records=G * a.n)
    except ValueError as e:
        sys.exit(str(e))
    dbformat.write(a.out, keys, names, G, a.n, a.k, ks)
    print("wrote %s: %d genomes x %d slots, K=%d, ks=%s" % (a.out, G, a.n, a.k, ks))


if __name__ == "__main__":
    main()
