#! /usr/bin/env python
"""Build the native database file (.mlgdb) straight from a CMash training HDF5 (the file select_db.py:69 of the reference
hands to StreamingQueryDNADatabase.py, e.g. data/cmash_db_n1000_k60.h5).

    python scripts/make_db_from_h5.py cmash_db_n1000_k60.h5 out.mlgdb [--k_range 30-60-10]

Reads the file with h5py where it is installed and with the built-in minimal reader (metalign_b200/h5min.py: the HDF5
flavour h5py writes by default) where it is not -- this repo's build image has no h5py.  Layout read: group `CountEstimators`, one sub-group per genome keyed by
the basename of its training file, dataset `kmers` (n fixed-length byte strings, '' for an unused slot), attribute `ksize`;
genomes in sorted-key order, which is the order MinHash.import_multiple_from_single_hdf5 -- and therefore
dump_kmers.py:7-14 and the query script -- see them in.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metalign_b200 import codec, dbformat  # noqa: E402


def _open(path):
    """(file object with h5py's mapping interface, function that reads a dataset)"""
    try:
        import h5py
        return h5py.File(path, "r"), (lambda d: d[...])
    except ImportError:
        from metalign_b200 import h5min
        return h5min.H5File(path), (lambda d: d.read())


def read_h5(path):
    f, read = _open(path)
    try:
        grp = f["CountEstimators"]
        names = sorted(grp.keys())
        K = n = None
        slots = []
        for name in names:
            g = grp[name]
            kmers = [k.decode() if isinstance(k, bytes) else str(k) for k in read(g["kmers"])]
            attrs = g.attrs
            k_here = int(attrs["ksize"]) if "ksize" in attrs else max((len(x) for x in kmers), default=0)
            if K is None:
                K, n = k_here, len(kmers)
            if k_here != K or len(kmers) != n:
                sys.exit("sketch %s has ksize %d / %d slots, expected %d / %d" % (name, k_here, len(kmers), K, n))
            slots.append(kmers)
    finally:
        f.close()
    return names, slots, K, n


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("h5")
    ap.add_argument("out")
    ap.add_argument("--k_range", default="30-60-10")
    a = ap.parse_args(argv)
    names, slots, K, n = read_h5(a.h5)
    lo, hi, step = (int(x) for x in a.k_range.split("-"))
    ks = [k for k in range(lo, hi + 1, step) if k <= K]
    buf = np.zeros((len(names) * n, K), dtype=np.uint8)
    for g, kmers in enumerate(slots):
        for j, kmer in enumerate(kmers):
            if kmer:
                if len(kmer) != K:
                    sys.exit("k-mer of length %d in sketch %s (ksize %d)" % (len(kmer), names[g], K))
                buf[g * n + j] = np.frombuffer(kmer.upper().encode(), dtype=np.uint8)
    keys = codec.ascii_slots_to_keys(buf, K)
    dbformat.write(a.out, keys.reshape(-1), names, len(names), n, K, ks)
    print("wrote %s: %d genomes x %d slots, K=%d, ks=%s" % (a.out, len(names), n, K, ks))


if __name__ == "__main__":
    main()
