#! /usr/bin/env python
"""Check a native database against the KMC database the reference ships beside its CMash training database:
the k-mers of data/cmash_db_n1000_k60_dump.kmc_* (scripts/select_db.py:44; made by `kmc -k60 -fa -ci0` from the dump of
every sketch, local_tests/retrain_and_test_metalign.sh:59-66) must be exactly the canonical forms of the non-empty
sketch slots of the native file.  Needs the SOURCE form of the .mlgdb (the built form does not carry the slots).

    python scripts/check_db_against_kmc.py data/cmash_db_n1000_k60_dump data/cmash_db_n1000_k60.src.mlgdb
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metalign_b200 import codec, dbformat, ingest  # noqa: E402


def main():
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    kkeys, info = ingest.read_kmc_database(sys.argv[1])
    h = dbformat.read_header(sys.argv[2])
    if info["k"] != h["K"]:
        sys.exit("k differs: KMC database %d, native database %d" % (info["k"], h["K"]))
    keys = dbformat.read_keys(sys.argv[2])
    keys = keys[keys[:, 0] != codec.EMPTY]
    canon = codec.canonical_keys(keys, h["K"])
    a = np.unique(canon.view([("hi", "<u8"), ("lo", "<u8")]))
    b = np.unique(np.ascontiguousarray(kkeys).view([("hi", "<u8"), ("lo", "<u8")]))
    only_native, only_kmc = np.setdiff1d(a, b).size, np.setdiff1d(b, a).size
    print("KMC database: %d k-mers (k=%d, layout %#x); native database: %d distinct canonical k-mers of %d non-empty slots"
          % (b.size, info["k"], info["version"], a.size, keys.shape[0]))
    if only_native or only_kmc:
        sys.exit("MISMATCH: %d k-mers only in the native database, %d only in the KMC database" % (only_native, only_kmc))
    print("identical k-mer sets")


if __name__ == "__main__":
    main()
