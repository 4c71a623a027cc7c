#! /usr/bin/env python
"""Turn a source .mlgdb (sketch keys, e.g. from scripts/make_db.py / make_sketch_db.py) into its BUILT form: the device
structures the query kernels use, so that every later `select_db.py` run loads the database with a file read instead of
rebuilding it on the GPU.  The analogue in the reference: MakeStreamingDNADatabase.py writes the trie next to the HDF5
once so that queries need not rebuild it (local_tests/retrain_and_test_metalign.sh:49).

    python scripts/build_db.py data/cmash_db_n1000_k60.src.mlgdb data/cmash_db_n1000_k60.mlgdb [--device 0]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metalign_b200.api import Context, Database  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("source")
    ap.add_argument("built")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    with Context(a.device) as ctx:
        t0 = time.perf_counter()
        db = Database.load(ctx, a.source)
        t1 = time.perf_counter()
        db.save(a.built)
        t2 = time.perf_counter()
        print("built %d genomes x %d slots (%d distinct k-mers) in %.2f s, wrote %s (%.2f GB) in %.2f s"
              % (db.G, db.n, db.n_distinct, t1 - t0, a.built, os.path.getsize(a.built) / 1e9, t2 - t1))
        db.close()


if __name__ == "__main__":
    main()
