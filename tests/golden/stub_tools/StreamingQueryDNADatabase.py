#!/usr/bin/env python3
"""Stub of CMash's StreamingQueryDNADatabase.py for the argv of scripts/select_db.py:73-76:
`<records.fa> <training.h5> <out.csv> 30-60-10 -c 0 -r 1000000 -v -f <prefilter.bf> --sensitive`.
Semantics = SURVEY.md 3.3 R4-R6 (oracle/oracle_py.py), Bloom prefilter modelled per MLG_STUB_GATE (exact | none).
The "HDF5" it opens is the stand-in written by tests/golden/make_ref_e2e_fixture.py: JSON with the sketch
names and k-mers (h5py is not installed here)."""
import argparse
import json
import os
import sys

import _common  # noqa: F401
import pandas as pd
from oracle import oracle_py

ap = argparse.ArgumentParser()
ap.add_argument("in_file"); ap.add_argument("reference_file"); ap.add_argument("out_file"); ap.add_argument("range")
ap.add_argument("-t", "--threads", type=int, default=0)
ap.add_argument("-c", "--containment_threshold", type=float, default=0.1)
ap.add_argument("-l", "--location_of_thresh", type=int, default=-1)
ap.add_argument("-r", "--reads_per_core", type=int, default=100000)
ap.add_argument("-f", "--filter_file", default=None)
ap.add_argument("-v", "--verbose", action="store_true")
ap.add_argument("--sensitive", action="store_true")
args = ap.parse_args()
if not args.sensitive:
    sys.exit("stub CMash: only --sensitive is modelled (scripts/select_db.py:76 always passes it)")
if args.filter_file is None or not os.path.exists(args.filter_file):
    sys.exit("stub CMash: prefilter file missing")
start, end, step = (int(x) for x in args.range.split("-"))
with open(args.reference_file) as f:
    ref = json.load(f)
K = ref["ksize"]
ks = [k for k in range(start, end + 1, step) if k <= K]
# import_multiple_from_single_hdf5 walks the groups sorted by name
order = sorted(range(len(ref["names"])), key=lambda i: ref["names"][i])
names = [ref["names"][i] for i in order]
sketches = [ref["sketches"][i] for i in order]
records = [s.upper() for s in _common.read_sequences(args.in_file, "fa")]
H = oracle_py.query_hits(records, sketches, ks, os.environ.get("MLG_STUB_GATE", "exact"))
num, den, ci = oracle_py.containment_table(H, sketches, ks, True)
df = pd.DataFrame({"k=%d" % k: [row[i] for row in ci] for i, k in enumerate(ks)}, index=names)
loc = df.columns[args.location_of_thresh]
out = df[df[loc] > args.containment_threshold].sort_values(loc, ascending=False)
out.to_csv(args.out_file, index=True, encoding="utf-8")
