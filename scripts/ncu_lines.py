#!/usr/bin/env python
"""Per-source-line instruction and stall profile of one kernel from an `ncu --set full --import-source on` report.

  python scripts/ncu_lines.py REPORT.ncu-rep [--top 40] [--kernel k1_minimizer]

Reads `ncu -i REPORT --page source --csv --print-source cuda,sass`: the CUDA-C view gives, per source line, the warp
instructions executed and the stall samples attributed to it.  Prints the lines by instruction count with their share
of the kernel, and the totals per file.
"""
import argparse
import csv
import io
import subprocess
from collections import defaultdict


def fnum(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--kernel", default="")
    a = ap.parse_args()
    cmd = ["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if a.kernel:
        cmd += ["-k", "regex:" + a.kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    per = []          # (file, line, src, inst, samples)
    fpath, hdr = "", None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or not r or not r[0].strip().isdigit():
            continue
        inst = fnum(r[hdr["Instructions Executed"]])
        smp = fnum(r[hdr["# Samples"]]) if "# Samples" in hdr else 0.0
        per.append((fpath, int(r[0]), r[1].strip(), inst, smp))
    tot = sum(p[3] for p in per) or 1.0
    tots = sum(p[4] for p in per) or 1.0
    byfile = defaultdict(float)
    for p in per:
        byfile[p[0]] += p[3]
    print("total warp instructions %.4g, stall samples %d" % (tot, tots))
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1]):
        print("  %-22s %5.1f %%" % (f, 100 * v / tot))
    print("%-20s %5s %7s %7s  %s" % ("file", "line", "inst %", "smpl %", "source"))
    for p in sorted(per, key=lambda p: -p[3])[: a.top]:
        print("%-20s %5d %7.2f %7.2f  %s" % (p[0], p[1], 100 * p[3] / tot, 100 * p[4] / tots, p[2][:110]))


if __name__ == "__main__":
    main()
