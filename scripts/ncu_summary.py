#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export of one probe-kernel launch and (optionally) write
profiles/k1_traffic.json, the DRAM-bytes-per-k-mer figure bench.py reports as `roofline.traffic`.

  python scripts/ncu_summary.py RAW.csv [--kmers N --genomes G --write-traffic profiles/k1_traffic.json]
"""
import argparse
import csv
import json
import os

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__thread_inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
]

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--kernel", default="k1_")
    ap.add_argument("--kmers", type=float, default=0)
    ap.add_argument("--genomes", type=int, default=0)
    ap.add_argument("--layout", type=int, default=1)
    ap.add_argument("--write-traffic", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "")
        if a.kernel not in name:
            continue
        print("kernel:", name[:110])
        for k in KEYS:
            if k in d:
                print("  %-62s %s %s" % (k, d[k], u[k]))
        st = [(k, num(d[k])) for k in hdr if k.startswith("smsp__average_warp") and "per_issue_active" in k]
        st = sorted([(k, v) for k, v in st if v is not None], key=lambda x: -x[1])
        print("  top stall reasons (warps per issue-active cycle):")
        for k, v in st[:8]:
            print("    %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        rd = num(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
        wr = num(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
        t = num(d["gpu__time_duration.sum"])
        tu = u["gpu__time_duration.sum"]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9) if tu in ("ns", "us", "ms", "s") else t * 1e-9
        print("  DRAM read+write %.3f GB in %.3f ms under ncu = %.0f GB/s" % ((rd + wr) / 1e9, t_s * 1e3, (rd + wr) / t_s / 1e9))
        if a.kmers:
            print("  per k-mer: %.2f DRAM bytes, %.1f warp instructions per 32 k-mers" % ((rd + wr) / a.kmers, num(d["smsp__inst_executed.sum"]) / (a.kmers / 32)))
        if a.write_traffic and a.kmers:
            out = {"dram_bytes_per_kmer": (rd + wr) / a.kmers, "genomes": a.genomes, "kernel": name.split("(")[0].split("::")[-1],
                   "source": os.path.relpath(a.raw), "kmers_in_capture": a.kmers, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "ncu_duration_ms": t_s * 1e3, "layout": a.layout,
                   "warp_instructions_per_32_kmers": num(d["smsp__inst_executed.sum"]) / (a.kmers / 32),
                   "issue_active_pct": num(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                   "warps_active_pct": num(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                   "dram_gbs_under_ncu": (rd + wr) / t_s / 1e9,
                   "top_stalls": {k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 3) for k, v in st[:5]}}
            with open(a.write_traffic, "w") as f:
                json.dump(out, f, indent=1)
            print("  wrote", a.write_traffic)
        break


if __name__ == "__main__":
    main()
