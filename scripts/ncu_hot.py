#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics + the hottest SASS instructions with their stall reasons.

usage: python scripts/ncu_hot.py report.ncu-rep [top_n] [kernel_regex]
"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed_pipe_xu.sum",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        print("== kernel:", name[:110])
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                print(f"   {h:75s} {v} {u}")
    src = run([rep, "--page", "source", "--csv"])
    # several kernels may be concatenated: split on the "Kernel Name" header rows
    blocks, cur = [], []
    for row in csv.reader(io.StringIO(src)):
        if row and row[0] == "Kernel Name":
            if cur:
                blocks.append(cur)
            cur = [row]
        else:
            cur.append(row)
    if cur:
        blocks.append(cur)
    for blk in blocks[:1]:
        print("== SASS hot spots:", blk[0][1][:100])
        h = blk[1]
        isamp = h.index("# Samples")
        iexec = h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        rows = [r for r in blk[2:] if len(r) > isamp and r[isamp].isdigit()]
        total = sum(int(r[isamp]) for r in rows) or 1
        print(f"   total samples {total}, SASS instructions {len(rows)}")
        agg = {}
        for r in rows:
            for i, c in stall_cols:
                agg[c] = agg.get(c, 0) + int(r[i] or 0)
        print("   stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        idx = {id(r): n for n, r in enumerate(rows)}
        for r in sorted(rows, key=lambda r: -int(r[isamp]))[:top]:
            st = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:3]
            sts = " ".join(f"{c}={n}" for n, c in st if n)
            print(f"   #{idx[id(r)]:4d} {100.0 * int(r[isamp]) / total:5.1f}% exec={r[iexec]:>9s} {r[1].strip()[:70]:70s} {sts}")


if __name__ == "__main__":
    main()
