#!/usr/bin/env python
"""Summarise an ncu report offline: key raw metrics + hottest source lines by warp-stall samples.
usage: python tools/ncu_hot.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
for d in rows[2:]:
    print("===", d[hdr.index("Kernel Name")][:70])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:70s} {d[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(float(d[i]), h) for i, h in enumerate(hdr) if "average_warps_issue_stalled" in h and "per_issue_active" in h]
    for v, h in sorted(st, reverse=True)[:7]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


lines = []
for r in csv.reader(io.StringIO(src)):
    if len(r) > 10 and r[0].isdigit():
        lines.append((num(r[6]), num(r[7]), int(r[0]), r[1].strip()))
tot = sum(x[0] for x in lines) or 1
toti = sum(x[1] for x in lines) or 1
print(f"--- hottest source lines by stall samples (of {tot} samples; {toti} warp instructions)")
for n, ni, ln, txt in sorted(lines, reverse=True)[:top]:
    print(f"  {100.0 * n / tot:5.1f}% smp {100.0 * ni / toti:5.1f}% inst  L{ln:<5d} {txt[:100]}")
print("--- most executed source lines")
for n, ni, ln, txt in sorted(lines, key=lambda x: -x[1])[:top]:
    print(f"  {100.0 * ni / toti:5.1f}% inst {100.0 * n / tot:5.1f}% smp  L{ln:<5d} {txt[:100]}")
