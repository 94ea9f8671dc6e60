#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + hottest SASS instructions (by stall samples) of each captured kernel.
usage: python tools/ncu_hot.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_active.avg", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
for r in rows[2:]:
    print("=====", r[hdr.index("Kernel Name")][:90])
    for h, u, v in zip(hdr, units, r):
        if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3):
            print(f"   {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for si, s in enumerate(secs):
    e = secs[si + 1] if si + 1 < len(secs) else len(rows)
    hdr = rows[s + 1]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [(int(r[isamp] or 0), r[ia], int(r[iex] or 0)) for r in rows[s + 2:e] if len(r) > isamp]
    tot = sum(d[0] for d in data) or 1
    print("===== hot SASS of", rows[s][1][:70], " samples", tot, " instrs", len(data))
    for i, (n, srcl, ex) in enumerate(data):
        if n > tot * minpct / 100:
            print(f"{i:5d} {100*n/tot:5.1f}% ex={ex:9d}  {srcl[:120]}")
