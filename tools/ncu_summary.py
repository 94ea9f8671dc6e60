#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one row per distinct (kernel, grid) with the metrics DESIGN.md argues from, then the
per-source-line hot spots of selected kernels (tools/ncu_lines.py; needs the library of the same build).
usage: python tools/ncu_summary.py report.ncu-rep [kernel-substring-for-line-profile ...] > profiles/rNN_ncu_step_kernels.md"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, scale=1.0, fmt="{:.1f}"):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return "–"
    try:
        return fmt.format(float(r[i].replace(",", "")) * scale)
    except ValueError:
        return r[i]


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
print(f"# `ncu --set full --clock-control none` — {os.path.basename(rep)}\n")
print("Eager (no CUDA graph) config-2 step of `bench.py`, caches flushed per kernel replay by ncu: times are cold-cache and serialised. "
      "`tensor` = `sm__pipe_tensor_cycles_active` (% of active cycles), `hmma inst` = `sm__inst_executed_pipe_tensor_subpipe_hmma` (% of peak), "
      "`red sectors` = `lts__t_sectors_srcunit_tex_op_red.sum`, stalls = warps stalled per issue-active cycle (top three).\n")
print("| kernel | grid × block | regs | µs | warps active % | issue active % | tensor % | hmma inst % | DRAM rd / wr MB | red sectors M | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
seen = set()
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = (name, r[col["launch__grid_size"]])
    if key in seen:
        continue
    seen.add(key)
    stalls = sorted(((float(r[col[h]] or 0), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall_cols), reverse=True)
    stalls = ", ".join(f"{n} {v:.1f}" for v, n in stalls if n != "selected")
    stalls = ", ".join(stalls.split(", ")[:3])
    short = name.split("(")[0].replace("void ", "")
    print(f"| `{short}` | {r[col['launch__grid_size']]} × {r[col['launch__block_size']]} | {get(r, 'launch__registers_per_thread', fmt='{:.0f}')} | "
          f"{get(r, 'gpu__time_duration.sum')} | {get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
          f"{get(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} | {get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
          f"{get(r, 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', fmt='{:.2f}')} | "
          f"{get(r, 'dram__bytes_read.sum')} / {get(r, 'dram__bytes_write.sum')} | {get(r, 'lts__t_sectors_srcunit_tex_op_red.sum', 1e-6, '{:.2f}')} | {stalls} |")
for k in sys.argv[2:]:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, k, "2.0"], capture_output=True, text=True).stdout
    print(f"\n## per-source-line cost: `{k}` (lines with ≥ 2 % of the stall samples or executed warp instructions)\n\n```\n{out.strip()}\n```")
