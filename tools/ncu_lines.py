#!/usr/bin/env python
"""Per-source-line cost of one kernel: joins the SASS page of an ncu report (stall samples, instructions executed) with the line
table of the same kernel in the built library (nvdisasm -g), by instruction order.  The library must be the build the report was taken from.
usage: python tools/ncu_lines.py report.ncu-rep <kernel substring> [min_pct] [launch index]"""
import csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
secs = [s for s in secs if kname in rows[s][1]]
s = secs[which]
all_secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
e = min([x for x in all_secs if x > s] + [len(rows)])
hdr = rows[s + 1]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [(r[ia].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in rows[s + 2:e] if len(r) > isamp]
full = rows[s][1]
print("kernel:", full[:100], " SASS instructions:", len(sass))
# line table from the library
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "nerf-vo_b200", "libnvo_b200.so")], cwd=tmp, capture_output=True)
base = re.sub(r"<.*", "", full.replace("void ", "")).split("(")[0].strip()
found = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if base not in dis:
        continue
    # split into functions
    funcs = re.split(r"\n\s*\.section\s+\.text\.", dis)
    for fn in funcs[1:]:
        name = fn.split(",")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if base not in dem:
            continue
        lines, cur = [], None
        for ln in fn.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur)
        if len(lines) == len(sass):
            found = (dem, lines)
            break
        else:
            print(f"  candidate {dem[:80]}: {len(lines)} instructions (report has {len(sass)})")
    if found:
        break
if not found:
    sys.exit("no function of the built library matches the report's instruction count: rebuild the commit the report was taken from")
dem, lines = found
agg = {}
for (txt, n, ex), loc in zip(sass, lines):
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += n; a[1] += ex; a[2] += 1
tot_s = sum(a[0] for a in agg.values()) or 1
tot_e = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}, warp instructions executed {tot_e}")
srcs = {}
for loc, a in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if loc is None or (100 * a[0] / tot_s < minpct and 100 * a[1] / tot_e < minpct):
        continue
    f, l = loc
    if f not in srcs:
        p = os.path.join(ROOT, "nerf-vo_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:110] if l - 1 < len(srcs[f]) else ""
    print(f"{f}:{l:4d} samples {100 * a[0] / tot_s:5.1f}%  executed {100 * a[1] / tot_e:5.1f}% ({a[1]:9d})  sass {a[2]:4d} | {text}")

# call sites of inlined helpers: contiguous SASS runs attributed to a header line, labelled by the nearest preceding line of the kernel's own file
if len(sys.argv) > 5 and sys.argv[5] == "sites":
    main_file = max(((f, sum(a[2] for (ff, _), a in agg.items() if ff == f)) for f in {k[0] for k in agg if k}), key=lambda t: t[1])[0]
    runs, prev_main, cur = [], None, None
    for (txt, n, ex), loc in zip(sass, lines):
        if loc and loc[0] == main_file:
            prev_main = loc
            cur = None
            continue
        if cur is None or cur[0] != prev_main:
            cur = [prev_main, loc, 0, 0]
            runs.append(cur)
        cur[2] += n; cur[3] += ex
    print(f"--- inlined-helper call sites (after {main_file} line) with >= {minpct}% of the samples")
    for pm, loc, n, ex in runs:
        if 100 * n / tot_s >= minpct:
            l = pm[1] if pm else 0
            text = srcs.get(main_file, [""] * l)[l - 1].strip()[:100] if pm and main_file in srcs else ""
            print(f"  after {main_file}:{l:4d} ({loc[0] if loc else '?'}:{loc[1] if loc else 0}) samples {100 * n / tot_s:5.1f}%  executed {ex:9d} | {text}")
