#!/usr/bin/env python
"""Regenerates profiles/<tag>_sass_summary.md (default tag r02) from `cuobjdump -sass nerf-vo_b200/libnvo_b200.so` (runs on the CPU box)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "nerf-vo_b200", "libnvo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda names: subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
MN = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "REDG", "LDG", "STG", "SHFL", "HMMA"]
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m:
        op = m.group(1)
        kernels[cur]["instrs"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                kernels[cur][k] += 1
names = demangle(list(kernels))
out = [f"# {TAG} — SASS evidence (`cuobjdump -sass nerf-vo_b200/libnvo_b200.so`, sm_100a; regenerate with `python tools/sass_summary.py`)", "",
       "Per kernel: instruction count and the mnemonics that prove the Blackwell-native path — `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld,",
       "`UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk (TMA engine, 1-D), `SYNCS` = mbarrier arrive/try_wait, `REDG` = red.global (v2/v4 f32 reductions),",
       "`HMMA` = legacy mma.sync (absent everywhere). Full listings of every kernel, one gzip per source file: `" + TAG + "_sass/*.sass.gz` (tools/dump_sass.sh).", "",
       "| kernel | instrs | " + " | ".join(MN) + " |", "|---|---|" + "---|" * len(MN)]
for (mangled, c), nm in zip(kernels.items(), names):
    short = re.sub(r"\(.*", "", nm).strip() or mangled
    out.append(f"| `{short}` | {c['instrs']} | " + " | ".join(str(c[k]) for k in MN) + " |")
open(os.path.join(ROOT, "profiles", TAG + "_sass_summary.md"), "w").write("\n".join(out) + "\n")
print(len(kernels), "kernels")
