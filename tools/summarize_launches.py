#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share.
usage: python tools/summarize_launches.py launches.csv [--md]"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = collections.OrderedDict()
for n, u in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += u
print(f"window = {len(rows)} launches, total {tot:.1f} us\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {n[:70]} | {c} | {u:.1f} | {u/c:.1f} | {100*u/tot:.1f}% |")
