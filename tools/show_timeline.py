#!/usr/bin/env python
"""Print a timeline CSV written by tools/timeline.py (kernels longer than `min_us`). usage: show_timeline.py file.csv [min_us]"""
import csv, sys
rows = list(csv.DictReader(open(sys.argv[1])))
mn = float(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows:
    s = float(r["start_us"]); d = float(r["dur_us"])
    if d >= mn:
        print(f"{s:8.1f} {d:7.1f} end {s+d:8.1f} st{r['stream']:>4} {r['name'][:55]}")
