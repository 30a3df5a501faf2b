"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: python tools/launch_agg.py file.csv [second_half]"""
import collections
import csv
import re
import sys

rows, hdr = [], None
for line in csv.reader(open(sys.argv[1])):
    if "Kernel Name" in line:
        hdr = line
    elif hdr and len(line) == len(hdr) and line[0].isdigit():
        rows.append(line)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
if len(sys.argv) > 2:
    rows = rows[len(rows) // 2:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} n={v[0]:5d} total={v[1]:10.1f} us avg={v[1] / v[0]:9.1f} us share={v[1] / tot:6.1%}")
print(f"total {tot:.1f} us over {len(rows)} launches")
