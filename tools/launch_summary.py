#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
    python tools/launch_summary.py gpurun_out/launches_bench_r1.csv "command line" > profiles/launches_bench_r1_summary.txt"""
import csv, sys
from collections import defaultdict
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    name = r["Kernel Name"].split("(")[0][:70]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(cold-cache, serialised launches: compare SHARES, not absolutes)")
print("%-72s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %12.1f %6.1f%%" % (k, n, us, 100 * us / tot))
