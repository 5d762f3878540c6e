#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total/mean time and the
kernel's SHARE of all profiled GPU time (ncu times are cold-cache and serialised, so shares - not absolutes - are what
is comparable with bench.py's CUDA-event numbers).

    python tools/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/launches_rNN_summary.csv
"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)                       # drop the argument list, keep template arguments
    name = name.replace("at::native::", "torch:")
    if len(name) > 110:
        name = name[:107] + "..."
    a = agg[name]
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e3                   # us
tot = sum(v[1] for v in agg.values())
for c in sys.argv[2:]:
    print("# " + c)
print(f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms of GPU time in the list")
print("kernel,launches,total_us,mean_us,share")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{n},{us:.1f},{us / n:.2f},{us / tot:.4f}")
