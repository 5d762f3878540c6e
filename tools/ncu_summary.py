#!/usr/bin/env python
"""Reduce `ncu -i X.ncu-rep --page raw --csv` of one profiled step (tools/ncu_step.py) to a per-launch table and a
per-kernel-family summary: time, DRAM bytes and % of peak, tensor-pipe %, L2 %, SM issue %, achieved occupancy.
    python tools/ncu_summary.py gpurun_out/full_raw.csv > profiles/ncu_full_rNN_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def num(r, k, default=0.0):
    if k not in ix or r[ix[k]] in ("", "n/a"):
        return default
    return float(r[ix[k]].replace(",", "")) * scale.get(units[ix[k]], 1.0)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("svrs::", "")


M = dict(us="gpu__time_duration.sum", rd="dram__bytes_read.sum", wr="dram__bytes_write.sum",
         dram="gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         tensor="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
         l2="lts__throughput.avg.pct_of_peak_sustained_elapsed",
         sm="sm__throughput.avg.pct_of_peak_sustained_elapsed",
         occ="sm__warps_active.avg.pct_of_peak_sustained_active")
fam = defaultdict(lambda: defaultdict(float))
print("# per launch: kernel,grid,block,us,dram_MB,dram_pct,tensor_pct,l2_pct,sm_pct,occupancy_pct")
for r in data:
    k = short(r[ix["Kernel Name"]])
    v = {a: num(r, b) for a, b in M.items()}
    mb = (v["rd"] + v["wr"]) / 1e6
    print(f"\"{k}\",\"{r[ix['Grid Size']]}\",\"{r[ix['Block Size']]}\",{v['us']:.2f},{mb:.2f},{v['dram']:.1f},{v['tensor']:.1f},{v['l2']:.1f},{v['sm']:.1f},{v['occ']:.1f}")
    f = fam[re.sub(r"<.*", "", k)]
    f["n"] += 1
    f["us"] += v["us"]
    f["mb"] += mb
    for a in ("dram", "tensor", "l2", "sm"):
        f[a] += v[a] * v["us"]          # time-weighted
tot = sum(f["us"] for f in fam.values())
print(f"# per kernel family (time-weighted utilisation): family,launches,total_us,share,dram_MB,dram_GBs,dram_pct,tensor_pct,l2_pct,sm_pct   [total {tot:.0f} us]")
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    u = f["us"]
    print(f"\"{k}\",{int(f['n'])},{u:.1f},{u / tot:.4f},{f['mb']:.1f},{f['mb'] / u * 1e3:.0f},{f['dram'] / u:.1f},{f['tensor'] / u:.1f},{f['l2'] / u:.1f},{f['sm'] / u:.1f}")
