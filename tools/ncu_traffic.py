#!/usr/bin/env python
"""Per-launch DRAM traffic of one kernel from an `ncu --set full` capture exported with `--page raw --csv`.

    python tools/ncu_traffic.py gpurun_out/conv_tc_r01_raw.csv conv_tc_kernel > profiles/traffic_r01.json

Writes {"kernel", "launches", "dram_bytes_per_launch", "mean_us", per-launch rows}; bench.py copies
dram_bytes_per_launch into roofline.traffic when the dominant kernel of the run is the same kernel."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
kern = sys.argv[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, k):
    return float(r[ix[k]].replace(",", "")) * scale.get(units[ix[k]], 1.0)


out = []
for r in data:
    if kern not in r[ix["Kernel Name"]]:
        continue
    out.append({
        "grid": r[ix["Grid Size"]],
        "us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")),
        "dram_read_bytes": val(r, "dram__bytes_read.sum"),
        "dram_write_bytes": val(r, "dram__bytes_write.sum"),
        "tensor_pipe_pct": float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
        "l2_throughput_pct": float(r[ix["lts__throughput.avg.pct_of_peak_sustained_elapsed"]]),
    })
n = len(out)
print(json.dumps({
    "kernel": kern, "launches": n,
    "command": "ncu --set full --clock-control none --profile-from-start off -k regex:%s -c 54 python tools/ncu_step.py" % kern,
    "dram_bytes_per_launch": sum(o["dram_read_bytes"] + o["dram_write_bytes"] for o in out) / max(n, 1),
    "mean_us": sum(o["us"] for o in out) / max(n, 1),
    "per_launch": out}, indent=1))
