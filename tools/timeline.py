#!/usr/bin/env python
"""Kernel timeline of GRAPH-REPLAYED training steps via torch.profiler (CUPTI): per-step span, sum of kernel time,
concurrency histogram and the per-kernel-family totals.  Writes gpurun_out/timeline.csv (name, stream, start_us, dur_us).
    python tools/timeline.py [--eager]"""
import collections
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-vae-rs_b200"))
import torch
import models
from dataset import synthetic_tiles
from svrs_native.trainer import FusedCondTrainer
from torch.profiler import profile, ProfilerActivity

use_graph = "--eager" not in sys.argv
torch.manual_seed(0)
model = models.Cond_SRVAE(2, 64).cuda()
model.set_compute_dtype(torch.bfloat16)
tr = FusedCondTrainer(model)
lr, hr = synthetic_tiles(8, seed=3)
lr, hr = lr.cuda(), hr.cuda()
for _ in range(6):
    tr.step_tiles(hr, lr, patch_size=64, use_graph=use_graph)
torch.cuda.synchronize()
NS = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS):
        tr.step_tiles(hr, lr, patch_size=64, use_graph=use_graph)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in evs), key=lambda r: r[1])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "timeline.csv"), "w") as f:
    f.write("name,start_us,dur_us\n")
    for n, t0, d in rows:
        f.write(f"\"{n[:90]}\",{t0:.3f},{d:.3f}\n")
starts = [r[1] for r in rows if "step_increment" in r[0] or "step_begin" in r[0]]
print(len(rows), "device activities;", len(starts), "steps")
if len(starts) >= 3:
    s0, s1 = starts[1], starts[2]
    step = [r for r in rows if s0 <= r[1] < s1]
    print(f"step span {s1 - s0:.1f} us, {len(step)} activities, sum of kernel time {sum(r[2] for r in step):.1f} us")
    pts = []
    for _, t0, d in step:
        pts += [(t0, 1), (t0 + d, -1)]
    pts.sort()
    lvl, last, hist = 0, pts[0][0], collections.Counter()
    for t, d in pts:
        hist[lvl] += t - last
        last = t
        lvl += d
    print("concurrency histogram (us at each number of running kernels):", {k: round(v, 1) for k, v in sorted(hist.items())})
    fam = collections.defaultdict(lambda: [0.0, 0])
    for n, _, d in step:
        k = n.split("(")[0].replace("void ", "").replace("svrs::", "")
        k = k.split("<")[0]
        fam[k][0] += d
        fam[k][1] += 1
    for k, (us, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:9.1f} us  x{n:<3d} {k}")
    print("---- timeline of the step (start, dur, name)")
    for n, t0, d in step:
        print(f"{t0 - s0:8.1f} {d:7.1f} {n[:70]}")
