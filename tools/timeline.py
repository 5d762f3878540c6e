#!/usr/bin/env python
"""Kernel timeline of graph-replayed training steps via torch.profiler (CUPTI): per-stream busy time, GPU idle gaps and
the kernels on the critical (last-finishing) chain.  Writes gpurun_out/timeline.csv (name, stream, start_us, dur_us)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-vae-rs_b200"))
import torch
import models
from dataset import synthetic_tiles, grid_patch_normalize
from svrs_native.trainer import FusedCondTrainer
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
model = models.Cond_SRVAE(2, 64).cuda()
model.set_compute_dtype(torch.bfloat16)
tr = FusedCondTrainer(model, torch.optim.Adam(model.parameters(), lr=1e-4), compute_dtype=torch.bfloat16)
lr, hr = synthetic_tiles(8, seed=3)
lr, hr = lr.cuda(), hr.cuda()
y, x = grid_patch_normalize(lr, 32), grid_patch_normalize(hr, 64)
for _ in range(5):
    tr.step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(x, y)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.name, getattr(e, "stream", -1) if hasattr(e, "stream") else -1, e.time_range.start, e.time_range.end - e.time_range.start) for e in evs), key=lambda r: r[2])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "timeline.csv"), "w") as f:
    f.write("name,stream,start_us,dur_us\n")
    for n, s, t0, d in rows:
        f.write(f"\"{n[:90]}\",{s},{t0:.3f},{d:.3f}\n")
print(len(rows), "device activities")
if rows:
    t_begin, t_end = rows[0][2], max(r[2] + r[3] for r in rows)
    print(f"span {(t_end - t_begin) / 3:.1f} us per step; sum of kernel time {sum(r[3] for r in rows) / 3:.1f} us per step")
