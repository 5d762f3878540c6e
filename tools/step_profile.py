"""Per-launch device-time table of one eager training step (CUDA events around every C-ABI call).
    python tools/step_profile.py [--dtype bf16] [--patches 128] [--top 40]
"""
import argparse
import os
os.environ.setdefault("SVRS_WGRAD_STREAM", "0")   # per-call event brackets only see the current stream
os.environ.setdefault("SVRS_BRANCH_STREAMS", "0")
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200")):
    sys.path.insert(0, p)
import torch

import models
from svrs_native import profile as prof
from svrs_native.lib import lib
from svrs_native.trainer import FusedCondTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--patches", type=int, default=128)
ap.add_argument("--top", type=int, default=45)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = models.Cond_SRVAE(2, 64).to(dev).train()
model.set_compute_dtype(torch.bfloat16 if a.dtype == "bf16" else torch.float32)
tr = FusedCondTrainer(model)
x = torch.rand(a.patches, 4, 64, 64, device=dev)
y = torch.rand(a.patches, 4, 32, 32, device=dev)
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
lib.timing = []
lib.timing_pad_cycles = int(os.environ.get("SVRS_PAD_CYCLES", "150000"))
steps = 3
for _ in range(steps):
    tr.step(x, y)
torch.cuda.synchronize()
rec, lib.timing = lib.timing, None
names = {n: [an for _, an in args] for n, (_, args) in lib.protos.items()}
agg = defaultdict(lambda: [0.0, 0, 0.0])
bykern = defaultdict(lambda: [0.0, 0, 0.0])
for name, args, e0, e1, kernels in rec:
    kk = bykern[kernels or name]
    kk[0] += e0.elapsed_time(e1) / steps
    kk[1] += 1
    kk[2] += prof._flops(name, args) / steps
    d = dict(zip(names[name], args))
    key = name.replace("svrs_", "")
    if "conv" in name:
        key += f" N{d['N']} {d['H']}x{d['W']} {d['Cin']}->{d['Cout']}" + (f" k{d['ksize']}" if "ksize" in d else "") + f" [{kernels}]"
    elif "M" in d and "C" in d:
        key += f" M{d['M']} C{d['C']}"
    r = agg[key]
    r[0] += e0.elapsed_time(e1) / steps
    r[1] += 1
    r[2] += prof._flops(name, args) / steps
tot = sum(v[0] for v in agg.values())
print(f"total kernel time per step (eager, event-timed): {tot:.3f} ms over {sum(v[1] for v in agg.values()) // steps} launches")
for k, (ms, n, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 and fl > 0 else 0
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n // steps:<3d} {tf:7.1f} TF/s  {k}")
byfn = defaultdict(lambda: [0.0, 0])
for name, args, e0, e1, _k in rec:
    byfn[name.replace("svrs_", "")][0] += e0.elapsed_time(e1) / steps
    byfn[name.replace("svrs_", "")][1] += 1
print("---- by kernel (library launch trace)")
for k, (ms, n, fl) in sorted(bykern.items(), key=lambda kv: -kv[1][0]):
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 and fl > 0 else 0
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n // steps:<3d} {tf:7.1f} TF/s  {k}")
print("---- by C-ABI function")
for k, (ms, n) in sorted(byfn.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n // steps:<3d} {k}")

# ---- pure GPU time per launch: re-issue every recorded call 8x inside a captured CUDA graph (no host gaps)
if os.environ.get("SVRS_REPLAY", "0") == "1":
    first = [r for r in rec[: len(rec) // steps]]
    keep = [x, y, tr]          # keep buffers alive; stale activations stay mapped in the caching allocator
    reps = 8
    table = []
    side = torch.cuda.Stream()
    for name, args, _, _, _k in first:
        fn = getattr(lib, name.replace("svrs_", ""))
        args = list(args)
        with torch.cuda.stream(side):
            args[-1] = side.cuda_stream
            fn(*args)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                args[-1] = torch.cuda.current_stream().cuda_stream
                for _ in range(reps):
                    fn(*args)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
        d = dict(zip(names[name], args))
        key = name.replace("svrs_", "")
        if "conv" in name:
            key += f" N{d['N']} {d['H']}x{d['W']} {d['Cin']}->{d['Cout']}" + (f" k{d['ksize']}" if "ksize" in d else "") + f" [{kernels}]"
        elif "M" in d and "C" in d:
            key += f" M{d['M']} C{d['C']}"
        table.append((e0.elapsed_time(e1) / reps, key, prof._flops(name, tuple(args))))
    tot2 = sum(t[0] for t in table)
    print(f"==== pure GPU time (graph-replayed, warm L2): {tot2:.3f} ms per step over {len(table)} launches")
    agg2 = defaultdict(lambda: [0.0, 0, 0.0])
    for ms, key, fl in table:
        agg2[key][0] += ms; agg2[key][1] += 1; agg2[key][2] += fl
    for k, (ms, n, fl) in sorted(agg2.items(), key=lambda kv: -kv[1][0])[:a.top]:
        tf = fl / (ms * 1e-3) / 1e12 if fl > 0 else 0
        print(f"{ms * 1e3:9.1f} us {100 * ms / tot2:5.1f}%  x{n:<3d} {tf:7.1f} TF/s  {k}")
    byfn2 = defaultdict(float)
    for ms, key, fl in table:
        byfn2[key.split(" ")[0]] += ms
    print("---- pure GPU time by function")
    for k, ms in sorted(byfn2.items(), key=lambda kv: -kv[1]):
        print(f"{ms * 1e3:9.1f} us {100 * ms / tot2:5.1f}%  {k}")
