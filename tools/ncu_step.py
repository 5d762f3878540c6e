"""One eager, single-stream training step of the bench workload between cudaProfilerStart / cudaProfilerStop, for
    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/full python tools/ncu_step.py
(every kernel of the step is captured once; ncu serialises the launches anyway).  --workload as bench.py."""
import argparse
import os
os.environ.setdefault("SVRS_WGRAD_STREAM", "0")
os.environ.setdefault("SVRS_BRANCH_STREAMS", "0")
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200")):
    sys.path.insert(0, p)
import torch

import models
from dataset import synthetic_tiles
from svrs_native.trainer import FusedCondTrainer, FusedVaeTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=8)
ap.add_argument("--model", default="cond")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = (models.Cond_SRVAE(2, 64) if a.model == "cond" else models.VAE(2, 64)).to(dev).train()
model.set_compute_dtype(torch.bfloat16)
tr = (FusedCondTrainer if a.model == "cond" else FusedVaeTrainer)(model)
lr, hr = synthetic_tiles(a.tiles, 256, seed=100)
lr, hr = lr.to(dev), hr.to(dev)
args = (hr, lr) if a.model == "cond" else (hr,)
for _ in range(3):
    tr.step_tiles(*args, patch_size=64, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step_tiles(*args, patch_size=64, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
