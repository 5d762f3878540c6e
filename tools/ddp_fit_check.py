"""Data parallel through the PUBLIC API (ADVICE r1 / VERDICT r1 weak #4).
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_fit_check.py --out /tmp/ddp.pt
    python tools/ddp_fit_check.py --compare /tmp/ddp.pt
The first command runs RandomCropLoader + model.fit() on R ranks (rank-aware loader shards, sync_bn, fp32, fused step, CUDA graph)
and saves rank 0's parameters; the second runs the same fit() in ONE process on the global batches and prints the largest
parameter difference.  No noise is injected: Philox is keyed by the global sample index, so both runs draw the same eps."""
import argparse
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vae-rs_b200")]
os.environ["SVRS_SYNC_BN"] = "1"
os.environ.setdefault("WANDB_MODE", "disabled")
import torch
import torch.distributed as dist

import models
import models.base as base_module
from dataset import RandomCropLoader, TileDataset, synthetic_tiles


class _Run:
    def log(self, *a, **k):
        pass

    def finish(self):
        pass


base_module.wandb.init = lambda *a, **k: _Run()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
os.chdir("/tmp")
GB, P, TILES, EPOCHS = 8, 64, 24, 2


def run(r, w):
    torch.manual_seed(0)
    m = models.Cond_SRVAE(2, P).to(dev)
    m.sync_bn = w > 1
    lr, hr = synthetic_tiles(TILES, 256, seed=4)
    train = RandomCropLoader(TileDataset(lr, hr), GB, P, dev, shuffle=True, seed=11, rank=r, world=w)
    val = RandomCropLoader(TileDataset(lr[:4], hr[:4]), 4, P, dev, shuffle=False, seed=12)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    m.fit(train_loader=train, val_loader=val, device=dev, optimizer=opt, epochs=EPOCHS, start_epoch=1, val_metrics_every=100)
    return {k: v.detach().float().cpu() for k, v in m.state_dict().items()}, float(m.gammax)


ap = argparse.ArgumentParser()
ap.add_argument("--out")
ap.add_argument("--compare")
args = ap.parse_args()
sd, gx = run(rank, world)
steps = EPOCHS * (TILES // GB)
if world > 1:
    flat = torch.cat([v.flatten() for k, v in sorted(sd.items()) if v.dtype.is_floating_point]).to(dev)
    f0 = flat.clone()
    dist.broadcast(f0, 0)
    spread = (flat - f0).abs().max()
    dist.all_reduce(spread, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({"sd": sd, "gammax": gx, "world": world}, args.out)
        print("DDP_FIT_CHECK ranks", world, "steps", steps, "param_spread_across_ranks", float(spread), "gammax", gx, flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)
if args.compare:
    ref = torch.load(args.compare)
    worst, worst_k = 0.0, None
    for k, v in sd.items():
        if not v.dtype.is_floating_point or k.endswith(("downsample.bias", "upsample.bias")):   # zero-gradient biases: Adam noise
            continue
        d = float((v - ref["sd"][k]).abs().max())
        if d > worst:
            worst, worst_k = d, k
    print(f"DDP_FIT_CHECK single process vs {ref['world']} ranks after {steps} steps (lr 1e-4, so a parameter moves <= {steps}e-4): "
          f"max |dp| {worst:.3e} ({worst_k}); gammax {gx:.8f} vs {ref['gammax']:.8f}", flush=True)
