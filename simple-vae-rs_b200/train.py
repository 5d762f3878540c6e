"""Command-line entry with the reference's flags (train.py:83-148) on the B200 kernel path.

    python train.py --dataset synthetic --model_type Cond_SRVAE -cr 2 --patch_size 64 --batch_size 128 --epochs 2

Extra flags (not in the reference): --dtype {fp32,bf16}.  Multi-GPU: launch with torchrun; each rank trains on
its shard and gradients are all-reduced over NCCL (svrs_native.trainer)."""
import argparse
import os
import time

import torch

import callbacks
import models
from dataset import init_dataloader

# (flags, argparse keywords) - the reference's CLI surface, one row per option
_FLAGS = (
    (("--pre_epochs",), dict(type=int, default=20, help="(unused, kept for CLI compatibility)")),
    (("--epochs",), dict(type=int, default=200)),
    (("--dataset",), dict(type=str, default="s2v")),
    (("--batch_size",), dict(type=int, default=16)),
    (("--patch_size",), dict(type=int, default=64)),
    (("--test",), dict(action="store_true")),
    (("--model_ckpt",), dict(type=str)),
    (("--val_metrics_every",), dict(type=int, default=5)),
    (("-cr", "--compression_ratio"), dict(type=float, default=1.5)),
    (("--model_type",), dict(type=str, default="Cond_SRVAE", choices=["Cond_SRVAE", "VAE"])),
    (("--dtype",), dict(type=str, default="fp32", choices=["fp32", "bf16"])),
)
_DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16}


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Train a VAE model.")
    for names, kw in _FLAGS:
        parser.add_argument(*names, **kw)
    return parser.parse_args(argv)


def _this_rank_device() -> torch.device:
    """One process per GPU: pick LOCAL_RANK's device and join the NCCL group when launched by torchrun."""
    if not torch.cuda.is_available():
        raise RuntimeError("svrs_b200 needs a CUDA device (B200 / sm_100a); there is no CPU path")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=dev)
    return dev


def _build_model(kind: str, cr: float, patch: int, job: str):
    hooks = [callbacks.ModelCheckpoint(job, "ckpt", monitor="Loss/val_loss", mode="min"),
             callbacks.EarlyStopping(patience=25, delta=0.01)]
    if kind == "Cond_SRVAE":
        return models.Cond_SRVAE(cr, patch, callbacks=hooks, slurm_job_id=job)
    if kind == "VAE":
        return models.VAE(cr, patch // 2, callbacks=hooks, slurm_job_id=job)      # the reference halves it (train.py:44)
    raise ValueError(f"Unknown model type: {kind}. Choose 'Cond_SRVAE' or 'VAE'.")


def main(args):
    dev = _this_rank_device()
    if args.compression_ratio <= 0:
        raise ValueError("Compression ratio must be a positive integer.")
    loaders = init_dataloader(args.dataset, args.batch_size, args.patch_size, device=dev)
    job = os.environ.get("SLURM_JOB_ID") or time.strftime("local_%Y%m%d-%H%M%S")
    net = _build_model(args.model_type, args.compression_ratio, args.patch_size, job).to(dev)
    net.set_compute_dtype(_DTYPES[args.dtype])
    if args.model_ckpt:
        net.load_state_dict(torch.load(args.model_ckpt, map_location=dev))
    evaluate_only = bool(args.test and args.model_ckpt)
    if not evaluate_only:
        net.fit(train_loader=loaders[0], val_loader=loaders[1], epochs=args.epochs, device=dev,
                optimizer=torch.optim.Adam(net.parameters(), lr=1e-4), start_epoch=1,
                val_metrics_every=args.val_metrics_every, slurm_job_id=job)
    net.task(loaders[1])


if __name__ == "__main__":
    cli = parse_args()
    print(cli)
    if cli.model_ckpt and not os.path.exists(cli.model_ckpt):
        raise FileNotFoundError(f"Model checkpoint {cli.model_ckpt} not found.")
    main(cli)
