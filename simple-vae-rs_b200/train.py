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


def main(args):
    if not torch.cuda.is_available():
        raise RuntimeError("svrs_b200 needs a CUDA device (B200 / sm_100a); there is no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    train_loader, val_loader = init_dataloader(args.dataset, args.batch_size, args.patch_size, device=device)
    cr = args.compression_ratio
    if cr <= 0:
        raise ValueError("Compression ratio must be a positive integer.")
    slurm_job_id = os.environ.get("SLURM_JOB_ID", f"local_{time.strftime('%Y%m%d-%H%M%S')}")
    callbacks_list = [callbacks.ModelCheckpoint(slurm_job_id, "ckpt", monitor="Loss/val_loss", mode="min"),
                      callbacks.EarlyStopping(patience=25, delta=0.01)]
    if args.model_type == "VAE":
        model = models.VAE(cr, args.patch_size // 2, callbacks=callbacks_list, slurm_job_id=slurm_job_id)
    elif args.model_type == "Cond_SRVAE":
        model = models.Cond_SRVAE(cr, args.patch_size, callbacks=callbacks_list, slurm_job_id=slurm_job_id)
    else:
        raise ValueError(f"Unknown model type: {args.model_type}. Choose 'Cond_SRVAE' or 'VAE'.")
    model.to(device)
    model.set_compute_dtype(torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    if args.model_ckpt:
        model.load_state_dict(torch.load(args.model_ckpt, map_location=device))
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4)
    if not (args.test and args.model_ckpt):
        model.fit(train_loader=train_loader, val_loader=val_loader, epochs=args.epochs, device=device,
                  optimizer=optimizer, start_epoch=1, val_metrics_every=args.val_metrics_every,
                  slurm_job_id=slurm_job_id)
    model.task(val_loader)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Train a VAE model.")
    p.add_argument("--pre_epochs", type=int, default=20, help="(unused, kept for CLI compatibility)")
    p.add_argument("--epochs", type=int, default=200)
    p.add_argument("--dataset", type=str, default="s2v")
    p.add_argument("--batch_size", type=int, default=16)
    p.add_argument("--patch_size", type=int, default=64)
    p.add_argument("--test", action="store_true")
    p.add_argument("--model_ckpt", type=str)
    p.add_argument("--val_metrics_every", type=int, default=5)
    p.add_argument("-cr", "--compression_ratio", type=float, default=1.5)
    p.add_argument("--model_type", type=str, default="Cond_SRVAE", choices=["Cond_SRVAE", "VAE"])
    p.add_argument("--dtype", type=str, default="fp32", choices=["fp32", "bf16"])
    return p.parse_args(argv)


if __name__ == "__main__":
    arguments = parse_args()
    print(arguments)
    if arguments.model_ckpt and not os.path.exists(arguments.model_ckpt):
        raise FileNotFoundError(f"Model checkpoint {arguments.model_ckpt} not found.")
    main(args=arguments)
