#!/usr/bin/env python
"""bench.py - headline benchmark: train patches/sec of the 64x64 SR CondVAE step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cond_grid|vae]

Workload at every N (weak scaling): BASELINE.json config 3 - CondVAE cr=2, P=64, grid mode: 8 synthetic 256x256
multispectral tiles per GPU -> 128 patches per GPU (64 tiles / 1024 patches at N=8), bf16 compute, fp32 master
weights, gradients SUM-all-reduced over NCCL.  One "step" = grid-patch gather + normalise, forward, ELBO, backward,
clip, Adam (the reference's models/base.py:103-107 body preceded by its dataset.py grid mode).

Prints ONE JSON line (rank 0).  `value` = device-timed whole-job patches/s with tiles resident in HBM;
`e2e` = the same through the public API with HOST (pinned) tile buffers: H2D copy of every step's tiles and a D2H read
of the loss inside the timed region.  `--impl reference` times the CPU restatement of the reference (oracle/) on the
host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "simple-vae-rs_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

TILES_PER_GPU = 8
PATCH = 64
CR = 2
# algorithmic FLOPs per patch of one CondVAE(cr=2,P=64) optimisation step (SURVEY 8.4 row d / BASELINE.md section 4)
FLOP_PER_PATCH_TRAIN = 8.13e9
FLOP_PER_PATCH_VAE = 4.44e9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(args, workload: str):
    """The reference's CPU implementation of the step (oracle restatement: same ATen CPU kernels, fp32), all host threads."""
    from oracle import ref_oracle as O
    import models
    torch.set_num_threads(os.cpu_count())
    B = 8                                    # bounded sample: one config-1 sized batch per step (SURVEY 8.4 row d)
    torch.manual_seed(0)
    if workload == "vae":
        m = models.VAE(CR, PATCH)
    else:
        m = models.Cond_SRVAE(CR, PATCH)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 4, PATCH, PATCH, generator=g)
    y = torch.rand(B, 4, PATCH // 2, PATCH // 2, generator=g)
    opt = O.AdamState()
    if workload == "vae":
        gam = {"gamma": torch.tensor(1.0)}
        Wd = (m.latent_size // 64) * (PATCH // 4) ** 2
        fn = lambda: O.vae_train_step(sd, gam, opt, CR, PATCH, x, torch.randn(B, Wd))
    else:
        gam = {"gammax": torch.tensor(1.0), "gammay": torch.tensor(1.0)}
        L, Lu = O.cond_latent_sizes(CR, PATCH)
        fn = lambda: O.cond_train_step(sd, gam, opt, CR, PATCH, x, y, torch.randn(B, Lu), torch.randn(B, L))
    for _ in range(max(1, min(args.warmup, 2))):
        fn()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=B / dt, ms=dt * 1e3, cores=os.cpu_count(), steps=steps,
                sample=f"{steps} steps x batch {B} patches, fp32, torch CPU ops, {os.cpu_count()} threads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="cond_grid", choices=["cond_grid", "vae"])
    ap.add_argument("--dtype", type=str, default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl_name = ("CondVAE cr=2 P=64 grid mode: 8 tiles (256x256x4) -> 128 patches per GPU" if args.workload == "cond_grid"
               else "VAE cr=2 P=64, 256 patches per GPU")
    metric = "train patches/sec (64x64 SR CondVAE, device-timed)"

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(args, args.workload)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": r["value"], "unit": "patches/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": wl_name, "note": "CPU arm runs a bounded sample: batch 8 patches per step"},
            "cpu_baseline": {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import models
    from dataset import grid_patch_normalize, synthetic_tiles
    from svrs_native.trainer import FusedCondTrainer, FusedVaeTrainer

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    if args.workload == "cond_grid":
        model = models.Cond_SRVAE(CR, PATCH).to(dev)
        per_gpu = TILES_PER_GPU * (256 // PATCH) ** 2
        flop_per_patch = FLOP_PER_PATCH_TRAIN
    else:
        model = models.VAE(CR, PATCH).to(dev)
        per_gpu = 256
        flop_per_patch = FLOP_PER_PATCH_VAE
    model.set_compute_dtype(dtype)
    model.train()
    Trainer = FusedCondTrainer if args.workload == "cond_grid" else FusedVaeTrainer
    tr = Trainer(model)
    tr.eng.rng.seed = 2026
    tr.eng.rng.sample_offset = rank * per_gpu           # partition-invariant eps (SURVEY 8.4 row e iii)
    use_graph = not args.no_graph

    # synthetic tiles: a few distinct tile sets rotated across steps; pinned host copies for the e2e leg
    n_sets = 4
    host_sets = []
    for s in range(n_sets):
        lr, hr = synthetic_tiles(TILES_PER_GPU if args.workload == "cond_grid" else per_gpu // 16, 256, seed=100 + 17 * rank + s)
        host_sets.append((lr.pin_memory(), hr.pin_memory()))
    dev_sets = [(lr.to(dev), hr.to(dev)) for lr, hr in host_sets]
    lr_buf, hr_buf = torch.empty_like(dev_sets[0][0]), torch.empty_like(dev_sets[0][1])
    h2d_bytes = lr_buf.numel() * 4 + hr_buf.numel() * 4

    def step_from_device(lr, hr):
        if args.workload == "cond_grid":
            y = grid_patch_normalize(lr, PATCH // 2)
            x = grid_patch_normalize(hr, PATCH)
            tr.rt.launches += 2
            return tr.step(x, y, use_graph=use_graph)
        x = grid_patch_normalize(hr, PATCH)
        tr.rt.launches += 1
        return tr.step(x, use_graph=use_graph)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up
    for i in range(args.warmup):
        step_from_device(*dev_sets[i % n_sets])
    barrier()

    # ---------------- timed region 1: device-resident inputs (value)
    l0 = tr.rt.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            out = step_from_device(*dev_sets[i % n_sets])
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = tr.rt.launches - l0
    last_loss = float(out[4])

    # ---------------- timed region 2: host buffers through the public API (e2e)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss_host = torch.empty(5, dtype=torch.float32).pin_memory()
    barrier()
    e2.record()
    # the package's own feed (dataset.TilePrefetcher): pinned host tiles -> device on a copy stream, double-buffered, so the
    # H2D copy of step i+1 overlaps step i; every step's copy and the D2H of its loss terms are inside the timed region
    from dataset import TilePrefetcher
    feed = TilePrefetcher((host_sets[i % n_sets] for i in range(args.steps)), dev)
    for lr_d, hr_d in feed:
        out = step_from_device(lr_d, hr_d)
        loss_host.copy_(out, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    total_patches = per_gpu * world * args.steps
    value = total_patches / (ms * 1e-3)
    e2e_value = total_patches / (ms_e2e * 1e-3)

    # ---------------- roofline of the dominant kernel family (eager, per-launch CUDA events; not in the timed region)
    pk = peaks()
    roof = None
    try:
        from svrs_native import profile as prof
        roof = prof.dominant_kernel_roofline(tr, step_from_device, dev_sets[0], pk, steps=2)
        # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (tools/ncu_traffic.py)
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r01.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("kernel") and tj["kernel"] in roof.get("kernel", ""):
                roof["traffic"] = tj["dram_bytes_per_launch"]
                roof["traffic_source"] = "profiles/traffic_r01.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    except Exception as ex:  # keep the headline number even if the profiling pass fails
        roof = {"error": repr(ex)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            class A: steps, warmup = 3, 1
            r = cpu_reference_arm(A, args.workload)
            cpu = {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        except Exception as ex:
            cpu = {"error": repr(ex)}

    if rank == 0:
        act_mb = 6.24 * 3 * per_gpu
        line = {
            "metric": metric, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": wl_name, "global_patches": per_gpu * world, "parallelism": f"dp{world}",
                       "cuda_graph": use_graph, "l2": f"working set per step ~{act_mb:.0f} MB of activations + 82 MB weights "
                       f"+ 247 MB Adam state per GPU, larger than the 126 MB L2; {n_sets} tile sets rotated",
                       "tensor_roofline_patches_per_s_per_gpu": pk["tf_sus"] * 1e12 / flop_per_patch,
                       "frac_of_step_roofline": (value / world) / (pk["tf_sus"] * 1e12 / flop_per_patch),
                       "last_loss": last_loss},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 20,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: communicators captured inside CUDA graphs make destroy_process_group() hang
        # (observed at N=2).  Everything is synchronised and flushed, so a hard exit is safe.
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
