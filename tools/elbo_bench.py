#!/usr/bin/env python
"""Bandwidth of the fused ELBO forward kernel (svrs_elbo_fwd) at CondVAE cr=2 P=64 for several batch sizes:
algorithmic bytes = 77,824 fp32 elements read per patch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-vae-rs_b200"))
import torch
from svrs_native.lib import lib
from svrs_native.elbo import elbo_forward

dev = "cuda"
gam = torch.ones(2, device=dev)
for B in (128, 1024, 4096):
    Wu, Wz = 2048, 8192
    xh, x = torch.rand(B, 4, 64, 64, device=dev), torch.rand(B, 4, 64, 64, device=dev)
    yh, y = torch.rand(B, 4, 32, 32, device=dev), torch.rand(B, 4, 32, 32, device=dev)
    eu, ez = torch.randn(B, 2 * Wu, device=dev) * 0.1, torch.randn(B, 2 * Wz, device=dev) * 0.1
    m3, l3 = torch.randn(B, Wz, device=dev) * 0.1, torch.randn(B, Wz, device=dev) * 0.1
    args = (xh, x, yh, y, eu[:, :Wu], eu[:, Wu:], ez[:, :Wz], ez[:, Wz:], m3, l3, gam, B)
    for _ in range(3):
        elbo_forward(*args)
    torch.cuda.synchronize()
    lib.timing, lib.timing_pad_cycles = [], 200000
    for _ in range(10):
        elbo_forward(*args)
    torch.cuda.synchronize()
    rec, lib.timing = lib.timing, None
    us = sorted(e0.elapsed_time(e1) * 1e3 for n, a, e0, e1, k in rec if n == "svrs_elbo_fwd")
    t = us[len(us) // 2]
    mb = 77824 * 4 * B / 1e6
    print(f"B={B:5d}  elbo_fwd median {t:8.1f} us  {mb:8.1f} MB  {mb / t * 1e3 / 1e3:5.2f} TB/s  ({mb / t / 6.5498 * 100:.0f}% of 6549.8 GB/s)")
