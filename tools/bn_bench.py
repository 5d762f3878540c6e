#!/usr/bin/env python
"""Micro-benchmark of the BatchNorm kernels at the bench workload's layer shapes (bf16): GB/s per kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-vae-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import lib, BF16, st

def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # L2 flush between shapes
for M, C in ((524288, 64), (131072, 128), (32768, 256), (131072, 64), (8192, 512)):
    x = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x)
    sums = torch.zeros(2 * C * 8, device="cuda", dtype=torch.float64)
    sc, sh, mu, iv, gam = (torch.rand(C, device="cuda") + 0.5 for _ in range(5))
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    mb = M * C * 2 / 1e6
    t = timeit(lambda: lib.bn_stats(x.data_ptr(), BF16, M, C, sums.data_ptr(), st()))
    print(f"M={M:7d} C={C:4d}  bn_stats      {t:7.1f} us  {mb / t * 1e3 / 1e3:6.2f} TB/s")
    t = timeit(lambda: lib.bn_apply(x.data_ptr(), y.data_ptr(), BF16, M, 0, C, sc.data_ptr(), sh.data_ptr(), 1, st()))
    print(f"M={M:7d} C={C:4d}  bn_apply      {t:7.1f} us  {2 * mb / t * 1e3 / 1e3:6.2f} TB/s")
    t = timeit(lambda: lib.bn_bwd_reduce(x.data_ptr(), dy.data_ptr(), BF16, M, 0, C, sc.data_ptr(), sh.data_ptr(), mu.data_ptr(), iv.data_ptr(), 1, sums.data_ptr(), st()))
    print(f"M={M:7d} C={C:4d}  bn_bwd_reduce {t:7.1f} us  {2 * mb / t * 1e3 / 1e3:6.2f} TB/s")
    t = timeit(lambda: lib.bn_bwd_apply(x.data_ptr(), dy.data_ptr(), y.data_ptr(), BF16, M, 0, C, sc.data_ptr(), sh.data_ptr(), mu.data_ptr(), iv.data_ptr(), gam.data_ptr(), 1, sums.data_ptr(), dg.data_ptr(), db.data_ptr(), st()))
    print(f"M={M:7d} C={C:4d}  bn_bwd_apply  {t:7.1f} us  {3 * mb / t * 1e3 / 1e3:6.2f} TB/s")
