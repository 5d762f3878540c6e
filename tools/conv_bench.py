"""Micro-benchmark of one conv layer through the C ABI (bf16): fprop / dgrad / wgrad, halo on/off.
    python tools/conv_bench.py N H W Cin Cout [ksize=3] [reps=50]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from helpers import BF16, lib, pack, st

N, H, W, Cin, Cout = (int(v) for v in sys.argv[1:6])
ks = int(sys.argv[6]) if len(sys.argv) > 6 else 3
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 50
only = os.environ.get("ONLY", "")
dev = "cuda"
torch.manual_seed(0)
s = 1 if ks == 3 else 2
x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
gy = torch.randn(N, H // s, W // s, Cout, device=dev).to(torch.bfloat16)
w = torch.randn(Cout, Cin, ks, ks, device=dev) * 0.1
b = torch.randn(Cout, device=dev)
pf, pb = pack(w, torch.bfloat16)
y = torch.empty_like(gy)
dx = torch.empty_like(x)
dw = torch.zeros_like(w)
dwp = dw.data_ptr() if os.environ.get('PACKED', '1') == '1' else None   # packed scratch (timing only)
flops = 2.0 * N * (H // s) * (W // s) * Cin * Cout * ks * ks


def timeit(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{label:28s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s")


for halo in ((1, 0) if ks == 3 else (1,)):
    lib.set_halo_mode(halo)
    tag = f"halo={halo}"
    if only in ("", "fprop"):
        timeit(lambda: lib.conv2d_fprop(x.data_ptr(), pf.data_ptr(), pb.data_ptr(), b.data_ptr(), y.data_ptr(), BF16, N, H, W, Cin, Cout, ks, 0, st()), f"fprop {tag}")
    if only in ("", "dgrad"):
        timeit(lambda: lib.conv2d_dgrad(gy.data_ptr(), pb.data_ptr(), pf.data_ptr(), dx.data_ptr(), BF16, N, H, W, Cin, Cout, ks, st()), f"dgrad {tag}")
for halo in ((1, 0) if ks == 3 else (1,)):
    lib.set_halo_mode(halo)
    if only in ("", "wgrad"):
        timeit(lambda: lib.conv2d_wgrad(x.data_ptr(), gy.data_ptr(), dw.data_ptr(), dwp, None, BF16, N, H, W, Cin, Cout, ks, 0, st()), f"wgrad halo={halo}")
lib.set_halo_mode(1)
