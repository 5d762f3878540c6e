"""Shared helpers for the GPU parity tests: everything calls the product through the C ABI (ctypes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simple-vae-rs_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from svrs_native.lib import BF16, F32, lib  # noqa: E402


def st():
    return torch.cuda.current_stream().cuda_stream


def dt(t):
    return F32 if t == torch.float32 else BF16


def nhwc(t, dtype=torch.float32):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous().float()


def pack(w, dtype, convT=False):
    """returns (pack_f, pack_b) for a conv / convT weight in torch layout (fp32, cuda)."""
    d0, d1 = w.shape[0], w.shape[1]
    kk = w.shape[2] * w.shape[3]
    p01 = torch.empty(w.numel(), device=w.device, dtype=dtype)
    p10 = torch.empty(w.numel(), device=w.device, dtype=dtype)
    lib.pack_weights(w.data_ptr(), d0, d1, kk, p01.data_ptr(), p10.data_ptr(), dt(dtype), st())
    return (p01, p10) if convT else (p10, p01)


def report(name, got, ref, rtol, atol=0.0):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    mx = err.max().item() if err.numel() else 0.0
    rel = mx / scale
    print(f"[parity] {name}: max|err|={mx:.3e} rel-to-max={rel:.3e} (ref max {scale:.3e})")
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    assert mx <= atol + rtol * scale, f"{name}: max|err| {mx:.3e} > {atol} + {rtol}*{scale:.3e}"
    return rel
