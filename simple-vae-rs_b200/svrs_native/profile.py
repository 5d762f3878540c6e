"""Per-kernel device timing of the training step (CUDA events around every C-ABI launch, eager mode) and the
roofline of the dominant kernel family for bench.py.  Not used on the timed path."""
from __future__ import annotations

from collections import defaultdict

import torch

from .lib import lib

_ARGS = {n: [a for _, a in args] for n, (_, args) in lib.protos.items()}


def _flops(name, a):
    """Algorithmic FLOPs of one conv launch from its C-ABI arguments (2*M*K*N*taps)."""
    d = dict(zip(_ARGS[name], a))
    if name in ("svrs_conv2d_fprop", "svrs_conv2d_dgrad", "svrs_conv2d_wgrad"):
        k = d["ksize"]
        s = 1 if k == 3 else 2
        return 2.0 * d["N"] * (d["H"] // s) * (d["W"] // s) * d["Cin"] * d["Cout"] * k * k
    if name in ("svrs_convT2d_fprop", "svrs_convT2d_dgrad", "svrs_convT2d_wgrad"):
        return 2.0 * d["N"] * d["H"] * d["W"] * d["Cin"] * d["Cout"] * 16
    return 0.0


def _bytes(name, a):
    """Algorithmic bytes of one conv launch: every operand read once, the result written once (bf16 activations and
    packed weights, fp32 weight gradient)."""
    d = dict(zip(_ARGS[name], a))
    if "Cin" not in d:
        return 0.0
    if name.startswith("svrs_convT2d"):
        kk, n_in, n_out = 16, d["N"] * d["H"] * d["W"] * d["Cin"], d["N"] * 4 * d["H"] * d["W"] * d["Cout"]
    else:
        k = d["ksize"]
        s = 1 if k == 3 else 2
        kk, n_in, n_out = k * k, d["N"] * d["H"] * d["W"] * d["Cin"], d["N"] * (d["H"] // s) * (d["W"] // s) * d["Cout"]
    w = d["Cin"] * d["Cout"] * kk
    if name.endswith("wgrad"):
        return 2.0 * (n_in + n_out) + 4.0 * w
    return 2.0 * (n_in + n_out + w)


def time_step(run_eager_step, steps: int = 2):
    """-> {(abi function, kernels it dispatched to): (total ms, calls, total flops)} averaged per step.  The kernel names
    come from the library's own launch trace (svrs_trace), so the attribution follows the real dispatch."""
    run_eager_step()                      # warm (allocator)
    torch.cuda.synchronize()
    lib.timing = []
    try:
        for _ in range(steps):
            run_eager_step()
        torch.cuda.synchronize()
        rec = lib.timing
    finally:
        lib.timing = None
    agg = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for name, a, e0, e1, kernels in rec:
        r = agg[(name, kernels)]
        r[0] += e0.elapsed_time(e1) / steps
        r[1] += 1.0 / steps
        r[2] += _flops(name, a) / steps
        r[3] += _bytes(name, a) / steps
    return {k: tuple(v) for k, v in agg.items()}


def dominant_kernel_roofline(tr, step_from_device, inputs, pk, steps: int = 2):
    import dataset

    def eager():
        lr, hr = inputs
        P = tr.eng.P
        if hasattr(tr.eng, "Wz"):
            y = dataset.grid_patch_normalize(lr, P // 2)
            x = dataset.grid_patch_normalize(hr, P)
            tr.step(x, y, use_graph=False)
        else:
            tr.step(dataset.grid_patch_normalize(hr, P), use_graph=False)

    side, tr.rt.wgrad_side = tr.rt.wgrad_side, False      # per-call event brackets only see the current stream,
    br, tr.rt.branch_streams = tr.rt.branch_streams, False  # and kernels must not overlap while they are being timed
    try:
        per = time_step(eager, steps)
    finally:
        tr.rt.wgrad_side, tr.rt.branch_streams = side, br
    fam = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    total_ms = sum(v[0] for v in per.values())
    for (name, kernels), (ms, n, fl, by) in per.items():
        f = fam[kernels or name]          # a wgrad call with a bias is "<wgrad kernel>,<colsum kernel>" in one bracket
        f[0] += ms; f[1] += n; f[2] += fl; f[3] += by
    conv = {k: v for k, v in fam.items() if v[2] > 0}
    top = max(conv.items(), key=lambda kv: kv[1][0])
    name, (ms, n, fl, by) = top
    achieved = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    conv_ms = sum(v[0] for v in conv.values())
    conv_fl = sum(v[2] for v in conv.values())
    shares = {k: round(v[0] / total_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:8]}
    return {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": pk["tf_sus"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_sus"], "traffic": None, "peak_source": pk["src"] + " bf16 sustained",
            "launches_per_step": n, "ms_per_step_in_kernel": ms, "us_per_launch": 1e3 * ms / n if n else 0.0,
            "algorithmic_flops_per_launch": fl / n if n else 0.0, "algorithmic_bytes_per_launch": by / n if n else 0.0, "ms_per_step_all_kernels_eager": total_ms,
            "all_conv_kernels": {"ms_per_step": conv_ms, "achieved": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0,
                                 "unit": "TFLOP/s"},
            "share_of_step_by_kernel": shares}
