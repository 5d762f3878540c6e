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


# Algorithmic HBM bytes of the bandwidth-bound entry points (every operand read once, every result written once; DESIGN
# section 3).  Returns 0 for calls that are not on this list.
def _hbm_bytes(name, a):
    d = dict(zip(_ARGS[name], a))
    es = lambda dt: 4 if dt == 0 else 2
    if name == "svrs_elbo_fwd":
        lat = 4 * d["B"] * (2 * d["W1"] + 4 * d["W2"])
        return d["n_x"] * (es(d["dt_x"]) + es(d["dt_tx"])) + d["n_y"] * (es(d["dt_y"]) + es(d["dt_ty"])) + lat
    if name == "svrs_elbo_bwd":
        lat = 4 * d["B"] * (2 * d["W1"] + 4 * d["W2"])
        return (d["n_x"] * (es(d["dt_x"]) + es(d["dt_tx"]) + es(d["dt_dx"])) + d["n_y"] * (es(d["dt_y"]) + es(d["dt_ty"]) + es(d["dt_dy"]))
                + 2 * lat)
    if name in ("svrs_reparam_fwd", "svrs_reparam_bwd"):
        return 4 * d["B"] * d["Wd"] * (3 if name.endswith("fwd") else 6)
    if name == "svrs_bn_stats":
        return d["M"] * d["C"] * es(d["dtype"])
    if name == "svrs_bn_apply_train" or name == "svrs_bn_apply":
        return 2 * d["M"] * d["C"] * es(d["dtype"])
    if name == "svrs_bn_bwd_reduce":
        return 2 * d["M"] * d["C"] * es(d["dtype"])
    if name == "svrs_bn_bwd_apply":
        return 3 * d["M"] * d["C"] * es(d["dtype"])
    if name == "svrs_sumsq":
        return 4 * d["n"]
    if name == "svrs_clip_adam":
        return 28 * d["n"]
    if name == "svrs_patch_gather_normalize":
        per = d["C"] * d["P"] * d["P"]
        outs = (4 if d["out_nchw_f32"] else 0) + (4 if d["out_nhwc_f32"] else 0) + (2 if d["out_nhwc_bf16"] else 0)
        return d["npatch"] * per * ((2 if d["src_is_i16"] else 4) + outs)
    return 0


def dominant_kernel_roofline(tr, inputs, pk, steps: int = 2, dtype=None, eager=None, rt=None):
    """-> (roofline of the dominant tensor-core kernel family, roofline_hbm of the bandwidth-bound kernels), both timed per
    launch with CUDA events in an eager single-stream pass at the bench batch.  tr = fused trainer (training workloads) or
    None with `eager` (a callable running one eager step) and `rt` (the engine's Runtime) for the inference workload."""
    if eager is None:
        def eager():
            lr, hr = inputs
            P = tr.eng.P
            if hasattr(tr.eng, "Wz"):
                tr.step_tiles(hr, lr, patch_size=P, use_graph=False)
            else:
                tr.step_tiles(hr, patch_size=P, use_graph=False)

    class _Holder:
        pass

    if tr is None:
        tr = _Holder()
        tr.rt = rt
    side, tr.rt.wgrad_side = tr.rt.wgrad_side, False      # per-call event brackets only see the current stream,
    br, tr.rt.branch_streams = tr.rt.branch_streams, False  # and kernels must not overlap while they are being timed
    pad, lib.timing_pad_cycles = lib.timing_pad_cycles, 100000
    hbm = defaultdict(lambda: [0.0, 0.0, 0.0])
    try:
        eager()
        torch.cuda.synchronize()
        lib.timing = []
        try:
            for _ in range(steps):
                eager()
            torch.cuda.synchronize()
            rec = lib.timing
        finally:
            lib.timing = None
    finally:
        tr.rt.wgrad_side, tr.rt.branch_streams = side, br
        lib.timing_pad_cycles = pad
    per = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for name, a, e0, e1, kernels in rec:
        ms = e0.elapsed_time(e1)
        r = per[(name, kernels)]
        r[0] += ms / steps
        r[1] += 1.0 / steps
        r[2] += _flops(name, a) / steps
        r[3] += _bytes(name, a) / steps
        hb = _hbm_bytes(name, a)
        if hb:
            h = hbm[name.replace("svrs_", "")]
            h[0] += ms / steps
            h[1] += 1.0 / steps
            h[2] += hb / steps
    # adam_multi: bytes from the flat buffer size (16 B read + 12 B written per parameter + 4 B of packs)
    n_par = tr.rt.store.total
    for (name, kernels), v in per.items():
        if name == "svrs_adam_multi":
            hbm["adam_multi"] = [v[0], v[1], 32.0 * n_par * v[1]]
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
    shares = {k: round(v[0] / total_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:10]}
    by_kernel = {k: {"ms_per_step": round(v[0], 4), "launches": round(v[1], 1), "tflops": round(v[2] / (v[0] * 1e-3) / 1e12, 1) if v[0] > 0 else 0.0}
                 for k, v in sorted(conv.items(), key=lambda kv: -kv[1][0])}
    roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": pk["tf_sus"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_sus"], "traffic": None, "peak_source": pk["src"] + " bf16 sustained",
            "launches_per_step": n, "ms_per_step_in_kernel": ms, "us_per_launch": 1e3 * ms / n if n else 0.0,
            "algorithmic_flops_per_launch": fl / n if n else 0.0, "algorithmic_bytes_per_launch": by / n if n else 0.0,
            "ms_per_step_all_kernels_eager": total_ms, "launches_per_step_all_kernels": sum(v[1] for v in per.values()),
            "all_conv_kernels": {"ms_per_step": conv_ms, "achieved": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0,
                                 "unit": "TFLOP/s"},
            "conv_kernels": by_kernel,
            "share_of_step_by_kernel": shares}
    roof_hbm = {"bound": "hbm", "peak": pk["hbm"], "unit": "GB/s", "peak_source": pk["src"] + " copy bandwidth",
                "note": "CUDA-event time per launch at the bench batch (eager, single stream); achieved = algorithmic bytes / time",
                "kernels": {k: {"us_per_launch": round(1e3 * v[0] / v[1], 2), "launches": round(v[1], 1),
                                "MB_per_launch": round(v[2] / v[1] / 1e6, 3),
                                "achieved": round(v[2] / (v[0] * 1e-3) / 1e9, 1),
                                "frac": round(v[2] / (v[0] * 1e-3) / 1e9 / pk["hbm"], 3)}
                            for k, v in sorted(hbm.items(), key=lambda kv: -kv[1][0]) if v[0] > 0 and v[1] > 0}}
    return roof, roof_hbm
