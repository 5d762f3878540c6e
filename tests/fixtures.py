"""Rebuild fixture weights / inputs / eps exactly as oracle/make_golden.py produced them (portable numpy PCG64)."""
import os

import torch

from oracle import ref_oracle as O


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"))


def build(fx_or_kind, cr=None, P=None, seed=0, device=None, dtype=torch.float32):
    """-> (model, sd) with the fixture's portable weights loaded; sd is a CPU fp32 copy for the oracle."""
    import models
    if isinstance(fx_or_kind, dict):
        kind, cr, P, seed = fx_or_kind["kind"], fx_or_kind["cr"], fx_or_kind["P"], fx_or_kind["seed_model"]
    else:
        kind = fx_or_kind
    m = models.Cond_SRVAE(cr, P) if kind == "cond" else models.VAE(cr, P)
    sd = O.portable_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    if device is not None:
        m.to(device)
        m.set_compute_dtype(dtype)
    return m, {k: v.clone() for k, v in sd.items()}


def checksum_ok(sd, ref):
    """(sum, abs-sum) digests; tolerant to summation order (CPU thread count / ISA differ between hosts)."""
    return all(torch.allclose(torch.stack([sd[k].double().sum(), sd[k].double().abs().sum()]), v, rtol=1e-9, atol=1e-9)
               for k, v in ref.items())


def grads_fp64(fx, sd, x, y, eps):
    """Parameter gradients of the first step evaluated by the oracle in float64 ('truth' for error budgets)."""
    d = lambda t: None if t is None else t.double()
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
    if fx["kind"] == "cond":
        gam = {"gammax": torch.tensor(1.0, dtype=torch.float64), "gammay": torch.tensor(1.0, dtype=torch.float64)}
        _, _, g = O.cond_train_step(sd64, gam, O.AdamState(), fx["cr"], fx["P"], d(x), d(y), d(eps[0]), d(eps[1]), return_grads=True)
    else:
        gam = {"gamma": torch.tensor(1.0, dtype=torch.float64)}
        _, _, g = O.vae_train_step(sd64, gam, O.AdamState(), fx["cr"], fx["P"], d(x), d(eps[0]), return_grads=True)
    return g


def check_grads(name, named_params, grads32, grads64, skip_suffixes, report):
    """CUDA gradients vs float64 truth, budgeted against the error the reference's own fp32 CPU arithmetic makes on the
    same quantity: err_cuda <= max(2e-4, 4 * err_reference_fp32), both relative to the tensor's max magnitude."""
    bad, worst = [], 0.0
    for k, p in named_params:
        if k.endswith(skip_suffixes):
            continue
        t64 = grads64[k]
        scale = float(t64.abs().max()) + 1e-30
        e_ref = float((grads32[k].double() - t64).abs().max()) / scale
        e_gpu = float((p.grad.detach().double().cpu() - t64).abs().max()) / scale
        worst = max(worst, e_gpu)
        flag = e_gpu > max(2e-4, 4 * e_ref)
        print(f"[parity] {name} grad {k}: cuda-vs-fp64 {e_gpu:.2e}  reference(fp32 CPU)-vs-fp64 {e_ref:.2e}{'  <-- FAIL' if flag else ''}")
        if flag:
            bad.append(k)
    print(f"[parity] {name}: worst CUDA parameter-gradient error vs fp64 truth (rel to max) = {worst:.3e}")
    assert not bad, f"gradient parity failed for {bad}"



def inputs(fx):
    """(x, y) of the fixture (stored, or regenerated from seed_data)."""
    if fx.get("x") is not None:
        return fx["x"], fx.get("y")
    r = O.PortableRng(fx["seed_data"])
    x = r.rand(fx["B"], 4, fx["P"], fx["P"])
    y = r.rand(fx["B"], 4, fx["P"] // 2, fx["P"] // 2) if fx["kind"] == "cond" else None
    return x, y


def eps_stream(fx, widths):
    """Generator of per-step eps lists in the reference's draw order (u then z for Cond_SRVAE)."""
    r = O.PortableRng(fx["seed_step"])
    while True:
        yield [r.randn(fx["B"], w) for w in widths]
