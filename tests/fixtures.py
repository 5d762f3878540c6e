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
    return all(torch.equal(torch.stack([sd[k].double().sum(), sd[k].double().abs().sum()]), v) for k, v in ref.items())


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
