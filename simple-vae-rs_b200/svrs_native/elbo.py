"""Fused Gaussian-ELBO callables (loss/vae_loss.py:5-13, loss/cond_vae_loss.py:5-58) as autograd Functions
over svrs_elbo_fwd / svrs_elbo_finalize / svrs_elbo_bwd."""
from __future__ import annotations

from typing import Optional

import torch

from .lib import BF16, F32, SvrsError, lib


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise SvrsError(f"unsupported dtype {t.dtype}")


def _rows(t: torch.Tensor) -> torch.Tensor:
    """Latent tensors [B, W] may be torch.chunk views (row stride 2W); the kernels take a row stride."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        t = t.reshape(t.shape[0], -1)
    if t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def _img(t: torch.Tensor, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    if like is not None and t.dtype != like.dtype:
        t = t.to(like.dtype)
    return t.contiguous()


def _gammas_dev(gx, gy, device) -> torch.Tensor:
    """gammas are plain 0-dim tensors that stay on the CPU in the reference."""
    vals = []
    for g in (gx, gy):
        if g is None:
            vals.append(torch.ones((), device=device))
        elif g.device != device:
            vals.append(g.detach().to(device=device, dtype=torch.float32))
        else:
            vals.append(g.detach().float())
    return torch.stack(vals)


def elbo_forward(recon_x, x, recon_y, y, mu1, lv1, mu2, lv2, mu3, lv3, gammas: torch.Tensor, B: int):
    """Raw launcher: returns (terms5, acc).  Any group may be None."""
    dev = gammas.device
    acc = torch.zeros(4, device=dev, dtype=torch.float64)
    terms = torch.empty(5, device=dev, dtype=torch.float32)
    nx = recon_x.numel() if recon_x is not None else 0
    ny = recon_y.numel() if recon_y is not None else 0
    lib.elbo_fwd(_p(recon_x), _p(x), _dt(recon_x) if nx else F32, _dt(x) if nx else F32, nx,
                 _p(recon_y), _p(y), _dt(recon_y) if ny else F32, _dt(y) if ny else F32, ny,
                 _p(mu1), _p(lv1), mu1.stride(0) if mu1 is not None else 0, mu1.shape[1] if mu1 is not None else 0,
                 _p(mu2), _p(lv2), mu2.stride(0) if mu2 is not None else 0,
                 _p(mu3), _p(lv3), mu3.stride(0) if mu3 is not None else 0, mu2.shape[1] if mu2 is not None else 0,
                 B, _p(acc), _st())
    lib.elbo_finalize(_p(acc), nx, ny, B, _p(gammas), _p(terms), _st())
    return terms, acc


class _CondLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon_x, x, recon_y, y, mu1, lv1, mu2, lv2, mu3, lv3, gammax, gammay):
        for t in (recon_x, x, recon_y, y, mu1, lv1, mu2, lv2, mu3, lv3):
            if not t.is_cuda:
                raise SvrsError("cond_loss: tensors must live on a CUDA device (no CPU fallback in svrs_b200)")
        dev = recon_x.device
        B = recon_x.shape[0]
        rx, ry = _img(recon_x), _img(recon_y)
        xx, yy = _img(x, rx), _img(y, ry)
        m1, l1, m2, l2, m3, l3 = (_rows(t) for t in (mu1, lv1, mu2, lv2, mu3, lv3))
        if l1.stride(0) != m1.stride(0):
            l1 = l1.contiguous(); m1 = m1.contiguous()
        if l2.stride(0) != m2.stride(0):
            l2 = l2.contiguous(); m2 = m2.contiguous()
        if l3.stride(0) != m3.stride(0):
            l3 = l3.contiguous(); m3 = m3.contiguous()
        gam = _gammas_dev(gammax, gammay, dev)
        terms, acc = elbo_forward(rx, xx, ry, yy, m1, l1, m2, l2, m3, l3, gam, B)
        ctx.save_for_backward(rx, xx, ry, yy, m1, l1, m2, l2, m3, l3, gam, acc)
        ctx.gdev = (gammax.device, gammay.device)
        ctx.B = B
        return terms[0], terms[1], terms[2], terms[3]

    @staticmethod
    def backward(ctx, g_msex, g_klu, g_msey, g_klz):
        rx, xx, ry, yy, m1, l1, m2, l2, m3, l3, gam, acc = ctx.saved_tensors
        dev = rx.device
        zero = torch.zeros((), device=dev)
        gout = torch.stack([(g if g is not None else zero).float().reshape(()) for g in (g_msex, g_klu, g_msey, g_klz)])
        d_rx, d_ry = torch.empty_like(rx), torch.empty_like(ry)
        d = [torch.empty(t.shape, device=dev, dtype=torch.float32) for t in (m1, l1, m2, l2, m3, l3)]
        dgam = torch.empty(2, device=dev, dtype=torch.float32)
        lib.elbo_bwd(_p(rx), _p(xx), _dt(rx), _dt(xx), rx.numel(), _p(d_rx), _dt(d_rx),
                     _p(ry), _p(yy), _dt(ry), _dt(yy), ry.numel(), _p(d_ry), _dt(d_ry),
                     _p(m1), _p(l1), m1.stride(0), m1.shape[1], _p(d[0]), _p(d[1]), d[0].stride(0),
                     _p(m2), _p(l2), m2.stride(0), _p(d[2]), _p(d[3]), d[2].stride(0),
                     _p(m3), _p(l3), m3.stride(0), m2.shape[1], _p(d[4]), _p(d[5]), d[4].stride(0),
                     ctx.B, _p(acc), _p(gam), _p(gout), _p(dgam), 0, _st())
        dgx = dgam[0].to(ctx.gdev[0])
        dgy = dgam[1].to(ctx.gdev[1])
        return d_rx, None, d_ry, None, d[0], d[1], d[2], d[3], d[4], d[5], dgx, dgy


class _BaseLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon_x, x, mu, logvar, gamma):
        for t in (recon_x, x, mu, logvar):
            if not t.is_cuda:
                raise SvrsError("base_loss: tensors must live on a CUDA device (no CPU fallback in svrs_b200)")
        dev = recon_x.device
        B = recon_x.shape[0]
        rx = _img(recon_x)
        xx = _img(x, rx)
        m1, l1 = _rows(mu), _rows(logvar)
        if l1.stride(0) != m1.stride(0):
            l1 = l1.contiguous(); m1 = m1.contiguous()
        gam = _gammas_dev(gamma, None, dev)
        terms, acc = elbo_forward(rx, xx, None, None, m1, l1, None, None, None, None, gam, B)
        ctx.save_for_backward(rx, xx, m1, l1, gam, acc)
        ctx.gdev = gamma.device
        ctx.B = B
        return terms[0], terms[1]

    @staticmethod
    def backward(ctx, g_mse, g_kld):
        rx, xx, m1, l1, gam, acc = ctx.saved_tensors
        dev = rx.device
        zero = torch.zeros((), device=dev)
        gout = torch.stack([(g if g is not None else zero).float().reshape(()) for g in (g_mse, g_kld, None, None)])
        d_rx = torch.empty_like(rx)
        dm, dl = (torch.empty(t.shape, device=dev, dtype=torch.float32) for t in (m1, l1))
        dgam = torch.empty(2, device=dev, dtype=torch.float32)
        lib.elbo_bwd(_p(rx), _p(xx), _dt(rx), _dt(xx), rx.numel(), _p(d_rx), _dt(d_rx),
                     None, None, F32, F32, 0, None, F32,
                     _p(m1), _p(l1), m1.stride(0), m1.shape[1], _p(dm), _p(dl), dm.stride(0),
                     None, None, 0, None, None, 0,
                     None, None, 0, 0, None, None, 0,
                     ctx.B, _p(acc), _p(gam), _p(gout), _p(dgam), 0, _st())
        return d_rx, None, dm, dl, dgam[0].to(ctx.gdev)


def cond_loss(recon_x, x, recon_y, y, mu1, logvar1, mu2, logvar2, mu3, logvar3, gammax, gammay):
    """loss/cond_vae_loss.py:5 - same positional order, returns (mse_x, kld_u, mse_y, kld_z)."""
    return _CondLossFn.apply(recon_x, x, recon_y, y, mu1, logvar1, mu2, logvar2, mu3, logvar3, gammax, gammay)


def base_loss(recon_x, x, mu, logvar, gamma):
    """loss/vae_loss.py:5 - returns (mse, kld)."""
    return _BaseLossFn.apply(recon_x, x, mu, logvar, gamma)
