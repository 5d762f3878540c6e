"""torch.autograd bridges so that the reference's unmodified call pattern
    outs = model(x, y); terms = cond_loss(...); loss.backward(); clip_grad_norm_; optimizer.step()
(models/base.py:103-107) works on top of the hand-written forward/backward chains.

Each model forward is ONE autograd node.  Its backward runs the explicit kernel chain, which accumulates
parameter gradients into the flat gradient buffer, and then publishes them as `param.grad` views."""
from __future__ import annotations

import torch

from .engine import CondEngine, VaeEngine


def _publish_grads(rt):
    store = rt.store
    for p in store.params:
        gv = store.grad_view(p)
        if p.grad is None:
            p.grad = gv.clone()
        else:
            p.grad.add_(gv)


class CondForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, eng: CondEngine, x, y, eps_u, eps_z, training: bool, need: bool):
        # `need` is decided by the caller: grad mode is always off inside Function.forward
        outs, ectx = eng.forward(x, y, eps_u, eps_z, training=training, save=need)
        ctx.eng, ctx.ectx = eng, ectx
        return outs["x_hat"], outs["y_hat"], outs["enc_z"], outs["enc_u"], outs["mu3"], outs["lv3"]

    @staticmethod
    def backward(ctx, d_xhat, d_yhat, d_enc_z, d_enc_u, d_mu3, d_lv3):
        eng = ctx.eng
        rt = eng.rt
        rt.zero_grads()
        cl = lambda t: None if t is None else t.contiguous().clone()
        eng.backward(ctx.ectx, d_xhat, d_yhat, cl(d_enc_z), cl(d_enc_u), d_mu3, cl(d_lv3))
        _publish_grads(rt)
        ctx.ectx = None
        return (None,) * 8


class VaeForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, eng: VaeEngine, x, eps, training: bool, need: bool):
        outs, ectx = eng.forward(x, eps, training=training, save=need)
        ctx.eng, ctx.ectx = eng, ectx
        return outs["x_hat"], outs["enc"]

    @staticmethod
    def backward(ctx, d_xhat, d_enc):
        eng = ctx.eng
        rt = eng.rt
        rt.zero_grads()
        eng.backward(ctx.ectx, d_xhat, None if d_enc is None else d_enc.contiguous().clone())
        _publish_grads(rt)
        ctx.ectx = None
        return (None,) * 6
