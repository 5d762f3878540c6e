"""Parameter containers for the two conv blocks of the reference (models/layers.py:217-297).

The blocks keep real nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d children under the reference's
attribute names (`conv`, `downsample` / `upsample`, `bn`) so that default initialisation, `state_dict()`
keys and shapes, `.to()` and optimizers behave exactly like the reference.  The arithmetic is NOT
torch's: models run these blocks through the sm_100a kernel library (svrs_native.engine), where
  down_block = conv3x3 s1 p1 -> conv4x4 s2 p1 -> BatchNorm -> ReLU
  up_block   = conv3x3 s1 p1 -> convT4x4 s2 p1 -> BatchNorm -> ReLU
are multi-tap implicit GEMMs with fused bias epilogues.  The reference's unused helpers
(downsample_sequence, upsample_sequence, self_attention, residual; layers.py:25-214,300-369) are dead
code in the reference and are intentionally not reproduced.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _Block(nn.Module):
    def __init__(self, with_relu: bool, with_bn: bool):
        super().__init__()
        self.with_relu = with_relu
        self.with_bn = with_bn
        self._rt = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Stand-alone use of a block (inference only; training goes through the owning model's fused
        forward/backward).  NCHW in, NCHW out, CUDA only."""
        from svrs_native.engine import Runtime, plan_sequential, _require_cuda

        _require_cuda(x, "input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("stand-alone block forward is inference-only; wrap the call in torch.no_grad()")
        if self._rt is None:
            net = plan_sequential(self.__class__.__name__, nn.Sequential(self))
            self._rt = (Runtime(self, [net]), net)
        rt, net = self._rt
        rt.ensure()
        rt.packs_dirty = True
        rt.pack_weights()
        n, c, h, w = x.shape
        xin = rt.to_nhwc(x.contiguous().float(), c * h * w, n, c, h, w)
        out, _ = rt.net_forward(net, xin, self.training, save=False)
        n, oh, ow, oc = out.shape
        res = torch.empty((n, oc, oh, ow), device=x.device, dtype=torch.float32)
        rt.to_nchw(out, res, oc * oh * ow)
        return res


class down_block(_Block):
    def __init__(self, in_channels, out_channels, with_relu=True, with_bn=True):
        super().__init__(with_relu, with_bn)
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.downsample = nn.Conv2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1)
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)


class up_block(_Block):
    def __init__(self, in_channels, out_channels, with_relu=True, with_bn=True):
        super().__init__(with_relu, with_bn)
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.upsample = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1)
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)


def conv3(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)
