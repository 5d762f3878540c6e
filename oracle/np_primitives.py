"""First-principles numpy definitions of the primitives the reference borrows from torch.

TEST INFRASTRUCTURE ONLY (see oracle/ref_oracle.py header).  The reference's arithmetic lives in
the third-party `torch==2.6.0` (pyproject.toml:18); these functions restate the *published*
definitions of those ops (torch.nn docs) with explicit loops / einsum in float64, so the CUDA
kernels are checked against the maths and not merely against another library build.  They are
O(N*K) python/numpy and meant for small cases only.
"""
from __future__ import annotations

import numpy as np


def conv2d(x, w, b, stride, pad):
    """nn.Conv2d (models/layers.py:231-236): out[n,co,oy,ox] = b[co] +
    sum_{ci,ky,kx} w[co,ci,ky,kx] * x[n,ci,oy*s-p+ky,ox*s-p+kx], zero padding."""
    N, C, H, W = x.shape
    Co, Ci, K, _ = w.shape
    assert Ci == C
    OH = (H + 2 * pad - K) // stride + 1
    OW = (W + 2 * pad - K) // stride + 1
    xp = np.zeros((N, C, H + 2 * pad, W + 2 * pad), dtype=np.float64)
    xp[:, :, pad:pad + H, pad:pad + W] = x
    out = np.zeros((N, Co, OH, OW), dtype=np.float64)
    for ky in range(K):
        for kx in range(K):
            patch = xp[:, :, ky:ky + stride * OH:stride, kx:kx + stride * OW:stride]
            out += np.einsum("nchw,oc->nohw", patch, w[:, :, ky, kx].astype(np.float64))
    return out + b.reshape(1, -1, 1, 1)


def conv_transpose2d_k4s2p1(x, w, b):
    """nn.ConvTranspose2d(k=4,s=2,p=1) (models/layers.py:275-277), weight [Cin,Cout,4,4]:
    out[n,co,2*iy-1+ky,2*ix-1+kx] += w[ci,co,ky,kx] * x[n,ci,iy,ix]."""
    N, Ci, H, W = x.shape
    _, Co, K, _ = w.shape
    full = np.zeros((N, Co, 2 * H + 2, 2 * W + 2), dtype=np.float64)   # index = o + 1
    for ky in range(K):
        for kx in range(K):
            contrib = np.einsum("nchw,co->nohw", x.astype(np.float64), w[:, :, ky, kx].astype(np.float64))
            full[:, :, ky:ky + 2 * H:2, kx:kx + 2 * W:2] += contrib
    return full[:, :, 1:1 + 2 * H, 1:1 + 2 * W] + b.reshape(1, -1, 1, 1)


def batchnorm_train(x, gamma, beta, eps=1e-5):
    """nn.BatchNorm2d in train mode: biased batch variance normalises; returns also the
    (mean, unbiased var) used for the running-stat update."""
    n = x.shape[0] * x.shape[2] * x.shape[3]
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))
    xhat = (x - mean.reshape(1, -1, 1, 1)) / np.sqrt(var.reshape(1, -1, 1, 1) + eps)
    return xhat * gamma.reshape(1, -1, 1, 1) + beta.reshape(1, -1, 1, 1), mean, var * n / (n - 1)


def adam_step(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (train.py:65) single-tensor update, t is the 1-based step."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - lr / (1 - b1 ** t) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps)
    return p, m, v


def clip_coef(grads, max_norm=1.0):
    """clip_grad_norm_ (models/base.py:106)."""
    total = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    return total, min(1.0, max_norm / (total + 1e-6))


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11) - the counter-based generator the device kernels use for
    eps.  counter: 4 uint32, key: 2 uint32 -> 4 uint32."""
    M0, M1 = 0xD2511F53, 0xCD9E8D57
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    c = [int(v) & 0xFFFFFFFF for v in counter]
    k = [int(v) & 0xFFFFFFFF for v in key]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c = [(hi1 ^ c[1] ^ k[0]) & 0xFFFFFFFF, lo1, (hi0 ^ c[3] ^ k[1]) & 0xFFFFFFFF, lo0]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def philox_normal4(seed, stream, idx4, step=0):
    """The device eps convention: element block idx4 (4 consecutive elements) of stream `stream`
    uses counter (idx4_lo, idx4_hi, stream, step), key (seed_lo, seed_hi); the 4 uint32 outputs feed two
    Box-Muller pairs: u1 = (r+1)*2^-32 in (0,1], u2 = r*2^-32 in [0,1);
    n0 = sqrt(-2 ln u1) cos(2 pi u2), n1 = sqrt(-2 ln u1) sin(2 pi u2)."""
    r = philox4x32_10([idx4 & 0xFFFFFFFF, idx4 >> 32, stream, step & 0xFFFFFFFF], [seed & 0xFFFFFFFF, seed >> 32])
    out = []
    for a, b in ((r[0], r[1]), (r[2], r[3])):
        u1 = (a + 1.0) * 2.0 ** -32
        u2 = b * 2.0 ** -32
        rad = np.sqrt(-2.0 * np.log(u1))
        out += [rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)]
    return out
