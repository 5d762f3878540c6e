"""CPU oracle for the (Cond_)SRVAE training-step hot path.  TEST INFRASTRUCTURE ONLY.

This module is a *functional restatement* (no nn.Module, no reference import) of
the algorithm the reference executes on its training path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product path (`simple-vae-rs_b200/`) never does.

Where the arithmetic lives: the reference delegates every op to the third-party
package `torch==2.6.0` (pyproject.toml:18, uv.lock:1038-1039; not vendored).  The
oracle therefore restates the reference's *call sequence* with `torch.nn.functional`
on CPU fp32 (the same ATen kernels the reference reaches on CPU), and
`oracle/np_primitives.py` restates the published definition of each primitive
(conv, transposed conv, batch-norm, Adam, clip) in numpy from first principles.

Parity pin: the reference's own tests hold NO numeric vectors for this path
(tests/test_models.py asserts shapes only; tests/test_training.py asserts
scheduler.last_epoch).  The oracle is pinned instead against outputs of the
reference itself, run in the build container by `oracle/make_golden.py`
(imports /root/reference unmodified with lpips/skimage/matplotlib stubbed) and
committed as fixtures under `tests/golden/`.

All citations are file:line into the reference tree.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5       # nn.BatchNorm2d default (models/layers.py:237,278)
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------
def cond_latent_sizes(cr: float, patch_size: int) -> Tuple[int, int]:
    """models/cond_vae.py:21-22."""
    latent = int((patch_size * patch_size * 4 / cr) // 256) * 256
    return latent, latent // 4


def vae_latent_size(cr: float, patch_size: int) -> int:
    """models/vae.py:29-31."""
    return int((patch_size * patch_size * 4 // cr) // 16) * 16


# ----------------------------------------------------------------------------
# blocks (models/layers.py)
# ----------------------------------------------------------------------------
class BNState:
    """Collects running-stat updates so the caller decides whether to apply them."""

    def __init__(self, sd: SD, training: bool, update_running: bool = True):
        self.sd = sd
        self.training = training
        self.update_running = update_running


def _bn(st: BNState, prefix: str, x: Tensor) -> Tensor:
    """nn.BatchNorm2d forward (models/layers.py:237,252-253 / 278,293-294).

    Train: batch statistics (biased variance) normalise; running stats are updated with
    momentum 0.1 and the *unbiased* variance; num_batches_tracked += 1.  Eval: running stats.
    """
    sd = st.sd
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if st.training:
        if st.update_running:
            sd[prefix + ".num_batches_tracked"] += 1
            out = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"],
                               True, BN_MOMENTUM, BN_EPS)
        else:
            out = F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"],
                               True, BN_MOMENTUM, BN_EPS)
        return out
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"],
                        False, BN_MOMENTUM, BN_EPS)


def down_block(st: BNState, p: str, x: Tensor) -> Tensor:
    """models/layers.py:240-256: conv3x3 s1 p1 -> conv4x4 s2 p1 -> BN -> ReLU."""
    sd = st.sd
    x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=1, padding=1)
    x = F.conv2d(x, sd[p + ".downsample.weight"], sd[p + ".downsample.bias"], stride=2, padding=1)
    x = _bn(st, p + ".bn", x)
    return F.relu(x)


def up_block(st: BNState, p: str, x: Tensor) -> Tensor:
    """models/layers.py:281-297: conv3x3 s1 p1 -> convT4x4 s2 p1 -> BN -> ReLU."""
    sd = st.sd
    x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=1, padding=1)
    x = F.conv_transpose2d(x, sd[p + ".upsample.weight"], sd[p + ".upsample.bias"], stride=2, padding=1)
    x = _bn(st, p + ".bn", x)
    return F.relu(x)


def conv3(sd: SD, p: str, x: Tensor) -> Tensor:
    """bare nn.Conv2d k3 s1 p1 + bias, NO activation (cond_vae.py:30-46 etc.)."""
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=1, padding=1)


# ----------------------------------------------------------------------------
# Cond_SRVAE sub-networks (models/cond_vae.py:27-231)
# ----------------------------------------------------------------------------
def encoder_y(st: BNState, y: Tensor) -> Tensor:
    """cond_vae.py:27-49."""
    h = down_block(st, "encoder_y.0", y)
    h = down_block(st, "encoder_y.1", h)
    for i in (2, 3, 4, 5):
        h = conv3(st.sd, f"encoder_y.{i}", h)
    return h.flatten(1)


def encoder_x(st: BNState, x: Tensor) -> Tensor:
    """cond_vae.py:83-108."""
    h = down_block(st, "encoder_x.0", x)
    h = down_block(st, "encoder_x.1", h)
    h = down_block(st, "encoder_x.2", h)
    for i in (3, 4, 5, 6):
        h = conv3(st.sd, f"encoder_x.{i}", h)
    return h.flatten(1)


def y_to_z(st: BNState, y: Tensor) -> Tensor:
    """cond_vae.py:146-165."""
    h = down_block(st, "y_to_z.0", y)
    h = down_block(st, "y_to_z.1", h)
    h = down_block(st, "y_to_z.2", h)
    h = conv3(st.sd, "y_to_z.3", h)
    h = conv3(st.sd, "y_to_z.4", h)
    return h.flatten(1)


def u_to_z(st: BNState, u: Tensor, latent_y: int, P: int) -> Tensor:
    """cond_vae.py:167-189 (Unflatten re-views the flat u buffer)."""
    h = u.unflatten(1, (latent_y // 16, P // 16, P // 16))
    h = conv3(st.sd, "u_to_z.1", h)
    h = conv3(st.sd, "u_to_z.2", h)
    return h.flatten(1)


def prior_head(st: BNState, name: str, j: Tensor, latent: int, P: int) -> Tensor:
    """cond_vae.py:191-231; logvar head ends in Hardtanh(-7, 7) (:230)."""
    h = j.unflatten(1, (latent * 2 // 16, P // 16, P // 16))
    h = conv3(st.sd, name + ".1", h)
    h = conv3(st.sd, name + ".2", h)
    h = h.flatten(1)
    if name == "logvar_u_y_to_z":
        h = F.hardtanh(h, -7.0, 7.0)
    return h


def decoder_y(st: BNState, u: Tensor, latent_y: int, P: int) -> Tensor:
    """cond_vae.py:51-81."""
    h = u.unflatten(1, (latent_y // 64, P // 8, P // 8))
    h = up_block(st, "decoder_y.1", h)
    h = up_block(st, "decoder_y.2", h)
    for i in (3, 4, 5, 6):
        h = conv3(st.sd, f"decoder_y.{i}", h)
    return torch.sigmoid(h)


def decoder_x(st: BNState, s: Tensor, latent: int, P: int) -> Tensor:
    """cond_vae.py:110-144."""
    h = s.unflatten(1, (latent * 2 // 64, P // 8, P // 8))
    h = up_block(st, "decoder_x.1", h)
    h = up_block(st, "decoder_x.2", h)
    h = up_block(st, "decoder_x.3", h)
    for i in (4, 5, 6, 7):
        h = conv3(st.sd, f"decoder_x.{i}", h)
    return torch.sigmoid(h)


def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """cond_vae.py:261-265 / vae.py:94-98 with eps injected instead of torch.randn_like."""
    std = torch.exp(0.5 * logvar)
    return mu + eps * std


def cond_forward(sd: SD, cr: float, P: int, x: Tensor, y: Tensor, eps_u: Tensor, eps_z: Tensor,
                 training: bool = True, update_running: bool = True):
    """Cond_SRVAE.forward (cond_vae.py:275-286).  Returns the 8-tuple in the reference's order.

    RNG draw order is u first, z second.  y_to_z runs TWICE (z_cond :239 and
    decode_x :271), so its BN running stats update twice per call.
    """
    L, Lu = cond_latent_sizes(cr, P)
    st = BNState(sd, training, update_running)
    mu_u, logvar_u = torch.chunk(encoder_y(st, y), 2, dim=1)          # :251-254
    u = reparameterize(mu_u, logvar_u, eps_u)                          # :277
    mu_z, logvar_z = torch.chunk(encoder_x(st, x), 2, dim=1)          # :256-259
    z = reparameterize(mu_z, logvar_z, eps_z)                          # :279
    # z_cond (:237-249)
    yz = y_to_z(st, y)
    uz = u_to_z(st, u, Lu, P)
    joint = torch.cat((yz, uz), dim=1)
    mu_z_uy = prior_head(st, "mu_u_y_to_z", joint, L, P)
    logvar_z_uy = prior_head(st, "logvar_u_y_to_z", joint, L, P)
    # decode_x (:270-273): y_to_z again
    y_enc = y_to_z(st, y)
    x_hat = decoder_x(st, torch.cat((y_enc, z), dim=1), L, P)
    y_hat = decoder_y(st, u, Lu, P)                                    # :267-268
    return x_hat, y_hat, mu_z, logvar_z, mu_u, logvar_u, mu_z_uy, logvar_z_uy


def cond_sample(sd: SD, cr: float, P: int, y: Tensor, eps_u: Tensor, eps_s: Tensor,
                training: bool = False):
    """Cond_SRVAE.sample (cond_vae.py:299-318): S posterior-predictive decodes for ONE LR patch.

    y: [1,4,P/2,P/2]; eps_u: [1,Lu']; eps_s: [S, L'].
    """
    L, Lu = cond_latent_sizes(cr, P)
    st = BNState(sd, training, update_running=training)
    mu_u, logvar_u = torch.chunk(encoder_y(st, y), 2, dim=1)
    u = reparameterize(mu_u, logvar_u, eps_u)
    yz = y_to_z(st, y)
    uz = u_to_z(st, u, Lu, P)
    joint = torch.cat((yz, uz), dim=1)
    mu3 = prior_head(st, "mu_u_y_to_z", joint, L, P)
    lv3 = prior_head(st, "logvar_u_y_to_z", joint, L, P)
    z = mu3 + eps_s * torch.exp(0.5 * lv3)
    yy = y.expand(z.size(0), -1, -1, -1)
    y_enc = y_to_z(st, yy)
    return decoder_x(st, torch.cat((y_enc, z), dim=1), L, P)


# ----------------------------------------------------------------------------
# VAE (models/vae.py:36-107)
# ----------------------------------------------------------------------------
def vae_forward(sd: SD, cr: float, P: int, x: Tensor, eps: Tensor, training: bool = True,
                update_running: bool = True):
    """VAE.forward (vae.py:103-107) -> (x_hat, mu, logvar)."""
    L = vae_latent_size(cr, P)
    st = BNState(sd, training, update_running)
    h = down_block(st, "encoder.0", x)
    h = down_block(st, "encoder.1", h)
    for i in (2, 3, 4, 5):
        h = conv3(sd, f"encoder.{i}", h)
    mu, logvar = h.flatten(1).chunk(2, dim=1)                          # vae.py:89-92
    z = reparameterize(mu, logvar, eps)
    h = z.unflatten(1, (L // 64, P // 4, P // 4))                      # vae.py:61-63
    h = up_block(st, "decoder.1", h)
    h = up_block(st, "decoder.2", h)
    for i in (3, 4, 5, 6):
        h = conv3(sd, f"decoder.{i}", h)
    return torch.sigmoid(h), mu, logvar


# ----------------------------------------------------------------------------
# losses (loss/vae_loss.py:5-13, loss/cond_vae_loss.py:5-58)
# ----------------------------------------------------------------------------
def base_loss(recon_x, x, mu, logvar, gamma):
    """loss/vae_loss.py:5-13.  NLL is a SUM over all elements, KL a MEAN over the batch (Q2)."""
    d = recon_x.shape[0] * recon_x.shape[1] * recon_x.shape[2] * recon_x.shape[3]
    mse = d * (F.mse_loss(recon_x, x, reduction="mean") / (2 * gamma.pow(2)) + gamma.log())
    kld = 0.5 * torch.sum(mu.pow(2) + logvar.exp() - 1 - logvar, dim=1).mean()
    return mse, kld


def cond_loss(recon_x, x, recon_y, y, mu1, logvar1, mu2, logvar2, mu3, logvar3, gammax, gammay):
    """loss/cond_vae_loss.py:39-58.  1 = u (q(u|y)), 2 = q(z|x), 3 = p(z|y,u)."""
    n_y = recon_y.numel()
    n_x = recon_x.numel()
    mse_y = n_y * (F.mse_loss(recon_y, y, reduction="mean") / (2 * gammay.pow(2)) + gammay.log())
    kld_u = 0.5 * torch.sum(mu1.pow(2) + logvar1.exp() - 1 - logvar1, dim=1).mean()
    mse_x = n_x * (F.mse_loss(recon_x, x, reduction="mean") / (2 * gammax.pow(2)) + gammax.log())
    kld_z = (0.5 * (torch.sum(logvar3 - logvar2 - 1, dim=1)
                    + torch.sum((logvar2 - logvar3).exp(), dim=1)
                    + torch.sum((mu2 - mu3).pow(2) * ((-logvar3).exp()), dim=1))).mean()
    return mse_x, kld_u, mse_y, kld_z


# ----------------------------------------------------------------------------
# optimisation step (models/base.py:103-107, train.py:65)
# ----------------------------------------------------------------------------
PARAM_SUFFIXES = (".weight", ".bias")


def param_keys(sd: SD) -> List[str]:
    """Keys of state_dict entries that are nn.Parameters (everything but BN buffers), in
    state_dict order == module.parameters() order."""
    return [k for k in sd if k.endswith(PARAM_SUFFIXES)]


def clip_coef(grads: List[Tensor], max_norm: float = 1.0) -> Tuple[Tensor, Tensor]:
    """torch.nn.utils.clip_grad_norm_ (models/base.py:106): L2 over per-tensor L2 norms;
    coef = min(1, max_norm / (total + 1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef


class AdamState:
    """torch.optim.Adam defaults (train.py:65): lr 1e-4, betas (.9,.999), eps 1e-8, no wd/amsgrad."""

    def __init__(self, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.t = 0
        self.m: Dict[str, Tensor] = {}
        self.v: Dict[str, Tensor] = {}

    def step(self, params: Dict[str, Tensor], grads: Dict[str, Tensor]):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        for k, p in params.items():
            g = grads[k]
            m = self.m.setdefault(k, torch.zeros_like(p))
            v = self.v.setdefault(k, torch.zeros_like(p))
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-(self.lr / bc1))


def cond_train_step(sd: SD, gam: Dict[str, Tensor], opt: AdamState, cr: float, P: int,
                    x: Tensor, y: Tensor, eps_u: Tensor, eps_z: Tensor,
                    max_norm: float = 1.0, return_grads: bool = False):
    """One optimisation step exactly as BaseVAE.fit runs it (models/base.py:103-107):
    zero_grad -> train_step (cond_vae.py:326-354) -> backward -> clip(module params only, Q3)
    -> Adam over params + {gammax, gammay}.  Mutates sd / gam / opt in place."""
    pk = param_keys(sd)
    leaves = {k: sd[k].detach().requires_grad_(True) for k in pk}
    work = dict(sd)
    work.update(leaves)
    gx = gam["gammax"].detach().requires_grad_(True)
    gy = gam["gammay"].detach().requires_grad_(True)
    outs = cond_forward(work, cr, P, x, y, eps_u, eps_z, training=True)
    x_hat, y_hat, mu_z, lv_z, mu_u, lv_u, mu3, lv3 = outs
    mse_x, kld_u, mse_y, kld_z = cond_loss(x_hat, x, y_hat, y, mu_u, lv_u, mu_z, lv_z, mu3, lv3, gx, gy)
    loss = mse_x + kld_u + mse_y + kld_z
    loss.backward()
    grads = {k: leaves[k].grad for k in pk}
    raw = {k: g.clone() for k, g in grads.items()} if return_grads else None
    total, coef = clip_coef([grads[k] for k in pk], max_norm)
    for k in pk:
        grads[k].mul_(coef)
    params = {k: sd[k] for k in pk}
    params["gammax"], params["gammay"] = gam["gammax"], gam["gammay"]
    grads["gammax"], grads["gammay"] = gx.grad, gy.grad
    with torch.no_grad():
        opt.step(params, grads)
    terms = dict(loss=loss.detach(), mse_x=mse_x.detach(), kld_u=kld_u.detach(),
                 mse_y=mse_y.detach(), kld_z=kld_z.detach(), grad_norm=total.detach())
    if return_grads:
        raw["gammax"], raw["gammay"] = gx.grad.clone(), gy.grad.clone()
        return terms, outs, raw
    return terms


def vae_train_step(sd: SD, gam: Dict[str, Tensor], opt: AdamState, cr: float, P: int,
                   x: Tensor, eps: Tensor, max_norm: float = 1.0, return_grads: bool = False):
    """VAE flavour of the step (vae.py:109-120 + base.py:103-107)."""
    pk = param_keys(sd)
    leaves = {k: sd[k].detach().requires_grad_(True) for k in pk}
    work = dict(sd)
    work.update(leaves)
    g = gam["gamma"].detach().requires_grad_(True)
    x_hat, mu, logvar = vae_forward(work, cr, P, x, eps, training=True)
    mse, kld = base_loss(x_hat, x, mu, logvar, g)
    loss = mse + kld
    loss.backward()
    grads = {k: leaves[k].grad for k in pk}
    raw = {k: t.clone() for k, t in grads.items()} if return_grads else None
    total, coef = clip_coef([grads[k] for k in pk], max_norm)
    for k in pk:
        grads[k].mul_(coef)
    params = {k: sd[k] for k in pk}
    params["gamma"] = gam["gamma"]
    grads["gamma"] = g.grad
    with torch.no_grad():
        opt.step(params, grads)
    terms = dict(loss=loss.detach(), mse=mse.detach(), kld=kld.detach(), grad_norm=total.detach())
    if return_grads:
        raw["gamma"] = g.grad.clone()
        return terms, (x_hat, mu, logvar), raw
    return terms


# ----------------------------------------------------------------------------
# grid patching + normalisation (dataset.py:220-247,265-274; utils.py:4-23)
# ----------------------------------------------------------------------------
def normalize_image(image: Tensor) -> Tensor:
    """utils.py:4-23: per-image per-channel min-max, (x-min)/(max-min+1e-5)."""
    if image.ndim == 3:
        mn = image.amin(dim=(1, 2), keepdim=True)
        mx = image.amax(dim=(1, 2), keepdim=True)
    elif image.ndim == 4:
        mn = image.amin(dim=(2, 3), keepdim=True)
        mx = image.amax(dim=(2, 3), keepdim=True)
    else:
        raise ValueError("Input image must be 3D or 4D tensor.")
    return (image - mn) / (mx - mn + 1e-5)


def select_crop(img: Tensor, patch_size: int, index: int) -> Tensor:
    """dataset.py:220-228: row-major patch `index` of a [C,H,W] tile."""
    num = img.shape[2] // patch_size
    row, col = index // num, index % num
    return img[:, row * patch_size:(row + 1) * patch_size, col * patch_size:(col + 1) * patch_size]


def grid_crop(img: Tensor, patch_size: int) -> Tensor:
    """dataset.py:230-247: all full patches of a [C,H,W] tile, row-major, stacked on dim 0."""
    _, h, w = img.shape
    out = []
    for row in range(0, h, patch_size):
        for col in range(0, w, patch_size):
            if row + patch_size <= h and col + patch_size <= w:
                out.append(img[:, row:row + patch_size, col:col + patch_size])
    return torch.stack(out, dim=0)


def grid_batch(lr_tiles: Tensor, hr_tiles: Tensor, patch_size: int) -> Tuple[Tensor, Tensor]:
    """Grid mode for a batch of tiles: per tile grid_crop (LR at patch_size//2, HR at patch_size;
    dataset.py:180-185), per-patch normalize_image (4-D path, utils.py:17-20), then grid_collate's
    torch.cat over tiles (dataset.py:265-274) -> tile-major, then row-major patch order.
    Returns (y_patches, x_patches) = (LR, HR), the (y, x) batch order of cond_vae.py:327."""
    ys, xs = [], []
    for t in range(lr_tiles.shape[0]):
        ys.append(normalize_image(grid_crop(lr_tiles[t], patch_size // 2)))
        xs.append(normalize_image(grid_crop(hr_tiles[t], patch_size)))
    return torch.cat(ys, dim=0), torch.cat(xs, dim=0)


# ----------------------------------------------------------------------------
# portable seeded tensors for fixtures
# ----------------------------------------------------------------------------
# torch's CPU generator is NOT bit-stable across hosts for large tensors (the vectorised fill depends on the
# CPU's ISA / thread split: fixtures minted in the build container did not reproduce on the GPU box).  Fixtures
# therefore derive weights, inputs and eps from numpy's PCG64, which is.
def portable_state_dict(template: SD, seed: int) -> SD:
    """Same distribution as torch's default Conv/ConvT init (kaiming_uniform(a=sqrt 5) == U(-1/sqrt(fan_in), .),
    fan_in = weight.size(1)*k*k); BatchNorm weight/bias/buffers keep their defaults."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out, bound = {}, None
    for k, v in template.items():
        if v.dim() == 4:
            bound = 1.0 / math.sqrt(v.shape[1] * v.shape[2] * v.shape[3])
            out[k] = torch.from_numpy(rng.uniform(-bound, bound, size=tuple(v.shape)).astype(np.float32))
        elif k.endswith(".bias") and ".bn." not in k:
            out[k] = torch.from_numpy(rng.uniform(-bound, bound, size=tuple(v.shape)).astype(np.float32))
        else:
            out[k] = v.detach().clone()
    return out


class PortableRng:
    def __init__(self, seed: int):
        import numpy as np
        self._rng = np.random.default_rng(seed)

    def rand(self, *shape) -> Tensor:
        import numpy as np
        return torch.from_numpy(self._rng.random(size=shape).astype(np.float32))

    def randn(self, *shape) -> Tensor:
        import numpy as np
        return torch.from_numpy(self._rng.standard_normal(size=shape).astype(np.float32))
