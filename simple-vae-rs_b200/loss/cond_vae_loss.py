"""cond_loss (loss/cond_vae_loss.py:5-58): negative ELBO terms of the conditional SR VAE.

    mse_y = n_y * (mean((recon_y - y)^2) / (2 gammay^2) + log gammay)          (:43-45)
    kld_u = 0.5 * mean_b sum_j (mu1^2 + exp(lv1) - 1 - lv1)                      (:46)
    mse_x = n_x * (mean((recon_x - x)^2) / (2 gammax^2) + log gammax)          (:47-49)
    kld_z = 0.5 * mean_b sum_j [(lv3 - lv2 - 1) + exp(lv2 - lv3) + (mu2 - mu3)^2 exp(-lv3)]   (:50-57)

1 = q(u|y), 2 = q(z|x), 3 = p(z|y,u).  Same positional signature and return order as the reference; the
arithmetic is ONE fused read-once reduction kernel (svrs_elbo_fwd) with a hand-written backward
(svrs_elbo_bwd), differentiable through torch.autograd.  CUDA tensors only.
"""
from svrs_native.elbo import cond_loss

__all__ = ["cond_loss"]
