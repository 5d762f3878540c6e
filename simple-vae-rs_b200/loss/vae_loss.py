"""base_loss (loss/vae_loss.py:5-13):  mse = d * (mean((recon_x - x)^2) / (2 gamma^2) + log gamma),
kld = 0.5 * mean_b sum_j (mu^2 + exp(logvar) - 1 - logvar).  Fused kernel + hand-written backward."""
from svrs_native.elbo import base_loss

__all__ = ["base_loss"]
