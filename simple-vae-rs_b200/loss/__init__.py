"""Public loss surface of the reference (loss/__init__.py:1-4)."""
from .cond_vae_loss import cond_loss
from .vae_loss import base_loss

__all__ = ["base_loss", "cond_loss"]
