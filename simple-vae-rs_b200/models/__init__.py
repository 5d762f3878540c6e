"""Public model surface of the reference (models/__init__.py:1-4)."""
from .cond_vae import Cond_SRVAE
from .vae import VAE

__all__ = ["VAE", "Cond_SRVAE"]
