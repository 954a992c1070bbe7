"""High-level inference API."""
from .mcmc import MCMC

__all__ = ["MCMC"]
