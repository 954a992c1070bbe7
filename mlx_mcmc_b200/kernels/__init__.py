"""MCMC sampling kernels (host front-ends of the CUDA kernels in csrc/)."""
from .metropolis import metropolis_hastings
from .hmc import hmc
from .nuts import nuts

__all__ = ["metropolis_hastings", "hmc", "nuts"]
