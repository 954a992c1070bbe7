"""Python front-ends of the three samplers.  Each keeps the signature of its reference counterpart and drives
libb200mcmc.so through `engine.launch_*`; the sampling itself happens in csrc/ (persistent kernels for pointwise
models, lock-step tcgen05 GEMM pipeline for `X @ beta` models)."""
from .hmc import hmc
from .metropolis import metropolis_hastings
from .nuts import nuts

__all__ = ("hmc", "metropolis_hastings", "nuts")
