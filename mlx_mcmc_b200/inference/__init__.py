"""User-facing driver of the B200 sampling path.

`MCMC` (inference/mcmc.py) keeps the reference's call convention -- ``MCMC(log_prob).run(...)``, ``summary()``,
``print_summary()`` -- and adds ``diagnostics()`` (R-hat / ESS computed on the device).
"""
from .mcmc import MCMC as MCMC

__all__ = ("MCMC",)
