"""Probability distributions (same six classes and export list as mlx_mcmc/distributions/__init__.py)."""
from .base import Distribution
from .normal import Normal
from .halfnormal import HalfNormal
from .beta import Beta
from .gamma import Gamma
from .exponential import Exponential
from .categorical import Categorical

__all__ = ["Distribution", "Normal", "HalfNormal", "Beta", "Gamma", "Exponential", "Categorical"]
