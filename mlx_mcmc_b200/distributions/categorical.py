"""Categorical distribution (reference: mlx_mcmc/distributions/categorical.py)."""
from __future__ import annotations

import numpy as np

from .. import core as mx
from ..tracer import Sym, UnsupportedOpError
from .base import Distribution, f32


class Categorical(Distribution):
    """Categorical over K classes from `probs` (normalised, logits = log p) or `logits` (probs by softmax;
    as in the reference, log_prob gathers the RAW logits, categorical.py:58-59,85).  Discrete and
    parameter-free on the sampling path: with concrete arguments log_prob is a constant that the tracer
    folds into the model; traced probs/logits are not lowered (no traceable reference model uses them)."""

    def __init__(self, probs=None, logits=None):
        if probs is None and logits is None:
            raise ValueError("Either probs or logits must be specified")
        if probs is not None and logits is not None:
            raise ValueError("Only one of probs or logits can be specified")
        for v in (probs, logits):
            if isinstance(v, Sym) or (isinstance(v, (list, tuple)) and any(isinstance(e, Sym) for e in v)):
                raise UnsupportedOpError("unsupported op: Categorical with traced probs/logits")
        if probs is not None:
            p = f32(probs)
            self.probs = p / np.sum(p)
            with np.errstate(divide="ignore"):
                self.logits = np.log(self.probs)
        else:
            self.logits = f32(logits)
            e = np.exp(self.logits - np.max(self.logits))
            self.probs = e / np.sum(e)
        self.num_categories = self.probs.shape[0]

    def log_prob(self, value):
        if isinstance(value, Sym):
            raise UnsupportedOpError("unsupported op: Categorical.log_prob of a traced (continuous) value")
        k = np.asarray(value).astype(np.int32)
        ok = (k >= 0) & (k < self.num_categories)
        return np.where(ok, self.logits[np.where(ok, k, 0)], np.float32(-np.inf)).astype(np.float32)

    def _device_sample_spec(self):
        return 6, 0.0, 0.0, np.asarray(self.probs, dtype=np.float64)   # B2M_SAMPLE_CATEGORICAL

    def sample(self, key, shape=()):
        u = mx.random.uniform(shape=shape, key=key)
        return np.sum(np.expand_dims(u, -1) > np.cumsum(self.probs), axis=-1).astype(np.int32)

    def entropy(self):
        with np.errstate(divide="ignore"):
            lp = np.where(self.probs > 0, np.log(self.probs), np.float32(0.0))
        return -np.sum(self.probs * lp)

    def mode(self):
        return np.argmax(self.probs).astype(np.int32)

    def __repr__(self):
        return f"Categorical(num_categories={self.num_categories})"
