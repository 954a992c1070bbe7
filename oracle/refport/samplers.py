"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the three samplers of korentomas/mlx-mcmc on the ``mlx.core``
stand-in (oracle/mlx_shim).  Same arithmetic, same random-key consumption order,
same float32/float64 split between array math and python bookkeeping as the
reference, so that with the same stand-in keys the draws are bit-identical to the
reference run here (asserted by oracle/make_golden.py against /root/reference).

Reference lines restated:
  metropolis_port   mlx_mcmc/kernels/metropolis.py:6-101
  hmc_port          mlx_mcmc/kernels/hmc.py:7-206   (grad :53-67, leapfrog :69-100,
                    hamiltonian :102-111, step :113-153, +-5% rule :164-170)
  nuts_port         mlx_mcmc/kernels/nuts.py:16-358 (no_u_turn :119-135, build_tree :137-218,
                    nuts_step :220-285, dual averaging :62-68,298-310,317-320)
  run_port          mlx_mcmc/inference/mcmc.py:38-189 (dispatch, Metropolis seed / seed+1 hand-off)

What is added on top of the reference (none of it changes a number): a ``Tape`` that
stores every random draw under a *slot* name -- (iteration, role, doubling, merge index)
-- plus the per-iteration decisions, so the CUDA path can be fed exactly these draws
and compared decision by decision.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import mlx.core as mx  # oracle stand-in
import numpy as np

MAX_ENERGY_ERROR = 1000.0  # nuts.py:13


# ---------------------------------------------------------------------------------------
@dataclass
class Tape:
    """Slot-addressed record of one sampler run (draws + decisions)."""

    normals: Dict[tuple, np.ndarray] = field(default_factory=dict)   # (it, param) -> z
    uniforms: Dict[tuple, float] = field(default_factory=dict)       # (it, role[, j[, merge]]) -> u
    iters: List[dict] = field(default_factory=list)                  # per-iteration decisions
    grad_evals: int = 0                                              # mx.grad calls
    value_evals: int = 0                                             # log_prob-only calls
    leapfrogs: int = 0


def merge_slot(leaf_index: int, level: int) -> int:
    """Post-order index of the merge that closes at `leaf_index` on tree level `level`
    (levels 0.. close in increasing order at one leaf).  #merges finished before leaf i
    equals i - popcount(i)."""
    return leaf_index - bin(leaf_index).count("1") + level


class _Target:
    """log p, its gradient, leapfrog and energy for a dict of named parameters."""

    def __init__(self, log_prob_fn: Callable, names: List[str], tape: Optional[Tape]):
        self.fn, self.names, self.tape = log_prob_fn, names, tape

    def logp(self, q):
        if self.tape is not None:
            self.tape.value_evals += 1
        return self.fn(q)

    def grad(self, q):
        if self.tape is not None:
            self.tape.grad_evals += 1
        names = self.names

        def positional(*vals):
            return self.fn(dict(zip(names, vals)))

        g = mx.grad(positional, argnums=list(range(len(names))))(*[q[n] for n in names])
        if not isinstance(g, tuple):
            g = (g,)
        return dict(zip(names, g))

    def leapfrog(self, q, p, eps):
        """Two half kicks with a fresh gradient each (hmc.py:69-100, nuts.py:89-111)."""
        if self.tape is not None:
            self.tape.leapfrogs += 1
        g0 = self.grad(q)
        p_half = {n: p[n] + 0.5 * eps * g0[n] for n in self.names}
        q1 = {n: q[n] + eps * p_half[n] for n in self.names}
        g1 = self.grad(q1)
        p1 = {n: p_half[n] + 0.5 * eps * g1[n] for n in self.names}
        return q1, p1

    def energy(self, q, p):
        kinetic = 0.5 * sum(mx.sum(v ** 2) for v in p.values())
        return -self.logp(q) + kinetic

    def draw_momentum(self, q, key, it):
        key, *subs = mx.random.split(key, len(self.names) + 1)
        p = {}
        for n, sk in zip(self.names, subs):
            p[n] = mx.random.normal(q[n].shape, key=sk)
            if self.tape is not None:
                self.tape.normals[(it, n)] = np.array(p[n], dtype=np.float32)
        return p, key


def _store(buf, q, names):
    for n in names:
        buf[n].append(float(q[n]) if q[n].size == 1 else q[n].tolist())


# ---------------------------------------------------------------------------------------
def metropolis_port(log_prob_fn, initial_params, num_samples=1000, proposal_scale=0.1,
                    random_seed=0, tape: Optional[Tape] = None):
    key = mx.random.key(random_seed)
    q = {k: mx.array(v) for k, v in initial_params.items()}
    names = list(q)
    tgt = _Target(log_prob_fn, names, tape)
    lp = tgt.logp(q)
    out = {n: [] for n in names}
    accepted = 0
    for it in range(num_samples):
        key, *subs = mx.random.split(key, len(names) + 1)
        prop = {}
        for n, sk in zip(names, subs):
            z = mx.random.normal(q[n].shape, key=sk)
            if tape is not None:
                tape.normals[(it, n)] = np.array(z, dtype=np.float32)
            prop[n] = q[n] + z * proposal_scale
        lp_prop = tgt.logp(prop)
        key, sk = mx.random.split(key)
        u = mx.random.uniform(key=sk)
        take = float(mx.log(u)) < float(lp_prop - lp)
        if tape is not None:
            tape.uniforms[(it, "accept")] = float(u)
            tape.iters.append({"accept": bool(take), "lp_prop": float(lp_prop)})
        if take:
            q, lp = prop, lp_prop
            accepted += 1
        for n in names:
            out[n].append(float(q[n]))
    return out, accepted / num_samples


# ---------------------------------------------------------------------------------------
def hmc_port(log_prob_fn, initial_params, num_samples=1000, num_warmup=1000, step_size=0.1,
             num_leapfrog_steps=10, adapt_step_size=True, target_accept=0.8, key=None,
             tape: Optional[Tape] = None):
    key = mx.random.key(0) if key is None else key
    q = {k: mx.array(v) for k, v in initial_params.items()}
    names = list(q)
    tgt = _Target(log_prob_fn, names, tape)
    out = {n: [] for n in names}

    def transition(q0, eps, key, it):
        p0, key = tgt.draw_momentum(q0, key, it)
        e0 = tgt.energy(q0, p0)
        qn, pn = q0, p0
        for _ in range(num_leapfrog_steps):
            qn, pn = tgt.leapfrog(qn, pn, eps)
        pn = {n: -v for n, v in pn.items()}
        e1 = tgt.energy(qn, pn)
        key, sk = mx.random.split(key)
        u = mx.random.uniform(shape=(), key=sk)
        ok = bool(float(mx.log(u) < -(e1 - e0)))
        if tape is not None:
            tape.uniforms[(it, "accept")] = float(u)
            tape.iters.append({"accept": ok, "H0": float(e0), "H1": float(e1), "eps": float(eps),
                               "q_prop": {n: np.array(qn[n], dtype=np.float32) for n in names}})
        return (qn if ok else q0), ok, key

    eps = step_size
    n_ok = n_all = 0
    for i in range(num_warmup):
        q, ok, key = transition(q, eps, key, i)
        n_ok += int(ok)
        n_all += 1
        if adapt_step_size and i > 10:          # +-5 % on the cumulative rate (hmc.py:164-170)
            eps = eps * 0.95 if (n_ok / n_all) < target_accept else eps * 1.05
    n_ok = n_all = 0
    for i in range(num_samples):
        q, ok, key = transition(q, eps, key, num_warmup + i)
        n_ok += int(ok)
        n_all += 1
        for n in names:
            out[n].append(float(q[n]))
    return {n: mx.array(out[n]) for n in names}, n_ok / n_all, eps


# ---------------------------------------------------------------------------------------
@dataclass
class _Tree:
    q_lo: dict
    q_hi: dict
    p_lo: dict
    p_hi: dict
    cand: dict
    n: int
    ok: bool
    alpha: float
    n_alpha: int


def nuts_port(log_prob_fn, initial_params, num_samples=1000, num_warmup=1000, step_size=0.1,
              max_tree_depth=10, adapt_step_size=True, target_accept=0.65, key=None,
              tape: Optional[Tape] = None):
    key = mx.random.key(0) if key is None else key
    q = {k: mx.array(v) for k, v in initial_params.items()}
    names = list(q)
    tgt = _Target(log_prob_fn, names, tape)
    out = {n: [] for n in names}

    def still_straight(q_lo, q_hi, p_lo, p_hi):
        span = {n: q_hi[n] - q_lo[n] for n in names}
        d_lo = sum(mx.sum(span[n] * p_lo[n]) for n in names)
        d_hi = sum(mx.sum(span[n] * p_hi[n]) for n in names)
        return float(d_lo) >= 0 and float(d_hi) >= 0

    def grow(qe, pe, u, v, depth, eps, q0, p0, rng, it, j_top, base):
        """Subtree of 2**depth leaves starting one step past the edge (qe, pe); `base` is the
        index of its first leaf inside doubling `j_top` (only used to name merge slots)."""
        if depth == 0:
            q1, p1 = tgt.leapfrog(qe, pe, v * eps)
            h1 = tgt.energy(q1, p1)
            log_slice = float(mx.log(u))
            n1 = 1 if log_slice <= float(-h1) else 0
            h0 = tgt.energy(q0, p0)                                 # recomputed per leaf (nuts.py:169)
            ok = log_slice < float(MAX_ENERGY_ERROR - h1)
            a = min(1.0, float(mx.exp(-h1 + h0)))                   # NaN -> 1.0 (python min)
            if tape is not None:
                tape.iters[-1]["leaves"].append({"j": j_top, "i": base, "H": float(h1), "n": n1, "s": bool(ok)})
            return _Tree(q1, q1, p1, p1, q1, n1, ok, a, 1)
        rng, k_first, k_second = mx.random.split(rng, 3)
        t = grow(qe, pe, u, v, depth - 1, eps, q0, p0, k_first, it, j_top, base)
        if t.ok:
            half = 1 << (depth - 1)
            if v == -1:
                t2 = grow(t.q_lo, t.p_lo, u, v, depth - 1, eps, q0, p0, k_second, it, j_top, base + half)
                t.q_lo, t.p_lo = t2.q_lo, t2.p_lo
            else:
                t2 = grow(t.q_hi, t.p_hi, u, v, depth - 1, eps, q0, p0, k_second, it, j_top, base + half)
                t.q_hi, t.p_hi = t2.q_hi, t2.p_hi
            rng, k_pick = mx.random.split(rng)
            w = float(mx.random.uniform(key=k_pick))
            if tape is not None:
                tape.uniforms[(it, "merge", j_top, merge_slot(base + 2 * half - 1, depth - 1))] = w
            if w < t2.n / max(t.n + t2.n, 1.0):
                t.cand = t2.cand
            t.n += t2.n
            t.alpha += t2.alpha
            t.n_alpha += t2.n_alpha
            t.ok = t2.ok and still_straight(t.q_lo, t.q_hi, t.p_lo, t.p_hi)
        return t

    def transition(q0, eps, it, key):
        p0, key = tgt.draw_momentum(q0, key, it)
        h0 = tgt.energy(q0, p0)
        key, sk = mx.random.split(key)
        us = mx.random.uniform(key=sk)
        log_u = float(-h0) + float(mx.log(us))            # float64 ...
        u = mx.exp(mx.array(log_u))                       # ... back to float32: underflows for H0 >~ 103
        if tape is not None:
            tape.uniforms[(it, "slice")] = float(us)
            tape.iters.append({"H0": float(h0), "eps": float(eps), "doublings": [], "leaves": []})
        q_lo = q_hi = q0
        p_lo = p_hi = p0
        cand = q0
        j, n, ok = 0, 1, True
        a_sum, a_cnt = 0.0, 0
        while ok and j < max_tree_depth:
            key, k_dir, k_tree, k_take = mx.random.split(key, 4)
            ud = float(mx.random.uniform(key=k_dir))
            v = 1 if ud < 0.5 else -1
            if v == -1:
                t = grow(q_lo, p_lo, u, v, j, eps, q0, p0, k_tree, it, j, 0)
                q_lo, p_lo = t.q_lo, t.p_lo
            else:
                t = grow(q_hi, p_hi, u, v, j, eps, q0, p0, k_tree, it, j, 0)
                q_hi, p_hi = t.q_hi, t.p_hi
            took = False
            ut = None
            if t.ok:
                ut = float(mx.random.uniform(key=k_take))
                if ut < min(1.0, t.n / max(n, 1.0)):
                    cand, took = t.cand, True
            n += t.n
            ok = t.ok and still_straight(q_lo, q_hi, p_lo, p_hi)
            a_sum += t.alpha
            a_cnt += t.n_alpha
            if tape is not None:
                tape.uniforms[(it, "dir", j)] = ud
                if ut is not None:
                    tape.uniforms[(it, "take", j)] = ut
                tape.iters[-1]["doublings"].append(
                    {"j": j, "v": v, "n_sub": t.n, "s_sub": bool(t.ok), "took": took, "s": bool(ok), "n": n})
            j += 1
        mean_alpha = a_sum / max(a_cnt, 1.0)
        if tape is not None:
            tape.iters[-1].update(depth=j, alpha=mean_alpha,
                                  q_new={nm: np.array(cand[nm], dtype=np.float32) for nm in names})
        return cand, mean_alpha, j, key

    mu = mx.log(10 * step_size)      # float32 array (nuts.py:63)
    eps = step_size
    eps_bar, h_bar = 1.0, 0.0
    gamma, t0, kappa = 0.05, 10.0, 0.75
    for m in range(num_warmup):
        q, a, _, key = transition(q, eps, m, key)
        if adapt_step_size:
            eta = 1.0 / (m + t0)
            h_bar = (1 - eta) * h_bar + eta * (target_accept - a)
            log_eps = mu - (math.sqrt(m + 1) / gamma) * h_bar
            log_eps = max(min(log_eps, 10.0), -10.0)
            eps = float(mx.exp(mx.array(log_eps)))
            w = float(m + 1) ** (-kappa)
            eps_bar = float(mx.exp(mx.array(w * math.log(eps) + (1 - w) * math.log(eps_bar))))
    if adapt_step_size:
        eps = eps_bar
    hits = 0
    for m in range(num_samples):
        q, a, _, key = transition(q, eps, m + num_warmup, key)
        _store(out, q, names)
        hits += int(a > 0.5)
    return {n: mx.array(out[n]) for n in names}, hits / num_samples, eps


# ---------------------------------------------------------------------------------------
def run_port(log_prob_fn, initial_params, num_samples=1000, num_warmup=1000, method="metropolis",
             proposal_scale=0.1, random_seed=0, tape: Optional[Tape] = None, **kw):
    """MCMC.run restated; returns (dict of numpy draws, acceptance_rate)."""
    if method == "hmc":
        s, a, _ = hmc_port(log_prob_fn, initial_params, num_samples=num_samples, num_warmup=num_warmup,
                           key=mx.random.key(random_seed), tape=tape, **kw)
        return {k: np.array(v) for k, v in s.items()}, a
    if method == "nuts":
        s, a, _ = nuts_port(log_prob_fn, initial_params, num_samples=num_samples, num_warmup=num_warmup,
                            key=mx.random.key(random_seed), tape=tape, **kw)
        return {k: np.array(v) for k, v in s.items()}, a
    if method != "metropolis":
        raise ValueError(f"Unknown sampling method: {method}")
    start = initial_params
    if num_warmup > 0:
        w, _ = metropolis_port(log_prob_fn, initial_params, num_samples=num_warmup,
                               proposal_scale=proposal_scale, random_seed=random_seed, **kw)
        start = {k: v[-1] for k, v in w.items()}
    s, a = metropolis_port(log_prob_fn, start, num_samples=num_samples, proposal_scale=proposal_scale,
                           random_seed=random_seed + 1 if num_warmup > 0 else random_seed, tape=tape, **kw)
    return {k: np.array(v) for k, v in s.items()}, a
