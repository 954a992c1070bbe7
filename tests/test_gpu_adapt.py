"""Round-2 features on the GPU, through the C ABI: constraint transforms, diagonal mass-matrix adaptation (SURVEY.md 8f
row 3), the radix-select quantiles behind MCMC.summary (8f row 2), and the decision-level parity of the tensor-core
NUTS path at a mid size (64 coefficients x 2048 observations) against the reference's own transitions.

The reference has no transforms and an identity mass matrix (README.md:165,220, PROGRESS.md:119 list both as planned),
so their checks are against float64 evaluations of the same maths and against exact / quadrature posteriors.
"""
import math

import numpy as np
import pytest
import torch

import mlx_mcmc_b200 as B
import mlx_mcmc_b200.core as mx
from mlx_mcmc_b200 import _cabi, workloads as W
from mlx_mcmc_b200.diagnostics import device_quantiles
from mlx_mcmc_b200.engine import ChainState, compile_model, launch_hmc, launch_nuts
from mlx_mcmc_b200.kernels._common import mass_from_window
from oracle.ns import ns as ons, value_and_grad   # checker only

pytestmark = pytest.mark.gpu


def mcse_ok(x, mean, sd, k=4.0):
    """x: (C, S) draws of independent chains; grand mean within k standard errors (from the spread of chain means) and
    pooled sd within k standard errors + 2 %."""
    x = np.asarray(x, dtype=np.float64)
    cm = x.mean(axis=1)
    se = cm.std(ddof=1) / math.sqrt(len(cm))
    ok_mean = abs(cm.mean() - mean) <= k * se + 1e-3 * sd
    v = x.var(axis=1, ddof=1)
    v_hat = v.mean() + cm.var(ddof=1)
    v_se = v.std(ddof=1) / math.sqrt(len(v))
    return ok_mean and abs(v_hat - sd ** 2) <= k * v_se + 0.03 * sd ** 2


# ------------------------------------------------------------------------------------------ transforms: value + gradient
@pytest.mark.parametrize("name", ["c1_normal", "c2_event_rate", "c5_ab_test", "t_halfnormal_scale", "t_regression_sigma"])
def test_transformed_logp_grad_matches_float64_chain_rule(cuda, name):
    """b2m_logp_grad of a model with transforms = log p(T(u)) + log|J| and its gradient by the chain rule, against the
    oracle's float64 evaluation of the untransformed density (both model classes)."""
    fn, init, _ = W.ALL_SMALL[name](B.ns)
    fo, _, _ = W.ALL_SMALL[name](ons)
    model = compile_model(fn, init, cache=False, transforms="auto")
    codes = model.tf_codes
    assert codes.any(), "the model has a constrained parameter"
    rng = np.random.default_rng(3)
    for _ in range(4):
        u = (model.pack(init, 1).cpu().numpy()[0] + 0.4 * rng.standard_normal(model.D)).astype(np.float32)
        u64 = u.astype(np.float64)
        theta = np.where(codes == 1, np.exp(u64), np.where(codes == 2, 1 / (1 + np.exp(-u64)), u64))
        params = {k: (theta[off] if not shp else theta[off:off + n]) for k, (off, n, shp) in model.layout.items()}
        lp64, g64 = value_and_grad(fo, params, "float64")
        g = np.concatenate([np.atleast_1d(g64[k]).ravel() for k in model.layout])
        jac = np.where(codes == 1, theta, np.where(codes == 2, theta * (1 - theta), 1.0))
        dlj = np.where(codes == 1, 1.0, np.where(codes == 2, 1 - 2 * theta, 0.0))
        with np.errstate(invalid="ignore", divide="ignore"):
            lj = np.where(codes == 1, u64, np.where(codes == 2, np.log(theta) + np.log1p(-theta), 0.0)).sum()
        want_lp, want_g = lp64 + lj, g * jac + dlj
        lp, gr = model.logp_grad(torch.from_numpy(np.tile(u, (3, 1))).cuda())
        assert abs(float(lp[1]) - want_lp) <= 2e-5 * max(1.0, abs(want_lp)), (name, float(lp[1]), want_lp)
        assert np.max(np.abs(gr[2].cpu().numpy() - want_g)) <= 2e-5 * max(1.0, np.max(np.abs(want_g))), (name, gr[2], want_g)


def test_transforms_are_off_by_default_and_draws_come_back_constrained(cuda):
    fn, init, meta = W.c2_event_rate(B.ns)
    kw = dict(num_samples=150, num_warmup=150, num_chains=512, key=mx.random.key(4))
    a, _ = B.hmc(fn, init, **kw)
    b, _ = B.hmc(fn, init, transforms=None, **kw)
    assert np.array_equal(a["rate"], b["rate"])                       # default = the reference's coordinates, bit for bit
    c, _ = B.hmc(fn, init, transforms="auto", adapt="dual_averaging", **kw)
    assert (c["rate"] > 0).all() and not np.array_equal(a["rate"], c["rate"])
    mean, sd = meta.post_shape / meta.post_rate, math.sqrt(meta.post_shape) / meta.post_rate
    assert mcse_ok(c["rate"], mean, sd)
    d, _ = B.nuts(fn, init, transforms="auto", compat="correct", **kw)
    assert (d["rate"] > 0).all() and mcse_ok(d["rate"], mean, sd)
    with pytest.raises(ValueError):
        B.nuts(fn, {"rate": -1.0}, transforms="auto", num_chains=2)   # the starting point must lie in the support


def _quadrature_posterior_c1(n):
    """float64 posterior moments of (mu, sigma) for the Normal(mu, sigma) model over n observations, by brute-force
    quadrature on a fine grid (priors Normal(0, 10), HalfNormal(5); tests/test_nuts.py:188-227 of the reference)."""
    y = W.data_c1(n).astype(np.float32).astype(np.float64)
    mu = np.linspace(y.mean() - 3.0, y.mean() + 3.0, 1201)
    sg = np.linspace(0.5, 6.0, 1401)
    M, S = np.meshgrid(mu, sg, indexing="ij")
    ss = ((y[None, None, :] - M[..., None]) ** 2).sum(-1)
    lp = -n * np.log(S) - 0.5 * ss / S ** 2 - 0.5 * (M / 10.0) ** 2 - 0.5 * (S / 5.0) ** 2
    w = np.exp(lp - lp.max())
    w /= w.sum()
    m_mu, m_sg = (w * M).sum(), (w * S).sum()
    return (m_mu, math.sqrt((w * (M - m_mu) ** 2).sum())), (m_sg, math.sqrt((w * (S - m_sg) ** 2).sum()))


def test_nuts_with_transforms_samples_the_reference_inference_problem(cuda):
    """The reference's own NUTS test model (tests/test_nuts.py:188-227: Normal(mu, sigma), 50 observations, the
    mx.array([...]) idiom) freezes the reference's sampler when a trajectory crosses sigma < 0.  In log
    coordinates for sigma, 1024 chains match the quadrature posterior within 4 MC standard errors."""
    fn, init, _ = W.c1_normal(B.ns, n=50, style="stack")
    (m_mu, s_mu), (m_sg, s_sg) = _quadrature_posterior_c1(50)
    s, rate, info = B.nuts(fn, init, num_samples=300, num_warmup=300, step_size=0.1, num_chains=1024, compat="correct",
                           transforms="auto", adapt_mass_matrix=True, key=mx.random.key(8), return_info=True)
    assert (s["sigma"] > 0).all() and np.isfinite(s["mu"]).all()
    assert mcse_ok(s["mu"], m_mu, s_mu), (s["mu"].mean(), m_mu, s["mu"].std(), s_mu)
    assert mcse_ok(s["sigma"], m_sg, s_sg), (s["sigma"].mean(), m_sg, s["sigma"].std(), s_sg)
    assert 0.5 < float(np.median(info.step_size)) < 3.0          # unit-scale coordinates after the metric is adapted
    assert info.depths.mean() < 3.5


# ------------------------------------------------------------------------------------------ mass matrix
def test_unit_mass_matrix_is_bit_identical_to_none_pointwise(cuda):
    """inv_mass = ones takes the mass-matrix code path of the pointwise kernels with every factor equal to 1.0f: same
    draws bit for bit as the reference's identity mass (the parity fixtures therefore cover the new arithmetic too)."""
    fn, init, _ = W.c1_normal(B.ns)
    model = compile_model(fn, init)
    ones = torch.ones(model.D, device="cuda")
    outs = []
    for im in (None, ones):
        st = ChainState(model, model.pack(init, 300), 0.02)
        dr = torch.empty(40, 300, model.D, device="cuda")
        launch_hmc(st, 40, 7, _cabi.ADAPT_NONE, 0.8, 5, 0, draws=dr, inv_mass=im)
        st2 = ChainState(model, model.pack(init, 300), 0.02)
        st2.da_state[:, 1] = 1.0
        dn = torch.empty(30, 300, model.D, device="cuda")
        launch_nuts(st2, 30, 6, _cabi.ADAPT_NONE, _cabi.COMPAT_CORRECT, 0.65, 5, 0, draws=dn, inv_mass=im)
        outs.append((dr, dn))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_mass_from_draws_matches_numpy(cuda):
    rng = np.random.default_rng(0)
    S, C, D = 37, 190, 70
    scale = 10.0 ** rng.uniform(-3, 3, D)
    x = (rng.standard_normal((S, C, D)) * scale + 5 * scale).astype(np.float32)
    model = compile_model(*W.c2_event_rate(B.ns)[:2])
    got = mass_from_window(model, torch.from_numpy(x).cuda()).cpu().numpy()
    n = S * C
    want = x.reshape(-1, D).astype(np.float64).var(axis=0) * (n / (n + 5.0)) + 1e-3 * 5.0 / (n + 5.0)
    assert np.max(np.abs(got - want) / want) < 1e-5


def test_mass_matrix_adaptation_on_a_badly_scaled_regression(cuda):
    """Coefficients on very different scales (column scales 0.03 ... 30 => posterior sds 1000:1): with the identity
    mass matrix NUTS needs a step size set by the narrowest direction and deep trees; the adapted diagonal metric
    recovers the posterior variances (within sampling error) and samples with shallow trees.  Both against the closed
    form N(m, V)."""
    n, d, C = 4000, 24, 512
    rng = np.random.default_rng(5)
    col = (10.0 ** np.linspace(-1.5, 1.5, d)).astype(np.float32)
    X = (rng.standard_normal((n, d)).astype(np.float32) * col)
    beta = (rng.standard_normal(d) / col).astype(np.float32)
    y = (X.astype(np.float64) @ beta + rng.standard_normal(n)).astype(np.float32)
    Xa, ya = mx.array(X), mx.array(y)

    def log_prob(p):
        return mx.sum(B.Normal(0, 100.0).log_prob(p["beta"])) + mx.sum(B.Normal(Xa @ p["beta"], 1.0).log_prob(ya))

    A = X.astype(np.float64).T @ X.astype(np.float64) + np.eye(d) / 100.0 ** 2
    V = np.linalg.inv(A)
    m = V @ (X.astype(np.float64).T @ y.astype(np.float64))
    sd = np.sqrt(np.diag(V))
    init = {"beta": np.zeros(d, dtype=np.float32)}
    # the chains start 60 posterior sds from the mode in EVERY scaled coordinate: the early windows see the transient
    # (drift inflates the variance of the slow coordinates, which is what speeds them up); 75 | 25-50-100-400 | 50
    kw = dict(num_samples=150, num_warmup=700, step_size=1e-3, max_tree_depth=10, num_chains=C, compat="correct",
              step_size_adaptation="pooled", key=mx.random.key(3), return_info=True)
    s, _, info = B.nuts(log_prob, init, adapt_mass_matrix=True, **kw)
    ratio = info.inv_mass / np.diag(V)
    print("inv_mass / posterior variance:", np.round(ratio, 2), "warm-up depth per window:",
          [round(float(info.warmup_depths[a:b].mean()), 2) for a, b in ((0, 75), (75, 100), (100, 150), (150, 250), (250, 650), (650, 700))])
    assert 0.6 < ratio.min() and ratio.max() < 1.6, (ratio.min(), ratio.max())
    for j in range(d):
        assert mcse_ok(s["beta"][:, :, j], m[j], sd[j], k=4.5), j
    assert info.depths.mean() <= 4.5, info.depths.mean()
    s0, _, info0 = B.nuts(log_prob, init, **{**kw, "num_samples": 20, "num_warmup": 100})
    assert info0.depths.mean() > info.depths.mean() + 2.0, (info0.depths.mean(), info.depths.mean())


def test_hmc_mass_matrix_on_device_matches_exact_posteriors(cuda):
    """fixed-L HMC with the adapted metric and transforms on the A/B model (two Beta posteriors, logit coordinates)."""
    fn, init, meta = W.c5_ab_test(B.ns)
    s, rate, info = B.hmc(fn, init, num_samples=300, num_warmup=300, step_size=0.05, num_leapfrog_steps=8, num_chains=2048,
                          adapt="dual_averaging", adapt_mass_matrix=True, transforms="auto", key=mx.random.key(5),
                          return_info=True)
    for name, (a, b) in (("p_A", meta.post_a), ("p_B", meta.post_b)):
        a, b = a + 0.0, b + 0.0      # posterior of Beta(1,1) prior x Beta(a, b) "likelihood" term: Beta(a, b) itself
        mean = a / (a + b)
        sd = math.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)))
        assert (s[name] > 0).all() and (s[name] < 1).all()
        assert mcse_ok(s[name], mean, sd), (name, s[name].mean(), mean, s[name].std(), sd)
    assert 0.6 < rate <= 1.0 and info.inv_mass.shape == (2,)


# ------------------------------------------------------------------------------------------ quantiles
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 300001])
def test_device_quantiles_match_numpy(cuda, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32) * 3 - 1
    if n > 10:
        x[::7] = x[0]                 # ties
        x[1::11] = -0.0
        x[3::13] = 0.0
    q = [0.0, 0.025, 0.5, 0.6180339, 0.975, 1.0]
    got = device_quantiles(torch.from_numpy(x).cuda(), q)
    want = np.percentile(x.astype(np.float64), [100 * v for v in q])
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7), (got, want)


# ------------------------------------------------------------------------------------------ tensor-core NUTS decisions
@pytest.mark.parametrize("copies", [1, 256])
def test_nuts_tree_decisions_match_reference_at_mid_size(cuda, copies):
    """64 coefficients x 2048 observations: every NUTS transition of the reference run replayed on the tcgen05 path
    (single-CTA tiles for one chain, CTA pairs for 256 identical chains) with the reference's draws injected --
    direction, n', s', candidate taken, U-turn and depth must be the reference's."""
    from test_gpu_parity import _replay_nuts
    _replay_nuts("nuts_regression_mid", 1, copies=copies, draw_tol=1e-4, alpha_tol=2e-2)


# ------------------------------------------------------------------------------------------ parity check 3 at the Target size
@pytest.mark.parametrize("mode", ["jitter", "mass_matrix"])
def test_c4_size_nuts_from_zero_matches_closed_form_posterior(cuda, mode):
    """BASELINE.json configs[3] through the user-facing call: 1000 coefficients x 100,000 observations, 4096 lock-step
    chains from beta = 0.  Posterior mean of every coefficient within 0.05 posterior sd of the closed-form N(m, V), pooled
    sd within 3 %, R-hat ~ 1 (the numbers tools/c4_full_run.py prints), and trees that stay shallow:
      jitter       60 warm-up + 40 kept transitions, identity mass, step_size_jitter = 0.2 (round 1's work-around for the
                   U-turn resonance at the step size the reference's dual averaging lands on);
      mass_matrix  150 + 40 with the windowed diagonal mass-matrix adaptation and NO jitter."""
    n, d, C = 100000, 1000, 4096
    fn, init, meta = W.regression(B.ns, n, d, seed=0)
    m = B.MCMC(fn)
    kw = dict(method="nuts", step_size=1e-3, num_chains=C, compat="correct", step_size_adaptation="pooled", return_torch=True,
              return_info=True, verbose=False, random_seed=1)
    if mode == "jitter":
        m.run(init, num_samples=40, num_warmup=60, step_size_jitter=0.2, **kw)
    else:
        m.run(init, num_samples=40, num_warmup=150, adapt_mass_matrix=True, **kw)
    diag = m.diagnostics(ess=False)["beta"]
    X = torch.from_numpy(meta.X).cuda().double()
    A = (X.T @ X + torch.eye(d, device="cuda", dtype=torch.float64) / meta.prior_scale ** 2)
    V = torch.linalg.inv(A)
    post_m = (V @ (X.T @ torch.from_numpy(meta.y).cuda().double())).cpu().numpy()
    sd = torch.sqrt(torch.diag(V)).cpu().numpy()
    err = np.abs(diag["mean"] - post_m) / sd
    ratio = diag["std"] / sd
    assert err.max() < 0.05, err.max()
    assert 0.97 < ratio.min() and ratio.max() < 1.03, (ratio.min(), ratio.max())
    assert np.nanmax(diag["rhat"]) < 1.05
    assert m.info.depths.mean() <= 5.0, m.info.depths.mean()
    if mode == "mass_matrix":      # the metric is the posterior variance (~1/N per coefficient); unit-scale step size
        assert 0.5 < np.median(m.info.inv_mass) * n < 2.0 and 0.1 < float(m.info.step_size[0]) < 1.0
