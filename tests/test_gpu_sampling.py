"""GPU: the public API end to end -- shapes, reproducibility, shard-independence of the random streams,
and parity check 3 (posterior moments within 4 Monte-Carlo standard errors of closed forms / the oracle)."""
import math
import os

import numpy as np
import pytest
import torch

import mlx_mcmc_b200 as B
import mlx_mcmc_b200.core as mx
from mlx_mcmc_b200 import workloads as W
from mlx_mcmc_b200.diagnostics import compute_ess, ess_geyer

pytestmark = pytest.mark.gpu


def mcse_ok(x, mean, sd, k=4.0):
    """Parity check 3 for draws x[chain, draw] against a known (mean, sd): the pooled mean and the pooled
    second central moment must be within k Monte-Carlo standard errors, the standard errors being
    estimated from the spread of the per-chain statistics (chains are independent, so this is valid
    whatever the within-chain autocorrelation -- including chains the reference's +-5 % rule left stuck)."""
    x = np.asarray(x, dtype=np.float64)
    C = x.shape[0]
    m_c = x.mean(axis=1)
    v_c = ((x - mean) ** 2).mean(axis=1)
    ok_mean = abs(m_c.mean() - mean) <= k * m_c.std(ddof=1) / math.sqrt(C)
    ok_var = abs(v_c.mean() - sd ** 2) <= k * v_c.std(ddof=1) / math.sqrt(C)
    return bool(ok_mean and ok_var)


# ------------------------------------------------------------------------------ API contract
def test_single_chain_shapes_and_types(cuda):
    fn, init, _ = W.c1_normal(B.ns)
    m = B.MCMC(fn)
    s = m.run(init, num_samples=200, num_warmup=100, method="hmc", step_size=0.01, num_leapfrog_steps=5, verbose=False)
    assert set(s) == {"mu", "sigma"} and s["mu"].shape == (200,) and isinstance(s["mu"], np.ndarray)
    assert isinstance(m.acceptance_rate, float) and 0.0 <= m.acceptance_rate <= 1.0
    summ = m.summary()
    assert set(summ["mu"]) == {"mean", "std", "median", "2.5%", "97.5%"}
    m.print_summary()
    s = m.run(init, num_samples=150, num_warmup=50, proposal_scale=0.3, verbose=False)      # default: metropolis
    assert s["sigma"].shape == (150,)
    fnv, initv, _ = W.t_vector_normal(B.ns, d=3)
    sv, rate = B.nuts(fnv, initv, num_samples=100, num_warmup=100, step_size=0.3, key=mx.random.key(1))
    assert sv["x"].shape == (100, 3)                                                           # nuts.py:338-339
    with pytest.raises(TypeError):
        m.run(init, num_samples=10, num_warmup=10, method="metropolis", step_size=0.1, verbose=False)


def test_multi_chain_shapes_and_torch_return(cuda):
    fn, init, _ = W.c2_event_rate(B.ns)
    s, rate = B.hmc(fn, init, num_samples=50, num_warmup=50, num_chains=64)
    assert s["rate"].shape == (64, 50)
    s, rate = B.hmc(fn, init, num_samples=50, num_warmup=50, num_chains=64, return_torch=True)
    assert isinstance(s["rate"], torch.Tensor) and s["rate"].is_cuda and tuple(s["rate"].shape) == (64, 50)


def test_same_key_same_draws_and_different_key_differs(cuda):
    fn, init, _ = W.t_normal_1d(B.ns, 0.0, 1.0)
    a, _ = B.nuts(fn, init, num_samples=100, num_warmup=100, key=mx.random.key(42))
    b, _ = B.nuts(fn, init, num_samples=100, num_warmup=100, key=mx.random.key(42))
    c, _ = B.nuts(fn, init, num_samples=100, num_warmup=100, key=mx.random.key(43))
    np.testing.assert_array_equal(a["mu"], b["mu"])          # tests/test_nuts.py:110-136, exact here
    assert not np.array_equal(a["mu"], c["mu"])
    a, _ = B.hmc(fn, init, num_samples=100, num_warmup=100, key=mx.random.key(7))
    b, _ = B.hmc(fn, init, num_samples=100, num_warmup=100, key=mx.random.key(7))
    np.testing.assert_array_equal(a["mu"], b["mu"])          # tests/test_hmc.py:148-177


@pytest.mark.parametrize("method", ["hmc", "nuts", "metropolis"])
def test_draws_do_not_depend_on_sharding_or_lanes(cuda, method):
    """chains [0,32) in one call == chains [0,16) + [16,32) in two calls == any lanes-per-chain choice"""
    fn, init, _ = W.c2_event_rate(B.ns)

    def run(nc, off, lanes):
        if method == "hmc":
            return B.hmc(fn, init, num_samples=40, num_warmup=40, num_chains=nc, chain_offset=off, lanes=lanes,
                         key=mx.random.key(3))[0]["rate"]
        if method == "nuts":
            return B.nuts(fn, init, num_samples=40, num_warmup=40, num_chains=nc, chain_offset=off, lanes=lanes,
                          max_tree_depth=6, key=mx.random.key(3))[0]["rate"]
        return B.metropolis_hastings(fn, init, num_samples=80, proposal_scale=0.3, num_chains=nc, chain_offset=off,
                                     lanes=lanes, random_seed=3)[0]["rate"]

    whole = run(32, 0, 1)
    halves = np.concatenate([run(16, 0, 1), run(16, 16, 1)])
    np.testing.assert_array_equal(whole, halves)
    wide = run(32, 0, 8)
    if method == "nuts":
        # lane-split sums differ in the last bits; a U-turn test sitting on a rounding boundary may then
        # resolve differently, after which that chain's path legitimately differs -- most chains agree
        same = np.mean(np.all(np.isclose(whole, wide, rtol=2e-3), axis=1))
        assert same > 0.5 and abs(np.mean(whole) - np.mean(wide)) < 0.1
    else:
        np.testing.assert_allclose(whole, wide, rtol=2e-3)


# ------------------------------------------------------------------------------ parity 3
def test_c2_hmc_posterior_matches_exact_gamma(cuda):
    fn, init, meta = W.c2_event_rate(B.ns)
    s, rate = B.hmc(fn, init, num_samples=1000, num_warmup=500, num_chains=4096, key=mx.random.key(0))
    x = s["rate"]
    mean, sd = meta.post_shape / meta.post_rate, math.sqrt(meta.post_shape) / meta.post_rate
    assert 0.5 < rate <= 1.0
    assert mcse_ok(x, mean, sd), (np.mean(x), mean, np.std(x), sd)


def test_c2_hmc_dual_averaging_hits_target(cuda):
    fn, init, meta = W.c2_event_rate(B.ns)
    s, rate, info = B.hmc(fn, init, num_samples=500, num_warmup=500, num_chains=2048, adapt="dual_averaging",
                          target_accept=0.8, return_info=True, key=mx.random.key(1))
    assert 0.7 < rate < 0.95, rate
    mean, sd = meta.post_shape / meta.post_rate, math.sqrt(meta.post_shape) / meta.post_rate
    assert abs(np.mean(s["rate"]) - mean) < 0.02 and abs(np.std(s["rate"]) - sd) < 0.02


def test_c5_metropolis_posterior_matches_exact_beta(cuda):
    fn, init, meta = W.c5_ab_test(B.ns)
    m = B.MCMC(fn)
    s = m.run(init, num_samples=2000, num_warmup=1000, proposal_scale=0.02, num_chains=8192, random_seed=0, verbose=False)
    assert 0.2 < m.acceptance_rate < 0.5          # the reference reports ~0.33
    for name, (a, b) in (("p_A", meta.post_a), ("p_B", meta.post_b)):
        mean, sd = a / (a + b), math.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)))
        x = s[name]
        assert mcse_ok(x, mean, sd), (name, np.mean(x), mean, np.std(x), sd)
        assert np.all((x > 0) & (x < 1))


def test_c1_metropolis_agrees_with_oracle_run(cuda):
    """config-1 model, reference's example-01/02 Metropolis settings: pooled CUDA chains vs one long oracle chain"""
    from oracle.ns import ns as ons, samplers
    fo, init, _ = W.c1_normal(ons)
    so, _ = samplers.run_port(fo, init, num_samples=3000, num_warmup=1000, proposal_scale=0.3, random_seed=42)
    fn, _, _ = W.c1_normal(B.ns)
    m = B.MCMC(fn)
    s = m.run(init, num_samples=1000, num_warmup=1000, proposal_scale=0.3, num_chains=1024, random_seed=42, verbose=False)
    for name in ("mu", "sigma"):
        o = np.asarray(so[name])
        ess_o = max(ess_geyer(o), 10.0)
        mcse = np.std(o) / math.sqrt(ess_o)
        assert abs(np.mean(s[name]) - np.mean(o)) <= 4 * mcse, (name, np.mean(s[name]), np.mean(o), mcse)
        assert abs(np.std(s[name]) - np.std(o)) <= 4 * np.std(o) / math.sqrt(2 * ess_o)


def test_nuts_prior_models_match_truth(cuda):
    """the reference's own NUTS tests (tests/test_nuts.py:13-54,88-108) over many chains"""
    fn, init, _ = W.t_normal_1d(B.ns, 5.0, 2.0)
    s, rate = B.nuts(fn, init, num_samples=500, num_warmup=500, step_size=0.5, num_chains=1024, key=mx.random.key(42))
    x = s["mu"]
    assert abs(np.mean(x) - 5.0) < 0.05 and abs(np.std(x) - 2.0) < 0.05 and rate > 0.5
    fn, init, _ = W.t_normal_2d(B.ns)
    s, _ = B.nuts(fn, init, num_samples=500, num_warmup=500, step_size=0.3, num_chains=1024, key=mx.random.key(123))
    assert abs(np.mean(s["mu1"])) < 0.05 and abs(np.mean(s["mu2"]) - 5.0) < 0.08
    assert abs(np.std(s["mu1"]) - 1.0) < 0.05 and abs(np.std(s["mu2"]) - 2.0) < 0.08
    fn, init, _ = W.t_halfnormal_scale(B.ns)
    s, _ = B.nuts(fn, init, num_samples=300, num_warmup=300, step_size=0.1, num_chains=256, key=mx.random.key(456))
    assert np.all(s["sigma"] > 0)


def test_nuts_correct_mode_c2_matches_exact_gamma(cuda):
    """compat='correct' (log-space slice) on a data model; the exact posterior is Gamma(52, 15.1)"""
    fn, init, meta = W.c2_event_rate(B.ns)
    s, rate, info = B.nuts(fn, init, num_samples=500, num_warmup=500, num_chains=2048, compat="correct",
                           return_info=True, key=mx.random.key(5))
    mean, sd = meta.post_shape / meta.post_rate, math.sqrt(meta.post_shape) / meta.post_rate
    x = s["rate"]
    assert abs(np.mean(x) - mean) < 0.02 and abs(np.std(x) - sd) < 0.02, (np.mean(x), mean, np.std(x), sd)
    assert info.grad_evals > 0 and info.depths.max() <= 10


def test_hmc_constrained_parameter_stays_positive(cuda):
    """tests/test_hmc.py:118-146 across chains: HalfNormal(2) draws are all positive"""
    fn, init, _ = W.t_halfnormal(B.ns, 2.0)
    s, rate = B.hmc(fn, init, num_samples=500, num_warmup=300, step_size=0.1, num_leapfrog_steps=10, num_chains=512)
    assert np.all(s["sigma"] > 0)
    assert 0.5 < np.mean(s["sigma"]) < 3.0


def test_ess_helpers(cuda):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4000)
    assert 2500 < compute_ess(x) <= 4000 * 1.2 and 2500 < ess_geyer(x) < 5000
    ar = np.zeros(4000)
    for i in range(1, 4000):
        ar[i] = 0.9 * ar[i - 1] + rng.standard_normal()
    assert 100 < ess_geyer(ar) < 500                          # n (1-rho)/(1+rho) = 210


# ------------------------------------------------------------------------------ GLM class (X @ beta)
def test_regression_logp_grad_medium_vs_float64(cuda):
    """README 'Medium' shape scaled down: N=2000, D=96, 300 chains at random positions vs float64 numpy"""
    from mlx_mcmc_b200.engine import compile_model
    fn, init, meta = W.regression(B.ns, 2000, 96, seed=1)
    model = compile_model(fn, init, cache=False)
    assert model.model_class == 1
    rng = np.random.default_rng(0)
    theta = (meta.beta_true[None, :] + 0.3 * rng.standard_normal((300, 96))).astype(np.float32)
    lp, gr = model.logp_grad(torch.from_numpy(theta).cuda())
    X, y, b = meta.X.astype(np.float64), meta.y.astype(np.float64), theta.astype(np.float64)
    r = y[None, :] - b @ X.T
    lp64 = (-0.5 * (r ** 2).sum(1) - 2000 * 0.5 * math.log(2 * math.pi)
            - 0.5 * (b ** 2).sum(1) / 100.0 - 96 * (0.5 * math.log(2 * math.pi) + math.log(10.0)))
    g64 = r @ X - b / 100.0
    assert np.max(np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64)) < 1e-5
    assert np.max(np.abs(gr.cpu().numpy() - g64)) / np.max(np.abs(g64)) < 1e-5


def test_regression_nuts_matches_closed_form_posterior(cuda):
    """parity check 3 on the regression model: compat='correct' NUTS vs the closed-form N(m, V)"""
    fn, init, meta = W.regression(B.ns, 500, 8, seed=2)
    m, V = W.regression_posterior(meta)
    s, rate, info = B.nuts(fn, init, num_samples=300, num_warmup=300, step_size=0.05, num_chains=512, compat="correct",
                           return_info=True, key=mx.random.key(4))
    x = s["beta"]                                   # (C, S, D)
    assert x.shape == (512, 300, 8)
    sd = np.sqrt(np.diag(V))
    for d in range(8):
        assert mcse_ok(x[:, :, d], m[d], sd[d]), (d, x[:, :, d].mean(), m[d], x[:, :, d].std(), sd[d])
    assert info.grad_evals > 0


def test_regression_nuts_pooled_step_size_and_compaction(cuda):
    """Pooled dual averaging: every chain ends warm-up with the same step size, the posterior still matches the
    closed form; and a chain's draws do not depend on which other chains share its (compacted) lock-step batch."""
    fn, init, meta = W.regression(B.ns, 800, 12, seed=7)
    m, V = W.regression_posterior(meta)
    s, rate, info = B.nuts(fn, init, num_samples=200, num_warmup=200, step_size=0.02, num_chains=384, compat="correct",
                           step_size_adaptation="pooled", return_info=True, key=mx.random.key(8))
    assert np.all(info.step_size == info.step_size[0]) and 1e-3 < info.step_size[0] < 1.0
    sd = np.sqrt(np.diag(V))
    for d in range(12):
        assert mcse_ok(s["beta"][:, :, d], m[d], sd[d]), d
    depths = info.depths
    assert depths.min() < depths.max()          # chains do stop at different depths, so batches were compacted
    # fixed step size, no adaptation: chains 0..127 alone vs inside a 384-chain batch.  A GLM-class chain's
    # arithmetic depends on its batch only through the centring point and the split-K order (float32 rounding),
    # so the first draws agree to rounding (later ones drift apart: the trajectories are chaotic)
    kw = dict(num_samples=2, num_warmup=1, step_size=float(info.step_size[0]), adapt_step_size=False, compat="correct",
              key=mx.random.key(9))
    a, _ = B.nuts(fn, init, num_chains=128, **kw)
    b, _ = B.nuts(fn, init, num_chains=384, **kw)
    close = np.all(np.abs(a["beta"][:, 0] - b["beta"][:128, 0]) < 1e-4 * (1 + np.abs(a["beta"][:, 0])), axis=1)
    assert close.mean() > 0.95, close.mean()
    with pytest.raises(ValueError):
        B.nuts(fn, init, num_chains=4, step_size_adaptation="global")


def test_regression_hmc_and_metropolis_run(cuda):
    fn, init, meta = W.regression(B.ns, 300, 4, seed=6)
    m, V = W.regression_posterior(meta)
    s, rate = B.hmc(fn, init, num_samples=400, num_warmup=300, step_size=0.02, num_leapfrog_steps=8, num_chains=256,
                    key=mx.random.key(2))
    assert s["beta"].shape == (256, 400, 4) and rate > 0.5
    sd = np.sqrt(np.diag(V))
    for d in range(4):
        assert mcse_ok(s["beta"][:, :, d], m[d], sd[d]), d
    s, rate = B.metropolis_hastings(fn, init, num_samples=300, proposal_scale=0.03, num_chains=64, random_seed=1)
    assert s["beta"].shape == (64, 300, 4) and 0.05 < rate < 0.95


def test_multi_gpu_sharding_when_two_gpus_are_visible(cuda):
    """tests/mgpu_check.py under torchrun on 2 GPUs: chain sharding is bit-equal to one GPU, observation
    sharding agrees to float32 rounding and keeps the ranks in lock-step.  Skipped on a 1-GPU box."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(root, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_device_diagnostics_match_the_oracle(cuda):
    """SURVEY.md 8(f) rows 1-2 on the device: per-series mean / var / ESS (both reference estimators, Geyer), pooled
    mean / std, R-hat -- against the oracle restatement of the reference's compute_ess / summary on the fixture series
    and on random multi-chain draws."""
    from oracle.refport import diagnostics as OD
    from mlx_mcmc_b200.diagnostics import device_series_stats, device_summary
    from util import golden
    g = golden("diagnostics")
    for n in (8, 57, 400, 1500):                                  # ragged lengths incl. the shortest the estimator takes
        rows = [r for r in g["ess"] if r["n"] == n]
        x = np.stack([np.asarray(r["x"], dtype=np.float32) for r in rows], axis=1)       # [S, series]
        d = torch.from_numpy(x[:, :, None].copy()).cuda()                                 # [S, C = series, D = 1]
        for mode, key in ((0, "ess06"), (1, "ess02")):
            st = device_series_stats(d, ess=True, ess_mode=mode)
            got = st["ess"][:, 0].cpu().numpy()
            for j, r in enumerate(rows):
                if r[key] is None:
                    continue
                assert abs(got[j] - r[key]) <= 2e-3 * abs(r[key]) + 1e-3, (n, r["kind"], mode, got[j], r[key])
        assert np.allclose(st["mean"][:, 0].cpu().numpy(), x.mean(0), rtol=1e-6, atol=1e-6)
        assert np.allclose(st["var"][:, 0].cpu().numpy(), x.astype(np.float64).var(0), rtol=1e-5, atol=1e-7)
        gey = st["ess_geyer"][:, 0].cpu().numpy()
        for j in range(x.shape[1]):
            want = ess_geyer(x[:, j])
            assert abs(gey[j] - want) <= 1e-3 * abs(want) + 1e-3, (n, rows[j]["kind"], gey[j], want)
    rng = np.random.default_rng(0)
    S, C, D = 300, 37, 5
    x = (rng.standard_normal((S, C, D)) * np.arange(1, D + 1) + 0.3 * rng.standard_normal((1, C, 1))).astype(np.float32)
    tab = device_summary(torch.from_numpy(x).cuda())
    for dd in range(D):
        assert abs(tab["mean"][dd] - x[:, :, dd].mean()) < 1e-5
        assert abs(tab["std"][dd] - x[:, :, dd].astype(np.float64).std()) < 1e-5 * (dd + 1)
        assert abs(tab["rhat"][dd] - OD.rhat(x[:, :, dd].T)) < 1e-5
        want = sum(float(OD.compute_ess(x[:, c, dd])) for c in range(C))
        assert abs(tab["ess"][dd] - want) <= 2e-3 * want
    # through the API: run(..., return_torch=True) then diagnostics()
    fn, init, meta = W.c2_event_rate(B.ns)
    m = B.MCMC(fn)
    m.run(init, num_samples=400, num_warmup=300, method="hmc", num_chains=64, return_torch=True, verbose=False,
          adapt="dual_averaging")
    t = m.diagnostics()["rate"]
    assert abs(t["mean"] - meta.post_shape / meta.post_rate) < 0.05 and 0.99 < t["rhat"] < 1.05 and t["ess_geyer"] > 64 * 50, t


def test_device_forward_sampling_moments_and_quantiles(cuda):
    """SURVEY.md 8(f) row 4.  The reference's own sampling tests are 10k-draw moment checks (tests/test_distributions.py:40-50,
    87-102; tests/test_new_distributions.py:46-59,107-120,162-175,235-256); here 400k device draws per distribution are held
    to the exact moments and to the exact quantiles (scipy) much more tightly, plus reproducibility and support."""
    from scipy import stats
    n = 400_000
    cases = [
        (B.Normal(1.5, 2.0), stats.norm(1.5, 2.0)),
        (B.HalfNormal(2.0), stats.halfnorm(scale=2.0)),
        (B.Exponential(3.0), stats.expon(scale=1 / 3.0)),
        (B.Gamma(2.0, 1.5), stats.gamma(2.0, scale=1 / 1.5)),
        (B.Gamma(0.4, 2.0), stats.gamma(0.4, scale=1 / 2.0)),            # shape < 1: the boosted branch
        (B.Gamma(50.0, 5.0), stats.gamma(50.0, scale=1 / 5.0)),
        (B.Beta(2.0, 5.0), stats.beta(2.0, 5.0)),
        (B.Beta(0.5, 0.5), stats.beta(0.5, 0.5)),
        (B.Beta(116.0, 886.0), stats.beta(116.0, 886.0)),                 # the A/B example's posterior
    ]
    qs = np.array([0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99])
    for dist, exact in cases:
        x = dist.sample_device(mx.random.key(11), (n,))
        assert x.is_cuda and x.dtype == torch.float32 and x.shape == (n,)
        h = x.cpu().numpy().astype(np.float64)
        assert np.isfinite(h).all() and h.min() >= exact.support()[0]
        se = exact.std() / math.sqrt(n)
        assert abs(h.mean() - exact.mean()) < 5 * se, (dist, h.mean(), exact.mean())
        assert abs(h.var() - exact.var()) < 0.02 * exact.var() + 1e-9, (dist, h.var(), exact.var())
        emp = np.quantile(h, qs)
        # quantile standard error ~ sqrt(q(1-q)/n) / pdf
        tol = 6 * np.sqrt(qs * (1 - qs) / n) / np.maximum(exact.pdf(exact.ppf(qs)), 1e-12)
        assert np.all(np.abs(emp - exact.ppf(qs)) < tol + 1e-6), (dist, emp, exact.ppf(qs))
        y = dist.sample_device(mx.random.key(11), (n,))
        z = dist.sample_device(mx.random.key(12), (n,))
        assert torch.equal(x, y) and not torch.equal(x, z)
    c = B.Categorical(probs=[0.2, 0.5, 0.3])
    k = c.sample_device(mx.random.key(5), (200, 1000))
    assert k.dtype == torch.int64 and k.shape == (200, 1000) and int(k.min()) == 0 and int(k.max()) == 2
    freq = torch.bincount(k.flatten(), minlength=3).double().cpu().numpy() / k.numel()
    assert np.allclose(freq, [0.2, 0.5, 0.3], atol=0.005)
    assert B.Normal(0, 1).sample_device(mx.random.key(1)).shape == ()
    with pytest.raises(TypeError):                       # a traced parameter cannot be sampled from
        from mlx_mcmc_b200.tracer import trace
        trace(lambda p: B.Normal(p["m"], 1.0).sample_device(mx.random.key(0), (3,)), {"m": 0.0})


def test_summary_on_device_draws_matches_numpy(cuda):
    """MCMC.summary() (mcmc.py:191-225) on device draws: same keys and numbers as numpy on the host copy."""
    fn, init, meta = W.c1_normal(B.ns)
    m = B.MCMC(fn)
    m.run(init, num_samples=300, num_warmup=300, method="hmc", step_size=0.05, num_chains=50, return_torch=True,
          verbose=False, adapt="dual_averaging")
    dev = m.summary(0.9)
    host = {k: v.cpu().numpy() for k, v in m.samples.items()}
    for name, x in host.items():
        want = {'mean': np.mean(x), 'std': np.std(x), 'median': np.median(x), '5.0%': np.percentile(x, 5.0),
                '95.0%': np.percentile(x, 95.0)}
        assert list(dev[name].keys()) == list(want.keys())
        for k, v in want.items():
            assert abs(dev[name][k] - v) <= 1e-4 * max(1.0, abs(v)), (name, k, dev[name][k], v)


def test_c3_size_nuts_matches_closed_form_posterior(cuda):
    """parity check 3 at the README 'Medium' size (100 coefficients x 10,000 observations, 1024 lock-step chains):
    posterior means and variances of every coefficient against the closed-form N(m, V) within 4 MC standard errors."""
    fn, init, meta = W.regression(B.ns, 10000, 100, seed=0)
    m, V = W.regression_posterior(meta)
    s, rate, info = B.nuts(fn, init, num_samples=120, num_warmup=150, step_size=0.005, max_tree_depth=8, num_chains=1024,
                           compat="correct", step_size_adaptation="pooled", return_info=True, return_torch=True,
                           key=mx.random.key(12))
    x = s["beta"].double()                                   # (C, S, D)
    C, S, D = x.shape
    sd = np.sqrt(np.diag(V))
    # chains are independent: the standard error of the grand mean comes from the spread of the chain means
    cm = x.mean(dim=1)                                       # (C, D)
    grand = cm.mean(dim=0).cpu().numpy()
    se = (cm.std(dim=0) / math.sqrt(C)).cpu().numpy()
    assert np.max(np.abs(grand - m) / se) < 4.5, np.max(np.abs(grand - m) / se)
    var_c = x.var(dim=1, unbiased=True)                      # per-chain variances, (C, D)
    v_hat = var_c.mean(dim=0).cpu().numpy() + cm.var(dim=0).cpu().numpy()
    v_se = (var_c.std(dim=0) / math.sqrt(C)).cpu().numpy() + 1e-12
    assert np.max(np.abs(v_hat - sd ** 2) / (4.5 * v_se + 0.02 * sd ** 2)) < 1.0
    assert np.all(info.step_size == info.step_size[0]) and info.grad_evals > 0


def test_step_size_jitter_keeps_the_posterior_and_is_off_by_default(cuda):
    """step_size_jitter (extension): 0 reproduces the un-jittered run bit for bit (pointwise class), a positive value
    changes the trajectories but not the target."""
    fn, init, meta = W.c2_event_rate(B.ns)
    kw = dict(num_samples=200, num_warmup=200, num_chains=256, compat="correct", key=mx.random.key(2))
    a, _ = B.nuts(fn, init, **kw)
    b, _ = B.nuts(fn, init, step_size_jitter=0.0, **kw)
    c, _ = B.nuts(fn, init, step_size_jitter=0.3, **kw)
    assert np.array_equal(a["rate"], b["rate"]) and not np.array_equal(a["rate"], c["rate"])
    mean, sd = meta.post_shape / meta.post_rate, math.sqrt(meta.post_shape) / meta.post_rate
    assert mcse_ok(c["rate"], mean, sd)
    with pytest.raises(ValueError):
        B.nuts(fn, init, step_size_jitter=1.5, num_chains=2)


def test_async_and_sync_nuts_schedules_agree(cuda):
    """The iteration-asynchronous schedule (default) and the synchronous one run the same per-chain algorithm with
    the same Philox slots: tree depths are identical and the first draws agree to the batch-dependent rounding of the
    contractions; both match the closed-form posterior."""
    fn, init, meta = W.regression(B.ns, 900, 10, seed=4)
    m, V = W.regression_posterior(meta)
    kw = dict(num_samples=60, num_warmup=80, step_size=0.02, num_chains=300, compat="correct", key=mx.random.key(6),
              return_info=True)
    a, ra, ia = B.nuts(fn, init, schedule="sync", **kw)
    b, rb, ib = B.nuts(fn, init, schedule="async", **kw)
    # per-chain dual averaging: every chain is an independent replica in both schedules
    same_depth = (ia.warmup_depths[:5] == ib.warmup_depths[:5]).mean()
    assert same_depth > 0.98, same_depth
    sd = np.sqrt(np.diag(V))
    for x in (a["beta"], b["beta"]):
        for d in range(10):
            assert mcse_ok(x[:, :, d], m[d], sd[d]), d
    assert abs(ra - rb) < 0.05 and ia.grad_evals > 0 and ib.grad_evals > 0
    # unequal work per chain: a few chains with a tiny step size run much deeper trees; the asynchronous schedule must
    # still deliver exactly num_samples draws per chain
    st_kw = dict(num_samples=12, num_warmup=1, adapt_step_size=False, step_size=0.02, num_chains=64, compat="correct",
                 max_tree_depth=7, key=mx.random.key(7), return_info=True)
    c, _, ic = B.nuts(fn, init, **st_kw)
    assert c["beta"].shape == (64, 12, 10) and np.isfinite(c["beta"]).all() and ic.depths.shape == (12, 64)
