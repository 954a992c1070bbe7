"""GPU: the CUDA path (through the C ABI) against the oracle and the committed golden fixtures.

Parity check 1: log p and gradient within 1e-5 relative (norm-wise) of the float64 arbiter and the
                reference's float32 values.
Parity check 2: with the reference's momentum / uniform draws injected, every accept decision and
                every NUTS tree decision is identical, states agree to float32 rounding.
"""
import numpy as np
import pytest
import torch

import mlx_mcmc_b200 as B
from mlx_mcmc_b200 import _cabi, workloads as W
from mlx_mcmc_b200.engine import ChainState, compile_model, launch_hmc, launch_mh, launch_nuts
from util import flat_params, golden, rel_err, tape_normals, tape_nuts, tape_uniforms

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north_star: "log_prob and gradient match within 1e-5 relative in fp32"


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype else t).cuda()


# ------------------------------------------------------------------------------ parity 1
@pytest.mark.parametrize("lanes", [1, 4, 32])
def test_logp_grad_vs_golden(cuda, lanes):
    g = golden("logp_grad")
    for model_name, rows in g.items():
        fn, init, _ = W.ALL_SMALL[model_name](B.ns)
        model = compile_model(fn, init)
        theta = np.stack([flat_params(model.layout, r["params"]) for r in rows])
        lp, gr = model.logp_grad(dev(theta), lanes=lanes)
        lp, gr = lp.cpu().numpy(), gr.cpu().numpy()
        for i, r in enumerate(rows):
            g64 = flat_params(model.layout, r["grad64"])
            g32 = flat_params(model.layout, r["grad32"])
            if np.isfinite(r["logp64"]):
                assert abs(lp[i] - r["logp64"]) <= TOL * max(abs(r["logp64"]), 1.0), (model_name, r["params"], lp[i])
                assert abs(lp[i] - r["logp32"]) <= TOL * max(abs(r["logp32"]), 1.0)
            elif np.isnan(r["logp64"]):
                assert np.isnan(lp[i]), (model_name, r["params"])
            else:
                assert lp[i] == r["logp64"], (model_name, r["params"], lp[i])       # -inf outside the support
            assert np.all(np.isfinite(gr[i]) == np.isfinite(g64)), (model_name, r["params"], gr[i], g64)
            ok = np.isfinite(g64)
            if ok.any():
                assert rel_err(gr[i][ok], g64[ok]) <= TOL, (model_name, r["params"], gr[i], g64)
                assert rel_err(gr[i][ok], g32[ok]) <= TOL


def test_logp_grad_random_points_vs_oracle(cuda):
    """seeded random points, CUDA vs the oracle's float64 evaluation"""
    from oracle.ns import ns as ons, value_and_grad
    rng = np.random.default_rng(0)
    for name in ("c1_normal", "c2_event_rate", "c5_ab_test", "t_vector_normal"):
        fn, init, _ = W.ALL_SMALL[name](B.ns)
        fo, _, _ = W.ALL_SMALL[name](ons)
        model = compile_model(fn, init)
        pts = []
        for _ in range(16):
            p = {}
            for k, v in init.items():
                base = np.asarray(v, dtype=np.float64)
                if name == "c5_ab_test":
                    p[k] = float(rng.uniform(0.02, 0.6))
                elif name == "c1_normal":
                    p[k] = float(rng.uniform(0.5, 8.0))
                elif name == "c2_event_rate":
                    p[k] = float(rng.uniform(0.2, 9.0))
                else:
                    p[k] = (base + rng.normal(size=base.shape)).tolist()
            pts.append(p)
        theta = np.stack([flat_params(model.layout, p) for p in pts])
        for lanes in (0, 1, 8):
            lp, gr = model.logp_grad(dev(theta), lanes=lanes)
            lp, gr = lp.cpu().numpy(), gr.cpu().numpy()
            G64 = []
            for i, p in enumerate(pts):
                lp64, g64 = value_and_grad(fo, p, "float64")
                assert abs(lp[i] - lp64) <= TOL * max(abs(lp64), 1.0), (name, p)
                G64.append(flat_params(model.layout, g64))
            assert rel_err(gr, np.stack(G64)) <= TOL, (name, lanes)       # norm-wise over the batch


def test_known_answers_through_cuda(cuda):
    """the reference's known-answer log_prob tests, evaluated by the CUDA kernels"""
    import math
    cases = [
        (lambda p: B.Normal(0, 1).log_prob(p["x"]), 0.0, -0.5 * math.log(2 * math.pi)),
        (lambda p: B.HalfNormal(1).log_prob(p["x"]), 0.0, math.log(2) - 0.5 * math.log(2 * math.pi)),
        (lambda p: B.HalfNormal(1).log_prob(p["x"]), -1.0, -math.inf),
        (lambda p: B.Beta(2, 2).log_prob(p["x"]), -0.1, -math.inf),
        (lambda p: B.Beta(2, 2).log_prob(p["x"]), 1.5, -math.inf),
        (lambda p: B.Beta(2, 2).log_prob(p["x"]), 0.0, -math.inf),
        (lambda p: B.Beta(2, 2).log_prob(p["x"]), 1.0, -math.inf),
        (lambda p: B.Beta(2, 2).log_prob(p["x"]), 0.5, math.log(1.5)),
        (lambda p: B.Gamma(2, 1).log_prob(p["x"]), -1.0, -math.inf),
        (lambda p: B.Gamma(2, 1).log_prob(p["x"]), 1.5, math.log(1.5) - 1.5),
        (lambda p: B.Exponential(2).log_prob(p["x"]), 0.0, math.log(2)),
        (lambda p: B.Exponential(2).log_prob(p["x"]), -1.0, -math.inf),
        (lambda p: B.Normal(0, 1).log_prob(p["x"]) + B.Categorical(probs=[0.2, 0.5, 0.3]).log_prob(1), 0.0,
         -0.5 * math.log(2 * math.pi) + math.log(0.5)),
    ]
    for fn, x, want in cases:
        model = compile_model(fn, {"x": 0.0}, cache=False)
        lp, gr = model.logp_grad(dev(np.array([[x]], dtype=np.float32)))
        got = float(lp.cpu()[0])
        if math.isinf(want):
            assert got == want and float(gr.cpu()[0, 0]) == 0.0
        else:
            assert abs(got - want) < 1e-5, (x, got, want)


# ------------------------------------------------------------------------------ parity 2: HMC / MH
@pytest.mark.parametrize("name", ["hmc_c1", "hmc_c2", "hmc_normal2d", "hmc_halfnormal"])
@pytest.mark.parametrize("lanes", [1, 8])
def test_hmc_injected_draws_match_reference(cuda, name, lanes):
    g = golden(name)
    kw = g["kwargs"]
    fn, init, _ = W.ALL_SMALL[g["model"]](B.ns)
    model = compile_model(fn, init)
    nw, ns_ = kw["num_warmup"], kw["num_samples"]
    tape = g["tape"]
    st = ChainState(model, model.pack(init, 1), kw["step_size"])
    zs = tape_normals(tape, model.layout, nw + ns_)
    us = tape_uniforms(tape, "accept", nw + ns_)
    en = torch.zeros(nw + ns_, 1, 2, device="cuda")
    ac = torch.zeros(nw + ns_, 1, dtype=torch.uint8, device="cuda")
    draws = torch.zeros(ns_, 1, model.D, device="cuda")
    launch_hmc(st, nw, kw["num_leapfrog_steps"], _cabi.ADAPT_REFERENCE, kw["target_accept"], 0, 0, lanes=lanes,
               inj_normal=dev(zs[:nw]), inj_uniform=dev(us[:nw]), trace_energy=en[:nw], trace_accept=ac[:nw])
    st.reset_counters()
    launch_hmc(st, ns_, kw["num_leapfrog_steps"], _cabi.ADAPT_NONE, kw["target_accept"], 0, nw, draws=draws, lanes=lanes,
               inj_normal=dev(zs[nw:]), inj_uniform=dev(us[nw:]), trace_energy=en[nw:], trace_accept=ac[nw:])
    torch.cuda.synchronize()
    ref_accept = np.array([it["accept"] for it in tape["iters"]], dtype=np.uint8)
    assert np.array_equal(ac.cpu().numpy()[:, 0], ref_accept)                       # every decision identical
    assert abs(float(st.step_size.cpu()[0]) - g["final_step_size"]) <= 1e-12 * g["final_step_size"]
    h0 = np.array([it["H0"] for it in tape["iters"]])
    h1 = np.array([it["H1"] for it in tape["iters"]])
    e = en.cpu().numpy()[:, 0]
    fin = np.isfinite(h1)
    assert rel_err(e[:, 0], h0) < 5e-5 and rel_err(e[fin, 1], h1[fin]) < 5e-5
    d = draws.cpu().numpy()[:, 0]
    for pname, (off, n, _) in model.layout.items():
        assert rel_err(d[:, off], np.asarray(g["draws"][pname])) < 1e-4, pname
    rate = float(st.n_accept.cpu()[0]) / ns_
    assert rate == g["accept_rate"]


@pytest.mark.parametrize("name", ["mh_c5", "mh_c1"])
def test_mh_injected_draws_match_reference(cuda, name):
    g = golden(name)
    kw = g["kwargs"]
    fn, init, _ = W.ALL_SMALL[g["model"]](B.ns)
    model = compile_model(fn, init)
    tape = g["tape"]
    start = tape["iters"][0]["start"]
    iters = tape["iters"][1:]
    n = kw["num_samples"]
    st = ChainState(model, model.pack(start, 1), 0.0)
    ac = torch.zeros(n, 1, dtype=torch.uint8, device="cuda")
    draws = torch.zeros(n, 1, model.D, device="cuda")
    launch_mh(st, n, kw["proposal_scale"], 0, 0, draws=draws, inj_normal=dev(tape_normals(tape, model.layout, n)),
              inj_uniform=dev(tape_uniforms(tape, "accept", n)), trace_accept=ac)
    torch.cuda.synchronize()
    assert np.array_equal(ac.cpu().numpy()[:, 0], np.array([it["accept"] for it in iters], dtype=np.uint8))
    d = draws.cpu().numpy()[:, 0]
    for pname, (off, _, _) in model.layout.items():
        assert rel_err(d[:, off], np.asarray(g["draws"][pname])) < 1e-6, pname
    assert float(st.n_accept.cpu()[0]) / n == g["accept_rate"]


# ------------------------------------------------------------------------------ parity 2: NUTS
def _dual_averaging_states(tape, nw, step_size, target=0.65):
    """(H_bar, eps_bar) BEFORE each warm-up iteration, recomputed on the host from the reference's recorded
    mean acceptance statistics with the recurrences of nuts.py:298-310 (float64 bookkeeping, float32 exp)."""
    import math
    f32 = np.float32
    mu = f32(np.log(f32(10.0 * step_size)))
    h_bar, eps_bar, out = 0.0, 1.0, []
    for m in range(nw):
        out.append((h_bar, eps_bar))
        a = tape["iters"][m]["alpha"]
        eta = 1.0 / (m + 10.0)
        h_bar = (1 - eta) * h_bar + eta * (target - a)
        le = f32(mu - f32((math.sqrt(m + 1) / 0.05) * h_bar))
        le = max(min(le, f32(10.0)), f32(-10.0))
        eps = float(np.exp(f32(le)))
        w = float(m + 1) ** -0.75
        eps_bar = float(np.exp(f32(w * math.log(eps) + (1 - w) * math.log(eps_bar))))
    out.append((h_bar, eps_bar))
    return out, float(mu)


@pytest.mark.parametrize("name", ["nuts_normal1d", "nuts_normal2d", "nuts_halfnormal_scale", "nuts_vector", "nuts_c2",
                                  "nuts_regression", "nuts_regression_sigma"])
@pytest.mark.parametrize("lanes", [1, 4])
def test_nuts_tree_decisions_match_reference(cuda, name, lanes):
    _replay_nuts(name, lanes)


@pytest.mark.parametrize("name", ["nuts_regression", "nuts_regression_sigma"])
def test_nuts_tree_decisions_match_reference_sync_schedule(cuda, name):
    """GLM class: the synchronous lock-step schedule (b2m_nuts_args.schedule = B2M_SCHED_SYNC) against the same
    reference transitions (the default, iteration-asynchronous fused-tick schedule is what the test above runs)."""
    _replay_nuts(name, 1, schedule=_cabi.SCHED_SYNC)


@pytest.mark.parametrize("name", ["nuts_regression", "nuts_regression_sigma"])
@pytest.mark.parametrize("path", ["tc", "simt"])
def test_nuts_tree_decisions_match_reference_unfused_paths(cuda, name, path):
    """GLM class: the tf32 / fp32 arithmetic paths keep the unfused asynchronous loop (tick, pack, K5, K6, finish as
    separate launches); the default fp16-encoded path runs the fused state kernel.  Same decisions either way."""
    _replay_nuts(name, 1, glm_path=path)


def _replay_nuts(name, lanes, schedule=0, glm_path="auto", copies=1, draw_tol=1e-5, alpha_tol=1e-4):
    """Every NUTS transition of the reference run is replayed on the GPU from the reference's own state
    (position, step size, dual-averaging state) with the reference's draws injected.  Replaying transition by
    transition keeps one-ulp differences of exp/log from being amplified by the step-size feedback loop, so
    the comparison is exact on decisions: direction, n', s', candidate taken, U-turn, depth."""
    g = golden(name)
    kw = g["kwargs"]
    fn, init, _ = W.ALL_SMALL[g["model"]](B.ns)
    model = compile_model(fn, init, glm_path=glm_path)
    nw, ns_, md = kw["num_warmup"], kw["num_samples"], kw["max_tree_depth"]
    tape = g["tape"]
    iters = tape["iters"]
    tot = nw + ns_
    inj_all = tape_nuts(tape, model.layout, tot, md)
    da, mu = _dual_averaging_states(tape, nw, kw["step_size"])
    q_prev = flat_params(model.layout, init)
    n_hits = 0
    for m in range(tot):
        it = iters[m]
        # `copies` identical chains fed the same draws: more than 128 rows puts the GLM contractions on CTA pairs
        st = ChainState(model, dev(np.repeat(q_prev[None, :], copies, axis=0)), it["eps"])
        h_bar, eps_bar = da[min(m, nw)]
        st.da_state[:, 0], st.da_state[:, 1], st.da_state[:, 2] = h_bar, eps_bar, mu
        inj = {k: dev(np.repeat(v[m:m + 1], copies, axis=1)) for k, v in inj_all.items()}
        tr = torch.full((1, copies, md, 6), -7, dtype=torch.int32, device="cuda")
        depth = torch.zeros(1, copies, dtype=torch.int32, device="cuda")
        alpha = torch.zeros(1, copies, device="cuda")
        h0 = torch.zeros(1, copies, device="cuda")
        draw = torch.zeros(1, copies, model.D, device="cuda")
        launch_nuts(st, 1, md, _cabi.ADAPT_DUAL_AVERAGING if m < nw else _cabi.ADAPT_NONE, _cabi.COMPAT_REFERENCE, 0.65,
                    0, m, draws=draw, depths=depth, alphas=alpha, lanes=lanes, inj=inj, trace_doubling=tr, trace_energy=h0,
                    schedule=schedule)
        torch.cuda.synchronize()
        q_new = flat_params(model.layout, it["q_new"])
        for c in sorted({0, copies - 1}):
            assert int(depth[0, c]) == it["depth"], (m, c, int(depth[0, c]), it["depth"])
            trc = tr.cpu().numpy()[0, c]
            for dbl in it["doublings"]:
                want = [dbl["v"], dbl["n_sub"], int(dbl["s_sub"]), int(dbl["took"]), int(dbl["s"]), dbl["n"]]
                assert trc[dbl["j"]].tolist() == want, (m, c, dbl, trc[dbl["j"]])
            assert abs(float(h0[0, c]) - it["H0"]) <= 1e-5 * max(1.0, abs(it["H0"]))
            assert abs(float(alpha[0, c]) - it["alpha"]) <= alpha_tol * max(it["alpha"], 1e-3), (m, float(alpha[0, c]), it["alpha"])
            assert rel_err(draw.cpu().numpy()[0, c], q_new) <= draw_tol + 1e-6 / max(np.max(np.abs(q_new)), 1e-30), (m, draw, q_new)
        if m < nw:   # dual averaging: step size for the next iteration and the running average
            want_eps = iters[m + 1]["eps"] if m + 1 < nw else None
            if want_eps is not None:
                assert abs(float(st.step_size[0]) - want_eps) <= 5e-4 * want_eps, (m, float(st.step_size[0]), want_eps)
            assert abs(float(st.da_state[0, 1]) - da[m + 1][1]) <= 5e-4 * da[m + 1][1]
        else:
            n_hits += int(st.n_accept[0])
        q_prev = q_new
    assert n_hits / ns_ == g["accept_rate"]
    assert abs(da[nw][1] - g["final_step_size"]) <= 1e-5 * g["final_step_size"]     # host recurrence == reference


@pytest.mark.parametrize("name", ["nuts_normal1d", "nuts_halfnormal_scale", "nuts_c2"])
def test_nuts_free_running_matches_reference(cuda, name):
    """The same runs free-running (own dual averaging, own state) with injected draws: for these models the
    whole 100+-iteration run reproduces the reference's draws."""
    g = golden(name)
    kw = g["kwargs"]
    fn, init, _ = W.ALL_SMALL[g["model"]](B.ns)
    model = compile_model(fn, init)
    nw, ns_, md = kw["num_warmup"], kw["num_samples"], kw["max_tree_depth"]
    tape = g["tape"]
    st = ChainState(model, model.pack(init, 1), kw["step_size"])
    st.da_state[:, 1] = 1.0
    st.da_state[:, 2] = float(np.log(np.float32(10.0 * kw["step_size"])))
    tot = nw + ns_
    inj = {k: dev(v) for k, v in tape_nuts(tape, model.layout, tot, md).items()}
    depths = torch.zeros(tot, 1, dtype=torch.int32, device="cuda")
    draws = torch.zeros(ns_, 1, model.D, device="cuda")

    def sl(lo, hi):
        return {k: v[lo:hi].contiguous() for k, v in inj.items()}

    launch_nuts(st, nw, md, _cabi.ADAPT_DUAL_AVERAGING, _cabi.COMPAT_REFERENCE, 0.65, 0, 0, depths=depths[:nw], inj=sl(0, nw))
    st.step_size.copy_(st.da_state[:, 1])
    st.n_accept.zero_()
    launch_nuts(st, ns_, md, _cabi.ADAPT_NONE, _cabi.COMPAT_REFERENCE, 0.65, 0, nw, draws=draws, depths=depths[nw:],
                inj=sl(nw, tot))
    torch.cuda.synchronize()
    assert np.array_equal(depths.cpu().numpy()[:, 0], np.array([it["depth"] for it in tape["iters"]]))
    assert abs(float(st.step_size.cpu()[0]) - g["final_step_size"]) <= 2e-5 * g["final_step_size"]
    d = draws.cpu().numpy()[:, 0]
    for pname, (off, n, shp) in model.layout.items():
        want = np.asarray(g["draws"][pname], dtype=np.float64).reshape(ns_, n)
        assert rel_err(d[:, off:off + n], want) < 2e-4, pname
    assert float(st.n_accept.cpu()[0]) / ns_ == g["accept_rate"]


@pytest.mark.parametrize("name", ["c1_normal", "c2_event_rate", "c5_ab_test", "t_halfnormal_scale", "t_vector_normal"])
def test_compact_and_general_pointwise_paths_agree_bit_for_bit(cuda, name):
    """Compact models (table in the kernel parameters, theta / gradient in registers) and the general path
    (shared-memory term table + mailbox) run the same arithmetic in the same order."""
    import mlx_mcmc_b200 as B
    import mlx_mcmc_b200.core as mx
    from mlx_mcmc_b200.engine import compile_model
    fn, init, _ = W.ALL_SMALL[name](B.ns)
    from mlx_mcmc_b200 import jit as J
    fast = compile_model(fn, init, cache=False)
    slow = compile_model(fn, init, cache=False, pointwise_path="general")
    spec = compile_model(fn, init, cache=False)
    assert J.specialize(spec, True) and spec.jit and spec.lib.b2m_model_has_module(spec.handle) == 1
    rng = np.random.default_rng(5)
    theta = fast.pack(init, 257) + torch.from_numpy(0.3 * rng.standard_normal((257, fast.D)).astype(np.float32)).cuda()
    for lanes in (1, 4):
        a, ga = fast.logp_grad(theta, lanes=lanes)
        for other in (slow, spec):      # general interpreter path; NVRTC-specialised kernels
            b, gb = other.logp_grad(theta, lanes=lanes)
            assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a, nan=7.0), torch.nan_to_num(b, nan=7.0))
            assert torch.equal(torch.nan_to_num(ga, nan=7.0), torch.nan_to_num(gb, nan=7.0))
    kw = dict(num_samples=40, num_warmup=30, num_chains=96, key=mx.random.key(3))
    if name != "t_vector_normal":           # hmc stores scalar parameters only in the reference; ours handles both
        sa, ra = B.hmc(fn, init, model=fast, **kw)
        for other in (slow, spec):
            sb, rb = B.hmc(fn, init, model=other, **kw)
            assert ra == rb and all(np.array_equal(sa[k], sb[k], equal_nan=True) for k in sa)
        if name != "c5_ab_test":
            ma, qa = B.metropolis_hastings(fn, init, num_samples=50, proposal_scale=0.05, num_chains=96, model=fast)
            mb, qb = B.metropolis_hastings(fn, init, num_samples=50, proposal_scale=0.05, num_chains=96, model=spec)
            assert qa == qb and all(np.array_equal(ma[k], mb[k], equal_nan=True) for k in ma)
    na, _ = B.nuts(fn, init, model=fast, **kw)
    for other in (slow, spec):
        nb, _ = B.nuts(fn, init, model=other, **kw)
        assert all(np.array_equal(na[k], nb[k], equal_nan=True) for k in na)


def test_edge_shapes_empty_single_and_one_by_one(cuda):
    """Edge inputs: an empty observation array (the likelihood sums to 0), a single observation, a 1 x 1 regression
    with one chain -- each against the oracle's float64 evaluation."""
    import mlx_mcmc_b200 as B
    from oracle.ns import ns as ons, value_and_grad

    def make(ns, y):
        def fn(p):
            lik = ns.mx.sum(ns.Exponential(p["rate"]).log_prob(ns.mx.array(y))) if len(y) else 0.0
            return ns.Gamma(2.0, 1.0).log_prob(p["rate"]) + lik
        return fn
    for y in (np.zeros(0, np.float32), np.array([0.7], np.float32), np.array([0.2, 1.1, 0.4], np.float32)):
        model = compile_model(make(B.ns, y), {"rate": 1.3}, cache=False)
        lp, g = model.logp_grad(model.pack({"rate": 1.3}, 3))
        lp64, g64 = value_and_grad(make(ons, y), {"rate": 1.3}, "float64")
        assert abs(float(lp[0]) - lp64) <= 1e-5 * max(1.0, abs(lp64)) and abs(float(g[0, 0]) - g64["rate"]) <= 1e-5 * max(1.0, abs(g64["rate"]))
        assert torch.equal(lp, lp[0].expand_as(lp))
    # 1 x 1 regression, one chain (everything is padding except one element)
    X = np.array([[2.0]], np.float32)
    yv = np.array([3.0], np.float32)

    def reg(ns):
        Xa, ya = ns.mx.array(X), ns.mx.array(yv)
        return lambda p: ns.mx.sum(ns.Normal(0, 10.0).log_prob(p["beta"])) + ns.mx.sum(ns.Normal(Xa @ p["beta"], 1.0).log_prob(ya))
    init = {"beta": np.array([0.4], np.float32)}
    gm = compile_model(reg(B.ns), init, cache=False)
    assert gm.model_class == 1
    lp, g = gm.logp_grad(gm.pack(init, 1))
    lp64, g64 = value_and_grad(reg(ons), init, "float64")
    assert abs(float(lp[0]) - lp64) <= 1e-5 * abs(lp64) and abs(float(g[0, 0]) - g64["beta"][0]) <= 1e-5 * abs(g64["beta"][0])
    s, _ = B.nuts(reg(B.ns), init, num_samples=50, num_warmup=50, num_chains=1, compat="correct", model=gm)
    assert s["beta"].shape == (50, 1) and np.isfinite(s["beta"]).all()
