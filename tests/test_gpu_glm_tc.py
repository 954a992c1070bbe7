"""GPU: the tcgen05 3xTF32 contractions (K5/K6) against float64 numpy and against the fp32 SIMT path."""
import math
import os

import numpy as np
import pytest
import torch

import mlx_mcmc_b200 as B
from mlx_mcmc_b200 import workloads as W
from mlx_mcmc_b200.engine import compile_model

pytestmark = pytest.mark.gpu


def _model(path, n, d, seed):
    old = os.environ.get("B2M_GLM_PATH")
    os.environ["B2M_GLM_PATH"] = path
    try:
        fn, init, meta = W.regression(B.ns, n, d, seed=seed)
        return compile_model(fn, init, cache=False), meta
    finally:
        if old is None:
            os.environ.pop("B2M_GLM_PATH", None)
        else:
            os.environ["B2M_GLM_PATH"] = old


def _float64(meta, theta):
    X, y, b = meta.X.astype(np.float64), meta.y.astype(np.float64), theta.astype(np.float64)
    n, d = X.shape
    r = y[None, :] - b @ X.T
    lp = (-0.5 * (r ** 2).sum(1) - n * 0.5 * math.log(2 * math.pi)
          - 0.5 * (b ** 2).sum(1) / 100.0 - d * (0.5 * math.log(2 * math.pi) + math.log(10.0)))
    return lp, r @ X - b / 100.0


@pytest.mark.parametrize("n,d,c", [(200, 5, 7), (1000, 100, 130), (5000, 96, 300), (3000, 300, 1024), (40000, 64, 256)])
def test_tc_logp_grad_vs_float64_and_simt(cuda, n, d, c):
    tc, meta = _model("tc", n, d, seed=n + d)
    simt, _ = _model("simt", n, d, seed=n + d)
    rng = np.random.default_rng(1)
    theta = (meta.beta_true[None, :] + 0.2 * rng.standard_normal((c, d))).astype(np.float32)
    t = torch.from_numpy(theta).cuda()
    lp_tc, g_tc = tc.logp_grad(t)
    lp_s, g_s = simt.logp_grad(t)
    torch.cuda.synchronize()
    lp64, g64 = _float64(meta, theta)
    for lp, g, tol in ((lp_tc, g_tc, 1e-5), (lp_s, g_s, 1e-5)):
        assert np.max(np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64)) < tol
        assert np.max(np.abs(g.cpu().numpy() - g64)) / np.max(np.abs(g64)) < tol
    # 3xTF32 with chunked promotion is as accurate as fp32 FMA tiles
    e_tc = np.max(np.abs(g_tc.cpu().numpy() - g64)) / np.max(np.abs(g64))
    e_s = np.max(np.abs(g_s.cpu().numpy() - g64)) / np.max(np.abs(g64))
    print(f"n={n} d={d} c={c}: grad err tc {e_tc:.2e} simt {e_s:.2e}; "
          f"logp err tc {np.max(np.abs(lp_tc.cpu().numpy() - lp64) / np.abs(lp64)):.2e}")
    assert e_tc < 3e-6


def test_tc_value_only_and_repeatability(cuda):
    tc, meta = _model("tc", 2000, 40, seed=3)
    theta = torch.from_numpy(np.tile(meta.beta_true, (64, 1)).astype(np.float32)).cuda()
    a, ga = tc.logp_grad(theta)
    b, _ = tc.logp_grad(theta, want_grad=False)
    c, gc = tc.logp_grad(theta)
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(ga, gc)      # deterministic, bit for bit
    assert torch.equal(a, a[0].expand_as(a))                                     # identical rows -> identical results
