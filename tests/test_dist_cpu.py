"""CPU: the host-side multi-GPU logic (mlx_mcmc_b200/dist.py) on a world_size-2 gloo group -- partitions tile,
the final gather restores rank order with unequal shard sizes, observation sharding slices the traced model."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import mlx_mcmc_b200 as B
from mlx_mcmc_b200 import dist as D, workloads as W
from mlx_mcmc_b200.tracer import OP_MATVEC, trace


def test_shard_chains_tiles_in_rank_order():
    for n in (1, 2, 7, 64, 65536, 4097):
        for world in (1, 2, 3, 8):
            blocks = [D.shard_chains(n, r, world) for r in range(world)]
            assert sum(c for c, _ in blocks) == n
            off = 0
            for c, o in blocks:
                assert o == off
                off += c
            assert max(c for c, _ in blocks) - min(c for c, _ in blocks) <= 1
    with pytest.raises(ValueError):
        D.shard_chains(4, 2, 2)


def test_shard_rows_cover_and_align():
    for n in (100, 10000, 100000, 12345):
        for world in (1, 2, 4, 8):
            rows = [D.shard_rows(n, r, world, multiple=256) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert a1 == b0 and a0 <= a1
            for r0, r1 in rows:
                assert r0 % 256 == 0 or r1 == r0        # every non-empty shard starts tile-aligned


def test_shard_observations_slices_only_the_likelihood():
    fn, init, meta = W.regression(B.ns, 1000, 8, seed=1)
    tr = trace(fn, init)
    parts = [D.shard_observations(tr, r, 4) for r in range(4)]
    lik = lambda t: next(x for x in t.terms if x.p0.kind == OP_MATVEC)  # noqa: E731
    assert sum(lik(p).length for p in parts) == 1000
    X = np.concatenate([p.arrays[lik(p).p0.a] for p in parts])
    y = np.concatenate([p.arrays[lik(p).x.a] for p in parts])
    assert np.array_equal(X, meta.X) and np.array_equal(y, meta.y)
    assert lik(tr).length == 1000 and tr.arrays[lik(tr).p0.a].shape == (1000, 8)      # the input is untouched
    for p in parts:                                                                     # priors are not sharded
        assert [t.length for t in p.terms if t.p0.kind != OP_MATVEC] == [t.length for t in tr.terms if t.p0.kind != OP_MATVEC]
    fn2, init2, _ = W.c2_event_rate(B.ns)
    with pytest.raises(ValueError):
        D.shard_observations(trace(fn2, init2), 0, 2)
    with pytest.raises(ValueError):
        D.shard_observations(trace(*W.regression(B.ns, 3, 2)[:2]), 3, 4)                 # a rank with no rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert D.rank_world() == (rank, world)
        n_total = 5                                           # 3 + 2: unequal shards
        count, off = D.shard_chains(n_total, rank, world)
        draws = {"mu": (np.arange(off, off + count, dtype=np.float32)[:, None] * 10 + np.arange(4, dtype=np.float32)[None]),
                 "beta": torch.arange(off, off + count, dtype=torch.float32)[:, None, None].expand(count, 4, 3).contiguous()}
        full = D.gather_draws(draws, count)
        assert isinstance(full["mu"], np.ndarray) and full["mu"].shape == (n_total, 4)
        assert np.array_equal(full["mu"][:, 0], 10 * np.arange(n_total))
        assert torch.equal(full["beta"][:, 0, 0], torch.arange(n_total, dtype=torch.float32))
        st = D.reduce_stats({"acc": float(rank + 1), "n": 1.0})
        assert st == {"acc": 3.0, "n": 2.0}
        assert D.reduce_stats({"t": float(rank)}, "max") == {"t": 1.0}
        with pytest.raises(ValueError):
            D.gather_draws({"mu": np.zeros((count + 1, 2), np.float32)}, count)
        # merge of a sliced call: the owner wrote its slice of zero-initialised outputs and advanced its counters
        n, per = 6, 3
        own = slice(rank * per, (rank + 1) * per)
        draws = torch.zeros(2, n, 3)
        draws[:, own] = torch.arange(n, dtype=torch.float32)[None, own, None] + 1.0
        before = torch.full((n,), 7, dtype=torch.int64)
        leaves = before.clone()
        leaves[own] += torch.arange(n)[own] + 1
        D.merge_slices((draws,), [(leaves, before)])
        assert torch.equal(draws[0, :, 0], torch.arange(n, dtype=torch.float32) + 1.0)
        assert torch.equal(leaves, 7 + torch.arange(n) + 1)
        with pytest.raises(ValueError):
            D.run_sharded(lambda p: 0, {"x": 0.0}, method="hmc", shard="obs", slice_state=True)
        # run_sharded refuses bad modes before touching a device
        with pytest.raises(ValueError):
            D.run_sharded(lambda p: 0, {"x": 0.0}, method="gibbs")
        with pytest.raises(ValueError):
            D.run_sharded(lambda p: 0, {"x": 0.0}, method="hmc", shard="rows")
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        td.destroy_process_group()


def test_gather_and_reduce_on_gloo_world_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, "ok"), (1, "ok")], res


def test_sliced_state_preconditions():
    """nuts(slice_state=True) never falls back silently: it needs an observation-sharded GLM model on > 1 ranks and a
    chain count the ranks (and the 256-row tiles) divide."""
    from types import SimpleNamespace as NS
    from mlx_mcmc_b200.kernels.nuts import _sliced_world
    comm2 = NS(world=2)
    assert _sliced_world(NS(_comm=comm2, model_class=1), 512, "async") == 2
    with pytest.raises(ValueError):
        _sliced_world(NS(_comm=comm2, model_class=1), 512, "sync")
    for model, chains in ((NS(_comm=None, model_class=1), 512), (NS(_comm=NS(world=1), model_class=1), 512),
                          (NS(_comm=comm2, model_class=0), 512), (NS(_comm=comm2, model_class=1), 384),
                          (NS(_comm=NS(world=3), model_class=1), 512)):
        with pytest.raises(ValueError):
            _sliced_world(model, chains, "async")
