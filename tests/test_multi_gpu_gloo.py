"""world_size-2 CPU test (gloo) of the host-side multi-GPU logic: shard assignment and the cross-shard merge of
FILO_Q_PARTIAL aggregates (filodb_b200/shard.py), checked against the oracle over the unsharded set of series."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from filodb_b200 import shard

T0, STEP, ROWS = 1_700_000_000_000, 15000, 120
N_SERIES, N_GROUPS = 24, 5
QUERY = (T0 + 300000, STEP, T0 + (ROWS - 1) * STEP, 300000)
AGGS = [("AGG_SUM", shard.AGG_SUM), ("AGG_AVG", shard.AGG_AVG), ("AGG_MIN", shard.AGG_MIN), ("AGG_MAX", shard.AGG_MAX),
        ("AGG_COUNT", shard.AGG_COUNT)]


def _series(i):
    rng = np.random.default_rng(1000 + i)
    ts = T0 + np.arange(ROWS, dtype=np.int64) * STEP
    v = 15 + np.sin(np.arange(1, ROWS + 1)) + rng.normal(0, 1, ROWS)
    v[rng.random(ROWS) < 0.05] = np.nan
    if i % 7 == 3:
        v[:] = np.nan                         # a series that contributes nothing
    return ts, v


def _group(i):
    return (i * 7 + 3) % N_GROUPS


def _store(o, ids):
    st = o.Store()
    for i in ids:
        ts, v = _series(i)
        st.add_series_rows(ts, v, [80, 40], val_mode=1, detect_drops=False)
    return st


def _partials(per_series, gids, agg):
    """FILO_Q_PARTIAL form (include/filo_b200.h) from per-series window results."""
    Tn = per_series.shape[1]
    ident = {shard.AGG_MIN: np.inf, shard.AGG_MAX: -np.inf}.get(agg, 0.0)
    vals = np.full((N_GROUPS, Tn), ident); cnts = np.zeros((N_GROUPS, Tn), np.int64)
    for row, g in zip(per_series, gids):
        ok = ~np.isnan(row)
        cnts[g] += ok
        if agg in (shard.AGG_SUM, shard.AGG_AVG):
            vals[g] += np.where(ok, row, 0.0)
        elif agg == shard.AGG_COUNT:
            vals[g] += ok
        elif agg == shard.AGG_MIN:
            vals[g] = np.minimum(vals[g], np.where(ok, row, np.inf))
        else:
            vals[g] = np.maximum(vals[g], np.where(ok, row, -np.inf))
    return vals, cnts


def _present(vals, cnts, agg):
    """What filo_present_partials does on the device (RowAggregator.present)."""
    out = vals / np.maximum(cnts, 1) if agg == shard.AGG_AVG else vals.copy()
    out[cnts == 0] = np.nan
    return out


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import oracle as o
        b, e = shard.series_range_of_rank(N_SERIES, rank, world)
        ids = list(range(b, e))
        st = _store(o, ids)
        per = st.query(o.FN_SUM_OVER_TIME, *QUERY)
        res = {}
        for name, agg in AGGS:
            v, c = _partials(per, [_group(i) for i in ids], agg)
            tv, tc = torch.from_numpy(v.copy()), torch.from_numpy(c.copy())
            shard.merge_partials(tv, tc, agg, dist)
            res[name] = _present(tv.numpy(), tc.numpy(), agg)
        ms = shard.max_over_ranks(10.0 + rank, dist, "cpu")
        q.put((rank, res, ms))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as ex:            # surface the failure in the parent instead of hanging it
        q.put((rank, repr(ex), None))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_assignment():
    assert shard.shards_of_rank(128, 3, 8) == list(range(3, 128, 8))
    all_shards = sorted(s for r in range(8) for s in shard.shards_of_rank(128, r, 8))
    assert all_shards == list(range(128))
    with pytest.raises(ValueError):
        shard.shards_of_rank(96, 0, 8)
    spans = [shard.series_range_of_rank(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard.series_range_of_rank(2, 3, 4) == (2, 2)


@pytest.mark.timeout(120)
def test_two_rank_merge_matches_unsharded_oracle(oracle):
    o = oracle
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    got = [q.get(timeout=100) for _ in range(world)]
    for p in procs: p.join(timeout=30)
    for rank, res, ms in got:
        assert not isinstance(res, str), "rank %d failed: %s" % (rank, res)
        assert ms == 11.0                      # max over ranks of (10 + rank)
    full = _store(o, range(N_SERIES))
    gids = np.array([_group(i) for i in range(N_SERIES)], np.int32)
    for name, agg in AGGS:
        exp = full.query(o.FN_SUM_OVER_TIME, *QUERY, aggr=getattr(o, name), group_ids=gids, n_groups=N_GROUPS)
        exp = np.asarray(exp[0] if isinstance(exp, tuple) else exp).reshape(N_GROUPS, -1)
        for rank, res, _ in got:
            a = res[name]
            assert a.shape == exp.shape
            assert (np.isnan(a) == np.isnan(exp)).all(), name
            m = ~np.isnan(exp)
            if agg in (shard.AGG_SUM, shard.AGG_AVG):
                np.testing.assert_allclose(a[m], exp[m], rtol=1e-9, atol=0)      # fold order differs across shards
            else:
                assert (a[m] == exp[m]).all(), name
