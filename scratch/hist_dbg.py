import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import hist as H, oracle as o
from filodb_b200 import capi
from tests.test_gpu_parity import _hist_series
def P(*a): print(*a, flush=True)
ctx = capi.Context(0)
rng = np.random.default_rng(21)
t0, rows = 1_700_000_000_000, 240
b = H.Buckets.geometric(2.0, 2.0, 12)
st = H.HistStore(b); S = 12
for s in range(S):
    ts = t0 + np.arange(rows, dtype=np.int64) * 15000
    st.add_series(ts, _hist_series(rng, rows, b.n, () if s % 3 else (77,)), [160, 80])
nch, addrs = st.all_info_addrs()
gids = np.arange(S, dtype=np.int32) % 4
tab = ctx.load_series(nch, addrs, group_ids=gids, n_groups=4, schema_flags=capi.SCHEMA_CUMULATIVE)
P("loaded")
queries = [(t0 + 300000, 15000, t0 + (rows - 1) * 15000, 300000), (t0 - 60000, 47000, t0 + rows * 15000 + 90000, 333333), (t0 + 2000000, 1, t0 + 2000000, 600000)]
for q in queries:
    for name in ("FN_RATE", "FN_INCREASE"):
        P("per-series", name, q[1]); ctx.query_hist(tab, getattr(capi, name), *q)
        P("agg", name, q[1]); ctx.query_hist(tab, getattr(capi, name), *q, aggr=capi.AGG_SUM, quantile=0.99)
P("sum_over_time"); ctx.query_hist(tab, capi.FN_SUM_OVER_TIME, *queries[0])
P("sum_over_time agg"); ctx.query_hist(tab, capi.FN_SUM_OVER_TIME, *queries[0], aggr=capi.AGG_SUM, quantile=0.5)
tab.free()
st2 = H.HistStore(b)
for s in range(5):
    obs = np.cumsum(rng.integers(0, 9, (rows, b.n)), axis=1).astype(np.int64)
    st2.add_series(t0 + np.arange(rows, dtype=np.int64) * 15000, obs, [100, 100, 40], sect=False)
nch2, addrs2 = st2.all_info_addrs()
tab2 = ctx.load_series(nch2, addrs2, schema_flags=0)
P("loaded delta")
for name in ("FN_RATE", "FN_INCREASE", "FN_SUM_OVER_TIME"):
    P("delta", name); ctx.query_hist(tab2, getattr(capi, name), *queries[0])
P("delta agg"); ctx.query_hist(tab2, capi.FN_RATE, *queries[0], aggr=capi.AGG_SUM, quantile=0.5)
P("done")
