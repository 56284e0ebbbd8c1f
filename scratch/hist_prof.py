"""Per-phase cycle profile of hist_scan_kernel (profiling build, -DFILO_HIST_PROF).  Run on the GPU box:
    python scratch/hist_prof.py [series]
Uses scratch/libfilo_b200_prof.so (built from the same sources with the profiling hooks compiled in)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import filodb_b200.capi as capi
capi.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfilo_b200_prof.so")
os.environ["FILO_HIST_V2"] = "0"          # the hooks live in the first kernel
from oracle import hist as H

S = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 400
K = min(S, 1024); nb = 20; ROWS = 480; RPC = 400; T0 = 1_700_000_000_000
b = H.Buckets.custom([2.0 * 3 ** i for i in range(nb - 1)] + [float("inf")])
rng = np.random.default_rng(42)
st = H.HistStore(b)
ts = T0 + np.arange(ROWS, dtype=np.int64) * 15000
for s in range(K):
    obs = np.zeros((ROWS, nb), np.int64)
    obs[np.arange(ROWS), (np.arange(ROWS) + s) % nb] = 1 + rng.integers(0, 3, ROWS)
    rows = np.cumsum(np.cumsum(obs, axis=1), axis=0)
    if s % 97 == 0: rows[300:] = np.cumsum(np.cumsum(obs[300:], axis=1), axis=0)
    st.add_series(ts, rows, [RPC, ROWS - RPC])
nch_k, addrs_k = st.all_info_addrs()
reps = (S + K - 1) // K
nch = np.tile(nch_k, reps)[:S].copy()
addrs = np.tile(addrs_k.reshape(K, -1), (reps, 1))[:S].reshape(-1).copy()
ctx = capi.Context(0)
tab = ctx.load_series(nch, addrs, schema_flags=capi.SCHEMA_CUMULATIVE)
L = capi.lib()
L.filo_debug_hist_prof.argtypes = [C.c_void_p, C.c_int]
names = ["stage record", "thread-0 chunk/section tables", "timestamps + section bases", "rows decode", "in-chunk corrections",
         "carried corr + window descriptors", "(window,bucket) rates", "item partial write"]
for label, kw in (("agg sum + quantile", dict(aggr=capi.AGG_SUM, quantile=0.99, want_values=False)),):
    for rep in range(2):
        out = np.zeros(16, np.uint64)
        L.filo_debug_hist_prof(out.ctypes.data, 1)
        t = time.perf_counter()
        ctx.query_hist(tab, capi.FN_RATE, T0, 15000, T0 + 7200000, 300000, **kw)
        dt = time.perf_counter() - t
        kns = ctx.last_stats["kernel_ns"]
        L.filo_debug_hist_prof(out.ctypes.data, 1)
    ctas = int(out[15]); tot = float(out[:12].sum())
    print("%s: S=%d kernel %.2f ms (wall %.2f ms), %d CTAs, %.0f cycles per series per CTA" % (label, S, kns / 1e6, dt * 1e3, ctas, tot / S))
    for i, n in enumerate(names):
        print("  %-36s %8.0f cycles/series  %5.1f %%" % (n, float(out[i]) / S, 100.0 * float(out[i]) / tot))
