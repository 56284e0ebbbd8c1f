"""Per-phase cycle profile of the histogram scan kernels (profiling build, -DFILO_HIST_PROF).
Build the variant here, then run on the GPU box:
    FILO_NVCC_EXTRA=-DFILO_HIST_PROF FILO_BUILD_OUT=scratch/libfilo_b200_prof.so python -m filodb_b200.build --force
    python scratch/hist_prof.py [series]          # second kernel, then the first (FILO_HIST_V2=0)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import filodb_b200.capi as capi
capi.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfilo_b200_prof.so")
from oracle import hist as H

S = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 2 * 600
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
L = capi.lib()
L.filo_debug_hist_prof.argtypes = [C.c_void_p, C.c_int]
L.filo_debug_hist2_prof.argtypes = [C.c_void_p, C.c_int]
names1 = ["stage record", "thread-0 chunk/section tables", "timestamps + section bases", "rows decode", "in-chunk corrections",
          "carried corr + window descriptors", "(window,bucket) rates", "item partial write"]
names2 = ["wait for the prefetched record", "thread-0 chunk/section tables", "timestamps + rows decode", "SectDelta base add (+ prefetch issue)",
          "corrections inside / across chunks", "windows: descriptors + rates + partial row"]
q = dict(aggr=capi.AGG_SUM, quantile=0.99, want_values=False)
for label, env, fnname, names in (("second kernel", "1", "filo_debug_hist2_prof", names2), ("first kernel", "0", "filo_debug_hist_prof", names1)):
    os.environ["FILO_HIST_V2"] = env
    tab = ctx.load_series(nch, addrs, schema_flags=capi.SCHEMA_CUMULATIVE)      # the work items are sized at load time
    fn = getattr(L, fnname)
    for rep in range(2):
        out = np.zeros(16, np.uint64)
        fn(out.ctypes.data, 1)
        ctx.query_hist(tab, capi.FN_RATE, T0, 15000, T0 + 7200000, 300000, **q)
        kns = ctx.last_stats["kernel_ns"]
        fn(out.ctypes.data, 1)
    tot = float(out[:12].sum())
    print("%s: S=%d kernel %.2f ms, %d CTAs, %.0f cycles per series per CTA" % (label, S, kns / 1e6, int(out[15]), tot / S))
    for i, n in enumerate(names):
        print("  %-44s %8.0f cycles/series  %5.1f %%" % (n, float(out[i]) / S, 100.0 * float(out[i]) / tot))
    tab.free()
