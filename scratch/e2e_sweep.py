# filo_scan_series pipeline sweep on one B200 box: one table / host mirror, then the C2 end-to-end step under different pipeline settings
# (slots, batch bytes, plan chunk), plus a device timeline of the default and of the best setting (FILO_SCAN_TRACE).
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import filodb_b200.capi as capi
S = int(os.environ.get("SWEEP_SERIES", "10000000"))
os.environ.setdefault("FILO_HOST_THREADS", str(max(4, min(64, bench.host_cores()["cores_usable"]))))
synth, fn_name, aggr_name, n_groups, desc = bench.WORKLOADS["c2"]
fn = getattr(capi, fn_name)
ctx = capi.Context(0)
tab = ctx.synth_table(S, bench.ROWS, bench.ROWS_PER_CHUNK, bench.T0_MS, bench.INTERVAL, n_groups=n_groups, seed=42, series_id_base=0, **synth)
start, step, end, window = bench.query_range("c2")
T = capi.num_windows(start, step, end)
arena, rec_off = tab.read_arena(0, S)
nch, addrs, keep = bench.host_chunk_infos(arena, rec_off, S)
tab.free()
hout = torch.empty(S * T, dtype=torch.float64).pin_memory(); hout_np = hout.numpy()
L = capi.lib()
ctx.host_register(arena)
def one():
    st_ = capi.Stats()
    ctx._check(L.filo_scan_series(ctx.h, S, nch.ctypes.data_as(C.c_void_p), addrs.ctypes.data_as(C.c_void_p), 0, 1, synth.get("schema_flags", 0),
                                  fn, start, step, end, window, hout_np.ctypes.data_as(C.c_void_p), C.byref(st_)))
def run(env, steps=2, trace=None):
    for k in ("FILO_SCAN_SLOTS", "FILO_SCAN_SLAB_MB", "FILO_SCAN_PLAN_CHUNK", "FILO_SCAN_OUT_MB", "FILO_SCAN_TRACE"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    one()
    t0 = time.perf_counter()
    for _ in range(steps): one()
    dt = (time.perf_counter() - t0) / steps
    if trace:
        os.environ["FILO_SCAN_TRACE"] = trace; one(); os.environ.pop("FILO_SCAN_TRACE")
    return dt
res = []
tag = os.environ.get("SWEEP_TAG", "d")
for i, env in enumerate([{}, {"FILO_SCAN_SLOTS": 4}] if tag != "full" else [{}, {"FILO_SCAN_SLOTS": 3}, {"FILO_SCAN_SLOTS": 4}, {"FILO_SCAN_SLOTS": 8}]):
    dt = run(env, steps=3, trace=("gpurun_out/scan_trace_%s%d.csv" % (tag, i)) if i == 0 else None)
    res.append({"tag": tag, "CUDA_DEVICE_MAX_CONNECTIONS": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "env": env, "s_per_step": round(dt, 4), "G_samples_per_s": round(S * bench.ROWS / dt / 1e9, 3)})
    print(res[-1], flush=True)
json.dump(res, open("gpurun_out/e2e_sweep_%s.json" % tag, "w"), indent=1)
ctx.host_unregister(arena)
