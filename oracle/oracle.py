"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding over oracle/_build/libfilo_oracle.so (the C++ restatement of FiloDB's chunk-scan +
range-function path, see filo_format.hpp / filo_query.hpp for the reference file:line citations).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfilo_oracle.so")

# RangeFn / AggrOp numbering == include/filo_b200.h
FN_LAST, FN_RATE, FN_INCREASE, FN_DELTA, FN_SUM_OVER_TIME, FN_AVG_OVER_TIME, FN_COUNT_OVER_TIME, \
    FN_MIN_OVER_TIME, FN_MAX_OVER_TIME, FN_TIMESTAMP, FN_STDDEV_OVER_TIME, FN_STDVAR_OVER_TIME, FN_CHANGES, FN_QUANTILE_OVER_TIME, \
    FN_ZSCORE, FN_HOLT_WINTERS, FN_PREDICT_LINEAR, FN_MAD_OVER_TIME, FN_PRESENT_OVER_TIME = range(19)
AGG_NONE, AGG_SUM, AGG_AVG, AGG_MIN, AGG_MAX, AGG_COUNT, AGG_TOPK, AGG_BOTTOMK = range(8)
VAL_OPTIMIZE, VAL_XOR, VAL_RAW = 0, 1, 2
TS_OPTIMIZE, TS_RAW = 0, 2


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", "Makefile"))]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _sig(_lib)
    return _lib


def _sig(L):
    i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    P = C.POINTER
    L.fo_last_error.restype = C.c_char_p
    for name, res, args in [
        ("fo_pack8", i32, [vp, vp, i32]), ("fo_unpack8", i32, [vp, i32, vp, P(i32)]),
        ("fo_pack_doubles", i32, [vp, i32, vp, i32]), ("fo_unpack_double_xor", i32, [vp, i32, vp, i32]),
        ("fo_pack_delta", i32, [vp, i32, vp, i32]), ("fo_unpack_delta", i32, [vp, i32, vp, i32]),
        ("fo_encode_timestamps", i32, [vp, i32, vp, i32]), ("fo_encode_longs", i32, [vp, i32, vp, i32]),
        ("fo_encode_doubles", i32, [vp, i32, i32, i32, vp, i32]),
        ("fo_encode_int_vector", i32, [vp, i32, i32, i32, vp, i32]),
        ("fo_minmax_to_nbits", None, [i32, i32, P(i32), P(i32)]),
        ("fo_long_length", i32, [vp]), ("fo_long_apply", i64, [vp, i32]),
        ("fo_long_binary_search", i32, [vp, i64]), ("fo_long_ceiling_index", i32, [vp, i64]),
        ("fo_long_sum", f64, [vp, i32, i32]),
        ("fo_int_length", i32, [vp]), ("fo_int_apply", i32, [vp, i32]), ("fo_int_sum", i64, [vp, i32, i32]),
        ("fo_double_length", i32, [vp]), ("fo_double_apply", f64, [vp, i32]),
        ("fo_double_sum", f64, [vp, i32, i32]), ("fo_double_count", i32, [vp, i32, i32]),
        ("fo_double_dropped", i32, [vp]), ("fo_total_bytes", i32, [vp]), ("fo_vector_type", i32, [vp]),
        ("fo_double_detect_drop", None, [vp, vp, vp]), ("fo_double_update_correction", None, [vp, vp, i32, vp]),
        ("fo_double_corrected_value", f64, [vp, i32, vp]), ("fo_double_drop_positions", i32, [vp, vp, i32]),
        ("fo_extrapolated_rate", f64, [i64, i64, i32, i64, f64, i64, f64, i32, i32]),
        ("fo_chunk_id", i64, [i64, i64]), ("fo_start_time_from_chunk_id", i64, [i64]),
        ("fo_store_new", vp, []), ("fo_store_free", None, [vp]), ("fo_store_num_series", i64, [vp]),
        ("fo_store_add_series", i64, [vp]),
        ("fo_store_add_chunk", i32, [vp, i64, vp, vp, i32, i32, i32, i32]),
        ("fo_store_add_chunk_raw", i32, [vp, i64, i64, i64, i32, vp, i32, vp, i32]),
        ("fo_store_add_series_rows", i32, [vp, vp, vp, i64, vp, i32, i32, i32, i32]),
        ("fo_store_add_from_arena", i64, [vp, vp, vp, i64]),
        ("fo_store_add_synth", i64, [vp, i64, i32, i32, i64, i32, i32, i32, i32, i32, i32, i32, C.c_uint64, i64, vp, i32]),
        ("fo_synth_group_ids", i32, [C.c_uint64, i64, i64, i32, vp]),
        ("fo_store_num_chunks", i32, [vp, i64]), ("fo_store_info_addrs", None, [vp, i64, vp]),
        ("fo_store_vector_bytes", i64, [vp, i64, i32, i32, vp, i64]),
        ("fo_store_algorithmic_bytes", i64, [vp]),
        ("fo_query", i32, [vp, i32, i32, i64, i64, i64, i64, i32, i32, i32, vp, i32, i32, i64, i64, vp, vp, vp]),
        ("fo_query2", i32, [vp, i32, i32, f64, f64, i64, i64, i64, i64, i32, i32, i32, vp, i32, i32, i64, i64, vp, vp, vp]),
        ("fo_store_add_chunk_longs", i32, [vp, i64, vp, vp, i32, i32, i32]),
        ("fo_num_windows", i32, [i64, i64, i64]),
        ("fo_serialize_result", i64, [vp, i64, i32, i64, i64, i64, i64, vp, i64, vp, vp, vp]),
        ("fo_result_rows", i64, [vp, i64, i32, i32, i64, i64, i64, i64, vp, vp, i64]),
        ("fo_sliding", None, [vp, vp, i64, i32, i32, i64, i64, i64, i64, vp]),
    ]:
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def last_error():
    return lib().fo_last_error().decode()


# ---- codecs -------------------------------------------------------------------------------------
def pack8(vals8):
    a = np.asarray(vals8, dtype=np.uint64)
    assert a.size == 8
    out = np.zeros(80, np.uint8)
    n = lib().fo_pack8(_p(a), _p(out), out.size)
    return bytes(out[:n])


def unpack8(data):
    b = np.frombuffer(bytes(data), np.uint8).copy()
    out = np.zeros(8, np.uint64)
    rem = C.c_int32(0)
    r = lib().fo_unpack8(_p(b), b.size, _p(out), C.byref(rem))
    return r, out, rem.value


def pack_doubles(vals):
    a = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.zeros(16 + 9 * a.size + 16, np.uint8)
    n = lib().fo_pack_doubles(_p(a), a.size, _p(out), out.size)
    assert n >= 0
    return bytes(out[:n])


def unpack_double_xor(data, n):
    b = np.frombuffer(bytes(data), np.uint8).copy()
    out = np.zeros(n, np.float64)
    r = lib().fo_unpack_double_xor(_p(b), b.size, _p(out), n)
    return r, out


def pack_delta(vals):
    a = np.ascontiguousarray(vals, dtype=np.int64)
    out = np.zeros(16 + 9 * a.size + 16, np.uint8)
    n = lib().fo_pack_delta(_p(a), a.size, _p(out), out.size)
    return bytes(out[:n])


def unpack_delta(data, n):
    b = np.frombuffer(bytes(data), np.uint8).copy()
    out = np.zeros(n, np.int64)
    r = lib().fo_unpack_delta(_p(b), b.size, _p(out), n)
    return r, out


def _enc(fn, a, *extra):
    cap = 64 + 9 * a.size
    out = np.zeros(cap, np.uint8)
    n = fn(_p(a), a.size, *extra, _p(out), cap)
    if n < 0:
        raise RuntimeError("encode failed: %s" % last_error())
    return out[:n].copy()


def encode_timestamps(ts):
    return _enc(lib().fo_encode_timestamps, np.ascontiguousarray(ts, dtype=np.int64))


def encode_longs(v):
    return _enc(lib().fo_encode_longs, np.ascontiguousarray(v, dtype=np.int64))


def encode_doubles(v, detect_drops=False, mode=VAL_OPTIMIZE):
    return _enc(lib().fo_encode_doubles, np.ascontiguousarray(v, dtype=np.float64), int(detect_drops), mode)


def encode_int_vector(v, nbits, signed):
    return _enc(lib().fo_encode_int_vector, np.ascontiguousarray(v, dtype=np.int32), nbits, int(signed))


def minmax_to_nbits(mn, mx):
    nb, sg = C.c_int32(), C.c_int32()
    lib().fo_minmax_to_nbits(mn, mx, C.byref(nb), C.byref(sg))
    return nb.value, bool(sg.value)


class Vec:
    """A BinaryVector held in a numpy byte buffer, with reader methods of the reference's VectorDataReaders."""

    def __init__(self, data):
        self.b = np.ascontiguousarray(np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data)
        self.p = _p(self.b)

    def total_bytes(self): return lib().fo_total_bytes(self.p)
    def vector_type(self): return lib().fo_vector_type(self.p)
    def dropped(self): return bool(lib().fo_double_dropped(self.p))
    # Long
    def long_length(self): return lib().fo_long_length(self.p)
    def long_apply(self, n): return lib().fo_long_apply(self.p, n)
    def binary_search(self, item): return lib().fo_long_binary_search(self.p, item)
    def ceiling_index(self, item): return lib().fo_long_ceiling_index(self.p, item)
    def long_sum(self, s, e): return lib().fo_long_sum(self.p, s, e)
    def longs(self): return [self.long_apply(i) for i in range(self.long_length())]
    # Int
    def int_length(self): return lib().fo_int_length(self.p)
    def int_apply(self, n): return lib().fo_int_apply(self.p, n)
    def int_sum(self, s, e): return lib().fo_int_sum(self.p, s, e)
    # Double
    def double_length(self): return lib().fo_double_length(self.p)
    def double_apply(self, n): return lib().fo_double_apply(self.p, n)
    def double_sum(self, s, e): return lib().fo_double_sum(self.p, s, e)
    def double_count(self, s, e): return lib().fo_double_count(self.p, s, e)
    def doubles(self): return [self.double_apply(i) for i in range(self.double_length())]

    def detect_drop(self, meta):
        m = np.array(meta_in(meta), np.float64); o = np.zeros(3)
        lib().fo_double_detect_drop(self.p, _p(m), _p(o))
        return meta_out(o)

    def update_correction(self, meta, force_corrected=True):
        m = np.array(meta_in(meta), np.float64); o = np.zeros(3)
        lib().fo_double_update_correction(self.p, _p(m), int(force_corrected), _p(o))
        return meta_out(o)

    def corrected_value(self, n, meta):
        m = np.array(meta_in(meta), np.float64)
        return lib().fo_double_corrected_value(self.p, n, _p(m))

    def drop_positions(self):
        out = np.zeros(4096, np.int32)
        n = lib().fo_double_drop_positions(self.p, _p(out), out.size)
        return list(out[:n])


def meta_in(meta):
    """meta: None (NoCorrection) or (lastValue, correction)"""
    return [0.0, 0.0, 0.0] if meta is None else [1.0, float(meta[0]), float(meta[1])]


def meta_out(o):
    return None if o[0] == 0 else (float(o[1]), float(o[2]))


def extrapolated_rate(ws, we, n, t1, v1, t2, v2, is_counter, is_rate):
    return lib().fo_extrapolated_rate(ws, we, n, t1, v1, t2, v2, int(is_counter), int(is_rate))


def num_windows(start, step, end):
    return lib().fo_num_windows(start, step, end)


# ---- store + query ------------------------------------------------------------------------------
class Store:
    """Series -> chunks with real ChunkSetInfo blocks (pointers inside), like a TimeSeriesPartition's chunk map."""

    def __init__(self):
        self.h = lib().fo_store_new()

    def __del__(self):
        try:
            lib().fo_store_free(self.h)
        except Exception:
            pass

    @property
    def num_series(self): return lib().fo_store_num_series(self.h)

    def add_series(self): return lib().fo_store_add_series(self.h)

    def add_chunk(self, series, ts, vals, val_mode=VAL_OPTIMIZE, detect_drops=False, ts_mode=TS_OPTIMIZE):
        ts = np.ascontiguousarray(ts, np.int64); vals = np.ascontiguousarray(vals, np.float64)
        assert ts.size == vals.size and ts.size > 0
        if lib().fo_store_add_chunk(self.h, series, _p(ts), _p(vals), ts.size, val_mode, int(detect_drops), ts_mode) != 0:
            raise RuntimeError(last_error())

    def add_chunk_longs(self, series, ts, vals, raw=False, ts_mode=TS_OPTIMIZE):
        """Long-column chunk: values through LongBinaryVector's appender + optimize() (DDV / const DDV / raw i64)."""
        ts = np.ascontiguousarray(ts, np.int64); vals = np.ascontiguousarray(vals, np.int64)
        assert ts.size == vals.size and ts.size > 0
        if lib().fo_store_add_chunk_longs(self.h, series, _p(ts), _p(vals), ts.size, int(raw), ts_mode) != 0:
            raise RuntimeError(last_error())

    def add_chunk_raw(self, series, start_time, end_time, num_rows, ts_bytes, val_bytes):
        t = np.ascontiguousarray(ts_bytes, np.uint8); v = np.ascontiguousarray(val_bytes, np.uint8)
        lib().fo_store_add_chunk_raw(self.h, series, start_time, end_time, num_rows, _p(t), t.size, _p(v), v.size)

    def add_series_rows(self, ts, vals, chunk_rows, val_mode=VAL_OPTIMIZE, detect_drops=False, ts_mode=TS_OPTIMIZE):
        ts = np.ascontiguousarray(ts, np.int64); vals = np.ascontiguousarray(vals, np.float64)
        cr = np.ascontiguousarray(chunk_rows, np.int32)
        assert cr.sum() == ts.size == vals.size
        if lib().fo_store_add_series_rows(self.h, _p(ts), _p(vals), ts.size, _p(cr), cr.size, val_mode, int(detect_drops), ts_mode) != 0:
            raise RuntimeError(last_error())
        return self.num_series - 1

    def add_from_arena(self, arena, rec_off, n_series):
        """arena: uint8 numpy array (kept alive by the store), rec_off: int64 offsets relative to arena[0]."""
        self._arena = arena; self._rec_off = np.ascontiguousarray(rec_off, np.int64)
        return lib().fo_store_add_from_arena(self.h, _p(arena), _p(self._rec_off), n_series)

    def add_synth(self, n_series, rows, rows_per_chunk=400, t0_ms=1_700_000_000_000, interval_ms=15000, ts_jitter_ms=0,
                  value_kind=0, value_enc=0, reset_period=0, nan_per_million=0, schema_flags=0, seed=42, series_id_base=0,
                  threads=1, **_ignored):
        """Same deterministic generator as the product's GPU generator (filo_synth_table), restated on the CPU."""
        st = np.sin(np.arange(1, rows + 1, dtype=np.float64))
        return lib().fo_store_add_synth(self.h, n_series, rows, rows_per_chunk, t0_ms, interval_ms, ts_jitter_ms, value_kind,
                                        value_enc, reset_period, nan_per_million, int(bool(schema_flags & 1)), seed,
                                        series_id_base, _p(st), threads)

    def num_chunks(self, series): return lib().fo_store_num_chunks(self.h, series)

    def info_addrs(self, series):
        out = np.zeros(self.num_chunks(series), np.uint64)
        lib().fo_store_info_addrs(self.h, series, _p(out))
        return out

    def all_info_addrs(self):
        """(n_chunks[int32 S], addrs[uint64 sum]) in the layout filo_load_series expects."""
        S = self.num_series
        nch = np.array([self.num_chunks(i) for i in range(S)], np.int32)
        addrs = np.concatenate([self.info_addrs(i) for i in range(S)]) if S else np.zeros(0, np.uint64)
        return nch, addrs

    def vector_bytes(self, series, chunk, col):
        out = np.zeros(1 << 16, np.uint8)
        n = lib().fo_store_vector_bytes(self.h, series, chunk, col, _p(out), out.size)
        if n < 0:
            out = np.zeros(-n, np.uint8)
            n = lib().fo_store_vector_bytes(self.h, series, chunk, col, _p(out), out.size)
        return out[:n].copy()

    def algorithmic_bytes(self): return lib().fo_store_algorithmic_bytes(self.h)

    def query(self, fn, start, step, end, window, cumulative=False, inclusive=True, aggr=AGG_NONE, k=0,
              group_ids=None, n_groups=1, threads=1, series_begin=0, series_end=-1, reuse_out=False, long_column=False, params=(0.0, 0.0)):
        S = (self.num_series if series_end < 0 else series_end) - series_begin
        T = num_windows(start, step, end)
        if aggr == AGG_NONE:
            if reuse_out and getattr(self, "_out", None) is not None and self._out.shape == (S, T):
                out = self._out        # timing loops: do not pay page faults of a fresh [S x T] buffer every call
            else:
                out = np.zeros((S, T), np.float64)
                if reuse_out: self._out = out
            aux = None
        elif aggr in (AGG_TOPK, AGG_BOTTOMK):
            out = np.zeros((n_groups, T, k), np.float64); aux = np.zeros((n_groups, T, k), np.int64)
        else:
            out = np.zeros((n_groups, T), np.float64)
            aux = np.zeros((n_groups, T), np.int64) if aggr == AGG_AVG else None
        g = np.ascontiguousarray(group_ids, np.int32) if group_ids is not None else None
        stats = np.zeros(3, np.int64)
        pr = tuple(params) + (0.0, 0.0)
        rc = lib().fo_query2(self.h, fn, int(bool(cumulative)) | (2 if long_column else 0), float(pr[0]), float(pr[1]), start, step, end, window,
                             int(inclusive), aggr, k, _p(g), n_groups, threads, series_begin, series_end, _p(out), _p(aux), _p(stats))
        if rc != 0:
            raise RuntimeError("oracle query failed: %s" % last_error())
        self.last_stats = {"samples_scanned": int(stats[0]), "bytes_scanned": int(stats[1]), "elapsed_ns": int(stats[2])}
        return (out, aux) if aux is not None else out


def sliding(ts, vals, fn, start, step, end, window, cumulative=False):
    ts = np.ascontiguousarray(ts, np.int64); vals = np.ascontiguousarray(vals, np.float64)
    out = np.zeros(num_windows(start, step, end), np.float64)
    lib().fo_sliding(_p(ts), _p(vals), ts.size, fn, int(cumulative), start, step, end, window, _p(out))
    return out


def synth_group_ids(seed, series_id_base, n, n_groups):
    out = np.zeros(n, np.int32)
    lib().fo_synth_group_ids(seed, series_id_base, n, n_groups, _p(out))
    return out


def serialize_result(values, start, step, end, now_ms=0):
    """SerializedRangeVector.apply over one shared RecordBuilder for the rows of `values` [n_rows, T] ->
    (containers uint8[n_containers, 4096], rows_serialized int32[n], start_record_no int32[n], first_container int64[n])."""
    v = np.ascontiguousarray(values, np.float64)
    n, T = v.shape
    cap = (n * T + 203) // 204 + 1
    out = np.zeros((cap, 4096), np.uint8)
    rs = np.zeros(n, np.int32); sr = np.zeros(n, np.int32); fc = np.zeros(n, np.int64)
    nc = lib().fo_serialize_result(_p(v), n, T, start, step, end, now_ms, _p(out), cap, _p(rs), _p(sr), _p(fc))
    if nc < 0:
        raise RuntimeError("fo_serialize_result: %s" % last_error())
    return out[:nc].copy(), rs, sr, fc


def result_rows(containers, rows_serialized, start_record_no, first_container, start, step, end):
    """SerializedRangeVector.rows of one range vector -> (ts int64[], values float64[])."""
    c = np.ascontiguousarray(containers, np.uint8)
    cap = max(int(rows_serialized), (end - start) // max(step, 1) + 1) + 1
    ts = np.zeros(cap, np.int64); vals = np.zeros(cap, np.float64)
    n = lib().fo_result_rows(_p(c), c.shape[0] if c.ndim == 2 else c.size // 4096, int(rows_serialized), int(start_record_no), int(first_container),
                             start, step, end, _p(ts), _p(vals), cap)
    return ts[:n].copy(), vals[:n].copy()
