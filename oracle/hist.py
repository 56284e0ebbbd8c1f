"""ORACLE (test infrastructure only): ctypes wrapper over oracle/hist_capi.cpp — FiloDB's histogram column path restated
on the CPU (filo_hist.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
import ctypes as C

import numpy as np

from .oracle import lib as _lib

GEOMETRIC, CUSTOM, EXP = 1, 2, 3
FMT_OTEL_DELTA = 9
ACK, INVALID_HISTOGRAM, BUCKET_SCHEMA_MISMATCH, VECTOR_TOO_SMALL = 0, 1, 2, 3
FMT_GEO_DELTA, FMT_GEO1_DELTA, FMT_CUSTOM_DELTA = 3, 4, 5
_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_ready = False


def lib():
    global _ready
    L = _lib()
    if not _ready:
        L.fo_hist_last_error.restype = C.c_char_p
        L.fo_hist_quantile.restype = _f64
        L.fo_hist_quantile.argtypes = [_i32, _f64, _f64, _i32, _vp, _i32, _vp, _f64]
        L.fo_hist_bucket_tops.restype = None; L.fo_hist_bucket_tops.argtypes = [_i32, _f64, _f64, _i32, _vp, _i32, _vp]
        L.fo_hist_serialize_buckets.argtypes = [_i32, _f64, _f64, _i32, _vp, _i32, _vp, _i32]
        L.fo_hist_parse_buckets.argtypes = [_vp, _i32, _vp, _i32]
        L.fo_hist_write_delta.argtypes = [_i32, _f64, _f64, _i32, _vp, _i32, _vp, _vp, _i32]
        L.fo_hist_blob_to_values.argtypes = [_vp, _vp, _i32]
        L.fo_hist_make_monotonic.restype = None; L.fo_hist_make_monotonic.argtypes = [_vp, _i32]
        L.fo_hist_appender_new.restype = _vp; L.fo_hist_appender_new.argtypes = [_i32, _i32]
        L.fo_hist_appender_free.restype = None; L.fo_hist_appender_free.argtypes = [_vp]
        L.fo_hist_appender_add.argtypes = [_vp, _vp, _i32]
        L.fo_hist_appender_length.argtypes = [_vp]
        L.fo_hist_appender_bytes.argtypes = [_vp, _vp, _i32]
        L.fo_hist_vec_info.argtypes = [_vp, _vp, _vp, _vp]
        L.fo_hist_vec_apply.argtypes = [_vp, _i32, _vp]
        L.fo_hist_vec_section_types.argtypes = [_vp, _vp, _i32]
        L.fo_hist_vec_sum.argtypes = [_vp, _i32, _i32, _vp]
        L.fo_hist_vec_detect_drop.argtypes = [_vp, _i32, _vp, _vp]
        L.fo_hist_vec_update_correction.argtypes = [_vp, _i32, _vp, _vp, _vp]
        L.fo_hist_vec_corrected.argtypes = [_vp, _i32, _i32, _vp, _vp]
        L.fo_hist_vec_apply_exp.argtypes = [_vp, _i32, _vp, _vp, _i32]
        L.fo_hist_vec_sum_exp.argtypes = [_vp, _i32, _i32, _vp, _vp, _i32]
        L.fo_hist_exp_add_scheme.argtypes = [_vp, _vp, _i32, _vp]
        L.fo_hist_exp_add_values.argtypes = [_vp, _vp, _vp, _vp]
        L.fo_hist_exp_add_no_correction.argtypes = [_vp, _vp, _vp, _vp, _vp, _i32]
        L.fo_hstore_new.restype = _vp
        L.fo_hstore_free.restype = None; L.fo_hstore_free.argtypes = [_vp]
        L.fo_hstore_add_series.restype = _i64; L.fo_hstore_add_series.argtypes = [_vp]
        L.fo_hstore_add_chunk.argtypes = [_vp, _i64, _vp, _i32, _i32, _f64, _f64, _i32, _vp, _i32, _vp, _i32, _i32]
        L.fo_hstore_num_chunks.restype = _i64; L.fo_hstore_num_chunks.argtypes = [_vp, _i64]
        L.fo_hstore_info_addrs.restype = None; L.fo_hstore_info_addrs.argtypes = [_vp, _i64, _vp]
        L.fo_hstore_vector_bytes.argtypes = [_vp, _i64, _i64, _vp, _i32]
        L.fo_hstore_query.argtypes = [_vp, _i32, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _i32, _i32, _f64, _vp, _vp, _vp]
        _ready = True
    return L


def _p(a):
    return None if a is None else a.ctypes.data


def _err():
    return lib().fo_hist_last_error().decode()


class Buckets:
    """GeometricBuckets(first, multiplier, n, minusOne) or CustomBuckets(les) (Histogram.scala:601-626, 874-899)."""

    def __init__(self, kind, n, first=0.0, mult=0.0, minus_one=False, les=None):
        self.kind, self.n, self.first, self.mult, self.minus_one = kind, n, float(first), float(mult), bool(minus_one)
        self.les = np.ascontiguousarray(les, np.float64) if les is not None else None

    @staticmethod
    def geometric(first, mult, n, minus_one=False):
        return Buckets(GEOMETRIC, n, first, mult, minus_one)

    @staticmethod
    def exponential(scale, start_index, num_positive):
        """Base2ExpHistogramBuckets(scale, startIndexPositiveBuckets, numPositiveBuckets) (Histogram.scala:684-866)."""
        return Buckets(EXP, num_positive + 1, float(scale), float(start_index))

    @property
    def scheme(self):
        return np.array([int(self.first), int(self.mult), self.n], np.int32)

    def can_accommodate(self, other):
        out = np.zeros(3, np.int32); a, b = self.scheme, other.scheme
        r = lib().fo_hist_exp_add_scheme(_p(a), _p(b), 180, _p(out))
        if r < 0: raise RuntimeError(_err())
        return bool(r & 1)

    def add(self, other, max_pos=180):
        out = np.zeros(3, np.int32); a, b = self.scheme, other.scheme
        if lib().fo_hist_exp_add_scheme(_p(a), _p(b), max_pos, _p(out)) < 0: raise RuntimeError(_err())
        return Buckets.exponential(int(out[0]), int(out[1]), int(out[2]) - 1)

    def add_values(self, our_values, other, other_values):
        v = np.array(our_values, np.float64); o = np.ascontiguousarray(other_values, np.float64); a, b = self.scheme, other.scheme
        if lib().fo_hist_exp_add_values(_p(a), _p(v), _p(b), _p(o)) != 0: raise RuntimeError(_err())
        return v

    def add_no_correction(self, values, other, other_values):
        """MutableHistogram(self, values).addNoCorrection(MutableHistogram(other, other_values)) -> (buckets, values)."""
        sch = self.scheme.copy(); a = np.ascontiguousarray(values, np.float64); o = np.ascontiguousarray(other_values, np.float64)
        out = np.zeros(512, np.float64); b = other.scheme
        n = lib().fo_hist_exp_add_no_correction(_p(sch), _p(a), _p(b), _p(o), _p(out), out.size)
        if n < 0: raise RuntimeError(_err())
        return Buckets.exponential(int(sch[0]), int(sch[1]), int(sch[2]) - 1), out[:n].copy()

    @staticmethod
    def custom(les):
        les = np.ascontiguousarray(les, np.float64)
        return Buckets(CUSTOM, les.size, les=les)

    def args(self):
        return (self.kind, self.first, self.mult, int(self.minus_one), _p(self.les), self.n)

    def tops(self):
        out = np.zeros(self.n, np.float64)
        lib().fo_hist_bucket_tops(*self.args(), _p(out))
        return out

    def serialize(self):
        out = np.zeros(65536, np.uint8)
        n = lib().fo_hist_serialize_buckets(*self.args(), _p(out), out.size)
        if n < 0: raise RuntimeError(_err())
        return out[:n].copy()

    def write_delta(self, values):
        """BinaryHistogram.writeDelta(buckets, values) -> blob bytes (HistogramVector.scala:171-198)."""
        v = np.ascontiguousarray(values, np.int64)
        out = np.zeros(65536 + 64, np.uint8)
        n = lib().fo_hist_write_delta(*self.args(), _p(v), _p(out), out.size)
        if n < 0: raise RuntimeError(_err())
        return out[:n].copy()

    def quantile(self, values, q):
        v = np.ascontiguousarray(values, np.float64)
        return float(lib().fo_hist_quantile(*self.args(), _p(v), float(q)))


def parse_buckets(defbytes, format_code):
    d = np.ascontiguousarray(defbytes, np.uint8)
    tops = np.zeros(8192, np.float64)
    n = lib().fo_hist_parse_buckets(_p(d), format_code, _p(tops), tops.size)
    if n < 0: raise RuntimeError(_err())
    return tops[:n].copy()


def blob_to_values(blob):
    b = np.ascontiguousarray(blob, np.uint8)
    out = np.zeros(8192, np.int64)
    n = lib().fo_hist_blob_to_values(_p(b), _p(out), out.size)
    if n < 0: raise RuntimeError(_err())
    return out[:n].copy()


def make_monotonic(values):
    v = np.array(values, np.float64)
    lib().fo_hist_make_monotonic(_p(v), v.size)
    return v


class Appender:
    """AppendableHistogramVector / AppendableSectDeltaHistVector (HistogramVector.scala:326-436, 489-545);
    sect=2: AppendableExpHistogramVector (ExpHistogramVector.scala:37-121)."""

    def __init__(self, sect, max_bytes):
        self.h = lib().fo_hist_appender_new(int(sect), max_bytes)

    def add(self, blob):
        b = np.ascontiguousarray(blob, np.uint8)
        r = lib().fo_hist_appender_add(self.h, _p(b), b.size)
        if r < 0: raise RuntimeError(_err())
        return r

    @property
    def length(self):
        return lib().fo_hist_appender_length(self.h)

    def bytes(self):
        out = np.zeros(1 << 20, np.uint8)
        n = lib().fo_hist_appender_bytes(self.h, _p(out), out.size)
        if n < 0: raise RuntimeError("vector larger than 1 MiB")
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None): lib().fo_hist_appender_free(self.h); self.h = None


class Reader:
    """RowHistogramReader / SectDeltaHistogramReader over vector bytes (HistogramVector.scala:557-738)."""

    def __init__(self, vec):
        self.vec = np.ascontiguousarray(vec, np.uint8)
        ln, nb, sect = _i32(), _i32(), _i32()
        if lib().fo_hist_vec_info(_p(self.vec), C.byref(ln), C.byref(nb), C.byref(sect)) != 0: raise RuntimeError(_err())
        self.length, self.num_buckets, self.sect = ln.value, nb.value, bool(sect.value)

    def __call__(self, i):
        out = np.zeros(self.num_buckets, np.int64)
        if lib().fo_hist_vec_apply(_p(self.vec), i, _p(out)) != 0: raise RuntimeError(_err())
        return out

    def section_types(self):
        out = np.zeros(4096, np.int32)
        n = lib().fo_hist_vec_section_types(_p(self.vec), _p(out), out.size)
        if n < 0: raise RuntimeError(_err())
        return out[:n].tolist()

    def apply_exp(self, i):
        """RowExpHistogramReader.apply -> ((scale, start, numPositive), values) (ExpHistogramVector.scala:172-189)."""
        sch = np.zeros(3, np.int32); out = np.zeros(256, np.int64)
        n = lib().fo_hist_vec_apply_exp(_p(self.vec), i, _p(sch), _p(out), out.size)
        if n < 0: raise RuntimeError(_err())
        return (int(sch[0]), int(sch[1]), int(sch[2]) - 1), out[:n].copy()

    def sum_exp(self, start, end):
        sch = np.zeros(3, np.int32); out = np.zeros(256, np.float64)
        n = lib().fo_hist_vec_sum_exp(_p(self.vec), start, end, _p(sch), _p(out), out.size)
        if n < 0: raise RuntimeError(_err())
        return (int(sch[0]), int(sch[1]), int(sch[2]) - 1), out[:n].copy()

    def sum(self, start, end):
        out = np.zeros(self.num_buckets, np.float64)
        if lib().fo_hist_vec_sum(_p(self.vec), start, end, _p(out)) != 0: raise RuntimeError(_err())
        return out

    def detect_drop(self, last=None, correction=None):
        """detectDropAndCorrection(meta) -> correction (None for NoCorrection)."""
        if last is None: return None
        l = np.ascontiguousarray(last, np.int64); c = np.array(correction, np.int64)
        if lib().fo_hist_vec_detect_drop(_p(self.vec), 1, _p(l), _p(c)) != 0: raise RuntimeError(_err())
        return c

    def update_correction(self, correction=None):
        """updateCorrection(meta) -> (lastValue, correction)."""
        c = None if correction is None else np.ascontiguousarray(correction, np.int64)
        ol, oc = np.zeros(self.num_buckets, np.int64), np.zeros(self.num_buckets, np.int64)
        if lib().fo_hist_vec_update_correction(_p(self.vec), int(c is not None), _p(c), _p(ol), _p(oc)) != 0: raise RuntimeError(_err())
        return ol, oc

    def corrected(self, n, correction=None):
        c = None if correction is None else np.ascontiguousarray(correction, np.int64)
        out = np.zeros(self.num_buckets, np.int64)
        if lib().fo_hist_vec_corrected(_p(self.vec), n, int(c is not None), _p(c), _p(out)) != 0: raise RuntimeError(_err())
        return out


class HistStore:
    """Histogram time series (timestamps + one histogram column), queried through the reference path restatement."""

    def __init__(self, buckets):
        self.b = buckets
        self.h = lib().fo_hstore_new()
        self.num_series = 0

    def add_series(self, ts, values, chunks, sect=True, max_bytes=15000):
        """values: int64 [rows, nb] cumulative bucket counts; chunks: rows per chunk."""
        si = lib().fo_hstore_add_series(self.h); self.num_series += 1
        ts = np.ascontiguousarray(ts, np.int64); v = np.ascontiguousarray(values, np.int64)
        off = 0
        for r in chunks:
            if lib().fo_hstore_add_chunk(self.h, si, _p(ts[off:off + r]), r, *self.b.args()[:5], self.b.n, _p(v[off:off + r]), int(sect), max_bytes) != 0:
                raise RuntimeError(_err())
            off += r
        return si

    def all_info_addrs(self):
        nch = np.array([lib().fo_hstore_num_chunks(self.h, i) for i in range(self.num_series)], np.int32)
        addrs = np.zeros(int(nch.sum()), np.uint64)
        off = 0
        for i in range(self.num_series):
            lib().fo_hstore_info_addrs(self.h, i, addrs[off:].ctypes.data); off += int(nch[i])
        return nch, addrs

    def vector_bytes(self, series, chunk):
        out = np.zeros(1 << 20, np.uint8)
        n = lib().fo_hstore_vector_bytes(self.h, series, chunk, _p(out), out.size)
        return out[:n].copy()

    def query(self, fn, start, step, end, window, cumulative=True, inclusive=True, aggr=False, group_ids=None, n_groups=1, q=float("nan")):
        """-> (values [rows, T, nb], empty [rows, T] bool[, quantile [rows, T]]) with rows = series or groups."""
        from .oracle import num_windows
        T = num_windows(start, step, end); nb = self.b.n
        rows = n_groups if aggr else self.num_series
        vals = np.zeros((rows, T, nb), np.float64); empty = np.zeros((rows, T), np.uint8)
        qs = np.zeros((rows, T), np.float64) if aggr else None
        g = np.ascontiguousarray(group_ids, np.int32) if group_ids is not None else None
        if lib().fo_hstore_query(self.h, fn, int(cumulative), start, step, end, window, int(inclusive), int(aggr), _p(g), n_groups, nb, float(q),
                                 _p(vals), _p(empty), _p(qs)) != 0:
            raise RuntimeError(_err())
        return (vals, empty.astype(bool), qs) if aggr else (vals, empty.astype(bool))

    def __del__(self):
        if getattr(self, "h", None): lib().fo_hstore_free(self.h); self.h = None
