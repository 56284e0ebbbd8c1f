"""Pins the histogram restatement (oracle/filo_hist.hpp) to the reference's own known answers:
core/src/test/scala/filodb.memory/format/vectors/HistogramTest.scala and HistogramVectorTest.scala."""
import numpy as np
import pytest

from oracle import hist as H

RAW = np.array([[10, 15, 17, 20, 25, 34, 76, 82],
                [6, 16, 26, 26, 36, 38, 56, 59],
                [11, 16, 26, 27, 33, 42, 46, 55],
                [4, 4, 5, 33, 35, 67, 91, 121]], np.int64)                       # HistogramTest.scala:8-13
INCR = np.cumsum(RAW, axis=0)                                                    # incrHistBuckets, :36-38
LAST_INCR = INCR[-1]
SCHEME = H.Buckets.geometric(1.0, 2.0, 8)                                        # bucketScheme, :6
CUSTOM = H.Buckets.custom([0.25, 0.5, 1.0, 2.5, 5.0, 10, float("inf")])          # customScheme, :7
CORR1 = np.array([1, 2, 3, 4, 5, 6, 7, 8], np.int64)                             # correction1, :49
CORR2 = np.array([2, 4, 6, 8, 10, 12, 14, 18], np.int64)                         # correction2, :50


def test_bucket_tops_and_serialization():
    """HistogramTest.scala:61-90."""
    assert H.Buckets.geometric(5.0, 3.0, 4).tops().tolist() == [5.0, 15.0, 45.0, 135.0]
    assert H.Buckets.geometric(2.0, 2.0, 8, minus_one=True).tops().tolist() == [1.0, 3.0, 7.0, 15.0, 31.0, 63.0, 127.0, 255.0]
    b1 = H.Buckets.geometric(5.0, 2.0, 4)
    assert len(b1.serialize()) == 2 + 2 + 8 + 8
    assert H.parse_buckets(b1.serialize(), H.FMT_GEO_DELTA).tolist() == b1.tops().tolist()
    b2 = H.Buckets.geometric(2.0, 2.0, 8, minus_one=True)
    assert len(b2.serialize()) == 20
    assert H.parse_buckets(b2.serialize(), H.FMT_GEO1_DELTA).tolist() == b2.tops().tolist()
    assert len(CUSTOM.serialize()) == 26
    assert H.parse_buckets(CUSTOM.serialize(), H.FMT_CUSTOM_DELTA).tolist() == CUSTOM.tops().tolist()


def test_quantile_known_answers():
    """HistogramTest.scala:52,109-120: exact shouldEqual values."""
    expected = [37.333333333333336, 10.8, 8.666666666666666, 28.75]
    for v, e in zip(RAW, expected):
        assert SCHEME.quantile(v.astype(float), 0.50) == e
    # cannot return more than the 2nd-to-last bucket top when the last bucket is +Inf
    assert CUSTOM.quantile(RAW[0][:7].astype(float), 0.95) == 10
    assert SCHEME.quantile(RAW[0].astype(float), -0.1) == float("-inf")
    assert SCHEME.quantile(RAW[0].astype(float), 1.1) == float("inf")
    assert np.isnan(SCHEME.quantile(np.zeros(8), 0.5))


def test_make_monotonic_with_nans():
    """HistogramTest.scala:480-495 shape: NaN and decreasing values take the running max."""
    nan = float("nan")
    assert H.make_monotonic([0, 3, 2, nan, 5, 4]).tolist() == [0, 3, 3, 3, 5, 5]
    assert H.make_monotonic([nan, nan, 1]).tolist() == [0, 0, 1]


def test_binary_histogram_blob_round_trip():
    """BinaryHistogram.writeDelta / toHistogram (HistogramVector.scala:84-96,171-198)."""
    for v in RAW:
        blob = SCHEME.write_delta(v)
        assert int(blob[0]) | (int(blob[1]) << 8) == len(blob) - 2 and blob[2] == H.FMT_GEO_DELTA
        assert int(blob[3]) | (int(blob[4]) << 8) == 18 and int(blob[5]) | (int(blob[6]) << 8) == 8
        assert H.blob_to_values(blob).tolist() == v.tolist()
    blob = CUSTOM.write_delta(RAW[1][:7])
    assert blob[2] == H.FMT_CUSTOM_DELTA and H.blob_to_values(blob).tolist() == RAW[1][:7].tolist()


def test_simple_vector_append_and_read():
    """HistogramVectorTest.scala:36-76."""
    app = H.Appender(False, 1024)
    for v in RAW:
        assert app.add(SCHEME.write_delta(v)) == H.ACK
    assert app.length == 4
    r = H.Reader(app.bytes())
    assert (r.length, r.num_buckets, r.sect) == (4, 8, False)
    for i in range(4):
        assert r(i).tolist() == RAW[i].tolist()
    app2 = H.Appender(False, 1024)
    for v in RAW:
        assert app2.add(CUSTOM.write_delta(v[:7])) == H.ACK
    r2 = H.Reader(app2.bytes())
    for i in range(4):
        assert r2(i).tolist() == RAW[i][:7].tolist()
    # sum(start, end): NaN-seeded MutableHistogram, addNoCorrection per row (:613-621)
    assert r.sum(0, 3).tolist() == RAW.sum(axis=0).astype(float).tolist()
    assert r.sum(1, 2).tolist() == (RAW[1] + RAW[2]).astype(float).tolist()


def test_schema_mismatch_and_invalid_blobs():
    """HistogramVectorTest.scala:367-381 (BucketSchemaMismatch) and addData validation (:366-371)."""
    app = H.Appender(False, 1024)
    assert app.add(SCHEME.write_delta(RAW[0])) == H.ACK
    assert app.add(H.Buckets.geometric(1.0, 2.0, 8, minus_one=True).write_delta(RAW[1])) == H.BUCKET_SCHEMA_MISMATCH
    assert app.add(CUSTOM.write_delta(RAW[1][:7])) == H.BUCKET_SCHEMA_MISMATCH
    assert app.add(np.zeros(3, np.uint8)) == H.INVALID_HISTOGRAM
    assert app.length == 1


def test_vector_too_small():
    app = H.Appender(True, 64)
    res = [app.add(SCHEME.write_delta(v)) for v in INCR]
    assert res[0] == H.ACK and H.VECTOR_TOO_SMALL in res


def test_sectdelta_append_read_and_update_correction():
    """HistogramVectorTest.scala:243-273."""
    app = H.Appender(True, 1024)
    for v in INCR:
        assert app.add(SCHEME.write_delta(v)) == H.ACK
    assert app.length == 4
    r = H.Reader(app.bytes())
    assert r.sect and r.length == 4
    for i in range(4):
        assert r(i).tolist() == INCR[i].tolist()
    last, corr = r.update_correction(None)
    assert last.tolist() == LAST_INCR.tolist() and corr.tolist() == [0] * 8
    last, corr = r.update_correction(CORR1)
    assert last.tolist() == LAST_INCR.tolist() and corr.tolist() == CORR1.tolist()


def test_sectdelta_detects_drops():
    """HistogramVectorTest.scala:333-362: one normal section, one drop section; corrections propagate."""
    app = H.Appender(True, 1024)
    for v in list(INCR) + list(INCR):
        assert app.add(SCHEME.write_delta(v)) == H.ACK
    assert app.length == 8
    r = H.Reader(app.bytes())
    assert r.section_types() == [0, 1]
    for i in range(4):
        assert r(i).tolist() == INCR[i].tolist() and r(4 + i).tolist() == INCR[i].tolist()
    last, corr = r.update_correction(None)
    assert last.tolist() == LAST_INCR.tolist() and corr.tolist() == LAST_INCR.tolist()
    last, corr = r.update_correction(CORR1)
    assert last.tolist() == LAST_INCR.tolist() and corr.tolist() == (CORR1 + LAST_INCR).tolist()


def test_detect_drop_at_chunk_start_and_corrected_value():
    """HistogramVectorTest.scala:442-487."""
    app = H.Appender(True, 1024)
    for v in INCR:
        app.add(SCHEME.write_delta(v))
    r = H.Reader(app.bytes())
    assert r.detect_drop(None) is None                                           # NoCorrection passes through
    assert r.detect_drop(CORR1, CORR2).tolist() == CORR2.tolist()                # first value >= last: unchanged
    assert r.detect_drop(LAST_INCR, CORR2).tolist() == (CORR2 + LAST_INCR).tolist()   # drop: correction += lastValue
    assert r.corrected(1, None).tolist() == INCR[1].tolist()
    assert r.corrected(1, CORR2).tolist() == (CORR2 + INCR[1]).tolist()
    app2 = H.Appender(True, 1024)
    for v in INCR:
        app2.add(SCHEME.write_delta(v))
    for v in INCR:
        app2.add(SCHEME.write_delta(v + 15))
    r2 = H.Reader(app2.bytes())
    incr5 = INCR[1] + 15 + LAST_INCR
    assert r2.corrected(5, None).tolist() == incr5.tolist()
    assert r2.corrected(5, CORR2).tolist() == (incr5 + CORR2).tolist()


def test_sections_roll_over_every_16_histograms():
    """AppendableSectDeltaHistVector.maxElementsPerSection = 16 (HistogramVector.scala:501): values survive section changes."""
    rng = np.random.default_rng(3)
    rows = np.cumsum(np.cumsum(rng.integers(0, 50, (70, 8)), axis=1), axis=0).astype(np.int64)
    app = H.Appender(True, 15000)
    for v in rows:
        assert app.add(SCHEME.write_delta(v)) == H.ACK
    r = H.Reader(app.bytes())
    assert r.section_types() == [0, 0, 0, 0, 0]
    for i in (0, 1, 15, 16, 17, 40, 69):
        assert r(i).tolist() == rows[i].tolist()


def _hist_rows(n, nb, rng, reset_at=None):
    inc = np.cumsum(rng.integers(0, 20, (n, nb)), axis=1)
    rows = np.cumsum(inc, axis=0).astype(np.int64)
    if reset_at is not None:
        rows[reset_at:] = np.cumsum(inc[reset_at:], axis=0)
    return rows


def test_hist_rate_matches_the_reference_expectation():
    """RateFunctionsSpec.scala:266-296: one window over 7 samples: rate(b) = (last_b - head_b) / (lastTime - headTime) * 1000."""
    rng = np.random.default_rng(11)
    rows = _hist_rows(10, 8, rng)
    ts = 100000 + np.arange(10, dtype=np.int64) * 10000
    st = H.HistStore(SCHEME)
    st.add_series(ts, rows, [10])
    from oracle import oracle as o
    start_ts, end_ts = 99500, 161000
    vals, empty = st.query(o.FN_RATE, end_ts, 100000, end_ts, end_ts - start_ts)
    assert vals.shape == (1, 1, 8) and not empty[0, 0]
    exp = (rows[6] - rows[0]).astype(float) / (160000 - 100000) * 1000
    np.testing.assert_allclose(vals[0, 0], exp, rtol=0, atol=1e-5)               # errorOk of the reference test


def test_hist_rate_with_drop_matches_the_reference_expectation():
    """RateFunctionsSpec.scala:298-331: the 8th sample is the first one again (a drop): corrected by the 7th."""
    rng = np.random.default_rng(12)
    rows7 = _hist_rows(7, 8, rng)
    rows = np.concatenate([rows7, rows7])
    ts = 100000 + np.arange(14, dtype=np.int64) * 10000
    st = H.HistStore(SCHEME)
    st.add_series(ts, rows, [14])
    from oracle import oracle as o
    start_ts, end_ts = 99500, 171000
    vals, empty = st.query(o.FN_RATE, end_ts, 110000, end_ts, end_ts - start_ts)
    last = rows7[0] + rows7[6]
    exp = (last - rows7[0]).astype(float) / (170000 - 100000) * 1000
    np.testing.assert_allclose(vals[0, 0], exp, rtol=0, atol=1e-5)


def test_hist_sum_aggregate_and_quantile():
    """sum(rate(h[..])) through HistSumRowAggregator then histogram_quantile: agrees with the per-series results folded in order."""
    from oracle import oracle as o
    rng = np.random.default_rng(13)
    t0 = 1_700_000_000_000
    ts = t0 + np.arange(120, dtype=np.int64) * 15000
    les = [2.0 * 3 ** i for i in range(9)] + [float("inf")]
    b = H.Buckets.custom(les)
    st = H.HistStore(b)
    for s in range(6):
        st.add_series(ts, _hist_rows(120, 10, rng, reset_at=70 if s == 2 else None), [80, 40])
    start, step, end, window = t0 + 300000, 60000, t0 + 119 * 15000, 300000
    per, empty = st.query(o.FN_RATE, start, step, end, window)
    assert not empty.any()
    agg, aempty, qs = st.query(o.FN_RATE, start, step, end, window, aggr=True, group_ids=[0, 1, 0, 1, 0, 1], n_groups=2, q=0.99)
    for g in (0, 1):
        members = [s for s in range(6) if s % 2 == g]
        for k in range(per.shape[1]):
            acc = per[members[0], k].copy()
            for s in members[1:]:
                acc = H.make_monotonic(acc + per[s, k])
            assert acc.tolist() == agg[g, k].tolist()
            assert qs[g, k] == b.quantile(agg[g, k], 0.99)
    assert np.isfinite(qs).all()
