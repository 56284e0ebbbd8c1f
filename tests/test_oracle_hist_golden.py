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


# ------------------------------------------------------------------ otel exponential buckets (Base2ExpHistogramBuckets)
def test_exp_buckets_serialize_tops_and_index_mapping():
    """HistogramTest.scala:76-84 (18-byte definition incl. start indexes beyond i16), :497-520 (tops), :545-551."""
    for b in (H.Buckets.exponential(3, -5, 16), H.Buckets.exponential(3, -9037032, 150)):
        d = b.serialize()
        assert len(d) == 18
        assert H.parse_buckets(d, H.FMT_OTEL_DELTA).tolist() == b.tops().tolist()
    b1 = H.Buckets.exponential(3, -5, 11)                                         # 0.707 .. 1.68
    t = b1.tops()
    assert len(t) == 12 and t[0] == 0.0
    assert t[1] == pytest.approx(0.7071067811865475, abs=1e-4) and t[-1] == pytest.approx(1.6817928305074294, abs=1e-4)
    assert t[-5 - -5 + 1 + 4] == 1.0                                              # bucketTop(bucketIndexToArrayIndex(-1)) shouldEqual 1.0
    rng = np.random.default_rng(5)
    for _ in range(200):
        scale, start, n = int(rng.integers(-20, 21)), int(rng.integers(-100000, 100001)), int(rng.integers(1, 181))
        with np.errstate(over="ignore", under="ignore"):
            base = np.float64(2.0) ** (np.float64(2.0) ** -scale)
            if not (base ** (start + 1) < 1e10 and base ** (start + n) < 1e10): continue
        tops = H.Buckets.exponential(scale, start, n).tops()
        for i in range(1, n + 1):
            assert tops[i] == pytest.approx(base ** (start + i), abs=1e-6)


def test_exp_quantile_known_answers():
    """HistogramTest.scala:120-131."""
    b = H.Buckets.exponential(3, -5, 11)
    v = np.arange(1, 13, dtype=np.float64)
    assert b.quantile(v, 0.5) == pytest.approx(1.0, abs=1e-5)
    assert b.quantile(v, 0.75) == pytest.approx(1.2968395546510099, abs=1e-5)
    assert b.quantile(v, 0.25) == pytest.approx(0.7711054127039704, abs=1e-5)
    assert b.quantile(v, 0.99) == pytest.approx(1.6643974694230492, abs=1e-5)
    assert b.quantile(v, 0.01) == 0.0                                             # zero bucket
    assert b.quantile(v, 0.085) == pytest.approx(0.014142135623730961, abs=1e-5)


def test_exp_scheme_add_and_add_values():
    """HistogramTest.scala:553-606, :623-642."""
    E = H.Buckets.exponential
    b1, b2, b3 = E(3, -5, 11), E(2, -2, 6), E(2, -4, 8)
    assert b2.tops()[-1] == pytest.approx(1.9999999999999998, abs=1e-4)
    assert not b1.can_accommodate(b2) and not b2.can_accommodate(b1)
    assert b3.can_accommodate(b1) and b3.can_accommodate(b2)
    badd = b1.add(b2)
    assert badd.n == 9 and int(badd.first) == 2
    assert badd.tops()[1] == pytest.approx(0.5946035575013606, abs=1e-4) and badd.tops()[-1] == pytest.approx(1.9999999999999998, abs=1e-4)
    assert badd.can_accommodate(b1) and badd.can_accommodate(b2)
    v = badd.add_values(np.zeros(9), b1, np.arange(12.0))
    assert v.tolist() == [0.0, 0.0, 1.0, 3.0, 5.0, 7.0, 9.0, 11.0, 11.0]
    v = badd.add_values(v, b2, np.arange(7.0))
    assert v.tolist() == [0.0, 0.0, 1.0, 4.0, 7.0, 10.0, 13.0, 16.0, 17.0]
    b4, b5 = E(5, 15, 36), E(3, 10, 6)
    assert b4.n == 37 and b4.tops()[1] == pytest.approx(1.414213562373094, abs=1e-4) and b4.tops()[-1] == pytest.approx(3.0183288551868377, abs=1e-4)
    assert b5.n == 7 and b5.tops()[1] == pytest.approx(2.59367910930202, abs=1e-4) and b5.tops()[-1] == pytest.approx(4.000000000000002, abs=1e-4)
    badd2 = badd.add(b4).add(b5)
    v2 = badd2.add_values(np.zeros(badd2.n), badd, v)
    assert v2.tolist() == [0.0, 0.0, 1.0, 4.0, 7.0, 10.0, 13.0, 16.0, 17.0, 17.0, 17.0, 17.0, 17.0, 17.0]
    v2 = badd2.add_values(v2, b4, np.arange(37.0))
    assert v2.tolist() == [0.0, 0.0, 1.0, 4.0, 7.0, 10.0, 14.0, 25.0, 34.0, 42.0, 50.0, 53.0, 53.0, 53.0]
    v2 = badd2.add_values(v2, b5, [0.0, 10.0, 11, 12, 13, 14, 15])
    assert v2.tolist() == [0.0, 0.0, 1.0, 4.0, 7.0, 10.0, 14.0, 25.0, 34.0, 42.0, 61.0, 66.0, 68.0, 68.0]
    # non-overlapping ranges of one scale; scale reduction under a bucket budget
    nb = E(3, -5, 11).add(E(3, 15, 11))
    assert nb.n == 32 and int(nb.mult) == -5
    a1 = E(6, -50, 21).add(E(6, 100, 26), max_pos=128)
    assert a1.scheme.tolist() == [5, -26, 91]
    assert a1.can_accommodate(E(6, -50, 21)) and a1.can_accommodate(E(6, 100, 26))
    a2 = E(6, -50, 21).add(E(6, 100, 26), max_pos=64)
    assert a2.scheme.tolist() == [4, -14, 47]


def test_exp_add_zero_only_histograms():
    """HistogramTest.scala:608-621."""
    b1, b2 = H.Buckets.exponential(20, 10, 0), H.Buckets.exponential(3, 10, 6)
    m2 = [1.0, 10.0, 11, 12, 13, 14, 15]
    rb, rv = b1.add_no_correction([1.0], b2, m2)
    assert rb.scheme.tolist() == b2.scheme.tolist() and rv.tolist() == [2.0, 10.0, 11, 12, 13, 14, 15]
    rb, rv = b2.add_no_correction(m2, b1, [1.0])
    assert rb.scheme.tolist() == b2.scheme.tolist() and rv.tolist() == [2.0, 10.0, 11, 12, 13, 14, 15]


# cumulative counts of the "real data" histogram of HistogramTest.scala:135-166 / ExpHistogramVectorTest.scala:64-95, scheme (3, -78, 126)
REAL_COUNTS = np.array([0] * 55 + [1] * 7 + [2, 2, 3, 3, 3, 3, 4, 5, 5, 5, 6, 6, 8, 8, 9, 9, 11, 12, 14, 15, 17, 19, 20, 22, 23, 26, 28, 31, 34, 37,
                                             41, 45, 48, 53, 58, 64, 70, 76, 84, 90, 99, 108, 118, 129, 140, 152, 167, 182, 199, 217, 237, 258,
                                             282, 308, 336, 367, 400, 435, 474, 517, 565, 617, 672, 732, 749], np.int64)


def test_exp_quantile_real_data():
    """HistogramTest.scala:133-176 (quantiles with the default min / max)."""
    b = H.Buckets.exponential(3, -78, 126)
    t = b.tops()
    assert t[1] == pytest.approx(0.0012664448775888738, rel=1e-12) and t[-1] == pytest.approx(64.00000000000009, rel=1e-12)
    v = REAL_COUNTS.astype(np.float64)
    assert b.quantile(v, 0.5) == pytest.approx(29.927691427444305, abs=1e-5)
    assert b.quantile(v, 0.99) == pytest.approx(61.602904581469566, abs=1e-5)
    assert b.quantile(v, 0.01) == pytest.approx(0.6916552392692796, abs=1e-5)


OTEL_EXP = [((3, -3, 1), [0, 3]), ((20, -3, 9), [0, 4, 5, 6, 7, 8, 9, 10, 11, 12]), ((20, -888388, 1), [0, 5])]   # ExpHistogramVectorTest.scala:35-39
EXP_VECTOR_HEX = ("5C00000009130300540003001800160009100002000300FDFFFFFF01000000000000000200031E00"
                  "1C000910000A001400FDFFFFFF0900000000000000FE00141111010300111800160009100002001400BC71F2FF0100000000000000020005")


def _exp_vector(max_bytes=1024, hists=OTEL_EXP):
    app = H.Appender(2, max_bytes)
    for sch, vals in hists:
        assert app.add(H.Buckets.exponential(*sch).write_delta(vals)) == H.ACK
    return app


def test_exp_vector_bytes_match_the_documented_example():
    """ExpHistogramVector.scala:19-35 gives the bytes of the vector holding the three histograms of ExpHistogramVectorTest.scala:35-39."""
    app = _exp_vector()
    assert app.length == 3
    assert app.bytes().tobytes().hex().upper() == EXP_VECTOR_HEX
    rd = H.Reader(app.bytes())
    assert rd.length == 3
    for i, (sch, vals) in enumerate(OTEL_EXP):
        got_s, got_v = rd.apply_exp(i)
        assert got_s == sch and got_v.tolist() == vals


def test_exp_vector_sum_and_capacity():
    """ExpHistogramVectorTest.scala:201-212 (sum over rows of different schemes), :96-116 (159 histograms of 127 buckets fit a 15 kB
    vector), :131-145 (575 one-observation histograms fit, the 576th does not), :117-129 (180 positive buckets)."""
    rd = H.Reader(_exp_vector().bytes())
    sch, vals = rd.sum_exp(0, 2)
    assert sch == (3, -8, 9)
    assert vals.tolist() == [0.0, 0.0, 5.0, 5.0, 5.0, 5.0, 8.0, 8.0, 14.0, 20.0]
    counts = REAL_COUNTS.copy()
    assert counts.size == 127
    scheme = H.Buckets.exponential(3, -78, 126)
    app = H.Appender(2, 15000)
    for _ in range(159):
        assert app.add(scheme.write_delta(counts)) == H.ACK
        counts = counts + 10
    assert H.Reader(app.bytes()).length == 159
    one = H.Buckets.exponential(20, 9037032, 1).write_delta([0, 1])
    app = H.Appender(2, 15000)
    for i in range(576):
        assert app.add(one) == (H.ACK if i < 575 else H.VECTOR_TOO_SMALL)
    assert H.Reader(app.bytes()).length == 575
    big = H.Buckets.exponential(5, -1, 180)
    c = np.zeros(181, np.int64); c[-1] = 2
    app = H.Appender(2, 15000)
    assert app.add(big.write_delta(c)) == H.ACK
    s, v = H.Reader(app.bytes()).apply_exp(0)
    assert s == (5, -1, 180) and v.tolist() == c.tolist()


def test_exp_vector_rejects_invalid_blobs_and_empty_reads():
    """ExpHistogramVectorTest.scala:12-21, :214-229."""
    app = H.Appender(2, 1024)
    assert app.add(np.frombuffer(b"monkeying" + bytes(32), np.uint8)) == H.INVALID_HISTOGRAM
    assert app.add(np.array([1, 0, 0, 0, 0, 0, 0, 0], np.uint8)) == H.INVALID_HISTOGRAM      # null histogram
    assert app.length == 0
    with pytest.raises(RuntimeError):
        H.Reader(app.bytes()).apply_exp(0)
