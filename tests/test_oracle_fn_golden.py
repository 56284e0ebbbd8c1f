"""Pins the oracle's restatement of the remaining chunked range functions (stddev / stdvar / zscore / changes / quantile_over_time /
mad_over_time / holt_winters / predict_linear / present_over_time) and of the Long-column (L) variants against the reference's own
known-answer tests.  Every case cites the reference test it restates (paths relative to /root/reference); data literals are the
reference tests' inputs and expected outputs."""
import math
import numpy as np
import pytest

NaN = float("nan")
T0, PUB = 100000, 10000          # AggrOverTimeFunctionsSpec.scala:161-162 (defaultStartTS, pubFreq)


def rv(o, data, chunk_rows=None, val_mode=0):
    """timeValueRV (AggrOverTimeFunctionsSpec.scala:198-201): samples at T0 + i * PUB, one chunk unless chunk_rows is given."""
    st = o.Store()
    if len(data):
        ts = T0 + np.arange(len(data)) * PUB
        st.add_series_rows(ts, data, chunk_rows or [len(data)], val_mode=val_mode)
    else:
        st.add_series()
    return st


def windows(n, ws, step):
    """chunkedWindowIt (AggrOverTimeFunctionsSpec.scala:222-232) for data.sliding(ws, step): (start, step_ms, end, window_ms, first rows)."""
    idx = [0]
    while idx[-1] + ws < n and idx[-1] + step < n:    # scala's sliding(): the next group exists while the last one has not reached the end
        idx.append(idx[-1] + step)
    window = (ws - 1) * PUB
    start = T0 + window
    return start, step * PUB, start + (len(idx) - 1) * step * PUB, window, idx


def test_changes_known_answer(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:536-548
    o = oracle
    st = rv(o, [1.5, 2.5, 3.5, 4.5, 5.5], val_mode=2)
    out = st.query(o.FN_CHANGES, 100000, 20000, 150000, 30000)[0]
    assert list(out) == [0.0, 2.0, 3.0]


def test_changes_constant_is_zero(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:981-989 ("should return 0 for changes on constant value")
    o = oracle
    st = rv(o, [1.0] * 6)                            # integral doubles -> const DDV value vector (DoubleLongWrap reader, slope 0)
    out = st.query(o.FN_CHANGES, 100000, 20000, 150000, 30000)[0]
    assert list(out) == [0.0, 0.0, 0.0]


def test_changes_multi_chunk_random(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:657-684: all-distinct shuffled data, changes == window length - 1, also across chunks
    o = oracle
    rng = np.random.default_rng(11)
    data = np.concatenate([[1.1, 1.5, 2.5, 3.5, 4.5, 5.5], np.arange(1, 241, dtype=float)])
    rng.shuffle(data)
    for chunk_rows in ([len(data)], [100, 100, len(data) - 200]):
        for val_mode in (0, 1, 2):
            st = rv(o, data, chunk_rows, val_mode)
            for _ in range(6):
                ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 55))
                start, stepms, end, window, idx = windows(len(data), ws, step)
                out = st.query(o.FN_CHANGES, start, stepms, end, window)[0]
                for k, i in enumerate(idx):
                    assert out[k] == len(data[i:i + ws]) - 1


def test_quantile_over_time_known_answers(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:550-606
    o = oracle
    qs = [0, 0.5, 0.75, 0.8, 1, -1, 2]
    two = [0, 0.5, 0.75, 0.8, 1, -math.inf, math.inf]
    three = [0, 1, 1.5, 1.6, 2, -math.inf, math.inf]
    uneven = [0, 1, 2.5, 2.8, 4, -math.inf, math.inf]
    for i, q in enumerate(qs):
        a = rv(o, [0.0, 1.0]).query(o.FN_QUANTILE_OVER_TIME, 110000, 120000, 150000, 30000, params=(q,))[0][0]
        assert a == pytest.approx(two[i], abs=1e-10)
        b = rv(o, [1.0, 0.0, 2.0]).query(o.FN_QUANTILE_OVER_TIME, 120000, 20000, 130000, 50000, params=(q,))[0][0]
        assert b == pytest.approx(three[i], abs=1e-10)
        c = rv(o, [0.0, 1.0, 4.0]).query(o.FN_QUANTILE_OVER_TIME, 120000, 20000, 130000, 30000, params=(q,))[0][0]
        assert c == pytest.approx(uneven[i], abs=1e-10)
    assert math.isnan(rv(o, []).query(o.FN_QUANTILE_OVER_TIME, 110000, 120000, 150000, 30000, params=(0.5,))[0][0])
    # median over sliding windows of 1..500
    data = np.arange(1, 501, dtype=float)
    st = rv(o, data)
    rng = np.random.default_rng(5)
    for _ in range(6):
        ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 55))
        start, stepms, end, window, idx = windows(500, ws, step)
        out = st.query(o.FN_QUANTILE_OVER_TIME, start, stepms, end, window, params=(0.5,))[0]
        for k, i in enumerate(idx):
            s = sorted(data[i:i + ws]); h = len(s) // 2
            assert out[k] == ((s[h - 1] + s[h]) / 2.0 if len(s) % 2 == 0 else s[h])


def _mad(s):
    s = sorted(s); n = len(s); h = n // 2
    med = (s[h - 1] + s[h]) / 2.0 if n % 2 == 0 else s[h]
    d = sorted(abs(med - v) for v in s)
    rank = 0.5 * (n - 1); lo = max(0, math.floor(rank)); hi = min(n - 1, lo + 1); w = rank - math.floor(rank)
    return d[lo] * (1 - w) + d[hi] * w


def test_mad_over_time_known_answers(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:608-654
    o = oracle
    a = rv(o, [9.0, 6.0, 4.0, 1.0, 1.0, 2.0, 2.0]).query(o.FN_MAD_OVER_TIME, 170000, 10000, 170000, 100000)[0][0]
    assert a == pytest.approx(1.0, abs=1e-10)
    assert math.isnan(rv(o, []).query(o.FN_MAD_OVER_TIME, 110000, 120000, 150000, 30000)[0][0])
    data = np.arange(1, 501, dtype=float)
    st = rv(o, data)
    rng = np.random.default_rng(6)
    for _ in range(6):
        ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 55))
        start, stepms, end, window, idx = windows(500, ws, step)
        out = st.query(o.FN_MAD_OVER_TIME, start, stepms, end, window)[0]
        for k, i in enumerate(idx):
            assert out[k] == _mad(list(data[i:i + ws]))


def _holt_winters(arr, sf=0.01, tf=0.1):
    # the spec's own model, AggrOverTimeFunctionsSpec.scala:695-713
    if len(arr) < 2:
        return NaN
    s0 = arr[0]; b0 = arr[1] - arr[0]
    for i in range(1, len(arr)):
        sm = sf * arr[i] + (1 - sf) * (s0 + b0)
        b0 = tf * (sm - s0) + (1 - tf) * b0
        s0 = sm
    return s0


def test_holt_winters_known_answers(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:686-759
    o = oracle
    for data in ([15900.0, 15920.0, 15940.0, 15960.0, 15980.0, 16000.0], [23850.0, 23880.0, 23910.0, 23940.0, 23970.0, 24000.0],
                 [31800.0, 31840.0, 31880.0, 31920.0, 31960.0, 32000.0], [-15900.0, -15920.0, -15940.0, -15960.0, -15980.0, -16000.0]):
        for val_mode in (0, 2):
            out = rv(o, data, val_mode=val_mode).query(o.FN_HOLT_WINTERS, 160000, 100000, 180000, 100000, params=(0.01, 0.1))[0]
            assert out[0] == _holt_winters(data)
    data = np.arange(1, 241, dtype=float)
    st = rv(o, data)
    rng = np.random.default_rng(7)
    for _ in range(6):
        ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 80))
        start, stepms, end, window, idx = windows(240, ws, step)
        out = st.query(o.FN_HOLT_WINTERS, start, stepms, end, window, params=(0.01, 0.1))[0]
        for k, i in enumerate(idx):
            exp = _holt_winters(list(data[i:i + ws]))
            if math.isnan(exp): assert math.isnan(out[k])
            else: assert out[k] == pytest.approx(exp, abs=1e-10)
    with pytest.raises(RuntimeError):
        st.query(o.FN_HOLT_WINTERS, start, stepms, end, window, params=(1.5, 0.1))


def _predict_linear(rows, t_end, duration):
    # the spec's own model, AggrOverTimeFunctionsSpec.scala:775-796 (simple linear regression over (t - windowEnd) / 1000)
    n = 0; sx = sy = sxy = sx2 = 0.0
    for t, v in rows:
        if not math.isnan(v):
            x = (t - t_end) / 1000.0
            sx += x; sy += v; sxy += x * v; sx2 += x * x; n += 1
    if n < 2: return NaN
    cov = sxy - sx * sy / n; var = sx2 - sx * sx / n
    slope = cov / var
    return slope * duration + (sy / n - slope * sx / n)


def test_predict_linear(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:761-808: chunked predict_linear == least-squares model over the window's samples
    o = oracle
    data = np.arange(1, 241, dtype=float) * 3.0 + np.sin(np.arange(240))
    st = rv(o, data, [150, 90], val_mode=2)
    rng = np.random.default_rng(8)
    for _ in range(6):
        ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 80))
        start, stepms, end, window, idx = windows(240, ws, step)
        out = st.query(o.FN_PREDICT_LINEAR, start, stepms, end, window, params=(10.0,))[0]
        for k, i in enumerate(idx):
            t_end = start + k * stepms
            rows = [(T0 + j * PUB, data[j]) for j in range(i, min(i + ws, 240))]
            assert out[k] == pytest.approx(_predict_linear(rows, t_end, 10.0), rel=1e-9)
    one = rv(o, [5.0]).query(o.FN_PREDICT_LINEAR, 100000, 10000, 100000, 50000, params=(10.0,))[0]
    assert math.isnan(one[0])


def _nn(a): return [v for v in a if not math.isnan(v)]


def _std_var_nan(a):
    n = _nn(a); avg = sum(n) / len(n)
    return sum(v * v for v in n) / len(n) - avg * avg


def test_var_stddev_zscore_present_with_nans(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:914-967 (sum / avg / stdvar / stddev / zscore / present_over_time with NaNs and empty input)
    o = oracle
    cases = [
        [15900.0, 15920.0, 15940.0, 15960.0, 15980.0, 16000.0, 16020.0],
        [-15900.0, -15920.0, -15940.0, -15960.0, -15980.0, -16000.0],
        [15900.0, 15920.0, 15940.0, 15960.0, 15980.0, 16000.0, NaN],
        [23850.0, 23880.0, 23910.0, 23940.0, 23970.0, 24000.0],
        [31800.0, 31840.0, 31880.0, 31920.0, 31960.0, 32000.0],
        [31800.0, 31840.0, 31880.0, NaN, 31920.0, 31960.0, 32000.0],
        [NaN, 31800.0, 31840.0, 31880.0, 31920.0, 31960.0, 32000.0],
        [NaN] * 7,
        [],
    ]
    q = (160000, 100000, 180000, 100000)
    for data in cases:
        st = rv(o, data, val_mode=2)
        nn = _nn(data)
        sm = st.query(o.FN_SUM_OVER_TIME, *q)[0][0]
        av = st.query(o.FN_AVG_OVER_TIME, *q)[0][0]
        var = st.query(o.FN_STDVAR_OVER_TIME, *q)[0][0]
        dev = st.query(o.FN_STDDEV_OVER_TIME, *q)[0][0]
        z = st.query(o.FN_ZSCORE, *q)[0][0]
        pr = st.query(o.FN_PRESENT_OVER_TIME, *q)[0][0]
        if not nn:
            assert all(math.isnan(x) for x in (sm, av, var, dev, z, pr))
            continue
        s = 0.0
        for v in nn: s += v
        assert sm == s and av == s / len(nn)
        assert var == _std_var_nan(data) and dev == math.sqrt(_std_var_nan(data))
        if math.isnan(data[-1]): assert math.isnan(z)          # z_score model, :149-159: NaN unless the last sample is a number
        else: assert z == (data[-1] - s / len(nn)) / math.sqrt(_std_var_nan(data))
        assert pr == 1.0


def test_var_stddev_sliding(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:504-534: chunked stdvar / stddev == data.sliding(w, step).map(stdVar)
    o = oracle
    data = np.arange(1, 501, dtype=float)
    rng = np.random.default_rng(10)
    for chunk_rows, val_mode in (([500], 0), ([200, 200, 100], 2)):
        st = rv(o, data, chunk_rows, val_mode)
        for _ in range(6):
            ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 55))
            start, stepms, end, window, idx = windows(500, ws, step)
            var = st.query(o.FN_STDVAR_OVER_TIME, start, stepms, end, window)[0]
            dev = st.query(o.FN_STDDEV_OVER_TIME, start, stepms, end, window)[0]
            for k, i in enumerate(idx):
                w = data[i:i + ws]
                # the chunked function sums per chunk then adds the chunk sums; one chunk == the spec's sequential model exactly
                if len(chunk_rows) == 1:
                    s = 0.0; s2 = 0.0
                    for v in w: s += v; s2 += v * v
                    avg = s / len(w)
                    assert var[k] == s2 / len(w) - avg * avg and dev[k] == math.sqrt(s2 / len(w) - avg * avg)
                else:
                    avg = float(np.sum(w)) / len(w)
                    assert var[k] == pytest.approx(float(np.sum(w * w)) / len(w) - avg * avg, rel=1e-9)


def test_long_column_variants(oracle):
    """Long-column (L) functions: AggrOverTimeFunctions.scala:60-77,99-116,574-585,924-938,1019-1028,1144-1183,1211-1225,1322-1359;
    RangeFunction.scala:319-339,696-703.  The reference has no known-answer test for them (parity unpinned beyond the readers, which
    LongVectorTest pins); checked here against the functions' definitions."""
    o = oracle
    rng = np.random.default_rng(12)
    n = 120
    ts = T0 + np.arange(n) * PUB
    cases = {"ddv": (np.cumsum(rng.integers(0, 50, n)) + 1000).astype(np.int64),
             "const": (7 + 3 * np.arange(n)).astype(np.int64),
             "flat": np.full(n, 42, np.int64),
             "raw": rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)}
    for name, vals in cases.items():
        st = o.Store(); si = st.add_series()
        st.add_chunk_longs(si, ts[:80], vals[:80], raw=(name == "raw"))
        st.add_chunk_longs(si, ts[80:], vals[80:], raw=(name == "raw"))
        ws, step = 25, 7
        start, stepms, end, window, idx = windows(n, ws, step)
        q = lambda fn, **kw: st.query(fn, start, stepms, end, window, long_column=True, **kw)[0]
        sm, cnt, mn, mx, last, avg = (q(f) for f in (o.FN_SUM_OVER_TIME, o.FN_COUNT_OVER_TIME, o.FN_MIN_OVER_TIME, o.FN_MAX_OVER_TIME, o.FN_LAST, o.FN_AVG_OVER_TIME))
        ch = q(o.FN_CHANGES); med = q(o.FN_QUANTILE_OVER_TIME, params=(0.5,)); sd = q(o.FN_STDDEV_OVER_TIME); sv = q(o.FN_STDVAR_OVER_TIME)
        for k, i in enumerate(idx):
            w = vals[i:i + ws]
            parts = [w[:max(0, 80 - i)], w[max(0, 80 - i):]]
            if name == "raw":      # LongVectorDataReader64.sum: sequential double adds per chunk
                tot = 0.0
                for p in parts:
                    if len(p):
                        s = 0.0
                        for v in p: s += float(v)
                        tot += s
                assert sm[k] == tot
            else:
                assert sm[k] == float(int(w.sum()))          # exact in double for these magnitudes
            assert cnt[k] == len(w) and mn[k] == float(w.min()) and mx[k] == float(w.max()) and last[k] == float(w[-1])
            assert math.isnan(avg[k])                         # AvgOverTimeChunkedFunctionL: sum starts as NaN and `sum += ...` keeps it NaN
            s = sorted(float(v) for v in w); h = len(s) // 2
            assert med[k] == (s[h] if len(s) % 2 else s[h - 1] * 0.5 + s[h] * 0.5)
            if name != "raw":
                fw = w.astype(float); a = fw.sum() / len(w)
                assert sv[k] == pytest.approx((fw * fw).sum() / len(w) - a * a, rel=1e-9, abs=1e-6)
                assert sd[k] == pytest.approx(math.sqrt(max(0.0, (fw * fw).sum() / len(w) - a * a)), rel=1e-9, abs=1e-3)
        # changes: the readers' own treatment of the previous chunk's last value (DeltaDeltaVector.scala:212-227,280-288; LongBinaryVector.scala:248-265)
        if name == "flat":
            # const DDV, slope 0, prev = NaN.toLong = 0 != 42 and ignorePrev = false: the first chunk of every window counts one change
            assert all(c == 1.0 for c in ch)
        if name == "const":
            for k, i in enumerate(idx):
                w = vals[i:i + ws]; nparts = (1 if i < 80 else 0) + (1 if i + len(w) > 80 else 0)
                assert ch[k] == (len(w) - nparts) + nparts        # (end - start) per chunk + 1 per chunk (prev differs every time)
    st = o.Store(); si = st.add_series(); st.add_chunk_longs(si, ts, cases["ddv"])
    with pytest.raises(RuntimeError):
        st.query(o.FN_RATE, T0, PUB, T0 + 10 * PUB, 5 * PUB, long_column=True)


def test_serialized_range_vector_known_answers(oracle):
    # core/src/test/scala/filodb.core/query/SerializedRangeVectorSpec.scala:40-66,68-94,124-146
    o = oracle
    vals = np.array([[NaN, 1.0, NaN, 3.0, NaN, 5.0, 6.0, NaN, NaN, NaN, NaN]])
    c, rs, sr, fc = o.serialize_result(vals, 0, 100, 1000)
    assert rs[0] == 4 and sr[0] == 0 and c.shape[0] == 1                   # numRowsSerialized 4
    assert int(c[0, :4].view(np.int32)[0]) == 12 + 4 * 20                   # estimateSerializedRowBytes 80: 4 records of 20 bytes
    assert int(c[0, 4:8].view(np.int32)[0]) == 1 << 24                      # version word
    ts, v = o.result_rows(c, rs[0], sr[0], fc[0], 0, 100, 1000)
    assert list(ts) == list(range(0, 1001, 100)) and [x for x in v if not math.isnan(x)] == [1.0, 3.0, 5.0, 6.0] and len(v) == 11
    # instant query (start == end): NaN rows are kept.  The spec feeds 11 raw rows with RvRange(1000, 100, 1000); the restatement takes rows on
    # the output grid, so the same rule is exercised with one row
    c1, rs1, _, _ = o.serialize_result(np.array([[NaN]]), 1000, 100, 1000)
    assert rs1[0] == 1 and int(c1[0, :4].view(np.int32)[0]) == 12 + 20
    # 201 range vectors through one shared builder: each keeps its own row count; containers hold 204 records
    many = np.tile(vals, (201, 1))
    c, rs, sr, fc = o.serialize_result(many, 0, 100, 1000)
    assert (rs == 4).all() and c.shape[0] == (201 * 4 + 203) // 204
    for i in (0, 50, 51, 52, 102, 200):
        ts, v = o.result_rows(c, rs[i], sr[i], fc[i], 0, 100, 1000)
        assert [x for x in v if not math.isnan(x)] == [1.0, 3.0, 5.0, 6.0] and len(v) == 11
    assert sr[51] == 204 and fc[51] == 0 and sr[52] == 4 and fc[52] == 1      # vector 51 starts after a full container: startRecordNo = its record count
