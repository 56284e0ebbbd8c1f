"""Pins the CPU oracle (oracle/) against the reference's own golden vectors / known-answer tests.

Every case cites the reference test it restates (paths relative to /root/reference).  The data literals are the
reference tests' inputs and expected outputs.
"""
import math
import numpy as np
import pytest

NaN = float("nan")
INEXACT = 0x80000000


def bs(x):
    """binarySearch results are Scala Ints: compare as unsigned 32-bit."""
    return x & 0xFFFFFFFF


# ------------------------------------------------------------------ NibblePack
def test_nibblepack_pack8_even_nibbles(oracle):
    # core/src/test/scala/filodb.memory/format/NibblePackTest.scala:12-30
    inputs = [0, 0x0000003322110000, 0x0000004433220000, 0x0000005544330000, 0x0000006655440000, 0, 0, 0]
    expected = bytes([0x1e, 0x54, 0x11, 0x22, 0x33, 0x22, 0x33, 0x44, 0x33, 0x44, 0x55, 0x44, 0x55, 0x66])
    assert oracle.pack8(inputs) == expected


def test_nibblepack_pack8_odd_nibbles_and_unpack(oracle):
    # NibblePackTest.scala:32-51, 53-76
    inputs = [0, 0x0000003322100000, 0x0000004433200000, 0x0000005544300000, 0x0000006655400000,
              0x0000007654300000, 0, 0]
    expected = bytes([0x3e, 0x45, 0x21, 0x32, 0x23, 0x33, 0x44, 0x43, 0x54, 0x45, 0x55, 0x66, 0x43, 0x65, 0x07])
    assert oracle.pack8(inputs) == expected
    rc, out, remaining = oracle.unpack8(expected)
    assert rc == 0 and remaining == 0
    assert list(out) == inputs


def test_nibblepack_delta_roundtrip(oracle):
    # NibblePackTest.scala:78-98
    for inputs in ([0, 1000, 1001, 1002, 1003, 2005, 2010, 3034, 4045, 5056, 6067, 7078], [10000, 1032583228027]):
        packed = oracle.pack_delta(inputs)
        rc, out = oracle.unpack_delta(packed, len(inputs))
        assert rc == 0 and list(out) == inputs


def test_nibblepack_doubles_roundtrip(oracle):
    # NibblePackTest.scala:100-111
    inputs = [0.0, 2.5, 5.0, 7.5, 8, 13.2, 18.9, 89, 101.1, 102.3]
    packed = oracle.pack_doubles(inputs)
    rc, out = oracle.unpack_double_xor(packed, len(inputs))
    assert rc == 0 and list(out) == inputs


def test_nibblepack_random_roundtrip(oracle):
    # NibblePackTest.scala property tests ("should pack and unpack random longs/doubles")
    rng = np.random.default_rng(7)
    for n in (1, 2, 7, 8, 9, 16, 17, 63, 200, 401):
        vals = rng.normal(0, 1000, n)
        vals[rng.random(n) < 0.1] = 0.0
        packed = oracle.pack_doubles(vals)
        rc, out = oracle.unpack_double_xor(packed, n)
        assert rc == 0
        assert out.tobytes() == np.asarray(vals, np.float64).tobytes()
        ints = np.cumsum(rng.integers(0, 1 << 40, n)).astype(np.int64)
        rc, out = oracle.unpack_delta(oracle.pack_delta(ints), n)
        assert rc == 0 and (out == ints).all()


# ------------------------------------------------------------------ Int vectors
def test_int_nbits_packing(oracle):
    # core/src/test/scala/filodb.memory/format/vectors/IntBinaryVectorTest.scala:89-114 (4-bit / 2-bit packing)
    v = oracle.Vec(oracle.encode_int_vector([0, 2, 1, 4, 3], 4, False))
    assert v.int_length() == 5 and [v.int_apply(i) for i in range(5)] == [0, 2, 1, 4, 3]
    assert v.total_bytes() == 8 + 3
    v = oracle.Vec(oracle.encode_int_vector([0, 2, 1, 3, 2], 2, False))
    assert v.int_length() == 5 and [v.int_apply(i) for i in range(5)] == [0, 2, 1, 3, 2]
    assert v.total_bytes() == 8 + 2
    # minMaxToNbitsSigned, IntBinaryVector.scala:161-177
    assert oracle.minmax_to_nbits(0, 3) == (2, False)
    assert oracle.minmax_to_nbits(0, 15) == (4, False)
    assert oracle.minmax_to_nbits(-100, 100) == (8, True)
    assert oracle.minmax_to_nbits(0, 200) == (8, False)
    assert oracle.minmax_to_nbits(-1000, 1000) == (16, True)
    assert oracle.minmax_to_nbits(0, 60000) == (16, False)
    assert oracle.minmax_to_nbits(-70000, 7) == (32, True)
    for nbits, signed, lo, hi in ((8, True, -128, 127), (8, False, 0, 255), (16, True, -32768, 32767),
                                  (16, False, 0, 65535), (32, True, -2**31, 2**31 - 1)):
        vals = np.random.default_rng(nbits).integers(lo, hi, 77, endpoint=True)
        v = oracle.Vec(oracle.encode_int_vector(vals, nbits, signed))
        assert v.int_length() == 77
        assert [v.int_apply(i) for i in range(77)] == list(vals)
        assert v.int_sum(3, 70) == int(vals[3:71].sum())


# ------------------------------------------------------------------ Long / DDV vectors
def test_long_optimize_ddv_nbits4(oracle):
    # LongVectorTest.scala:98-123: Seq(0,2,1,4,3) -> DDV, nbits=4, 28 + 3 bytes
    orig = [0, 2, 1, 4, 3]
    v = oracle.Vec(oracle.encode_longs(orig))
    assert v.vector_type() & 0xff == 0x08
    assert v.long_length() == 5 and v.longs() == orig
    assert v.total_bytes() == 28 + 3
    assert v.long_sum(0, 4) == 2 + 1 + 4 + 3


def test_long_ddv_const(oracle):
    # LongVectorTest.scala:124-137
    start = 1700000000123
    orig = [i * 10000 + start for i in range(51)]
    v = oracle.Vec(oracle.encode_longs(orig))
    assert v.total_bytes() == 24 and v.longs() == orig
    assert v.long_sum(0, 13) == float(sum(orig[:14]))


def test_long_binary_search_raw_and_ddv(oracle):
    # LongVectorTest.scala:185-225
    orig = [1000, 2001, 2999, 5123, 5250, 6004, 7678]
    o = oracle
    st = o.Store(); s = st.add_series(); st.add_chunk(s, orig, [0.0] * len(orig), ts_mode=o.TS_RAW)
    r = o.Vec(st.vector_bytes(s, 0, 0))
    assert bs(r.binary_search(0)) == INEXACT | 0
    assert bs(r.binary_search(999)) == INEXACT | 0
    assert r.binary_search(1000) == 0
    assert bs(r.binary_search(3000)) == INEXACT | 3
    assert bs(r.binary_search(7677)) == INEXACT | 6
    assert r.binary_search(7678) == 6
    assert bs(r.binary_search(7679)) == INEXACT | 7
    assert [r.ceiling_index(x) for x in (0, 999, 1000, 3000, 7677, 7678, 7679)] == [-1, -1, 0, 2, 5, 6, 6]
    d = o.Vec(o.encode_longs(orig))
    assert d.vector_type() == ((0x08 << 8) | 0x08)      # DeltaDeltaDataReader
    assert bs(d.binary_search(0)) == INEXACT | 0
    assert bs(d.binary_search(999)) == INEXACT | 0
    assert d.binary_search(1000) == 0
    assert bs(d.binary_search(3000)) == INEXACT | 3
    assert d.binary_search(5123) == 3
    assert d.binary_search(5250) == 4
    assert bs(d.binary_search(6003)) == INEXACT | 5
    assert d.binary_search(6004) == 5
    assert bs(d.binary_search(7677)) == INEXACT | 6
    assert d.binary_search(7678) == 6
    assert bs(d.binary_search(7679)) == INEXACT | 7
    assert d.ceiling_index(7679) == 6


def test_long_binary_search_slope0(oracle):
    # LongVectorTest.scala:227-238
    v = oracle.Vec(oracle.encode_longs([1000 + (x // 5) for x in range(16)]))
    assert v.vector_type() == ((0x08 << 8) | 0x08)
    assert bs(v.binary_search(999)) == INEXACT | 0
    assert v.binary_search(1000) == 0
    assert v.binary_search(1001) == 9


def test_long_binary_search_ddv_const(oracle):
    # LongVectorTest.scala:240-264
    start = 1700000000000
    orig = [i * 10000 + start for i in range(51)]
    v = oracle.Vec(oracle.encode_longs(orig))
    assert v.total_bytes() == 24
    assert bs(v.binary_search(start - 1)) == INEXACT | 0
    assert v.binary_search(start) == 0
    assert bs(v.binary_search(start + 1)) == INEXACT | 1
    assert bs(v.binary_search(start + 100001)) == INEXACT | 11
    assert bs(v.binary_search(start + len(orig) * 10000)) == INEXACT | len(orig)
    v2 = oracle.Vec(oracle.encode_longs([1000] * 16))
    assert v2.total_bytes() == 24
    assert bs(v2.binary_search(999)) == INEXACT | 0
    assert v2.binary_search(1000) == 0
    assert bs(v2.binary_search(1001)) == INEXACT | 16


def test_long_binary_search_random(oracle):
    # LongVectorTest.scala:266-301 (property test)
    rng = np.random.default_rng(11)
    for maxval in (1000, 5000, 30000):
        for _ in range(10):
            n = int(rng.integers(3, 120))
            longs = np.concatenate([[10000], 10000 + np.cumsum(rng.integers(10, maxval, n))]).astype(np.int64)
            v = oracle.Vec(oracle.encode_longs(longs))
            for num in rng.integers(0, int(longs[-1]) * 3, 60):
                out = v.binary_search(int(num))
                idx = int(np.searchsorted(longs, num, side="left"))
                if idx < len(longs):
                    assert bs(out) == (idx if longs[idx] == num else (idx | INEXACT))
                else:
                    assert bs(out) == (INEXACT | len(longs))


def test_timestamp_approx_const(oracle):
    # LongVectorTest.scala:353-382: TimestampAppendingVector uses const DDV when within +/-250 ms of the slope line
    start = 1700000000000
    orig = [i * 10000 + start for i in range(50)]
    jitter = list(orig); jitter[5] += 200; jitter[17] -= 249
    v = oracle.Vec(oracle.encode_timestamps(jitter))
    assert v.total_bytes() == 24
    assert v.longs() == orig          # approximated back onto the line
    jitter[9] += 251
    v = oracle.Vec(oracle.encode_timestamps(jitter))
    assert v.total_bytes() > 24 and v.longs() == jitter
    # a plain LongAppendingVector never approximates
    v = oracle.Vec(oracle.encode_longs([o + (1 if i == 3 else 0) for i, o in enumerate(orig)]))
    assert v.total_bytes() > 24


# ------------------------------------------------------------------ Double vectors
def test_double_optimize_integral_to_ddv_const(oracle):
    # DoubleVectorTest.scala:77-87 and :147-157
    v = oracle.Vec(oracle.encode_doubles([float(i) for i in range(10)]))
    assert v.total_bytes() == 24 and v.double_length() == 10
    assert v.doubles() == [float(i) for i in range(10)]
    v = oracle.Vec(oracle.encode_doubles([float(i) for i in range(100000, 100005)]))
    assert v.total_bytes() == 24 and v.double_apply(2) == 100002.0


def test_double_edge_case_not_const(oracle):
    # DoubleVectorTest.scala:135-145: 55, 60, 60 ... must NOT become a const DDV
    orig = [55.0, 60.0] + [60.0] * 10
    v = oracle.Vec(oracle.encode_doubles(orig))
    assert v.total_bytes() > 24 and v.doubles() == orig


def test_double_sum_ignores_nan(oracle):
    # DoubleVectorTest.scala:171-196
    orig = [1000, 2001.1, 2999.99, 5123.4, 5250, 6004, 7678]
    v = oracle.Vec(oracle.encode_doubles(orig + [NaN], mode=2))
    assert v.double_sum(2, len(orig) - 1) == sum(orig[2:])
    assert v.double_sum(2, len(orig)) == sum(orig[2:])
    assert v.double_count(0, len(orig)) == len(orig)
    # all NaN -> NaN (DoubleVector.scala:243-253)
    allnan = oracle.Vec(oracle.encode_doubles([NaN, NaN, NaN], mode=2))
    assert math.isnan(allnan.double_sum(0, 2)) and allnan.double_count(0, 2) == 0


def test_counter_drop_flag_and_positions(oracle):
    # DoubleVectorTest.scala:101-133
    orig = [3904.0, 3904.0, 3905.0, 3908.0, 3909.0, NaN, 3910.0, 3912.0, 3914.0, 3914.0, 3915.0, NaN,
            3905.0, 3906.0, 3907.0, 3908.0, 3909.0, 3910.0]
    v = oracle.Vec(oracle.encode_doubles(orig, detect_drops=True))
    assert v.dropped() and v.double_length() == len(orig)
    assert v.drop_positions() == [i for i, x in enumerate(orig) if math.isnan(x)]
    end_nan = [3904.0, 3904.0, 3905.0, 3908.0, 3909.0, 3910.0, 3912.0, 3914.0, 3914.0, 3915.0, 3916.0, 3917.0, 3918.0,
               3919.0, 3920.0, 3922.0, NaN]
    assert oracle.Vec(oracle.encode_doubles(end_nan, detect_drops=True)).dropped()
    # :89-99 detectDropAndCorrection / updateCorrection with NaN first
    v = oracle.Vec(oracle.encode_doubles([NaN, 3904.0, 3904.0, 3905.0, 3908.0, 3909.0], detect_drops=True, mode=2))
    assert v.detect_drop((300.0, 0.0)) == (300.0, 300.0)
    assert v.update_correction((300.0, 0.0), force_corrected=False) == (3909.0, 0.0)
    v = oracle.Vec(oracle.encode_doubles([NaN, 3904.0, 3904.0, 3905.0, NaN, NaN], detect_drops=True, mode=2))
    assert v.detect_drop((300.0, 0.0)) == (300.0, 300.0)
    assert v.update_correction((300.0, 0.0), force_corrected=False) == (3905.0, 0.0)


def test_counter_correction_golden(oracle):
    # DoubleVectorTest.scala:273-358: [101,102.5,9,13.3,21.1]
    data = [101, 102.5, 9, 13.3, 21.1]
    v = oracle.Vec(oracle.encode_doubles(data, detect_drops=True))
    assert v.dropped()
    assert v.corrected_value(0, None) == 101 and v.corrected_value(1, None) == 102.5
    assert v.corrected_value(2, None) == 111.5
    assert v.corrected_value(4, None) == 123.6
    assert v.drop_positions() == [2]
    assert v.update_correction((999.9, 50.0)) == (21.1, 152.5)
    # with a carried-over correction
    assert v.corrected_value(2, (0.0, 50.0)) == 161.5
    # non-dropped chunk: plain add (DoubleVector.scala:203-207), detect drop across chunk boundary (:177-187)
    v2 = oracle.Vec(oracle.encode_doubles([5.0, 7.5, 9.25], detect_drops=True))
    assert not v2.dropped()
    assert v2.detect_drop((21.1, 152.5)) == (21.1, 152.5 + 21.1)
    assert v2.detect_drop((4.0, 1.0)) == (4.0, 1.0)
    assert v2.corrected_value(1, (21.1, 10.0)) == 17.5
    assert v2.update_correction((21.1, 10.0)) == (9.25, 10.0)
    assert v2.update_correction(None) == (9.25, 0.0)


def test_xor_container_matches_plain_semantics(oracle):
    # Our container: payload pinned (== packDoubles), semantics == unpackDoubleXOR -> DoubleVectorDataReader64
    rng = np.random.default_rng(3)
    vals = 15 + np.sin(np.arange(1, 402)) + rng.normal(0, 1, 401)
    vals[[5, 77, 400]] = NaN
    x = oracle.Vec(oracle.encode_doubles(vals, mode=1))
    r = oracle.Vec(oracle.encode_doubles(vals, mode=2))
    assert x.double_length() == r.double_length() == 401
    assert np.asarray(x.doubles()).tobytes() == np.asarray(r.doubles()).tobytes()
    for s, e in ((0, 400), (3, 90), (77, 77), (5, 5)):
        a, b = x.double_sum(s, e), r.double_sum(s, e)
        assert (math.isnan(a) and math.isnan(b)) or a == b
        assert x.double_count(s, e) == r.double_count(s, e)
    # payload bytes are exactly NibblePack.packDoubles
    po = int(np.frombuffer(x.b[14:16].tobytes(), np.uint16)[0])
    packed = oracle.pack_doubles(vals)
    assert x.b[po:po + len(packed)].tobytes() == packed


# ------------------------------------------------------------------ chunkID
def test_chunk_id(oracle):
    # core/src/main/scala/filodb.core/store/package.scala:112-130
    cid = oracle.lib().fo_chunk_id(1700000000000, 1700000123)
    assert oracle.lib().fo_start_time_from_chunk_id(cid) == 1700000000000
    assert cid < 0       # "Chunk ids will be negative until the year ~2039"


# ------------------------------------------------------------------ extrapolatedRate / rate known answers
PROM_SAMPLES = [(1548191486000, 84.0), (1548191496000, 152.0), (1548191506000, 195.0), (1548191516000, 222.0),
                (1548191526000, 245.0), (1548191536000, 251.0), (1548191546000, 329.0), (1548191556000, 374.0),
                (1548191566000, 431.0)]
PROM_EXPECTED = {1548191496000: 0.34, 1548191511000: 0.555, 1548191526000: 0.60375, 1548191541000: 0.668,
                 1548191556000: 1.0357142857142858}


def _store_one(o, samples, chunk_rows=None, detect_drops=True, val_mode=0):
    st = o.Store()
    ts = [t for t, _ in samples]; vs = [v for _, v in samples]
    st.add_series_rows(ts, vs, chunk_rows or [len(ts)], val_mode=val_mode, detect_drops=detect_drops)
    return st


@pytest.mark.parametrize("val_mode", [0, 1, 2])
def test_rate_matches_prometheus(oracle, val_mode):
    # query/src/test/scala/filodb/query/exec/WindowIteratorSpec.scala:219-255
    o = oracle
    st = _store_one(o, PROM_SAMPLES, val_mode=val_mode)
    start, step, end, window = 1548191496000, 15000, 1548191796000, 300000
    out = st.query(o.FN_RATE, start, step, end, window, cumulative=True)[0]
    for k, v in enumerate(out):
        t = start + k * step
        if t in PROM_EXPECTED:
            assert v == pytest.approx(PROM_EXPECTED[t], abs=1e-10)
    sl = o.sliding([t for t, _ in PROM_SAMPLES], [v for _, v in PROM_SAMPLES], o.FN_RATE, start, step, end, window, cumulative=True)
    for a, b in zip(out, sl):
        assert (math.isnan(a) and math.isnan(b)) or a == b
    # instant query, step adjusted to 1 (:286-323)
    one = st.query(o.FN_RATE, 1548191796000, 1, 1548191796000, 300000, cumulative=True)
    assert one.shape == (1, 1)


def test_rate_nan_end_of_series_marker(oracle):
    # WindowIteratorSpec.scala:257-284  -> 0.5870753512132821 exactly
    o = oracle
    samples = [(1614821996000, NaN), (1614821996100, 489.0), (1614821997000, NaN), (1614822566000, 19.0),
               (1614822596000, 26.0), (1614822626000, 26.0), (1614822656000, 26.0), (1614822686000, 26.0),
               (1614822716000, 26.0), (1614822717000, NaN), (1614822866000, 5.0)]
    st = _store_one(o, samples)
    out = st.query(o.FN_RATE, 1614822880000, 15000, 1614822880000, 900000, cumulative=True)
    assert out[0, 0] == 0.5870753512132821
    sl = o.sliding([t for t, _ in samples], [v for _, v in samples], o.FN_RATE, 1614822880000, 15000, 1614822880000, 900000, True)
    assert sl[0] == 0.5870753512132821


COUNTER_SAMPLES = [(8072000, 4419.0), (8082100, 4511.0), (8092196, 4614.0), (8102215, 4724.0), (8112223, 4909.0),
                   (8122388, 4948.0), (8132570, 5000.0), (8142822, 5095.0), (8152858, 5102.0), (8162999, 5201.0)]


def test_rate_functions_spec(oracle):
    # query/src/test/scala/filodb/query/exec/rangefn/RateFunctionsSpec.scala:58-178
    o = oracle
    err = 1e-7
    q = COUNTER_SAMPLES
    startTs, endTs = 8071950, 8163070
    st = _store_one(o, q)
    expected = (q[-1][1] - q[0][1]) / (q[-1][0] - q[0][0]) * 1000
    assert st.query(o.FN_RATE, endTs, 10000, endTs, endTs - startTs, cumulative=True)[0, 0] == pytest.approx(expected, abs=err)

    # reset at chunk boundary (:72-92)
    chunk2 = [(8173000, 325.0), (8183000, 511.0), (8193000, 614.0), (8203000, 724.0), (8213000, 909.0)]
    st = _store_one(o, q + chunk2, chunk_rows=[10, 5])
    endTs2 = 8213070
    expected = (chunk2[-1][1] + q[-1][1] - q[0][1]) / (chunk2[-1][0] - q[0][0]) * 1000
    assert st.query(o.FN_RATE, endTs2, 10000, endTs2, endTs2 - startTs, cumulative=True)[0, 0] == pytest.approx(expected, abs=err)

    # NaN at the beginning of the 2nd chunk (:94-115)
    chunk2n = [(8173000, NaN)] + chunk2[1:]
    st = _store_one(o, q + chunk2n, chunk_rows=[10, 5])
    assert st.query(o.FN_RATE, endTs2, 10000, endTs2, endTs2 - startTs, cumulative=True)[0, 0] == pytest.approx(expected, abs=err)

    # drops in the middle of chunks (:117-158)
    reset1 = [(8072000, 4419.0), (8082100, 4511.0), (8092196, 4614.0), (8102215, 4724.0), (8112223, 4909.0),
              (8122388, 948.0), (8132570, 1000.0), (8142822, 1095.0), (8152858, 1102.0), (8162999, 1201.0)]
    reset2 = [(8173000, 1325.0), (8183000, 1511.0), (8193000, 214.0), (8203000, 324.0), (8213000, 409.0)]
    corrections = reset1[4][1] + reset2[1][1]
    expected = (reset2[-1][1] + corrections - reset1[0][1]) / (reset2[-1][0] - reset1[0][0]) * 1000
    for rows in ([10, 5], [15]):
        st = _store_one(o, reset1 + reset2, chunk_rows=rows)
        assert st.query(o.FN_RATE, endTs2, 10000, endTs2, endTs2 - startTs, cumulative=True)[0, 0] == pytest.approx(expected, abs=err)

    # one sample in window -> NaN (:160-167)
    st = _store_one(o, q)
    assert math.isnan(st.query(o.FN_RATE, 8103215, 10000, 8103215, 2000, cumulative=True)[0, 0])
    # flat counter -> 0.0 (:169-178)
    st = _store_one(o, [(t, q[0][1]) for t, _ in q])
    assert st.query(o.FN_RATE, endTs, 10000, endTs, endTs - startTs, cumulative=True)[0, 0] == 0.0


def test_rate_chunked_equals_sliding_random(oracle):
    # RateFunctionsSpec.scala:181-212: chunked == sliding for random window/step, across chunk boundaries (maxChunkSize=200)
    o = oracle
    rng = np.random.default_rng(5)
    data = (np.arange(1, 501) * 10 + rng.integers(0, 10, 500)).astype(float)
    ts = 100000 + np.arange(500) * 10000
    st = o.Store(); st.add_series_rows(ts, data, [200, 200, 100], detect_drops=True)
    for _ in range(10):
        ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 55))
        window = (ws - 1) * 10000
        start = 100000 + window
        nwin = len(range(0, 500 - ws + 1, step))
        end = start + (nwin - 1) * step * 10000
        for fn in (o.FN_RATE, o.FN_INCREASE):
            ch = st.query(fn, start, step * 10000, end, window, cumulative=True)[0]
            sl = o.sliding(ts, data, fn, start, step * 10000, end, window, cumulative=True)
            assert ch.tobytes() == sl.tobytes()


# ------------------------------------------------------------------ *_over_time known answers
OT_SAMPLES = [(100000, 1.0), (153000, 2.0), (250000, 3.0), (270000, 4.0), (280000, 5.0), (360000, 6.0), (430000, 7.0),
              (690000, 8.0), (700000, 9.0), (710000, NaN)]


def _non_nan(o, st, fn, start, step, end, window):
    out = st.query(fn, start, step, end, window)[0]
    return [(start + k * step, v) for k, v in enumerate(out) if not math.isnan(v)]


def test_over_time_known_answers(oracle):
    # WindowIteratorSpec.scala:180-217 (sum), :466-500 (avg), :502-535 (count), :594-630 (min), :632-665 (max)
    o = oracle
    st = _store_one(o, OT_SAMPLES, detect_drops=False)
    assert _non_nan(o, st, o.FN_SUM_OVER_TIME, 50000, 100000, 1100000, 100000) == \
        [(150000, 1.0), (250000, 5.0), (350000, 12.0), (450000, 13.0), (750000, 17.0)]
    assert _non_nan(o, st, o.FN_AVG_OVER_TIME, 50000, 100000, 700000, 100000) == \
        [(150000, 1.0), (250000, 2.5), (350000, 4.0), (450000, 6.5)]
    assert _non_nan(o, st, o.FN_COUNT_OVER_TIME, 50000, 100000, 700000, 100000) == \
        [(150000, 1.0), (250000, 2.0), (350000, 3.0), (450000, 2.0)]
    assert _non_nan(o, st, o.FN_MIN_OVER_TIME, 50000, 100000, 700000, 100000) == \
        [(150000, 1.0), (250000, 2.0), (350000, 3.0), (450000, 6.0)]
    assert _non_nan(o, st, o.FN_MAX_OVER_TIME, 50000, 100000, 700000, 100000) == \
        [(150000, 1.0), (250000, 3.0), (350000, 5.0), (450000, 7.0)]


def test_downsampled_avg_and_count_known_answers(oracle):
    """WindowIteratorSpec.scala:540-584 ("should calculate query results from downsampled data"): over a downsample schema avg_over_time
    becomes AvgWithSumAndCountOverTime = sum_over_time(sum column) / sum_over_time(count column) and count_over_time becomes sum_over_time of
    the count column (RangeFunction.downsampleRangeFunction, RangeFunction.scala:272-279; AvgWithSumAndCountOverTimeFuncD,
    AggrOverTimeFunctions.scala:820-854).  This is the quotient filo_query_avg_sum_count forms on the device."""
    o = oracle
    # (timestamp, min, max, sum, count, avg) rows of the reference test
    rows = [(100000, 2.0, 5.0, 20.0, 5.0, 2.8), (153000, 1.0, 6.0, 18.0, 3.0, 1.4), (250000, 3.0, 7.0, 21.0, 5.0, 5.0), (270000, 2.0, 10.0, 22.0, 4.0, 6.0),
            (280000, 1.5, 2.0, 10.0, 6.0, 1.75), (360000, 0.6, 7.0, 23.0, 7.0, 2.0), (430000, 7.0, 10.0, 60.0, 5.0, 8.0), (690000, 1.8, 5.0, 25.0, 7.0, 3.0),
            (700000, 4.9, 12.0, 80.0, 10.0, 10.0), (710000, 0.1, 3.0, 10.0, 10.0, 1.0)]
    sums = _store_one(o, [(r[0], r[3]) for r in rows], detect_drops=False)
    cnts = _store_one(o, [(r[0], r[4]) for r in rows], detect_drops=False)
    q = (50000, 100000, 750000, 100000)
    num = sums.query(o.FN_SUM_OVER_TIME, *q)[0]; den = cnts.query(o.FN_SUM_OVER_TIME, *q)[0]
    with np.errstate(all="ignore"):
        avg = num / den
    got = [(q[0] + k * q[1], v) for k, v in enumerate(avg) if not math.isnan(v)]
    assert got == [(150000, 4.0), (250000, 4.875), (350000, 3.533333333333333), (450000, 6.916666666666667), (750000, 4.2592592592592595)]
    assert _non_nan(o, cnts, o.FN_SUM_OVER_TIME, *q) == [(150000, 5.0), (250000, 8.0), (350000, 15.0), (450000, 12.0), (750000, 27.0)]


def test_last_sample_staleness(oracle):
    # WindowIteratorSpec.scala:325-368 (window 180000) and :370-431 (5 min) and :433-464
    o = oracle
    samples = [(1540832354000, 1.0), (1540835954000, 2.0), (1540839554000, 3.0), (1540843154000, 4.0),
               (1540846754000, 237.0), (1540850354000, 330.0)]
    st = _store_one(o, samples, detect_drops=False)
    res = _non_nan(o, st, o.FN_LAST, 1540845090000, 15000, 1540855905000, 180000)
    exp = [(1540846755000 + 15000 * i, 237.0) for i in range(12)] + [(1540850355000 + 15000 * i, 330.0) for i in range(12)]
    assert res == exp
    res = _non_nan(o, st, o.FN_LAST, 1540845090000, 15000, 1540855905000, 300000)
    exp = [(1540846755000 + 15000 * i, 237.0) for i in range(20)] + [(1540850355000 + 15000 * i, 330.0) for i in range(20)]
    assert res == exp
    st = _store_one(o, [(100000, 100.0), (153000, 160.0), (200000, 200.0)], detect_drops=False)
    assert _non_nan(o, st, o.FN_LAST, 100000, 100000, 600000, 300001) == \
        [(100000, 100.0), (200000, 200.0), (300000, 200.0), (400000, 200.0), (500000, 200.0)]


def test_sum_over_time_random_windows(oracle):
    # rangefn/AggrOverTimeFunctionsSpec.scala:287-303: chunked sum == data.sliding(w, step).map(_.sum), 2 chunks of <=200
    o = oracle
    rng = np.random.default_rng(9)
    data = np.arange(1, 241, dtype=float)      # -> const DDV value vector (closed-form sum path)
    ts = 100000 + np.arange(240) * 10000
    for val_mode, vals in ((0, data), (2, data + rng.random(240)), (1, data + rng.random(240))):
        st = o.Store(); st.add_series_rows(ts, vals, [200, 40], val_mode=val_mode)
        for _ in range(8):
            ws = int(rng.integers(10, 110)); step = int(rng.integers(5, 55))
            window = (ws - 1) * 10000; start = 100000 + window
            idx = list(range(0, 240 - ws + 1, step))
            end = start + (len(idx) - 1) * step * 10000
            out = st.query(o.FN_SUM_OVER_TIME, start, step * 10000, end, window)[0]
            # window sums: sequential per chunk, then chunk sums added (AggrOverTimeFunctions.scala:568-570)
            for k, i in enumerate(idx):
                rows = vals[i:i + ws]
                c1 = [v for j, v in enumerate(rows) if i + j < 200]; c2 = [v for j, v in enumerate(rows) if i + j >= 200]
                exp = 0.0
                if val_mode == 0:
                    exp = float(sum(rows))
                    assert out[k] == exp
                else:
                    tot = None
                    for part in (c1, c2):
                        if part:
                            s = 0.0
                            for v in part: s += v
                            tot = s if tot is None else tot + s
                    assert out[k] == tot
            mn = st.query(o.FN_MIN_OVER_TIME, start, step * 10000, end, window)[0]
            mx = st.query(o.FN_MAX_OVER_TIME, start, step * 10000, end, window)[0]
            cnt = st.query(o.FN_COUNT_OVER_TIME, start, step * 10000, end, window)[0]
            avg = st.query(o.FN_AVG_OVER_TIME, start, step * 10000, end, window)[0]
            for k, i in enumerate(idx):
                assert mn[k] == vals[i:i + ws].min() and mx[k] == vals[i:i + ws].max() and cnt[k] == ws
                assert avg[k] == out[k] / ws


def test_sum_nan_repoison_quirk(oracle):
    # AggrOverTimeFunctions.scala:568-570: an all-NaN chunk re-poisons a non-NaN running sum; avg :1000
    o = oracle
    ts = [1000, 2000, 3000, 4000]
    st = o.Store(); st.add_series_rows(ts, [1.0, 2.0, NaN, NaN], [2, 2], val_mode=2)
    assert math.isnan(st.query(o.FN_SUM_OVER_TIME, 4000, 1000, 4000, 3000)[0, 0])
    assert math.isnan(st.query(o.FN_AVG_OVER_TIME, 4000, 1000, 4000, 3000)[0, 0])
    assert st.query(o.FN_COUNT_OVER_TIME, 4000, 1000, 4000, 3000)[0, 0] == 2.0
    st = o.Store(); st.add_series_rows(ts, [NaN, NaN, 1.0, 2.0], [2, 2], val_mode=2)
    assert st.query(o.FN_SUM_OVER_TIME, 4000, 1000, 4000, 3000)[0, 0] == 3.0
    # count: NaN when no chunk has rows in the window, 0.0 when rows exist but all NaN (:953-956)
    assert math.isnan(st.query(o.FN_COUNT_OVER_TIME, 500, 1000, 500, 100)[0, 0])
    assert st.query(o.FN_COUNT_OVER_TIME, 2000, 1000, 2000, 1000)[0, 0] == 0.0
    assert math.isnan(st.query(o.FN_AVG_OVER_TIME, 2000, 1000, 2000, 1000)[0, 0])


# ------------------------------------------------------------------ across-series aggregators
def test_aggregators_vs_transpose_and_fold(oracle):
    # query/src/test/scala/filodb/query/exec/AggrOverRangeVectorsSpec.scala:31-206 (sum/min/max/count/avg vs fold), :208-332 (NaN)
    o = oracle
    rng = np.random.default_rng(21)
    S, n = 12, 60
    ts = 100000 + np.arange(n) * 10000
    st = o.Store(); rows = []
    for s in range(S):
        v = rng.random(n) * 100
        v[rng.random(n) < 0.15] = NaN
        if s == 3: v[:] = NaN
        st.add_series_rows(ts, v, [n], val_mode=2); rows.append(v)
    groups = np.array([s % 3 for s in range(S)], np.int32)
    start, step, end, window = 100000, 10000, 100000 + (n - 1) * 10000, 0 + 5000
    per = st.query(o.FN_LAST, start, step, end, window)
    T = per.shape[1]
    for aggr in (o.AGG_SUM, o.AGG_MIN, o.AGG_MAX, o.AGG_COUNT):
        out = st.query(o.FN_LAST, start, step, end, window, aggr=aggr, group_ids=groups, n_groups=3)
        for g in range(3):
            for t in range(T):
                col = [per[s, t] for s in range(S) if groups[s] == g and not math.isnan(per[s, t])]
                if not col:
                    assert math.isnan(out[g, t])
                elif aggr == o.AGG_SUM:
                    acc = 0.0
                    for x in col: acc += x
                    assert out[g, t] == acc
                elif aggr == o.AGG_MIN: assert out[g, t] == min(col)
                elif aggr == o.AGG_MAX: assert out[g, t] == max(col)
                else: assert out[g, t] == len(col)
    avg, cnt = st.query(o.FN_LAST, start, step, end, window, aggr=o.AGG_AVG, group_ids=groups, n_groups=3)
    for g in range(3):
        for t in range(T):
            col = [per[s, t] for s in range(S) if groups[s] == g and not math.isnan(per[s, t])]
            assert cnt[g, t] == len(col)
            if col: assert avg[g, t] == pytest.approx(sum(col) / len(col), rel=1e-12)
            else: assert math.isnan(avg[g, t])
    # topk / bottomk (:533-552, 601-633): k largest non-NaN per window, ascending in the row
    for aggr, rev in ((o.AGG_TOPK, True), (o.AGG_BOTTOMK, False)):
        vals, ids = st.query(o.FN_LAST, start, step, end, window, aggr=aggr, k=3, group_ids=np.zeros(S, np.int32), n_groups=1)
        for t in range(T):
            col = sorted([per[s, t] for s in range(S) if not math.isnan(per[s, t])], reverse=rev)[:3]
            got = [v for v, i in zip(vals[0, t], ids[0, t]) if i >= 0]
            assert sorted(got, reverse=rev) == col
            assert got == sorted(got, reverse=not rev)
            for v, i in zip(vals[0, t], ids[0, t]):
                if i >= 0: assert per[i, t] == v


def test_aggregators_literal_known_answers(oracle):
    # AggrOverRangeVectorsSpec.scala:421-470 ("should return NaN when all values are NaN for a timestamp"): literal expectations
    o = oracle
    ts = np.array([1000, 2000], np.int64)
    st = o.Store()
    for v in (5.6, 4.4, 5.4):
        st.add_series_rows(ts, np.array([NaN, v]), [2], val_mode=0)
    g = np.zeros(3, np.int32)
    q = lambda aggr, **kw: st.query(o.FN_LAST, 1000, 1000, 2000, 500, aggr=aggr, group_ids=g, n_groups=1, **kw)
    s = q(o.AGG_SUM);   assert math.isnan(s[0, 0]) and s[0, 1] == 15.4
    m = q(o.AGG_MIN);   assert math.isnan(m[0, 0]) and m[0, 1] == 4.4
    x = q(o.AGG_MAX);   assert math.isnan(x[0, 0]) and x[0, 1] == 5.6
    c = q(o.AGG_COUNT); assert math.isnan(c[0, 0]) and c[0, 1] == 3.0
    a, n = q(o.AGG_AVG); assert math.isnan(a[0, 0]) and abs(a[0, 1] - 5.133333333333333) < 1e-9 and n[0, 1] == 3   # the spec compares with |d| < error
    assert a[0, 1] == ((5.6 * 1 + 4.4 * 1) / 2 * 2 + 5.4 * 1) / 3                     # AvgRowAggregator.scala:38-46 running mean, exact
    vals, ids = q(o.AGG_BOTTOMK, k=2)
    assert all(i < 0 for i in ids[0, 0]) and sorted(v for v, i in zip(vals[0, 1], ids[0, 1]) if i >= 0) == [4.4, 5.4]
    vals, ids = q(o.AGG_TOPK, k=2)
    assert all(i < 0 for i in ids[0, 0]) and sorted(v for v, i in zip(vals[0, 1], ids[0, 1]) if i >= 0) == [5.4, 5.6]
