"""Host-side helpers that restate reference encodings outside the CUDA library (no GPU needed): bucket definitions built in Python for the
histogram entry points, and the numpy models of the device generators that bench.py / the tests rebuild oracle inputs from."""
import numpy as np
import pytest


def test_bucket_definitions_match_the_reference_serialisation(oracle):
    # HistogramBuckets.serialize: GeometricBuckets Histogram.scala:609-617, CustomBuckets :878-884 (NibblePack.packDoubles of the tops)
    from filodb_b200 import capi
    from oracle import hist as H
    for first, mult, n, minus_one in ((2.0, 3.0, 20, False), (1.0, 2.0, 8, False), (0.5, 1.5, 64, False)):
        d, fmt = capi.geometric_bucket_def(first, mult, n)
        assert fmt == 3 and (d == H.Buckets.geometric(first, mult, n).serialize()).all()
    rng = np.random.default_rng(3)
    for les in ([2.0 * 3 ** i for i in range(19)] + [float("inf")], [0.5 * 2 ** i for i in range(12)] + [float("inf")], [1.0, 2.5, 7.25],
                list(np.cumsum(rng.random(33)) * 1e3), [5.0]):
        d, fmt = capi.custom_bucket_def(les)
        ref = H.Buckets.custom(les).serialize()
        assert fmt == 5 and d.size == ref.size and (d == ref).all(), les


def test_numpy_generator_models_agree():
    # bench.gen_hist_series_np (vectorised) == tests/synth_ref.gen_hist_series (row by row): both model hist_row in synth_kernels.cu
    import bench
    from tests import synth_ref as sr
    for seed, gid, rows, nb, reset in ((42, 0, 480, 20, 97), (42, 97, 480, 20, 97), (5, 194, 230, 13, 97), (7, 12345, 64, 8, 0)):
        assert (bench.gen_hist_series_np(seed, gid, rows, nb, reset) == sr.gen_hist_series(seed, gid, rows, nb, reset)).all()
    g = bench.synth_group_ids(42, 1000, 64, 7)
    assert list(g) == [sr.group_id(42, 1000 + i, 7) for i in range(64)]
