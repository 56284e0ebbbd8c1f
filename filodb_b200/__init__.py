"""filodb_b200 — B200-native chunk scan + PromQL range-vector aggregation for FiloDB (one hot path, nothing else).

The product is the C-ABI shared library `libfilo_b200.so` (include/filo_b200.h) holding hand-written sm_100a kernels.
This package is the thin host-side mirror used by the tests and bench: `capi` binds the C-ABI with ctypes, `exec`
mirrors the reference's operator surface (PeriodicSamplesMapper / AggregateMapReduce / RangeVectorTransformer).
There is no CPU fallback: importing `capi` without the built library raises.
"""
from . import capi  # noqa: F401
from . import exec as exec_  # noqa: F401  (operator mirror: PeriodicSamplesMapper / AggregateMapReduce / FusedGpuExec)
