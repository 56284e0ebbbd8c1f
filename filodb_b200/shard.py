"""Host-side multi-GPU logic: FiloDB shard -> GPU assignment and the one cross-shard exchange of an aggregate query.

One process per GPU.  Series are independent until the across-series aggregate, so the data path has no
collective for per-series queries; aggregates merge `[G x T]` partials (FILO_Q_PARTIAL form, include/filo_b200.h) with
one all-reduce, the role `LocalPartitionReduceAggregateExec` + `RowAggregator.reduceAggregate` play in the reference
(query/exec/AggrOverRangeVectors.scala:119-182; aggregator/*RowAggregator.scala).

Works on CUDA tensors over NCCL (product) and on CPU tensors over gloo (tests/test_multi_gpu_gloo.py).
"""
from __future__ import annotations

AGG_NONE, AGG_SUM, AGG_AVG, AGG_MIN, AGG_MAX, AGG_COUNT, AGG_TOPK, AGG_BOTTOMK = range(8)


def shards_of_rank(num_shards: int, rank: int, world: int) -> list[int]:
    """GPU g owns shards {s : s mod nGPU == g}: FiloDB's spread bits are the upper bits of the shard number
    (coordinator/ShardMapper.scala:26-48,93-102), so the modulo spreads one shard key's shards over all GPUs."""
    if num_shards & (num_shards - 1):
        raise ValueError("numShards must be a power of two (ShardMapper.scala:26-31)")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return [s for s in range(num_shards) if s % world == rank]


def series_range_of_rank(n_series_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) series-id range of a rank when series (not shards) are split evenly (synthetic bench)."""
    per = (n_series_total + world - 1) // world
    b = min(n_series_total, rank * per)
    return b, min(n_series_total, b + per)


def merge_partials(values, counts, aggr_op: int, dist) -> None:
    """In-place cross-rank merge of FILO_Q_PARTIAL results: values [G*T] f64, counts [G*T] i64.
    sum/avg/count: Σ values, Σ counts; min/max: min/max of values (identity ±Inf), Σ counts."""
    if aggr_op in (AGG_SUM, AGG_AVG, AGG_COUNT):
        dist.all_reduce(values, op=dist.ReduceOp.SUM)
    elif aggr_op == AGG_MIN:
        dist.all_reduce(values, op=dist.ReduceOp.MIN)
    elif aggr_op == AGG_MAX:
        dist.all_reduce(values, op=dist.ReduceOp.MAX)
    else:
        raise ValueError("aggregate %d has no all-reduce merge (topk merges gathered candidates)" % aggr_op)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)


def max_over_ranks(x: float, dist, device) -> float:
    """Device-timed durations are reported as the max over ranks."""
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
