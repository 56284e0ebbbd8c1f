// sm_100a kernels: fused chunk decode + windowed range function (+ across-series aggregate), one warp per series.
// See scan_device.cuh for the per-window semantics and DESIGN.md for the layout / roofline discussion.
#include <cuda_runtime.h>
#include <stdint.h>
#define FILO_DEV_ERR_TS_WIRE 1
#define FILO_DEV_ERR_VAL_WIRE 2
#define FILO_DEV_ERR_EMPTY 3
#define FILO_DEV_ERR_SCRATCH 4
#include "scan_device.cuh"
#include "kernels.h"
#include "scan_fast.cuh"
#include "scan_tile.cuh"
#include "scan_wp.cuh"
#include "scan_wp_ctr.cuh"

namespace filo {

__device__ __forceinline__ void report_error(int* d_err, int code, int64_t series) {
  if (atomicCAS(&d_err[0], 0, code) == 0) { d_err[1] = (int)(series & 0x7fffffff); d_err[2] = (int)(series >> 31); }
}

// Resolve all chunks of one series that intersect [start - window, end] (TimeSeriesPartition.infos(start,end),
// TimeSeriesPartition.scala:365-366 + ChunkSetInfo.intersection :99-108).  Returns number of resolved chunks (D[0..n)).
__device__ __forceinline__ int resolve_series(const uint8_t* rec, const QueryParams& q, uint8_t* scratch, uint32_t scratch_bytes,
                                              bool need_corrected, int lane, int& err, int64_t& rows_scanned, int64_t& bytes_scanned) {
  const RecordHeader* h = reinterpret_cast<const RecordHeader*>(rec);
  const ChunkEntry* E = reinterpret_cast<const ChunkEntry*>(rec + sizeof(RecordHeader));
  const int nch = (int)h->n_chunks;
  const int64_t t1 = q.start - q.window, t2 = q.end;
  // chunks are time ordered: [cLo, cHi) = those with end_time >= t1 and start_time <= t2
  int cLo = 0; while (cLo < nch && E[cLo].end_time < t1) ++cLo;
  int cHi = cLo; while (cHi < nch && E[cHi].start_time <= t2) ++cHi;
  if (t1 > t2) cHi = cLo;
  const int n = cHi - cLo;
  err = 0;
  if (n <= 0) return 0;
  // scratch need: descriptors + decoded rows (host sized scratch_bytes for the worst series; double-check)
  ChunkDesc* D = reinterpret_cast<ChunkDesc*>(scratch);
  ScratchCursor sc; sc.p = scratch + align_up((uint32_t)n * (uint32_t)sizeof(ChunkDesc), 16);
  uint32_t need = (uint32_t)(sc.p - scratch) + (uint32_t)h->n_rows * 8u * (need_corrected ? 3u : 2u);
  if (need > scratch_bytes) { err = FILO_DEV_ERR_SCRATCH; return 0; }
  for (int c = 0; c < n; ++c) {
    int e = resolve_chunk(rec, &E[cLo + c], &D[c], sc, need_corrected, lane, false, q.long_values != 0);
    if (e) { err = e; return 0; }
  }
  // CountingChunkInfoIterator (ChunkSetInfo.scala:336-380): chunks pulled by the window iterator
  if (lane == 0) {
    const int64_t lastEnd = q.start + (int64_t)(q.T - 1) * q.step;
    int f = 0; while (f < n - 1 && D[f].end_time < lastEnd) ++f;
    for (int c = 0; c <= f; ++c) {
      rows_scanned += D[c].num_rows;
      bytes_scanned += (int64_t)ld32(rec + E[cLo + c].ts_off) + 4 + (int64_t)ld32(rec + E[cLo + c].val_off) + 4;
    }
  }
  return n;
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 1: PeriodicSamplesMapper without aggregate — out[series * T + k]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_series_kernel(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, int64_t n_series,
                   QueryParams q, double* __restrict__ out,
                   uint8_t* gscratch, uint32_t scratch_bytes, int use_smem,
                   unsigned long long* d_counters, int* d_err) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * SCAN_WARPS + warp, nw = (int64_t)gridDim.x * SCAN_WARPS;
  uint8_t* scratch = use_smem ? smem + (size_t)warp * scratch_bytes : gscratch + (size_t)gw * scratch_bytes;
  const int fn = q.fn;
  const bool need_corrected = ((fn == FN_RATE || fn == FN_INCREASE) && q.cumulative);
  int64_t rows = 0, bytes = 0;
  for (int64_t i = gw; i < n_series; i += nw) {
    const uint8_t* rec = arena + rec_off[i];
    int err;
    const int n = resolve_series(rec, q, scratch, scratch_bytes, need_corrected, lane, err, rows, bytes);
    if (err) { if (lane == 0) report_error(d_err, err, i); continue; }
    const ChunkDesc* D = reinterpret_cast<const ChunkDesc*>(scratch);
    double* o = out + (size_t)i * q.T;
    for (int k = lane; k < q.T; k += 32) o[k] = eval_window(D, 0, n, q, k);
    __syncwarp();
  }
  if (lane == 0 && (rows | bytes)) { atomicAdd(&d_counters[0], (unsigned long long)rows); atomicAdd(&d_counters[1], (unsigned long long)bytes); }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2: fused PeriodicSamplesMapper + AggregateMapReduce map/reduce phase.
// Work item = run of <= SEG consecutive series (in group-sorted order) of ONE group.  The warp folds the item's series
// into per-window accumulators (shared memory, or global scratch when T is large) and writes one partial row:
//   pval[item*T + k] = Σ non-NaN (SUM/AVG/COUNT) | min | max ;  pcnt[item*T + k] = number of non-NaN inputs
// A second kernel folds the partial rows of each group in item order (deterministic, atomics-free).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_agg_kernel(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, const int32_t* __restrict__ order,
                const int64_t* __restrict__ item_begin, int64_t n_items,
                QueryParams q, int agg_op, double* __restrict__ pval, uint32_t* __restrict__ pcnt,
                uint8_t* gscratch, uint32_t scratch_bytes, uint32_t acc_bytes, int use_smem,
                unsigned long long* d_counters, int* d_err) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * SCAN_WARPS + warp, nw = (int64_t)gridDim.x * SCAN_WARPS;
  const uint32_t per_warp = scratch_bytes + acc_bytes;
  uint8_t* base = use_smem ? smem + (size_t)warp * per_warp : gscratch + (size_t)gw * per_warp;
  double* acc = reinterpret_cast<double*>(base);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(base + (size_t)q.T * 8);
  uint8_t* scratch = base + acc_bytes;
  const int fn = q.fn;
  const bool need_corrected = ((fn == FN_RATE || fn == FN_INCREASE) && q.cumulative);
  const double ident = agg_op == AGG_MIN ? __longlong_as_double(0x7ff0000000000000LL)
                     : agg_op == AGG_MAX ? __longlong_as_double(0xfff0000000000000LL) : 0.0;
  int64_t rows = 0, bytes = 0;
  for (int64_t it = gw; it < n_items; it += nw) {
    for (int k = lane; k < q.T; k += 32) { acc[k] = ident; cnt[k] = 0; }
    const int64_t b = item_begin[it], e = item_begin[it + 1];
    for (int64_t pos = b; pos < e; ++pos) {
      const int64_t i = order ? order[pos] : pos;
      const uint8_t* rec = arena + rec_off[i];
      int err;
      const int n = resolve_series(rec, q, scratch, scratch_bytes, need_corrected, lane, err, rows, bytes);
      if (err) { if (lane == 0) report_error(d_err, err, i); continue; }
      const ChunkDesc* D = reinterpret_cast<const ChunkDesc*>(scratch);
      for (int k = lane; k < q.T; k += 32) {
        const double v = eval_window(D, 0, n, q, k);
        if (v == v) {                                       // RowAggregators skip NaN (SumRowAggregator.scala:22-29 ...)
          double a = acc[k];
          if (agg_op == AGG_MIN) a = v < a ? v : a;
          else if (agg_op == AGG_MAX) a = v > a ? v : a;
          else if (agg_op == AGG_COUNT) a = a;             // count only
          else a += v;
          acc[k] = a; cnt[k] += 1;
        }
      }
      __syncwarp();
    }
    double* pv = pval + (size_t)it * q.T; uint32_t* pc = pcnt + (size_t)it * q.T;
    for (int k = lane; k < q.T; k += 32) { pv[k] = acc[k]; pc[k] = cnt[k]; }
    __syncwarp();
  }
  if (lane == 0 && (rows | bytes)) { atomicAdd(&d_counters[0], (unsigned long long)rows); atomicAdd(&d_counters[1], (unsigned long long)bytes); }
}

// Fold the partial rows of each group (items [gis[g], gis[g+1])) in item order.  Block = (32 windows) x (8 item lanes);
// thread (kk, j) folds items j, j+8, ... sequentially, then the 8 lanes are folded in fixed order -> deterministic.
// partial_out: values/counts in mergeable form; otherwise presented (NaN when count == 0; Σ/n for AVG; n for COUNT).
__global__ void __launch_bounds__(256)
merge_partials_kernel(const double* __restrict__ pval, const uint32_t* __restrict__ pcnt, const int64_t* __restrict__ gis,
                      int n_groups, int T, int agg_op, int partial_out, double* __restrict__ out_val, int64_t* __restrict__ out_cnt) {
  __shared__ double sv[8][33]; __shared__ unsigned long long sc[8][33];
  const int kk = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int ktiles = (T + 31) / 32;
  const int g = blockIdx.x / ktiles, k = (blockIdx.x % ktiles) * 32 + kk;
  if (g >= n_groups) return;
  const double ident = agg_op == AGG_MIN ? __longlong_as_double(0x7ff0000000000000LL)
                     : agg_op == AGG_MAX ? __longlong_as_double(0xfff0000000000000LL) : 0.0;
  double a = ident; unsigned long long c = 0;
  if (k < T) {
    for (int64_t it = gis[g] + j; it < gis[g + 1]; it += 8) {
      const double v = pval[(size_t)it * T + k]; const uint32_t n = pcnt[(size_t)it * T + k];
      if (n) {
        if (agg_op == AGG_MIN) a = v < a ? v : a; else if (agg_op == AGG_MAX) a = v > a ? v : a; else a += v;
        c += n;
      }
    }
  }
  sv[j][kk] = a; sc[j][kk] = c;
  __syncthreads();
  if (j == 0 && k < T) {
    for (int jj = 1; jj < 8; ++jj) {
      const double v = sv[jj][kk]; const unsigned long long n = sc[jj][kk];
      if (n) { if (agg_op == AGG_MIN) a = v < a ? v : a; else if (agg_op == AGG_MAX) a = v > a ? v : a; else a += v; c += n; }
    }
    const size_t o = (size_t)g * T + k;
    if (partial_out) { out_val[o] = a; if (out_cnt) out_cnt[o] = (int64_t)c; }
    else {
      const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
      double r;
      if (c == 0) r = NaNv;
      else if (agg_op == AGG_AVG) r = a / (double)c;
      else if (agg_op == AGG_COUNT) r = (double)c;
      else r = a;
      out_val[o] = r; if (out_cnt) out_cnt[o] = (int64_t)c;
    }
  }
}

// present after a cross-GPU merge of partials
__global__ void present_kernel(int agg_op, int64_t n, const double* __restrict__ vals, const int64_t* __restrict__ cnts, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = cnts[i]; const double v = vals[i];
  double r;
  if (c == 0) r = __longlong_as_double(0x7ff8000000000000LL);
  else if (agg_op == AGG_AVG) r = v / (double)c;
  else if (agg_op == AGG_COUNT) r = (double)c;
  else r = v;
  out[i] = r;
}

// topk / bottomk over per-series results (TopBottomKRowAggregator.scala:84-95): one thread per (group, window) scans the
// group's series in sorted (arrival) order keeping the k best non-NaN values; ties keep the earlier series.
// Output row in the reference's dequeue order: topk ascending, bottomk descending; empty slots = ±Double.MaxValue, id -1.
__global__ void topk_kernel(const double* __restrict__ per_series, const int32_t* __restrict__ order, const int64_t* __restrict__ group_start,
                            int n_groups, int T, int kk, int bottom, double* __restrict__ out_val, int64_t* __restrict__ out_id) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (int64_t)n_groups * T) return;
  const int g = (int)(tid / T), t = (int)(tid % T);
  double bv[FILO_MAX_TOPK]; int64_t bi[FILO_MAX_TOPK]; int n = 0;    // sorted best-first
  for (int64_t pos = group_start[g]; pos < group_start[g + 1]; ++pos) {
    const int64_t s = order ? order[pos] : pos;
    const double v = per_series[(size_t)s * T + t];
    if (v != v) continue;
    // insertion position: after all elements at least as good (ties keep earlier arrival ahead)
    int p = n;
    while (p > 0 && (bottom ? (v < bv[p - 1]) : (v > bv[p - 1]))) --p;
    if (p >= kk) continue;
    const int last = n < kk ? n : kk - 1;
    for (int m = last; m > p; --m) { bv[m] = bv[m - 1]; bi[m] = bi[m - 1]; }
    bv[p] = v; bi[p] = s; if (n < kk) ++n;
  }
  double* ov = out_val + (size_t)tid * kk; int64_t* oi = out_id + (size_t)tid * kk;
  for (int m = 0; m < kk; ++m) {
    if (m < n) { ov[m] = bv[n - 1 - m]; oi[m] = bi[n - 1 - m]; }          // worst of the kept first (dequeue order)
    else { ov[m] = bottom ? 1.7976931348623157e308 : -1.7976931348623157e308; oi[m] = -1; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// group bookkeeping (device side, so that synthetic tables never leave the GPU)
// ---------------------------------------------------------------------------------------------------------------
__global__ void iota_kernel(int32_t* a, int64_t n) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = (int32_t)i; }
// group_start[g] = lower_bound(sorted_keys, g)
__global__ void group_bounds_kernel(const int32_t* __restrict__ sorted_keys, int64_t n, int n_groups, int64_t* __restrict__ group_start) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_groups) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (sorted_keys[m] < g) lo = m + 1; else hi = m; }
  group_start[g] = lo;
}
// items per group -> gis (exclusive scan done by caller with cub); fill item_begin
__global__ void group_item_count_kernel(const int64_t* __restrict__ group_start, int n_groups, int seg, int64_t* __restrict__ cnt) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) { int64_t n = group_start[g + 1] - group_start[g]; cnt[g] = (n + seg - 1) / seg; }
}
__global__ void fill_items_kernel(const int64_t* __restrict__ group_start, const int64_t* __restrict__ gis, int n_groups, int seg,
                                  int64_t n_items, int64_t n_series, int64_t* __restrict__ item_begin) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) {
    const int64_t b = group_start[g], e = group_start[g + 1];
    int64_t it = gis[g];
    for (int64_t p = b; p < e; p += seg) item_begin[it++] = p;
  }
  if (g == 0) item_begin[n_items] = n_series;
}

// ---------------------------------------------------------------------------------------------------------------
// v2 kernels (scan_fast.cuh): TMA-staged records + blocked window reductions.  Per-warp shared memory:
//   [mbarrier 16 B][record staging buffer rec_cap][result transpose stage][accumulators (agg only)][decode scratch]
// One staging buffer per warp: the record is dead as soon as its chunks are resolved into scratch, so the bulk copy of the
// NEXT series is issued right after resolve and overlaps the whole window phase of the current one.
// ---------------------------------------------------------------------------------------------------------------
struct WarpStage {
  uint64_t* bar; uint8_t* buf; uint32_t cap; uint32_t parity; bool staged;
  const uint8_t* arena; const int64_t* rec_off;
  __device__ __forceinline__ void issue(int64_t series, int lane) {        // all lanes call; lane 0 issues
    const int64_t o = rec_off[series];
    const uint32_t bytes = (uint32_t)(rec_off[series + 1] - o);
    staged = cap != 0 && bytes <= cap;
    if (staged && lane == 0) { mbar_expect_tx(bar, bytes); tma_load_1d(buf, arena + o, bytes, bar); }
  }
  __device__ __forceinline__ const uint8_t* acquire(int64_t series) {       // record of `series`, previously issued
    if (!staged) return arena + rec_off[series];
    mbar_wait(bar, parity); parity ^= 1;
    return buf;
  }
};

template <int CLS>
__global__ void __launch_bounds__(FAST_WARPS * 32, FAST_MIN_CTAS)
scan_series_kernel_v2(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, int64_t n_series,
                      QueryParams q, double* __restrict__ out, uint32_t rec_cap, uint32_t scratch_bytes,
                      unsigned long long* d_counters, int* d_err,
                      const int64_t* __restrict__ list, const unsigned long long* __restrict__ list_count) {
  if (list) n_series = (int64_t)*list_count;             // fallback pass of the tile kernel: only the listed series
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * FAST_WARPS + warp, nw = (int64_t)gridDim.x * FAST_WARPS;
  const uint32_t per_warp = WARP_HDR_BYTES + rec_cap + STAGE_BYTES + scratch_bytes;
  uint8_t* base = smem + (size_t)warp * per_warp;
  WarpStage st{reinterpret_cast<uint64_t*>(base), base + WARP_HDR_BYTES, rec_cap, 0, false, arena, rec_off};
  double* stage = reinterpret_cast<double*>(base + WARP_HDR_BYTES + rec_cap);
  uint8_t* scratch = base + WARP_HDR_BYTES + rec_cap + STAGE_BYTES;
  if (lane == 0) { mbar_init(st.bar, 1); mbar_fence_init(); }
  __syncwarp();
  int64_t rows = 0, bytes = 0;
  int64_t ii = gw;
  if (ii < n_series) st.issue(list ? list[ii] : ii, lane);
  for (; ii < n_series; ii += nw) {
    const int64_t i = list ? list[ii] : ii;
    const uint8_t* rec = st.acquire(i);
    double* o = out + (size_t)i * q.T;
    int err;
    const int64_t inext_i = ii + nw;
    const bool has_next = inext_i < n_series;
    const int64_t inext = has_next ? (list ? list[inext_i] : inext_i) : 0;
    process_series<CLS>(rec, q, scratch, scratch_bytes, stage, lane, err, rows, bytes,
                   [&](int k, double v, bool valid) { if (valid) o[k] = v; },
                   [&]() { if (has_next) st.issue(inext, lane); });
    if (err) {
      if (lane == 0) report_error(d_err, err, i);
      __syncwarp();
      if (has_next) st.issue(inext, lane);      // process_series returned before its release hook ran
    }
    __syncwarp();
  }
  if (lane == 0 && (rows | bytes)) { atomicAdd(&d_counters[0], (unsigned long long)rows); atomicAdd(&d_counters[1], (unsigned long long)bytes); }
}

template <int CLS>
__global__ void __launch_bounds__(FAST_WARPS * 32, FAST_MIN_CTAS)
scan_agg_kernel_v2(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, const int32_t* __restrict__ order,
                   const int64_t* __restrict__ item_begin, int64_t n_items,
                   QueryParams q, int agg_op, double* __restrict__ pval, uint32_t* __restrict__ pcnt,
                   uint32_t rec_cap, uint32_t scratch_bytes, uint32_t acc_bytes,
                   unsigned long long* d_counters, int* d_err,
                   const int64_t* __restrict__ list, const unsigned long long* __restrict__ list_count) {
  if (list) n_items = (int64_t)*list_count;              // fallback pass of the tile kernel: only the listed items
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * FAST_WARPS + warp, nw = (int64_t)gridDim.x * FAST_WARPS;
  const uint32_t per_warp = WARP_HDR_BYTES + rec_cap + STAGE_BYTES + acc_bytes + scratch_bytes;
  uint8_t* base = smem + (size_t)warp * per_warp;
  WarpStage st{reinterpret_cast<uint64_t*>(base), base + WARP_HDR_BYTES, rec_cap, 0, false, arena, rec_off};
  double* stage = reinterpret_cast<double*>(base + WARP_HDR_BYTES + rec_cap);
  double* acc = reinterpret_cast<double*>(base + WARP_HDR_BYTES + rec_cap + STAGE_BYTES);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(base + WARP_HDR_BYTES + rec_cap + STAGE_BYTES + (size_t)q.T * 8);
  uint8_t* scratch = base + WARP_HDR_BYTES + rec_cap + STAGE_BYTES + acc_bytes;
  if (lane == 0) { mbar_init(st.bar, 1); mbar_fence_init(); }
  __syncwarp();
  const double ident = agg_op == AGG_MIN ? __longlong_as_double(0x7ff0000000000000LL)
                     : agg_op == AGG_MAX ? __longlong_as_double(0xfff0000000000000LL) : 0.0;
  int64_t rows = 0, bytes = 0;
  // flattened (item, position) walk so that the next series to prefetch is always known
  int64_t it = gw;
  int64_t pos = 0, pend = 0;
  auto real_item = [&](int64_t x) -> int64_t { return list ? list[x] : x; };
  auto advance_item = [&]() { while (it < n_items) { const int64_t ri = real_item(it); pos = item_begin[ri]; pend = item_begin[ri + 1]; if (pos < pend) return true; it += nw; } return false; };
  bool have = advance_item();
  if (have) st.issue(order ? order[pos] : pos, lane);
  while (have) {
    for (int k = lane; k < q.T; k += 32) { acc[k] = ident; cnt[k] = 0; }
    __syncwarp();
    const int64_t my_item = real_item(it);
    while (true) {
      const int64_t i = order ? order[pos] : pos;
      // next series in walk order
      int64_t npos = pos + 1, nit = it, npend = pend; bool nhave = true;
      if (npos >= pend) { nit = it + nw; nhave = false; while (nit < n_items) { const int64_t ri = real_item(nit); npos = item_begin[ri]; npend = item_begin[ri + 1]; if (npos < npend) { nhave = true; break; } nit += nw; } }
      const int64_t inext = nhave ? (order ? order[npos] : npos) : -1;
      const uint8_t* rec = st.acquire(i);
      int err;
      process_series<CLS>(rec, q, scratch, scratch_bytes, stage, lane, err, rows, bytes,
                     [&](int k, double v, bool valid) {
                       if (valid && v == v) {              // RowAggregators skip NaN (SumRowAggregator.scala:22-29 ...)
                         double a = acc[k];
                         if (agg_op == AGG_MIN) a = v < a ? v : a;
                         else if (agg_op == AGG_MAX) a = v > a ? v : a;
                         else if (agg_op != AGG_COUNT) a += v;
                         acc[k] = a; cnt[k] += 1;
                       }
                     },
                     [&]() { if (inext >= 0) st.issue(inext, lane); });
      if (err) {
        if (lane == 0) report_error(d_err, err, i);
        __syncwarp();
        if (inext >= 0) st.issue(inext, lane);
      }
      __syncwarp();
      const bool same_item = nhave && nit == it;
      pos = npos; pend = npend; it = nit; have = nhave;
      if (!same_item) break;
    }
    double* pv = pval + (size_t)my_item * q.T; uint32_t* pc = pcnt + (size_t)my_item * q.T;
    for (int k = lane; k < q.T; k += 32) { pv[k] = acc[k]; pc[k] = cnt[k]; }
    __syncwarp();
  }
  if (lane == 0 && (rows | bytes)) { atomicAdd(&d_counters[0], (unsigned long long)rows); atomicAdd(&d_counters[1], (unsigned long long)bytes); }
}

#ifndef FILO_CUSIM      // the launchers need nvcc; the emulation build (tests/cpp) calls the kernels through cusim::launch
#ifdef FILO_TILE_PROF
extern "C" int filo_debug_tile_prof(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_tile_prof, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset) { unsigned long long z[16] = {}; e = cudaMemcpyToSymbol(g_tile_prof, z, sizeof z); }
  return (int)e;
}
#endif
// ---------------------------------------------------------------------------------------------------------------
// host-callable launchers
// ---------------------------------------------------------------------------------------------------------------
template <int CLS>
static cudaError_t launch_series_v2_cls(const ScanLaunch& L, double* out, uint32_t rec_cap, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(scan_series_kernel_v2<CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan_series_kernel_v2<CLS><<<L.grid, FAST_WARPS * 32, smem, L.stream>>>(L.arena, L.rec_off, L.n_series, L.q, out, rec_cap, L.scratch_bytes,
                                                                           L.d_counters, L.d_err, L.list, L.list_count);
  return cudaGetLastError();
}
cudaError_t launch_scan_series_v2(const ScanLaunch& L, double* out, uint32_t rec_cap) {
  const size_t smem = (size_t)(WARP_HDR_BYTES + rec_cap + STAGE_BYTES + L.scratch_bytes) * FAST_WARPS;
  switch (fn_class_of(L.q.fn, L.q.cumulative, L.q.long_values)) {
    case CLASS_SUM: return launch_series_v2_cls<CLASS_SUM>(L, out, rec_cap, smem);
    case CLASS_MINMAX: return launch_series_v2_cls<CLASS_MINMAX>(L, out, rec_cap, smem);
    case CLASS_COUNTER: return launch_series_v2_cls<CLASS_COUNTER>(L, out, rec_cap, smem);
    default: return launch_series_v2_cls<CLASS_POINT>(L, out, rec_cap, smem);
  }
}
template <int CLS>
static cudaError_t launch_agg_v2_cls(const ScanLaunch& L, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                                     double* pval, uint32_t* pcnt, uint32_t acc_bytes, uint32_t rec_cap, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(scan_agg_kernel_v2<CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan_agg_kernel_v2<CLS><<<L.grid, FAST_WARPS * 32, smem, L.stream>>>(L.arena, L.rec_off, order, item_begin, n_items, L.q, agg_op, pval, pcnt,
                                                                        rec_cap, L.scratch_bytes, acc_bytes, L.d_counters, L.d_err, L.list, L.list_count);
  return cudaGetLastError();
}
cudaError_t launch_scan_agg_v2(const ScanLaunch& L, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                               double* pval, uint32_t* pcnt, uint32_t acc_bytes, uint32_t rec_cap) {
  const size_t smem = (size_t)(WARP_HDR_BYTES + rec_cap + STAGE_BYTES + acc_bytes + L.scratch_bytes) * FAST_WARPS;
  switch (fn_class_of(L.q.fn, L.q.cumulative, L.q.long_values)) {
    case CLASS_SUM: return launch_agg_v2_cls<CLASS_SUM>(L, order, item_begin, n_items, agg_op, pval, pcnt, acc_bytes, rec_cap, smem);
    case CLASS_MINMAX: return launch_agg_v2_cls<CLASS_MINMAX>(L, order, item_begin, n_items, agg_op, pval, pcnt, acc_bytes, rec_cap, smem);
    case CLASS_COUNTER: return launch_agg_v2_cls<CLASS_COUNTER>(L, order, item_begin, n_items, agg_op, pval, pcnt, acc_bytes, rec_cap, smem);
    default: return launch_agg_v2_cls<CLASS_POINT>(L, order, item_begin, n_items, agg_op, pval, pcnt, acc_bytes, rec_cap, smem);
  }
}
struct TileAggArgs { const int32_t* order; const int64_t* item_begin; int64_t n_items; int agg_op; double* pval; uint32_t* pcnt; };
template <int CLS, int FN, bool AGG, int DEC>
static cudaError_t launch_tile_dec(const ScanLaunch& L, double* out, const TileSmem& T, int64_t* fallback_list, unsigned long long* fallback_count,
                                   const TileAggArgs& A) {
  cudaError_t e = cudaFuncSetAttribute(scan_tile_kernel<CLS, FN, AGG, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T.total);
  if (e != cudaSuccess) return e;
  scan_tile_kernel<CLS, FN, AGG, DEC><<<L.grid, TILE_LAUNCH_THREADS, T.total, L.stream>>>(L.arena, L.rec_off, L.n_series, L.q, out, T, fallback_list, fallback_count,
      L.d_counters, L.d_err, A.order, A.item_begin, A.n_items, A.agg_op, A.pval, A.pcnt);
  return cudaGetLastError();
}
template <int CLS, int FN, bool AGG>
static cudaError_t launch_tile_fn(const ScanLaunch& L, double* out, const TileSmem& T, int64_t* fallback_list, unsigned long long* fallback_count,
                                  const TileAggArgs& A) {
  if (T.opts & TILE_OPT_WARPDEC) return launch_tile_dec<CLS, FN, AGG, 1>(L, out, T, fallback_list, fallback_count, A);      // experimental per-warp decode
  return launch_tile_dec<CLS, FN, AGG, 0>(L, out, T, fallback_list, fallback_count, A);
}
template <bool AGG>
static cudaError_t launch_tile_any(const ScanLaunch& L, double* out, const TileSmem& T, int64_t* fallback_list, unsigned long long* fallback_count,
                                   const TileAggArgs& A) {
  if (fn_class_of(L.q.fn, L.q.cumulative, L.q.long_values) == CLASS_COUNTER) {
    switch (L.q.fn) {
      case FN_RATE: return launch_tile_fn<CLASS_COUNTER, FN_RATE, AGG>(L, out, T, fallback_list, fallback_count, A);
      case FN_INCREASE: return launch_tile_fn<CLASS_COUNTER, FN_INCREASE, AGG>(L, out, T, fallback_list, fallback_count, A);
      default: return launch_tile_fn<CLASS_COUNTER, FN_DELTA, AGG>(L, out, T, fallback_list, fallback_count, A);
    }
  }
  switch (L.q.fn) {
    case FN_RATE: return launch_tile_fn<CLASS_SUM, FN_RATE, AGG>(L, out, T, fallback_list, fallback_count, A);
    case FN_AVG: return launch_tile_fn<CLASS_SUM, FN_AVG, AGG>(L, out, T, fallback_list, fallback_count, A);
    case FN_COUNT: return launch_tile_fn<CLASS_SUM, FN_COUNT, AGG>(L, out, T, fallback_list, fallback_count, A);
    default: return launch_tile_fn<CLASS_SUM, FN_SUM, AGG>(L, out, T, fallback_list, fallback_count, A);      // FN_SUM, FN_INCREASE on a delta schema
  }
}
cudaError_t launch_scan_tile(const ScanLaunch& L, double* out, const TileSmem& T, int64_t* fallback_list, unsigned long long* fallback_count) {
  return launch_tile_any<false>(L, out, T, fallback_list, fallback_count, TileAggArgs{nullptr, nullptr, 0, 0, nullptr, nullptr});
}
// fused across-series aggregate: one partial row per item (same contract as launch_scan_agg_v2); items with an irregular series
// are appended to fallback_list
cudaError_t launch_scan_tile_agg(const ScanLaunch& L, const TileSmem& T, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                                 double* pval, uint32_t* pcnt, int64_t* fallback_list, unsigned long long* fallback_count) {
  return launch_tile_any<true>(L, nullptr, T, fallback_list, fallback_count, TileAggArgs{order, item_begin, n_items, agg_op, pval, pcnt});
}
// v4 warp-pipeline kernel (scan_wp.cuh): one CTA of W.warps warps per SM; declined series go to fallback_list
template <int FN, int NW>
static cudaError_t launch_wp_nw(const ScanLaunch& L, double* out, const WpSmem& W, int64_t* fallback_list, unsigned long long* fallback_count) {
  const size_t smem = (size_t)W.per_warp * W.warps;
  cudaError_t e = cudaFuncSetAttribute(scan_wp_sum_kernel<FN, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan_wp_sum_kernel<FN, NW><<<L.grid, W.warps * 32, smem, L.stream>>>(L.arena, L.rec_off, L.n_series, L.q, out, W, fallback_list, fallback_count, L.d_counters, L.d_err);
  return cudaGetLastError();
}
template <int FN>
static cudaError_t launch_wp_fn(const ScanLaunch& L, double* out, const WpSmem& W, int64_t* fallback_list, unsigned long long* fallback_count) {
  // the register budget follows the warps per CTA: 128 registers up to 16 warps, 96 up to 20
  if (W.warps <= (uint32_t)WP_MAX_WARPS) return launch_wp_nw<FN, WP_MAX_WARPS>(L, out, W, fallback_list, fallback_count);
  return launch_wp_nw<FN, WP_MAX_WARPS_ALIAS>(L, out, W, fallback_list, fallback_count);
}
cudaError_t launch_scan_wp(const ScanLaunch& L, double* out, const WpSmem& W, int64_t* fallback_list, unsigned long long* fallback_count) {
  switch (L.q.fn) {
    case FN_RATE: return launch_wp_fn<FN_RATE>(L, out, W, fallback_list, fallback_count);
    case FN_AVG: return launch_wp_fn<FN_AVG>(L, out, W, fallback_list, fallback_count);
    case FN_COUNT: return launch_wp_fn<FN_COUNT>(L, out, W, fallback_list, fallback_count);
    default: return launch_wp_fn<FN_SUM>(L, out, W, fallback_list, fallback_count);      // FN_SUM, FN_INCREASE on a delta schema
  }
}
// v4 counter-class kernel (scan_wp_ctr.cuh): per-series rows (order == nullptr, n_items == 0) or one partial row per work item
template <int FN, bool AGG, int NW, bool IRR>
static cudaError_t launch_wp_ctr_nw(const ScanLaunch& L, double* out, const WpCtrSmem& W, int64_t* fallback_list, unsigned long long* fallback_count,
                                    const TileAggArgs& A) {
  const size_t smem = (size_t)W.per_warp * W.warps + sizeof(TileCtrTab) * (TILE_CTR_TABMAX + 1);
  cudaError_t e = cudaFuncSetAttribute(scan_wp_ctr_kernel<FN, AGG, NW, IRR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scan_wp_ctr_kernel<FN, AGG, NW, IRR><<<L.grid, W.warps * 32, smem, L.stream>>>(L.arena, L.rec_off, L.n_series, L.q, out, W, fallback_list, fallback_count,
      L.d_counters, L.d_err, A.order, A.item_begin, A.n_items, A.agg_op, A.pval, A.pcnt);
  return cudaGetLastError();
}
template <int FN, bool AGG>
static cudaError_t launch_wp_ctr_fn(const ScanLaunch& L, double* out, const WpCtrSmem& W, int64_t* fallback_list, unsigned long long* fallback_count, const TileAggArgs& A) {
  if (W.tsr != 0) return launch_wp_ctr_nw<FN, AGG, 16, true>(L, out, W, fallback_list, fallback_count, A);      // irregular timestamps: the larger shared-memory footprint keeps it at <= 16 warps
  if (W.warps <= 16) return launch_wp_ctr_nw<FN, AGG, 16, false>(L, out, W, fallback_list, fallback_count, A);
  return launch_wp_ctr_nw<FN, AGG, WP_CTR_MAX_WARPS, false>(L, out, W, fallback_list, fallback_count, A);
}
template <bool AGG>
static cudaError_t launch_wp_ctr_any(const ScanLaunch& L, double* out, const WpCtrSmem& W, int64_t* fallback_list, unsigned long long* fallback_count, const TileAggArgs& A) {
  switch (L.q.fn) {
    case FN_RATE: return launch_wp_ctr_fn<FN_RATE, AGG>(L, out, W, fallback_list, fallback_count, A);
    case FN_INCREASE: return launch_wp_ctr_fn<FN_INCREASE, AGG>(L, out, W, fallback_list, fallback_count, A);
    default: return launch_wp_ctr_fn<FN_DELTA, AGG>(L, out, W, fallback_list, fallback_count, A);
  }
}
cudaError_t launch_scan_wp_ctr(const ScanLaunch& L, double* out, const WpCtrSmem& W, int64_t* fallback_list, unsigned long long* fallback_count) {
  return launch_wp_ctr_any<false>(L, out, W, fallback_list, fallback_count, TileAggArgs{nullptr, nullptr, 0, 0, nullptr, nullptr});
}
cudaError_t launch_scan_wp_ctr_agg(const ScanLaunch& L, const WpCtrSmem& W, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                                   double* pval, uint32_t* pcnt, int64_t* fallback_list, unsigned long long* fallback_count) {
  return launch_wp_ctr_any<true>(L, nullptr, W, fallback_list, fallback_count, TileAggArgs{order, item_begin, n_items, agg_op, pval, pcnt});
}
size_t v2_smem_per_warp(uint32_t rec_cap, uint32_t scratch_bytes, uint32_t acc_bytes) { return WARP_HDR_BYTES + (size_t)rec_cap + STAGE_BYTES + acc_bytes + scratch_bytes; }
cudaError_t launch_scan_series(const ScanLaunch& L, double* out) {
  size_t smem = L.use_smem ? (size_t)L.scratch_bytes * SCAN_WARPS : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(scan_series_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  scan_series_kernel<<<L.grid, SCAN_WARPS * 32, smem, L.stream>>>(L.arena, L.rec_off, L.n_series, L.q, out,
      L.gscratch, L.scratch_bytes, L.use_smem, L.d_counters, L.d_err);
  return cudaGetLastError();
}
cudaError_t launch_scan_agg(const ScanLaunch& L, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                            double* pval, uint32_t* pcnt, uint32_t acc_bytes) {
  size_t smem = L.use_smem ? (size_t)(L.scratch_bytes + acc_bytes) * SCAN_WARPS : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(scan_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  scan_agg_kernel<<<L.grid, SCAN_WARPS * 32, smem, L.stream>>>(L.arena, L.rec_off, order, item_begin, n_items, L.q, agg_op,
      pval, pcnt, L.gscratch, L.scratch_bytes, acc_bytes, L.use_smem, L.d_counters, L.d_err);
  return cudaGetLastError();
}
cudaError_t launch_merge_partials(const double* pval, const uint32_t* pcnt, const int64_t* gis, int n_groups, int T, int agg_op,
                                  int partial_out, double* out_val, int64_t* out_cnt, cudaStream_t s) {
  const int ktiles = (T + 31) / 32;
  merge_partials_kernel<<<n_groups * ktiles, 256, 0, s>>>(pval, pcnt, gis, n_groups, T, agg_op, partial_out, out_val, out_cnt);
  return cudaGetLastError();
}
cudaError_t launch_present(int agg_op, int64_t n, const double* vals, const int64_t* cnts, double* out, cudaStream_t s) {
  present_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(agg_op, n, vals, cnts, out);
  return cudaGetLastError();
}
cudaError_t launch_topk(const double* per_series, const int32_t* order, const int64_t* group_start, int n_groups, int T, int k, int bottom,
                        double* out_val, int64_t* out_id, cudaStream_t s) {
  const int64_t n = (int64_t)n_groups * T;
  topk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(per_series, order, group_start, n_groups, T, k, bottom, out_val, out_id);
  return cudaGetLastError();
}
cudaError_t launch_iota(int32_t* a, int64_t n, cudaStream_t s) {
  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, n); return cudaGetLastError();
}
cudaError_t launch_group_bounds(const int32_t* sorted_keys, int64_t n, int n_groups, int64_t* group_start, cudaStream_t s) {
  group_bounds_kernel<<<(n_groups + 1 + 127) / 128, 128, 0, s>>>(sorted_keys, n, n_groups, group_start); return cudaGetLastError();
}
cudaError_t launch_group_item_count(const int64_t* group_start, int n_groups, int seg, int64_t* cnt, cudaStream_t s) {
  group_item_count_kernel<<<(n_groups + 127) / 128, 128, 0, s>>>(group_start, n_groups, seg, cnt); return cudaGetLastError();
}
cudaError_t launch_fill_items(const int64_t* group_start, const int64_t* gis, int n_groups, int seg, int64_t n_items, int64_t n_series,
                              int64_t* item_begin, cudaStream_t s) {
  fill_items_kernel<<<(n_groups + 127) / 128, 128, 0, s>>>(group_start, gis, n_groups, seg, n_items, n_series, item_begin);
  return cudaGetLastError();
}

#endif // FILO_CUSIM

} // namespace filo
