// v2 scan path: TMA-staged records, register-blocked window reductions, lane-coalesced result sink.
//
// Per series (one warp):
//   1. the record (chunk pages) arrives in shared memory by one cp.async.bulk (TMA 1-D bulk copy) that was issued while the
//      previous series was being reduced (one staging buffer + one mbarrier per warp);
//   2. chunks are resolved/decoded into per-warp scratch (scan_device.cuh); afterwards the staging buffer is dead and the
//      bulk copy of the NEXT series is issued;
//   3. "single-chunk" windows — windows whose contributing rows all belong to ONE chunk with const-DDV timestamps whose slope
//      equals the query step (the vast majority in practice; the row range may be clamped by the chunk's first/last row) —
//      are reduced BLK_R at a time per lane: the lane walks rows s0 .. s0+Wr+R-1 once and feeds each row to the accumulators
//      of the windows containing it.  Every accumulator still sees its rows in row order starting from 0.0, i.e. exactly the
//      reference's sequential sum (DoubleVector.scala:243-253), but a row is loaded once per R windows.  Rows outside the
//      chunk are fed as +0.0, which is an exact no-op for an accumulator that started at +0.0.  R = 15 (odd): lane base rows
//      are 15 apart, so the 64-bit shared-memory reads of a warp are bank-conflict-free in the plain linear layout;
//   4. the remaining windows (rows from two chunks, irregular timestamps) are enumerated densely and take the general path
//      (eval_window), which restates the reference state machine literally.
// Results leave through a Sink called by all lanes (lane-consecutive windows on the bulk path: coalesced stores).
#pragma once
#include "scan_device.cuh"

namespace filo {

constexpr int BLK_R = 15;                // windows per lane in the blocked reduction (odd => conflict-free linear layout)
constexpr int STAGE_PITCH = 33;          // doubles; stage[j * 33 + lane]
constexpr int STAGE_VALS_BYTES = BLK_R * STAGE_PITCH * 8;             // 3960
constexpr int STAGE_BYTES = (STAGE_VALS_BYTES + 256 + 127) / 128 * 128;   // + per-block {k0, nwin} for the transposed read-out
constexpr int WARP_HDR_BYTES = 128;      // mbarrier slot; keeps every per-warp region 128-byte aligned (TMA destination needs 16)

// ------------------------------------------------------------------------------------------------ TMA / mbarrier
// FILO_CUSIM: the kernels compiled for the host on top of tests/cpp/cusim.h (fiber-per-thread SIMT emulation, test infrastructure);
// the PTX helpers below are the only place where the two builds differ.
#ifdef FILO_CUSIM
inline void mbar_init(uint64_t* bar, int count) { cusim::mbar_init(bar, (uint32_t)count); }
inline void mbar_fence_init() {}
inline void mbar_arrive(uint64_t* bar) { cusim::mbar_arrive(bar); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { cusim::mbar_expect_tx(bar, bytes); }
inline void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { cusim::tma_load(dst, src, bytes, bar); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { cusim::mbar_wait(bar, parity); }
inline void mbar_wait_parked(uint64_t* bar, uint32_t parity) { cusim::mbar_wait(bar, parity); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// same, for waits that may last microseconds: the suspend-time hint lets the hardware park the thread instead of spinning
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
  } while (!ok);
}

#endif

// ------------------------------------------------------------------------------------------------ integer helpers
// Division by the (kernel-invariant) query step: double reciprocal + one correction, exact for |a| < 2^31, d < 2^31.
struct StepDiv {
  int64_t d; double inv;
  __device__ __forceinline__ void init(int64_t dd) { d = dd; inv = 1.0 / (double)dd; }
  __device__ __forceinline__ uint32_t udiv(uint32_t n) const {
    uint32_t q = __double2uint_rz(__dmul_rn((double)n, inv));
    const int32_t r = (int32_t)(n - q * (uint32_t)d);
    if (r < 0) --q; else if ((uint32_t)r >= (uint32_t)d) ++q;
    return q;
  }
  __device__ __forceinline__ int64_t floor_div(int64_t a) const {
    if (d < 0x7fffffffLL) {
      if (a >= 0 && a < 0x7fffffffLL) return (int64_t)udiv((uint32_t)a);
      if (a < 0 && a > -0x3fffffffLL) return -(int64_t)udiv((uint32_t)(-a) + (uint32_t)d - 1);
    }
    int64_t q = a / d; if ((a % d) != 0 && a < 0) --q; return q;
  }
  __device__ __forceinline__ int64_t ceil_div(int64_t a) const { return -floor_div(-a); }
};

// x / y for a loop-invariant y, with rcp = RN(1/y) precomputed: q0 = RN(x*rcp); r = x - q0*y (exact, FMA); q = RN(q0 + r*rcp).
// This is the final correction step of the IEEE division sequence (Markstein) and returns the correctly rounded quotient
// whenever q0 is a normal number well inside the exponent range; anything else takes the real division.  y > 0 is assumed.
// IEEE division kept out of line: at the call sites below it is the rare branch, and an inlined division (about 40 instructions) gets
// if-converted into the common path, where it is issued for every window with its results predicated off
#ifdef FILO_CUSIM
inline double ddiv_rare(double x, double y) { return x / y; }
#else
static __device__ __noinline__ double ddiv_rare(double x, double y) { return x / y; }
#endif
__device__ __forceinline__ double div_invariant(double x, double y, double rcp) {
  const double q0 = __dmul_rn(x, rcp);
  const uint32_t ex = ((uint32_t)__double2hiint(q0) >> 20) & 0x7ff;     // biased exponent
  if (ex > 64u && ex < 1983u) { const double r = __fma_rn(-q0, y, x); return __fma_rn(r, rcp, q0); }
  if (x == 0.0) return q0;                 // +-0 / y (y > 0, finite): the product already has the quotient's sign
  return ddiv_rare(x, y);
}

// Single-chunk window interval [kA, kB] of chunk c (const-DDV timestamps with 0 < slope == step).  Window k belongs to it
// iff, all affine in k (SURVEY.md Appendix B / DESIGN.md §3.3):
//   wEnd_k   >= ts[0]                        the (unclamped) row range reaches into the chunk from below
//   wStart_k <= ts[nrows-1]                  ... and from above
//   wStart_k >  max(endTime_{c-1}, last ts of c-1)   previous chunk neither in the window's chunk set nor contributing rows
//   wEnd_k   <  first ts of chunk c+1        next chunk contributes no row (it may be in the chunk set; it is a no-op there)
//   wStart_k <= endTime_c                    chunk c still in the chunk set (ChunkSetInfo.scala:481-483)
// sA = unclamped first row of window kA (may be negative), Wr = last - first row of every window.
// Lanes 0..4 (then 0..1) each do one division; results are broadcast.
__device__ __forceinline__ void chunk_interval(ChunkDesc* D, int n, int c, const QueryParams& q, const StepDiv& sd, int lane) {
  ChunkDesc& d = D[c];
  int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
  const int64_t S0 = q.start - winDur, E0 = q.start;
  const int64_t slope = d.ts_slope, init = d.ts_init;
  int nrows = d.num_rows < d.ts_len ? d.num_rows : d.ts_len; if (d.val_len < nrows) nrows = d.val_len;
  int64_t v = 0;
  if (lane == 0) v = sd.ceil_div(init - E0);                                           // wEnd_k >= ts[0]
  else if (lane == 1) {                                                                 // wStart_k > prevMax
    v = 0;
    if (c > 0) {
      const ChunkDesc& p = D[c - 1];
      int64_t pm = p.end_time; const int64_t plast = ts_apply(p, p.ts_len - 1); if (plast > pm) pm = plast;
      v = sd.ceil_div(pm + 1 - S0);
    }
  }
  else if (lane == 2) v = sd.floor_div(init + (int64_t)(nrows - 1) * slope - S0);       // wStart_k <= ts[nrows-1]
  else if (lane == 3) v = (c + 1 < n) ? sd.floor_div(D[c + 1].ts_init - 1 - E0) : (int64_t)q.T;   // wEnd_k < next first ts
  else if (lane == 4) v = sd.floor_div(d.end_time - S0);                                // wStart_k <= endTime_c
  int64_t kA = __shfl_sync(0xffffffffu, v, 0);
  { const int64_t t = __shfl_sync(0xffffffffu, v, 1); if (t > kA) kA = t; }
  int64_t kB = __shfl_sync(0xffffffffu, v, 2);
  { const int64_t t = __shfl_sync(0xffffffffu, v, 3); if (t < kB) kB = t; }
  { const int64_t t = __shfl_sync(0xffffffffu, v, 4); if (t < kB) kB = t; }
  if (kA < 0) kA = 0;
  if (kB > q.T - 1) kB = q.T - 1;
  int64_t w = 0;
  if (kA <= kB) {
    if (lane == 0) w = sd.ceil_div(S0 + kA * q.step - init);          // unclamped first row of window kA (slope == step)
    else if (lane == 1) w = sd.floor_div(E0 + kA * q.step - init);    // unclamped last row of window kA
  }
  const int64_t sA = __shfl_sync(0xffffffffu, w, 0), eA = __shfl_sync(0xffffffffu, w, 1);
  if (lane == 0) {
    if (kA <= kB && eA >= sA) { d.kA = (int32_t)kA; d.kB = (int32_t)kB; d.sA = (int32_t)sA; d.Wr = (int32_t)(eA - sA); }
    else { d.kA = 0; d.kB = -1; d.sA = 0; d.Wr = 0; }
    d.nrows_eff = nrows;
  }
}

// ------------------------------------------------------------------------------------------------ blocked reductions
// Lane's block: windows j = 0..R-1, window j takes rows r0 + i for i in [j, j + Wr], in row order.  Rows outside [0, nrows)
// are fed as +0.0 (exact no-op).  Requires Wr >= BLK_R - 1.  With CHECK_NAN, NaN rows are skipped and counted out.
// CHECK_BOUNDS = false: the caller guarantees that every row a valid window reads outside [0, nrows) holds +0.0 and that
// rows up to BLK_R - 1 past the last valid window are readable.  NEED_CNT = false: cnt[] is not produced.
template <int R, bool CHECK_NAN, bool CHECK_BOUNDS, bool NEED_CNT>
__device__ __forceinline__ void blocked_sum_r(const double* __restrict__ slots, int r0, int nrows, int Wr, double acc[R], int cnt[R]) {
  static_assert(!CHECK_NAN || (CHECK_BOUNDS && NEED_CNT), "NaN-aware sums count rows and must not see padding");
  auto load = [&](int i, int& ok) -> double {
    const int r = r0 + i;
    if (!CHECK_BOUNDS) { ok = 1; return slots[r]; }
    ok = ((unsigned)r < (unsigned)nrows) ? 1 : 0;
    double v = 0.0;
    if (ok) v = slots[r];
    if (CHECK_NAN) { if (v != v) { v = 0.0; ok = 0; } }     // DoubleVector.scala:243-253
    return v;
  };
#pragma unroll
  for (int j = 0; j < R; ++j) { acc[j] = 0.0; cnt[j] = 0; }
#pragma unroll
  for (int i = 0; i < R - 1; ++i) {                    // ramp-up: row i feeds windows 0..i
    int ok; const double v = load(i, ok);
#pragma unroll
    for (int j = 0; j <= i; ++j) { acc[j] += v; if (CHECK_NAN) cnt[j] += ok; }
  }
  for (int i = R - 1; i <= Wr; ++i) {                  // steady state: every window
    int ok; const double v = load(i, ok);
#pragma unroll
    for (int j = 0; j < R; ++j) { acc[j] += v; if (CHECK_NAN) cnt[j] += ok; }
  }
#pragma unroll
  for (int t = 1; t < R; ++t) {                        // ramp-down: row Wr + t feeds windows t..R-1
    int ok; const double v = load(Wr + t, ok);
#pragma unroll
    for (int j = t; j < R; ++j) { acc[j] += v; if (CHECK_NAN) cnt[j] += ok; }
  }
  if (!CHECK_NAN) {
#pragma unroll
    for (int j = 0; j < R; ++j) {                      // rows of window j inside the chunk
      if (NEED_CNT) {
        int lo = r0 + j; if (lo < 0) lo = 0;
        int hi = r0 + j + Wr; if (hi > nrows - 1) hi = nrows - 1;
        cnt[j] = hi >= lo ? hi - lo + 1 : 0;
      } else cnt[j] = 1;                                   // a blocked window always has a row inside its chunk
    }
  }
}

template <bool CHECK_NAN, bool CHECK_BOUNDS = true, bool NEED_CNT = true>
__device__ __forceinline__ void blocked_sum(const double* __restrict__ slots, int r0, int nrows, int Wr, double acc[BLK_R], int cnt[BLK_R]) {
  blocked_sum_r<BLK_R, CHECK_NAN, CHECK_BOUNDS, NEED_CNT>(slots, r0, nrows, Wr, acc, cnt);
}

template <bool IS_MIN>
__device__ __forceinline__ void blocked_minmax(const ChunkDesc& c, int r0, int nrows, int Wr, double acc[BLK_R]) {
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int j = 0; j < BLK_R; ++j) acc[j] = NaNv;
  for (int i = 0; i <= Wr + BLK_R - 1; ++i) {
    const int r = r0 + i;
    const double v = ((unsigned)r < (unsigned)nrows) ? slot_value(c, r) : NaNv;     // NaN is ignored by min/maxIgnoreNaN
#pragma unroll
    for (int j = 0; j < BLK_R; ++j)
      if (i >= j && i <= j + Wr) acc[j] = IS_MIN ? min_ignore_nan(acc[j], v) : max_ignore_nan(acc[j], v);
  }
}

// finalize one single-chunk window of a SUM-class function from (chunk sum over non-NaN rows, non-NaN count)
struct SumFinish {
  int fn; double div, rcp;
  __device__ __forceinline__ void init(const QueryParams& q) {
    fn = q.fn;
    int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
    // (windowEnd - curWindowStart) is the same for every window: RateFunctions.scala:436-442
    div = (double)(q.inclusive ? winDur : winDur + 1); rcp = 1.0 / div;
  }
  __device__ __forceinline__ double operator()(double cs, int nn) const {
    const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
    // exactly one contributing chunk: sum = isNaN(cs) ? NaN : 0 + cs, count = nn   (AggrOverTimeFunctions.scala:568-570)
    const double sum = nn ? cs : NaNv;
    if (fn == FN_RATE) return __dmul_rn(div_invariant(sum, div, rcp), 1000.0);
    if (fn == FN_SUM || fn == FN_INCREASE) return sum;
    if (fn == FN_AVG) return nn > 0 ? sum / (double)nn : sum;            // AggrOverTimeFunctions.scala:1000
    return (double)nn;                                                   // FN_COUNT
  }
};

// ------------------------------------------------------------------------------------------------ per-series driver
// Sink: void operator()(int k, double v, bool valid) — called by all 32 lanes.
// CLS is the compile-time function class (one kernel instantiation per class keeps the code in the instruction cache).
template <int CLS, class Sink, class RecRelease>
__device__ __forceinline__ void process_series(const uint8_t* rec, const QueryParams& q, uint8_t* scratch, uint32_t scratch_bytes,
                                               double* stage, int lane, int& err, int64_t& rows_scanned, int64_t& bytes_scanned,
                                               Sink&& sink, RecRelease&& release) {
  const RecordHeader* h = reinterpret_cast<const RecordHeader*>(rec);
  const ChunkEntry* E = reinterpret_cast<const ChunkEntry*>(rec + sizeof(RecordHeader));
  const int nch = (int)h->n_chunks;
  const int64_t t1 = q.start - q.window, t2 = q.end;
  int cLo = 0; while (cLo < nch && E[cLo].end_time < t1) ++cLo;
  int cHi = cLo; while (cHi < nch && E[cHi].start_time <= t2) ++cHi;
  if (t1 > t2) cHi = cLo;
  const int n = cHi - cLo;
  err = 0;
  const bool need_corrected = (CLS == CLASS_COUNTER) && q.fn != FN_DELTA;
  ChunkDesc* D = reinterpret_cast<ChunkDesc*>(scratch);
  bool regular = false;
  if (n > 0) {
    ScratchCursor sc; sc.p = scratch + align_up((uint32_t)n * (uint32_t)sizeof(ChunkDesc), 16);
    const uint32_t need = (uint32_t)(sc.p - scratch) + (uint32_t)h->n_rows * 8u *
                          (((h->flags & REC_ALL_TS_CONST) ? 1u : 2u) + ((need_corrected && (h->flags & REC_ANY_DROP)) ? 1u : 0u));
    if (need > scratch_bytes) { err = FILO_DEV_ERR_SCRATCH; return; }
    // regular <=> every chunk in range has const-DDV timestamps with slope == step (rows advance one per window)
    regular = (h->flags & REC_ALL_TS_CONST) != 0 && CLS != CLASS_POINT;
    if (regular) {
      for (int c = 0; c < n; ++c) {
        const uint8_t* tv = rec + E[cLo + c].ts_off;
        if ((int64_t)(int32_t)ld32(tv + 20) != q.step) { regular = false; break; }
      }
    }
    for (int c = 0; c < n; ++c) {
      const int e = resolve_chunk(rec, &E[cLo + c], &D[c], sc, need_corrected, lane, true, q.long_values != 0);
      if (e) { err = e; return; }
    }
    if (lane == 0) {                                        // CountingChunkInfoIterator, ChunkSetInfo.scala:336-380
      const int64_t lastEnd = q.start + (int64_t)(q.T - 1) * q.step;
      int f = 0; while (f < n - 1 && D[f].end_time < lastEnd) ++f;
      for (int c = 0; c <= f; ++c) {
        rows_scanned += D[c].num_rows;
        bytes_scanned += (int64_t)ld32(rec + E[cLo + c].ts_off) + 4 + (int64_t)ld32(rec + E[cLo + c].val_off) + 4;
      }
    }
    if (regular) {
      StepDiv sd; sd.init(q.step);
      for (int c = 0; c < n; ++c) chunk_interval(D, n, c, q, sd, lane);
    }
  }
  __syncwarp();
  release();          // every byte of the record that is still needed now lives in scratch: the staging buffer may be refilled

  if (!regular) {
    for (int k0 = 0; k0 < q.T; k0 += 32) {
      const int k = k0 + lane;
      double v = 0.0;
      if (k < q.T) v = eval_window(D, 0, n, q, k);
      sink(k, v, k < q.T);
    }
    __syncwarp();
    return;
  }

  // ---- phase 1: single-chunk windows
  SumFinish fin; fin.init(q);
  if (CLS == CLASS_SUM) {
    // (a) all chunks whose windows can be blocked (double slots, Wr >= R-1, same Wr) share ONE block list, so that the lanes of
    //     a warp stay busy even when a chunk has few windows
    int uniWr = -1, total_blocks = 0; bool any_nan = false;
    for (int c = 0; c < n; ++c) {
      const ChunkDesc& d = D[c];
      const bool elig = d.kA <= d.kB && !d.val_is_long && d.Wr >= BLK_R - 1 && (uniWr < 0 || d.Wr == uniWr);
      if (elig) { uniWr = d.Wr; any_nan |= d.has_nan != 0; }
      if (lane == 0) { D[c].blk0 = total_blocks; D[c].blk_n = elig ? (d.kB - d.kA + BLK_R) / BLK_R : 0; }
      if (elig) total_blocks += (d.kB - d.kA + BLK_R) / BLK_R;
    }
    __syncwarp();
    int* stage_k0 = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(stage) + STAGE_VALS_BYTES);
    int* stage_nw = stage_k0 + 32;
    for (int B0 = 0; B0 < total_blocks; B0 += 32) {
      const int B = B0 + lane;
      double acc[BLK_R]; int cnt[BLK_R];
      int k0 = 0, nw = 0;
      if (B < total_blocks) {
        int c = 0;
        if (n <= 4) { while (c + 1 < n && B >= D[c].blk0 + D[c].blk_n) ++c; }
        else { int lo = 0, hi = n - 1; while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (D[m].blk0 <= B) lo = m; else hi = m - 1; } c = lo;
               while (c > 0 && D[c].blk_n == 0) --c; }
        const ChunkDesc& d = D[c];
        const int b = B - d.blk0;
        const int r0 = d.sA + b * BLK_R;
        const double* slots = reinterpret_cast<const double*>(d.val_slots);
        if (any_nan) blocked_sum<true>(slots, r0, d.nrows_eff, uniWr, acc, cnt); else blocked_sum<false>(slots, r0, d.nrows_eff, uniWr, acc, cnt);
#pragma unroll
        for (int j = 0; j < BLK_R; ++j) acc[j] = fin(acc[j], cnt[j]);
        k0 = d.kA + b * BLK_R; nw = d.kB - k0 + 1; if (nw > BLK_R) nw = BLK_R;
      }
      // transpose through shared memory so that results leave lane-consecutive
#pragma unroll
      for (int j = 0; j < BLK_R; ++j) stage[j * STAGE_PITCH + lane] = acc[j];
      stage_k0[lane] = k0; stage_nw[lane] = nw;
      __syncwarp();
#pragma unroll
      for (int m = 0; m < BLK_R; ++m) {
        const int off = m * 32 + lane;
        const int bl = off / BLK_R, j = off - bl * BLK_R;
        sink(stage_k0[bl] + j, stage[j * STAGE_PITCH + bl], j < stage_nw[bl]);
      }
      __syncwarp();
    }
    // (b) the rest (DDV-as-long values: closed form + exact Long residual sum, DeltaDeltaVector.scala:190-194; short windows;
    //     a chunk whose Wr differs): one window per lane, the clamped row range is known without any search
    for (int c = 0; c < n; ++c) {
      const ChunkDesc& d = D[c];
      if (d.kA > d.kB || d.blk_n != 0) continue;
      const int nwin = d.kB - d.kA + 1;
      for (int w0 = 0; w0 < nwin; w0 += 32) {
        const int w = w0 + lane;
        double res = 0.0;
        if (w < nwin) {
          int s = d.sA + w, e = s + d.Wr; if (s < 0) s = 0; if (e > d.nrows_eff - 1) e = d.nrows_eff - 1;
          int cnt = 0; double cs = 0.0;
          if (s <= e) cs = chunk_sum(d, s, e, cnt);
          res = fin(cs, cnt);
        }
        sink(d.kA + w, res, w < nwin);
      }
    }
  } else {
    for (int c = 0; c < n; ++c) {
      const ChunkDesc& d = D[c];
      const int kA = d.kA, kB = d.kB;
      if (kA > kB) continue;
      const int Wr = d.Wr;
      const int nwin = kB - kA + 1;
      if (CLS == CLASS_COUNTER) {
        // single chunk, no correction carried in (RangeFunction.scala:138-163 with correctionMeta == NoCorrection)
        int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
        for (int w0 = 0; w0 < nwin; w0 += 32) {
          const int w = w0 + lane;
          double res = __longlong_as_double(0x7ff8000000000000LL);
          if (w < nwin) {
            const int k = kA + w;
            int s = d.sA + w, e = s + Wr; if (s < 0) s = 0; if (e > d.nrows_eff - 1) e = d.nrows_eff - 1;
            const bool skip = (q.fn != FN_DELTA) && s == 0 && e == 0 && is_nan(slot_value(d, 0));
            if (!skip && e > s) {
              const int64_t tS = d.ts_init + (int64_t)d.ts_slope * s, tE = d.ts_init + (int64_t)d.ts_slope * e;
              double loV, hiV;
              if (q.fn == FN_DELTA || !d.dropped) { loV = slot_value(d, s); hiV = slot_value(d, e); }
              else { loV = d.corr_slots[s]; hiV = d.corr_slots[e]; }
              const int64_t wEnd = q.start + (int64_t)k * q.step, wStart = wEnd - winDur;
              const int64_t cws = q.inclusive ? wStart : wStart - 1;
              res = extrapolated_rate(cws, wEnd, e - s + 1, tS, loV, tE, hiV, q.fn != FN_DELTA, q.fn == FN_RATE);
            }
          }
          sink(kA + w, res, w < nwin);
        }
        continue;
      }
      // CLASS_MINMAX: blocked, lane handles BLK_R consecutive windows
      const int nblk = (nwin + BLK_R - 1) / BLK_R;
      for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int b = b0 + lane;
        double acc[BLK_R];
        if (b < nblk) {
          const int r0 = d.sA + b * BLK_R;
          if (q.fn == FN_MIN) blocked_minmax<true>(d, r0, d.nrows_eff, Wr, acc); else blocked_minmax<false>(d, r0, d.nrows_eff, Wr, acc);
        }
#pragma unroll
        for (int j = 0; j < BLK_R; ++j) stage[j * STAGE_PITCH + lane] = acc[j];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < BLK_R; ++m) {
          const int off = m * 32 + lane;
          const int bl = off / BLK_R, j = off - bl * BLK_R;
          const int w = b0 * BLK_R + off;
          sink(kA + w, stage[j * STAGE_PITCH + bl], w < nwin);
        }
        __syncwarp();
      }
    }
  }
  // ---- phase 2: the windows outside every single-chunk interval, enumerated densely, through the literal state machine
  int total = 0;
  { int prev = -1; for (int c = 0; c < n; ++c) { if (D[c].kA <= D[c].kB) { total += D[c].kA - prev - 1; prev = D[c].kB; } } total += q.T - prev - 1; }
  for (int u0 = 0; u0 < total; u0 += 32) {
    int u = u0 + lane;
    const bool todo = u < total;
    int k = 0;
    if (todo) {                                             // u-th uncovered window
      int prev = -1; bool found = false;
      for (int c = 0; c < n && !found; ++c) {
        if (D[c].kA > D[c].kB) continue;
        const int gap = D[c].kA - prev - 1;
        if (u < gap) { k = prev + 1 + u; found = true; } else { u -= gap; prev = D[c].kB; }
      }
      if (!found) k = prev + 1 + u;
    }
    double v = 0.0;
    if (todo) v = eval_window(D, 0, n, q, k);
    sink(k, v, todo);
  }
  __syncwarp();
}

} // namespace filo
