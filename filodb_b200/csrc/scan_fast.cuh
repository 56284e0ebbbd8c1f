// v2 scan path: TMA-staged records, register-blocked window reductions, lane-coalesced result sink.
//
// Per series (one warp):
//   1. the record (chunk pages) arrives in shared memory by one cp.async.bulk (TMA 1-D bulk copy) that was issued while the
//      previous series was being reduced (one staging buffer + one mbarrier per warp);
//   2. chunks are resolved/decoded into per-warp scratch (scan_device.cuh); afterwards the staging buffer is dead and the
//      bulk copy of the NEXT series is issued;
//   3. "interior" windows — windows whose only contributing rows are an unclamped row range [s, s+Wr] of ONE chunk with
//      const-DDV timestamps whose slope equals the query step (the vast majority in practice) — are reduced BLK_R at a time
//      per lane: the lane walks rows s0 .. s0+Wr+R-1 once and feeds each row to the accumulators of the windows containing
//      it.  Every accumulator still sees its rows in row order starting from 0.0, i.e. exactly the reference's sequential sum
//      (DoubleVector.scala:243-253), but a row is loaded once per R windows.  R = 15 (odd): lane base rows are 15 apart, so the
//      64-bit shared-memory reads of a warp are bank-conflict-free in the plain linear layout;
//   4. the remaining windows (series start/end, chunk boundaries, irregular timestamps) are enumerated densely and take the
//      general path (eval_window), which restates the reference state machine literally.
// Results leave through a Sink called by all lanes (lane-consecutive windows on the bulk path: coalesced stores).
#pragma once
#include "scan_device.cuh"

namespace filo {

constexpr int BLK_R = 15;                // windows per lane in the blocked reduction (odd => conflict-free linear layout)
constexpr int STAGE_PITCH = 33;          // doubles; stage[j * 33 + lane]
constexpr int STAGE_BYTES = (BLK_R * STAGE_PITCH * 8 + 127) / 128 * 128;
constexpr int WARP_HDR_BYTES = 128;      // mbarrier slot; keeps every per-warp region 128-byte aligned (TMA destination needs 16)

// ------------------------------------------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

// ------------------------------------------------------------------------------------------------ integer helpers
__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {        // b > 0
  if (a >= 0 && a <= 0xffffffffLL && b <= 0xffffffffLL) return (int64_t)((uint32_t)a / (uint32_t)b);
  if (a < 0 && -a <= 0x7fffffffLL && b <= 0x7fffffffLL) { const uint32_t na = (uint32_t)(-a), ub = (uint32_t)b; return -(int64_t)((na + ub - 1) / ub); }
  int64_t q = a / b; if ((a % b) != 0 && a < 0) --q; return q;
}
__device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) {         // b > 0
  if (a > 0 && a <= 0x7fffffffLL && b <= 0x7fffffffLL) return (int64_t)(((uint32_t)a + (uint32_t)b - 1) / (uint32_t)b);
  return -floor_div(-a, b);
}

// x / y for a loop-invariant y, with rcp = RN(1/y) precomputed: q0 = RN(x*rcp); r = x - q0*y (exact, FMA); q = RN(q0 + r*rcp).
// This is the final correction step of the IEEE division sequence (Markstein) and returns the correctly rounded quotient
// whenever q0 is a normal number well inside the exponent range; anything else takes the real division.
__device__ __forceinline__ double div_invariant(double x, double y, double rcp) {
  const double q0 = __dmul_rn(x, rcp);
  const double aq = fabs(q0);
  if (aq > 1e-290 && aq < 1e290) { const double r = __fma_rn(-q0, y, x); return __fma_rn(r, rcp, q0); }
  return x / y;
}

enum { CLASS_SUM = 0, CLASS_MINMAX = 1, CLASS_POINT = 2, CLASS_COUNTER = 3 };
__host__ __device__ __forceinline__ int fn_class_of(int fn, int cumulative) {
  switch (fn) {
    case FN_SUM: case FN_AVG: case FN_COUNT: return CLASS_SUM;
    case FN_RATE: case FN_INCREASE: return cumulative ? CLASS_COUNTER : CLASS_SUM;
    case FN_DELTA: return CLASS_COUNTER;
    case FN_MIN: case FN_MAX: return CLASS_MINMAX;
    default: return CLASS_POINT;
  }
}

// Interior window interval of chunk c.  Only for const-DDV timestamps with 0 < slope == step.  D[0..n) resolved chunks.
// Window k is interior to chunk c iff (all affine in k, SURVEY.md Appendix B / DESIGN.md §3.3):
//   wStart_k >= ts[0]                      rows unclamped from below
//   wStart_k >  max(endTime_{c-1}, last ts of c-1)   previous chunk neither in the window's chunk set nor contributing rows
//   wEnd_k   <  ts[0] + nrows*slope        rows unclamped from above
//   wEnd_k   <  first ts of chunk c+1      next chunk contributes no row (it may be in the chunk set; it is a no-op there)
//   wStart_k <= endTime_c                  chunk c still in the chunk set (ChunkSetInfo.scala:481-483)
// Lanes 0..2 (then 0..1) each do one division; results are broadcast.
__device__ __forceinline__ void interior_interval(ChunkDesc* D, int n, int c, const QueryParams& q, int lane) {
  ChunkDesc& d = D[c];
  int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
  const int64_t S0 = q.start - winDur, E0 = q.start, step = q.step;
  const int64_t slope = d.ts_slope, init = d.ts_init;
  int nrows = d.num_rows < d.ts_len ? d.num_rows : d.ts_len; if (d.val_len < nrows) nrows = d.val_len;
  int64_t L = init;
  if (c > 0) {
    const ChunkDesc& p = D[c - 1];
    int64_t pm = p.end_time; const int64_t plast = ts_apply(p, p.ts_len - 1); if (plast > pm) pm = plast;
    if (pm + 1 > L) L = pm + 1;
  }
  int64_t U = init + (int64_t)nrows * slope - 1;
  if (c + 1 < n) { const int64_t nfirst = D[c + 1].ts_init; if (nfirst - 1 < U) U = nfirst - 1; }
  int64_t v = 0;
  if (lane == 0) v = ceil_div(L - S0, step);
  else if (lane == 1) v = floor_div(U - E0, step);
  else if (lane == 2) v = floor_div(d.end_time - S0, step);
  int64_t kA = __shfl_sync(0xffffffffu, v, 0), kB = __shfl_sync(0xffffffffu, v, 1);
  const int64_t kB2 = __shfl_sync(0xffffffffu, v, 2);
  if (kB2 < kB) kB = kB2;
  if (kA < 0) kA = 0;
  if (kB > q.T - 1) kB = q.T - 1;
  int64_t w = 0;
  if (kA <= kB) {
    if (lane == 0) w = ceil_div(S0 + kA * step - init, slope);          // first row of window kA
    else if (lane == 1) w = floor_div(E0 + kA * step - init, slope);    // last row of window kA
  }
  const int64_t sA = __shfl_sync(0xffffffffu, w, 0), eA = __shfl_sync(0xffffffffu, w, 1);
  if (lane == 0) {
    if (kA <= kB && eA >= sA) { d.kA = (int32_t)kA; d.kB = (int32_t)kB; d.sA = (int32_t)sA; d.Wr = (int32_t)(eA - sA); }
    else { d.kA = 0; d.kB = -1; d.sA = 0; d.Wr = 0; }
  }
}

// ------------------------------------------------------------------------------------------------ blocked reductions
// `base` points at the lane's first row; rows beyond `nvalid-1` (tail block only) are clamped and feed discarded windows.
// Accumulator j (window k0 + j) takes rows i in [j, j + Wr], in row order.  Requires Wr >= BLK_R - 1.
template <bool CHECK_NAN>
__device__ __forceinline__ void blocked_sum(const double* __restrict__ base, int nvalid, int Wr, double acc[BLK_R], int cnt[BLK_R]) {
  auto load = [&](int i, int& ok) -> double {
    double v = base[i < nvalid ? i : nvalid - 1];
    if (CHECK_NAN) { ok = (v == v) ? 1 : 0; if (!ok) v = 0.0; } else ok = 1;   // NaN rows are skipped (DoubleVector.scala:243-253)
    return v;
  };
#pragma unroll
  for (int j = 0; j < BLK_R; ++j) { acc[j] = 0.0; cnt[j] = 0; }
#pragma unroll
  for (int i = 0; i < BLK_R - 1; ++i) {                    // ramp-up: row i feeds windows 0..i
    int ok; const double v = load(i, ok);
#pragma unroll
    for (int j = 0; j <= i; ++j) { acc[j] += v; if (CHECK_NAN) cnt[j] += ok; }
  }
  for (int i = BLK_R - 1; i <= Wr; ++i) {                  // steady state: every window
    int ok; const double v = load(i, ok);
#pragma unroll
    for (int j = 0; j < BLK_R; ++j) { acc[j] += v; if (CHECK_NAN) cnt[j] += ok; }
  }
#pragma unroll
  for (int t = 1; t < BLK_R; ++t) {                        // ramp-down: row Wr + t feeds windows t..R-1
    int ok; const double v = load(Wr + t, ok);
#pragma unroll
    for (int j = t; j < BLK_R; ++j) { acc[j] += v; if (CHECK_NAN) cnt[j] += ok; }
  }
  if (!CHECK_NAN) {
#pragma unroll
    for (int j = 0; j < BLK_R; ++j) cnt[j] = Wr + 1;
  }
}

template <bool IS_MIN>
__device__ __forceinline__ void blocked_minmax(const ChunkDesc& c, int r0, int Wr, double acc[BLK_R]) {
  const int rmax = c.val_len - 1;
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int j = 0; j < BLK_R; ++j) acc[j] = NaNv;
  for (int i = 0; i <= Wr + BLK_R - 1; ++i) {
    int r = r0 + i; if (r > rmax) r = rmax;
    const double v = slot_value(c, r);
#pragma unroll
    for (int j = 0; j < BLK_R; ++j)
      if (i >= j && i <= j + Wr) acc[j] = IS_MIN ? min_ignore_nan(acc[j], v) : max_ignore_nan(acc[j], v);
  }
}

// finalize one interior window of a SUM-class function from (chunk sum over non-NaN rows, non-NaN count)
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
    switch (fn) {
      case FN_SUM: case FN_INCREASE: return sum;
      case FN_RATE: return __dmul_rn(div_invariant(sum, div, rcp), 1000.0);
      case FN_AVG: return nn > 0 ? sum / (double)nn : sum;            // AggrOverTimeFunctions.scala:1000
      case FN_COUNT: return (double)nn;
    }
    return NaNv;
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
      const int e = resolve_chunk(rec, &E[cLo + c], &D[c], sc, need_corrected, lane, true);
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
    if (regular) for (int c = 0; c < n; ++c) interior_interval(D, n, c, q, lane);
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

  // ---- phase 1: interior windows, chunk by chunk
  SumFinish fin; fin.init(q);
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
          const int k = kA + w, s = d.sA + w, e = s + Wr;
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
    if (CLS == CLASS_SUM && (d.val_is_long || Wr < BLK_R - 1)) {
      // DoubleLongWrapDataReader.sum is a closed form + exact Long residual sum (DeltaDeltaVector.scala:190-194), and short
      // windows do not amortise a block: one window per lane, rows s..s+Wr known without any search
      for (int w0 = 0; w0 < nwin; w0 += 32) {
        const int w = w0 + lane;
        double res = 0.0;
        if (w < nwin) { int cnt; const double cs = chunk_sum(d, d.sA + w, d.sA + w + Wr, cnt); res = fin(cs, cnt); }
        sink(kA + w, res, w < nwin);
      }
      continue;
    }
    // blocked reductions: lane handles BLK_R consecutive windows
    const int nblk = (nwin + BLK_R - 1) / BLK_R;
    for (int b0 = 0; b0 < nblk; b0 += 32) {
      const int b = b0 + lane;
      double acc[BLK_R]; int cnt[BLK_R];
      if (b < nblk) {
        const int r0 = d.sA + b * BLK_R;
        if (CLS == CLASS_SUM) {
          const double* base = reinterpret_cast<const double*>(d.val_slots) + r0;
          if (d.has_nan) blocked_sum<true>(base, d.val_len - r0, Wr, acc, cnt); else blocked_sum<false>(base, d.val_len - r0, Wr, acc, cnt);
#pragma unroll
          for (int j = 0; j < BLK_R; ++j) acc[j] = fin(acc[j], cnt[j]);
        } else if (q.fn == FN_MIN) blocked_minmax<true>(d, r0, Wr, acc);
        else blocked_minmax<false>(d, r0, Wr, acc);
      }
      // transpose through shared memory so that results leave lane-consecutive
#pragma unroll
      for (int j = 0; j < BLK_R; ++j) stage[j * STAGE_PITCH + lane] = acc[j];
      __syncwarp();
#pragma unroll
      for (int m = 0; m < BLK_R; ++m) {
        const int off = m * 32 + lane;                      // window offset within this group of 32*R
        const int bl = off / BLK_R, j = off - bl * BLK_R;
        const int w = b0 * BLK_R + off;
        sink(kA + w, stage[j * STAGE_PITCH + bl], w < nwin);
      }
      __syncwarp();
    }
  }
  // ---- phase 2: the windows outside every interior interval, enumerated densely, through the literal state machine
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
