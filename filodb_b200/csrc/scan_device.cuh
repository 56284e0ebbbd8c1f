// Device-side decode of FiloDB BinaryVectors + per-window chunked range functions (sm_100a).
// One warp owns one time series.  Stage 1 (warp-cooperative): resolve every chunk of the series into a ChunkDesc and
// decode what is not directly addressable (DDV timestamps, DDV-as-long values, XOR-NibblePack doubles, counter-corrected
// rows) into per-warp scratch.  Stage 2: lanes take output windows round-robin and run the reference's
// ChunkedRangeFunction state machine over the chunks of the window, in the reference's operation order.
//
// Reference semantics restated here (file:line under /root/reference):
//   row search      LongBinaryVector.scala:145-169, DeltaDeltaVector.scala:159-188, 245-253
//   chunk selection store/ChunkSetInfo.scala:467-511 (WindowedChunkIterator.nextWindow)
//   range functions rangefn/RangeFunction.scala:131-242,595-724; RateFunctions.scala:72-111,230-322,424-445;
//                   AggrOverTimeFunctions.scala:40-116,553-572,924-1015; QueryUtils.scala:109-123
//   value readers   vectors/DoubleVector.scala:177-207,234-281,325-391,544-567; DeltaDeltaVector.scala:190-194,259-268
//   NibblePack      NibblePack.scala:395-447 (unpack8), 374-384 (unpackDoubleXOR)
// Compiled with -fmad=false: the JVM never contracts a*b+c, and parity is bit-exact per series.
#pragma once
#ifndef FILO_DEV_ERR_TS_WIRE
#define FILO_DEV_ERR_TS_WIRE 1
#define FILO_DEV_ERR_VAL_WIRE 2
#define FILO_DEV_ERR_EMPTY 3
#define FILO_DEV_ERR_SCRATCH 4
#endif
#include <stdint.h>
#include <cuda_runtime.h>
#include "filo_record.h"
#include "scan_params.h"

namespace filo {


struct __align__(16) ChunkDesc {
  int64_t start_time, end_time;
  const int64_t* ts_slots;      // nullptr => closed form ts_init + (int32)(ts_slope * n)   (const DDV)
  const void*    val_slots;     // f64 slots, or i64 slots when val_is_long
  const double*  corr_slots;    // counter-corrected rows of a dropped chunk (CorrectingDoubleVectorReader.corrected)
  int64_t ts_init, val_init;
  int32_t ts_slope, val_slope;
  int32_t num_rows, ts_len, val_len;
  uint8_t val_is_long, dropped, has_nan, fast32;  // has_nan: some double slot is NaN (always 1 when unknown)
  double upd_last, upd_corr;    // dropped chunk: last non-NaN value (or 0) and the chunk's total correction
  double first_val, last_val;   // apply(0), apply(len-1)
  int32_t kA, kB;               // windows [kA, kB] whose only contributing rows are an unclamped row range of this chunk
  int32_t sA;                   // first row of window kA (rows advance by one per window in that interval)
  int32_t Wr;                   // last row - first row of every window in [kA, kB]
  int32_t blk0, blk_n;          // blocked-reduction work list: first block index / number of blocks of this chunk
  int32_t nrows_eff;            // min(num_rows, ts_len, val_len)
  int32_t val_kind;             // VK_F64 / VK_DDV / VK_DDV_CONST / VK_RAW_I64 (reader-specific behaviour: changes(), Long sums)
};
enum { VK_F64 = 0, VK_DDV = 1, VK_DDV_CONST = 2, VK_RAW_I64 = 3 };
static_assert(sizeof(ChunkDesc) == 144, "ChunkDesc size");

// ------------------------------------------------------------------------------------------------ loads
__device__ __forceinline__ uint32_t ld32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint64_t ld64(const uint8_t* p) { return *reinterpret_cast<const uint64_t*>(p); }
__device__ __forceinline__ uint64_t ld64_a4(const uint8_t* p) {           // 4-byte aligned 64-bit field
  return (uint64_t)ld32(p) | ((uint64_t)ld32(p + 4) << 32);
}
__device__ __forceinline__ bool is_nan(double v) { return v != v; }

// IntBinaryVector readers (IntBinaryVector.scala:306-457); `in` = inner vector start (4-byte aligned)
__device__ __forceinline__ int32_t int_apply(const uint8_t* in, int nbits, bool sgn, int n) {
  const uint8_t* d = in + 8;
  switch (nbits) {
    case 32: return (int32_t)ld32(d + 4 * (size_t)n);
    case 16: { uint16_t h = *reinterpret_cast<const uint16_t*>(d + 2 * (size_t)n); return sgn ? (int32_t)(int16_t)h : (int32_t)h; }
    case 8:  { uint8_t b = d[n]; return sgn ? (int32_t)(int8_t)b : (int32_t)b; }
    case 4:  return ((int32_t)(int8_t)d[n >> 1] >> ((n & 1) * 4)) & 0x0f;
    case 2:  return ((int32_t)(int8_t)d[n >> 2] >> ((n & 3) * 2)) & 0x03;
  }
  return 0;
}
__device__ __forceinline__ int int_length(const uint8_t* in) {               // IntBinaryVector.scala:248-250
  uint32_t w = ld32(in + 4);
  int nbits = (w >> 16) & 0x7f, bs = (w >> 24) & 0x3f;
  int nb = (int)ld32(in);
  return ((nb - 4) * 8 + (bs != 0 ? bs - 8 : 0)) / nbits;
}

// ------------------------------------------------------------------------------------------------ XOR decode
// Warp-cooperative decode of a FiloXorDoubleVector into out[0..n).  Lane g handles NibblePack group g (8 values):
// field extraction is independent per group thanks to the group-offset table; the XOR chain is a warp prefix-XOR.
__device__ __forceinline__ bool xor_decode_warp(const uint8_t* v, double* out, int lane) {
  bool any_nan = false;
  const int n = (int)ld32(v + XOR_OFF_N);
  const uint32_t w12 = ld32(v + XOR_OFF_NGROUPS);
  const int ng = w12 & 0xffff, payloadOff = w12 >> 16;
  const uint8_t* payload = v + payloadOff;
  if (n <= 0) return false;
  uint64_t carry = ld64(payload);
  if (lane == 0) { reinterpret_cast<uint64_t*>(out)[0] = carry; any_nan = (carry & 0x7fffffffffffffffull) > 0x7ff0000000000000ull; }
  const uint8_t* groups = payload + 8;
  const uint16_t* tab = reinterpret_cast<const uint16_t*>(v + XOR_OFF_GROUPTAB);
  for (int g0 = 0; g0 < ng; g0 += 32) {
    const int g = g0 + lane;
    uint64_t d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 0;
    if (g < ng) {
      const uint8_t* gp = groups + tab[g];
      const uint32_t mask = gp[0];
      if (mask != 0) {
        const uint32_t hdr = gp[1];
        const int numBits = ((hdr >> 4) + 1) * 4;
        const int tz = (hdr & 0x0f) * 4;
        const uint64_t fmask = numBits >= 64 ? ~0ull : ((1ull << numBits) - 1);
        const uint8_t* data = gp + 2;
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(data);
        uint64_t bit = (uint64_t)(a0 & 7) * 8;               // bit offset from the aligned base
        const uint8_t* base = reinterpret_cast<const uint8_t*>(a0 & ~(uintptr_t)7);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (mask & (1u << i)) {
            const uint8_t* wp = base + ((bit >> 6) << 3);
            const int off = (int)(bit & 63);
            uint64_t w0 = ld64(wp);
            uint64_t val = w0 >> off;
            if (off + numBits > 64) val |= ld64(wp + 8) << (64 - off);
            d[i] = (val & fmask) << tz;
            bit += numBits;
          }
        }
      }
    }
    uint64_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x ^= d[i]; d[i] = x; }
    uint64_t incl = x;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl ^= t;
    }
    const uint64_t basev = carry ^ incl ^ x;
    if (g < ng) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int idx = 1 + g * 8 + i;
        if (idx < n) {
          const uint64_t bits = basev ^ d[i];
          reinterpret_cast<uint64_t*>(out)[idx] = bits;
          any_nan |= ((uint32_t)(bits >> 32) & 0x7ff00000u) == 0x7ff00000u;    // NaN or Inf: conservative, 2 instructions
        }
      }
    }
    carry ^= __shfl_sync(0xffffffffu, incl, 31);
  }
  return any_nan;
}

// ------------------------------------------------------------------------------------------------ chunk resolve
struct ScratchCursor { uint8_t* p; };

__device__ __forceinline__ double slot_value(const ChunkDesc& c, int r) {
  if (c.val_is_long) return (double)reinterpret_cast<const int64_t*>(c.val_slots)[r];   // DoubleLongWrapDataReader.apply
  return reinterpret_cast<const double*>(c.val_slots)[r];
}

// Resolves chunk `e` of the record into `d`; decodes into scratch as needed.  Returns an error code (0 = ok).
// All lanes call it with identical arguments; stores to *d are done by lane 0 and published with __syncwarp.
__device__ __forceinline__ int resolve_chunk(const uint8_t* rec, const ChunkEntry* e, ChunkDesc* d, ScratchCursor& sc,
                                             bool need_corrected, int lane, bool copy_all = false, bool long_col = false) {
  bool lane_nan = false; bool nan_known = false;
  const uint8_t* tv = rec + e->ts_off;
  const uint8_t* vv = rec + e->val_off;
  const uint32_t tw = ld32(tv + 4), vw = ld32(vv + 4);
  const int twire = tw & 0xffff, vwire = vw & 0xffff;
  int err = 0;
  // ---- timestamps (LongBinaryVector.scala:60-67)
  const int64_t* ts_slots = nullptr; int64_t ts_init = 0; int32_t ts_slope = 0; int32_t ts_len = 0;
  if (twire == WIRE_DDV_CONST) {                         // DeltaDeltaVector.scala:89-106
    ts_len = (int32_t)ld32(tv + 8); ts_init = (int64_t)ld64_a4(tv + 12); ts_slope = (int32_t)ld32(tv + 20);
  } else if (twire == WIRE_RAW64) {
    ts_len = ((int32_t)ld32(tv) - 4) / 8; ts_slots = reinterpret_cast<const int64_t*>(tv + 8);
    if (copy_all) {                                      // the staged record buffer is recycled right after resolve
      int64_t* slots = reinterpret_cast<int64_t*>(sc.p);
      for (int r = lane; r < ts_len; r += 32) slots[r] = ts_slots[r];
      ts_slots = slots; sc.p += (size_t)ts_len * 8;
    }
  } else if (twire == WIRE_DDV) {                        // DeltaDeltaVector.scala:138-156
    const uint8_t* in = tv + 20;
    const uint32_t iw = ld32(in + 4);
    const int nbits = (iw >> 16) & 0x7f; const bool sgn = (iw >> 23) & 1;
    ts_len = int_length(in);
    const int64_t init = (int64_t)ld64(tv + 8); const int64_t slope = (int32_t)ld32(tv + 16);
    int64_t* slots = reinterpret_cast<int64_t*>(sc.p);
    for (int r = lane; r < ts_len; r += 32) slots[r] = init + slope * r + (int64_t)int_apply(in, nbits, sgn, r);
    ts_slots = slots; sc.p += (size_t)ts_len * 8;
  } else err = FILO_DEV_ERR_TS_WIRE;
  // ---- values (DoubleVector.scala:62-71)
  const void* val_slots = nullptr; int64_t val_init = 0; int32_t val_slope = 0; int32_t val_len = 0; bool is_long = false;
  const bool dropped = (vw >> 31) & 1;                   // PrimitiveVectorReader.dropped, BinaryVector.scala:530-531
  int val_kind = VK_F64;
  if (vwire == WIRE_XOR && long_col) err = err ? err : FILO_DEV_ERR_VAL_WIRE;      // not a LongBinaryVector
  else if (vwire == WIRE_RAW64 && long_col) {            // LongVectorDataReader64: 64-bit longs (LongBinaryVector.scala:192-206)
    is_long = true; val_kind = VK_RAW_I64; nan_known = true;
    val_len = ((int32_t)ld32(vv) - 4) / 8;
    if (!copy_all) val_slots = vv + 8;
    else {
      uint64_t* slots = reinterpret_cast<uint64_t*>(sc.p);
      const uint64_t* src = reinterpret_cast<const uint64_t*>(vv + 8);
      for (int r = lane; r < val_len; r += 32) slots[r] = src[r];
      val_slots = slots; sc.p += (size_t)val_len * 8;
    }
  } else if (vwire == WIRE_RAW64) {
    val_len = ((int32_t)ld32(vv) - 4) / 8;
    if (!copy_all) val_slots = vv + 8;                   // addressable in place
    else {                                               // the staged record buffer is recycled right after resolve
      uint64_t* slots = reinterpret_cast<uint64_t*>(sc.p);
      const uint64_t* src = reinterpret_cast<const uint64_t*>(vv + 8);
      for (int r = lane; r < val_len; r += 32) { const uint64_t b = src[r]; slots[r] = b; lane_nan |= (b & 0x7fffffffffffffffull) > 0x7ff0000000000000ull; }
      nan_known = true;
      val_slots = slots; sc.p += (size_t)val_len * 8;
    }
  } else if (vwire == WIRE_DDV_CONST) {
    is_long = true; val_kind = VK_DDV_CONST; val_len = (int32_t)ld32(vv + 8); val_init = (int64_t)ld64_a4(vv + 12); val_slope = (int32_t)ld32(vv + 20);
    int64_t* slots = reinterpret_cast<int64_t*>(sc.p);
    for (int r = lane; r < val_len; r += 32) slots[r] = val_init + (int64_t)(int32_t)((uint32_t)val_slope * (uint32_t)r);
    val_slots = slots; sc.p += (size_t)val_len * 8; nan_known = true;
  } else if (vwire == WIRE_DDV) {
    is_long = true; val_kind = VK_DDV;
    const uint8_t* in = vv + 20;
    const uint32_t iw = ld32(in + 4);
    const int nbits = (iw >> 16) & 0x7f; const bool sgn = (iw >> 23) & 1;
    val_len = int_length(in);
    val_init = (int64_t)ld64(vv + 8); val_slope = (int32_t)ld32(vv + 16);
    int64_t* slots = reinterpret_cast<int64_t*>(sc.p);
    for (int r = lane; r < val_len; r += 32) slots[r] = val_init + (int64_t)val_slope * r + (int64_t)int_apply(in, nbits, sgn, r);
    val_slots = slots; sc.p += (size_t)val_len * 8; nan_known = true;
  } else if (vwire == WIRE_XOR) {
    val_len = (int32_t)ld32(vv + XOR_OFF_N);
    double* slots = reinterpret_cast<double*>(sc.p);
    lane_nan = xor_decode_warp(vv, slots, lane); nan_known = true;
    val_slots = slots; sc.p += (size_t)val_len * 8;
  } else err = err ? err : FILO_DEV_ERR_VAL_WIRE;
  const unsigned nan_ballot = __ballot_sync(0xffffffffu, lane_nan);
  __syncwarp();
  if (err) return err;
  if (val_len <= 0 || ts_len <= 0) return FILO_DEV_ERR_EMPTY;
  if (lane == 0) {
    d->start_time = e->start_time; d->end_time = e->end_time;
    d->ts_slots = ts_slots; d->val_slots = val_slots; d->corr_slots = nullptr;
    d->ts_init = ts_init; d->val_init = val_init; d->ts_slope = ts_slope; d->val_slope = val_slope;
    d->num_rows = e->num_rows; d->ts_len = ts_len; d->val_len = val_len;
    d->val_is_long = is_long; d->dropped = dropped; d->has_nan = (uint8_t)((!nan_known || nan_ballot != 0) ? 1 : 0);
    d->kA = 0; d->kB = -1; d->sA = 0; d->Wr = 0; d->blk0 = 0; d->blk_n = 0; d->nrows_eff = 0; d->val_kind = val_kind;
    // row search may use 32-bit arithmetic when no Int wrap can occur in slope * n (DeltaDeltaVector.scala:241-253)
    d->fast32 = (ts_slots == nullptr && ts_slope > 0 && (int64_t)ts_slope * ((int64_t)ts_len + 1) < 0x7fffffffLL) ? 1 : 0;
    d->upd_last = 0; d->upd_corr = 0;
  }
  __syncwarp();
  if (lane == 0) { d->first_val = slot_value(*d, 0); d->last_val = slot_value(*d, val_len - 1); }
  // ---- CorrectingDoubleVectorReader.corrected / updateCorrection (DoubleVector.scala:325-342, 375-391)
  if (dropped && need_corrected) {
    double* cs = reinterpret_cast<double*>(sc.p);
    double acc = 0.0;                                     // _correction
    double lastNonNaN = 0.0; bool haveLast = false;
    for (int r0 = 0; r0 < val_len; r0 += 32) {
      const int r = r0 + lane;
      double v = 0.0, prev = -1.7976931348623157e308;     // Double.MinValue
      bool nan_v = true;
      if (r < val_len) {
        double raw = slot_value(*d, r); nan_v = is_nan(raw); v = nan_v ? 0.0 : raw;
        if (r > 0) { double pr = slot_value(*d, r - 1); prev = is_nan(pr) ? 0.0 : pr; }
      }
      const bool isdrop = (r < val_len) && (v < prev);
      unsigned m = __ballot_sync(0xffffffffu, isdrop);
      double mine = acc;
      while (m) {                                         // serial over drops: keeps the reference's add order
        const int b = __ffs(m) - 1; m &= m - 1;
        acc += __shfl_sync(0xffffffffu, prev, b);
        if (lane >= b) mine = acc;
      }
      if (r < val_len) cs[r] = v + mine;
      const unsigned nn = __ballot_sync(0xffffffffu, (r < val_len) && !nan_v);
      if (nn) { const int hb = 31 - __clz(nn); lastNonNaN = __shfl_sync(0xffffffffu, v, hb); haveLast = true; }
    }
    sc.p += (size_t)val_len * 8;
    if (lane == 0) { d->corr_slots = cs; d->upd_last = haveLast ? lastNonNaN : 0.0; d->upd_corr = acc; }
  }
  __syncwarp();
  return 0;
}

// ------------------------------------------------------------------------------------------------ row search
// first row with ts >= t (== len if none): binarySearch(...) & 0x7fffffff;  *exact = bit 31 clear
__device__ __forceinline__ int ts_search(const ChunkDesc& c, int64_t item, bool& exact) {
  if (c.ts_slots == nullptr) {                            // DeltaDeltaConstDataReader.binarySearch, DeltaDeltaVector.scala:245-253
    const int32_t len = c.ts_len;
    if (c.fast32) {
      // slope > 0 and slope*(len+1) < 2^31: same quotient as the reference's truncating Long division, in 32-bit arithmetic
      const int64_t diff = item - c.ts_init;
      if (diff <= 0) { exact = (diff == 0); return 0; }
      const uint32_t slope = (uint32_t)c.ts_slope;
      if (diff > (int64_t)(len - 1) * (int64_t)slope) { exact = false; return len; }
      const uint32_t g = ((uint32_t)diff + slope - 1) / slope;
      exact = (g * slope == (uint32_t)diff);
      return (int)g;
    }
    const int64_t slope = (int64_t)c.ts_slope;
    int32_t guess;
    if (slope == 0) guess = (item <= c.ts_init) ? 0 : len;
    else guess = (int32_t)((item - c.ts_init + (slope - 1)) / slope);
    if (guess < 0) { exact = false; return 0; }
    if (guess >= len) { exact = false; return len; }
    const int64_t at = c.ts_init + (int64_t)(int32_t)((uint32_t)c.ts_slope * (uint32_t)guess);
    exact = (item == at);
    return guess;
  }
  // decoded / raw timestamps are strictly increasing (TimeSeriesPartition.ingest drops ts <= last, TimeSeriesPartition.scala:135-136)
  int lo = 0, hi = c.ts_len;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (c.ts_slots[mid] < item) lo = mid + 1; else hi = mid; }
  exact = (lo < c.ts_len) && (c.ts_slots[lo] == item);
  return lo;
}
__device__ __forceinline__ int64_t ts_apply(const ChunkDesc& c, int n) {
  if (c.ts_slots == nullptr) return c.ts_init + (int64_t)(int32_t)((uint32_t)c.ts_slope * (uint32_t)n);
  return c.ts_slots[n];
}

// ------------------------------------------------------------------------------------------------ reducers
__device__ __forceinline__ double min_ignore_nan(double a, double b) { if (a != a) return b; if (b != b) return a; return a < b ? a : b; }
__device__ __forceinline__ double max_ignore_nan(double a, double b) { if (a != a) return b; if (b != b) return a; return a > b ? a : b; }

__device__ __forceinline__ double slope_sum(int64_t initVal, int32_t slope, int start, int end) {     // DeltaDeltaVector.scala:265-268
  const int32_t len = end - start + 1;
  const int64_t a = initVal + (int64_t)start * (int64_t)slope;
  const int32_t half = (int32_t)((uint32_t)(end - start) * (uint32_t)len) / 2;
  return (double)len * (double)a + (double)((int64_t)half * (int64_t)slope);
}

// chunk sum over rows [s,e] (DoubleVectorDataReader64.sum / DoubleLongWrapDataReader.sum); *cnt = non-NaN count
__device__ __forceinline__ double chunk_sum(const ChunkDesc& c, int s, int e, int& cnt) {
  if (c.val_kind == VK_RAW_I64) {                         // LongVectorDataReader64.sum: sequential double adds (LongBinaryVector.scala:211-222)
    const int64_t* lv = reinterpret_cast<const int64_t*>(c.val_slots);
    double sum = 0.0;
    for (int r = s; r <= e; ++r) sum += (double)lv[r];
    cnt = e - s + 1;
    return sum;
  }
  if (c.val_is_long) {
    const int64_t* lv = reinterpret_cast<const int64_t*>(c.val_slots);
    int64_t resid = 0;
    for (int r = s; r <= e; ++r) resid += lv[r] - (c.val_init + (int64_t)c.val_slope * r);
    cnt = e - s + 1;
    return slope_sum(c.val_init, c.val_slope, s, e) + (double)resid;
  }
  const double* dv = reinterpret_cast<const double*>(c.val_slots);
  double sum = 0.0; int n = 0;
  // NaN-seeded sum that skips NaN == (0.0 + v1 + v2 ...) over non-NaN values, NaN if there are none
  for (int r = s; r <= e; ++r) { double v = dv[r]; if (v == v) { sum += v; ++n; } }
  cnt = n;
  return n ? sum : __longlong_as_double(0x7ff8000000000000LL);
}

__device__ __forceinline__ double extrapolated_rate(int64_t windowStart, int64_t windowEnd, int32_t numSamples,
                                                    int64_t t1, double v1, int64_t t2, double v2, bool isCounter, bool isRate) {
  // RateFunctions.scala:72-111 — same operation order
  double durationToStart = (double)(t1 - windowStart) / 1000.0;
  const double durationToEnd = (double)(windowEnd - t2) / 1000.0;
  const double sampledInterval = (double)(t2 - t1) / 1000.0;
  const double avgDur = sampledInterval / ((double)numSamples - 1.0);
  const double delta = v2 - v1;
  if (isCounter && delta > 0 && v1 >= 0) {
    const double durationToZero = sampledInterval * (v1 / delta);
    if (durationToZero < durationToStart) durationToStart = durationToZero;
  }
  const double thr = avgDur * 1.1;
  double ext = sampledInterval;
  ext += (durationToStart < thr) ? durationToStart : avgDur / 2.0;
  ext += (durationToEnd < thr) ? durationToEnd : avgDur / 2.0;
  const double scaledDelta = delta * (ext / sampledInterval);
  return isRate ? (scaledDelta / (double)(windowEnd - windowStart) * 1000.0) : scaledDelta;
}

// ------------------------------------------------------------------------------------------------ extended functions
// The remaining chunked range functions (AggrOverTimeFunctions.scala:1082-1604, RangeFunction.scala:725-748) and the Long-column
// variants (AggrOverTimeFunctions.scala:60-116,574-585,924-938,1019-1028,1144-1183,1211-1225,1322-1359; RangeFunction.scala:696-703).
// Out of line: the hot kernels keep their register budget; these functions run window by window (one lane per window).
__device__ __forceinline__ bool chunk_rows(const ChunkDesc& c, int64_t wStart, int64_t wEnd, int& s, int& e) {
  bool ex; s = ts_search(c, wStart, ex);
  const int idx = ts_search(c, wEnd, ex);
  e = ex ? idx : idx - 1; if (e > c.num_rows - 1) e = c.num_rows - 1;
  return s <= e;
}
__device__ __forceinline__ void window_chunk_set(const ChunkDesc* D, int cLo, int cHi, int64_t wStart, int64_t wEnd, int& a, int& f) {
  a = cLo;
  while (a < cHi && D[a].end_time < wStart) ++a;
  f = a; while (f < cHi - 1 && D[f].end_time < wEnd) ++f;
}
__device__ __forceinline__ int64_t d2l_jvm(double d) {       // JVM d2l: NaN -> 0, saturating
  if (d != d) return 0;
  if (d >= 9223372036854775807.0) return INT64_MAX;
  if (d <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)d;
}
__device__ __forceinline__ int64_t slot_long(const ChunkDesc& c, int r) { return reinterpret_cast<const int64_t*>(c.val_slots)[r]; }
// LongVectorDataReader.changes per reader (DeltaDeltaVector.scala:212-227, 280-288; LongBinaryVector.scala:248-265)
__device__ __forceinline__ void long_changes(const ChunkDesc& c, int s, int e, int64_t prev, bool ignorePrev, int64_t& ch, int64_t& last) {
  if (c.val_kind == VK_DDV_CONST) {
    const int64_t firstValue = slot_long(c, s); last = slot_long(c, e);
    ch = (!ignorePrev && prev != firstValue) ? 1 : 0;
    if (c.val_slope != 0) ch += (int64_t)(e - s);
    return;
  }
  int64_t prevVector = prev; ch = 0;
  for (int i = s; i <= e; ++i) {
    const int64_t cur = slot_long(c, i);
    if (i == s && (ignorePrev || c.val_kind == VK_RAW_I64)) prevVector = cur;      // the raw 64-bit reader always re-seeds prev
    if (prevVector != cur) ch += 1;
    prevVector = cur;
  }
  last = prevVector;
}
// order-preserving key of a double (Double.compare order: -0.0 < 0.0, NaN above +Inf) and its inverse
__device__ __forceinline__ uint64_t dkey(double v) { const uint64_t b = (uint64_t)__double_as_longlong(v); return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull); }
__device__ __forceinline__ double dkey_inv(uint64_t k) { const uint64_t b = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k; return __longlong_as_double((long long)b); }

// values of a window: every row of the chunks' row ranges, NaN dropped for double columns.  MODE 0: the value, MODE 1: |ref - value|
template <int MODE>
__device__ __forceinline__ void win_count_le(const ChunkDesc* D, int a, int f, int cHi, int64_t wStart, int64_t wEnd, bool long_col, double ref,
                                             uint64_t K, int& n_le, uint64_t& next_above) {
  n_le = 0; next_above = ~0ull;
  for (int ci = a; ci <= f && ci < cHi; ++ci) {
    const ChunkDesc& c = D[ci]; int s, e;
    if (!chunk_rows(c, wStart, wEnd, s, e)) continue;
    for (int r = s; r <= e; ++r) {
      double v = slot_value(c, r);
      if (!long_col && is_nan(v)) continue;
      if (MODE == 1) v = fabs(ref - v);
      const uint64_t kk = dkey(v);
      if (kk <= K) ++n_le; else if (kk < next_above) next_above = kk;
    }
  }
}
// the values with sorted indices lo and hi = min(n - 1, lo + 1) (java.util.Arrays.sort order) by bisection over the key space
template <int MODE>
__device__ __forceinline__ void win_select(const ChunkDesc* D, int a, int f, int cHi, int64_t wStart, int64_t wEnd, bool long_col, double ref,
                                           int lo, int hi, double& vlo, double& vhi) {
  uint64_t L = 0, H = ~0ull; int n_le; uint64_t above;
  while (L < H) {                                        // smallest K with count(key <= K) >= lo + 1
    const uint64_t M = L + ((H - L) >> 1);
    win_count_le<MODE>(D, a, f, cHi, wStart, wEnd, long_col, ref, M, n_le, above);
    if (n_le >= lo + 1) H = M; else L = M + 1;
  }
  win_count_le<MODE>(D, a, f, cHi, wStart, wEnd, long_col, ref, L, n_le, above);
  vlo = dkey_inv(L);
  vhi = (hi == lo || n_le >= hi + 1) ? vlo : dkey_inv(above);
}

#ifdef FILO_CUSIM
inline double eval_window_ext(
#else
static __device__ __noinline__ double eval_window_ext(
#endif
    const ChunkDesc* D, int cLo, int cHi, const QueryParams& q, int k, int a, int f,
                                               int64_t wStart, int64_t wEnd) {
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
  const int fn = q.fn;
  const bool lng = q.long_values != 0;
  switch (fn) {
    case FN_LAST: case FN_PRESENT: {                      // LastSampleChunkedFunctionL :696-703; PresentOverTimeChunkedFunctionD :725-748
      int64_t lastTs = -1; double lastVal = NaNv;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        const ChunkDesc& c = D[ci];
        bool ex; const int idx = ts_search(c, wEnd, ex);
        int e = ex ? idx : idx - 1; if (e > c.num_rows - 1) e = c.num_rows - 1;
        if (e < 0) continue;
        const int64_t t = ts_apply(c, e);
        if (!(t >= wStart && t > lastTs)) continue;
        if (fn == FN_LAST) { lastTs = t; lastVal = slot_value(c, e); }
        else {
          const double dv = slot_value(c, e);
          if (is_nan(dv)) { if (e > 0) { lastTs = t; lastVal = is_nan(slot_value(c, e - 1)) ? NaNv : 1.0; } }
          else { lastTs = t; lastVal = 1.0; }
        }
      }
      return lastVal;
    }
    case FN_COUNT: {                                      // CountOverTimeChunkedFunction :924-938 (Long columns)
      int32_t count = 0;
      for (int ci = a; ci <= f && ci < cHi; ++ci) { int s, e; if (chunk_rows(D[ci], wStart, wEnd, s, e)) count += e - s + 1; }
      return (double)count;
    }
    case FN_SUM: case FN_AVG: {                           // SumOverTimeChunkedFunctionL :574-585; AvgOverTimeChunkedFunctionL :1019-1028
      double sum = NaNv; int32_t count = 0;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        int s, e; if (!chunk_rows(D[ci], wStart, wEnd, s, e)) continue;
        int cnt; const double cs = chunk_sum(D[ci], s, e, cnt);
        if (fn == FN_SUM && is_nan(sum)) sum = 0.0;       // the Avg variant never clears its NaN seed: avg_over_time of a Long column is NaN
        sum += cs; count += e - s + 1;
      }
      if (fn == FN_SUM) return sum;
      return count > 0 ? sum / (double)count : (is_nan(sum) ? sum : 0.0);
    }
    case FN_MIN: case FN_MAX: {                           // Min/MaxOverTimeChunkedFunctionL :60-116
      int64_t m = fn == FN_MIN ? INT64_MAX : INT64_MIN;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        int s, e; if (!chunk_rows(D[ci], wStart, wEnd, s, e)) continue;
        for (int r = s; r <= e; ++r) { const int64_t v = slot_long(D[ci], r); m = fn == FN_MIN ? (v < m ? v : m) : (v > m ? v : m); }
      }
      return (double)m;
    }
    case FN_STDDEV: case FN_STDVAR: case FN_ZSCORE: {
      if (lng) {                                          // VarOverTimeChunkedFunctionL :1144-1183
        double sum = 0.0, sq = 0.0; int32_t count = 0;
        for (int ci = a; ci <= f && ci < cHi; ++ci) {
          int s, e; if (!chunk_rows(D[ci], wStart, wEnd, s, e)) continue;
          double _sum = 0.0, _sq = 0.0;
          for (int r = s; r <= e; ++r) { const double v = (double)slot_long(D[ci], r); _sum += v; _sq += v * v; }
          count += e - s + 1; sum += _sum; sq += _sq;
        }
        const double avg = count > 0 ? sum / (double)count : 0.0;
        const double var = sq / (double)count - avg * avg;
        return fn == FN_STDDEV ? sqrt(var) : var;
      }
      double sum = NaNv, sq = NaNv; int32_t count = 0;    // VarOverTimeChunkedFunctionD :1082-1142, ZScoreChunkedFunctionD :1592-1604
      double lastSample = NaNv; bool haveLast = false;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        int s, e; if (!chunk_rows(D[ci], wStart, wEnd, s, e)) continue;
        double cs = NaNv, cq = NaNv; int cc = 0;
        for (int r = s; r <= e; ++r) {
          const double v = slot_value(D[ci], r);
          if (!is_nan(v)) {
            if (is_nan(cs)) cs = 0.0;
            if (is_nan(cq)) cq = 0.0;
            if (r == e) { lastSample = v; haveLast = true; }
            cs += v; cq += v * v; cc += 1;
          }
        }
        if (!is_nan(cs) && is_nan(sum)) sum = 0.0;
        sum += cs;
        if (!is_nan(cq) && is_nan(sq)) sq = 0.0;
        sq += cq;
        count += cc;
      }
      if (count <= 0) return is_nan(sum) ? sum : 0.0;
      const double avg = sum / (double)count;
      if (fn == FN_STDVAR) return (sq / (double)count) - (avg * avg);
      if (fn == FN_STDDEV) return sqrt((sq / (double)count) - (avg * avg));
      // zscore: lastSample is not cleared between windows (reset() leaves it): when this window's end rows are all NaN the value of the
      // closest earlier window that saw a number at a chunk's end row is still there
      for (int kk = k - 1; kk >= 0 && !haveLast; --kk) {
        const int64_t we = q.start + (int64_t)kk * q.step, ws = we - (wEnd - wStart);
        int a2, f2; window_chunk_set(D, cLo, cHi, ws, we, a2, f2);
        for (int ci = a2; ci <= f2 && ci < cHi; ++ci) {
          int s, e; if (!chunk_rows(D[ci], ws, we, s, e)) continue;
          const double v = slot_value(D[ci], e);
          if (!is_nan(v)) { lastSample = v; haveLast = true; }
        }
      }
      const double stdDev = sqrt(sq / (double)count - avg * avg);
      return (lastSample - avg) / stdDev;
    }
    case FN_CHANGES: {                                    // ChangesChunkedFunctionD / L :1185-1225
      double changes = NaNv, prev = NaNv;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        const ChunkDesc& c = D[ci];
        int s, e; if (!chunk_rows(c, wStart, wEnd, s, e)) continue;
        if (is_nan(changes)) changes = 0.0;
        if (c.val_is_long) {                              // Long column, or the DoubleLongWrap reader of a double column (DoubleVector.scala:559-566)
          int64_t ch, last;
          long_changes(c, s, e, d2l_jvm(prev), lng ? false : is_nan(prev), ch, last);
          changes += (double)ch; prev = (double)last;
        } else {                                          // DoubleVectorDataReader64.changes, DoubleVector.scala:283-303
          double ch = 0.0, pv = prev;
          for (int r = s; r <= e; ++r) {
            const double v = slot_value(c, r);
            if (!is_nan(v) && pv != v && !is_nan(pv)) ch += 1.0;
            pv = v;
          }
          changes += ch; prev = pv;
        }
      }
      return changes;
    }
    case FN_QUANTILE: case FN_MAD: {                      // Quantile :1227-1344, MedianAbsoluteDeviation :1248-1359
      int n = 0; bool visited = false;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        const ChunkDesc& c = D[ci]; int s, e;
        if (!chunk_rows(c, wStart, wEnd, s, e)) continue;
        visited = true;
        if (lng) n += e - s + 1;
        else for (int r = s; r <= e; ++r) n += is_nan(slot_value(c, r)) ? 0 : 1;
      }
      const double qv = fn == FN_MAD ? 0.5 : q.p0;
      if (fn == FN_QUANTILE && qv < 0) return visited ? __longlong_as_double(0xfff0000000000000LL) : NaNv;
      if (fn == FN_QUANTILE && qv > 1) return visited ? __longlong_as_double(0x7ff0000000000000LL) : NaNv;
      if (n == 0) return NaNv;
      const double rank = qv * (double)(n - 1);           // QuantileOverTimeFunction.calculateRank :399-407
      const double fl = floor(rank);
      const double lower = fl > 0.0 ? fl : 0.0;
      const double upper = (lower + 1 < (double)(n - 1)) ? lower + 1 : (double)(n - 1);
      const double weight = rank - fl;
      const int lo = (int)lower, hi = (int)upper;
      double vlo, vhi;
      win_select<0>(D, a, f, cHi, wStart, wEnd, lng, 0.0, lo, hi, vlo, vhi);
      const double res = vlo * (1 - weight) + vhi * weight;
      if (fn == FN_QUANTILE) return res;
      double dlo, dhi;
      win_select<1>(D, a, f, cHi, wStart, wEnd, lng, res, lo, hi, dlo, dhi);
      return dlo * (1 - weight) + dhi * weight;
    }
    case FN_HOLT_WINTERS: {                               // HoltWintersChunkedFunctionD :1393-1453 (see oracle/filo_query.hpp addHoltWinters)
      const double sf = q.p0, tf = q.p1;
      double b0 = NaNv, s0 = NaNv, nextvalue = NaNv, smoothed = NaNv;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        const ChunkDesc& c = D[ci];
        int s, e; if (!chunk_rows(c, wStart, wEnd, s, e)) continue;
        int pos = s, rowNum = s;
        if (is_nan(s0) && is_nan(b0)) {
          double _s0 = NaNv, _b0 = NaNv; int cur = s;
          while (cur <= e && is_nan(_s0)) { const double nv = slot_value(c, pos++); if (!is_nan(nv)) _s0 = nv; ++cur; }
          while (cur <= e && is_nan(_b0)) { const double nv = slot_value(c, pos++); if (!is_nan(nv)) _b0 = nv; ++cur; }
          nextvalue = _b0; b0 = _b0 - _s0; rowNum = cur - 1; s0 = _s0;
        } else if (is_nan(b0)) {
          double _b0 = NaNv; int cur = s;
          while (cur <= e && is_nan(_b0)) { const double nv = slot_value(c, pos++); if (!is_nan(nv)) _b0 = nv; ++cur; }
          nextvalue = _b0; b0 = _b0 - s0; rowNum = cur - 1;
        } else nextvalue = slot_value(c, pos++);
        if (!is_nan(b0)) {
          while (rowNum <= e) {
            if (!is_nan(nextvalue)) {
              const double _s0 = sf * nextvalue + (1 - sf) * (s0 + b0);
              b0 = tf * (_s0 - s0) + (1 - tf) * b0;
              s0 = _s0;
            }
            nextvalue = (pos <= e) ? slot_value(c, pos) : NaNv; ++pos;
            ++rowNum;
          }
          smoothed = s0;
        }
      }
      return smoothed;
    }
    case FN_PREDICT_LINEAR: {                             // PredictLinearChunkedFunctionD / L :1496-1590
      double sumX = NaNv, sumY = NaNv, sumXY = NaNv, sumX2 = NaNv; int32_t counter = 0;
      for (int ci = a; ci <= f && ci < cHi; ++ci) {
        const ChunkDesc& c = D[ci];
        int s, e; if (!chunk_rows(c, wStart, wEnd, s, e)) continue;
        for (int r = s; r <= e; ++r) {
          const double v = slot_value(c, r);
          if (!lng && is_nan(v)) continue;
          const double x = (double)(ts_apply(c, r) - wEnd) / 1000.0;
          if (is_nan(sumY)) { sumY = v; sumX = x; sumXY = x * v; sumX2 = x * x; }
          else { sumY += v; sumX += x; sumXY += x * v; sumX2 += x * x; }
          counter += 1;
        }
      }
      const double covXY = sumXY - sumX * sumY / (double)counter;
      const double varX = sumX2 - sumX * sumX / (double)counter;
      const double slope = covXY / varX;
      const double intercept = sumY / (double)counter - slope * sumX / (double)counter;
      return counter >= 2 ? slope * q.p0 + intercept : NaNv;
    }
  }
  return NaNv;
}

// One output window of one series.  D[cLo..cHi) are the chunks that intersect [start - window, end].
__device__ __forceinline__ double eval_window(const ChunkDesc* D, int cLo, int cHi, const QueryParams& q, int k) {
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
  const int64_t wEnd = q.start + (int64_t)k * q.step;
  int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
  const int64_t wStart = wEnd - winDur;
  // chunk set of the window (WindowedChunkIterator.nextWindow, closed form under time-ordered chunks)
  int a = cLo, f;
  if (cHi - cLo <= 8) {
    while (a < cHi && D[a].end_time < wStart) ++a;
    f = a; while (f < cHi - 1 && D[f].end_time < wEnd) ++f;
  } else {
    int lo = cLo, hi = cHi; while (lo < hi) { int m = (lo + hi) >> 1; if (D[m].end_time < wStart) lo = m + 1; else hi = m; } a = lo;
    lo = a; hi = cHi; while (lo < hi) { int m = (lo + hi) >> 1; if (D[m].end_time < wEnd) lo = m + 1; else hi = m; }
    f = lo < cHi - 1 ? lo : cHi - 1;
  }
  const int fn = q.fn;
  if (fn >= FN_STDDEV || q.long_values) return eval_window_ext(D, cLo, cHi, q, k, a, f, wStart, wEnd);
  const bool counterPath = ((fn == FN_RATE || fn == FN_INCREASE) && q.cumulative) || fn == FN_DELTA;

  if (fn == FN_LAST || fn == FN_TIMESTAMP) {              // RangeFunction.scala:603-613, 708-716
    int64_t lastTs = -1; double lastVal = NaNv, tsVal = NaNv;
    for (int ci = a; ci <= f && ci < cHi; ++ci) {
      const ChunkDesc& c = D[ci];
      bool ex; int idx = ts_search(c, wEnd, ex);
      int e = ex ? idx : idx - 1; if (e > c.num_rows - 1) e = c.num_rows - 1;
      if (e >= 0) {
        const int64_t t = ts_apply(c, e);
        if (fn == FN_TIMESTAMP) tsVal = (double)t / 1000.0;
        else if (t >= wStart && t > lastTs) { lastTs = t; lastVal = slot_value(c, e); }
      }
    }
    return fn == FN_TIMESTAMP ? tsVal : lastVal;
  }

  if (counterPath) {                                      // CounterChunkedRangeFunction + ChunkedRateFunctionBase
    int32_t numSamples = 0; int64_t loT = INT64_MAX, hiT = 0; double loV = NaNv, hiV = NaNv;
    bool some = false; double corrLast = 0.0, corr = 0.0; // correctionMeta
    for (int ci = a; ci <= f && ci < cHi; ++ci) {
      const ChunkDesc& c = D[ci];
      bool ex; const int s = ts_search(c, wStart, ex);
      int idx = ts_search(c, wEnd, ex);
      int e = ex ? idx : idx - 1; if (e > c.num_rows - 1) e = c.num_rows - 1;
      // detectDropAndCorrection, DoubleVector.scala:177-187
      if (some) { const double first = c.first_val; if (is_nan(first) || first < corrLast) corr = corr + corrLast; }
      if (s <= e) {
        const int64_t tS = ts_apply(c, s), tE = ts_apply(c, e);
        bool skip = false;
        if (fn != FN_DELTA && s == 0 && e == 0 && is_nan(slot_value(c, 0))) skip = true;    // RateFunctions.scala:255-256
        if (!skip && (tS < loT || tE > hiT)) {
          numSamples += e - s + 1;
          if (tS < loT) {
            loT = tS;
            if (fn == FN_DELTA) loV = slot_value(c, s);
            else { const double base = (c.dropped ? c.corr_slots[s] : slot_value(c, s)); loV = some ? base + corr : base; }
          }
          if (tE > hiT) {
            hiT = tE;
            if (fn == FN_DELTA) hiV = slot_value(c, e);
            else { const double base = (c.dropped ? c.corr_slots[e] : slot_value(c, e)); hiV = some ? base + corr : base; }
          }
        }
      }
      // updateCorrection, DoubleVector.scala:190-195 / 375-391
      if (c.dropped) { corrLast = c.upd_last; corr = (some ? corr : 0.0) + c.upd_corr; }
      else { corrLast = c.last_val; corr = some ? corr : 0.0; }
      some = true;
    }
    if (hiT > loT) {
      const int64_t cws = q.inclusive ? wStart : wStart - 1;
      return extrapolated_rate(cws, wEnd, numSamples, loT, loV, hiT, hiV, fn != FN_DELTA, fn == FN_RATE);
    }
    return NaNv;
  }

  // TimeRangeFunction family: sum / avg / count / min / max / delta-schema rate+increase
  double sum = NaNv; int32_t count = 0; double countD = NaNv; double mn = NaNv, mx = NaNv;
  for (int ci = a; ci <= f && ci < cHi; ++ci) {
    const ChunkDesc& c = D[ci];
    bool ex; const int s = ts_search(c, wStart, ex);
    int idx = ts_search(c, wEnd, ex);
    int e = ex ? idx : idx - 1; if (e > c.num_rows - 1) e = c.num_rows - 1;
    if (s > e) continue;
    if (fn == FN_MIN || fn == FN_MAX) {
      if (fn == FN_MIN) for (int r = s; r <= e; ++r) mn = min_ignore_nan(mn, slot_value(c, r));
      else for (int r = s; r <= e; ++r) mx = max_ignore_nan(mx, slot_value(c, r));
    } else {
      int cnt; const double cs = chunk_sum(c, s, e, cnt);
      if (fn == FN_COUNT) { if (is_nan(countD)) countD = 0.0; countD += (double)cnt; }
      else {                                              // AggrOverTimeFunctions.scala:568-570
        if (!is_nan(cs) && is_nan(sum)) sum = 0.0;
        sum += cs;
        count += cnt;
      }
    }
  }
  switch (fn) {
    case FN_SUM: case FN_INCREASE: return sum;
    case FN_RATE: { const int64_t cws = q.inclusive ? wStart : wStart - 1; return sum / (double)(wEnd - cws) * 1000.0; }
    case FN_AVG: return count > 0 ? sum / (double)count : (is_nan(sum) ? sum : 0.0);      // AggrOverTimeFunctions.scala:1000
    case FN_COUNT: return countD;
    case FN_MIN: return mn;
    case FN_MAX: return mx;
  }
  return NaNv;
}

} // namespace filo
