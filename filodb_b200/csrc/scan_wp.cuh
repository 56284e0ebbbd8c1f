// v4 scan path ("wp": warp pipeline).  One warp owns one series end to end; warps never synchronise with each other.
//
// Each warp runs its own three-buffer pipeline in shared memory:
//   R  the series' record (ChunkSetInfo entries + BinaryVectors verbatim), filled by ONE cp.async.bulk (TMA 1-D) per series that is
//      issued as soon as the previous record has been decoded, i.e. it is in flight during the previous series' window phase;
//   V  the decoded rows, laid out per chunk with zero rows in between so that clamped windows read +0.0 instead of testing bounds
//      (x + 0.0 is exact for an accumulator that started at +0.0), skewed by one pad slot per 8 rows: both the 8-byte row stores of the
//      group decode (lane stride 8 rows) and the 8-byte row loads of the window blocks (lane stride 8 windows) then walk the banks with
//      an odd stride of 9 words -- conflict-free;
//   O  the series' T results, leaving with one cp.async.bulk store that overlaps the next series' decode.
// Phases of a series (all 32 lanes, only __syncwarp between them):
//   setup    lane c = chunk c: header parse, regularity checks, window plan (touch interval, block list, row positions).  The plan
//            depends on (init, nrows, endTime) of the chunks only, so it is reused while consecutive series share those (memo);
//   decode   lane = NibblePack group (two groups per lane): branch-free field extraction, XOR prefix inside the group, warp-wide
//            XOR scan over the group totals, rows stored once;
//   windows  item = (chunk, block of 8 windows), two items per lane: register-blocked sequential sums in the reference's row order
//            (DoubleVector.scala:243-253, AggrOverTimeFunctions.scala:560-571).  A window that takes rows from two chunks gets one
//            partial sum from each chunk's block list; a short fix-up pass adds them in chunk order.
// Anything outside this fast path (irregular timestamps, DDV-long values, > 4 chunks, NaN / Inf / denormal / zero values, windows
// shorter than 9 rows, windows over three chunks ...) is appended to the fallback list and answered by the v2 kernel into the same
// output buffer, exactly as the tile kernel does.
#pragma once
#include "scan_tile.cuh"
#include "scan_wp_layout.h"

namespace filo {

// shared-memory word loads by byte offset (keeps the field extraction in the shared window: LDS, not generic loads)
#ifdef FILO_CUSIM
__device__ __forceinline__ uint32_t wp_soff(const void* p) { return (uint32_t)(reinterpret_cast<const uint8_t*>(p) - smem); }
__device__ __forceinline__ uint32_t wp_lds32(uint32_t off) { uint32_t v; std::memcpy(&v, smem + off, 4); return v; }
#else
__device__ __forceinline__ uint32_t wp_soff(const void* p) { return smem_u32(p); }
__device__ __forceinline__ uint32_t wp_lds32(uint32_t off) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(off)); return v; }
#endif

__device__ __forceinline__ int wp_vidx(int p) { return p + (p >> 3); }
// results are written once and not read again by the scan: streaming store
#ifdef FILO_CUSIM
__device__ __forceinline__ void wp_store_result(double* g, double v) { *g = v; }
#else
__device__ __forceinline__ void wp_store_result(double* g, double v) { __stcs(g, v); }
#endif

// finish of one window of a SUM-class function from (sum over the rows, number of rows); values are known to be finite, normal and
// of moderate magnitude (the decode checked), so the invariant division needs no range test
template <int FN>
__device__ __forceinline__ double wp_finish(double cs, int nn, double div, double rcp, double scale, int nfull, double rcpn, bool raw) {
  // raw blocks pass div = rcp = scale = 1: the sequence below then returns cs itself, bit for bit
  if (FN == FN_RATE) { const double q0 = __dmul_rn(cs, rcp); const double r = __fma_rn(-q0, div, cs); return __dmul_rn(__fma_rn(r, rcp, q0), scale); }
  if (FN == FN_COUNT) return (double)nn;
  if (FN == FN_AVG) {
    if (raw) return cs;
    if (nn == nfull) { const double q0 = __dmul_rn(cs, rcpn); const double r = __fma_rn(-q0, (double)nfull, cs); return __fma_rn(r, rcpn, q0); }
    return cs / (double)nn;
  }
  return cs;
}

// the last two row groups of a block: rows 8q .. Wr + 7 with Wr = 8q + U.  Row t of group q feeds windows max(0, t - U) .. 7, row t < U of
// group q + 1 feeds windows 8 + t - U .. 7 (window w takes rows w .. w + Wr of the block)
template <int U>
__device__ __forceinline__ void wp_block_tail(const double* __restrict__ pa, const double* __restrict__ pb, double a[WP_R], double b[WP_R]) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const double va = pa[t], vb = pb[t];
#pragma unroll
    for (int w = 0; w < 8; ++w) if (w >= t - U) { a[w] += va; b[w] += vb; }
  }
#pragma unroll
  for (int t = 0; t < U; ++t) {
    const double va = pa[9 + t], vb = pb[9 + t];
#pragma unroll
    for (int w = 0; w < 8; ++w) if (w >= 8 + t - U) { a[w] += va; b[w] += vb; }
  }
}

// two blocks per lane: a[w] / b[w] = sum of rows w .. w + Wr (in row order, starting from +0.0) of the block at pa / pb.  Wr >= 8.
__device__ __forceinline__ void wp_block_pair(const double* __restrict__ pa, const double* __restrict__ pb, int Wr, double a[WP_R], double b[WP_R]) {
#pragma unroll
  for (int w = 0; w < 8; ++w) { a[w] = 0.0; b[w] = 0.0; }
#pragma unroll
  for (int t = 0; t < 8; ++t) {                        // group 0: row t feeds windows 0 .. t
    const double va = pa[t], vb = pb[t];
#pragma unroll
    for (int w = 0; w <= t; ++w) { a[w] += va; b[w] += vb; }
  }
  const int q = Wr >> 3;
  pa += 9; pb += 9;
  for (int m = 1; m < q; ++m, pa += 9, pb += 9) {      // full groups: every window
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const double va = pa[t], vb = pb[t];
#pragma unroll
      for (int w = 0; w < 8; ++w) { a[w] += va; b[w] += vb; }
    }
  }
  switch (Wr & 7) {
    case 0: wp_block_tail<0>(pa, pb, a, b); break;
    case 1: wp_block_tail<1>(pa, pb, a, b); break;
    case 2: wp_block_tail<2>(pa, pb, a, b); break;
    case 3: wp_block_tail<3>(pa, pb, a, b); break;
    case 4: wp_block_tail<4>(pa, pb, a, b); break;
    case 5: wp_block_tail<5>(pa, pb, a, b); break;
    case 6: wp_block_tail<6>(pa, pb, a, b); break;
    default: wp_block_tail<7>(pa, pb, a, b); break;
  }
}

// per-series parse, lane c = chunk c of the chunks in range: what the record says about the chunk, and whether the series qualifies
struct WpParsed {
  bool regular, have, any_raw; bool irr;            // irr: some chunk's timestamps are not on the query's step grid (DDV residuals or slope != step)
  int tslope; uint32_t toff;                         // timestamp vector: slope, byte offset in R
  int n, cLo;
  int64_t init, end_time; int nrows, num_rows, vbytes, ng, vwire, dropped, tlen, grp_base, ngroups; uint32_t voff, w12;
};
// STRICT (SUM class): endTime covers the chunk's rows and lies before the next chunk's first row, so that "has a row in the window" and
// "is in the window's chunk set" (ChunkSetInfo.scala:481-510) coincide; the counter class evaluates the chunk set itself
template <bool STRICT, bool IRR = false>
__device__ __forceinline__ WpParsed wp_parse(const uint8_t* R, const QueryParams& q, bool staged, int lane) {
  const unsigned FULL = 0xffffffffu;
  WpParsed P;
  bool regular = staged;
  int n = 0, cLo = 0;
  if (staged) {
    const RecordHeader* h = reinterpret_cast<const RecordHeader*>(R);
    const ChunkEntry* Eall = reinterpret_cast<const ChunkEntry*>(R + sizeof(RecordHeader));
    const int nch = (int)h->n_chunks;
    regular = nch <= 32 && (IRR || (h->flags & REC_ALL_TS_CONST) != 0);
    const int64_t t1 = q.start - q.window, t2 = q.end;
    bool below = false, within = false;
    if (regular && lane < nch) { below = Eall[lane].end_time < t1; within = !below && Eall[lane].start_time <= t2; }
    const unsigned mb = __ballot_sync(FULL, below), mw = __ballot_sync(FULL, within);
    cLo = __ffs((int)~mb) - 1; if (cLo < 0) cLo = 32;                         // chunks are time-ordered: `below` is a prefix
    const unsigned rest = cLo < 32 ? (mw >> cLo) : 0u;
    n = __ffs((int)~rest) - 1; if (n < 0) n = 32;
    if (n > WP_MAXC) regular = false;
  }
  const int c = lane;
  bool have = regular && c < n;
  int64_t init = 0, end_time = 0; int nrows = 0, num_rows = 0, vbytes = 0, ng = 0, vwire = 0, dropped = 0, tlen = 0; uint32_t voff = 0, w12 = 0;
  bool okc = true, irrc = false; int slope = 0; uint32_t toff = 0;
  if (have) {
    const ChunkEntry& e = reinterpret_cast<const ChunkEntry*>(R + sizeof(RecordHeader))[cLo + c];
    const uint8_t* tv = R + e.ts_off; const uint8_t* vv = R + e.val_off;
    const uint32_t vw4 = ld32(vv + 4);
    vwire = (int)(vw4 & 0xffff); dropped = (int)((vw4 >> 31) & 1);
    toff = e.ts_off;
    const int twire = (int)(ld32(tv + 4) & 0xffff);
    if (!IRR || twire == WIRE_DDV_CONST) { tlen = (int)ld32(tv + 8); init = (int64_t)ld64_a4(tv + 12); slope = (int)ld32(tv + 20); if (IRR && twire != WIRE_DDV_CONST) okc = false; }
    else if (twire == WIRE_DDV) {                          // DeltaDeltaVector.scala:138-156: +8 init, +16 slope, +20 IntBinaryVector of residuals
      init = (int64_t)ld64(tv + 8); slope = (int)ld32(tv + 16); tlen = int_length(tv + 20); irrc = true;
      const int nb = (int)((ld32(tv + 24) >> 16) & 0x7f);
      if (!(nb == 2 || nb == 4 || nb == 8 || nb == 16 || nb == 32)) okc = false;
    } else okc = false;
    if (IRR && (int64_t)slope != q.step) irrc = true;
    end_time = e.end_time; num_rows = e.num_rows; voff = e.val_off;
    vbytes = (int)ld32(tv) + 4 + (int)ld32(vv) + 4;
    int vlen = 0;
    if (vwire == WIRE_XOR) { vlen = (int)ld32(vv + XOR_OFF_N); w12 = ld32(vv + XOR_OFF_NGROUPS); ng = (int)(w12 & 0xffff); if (ng != (vlen + 6) / 8) okc = false; }
    else if (vwire == WIRE_RAW64) vlen = ((int)ld32(vv) - 4) / 8;
    else okc = false;
    if ((!IRR && (int64_t)slope != q.step) || slope <= 0 || tlen <= 0 || vlen <= 0 || num_rows <= 0) okc = false;
    nrows = num_rows < tlen ? num_rows : tlen;
    if (vlen != nrows) okc = false;                       // the decode writes every row of the vector
    if (STRICT && end_time < init + (int64_t)(nrows - 1) * q.step) okc = false;
  }
  {
    const int64_t endp = __shfl_up_sync(FULL, end_time, 1);
    if (STRICT && have && c > 0 && !(endp < init)) okc = false;
    if (!__all_sync(FULL, okc)) regular = false;
  }
  have = have && regular;
  if (!have) { ng = 0; nrows = 0; }
  // group slots: exclusive prefix over the chunks
  int grp_base = ng;
  { const int a0 = __shfl_sync(FULL, ng, 0), a1 = __shfl_sync(FULL, ng, 1), a2 = __shfl_sync(FULL, ng, 2), a3 = __shfl_sync(FULL, ng, 3);
    grp_base = (c > 0 ? a0 : 0) + (c > 1 ? a1 : 0) + (c > 2 ? a2 : 0);
    if (a0 + a1 + a2 + a3 > WP_MAXG) regular = false; }
  P.ngroups = __shfl_sync(FULL, grp_base + ng, WP_MAXC - 1);
  P.any_raw = __any_sync(FULL, have && vwire == WIRE_RAW64);
  P.irr = IRR && __any_sync(FULL, have && irrc); P.tslope = slope; P.toff = toff;
  P.regular = regular; P.have = have && regular; P.n = n; P.cLo = cLo; P.init = init; P.end_time = end_time; P.nrows = nrows; P.num_rows = num_rows;
  P.vbytes = vbytes; P.ng = ng; P.vwire = vwire; P.dropped = dropped; P.tlen = tlen; P.grp_base = grp_base; P.voff = voff; P.w12 = w12;
  return P;
}

// Decode of a series into V: lane = NibblePack group (slots lane and lane + 32, described by dd_dst / dd_inf, see the plan); raw f64
// vectors are copied.  Returns the AND over every value v of hi(v) ^ (hi(v) << 1): bit 30 is set while every value has exponent bits
// 10 and 9 different, i.e. 2^-511 <= |v| < 2^513 (finite, normal, not zero).  DROPS (counter class): counter drops inside drop-flagged
// chunks (DoubleVector.scala:330-340) are recorded as (row, amount) in DR[chunk]; row r drops when (NaN -> 0) of it is below
// (NaN -> 0) of row r - 1, the amount is the value before the drop.
template <bool DROPS>
__device__ __forceinline__ uint32_t wp_decode(const uint8_t* R, double* V, const WpChunk* CD, uint64_t* xtab, const int dd_dst[2], const int dd_inf[2],
                                              int n, bool any_raw, int lane, TileDrops* DR) {
  uint32_t okbits = 0xffffffffu;
  uint64_t d[2][8];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const bool active = dd_inf[jj] & 1;
    const WpChunk& ch = CD[(dd_inf[jj] >> 1) & 3];
    const uint8_t* gp = R;
    if (active) gp = R + ch.grp_off + reinterpret_cast<const uint16_t*>(R + ch.tab_off)[dd_inf[jj] >> 8];
    const uint32_t mask = active ? gp[0] : 0u;
    const uint32_t hdr = gp[1];
    const uint32_t numBits = ((hdr >> 4) + 1) * 4;
    const uint32_t tz = (hdr & 0x0f) * 4;
    const uint64_t fm = (~0ull >> (64 - numBits)) << tz;       // the field's bits in the value
    // bit address (in shared memory) of the 64-bit window that has field 0 at bit tz
    uint32_t xb = (wp_soff(gp + 2) << 3) - tz;
    uint64_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool on = (mask >> i) & 1u;
      const uint32_t wa = (xb >> 3) & ~3u;
      const uint32_t w0 = wp_lds32(wa), w1 = wp_lds32(wa + 4), w2 = wp_lds32(wa + 8);
      const uint32_t lo = __funnelshift_r(w0, w1, xb), hi = __funnelshift_r(w1, w2, xb);
      const uint64_t f = on ? fm : 0ull;
      x ^= (((uint64_t)hi << 32) | lo) & f;                    // running XOR of the fields, already shifted by tz
      xb += on ? numBits : 0u;
      d[jj][i] = x;
    }
  }
  // exclusive XOR scan of the group totals over the 64 slots
  uint64_t i0x = d[0][7], i1x = d[1][7];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t y0 = shfl_up_u64(i0x, o), y1 = shfl_up_u64(i1x, o);
    if (lane >= o) { i0x ^= y0; i1x ^= y1; }
  }
  const uint64_t tot0 = shfl_u64(i0x, 31);
  const uint64_t ex0 = i0x ^ d[0][7], ex1 = i1x ^ d[1][7] ^ tot0;
  xtab[lane] = ex0; xtab[32 + lane] = ex1;
  __syncwarp();
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const bool active = dd_inf[jj] & 1;
    const int ci = (dd_inf[jj] >> 1) & 3;
    const WpChunk& ch = CD[ci];
    // value before the group = first ^ (prefix at the slot) ^ (prefix at the chunk's first slot)
    const uint64_t pre = ch.first ^ (jj ? ex1 : ex0) ^ xtab[active ? ch.grp_base : 0];
    uint64_t* dst = reinterpret_cast<uint64_t*>(V) + dd_dst[jj];
    const int t = (dd_inf[jj] >> 3) & 7;                       // rows t' with t + t' >= 8 sit one pad slot further
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t b = d[jj][i] ^ pre;
        dst[i + ((t + i) >> 3)] = b;
        const uint32_t h = (uint32_t)(b >> 32);
        okbits &= h ^ (h << 1);
      }
      if ((dd_inf[jj] >> 8) == 0) { dst[t == 0 ? -2 : -1] = ch.first; const uint32_t h = (uint32_t)(ch.first >> 32); okbits &= h ^ (h << 1); }
      if (DROPS && ch.dropped) {
        const int g = dd_inf[jj] >> 8;
        const int nleft = ch.nrows - 1 - g * 8;                // rows past the chunk are not data
        double prevv = nan0(__longlong_as_double((long long)pre));
        TileDrops& D = DR[ci];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double cur = nan0(__longlong_as_double((long long)(d[jj][i] ^ pre)));
          if (i < nleft && cur < prevv) {                      // rare: position and amount (the value before the drop)
            const int at = atomicAdd(&D.n, 1);
            if (at < TILE_MAXDROP) { D.pos[at] = 1 + g * 8 + i; D.amt[at] = prevv; }
          }
          prevv = cur;
        }
      }
    }
  }
  // raw f64 vectors: plain copy
  for (int ci = 0; any_raw && ci < n; ++ci) {
    const WpChunk& ch = CD[ci];
    if (ch.wire != WIRE_RAW64) continue;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(R + ch.val_off + 8);
    const bool drp = DROPS && ch.dropped;
    for (int r = lane; r < ch.nrows; r += 32) {
      const uint64_t b = src[r];
      reinterpret_cast<uint64_t*>(V)[wp_vidx(ch.rowpos + r)] = b;
      const uint32_t h = (uint32_t)(b >> 32);
      okbits &= h ^ (h << 1);
      if (drp && r > 0) {
        const double cur = nan0(__longlong_as_double((long long)b)), prevv = nan0(__longlong_as_double((long long)src[r - 1]));
        if (cur < prevv) { TileDrops& D = DR[ci]; const int at = atomicAdd(&D.n, 1); if (at < TILE_MAXDROP) { D.pos[at] = r; D.amt[at] = prevv; } }
      }
    }
  }
  return okbits;
}

// window block `it` of the plan: V index of its first row, byte offset (inside the warp's region) of its first result slot, and
// inf = (jEnd + 1) | raw << 4 | skew phase of the first slot << 5 | chunk << 8 | first window << 10, where slots 0 .. jEnd are stored
__device__ __forceinline__ void wp_item(const WpChunk* CD, const WpSmem& L, int it, int items, int psi, int& pp, int& op, int& inf) {
  const bool active = it < items;
  int ci = 0;                                             // the last chunk with blocks whose blk0 <= it
  if (CD[1].nblk > 0 && it >= CD[1].blk0) ci = 1;
  if (CD[2].nblk > 0 && it >= CD[2].blk0) ci = 2;
  if (CD[3].nblk > 0 && it >= CD[3].blk0) ci = 3;
  const WpChunk& ch = CD[ci];
  const int b = active ? it - ch.blk0 : 0;
  const int k0 = ch.kT0 + WP_R * b;
  pp = ch.vidx0 + 9 * b;
  const bool tojz = b < ch.jzb;                           // raw sums to J (every slot of the block exists there)
  const bool raw = tojz || b >= ch.tb;
  op = tojz ? (int)WP_OFF_J + 8 * (ch.joff + WP_R * b) : (int)L.out + 8 * (k0 + ((k0 + psi) >> 3));
  int jEnd = !active ? -1 : tojz ? WP_R - 1 : ch.kT1 - k0;
  if (jEnd > WP_R - 1) jEnd = WP_R - 1;
  const int ot = tojz ? 0 : (k0 + psi) & 7;               // slots j with ot + j >= 8 sit one pad slot further
  inf = (jEnd + 1) | ((raw ? 1 : 0) << 4) | (ot << 5) | (ci << 8) | (k0 << 10);
}

// ---------------------------------------------------------------------------------------------------------------------
// SUM-class kernel: sum / avg / count_over_time, rate / increase on delta-temporality schemas.  No across-series aggregate.
// ---------------------------------------------------------------------------------------------------------------------
template <int FN, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
scan_wp_sum_kernel(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, int64_t n_series, QueryParams q,
                   double* __restrict__ out, WpSmem L, int64_t* __restrict__ fallback_list, unsigned long long* __restrict__ fallback_count,
                   unsigned long long* d_counters, int* d_err) {
  extern __shared__ __align__(128) uint8_t smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* wb = smem + (size_t)warp * L.per_warp;
  uint64_t* bar = reinterpret_cast<uint64_t*>(wb);
  WpChunk* CD = reinterpret_cast<WpChunk*>(wb + WP_OFF_DESC);
  uint64_t* xtab = reinterpret_cast<uint64_t*>(wb + WP_OFF_J);     // decode: exclusive XOR prefix per group slot (dead before J is written)
  double* J = reinterpret_cast<double*>(wb + WP_OFF_J);
  uint8_t* R = wb + WP_OFF_REC;
  double* V = reinterpret_cast<double*>(wb + L.vals);
  double* O = reinterpret_cast<double*>(wb + L.out);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncwarp();

  int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
  const double fdiv = (double)(q.inclusive ? winDur : winDur + 1), frcp = 1.0 / fdiv;     // RateFunctions.scala:436-442
  const int64_t S0 = q.start - winDur, E0 = q.start;
  const int64_t lastEnd = q.start + (int64_t)(q.T - 1) * q.step;
  StepDiv sd; sd.init(q.step);
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);

  // memo of the window plan (lane c holds chunk c's key; the plan itself stays in CD)
  int64_t m_init = 0, m_end = 0; int m_nrows = -1, m_n = -1, m_wire = -1; bool m_ok = false;
  // per-lane work items of the plan: two decode slots and the two window blocks of the first pass (see the plan)
  int dd_dst[2] = {0, 0}, dd_inf[2] = {0, 0}, wi_pp[2] = {0, 0}, wi_op[2] = {0, 0}, wi_inf[2] = {0, 0};
  int gz[3] = {-1, -1, -1}; bool gz_all = true;      // this lane's zero rows (V indices) when the plan has at most 96 of them
  int p_Wr = 0, p_items = 0, p_nfull = 0, p_psi = 0; double p_rcpn = 0.0; bool p_gaps = true, p_oal = false;
  int64_t rows_scanned = 0, bytes_scanned = 0;
  uint32_t parity = 0;

  auto issue = [&](int64_t off, uint32_t sz) {            // lane 0: fetch a record into R
    mbar_expect_tx(bar, sz);
    tma_load_1d(R, arena + off, sz, bar);
  };
  int64_t cur_off = 0; uint32_t cur_sz = 0;
  if (s < n_series) { cur_off = rec_off[s]; cur_sz = (uint32_t)(rec_off[s + 1] - cur_off); }
  if (s < n_series && cur_sz <= L.rec_cap && lane == 0) issue(cur_off, cur_sz);

  for (; s < n_series; s += nwarps) {
    const int64_t sn = s + nwarps;
    int64_t nxt_off = 0; uint32_t nxt_sz = 0;
    if (sn < n_series) { nxt_off = rec_off[sn]; nxt_sz = (uint32_t)(rec_off[sn + 1] - nxt_off); }
    const bool staged = cur_sz <= L.rec_cap;
    if (staged) { mbar_wait(bar, parity); parity ^= 1; }
    // ------------------------------------------------------------------------------------------------ setup (lane c = chunk c)
    const WpParsed P = wp_parse<true>(R, q, staged, lane);
    bool regular = P.regular;
    const bool have = P.have; const int n = P.n, c = lane;
    const int64_t init = P.init, end_time = P.end_time;
    const int nrows = P.nrows, num_rows = P.num_rows, vbytes = P.vbytes, ng = P.ng, vwire = P.vwire, grp_base = P.grp_base, ngroups = P.ngroups;
    const uint32_t voff = P.voff, w12 = P.w12; const bool any_raw = P.any_raw;
    // ---- window plan, reused while the chunk shapes repeat
    const bool samec = !(c < n) || (init == m_init && nrows == m_nrows && end_time == m_end && vwire == m_wire);
    const bool same_all = __all_sync(FULL, samec);
    const bool same = m_ok && n == m_n && same_all;
    if (regular && !same) {
      m_init = init; m_end = end_time; m_nrows = nrows; m_n = n; m_wire = vwire; m_ok = false;
      int64_t s0 = 0, e0 = 0;
      if (have) { s0 = sd.ceil_div(S0 - init); e0 = sd.floor_div(E0 - init); }
      const int Wr = (int)(e0 - s0);
      int64_t kT0 = -e0; if (kT0 < 0) kT0 = 0;
      int64_t kT1 = (int64_t)(nrows - 1) - s0; if (kT1 > q.T - 1) kT1 = q.T - 1;
      const bool touch = have && kT0 <= kT1;
      if (!touch) { kT0 = 0x3fffffff; kT1 = -1; }
      const int Wr0 = __shfl_sync(FULL, Wr, 0);
      bool okp = !have || Wr == Wr0;
      if (Wr0 < 8 || (uint32_t)Wr0 + 1 > L.jcap) okp = false;
      // touched chunks must be contiguous, and a window may take rows from at most two chunks
      const unsigned tm = __ballot_sync(FULL, touch) & 0xfu;
      if (tm != 0) { const unsigned lowbit = tm & (0u - tm); const unsigned filled = tm + lowbit; if ((filled & (filled - 1)) != 0) okp = false; }
      const int64_t kT1p = __shfl_up_sync(FULL, kT1, 1), kT1pp = __shfl_up_sync(FULL, kT1, 2), kT0n = __shfl_down_sync(FULL, kT0, 1);
      const bool prev_t = c > 0 && ((tm >> (c - 1)) & 1u), next_t = c + 1 < WP_MAXC && ((tm >> (c + 1)) & 1u);
      if (touch && c >= 2 && ((tm >> (c - 2)) & 1u) && !(kT1pp < kT0)) okp = false;
      int64_t ownLo = kT0, ownHi = kT1;
      if (touch && prev_t && kT1p + 1 > ownLo) ownLo = kT1p + 1;
      if (touch && next_t && kT0n - 1 < ownHi) ownHi = kT0n - 1;
      const int hs = touch ? (int)(ownLo - kT0) : 0;
      const int nblk = touch ? (int)((kT1 - kT0 + WP_R) / WP_R) : 0;
      // whole blocks: [0, jzb) raw to J (the head share, rounded up), [tb, nblk) raw to O (from the block of ownHi + 1), own blocks in between
      const int jzb = (hs + WP_R - 1) / WP_R;
      const int tb = (touch && ownHi < kT1) ? (int)((ownHi + 1 - kT0) / WP_R) : nblk;
      if (touch && jzb > tb) okp = false;
      int blk0, items, joff, jtot, cov;
      { const int a0 = __shfl_sync(FULL, nblk, 0), a1 = __shfl_sync(FULL, nblk, 1), a2 = __shfl_sync(FULL, nblk, 2), a3 = __shfl_sync(FULL, nblk, 3);
        blk0 = (c > 0 ? a0 : 0) + (c > 1 ? a1 : 0) + (c > 2 ? a2 : 0); items = a0 + a1 + a2 + a3; }
      { const int z = jzb * WP_R;
        const int a0 = __shfl_sync(FULL, z, 0), a1 = __shfl_sync(FULL, z, 1), a2 = __shfl_sync(FULL, z, 2), a3 = __shfl_sync(FULL, z, 3);
        joff = (c > 0 ? a0 : 0) + (c > 1 ? a1 : 0) + (c > 2 ? a2 : 0); jtot = a0 + a1 + a2 + a3; }
      { const int z = touch ? (int)(kT1 - kT0 + 1) - hs : 0;      // windows this chunk is the first to touch
        cov = z;
#pragma unroll
        for (int o = 1; o < WP_MAXC; o <<= 1) cov += __shfl_xor_sync(FULL, cov, o); }
      if ((uint32_t)jtot > L.jcap) okp = false;
      if (L.alias && items > 64) okp = false;                 // O takes V's place: every block is summed before the first result is stored
      // row positions: chunk after chunk, Wr .. Wr + 7 zero rows in between, every chunk's block 0 at a multiple of 8
      const int fr = touch ? (int)(s0 + kT0) : 0;              // first row of block 0 (may be negative: zero rows in front)
      int rowpos = 0;
      {
        int base = 0;                                           // first position this chunk's rows may take
#pragma unroll
        for (int cc = 0; cc < WP_MAXC; ++cc) {
          int x;
          if (cc == 0) { x = fr < 0 ? -fr : ((8 - (fr & 7)) & 7); }
          else { x = base + ((-(base + fr)) & 7); }
          if (c == cc) rowpos = x;
          const int nb = __shfl_sync(FULL, x + nrows + Wr0, cc);   // (lane cc's own x is the valid one)
          base = nb;
        }
      }
      const int pend = __shfl_sync(FULL, rowpos + nrows, n > 0 ? n - 1 : 0) + Wr0 + 8;
      if ((uint32_t)(pend + (pend >> 3) + 2) > L.vcap) okp = false;
      if (!__all_sync(FULL, okp)) regular = false;
      if (regular) {
        if (c < WP_MAXC) {
          WpChunk& d = CD[c];
          d.kT0 = (int)kT0; d.kT1 = (int)kT1; d.ownLo = (int)ownLo; d.ownHi = (int)ownHi; d.blk0 = blk0; d.nblk = nblk;
          d.vidx0 = wp_vidx(rowpos + fr); d.rowpos = rowpos; d.nrows = nrows; d.s0 = (int)s0; d.e0 = (int)e0; d.joff = joff; d.hs = hs; d.jzb = jzb; d.tb = tb;
        }
        p_gaps = __shfl_sync(FULL, cov, 0) != q.T;
        // O is skewed like V (one pad slot per 8 windows, block starts of the first touched chunk on the 9-word grid): the 8-byte result
        // stores of a warp (lane stride 8 windows) then spread over the banks.  p_oal: every chunk's blocks start on that grid
        { const int first_t = tm ? __ffs((int)tm) - 1 : 0;
          p_psi = (-(int)__shfl_sync(FULL, (int)(touch ? kT0 : 0), first_t)) & 7;
          p_oal = __all_sync(FULL, !touch || (((int)kT0 + p_psi) & 7) == 0); }
        p_Wr = Wr0; p_items = items; p_nfull = Wr0 + 1; p_rcpn = 1.0 / (double)(Wr0 + 1);
        __syncwarp();
        // zero rows: in front of chunk 0, between chunks, behind the last chunk (+ slack the last block's unused windows read).  They are
        // written again for every series (the group decode runs up to 7 rows past a chunk; with O in V's place the results land on them)
        {
          int tot = 0;
          gz[0] = gz[1] = gz[2] = -1;
          for (int g = 0; g <= n; ++g) {
            const int g0 = g == 0 ? 0 : CD[g - 1].rowpos + CD[g - 1].nrows;
            const int g1 = g == n ? pend : CD[g].rowpos;
#pragma unroll
            for (int u = 0; u < 3; ++u) { const int i = u * 32 + lane - tot; if (i >= 0 && i < g1 - g0) gz[u] = wp_vidx(g0 + i); }
            tot += g1 - g0;
          }
          gz_all = tot <= 96;
        }
        // this lane's work items (they stay valid with the plan): decode slots lane, lane + 32 and window blocks lane, lane + 32
        {
          const int gb1 = __shfl_sync(FULL, have ? grp_base : 0x7fffffff, 1), gb2 = __shfl_sync(FULL, have ? grp_base : 0x7fffffff, 2),
                    gb3 = __shfl_sync(FULL, have ? grp_base : 0x7fffffff, 3);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int slot = jj * 32 + lane;
            const bool active = slot < ngroups;
            const int ci = active ? (slot >= gb1 ? 1 : 0) + (slot >= gb2 ? 1 : 0) + (slot >= gb3 ? 1 : 0) : 0;
            const int gbc = ci == 0 ? 0 : ci == 1 ? gb1 : ci == 2 ? gb2 : gb3;
            const int g = active ? slot - gbc : 0;
            const int pq = CD[ci].rowpos + 1 + g * 8;
            dd_dst[jj] = wp_vidx(pq);
            dd_inf[jj] = (active ? 1 : 0) | (ci << 1) | ((pq & 7) << 3) | (g << 8);      // active, chunk, skew phase of the first row, group in chunk
          }
#pragma unroll
          for (int X = 0; X < 2; ++X) wp_item(CD, L, X * 32 + lane, items, p_psi, wi_pp[X], wi_op[X], wi_inf[X]);
        }
        m_ok = true;
      }
    }
    if (!regular) {
      // declined: the v2 kernel answers this series
      if (lane == 0) { const unsigned long long slot = atomicAdd(fallback_count, 1ull); fallback_list[slot] = s; }
      __syncwarp();
      if (sn < n_series && nxt_sz <= L.rec_cap && lane == 0) issue(nxt_off, nxt_sz);
      cur_off = nxt_off; cur_sz = nxt_sz;
      continue;
    }
    // per-series parts of the descriptors
    if (c < WP_MAXC) {
      WpChunk& d = CD[c];
      d.grp_base = have ? grp_base : 0x7fffffff; d.ng = ng; d.wire = vwire; d.val_off = voff; d.dropped = P.dropped;
      if (have && vwire == WIRE_XOR) { const uint32_t po = w12 >> 16; d.first = ld64(R + voff + po); d.grp_off = voff + po + 8; d.tab_off = voff + XOR_OFF_GROUPTAB; }
      else { d.first = have ? ld64(R + voff + 8) : 0ull; d.grp_off = 0; d.tab_off = 0; }
    }
    // scan counters (CountingChunkInfoIterator, ChunkSetInfo.scala:336-380): every chunk in range is pulled, except one that starts
    // after the last window end
    int cnt_rows = 0, cnt_bytes = 0;
    { const int64_t endp = __shfl_up_sync(FULL, end_time, 1);
      if (have && !(c > 0 && !(endp < lastEnd))) { cnt_rows = num_rows; cnt_bytes = vbytes; } }
#pragma unroll
    for (int o = 1; o < WP_MAXC; o <<= 1) { cnt_rows += __shfl_xor_sync(FULL, cnt_rows, o); cnt_bytes += __shfl_xor_sync(FULL, cnt_bytes, o); }
    __syncwarp();
    // ------------------------------------------------------------------------------------------------ decode
    const uint32_t okbits = wp_decode<false>(R, V, CD, xtab, dd_dst, dd_inf, n, any_raw, lane, nullptr);
    const bool vals_ok = __all_sync(FULL, (okbits >> 30) & 1u);
    __syncwarp();
    // R is dead: fetch the next record behind the window phase
    if (sn < n_series && nxt_sz <= L.rec_cap && lane == 0) issue(nxt_off, nxt_sz);
    cur_off = nxt_off; cur_sz = nxt_sz;
    // zero rows (the last group of an XOR chunk decoded up to 7 rows past the chunk; results of the previous series when O is in V's place)
    if (gz_all) {
#pragma unroll
      for (int u = 0; u < 3; ++u) if (gz[u] >= 0) V[gz[u]] = 0.0;
    } else {
      const int pend = CD[n - 1].rowpos + CD[n - 1].nrows + p_Wr + 8;
      for (int g = 0; g <= n; ++g) {
        const int g0 = g == 0 ? 0 : CD[g - 1].rowpos + CD[g - 1].nrows;
        const int g1 = g == n ? pend : CD[g].rowpos;
        for (int pz = g0 + lane; pz < g1; pz += 32) V[wp_vidx(pz)] = 0.0;
      }
    }
    if (!vals_ok) {
      // NaN / Inf / zero / denormal / very large or small values: the literal kernel answers (it needs the NaN-aware sums)
      if (lane == 0) { const unsigned long long slot = atomicAdd(fallback_count, 1ull); fallback_list[slot] = s; }
      __syncwarp();
      continue;
    }
    if (lane == 0) { rows_scanned += cnt_rows; bytes_scanned += cnt_bytes; }
    __syncwarp();
    // ------------------------------------------------------------------------------------------------ windows
    const int psi = p_psi;
    auto oidx = [&](int k) -> int { return k + ((k + psi) >> 3); };
    {
      const int Wr = p_Wr;
      for (int it0 = 0; it0 < p_items; it0 += 64) {
        int ipp[2], iop[2], iinf[2];
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          if (it0 == 0) { ipp[X] = wi_pp[X]; iop[X] = wi_op[X]; iinf[X] = wi_inf[X]; }
          else wp_item(CD, L, it0 + X * 32 + lane, p_items, psi, ipp[X], iop[X], iinf[X]);
        }
        const double* pp[2]; double* op[2]; int jEnd[2], ot[2]; bool rawm[2];
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          pp[X] = V + ipp[X]; op[X] = reinterpret_cast<double*>(wb + iop[X]);
          jEnd[X] = (iinf[X] & 15) - 1; rawm[X] = (iinf[X] >> 4) & 1; ot[X] = (iinf[X] >> 5) & 7;
        }
        double a[WP_R], bb[WP_R];
        wp_block_pair(pp[0], pp[1], Wr, a, bb);
        __syncwarp();                                       // (O may sit on V: every lane has read its rows)
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          const double dv = rawm[X] ? 1.0 : fdiv, rc = rawm[X] ? 1.0 : frcp, sc = rawm[X] ? 1.0 : 1000.0;
#pragma unroll
          for (int j = 0; j < WP_R; ++j) {
            const double raw = X ? bb[j] : a[j];
            int nn = 1;
            if (FN == FN_AVG || FN == FN_COUNT) {
              const WpChunk& ch = CD[(iinf[X] >> 8) & 3];
              int lo = ch.s0 + (iinf[X] >> 10) + j; const int hi0 = lo + Wr; if (lo < 0) lo = 0;
              const int hi = hi0 > ch.nrows - 1 ? ch.nrows - 1 : hi0;
              nn = hi - lo + 1;
            }
            const double fin = wp_finish<FN>(raw, nn, dv, rc, sc, p_nfull, p_rcpn, rawm[X]);
            if (X) bb[j] = fin; else a[j] = fin;
          }
        }
        if (p_oal) {                                        // every block starts on O's 9-word grid: constant store offsets
#pragma unroll
          for (int j = 0; j < WP_R; ++j) { if (j <= jEnd[0]) op[0][j] = a[j]; if (j <= jEnd[1]) op[1][j] = bb[j]; }
        } else {
#pragma unroll
          for (int j = 0; j < WP_R; ++j) {
            if (j <= jEnd[0]) op[0][j + ((ot[0] + j) >> 3)] = a[j];
            if (j <= jEnd[1]) op[1][j + ((ot[1] + j) >> 3)] = bb[j];
          }
        }
      }
      __syncwarp();
      // raw blocks: a window with rows from two chunks is (0 + partial of the earlier chunk) + partial of the later one
      // (AggrOverTimeFunctions.scala:560-571); own windows that sit in a raw block are finished here as well
      for (int ci = 0; ci < n; ++ci) {
        const WpChunk& ch = CD[ci];
        if (ch.jzb == 0 && ch.tb >= ch.nblk) continue;
        auto rows_in = [&](const WpChunk& x, int k) -> int {
          int lo = x.s0 + k; if (lo < 0) lo = 0; int hi = x.s0 + k + Wr; if (hi > x.nrows - 1) hi = x.nrows - 1;
          return hi - lo + 1;
        };
        // head: blocks [0, jzb)
        const int nh = ch.jzb * WP_R < ch.kT1 - ch.kT0 + 1 ? ch.jzb * WP_R : ch.kT1 - ch.kT0 + 1;
        for (int i = lane; i < nh; i += 32) {
          const int k = ch.kT0 + i;
          double v = J[ch.joff + i]; int nn = 1;
          if (FN == FN_AVG || FN == FN_COUNT) nn = rows_in(ch, k);
          if (i < ch.hs) { v = O[oidx(k)] + v; if (FN == FN_AVG || FN == FN_COUNT) nn += rows_in(CD[ci - 1], k); }
          O[oidx(k)] = wp_finish<FN>(v, nn, fdiv, frcp, 1000.0, p_nfull, p_rcpn, false);
        }
        // tail: own windows of block tb
        for (int k = ch.kT0 + ch.tb * WP_R + lane; k <= ch.ownHi && ch.tb < ch.nblk; k += 32) {
          int nn = 1;
          if (FN == FN_AVG || FN == FN_COUNT) nn = rows_in(ch, k);
          O[oidx(k)] = wp_finish<FN>(O[oidx(k)], nn, fdiv, frcp, 1000.0, p_nfull, p_rcpn, false);
        }
      }
      // windows without rows: NaN (no chunk contributes: AggrOverTimeFunctions.scala:560-571 leaves the NaN seed)
      if (p_gaps) {
        int prev = -1;
        for (int ci = 0; ci <= n; ++ci) {
          int gend = q.T;
          if (ci < n) { if (CD[ci].nblk == 0) continue; gend = CD[ci].kT0; }
          for (int k = prev + 1 + lane; k < gend; k += 32) O[oidx(k)] = NaNv;
          if (ci < n) prev = CD[ci].kT1;
        }
      }
    }
    // ------------------------------------------------------------------------------------------------ result row
    __syncwarp();
    {
      // lane-consecutive windows: 256 contiguous bytes per store instruction; O index of window lane + 32 m = oidx(lane) + 36 m
      double* gp = out + (size_t)s * q.T + lane;
      const double* sp = O + oidx(lane);
      int iters = (q.T - lane + 31) >> 5;
      for (; iters >= 4; iters -= 4, gp += 128, sp += 144) {
        const double v0 = sp[0], v1 = sp[36], v2 = sp[72], v3 = sp[108];
        wp_store_result(gp, v0); wp_store_result(gp + 32, v1); wp_store_result(gp + 64, v2); wp_store_result(gp + 96, v3);
      }
      for (; iters > 0; --iters, gp += 32, sp += 36) wp_store_result(gp, *sp);
    }
    __syncwarp();
  }
  if (lane == 0) {
    if (rows_scanned | bytes_scanned) { atomicAdd(&d_counters[0], (unsigned long long)rows_scanned); atomicAdd(&d_counters[1], (unsigned long long)bytes_scanned); }
  }
}

} // namespace filo
