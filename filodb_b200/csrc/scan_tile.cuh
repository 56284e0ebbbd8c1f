// v3 scan path: CTA-tile kernel for the regular case (the BASELINE workloads).
//
// A CTA processes tiles of TILE_NS consecutive series.  Records of consecutive series are adjacent in the arena, so a tile's
// chunk pages arrive with ONE cp.async.bulk (TMA) into shared memory, and the tile's [TILE_NS x T] results leave with ONE
// cp.async.bulk store.  Between the two, all 256 threads work on uniform work items:
//   setup    warp w resolves series w: chunk range, regularity, single-chunk window intervals (scan_fast.cuh definitions)
//   decode   item = (series, NibblePack group): field extraction + local XOR prefix; a per-series segmented prefix over the
//            group totals; an apply pass over the same items.  Raw f64 vectors are copied.  NaN/Inf presence is recorded.
//   windows  item = (series, block of BLK_R single-chunk windows): register-blocked sequential sums (exact reference order);
//            item = (series, other window): literal per-chunk fold for windows that take rows from two chunks.
// A series is "regular" when every chunk in range has const-DDV timestamps with slope == step and XOR/raw double values,
// with at most TILE_MAXC chunks and TILE_MAXG NibblePack groups; anything else is appended to a fallback list that the
// generic v2 kernel processes afterwards (same output buffer), so the result is always complete and identical.
#pragma once
#include "scan_fast.cuh"
#include "scan_tile_layout.h"

namespace filo {

__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// compile-time specialised finish of one single-chunk window (SumFinish of scan_fast.cuh with FN known)
template <int FN>
__device__ __forceinline__ double tile_finish(double cs, int nn, double div, double rcp) {
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
  if (FN == FN_COUNT) return (double)nn;
  const double sum = nn ? cs : NaNv;
  if (FN == FN_RATE) return __dmul_rn(div_invariant(sum, div, rcp), 1000.0);
  if (FN == FN_AVG) return nn > 0 ? sum / (double)nn : sum;            // AggrOverTimeFunctions.scala:1000
  return sum;                                                          // FN_SUM, FN_INCREASE (delta schema)
}

// literal per-chunk fold for one window of a regular series (TimeRangeFunction family on const-DDV timestamps):
// chunk-set membership ChunkSetInfo.scala:481-510, row range RangeFunction.scala:185-190, fold AggrOverTimeFunctions.scala:560-571.
// Rows advance one per window, so the unclamped row range of window k is [s0 + k, e0 + k] (no search, no division).
template <int FN>
__device__ __forceinline__ double tile_eval_window(const TileSeries& S, const double* vals, const QueryParams& q, double div, int k) {
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
  int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
  const int64_t wEnd = q.start + (int64_t)k * q.step, wStart = wEnd - winDur;
  double sum = NaNv; int cnt = 0; bool anyrows = false;
  for (int c = 0; c < S.n; ++c) {
    const TileChunk& ch = S.c[c];
    if (ch.end_time < wStart) continue;
    if (c > 0 && !(S.c[c - 1].end_time < wEnd)) continue;
    int su = ch.s0 + k; if (su < 0) su = 0;
    int eu = ch.e0 + k; if (eu > ch.nrows - 1) eu = ch.nrows - 1;
    if (su > eu) continue;
    const double* v = vals + ch.row_base;
    double cs = 0.0; int nn = 0;
    for (int r = su; r <= eu; ++r) { const double x = v[r]; if (x == x) { cs += x; ++nn; } }
    anyrows = true;
    const double csn = nn ? cs : NaNv;
    if (nn && sum != sum) sum = 0.0;
    sum += csn; cnt += nn;
  }
  if (FN == FN_RATE) return sum / div * 1000.0;
  if (FN == FN_AVG) return cnt > 0 ? sum / (double)cnt : (sum != sum ? sum : 0.0);
  if (FN == FN_COUNT) return anyrows ? (double)cnt : NaNv;
  return sum;
}

// ---------------------------------------------------------------------------------------------------------------------
// Tile kernel, SUM class (sum/avg/count_over_time, rate/increase on delta schemas), no across-series aggregate.
// ---------------------------------------------------------------------------------------------------------------------
template <int FN>
__global__ void __launch_bounds__(TILE_THREADS, 2)
scan_tile_sum_kernel(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, int64_t n_series,
                     QueryParams q, double* __restrict__ out, TileSmem L,
                     int64_t* __restrict__ fallback_list, unsigned long long* __restrict__ fallback_count,
                     unsigned long long* d_counters, int* d_err) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* recbuf = smem + L.rec;
  double* vals = reinterpret_cast<double*>(smem + L.vals);
  double* otile = reinterpret_cast<double*>(smem + L.out);
  TileSeries* SD = reinterpret_cast<TileSeries*>(smem + L.desc);
  uint64_t* gtot = reinterpret_cast<uint64_t*>(smem + L.gtot);
  TileMeta* M = reinterpret_cast<TileMeta*>(smem + L.meta);
  const int64_t n_tiles = (n_series + TILE_NS - 1) / TILE_NS;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  StepDiv sd; sd.init(q.step);
  int64_t winDur = q.inclusive ? q.window : q.window - 1; if (winDur < 0) winDur = 0;
  const double fdiv = (double)(q.inclusive ? winDur : winDur + 1), frcp = 1.0 / fdiv;     // RateFunctions.scala:436-442
  const int64_t S0 = q.start - winDur, E0 = q.start;
  uint32_t parity = 0;
  int64_t rows_scanned = 0, bytes_scanned = 0;
  const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;

  auto issue_tile = [&](int64_t t) -> bool {          // returns whether the tile is staged through TMA
    const int64_t i0 = t * TILE_NS, i1 = (i0 + TILE_NS < n_series) ? i0 + TILE_NS : n_series;
    const int64_t o = rec_off[i0];
    const uint32_t bytes = (uint32_t)(rec_off[i1] - o);
    if (bytes > L.rec_cap - 64) return false;
    if (tid == 0) { mbar_expect_tx(bar, bytes); tma_load_1d(recbuf, arena + o, bytes, bar); }
    return true;
  };
  int64_t t = blockIdx.x;
  bool staged = (t < n_tiles) ? issue_tile(t) : false;

  for (; t < n_tiles; t += gridDim.x) {
    const int64_t i0 = t * TILE_NS;
    const int ns = (int)((i0 + TILE_NS < n_series ? i0 + TILE_NS : n_series) - i0);
    const int64_t tile_base = rec_off[i0];
    if (staged) { mbar_wait(bar, parity); parity ^= 1; }
    // ------------------------------------------------------------------ setup: warp w <-> series w
    if (warp < TILE_NS) {
      TileSeries& S = SD[warp];
      if (warp >= ns) { if (lane == 0) { S.n = 0; S.regular = 2; S.nblocks = 0; S.nrest = 0; S.ngroups = 0; S.nrows = 0; } }
      else {
        const uint32_t roff = (uint32_t)(rec_off[i0 + warp] - tile_base);
        const uint8_t* rec = staged ? recbuf + roff : arena + tile_base + roff;
        const RecordHeader* h = reinterpret_cast<const RecordHeader*>(rec);
        const ChunkEntry* E = reinterpret_cast<const ChunkEntry*>(rec + sizeof(RecordHeader));
        const int nch = (int)h->n_chunks;
        const int64_t t1 = q.start - q.window, t2 = q.end;
        int cLo = 0; while (cLo < nch && E[cLo].end_time < t1) ++cLo;
        int cHi = cLo; while (cHi < nch && E[cHi].start_time <= t2) ++cHi;
        if (t1 > t2) cHi = cLo;
        const int n = cHi - cLo;
        bool regular = staged && n <= TILE_MAXC && (n == 0 || (h->flags & REC_ALL_TS_CONST));
        int ngroups = 0, nrows_tot = 0; bool any_raw = false;
        if (regular) {
          for (int c = 0; c < n; ++c) {
            const ChunkEntry& e = E[cLo + c];
            const uint8_t* tv = rec + e.ts_off; const uint8_t* vv = rec + e.val_off;
            const int vwire = ld32(vv + 4) & 0xffff;
            const int tlen = (int)ld32(tv + 8); const int64_t init = (int64_t)ld64_a4(tv + 12); const int slope = (int)ld32(tv + 20);
            int vlen, ng = 0;
            if (vwire == WIRE_XOR) { vlen = (int)ld32(vv + XOR_OFF_N); ng = (int)(ld32(vv + XOR_OFF_NGROUPS) & 0xffff); }
            else if (vwire == WIRE_RAW64) { vlen = ((int)ld32(vv) - 4) / 8; any_raw = true; }
            else { regular = false; break; }
            if ((int64_t)slope != q.step || tlen <= 0 || vlen <= 0) { regular = false; break; }
            int nrows = e.num_rows < tlen ? e.num_rows : tlen; if (vlen < nrows) nrows = vlen;
            if (lane == 0) {
              TileChunk& ch = S.c[c];
              ch.init = init; ch.end_time = e.end_time; ch.nrows = nrows; ch.row_base = nrows_tot;
              ch.val_off = roff + e.val_off; ch.wire = vwire; ch.ngroups = ng; ch.grp_base = ngroups; ch.has_nan = 0;
              ch.tlen = tlen; ch.vlen = vlen;
            }
            ngroups += ng; nrows_tot += vlen;
            if (lane == 0) {      // CountingChunkInfoIterator, ChunkSetInfo.scala:336-380 (every chunk in range is pulled when regular)
              rows_scanned += e.num_rows;
              bytes_scanned += (int64_t)ld32(tv) + 4 + (int64_t)ld32(vv) + 4;
            }
          }
          if (ngroups > TILE_MAXG || nrows_tot + 2 > (int)L.vals_pitch) regular = false;
        }
        __syncwarp();
        if (regular) {
          // single-chunk window intervals (same definition as scan_fast.cuh chunk_interval)
          int nblocks = 0, covered = 0;
          const int64_t lastEnd = q.start + (int64_t)(q.T - 1) * q.step;
          for (int c = 0; c < n; ++c) {
            TileChunk& ch = S.c[c];
            int64_t v = 0;
            if (lane == 0) v = sd.ceil_div(ch.init - E0);
            else if (lane == 1) {
              if (c > 0) { const TileChunk& p = S.c[c - 1]; int64_t pm = p.end_time; const int64_t pl = p.init + (int64_t)(p.tlen - 1) * q.step; if (pl > pm) pm = pl; v = sd.ceil_div(pm + 1 - S0); }
            }
            else if (lane == 2) v = sd.floor_div(ch.init + (int64_t)(ch.nrows - 1) * q.step - S0);
            else if (lane == 3) v = (c + 1 < n) ? sd.floor_div(S.c[c + 1].init - 1 - E0) : (int64_t)q.T;
            else if (lane == 4) v = sd.floor_div(ch.end_time - S0);
            else if (lane == 5) v = sd.ceil_div(S0 - ch.init);            // s0: unclamped first row of window 0
            else if (lane == 6) v = sd.floor_div(E0 - ch.init);           // e0: unclamped last row of window 0
            int64_t kA = __shfl_sync(0xffffffffu, v, 0);
            { const int64_t x = __shfl_sync(0xffffffffu, v, 1); if (x > kA) kA = x; }
            int64_t kB = __shfl_sync(0xffffffffu, v, 2);
            { const int64_t x = __shfl_sync(0xffffffffu, v, 3); if (x < kB) kB = x; }
            { const int64_t x = __shfl_sync(0xffffffffu, v, 4); if (x < kB) kB = x; }
            const int64_t s0 = __shfl_sync(0xffffffffu, v, 5), e0 = __shfl_sync(0xffffffffu, v, 6);
            if (kA < 0) kA = 0;
            if (kB > q.T - 1) kB = q.T - 1;
            const int64_t sA = s0 + kA, eA = e0 + kA;
            const bool ok = kA <= kB && eA >= sA;
            const int Wr = ok ? (int)(eA - sA) : 0;
            const int nwin = ok ? (int)(kB - kA + 1) : 0;
            // blocked only when the windows are long enough to amortise a block; short windows go through the per-window path
            const bool blocked = ok && Wr >= BLK_R - 1;
            const int nb = blocked ? (nwin + BLK_R - 1) / BLK_R : 0;
            if (lane == 0) {
              ch.kA = blocked ? (int)kA : 0; ch.kB = blocked ? (int)kB : -1; ch.sA = (int)sA; ch.Wr = Wr; ch.blk0 = nblocks; ch.blk_n = nb;
              ch.s0 = (int)s0; ch.e0 = (int)e0;
            }
            nblocks += nb; if (blocked) covered += nwin;
            // a chunk that the window iterator never pulls (it starts after the last window end) is not counted as scanned
            if (lane == 0 && c > 0 && !(S.c[c - 1].end_time < lastEnd)) {
              const ChunkEntry& e = E[cLo + c];
              rows_scanned -= e.num_rows; bytes_scanned -= (int64_t)ld32(rec + e.ts_off) + 4 + (int64_t)ld32(rec + e.val_off) + 4;
            }
          }
          if (lane == 0) { S.n = n; S.regular = 1; S.rec_off = (int)roff; S.nblocks = nblocks; S.nrest = q.T - covered; S.ngroups = ngroups; S.nrows = nrows_tot; S.pad = any_raw; }
        } else if (lane == 0) {
          S.n = 0; S.regular = 0; S.nblocks = 0; S.nrest = 0; S.ngroups = 0; S.nrows = 0; S.pad = 0;
          const unsigned long long slot = atomicAdd(fallback_count, 1ull);
          fallback_list[slot] = i0 + warp;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {                              // tile work-list prefixes (read after the next barriers)
      int p = 0, r = 0, raw = 0, allreg = 1;
      for (int s = 0; s < TILE_NS; ++s) {
        M->pref[s] = p; M->rpref[s] = r;
        if (SD[s].regular == 1) { p += SD[s].nblocks; r += SD[s].nrest; raw |= SD[s].pad; }
        if (s < ns && SD[s].regular != 1) allreg = 0;
      }
      M->pref[TILE_NS] = p; M->rpref[TILE_NS] = r; M->any_nan = 0; M->any_raw = raw; M->all_regular = allreg;
    }
    // ------------------------------------------------------------------ decode pass 1: fields + local prefix, group totals
    // item = (group slot, series) with the series index fastest: neighbouring lanes write to different series' rows, which
    // spreads the 8-value stores over the shared-memory banks
#pragma unroll 2
    for (int it = tid; it < TILE_NS * TILE_MAXG; it += TILE_THREADS) {
      const int s = it & (TILE_NS - 1), slot = it >> 3;
      const TileSeries& S = SD[s];
      if (S.regular != 1 || slot >= S.ngroups) continue;
      int c = 0; while (c + 1 < S.n && slot >= S.c[c + 1].grp_base) ++c;
      const TileChunk& ch = S.c[c];
      const int g = slot - ch.grp_base;
      const uint8_t* v = recbuf + ch.val_off;
      const uint32_t w12 = ld32(v + XOR_OFF_NGROUPS);
      const uint8_t* groups = v + (w12 >> 16) + 8;
      const uint16_t* tab = reinterpret_cast<const uint16_t*>(v + XOR_OFF_GROUPTAB);
      const uint8_t* gp = groups + tab[g];
      uint64_t d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = 0;
      const uint32_t mask = gp[0];
      if (mask != 0) {
        const uint32_t hdr = gp[1];
        const int numBits = ((hdr >> 4) + 1) * 4;
        const int tz = (hdr & 0x0f) * 4;
        const uint64_t fmask = numBits >= 64 ? ~0ull : ((1ull << numBits) - 1);
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(gp + 2);
        uint32_t bit = (uint32_t)(a0 & 7) * 8;
        const uint8_t* base = reinterpret_cast<const uint8_t*>(a0 & ~(uintptr_t)7);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (mask & (1u << i)) {
            const uint8_t* wp = base + ((bit >> 6) << 3);
            const int off = (int)(bit & 63);
            uint64_t val = ld64(wp) >> off;
            if (off + numBits > 64) val |= ld64(wp + 8) << (64 - off);
            d[i] = (val & fmask) << tz;
            bit += numBits;
          }
        }
      }
      uint64_t x = 0;
      uint64_t* dst = reinterpret_cast<uint64_t*>(vals + (size_t)s * L.vals_pitch + ch.row_base) + 1 + g * 8;
      const int nleft = ch.vlen - 1 - g * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) { x ^= d[i]; if (i < nleft) dst[i] = x; }
      gtot[s * TILE_MAXG + slot] = x;
    }
    __syncthreads();
    // raw f64 vectors: plain copy (+ NaN/Inf presence)
    if (M->any_raw) {
      for (int s = 0; s < TILE_NS; ++s) {
        const TileSeries& S = SD[s];
        if (S.regular != 1) continue;
        for (int c = 0; c < S.n; ++c) {
          const TileChunk& ch = S.c[c];
          if (ch.wire != WIRE_RAW64) continue;
          const uint64_t* src = reinterpret_cast<const uint64_t*>(recbuf + ch.val_off + 8);
          uint64_t* dst = reinterpret_cast<uint64_t*>(vals + (size_t)s * L.vals_pitch + ch.row_base);
          bool nan = false;
          for (int r = tid; r < ch.vlen; r += TILE_THREADS) { const uint64_t b = src[r]; dst[r] = b; nan |= ((uint32_t)(b >> 32) & 0x7ff00000u) == 0x7ff00000u; }
          if (nan) M->any_nan = 1;
        }
      }
    }
    // ------------------------------------------------------------------ decode pass 2: per-series segmented prefix of group totals
    if (warp < TILE_NS && SD[warp].regular == 1) {
      const TileSeries& S = SD[warp];
      for (int c = 0; c < S.n; ++c) {
        const TileChunk& ch = S.c[c];
        if (ch.wire != WIRE_XOR) continue;
        const uint8_t* v = recbuf + ch.val_off;
        uint64_t carry = ld64(v + (ld32(v + XOR_OFF_NGROUPS) >> 16));        // first value of the chunk
        if (lane == 0) {
          reinterpret_cast<uint64_t*>(vals + (size_t)warp * L.vals_pitch + ch.row_base)[0] = carry;
          if (((uint32_t)(carry >> 32) & 0x7ff00000u) == 0x7ff00000u) M->any_nan = 1;
        }
        for (int g0 = 0; g0 < ch.ngroups; g0 += 32) {
          const int g = g0 + lane;
          const uint64_t x = g < ch.ngroups ? gtot[warp * TILE_MAXG + ch.grp_base + g] : 0;
          uint64_t incl = x;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) { const uint64_t y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl ^= y; }
          if (g < ch.ngroups) gtot[warp * TILE_MAXG + ch.grp_base + g] = carry ^ incl ^ x;   // value before the group
          carry ^= __shfl_sync(0xffffffffu, incl, 31);
        }
      }
    }
    __syncthreads();
    // the record bytes are dead now: prefetch the next tile into the staging buffer while this one is reduced
    const int64_t tnext = t + gridDim.x;
    bool staged_next = false;
    if (tnext < n_tiles) staged_next = issue_tile(tnext);
    // ------------------------------------------------------------------ decode pass 3: apply the prefix (same items as pass 1)
#pragma unroll 2
    for (int it = tid; it < TILE_NS * TILE_MAXG; it += TILE_THREADS) {
      const int s = it & (TILE_NS - 1), slot = it >> 3;
      const TileSeries& S = SD[s];
      if (S.regular != 1 || slot >= S.ngroups) continue;
      int c = 0; while (c + 1 < S.n && slot >= S.c[c + 1].grp_base) ++c;
      const TileChunk& ch = S.c[c];
      const int g = slot - ch.grp_base;
      const uint64_t pre = gtot[s * TILE_MAXG + slot];
      uint64_t* dst = reinterpret_cast<uint64_t*>(vals + (size_t)s * L.vals_pitch + ch.row_base) + 1 + g * 8;
      const int nleft = ch.vlen - 1 - g * 8;
      uint32_t hi_or = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < nleft) {
        const uint64_t b = dst[i] ^ pre; dst[i] = b;
        hi_or |= (((uint32_t)(b >> 32) & 0x7ff00000u) == 0x7ff00000u) ? 1u : 0u;      // NaN or Inf: conservative
      }
      if (hi_or) M->any_nan = 1;
    }
    if (tid == 0) tma_store_wait_read();       // the previous tile's bulk store must have finished reading `otile`
    __syncthreads();
    // ------------------------------------------------------------------ windows: blocked single-chunk windows
    {
      const bool any_nan = M->any_nan != 0;
      const int nitems = M->pref[TILE_NS];
      for (int it = tid; it < nitems; it += TILE_THREADS) {
        int s = 0;
#pragma unroll
        for (int j = 1; j < TILE_NS; ++j) if (it >= M->pref[j]) s = j;
        const TileSeries& S = SD[s];
        const int B = it - M->pref[s];
        int c = 0; while (c + 1 < S.n && B >= S.c[c].blk0 + S.c[c].blk_n) ++c;
        const TileChunk& ch = S.c[c];
        const int b = B - ch.blk0;
        const int r0 = ch.sA + b * BLK_R;
        const double* slots = vals + (size_t)s * L.vals_pitch + ch.row_base;
        double acc[BLK_R]; int cnt[BLK_R];
        if (any_nan) blocked_sum<true>(slots, r0, ch.nrows, ch.Wr, acc, cnt); else blocked_sum<false>(slots, r0, ch.nrows, ch.Wr, acc, cnt);
        const int k0 = ch.kA + b * BLK_R;
        int nw = ch.kB - k0 + 1; if (nw > BLK_R) nw = BLK_R;
        double* o = otile + (size_t)s * L.out_pitch + k0;
#pragma unroll
        for (int j = 0; j < BLK_R; ++j) if (j < nw) o[j] = tile_finish<FN>(acc[j], cnt[j], fdiv, frcp);
      }
      // ---------------------------------------------------------------- windows: everything else (chunk junctions, short windows)
      const int nrest = M->rpref[TILE_NS];
      for (int it = tid; it < nrest; it += TILE_THREADS) {
        int s = 0;
#pragma unroll
        for (int j = 1; j < TILE_NS; ++j) if (it >= M->rpref[j]) s = j;
        const TileSeries& S = SD[s];
        int u = it - M->rpref[s];
        int prev = -1; bool found = false;          // u-th window not covered by a blocked interval
        for (int c = 0; c < S.n && !found; ++c) {
          if (S.c[c].kA > S.c[c].kB) continue;
          const int gap = S.c[c].kA - prev - 1;
          if (u < gap) found = true; else { u -= gap; prev = S.c[c].kB; }
        }
        const int k = prev + 1 + u;
        otile[(size_t)s * L.out_pitch + k] = tile_eval_window<FN>(S, vals + (size_t)s * L.vals_pitch, q, fdiv, k);
      }
    }
    fence_async_smem();        // make this thread's writes to the output tile visible to the async proxy (bulk store below)
    __syncthreads();
    // ------------------------------------------------------------------ results: one bulk store for the tile (regular rows only)
    {
      double* gout = out + (size_t)i0 * q.T;
      const uint32_t bytes = (uint32_t)ns * (uint32_t)q.T * 8u;
      if (M->all_regular && out_aligned && (bytes & 15) == 0 && (((size_t)i0 * q.T * 8) & 15) == 0) {
        if (tid == 0) tma_store_1d(gout, otile, bytes);
      } else {
        for (int s = 0; s < ns; ++s) {
          if (SD[s].regular != 1) continue;
          for (int k = tid; k < q.T; k += TILE_THREADS) gout[(size_t)s * q.T + k] = otile[(size_t)s * L.out_pitch + k];
        }
        __syncthreads();
      }
    }
    staged = staged_next;
  }
  if (tid == 0) tma_store_wait_read();
  if (lane == 0 && (rows_scanned | bytes_scanned)) {
    atomicAdd(&d_counters[0], (unsigned long long)rows_scanned); atomicAdd(&d_counters[1], (unsigned long long)bytes_scanned);
  }
}

} // namespace filo
