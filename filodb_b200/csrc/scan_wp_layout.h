// Shared-memory layout of the v4 warp-pipeline kernel (scan_wp.cuh); plain structs, usable from host code.
#pragma once
#include <stdint.h>
#include "filo_record.h"
#include "scan_params.h"
#include "scan_tile_layout.h"
namespace filo {
constexpr int WP_MAXC = 4;             // chunks in range per series on this path
constexpr int WP_MAXG = 64;            // NibblePack groups per series (two per lane)
constexpr int WP_R = 8;                // windows per block
constexpr int WP_MAX_WARPS = 16;       // warps per CTA (one CTA per SM) when O has its own region
#ifndef FILO_WP_MAX_WARPS_ALIAS
#define FILO_WP_MAX_WARPS_ALIAS 20
#endif
constexpr int WP_MAX_WARPS_ALIAS = FILO_WP_MAX_WARPS_ALIAS; // ... when O takes V's place (single-pass plans): 20 warps = 96 registers per thread (22 = 88 registers, spills: A/B in profiles/r2)

struct WpChunk {                       // per warp, per chunk in range (shared memory)
  uint64_t first;                      // XOR vectors: bits of the first value
  int32_t kT0, kT1;                    // windows whose (unclamped) row range meets the chunk's rows, clipped to [0, T)
  int32_t ownLo, ownHi;                // ... of which only this chunk contributes to [ownLo, ownHi]
  int32_t blk0, nblk;                  // block items of this chunk: [blk0, blk0 + nblk)
  int32_t vidx0;                       // V index of the first row of block 0
  int32_t rowpos;                      // V position (before skewing) of row 0
  int32_t nrows, s0, e0;               // rows; unclamped first / last row of window 0
  int32_t grp_base, ng, wire;          // group slots [grp_base, grp_base + ng)
  uint32_t grp_off, tab_off, val_off;  // byte offsets in R: first group, u16 group table, value vector
  int32_t joff, hs;                    // head share: windows [kT0, ownLo) also take rows from the previous chunk
  int32_t jzb, tb;                     // blocks [0, jzb) (the head share rounded up to whole blocks) leave their raw sums at J[joff ..]; blocks [tb, nblk)
                                       // (from the block that holds ownHi + 1) leave raw sums in O; the blocks in between hold own windows only
  int32_t dropped;                     // counter drop flag of the value vector (PrimitiveVectorReader.dropped, BinaryVector.scala:530-531)
};
static_assert(sizeof(WpChunk) == 96, "WpChunk");

// fixed part of a warp's region: mbarrier, descriptors, J -- compile-time offsets keep the kernel's address arithmetic in immediates
constexpr uint32_t WP_OFF_DESC = 16, WP_OFF_J = WP_OFF_DESC + WP_MAXC * 96, WP_J_DOUBLES = 128, WP_OFF_REC = WP_OFF_J + WP_J_DOUBLES * 8;
struct WpSmem {                        // byte offsets inside a warp's region, all multiples of 16
  uint32_t desc, jbuf, rec, vals, out, per_warp;
  uint32_t rec_cap, vcap /*doubles*/, jcap /*doubles*/, ocap /*doubles*/;
  uint32_t warps;                      // warps per CTA
  uint32_t alias;                      // O lives in V's region: every block of a series is summed (one pass of <= 64 blocks) before the first result is stored
};
// wrows = window / step + 1: the most rows a window can span
FILO_HD inline WpSmem wp_layout(uint32_t max_rec_bytes, uint32_t max_rows, uint32_t max_chunks, uint32_t T, uint32_t wrows, bool alias) {
  WpSmem L;
  if (max_chunks > (uint32_t)WP_MAXC) max_chunks = WP_MAXC;
  L.rec_cap = align_up(max_rec_bytes + 16, 16);
  const uint32_t P = max_rows + (max_chunks + 1) * (wrows + 7) + 16;         // positions: rows + zero gaps + slack
  L.vcap = align_up(P + P / 8 + 2, 2);
  L.jcap = WP_J_DOUBLES;               // raw sums of the windows two chunks share (a plan that needs more is declined); also the XOR prefix table of the decode (64 words)
  (void)wrows;
  L.ocap = align_up(T + T / 8 + 4, 2);                                       // skewed like V: one pad slot per 8 windows
  if (alias && L.vcap < L.ocap) L.vcap = L.ocap;
  static_assert(sizeof(WpChunk) == 96, "WP_OFF_J");
  uint32_t o = WP_OFF_REC;
  L.desc = WP_OFF_DESC; L.jbuf = WP_OFF_J;
  L.rec = o; o += L.rec_cap;
  L.vals = o; o += L.vcap * 8;
  if (alias) L.out = L.vals; else { L.out = o; o += L.ocap * 8; }
  L.per_warp = align_up(o, 16);
  L.warps = 0;
  L.alias = alias ? 1u : 0u;
  return L;
}
// an upper bound of the blocks of a series: sum over chunks of ceil(touched windows / 8), with at most wrows - 1 windows shared per junction
FILO_HD inline uint32_t wp_max_items(uint32_t max_chunks, uint32_t T, uint32_t wrows) {
  if (max_chunks > (uint32_t)WP_MAXC) max_chunks = WP_MAXC;
  if (max_chunks == 0) max_chunks = 1;
  return (T + (max_chunks - 1) * (wrows - 1) + 7 * max_chunks) / 8;
}

constexpr int WP_CTR_MAX_WARPS = 20;   // counter kernel: warps per CTA at the lower register budget
struct WpCtrChunk {                    // per warp, per chunk in range: plan (window intervals, extrapolation constants)
  int64_t init, end_time;
  int32_t nrows, s0, e0, rowpos;
  int32_t kA, kB;                      // unclamped single-chunk windows with at least two samples ([0, -1] if none)
  int32_t kA2, kB2;                    // every single-chunk window of the chunk, clamped ones included
  TileCtr kc;                          // per-series: kc.dropped
};
static_assert(sizeof(WpCtrChunk) == 112, "WpCtrChunk");

struct WpCtrSmem {                     // byte offsets inside a warp's region (multiples of 16); R sits at WP_OFF_REC, xtab at WP_OFF_J, drops behind it
  uint32_t vals, kc, acc, nbad, per_warp, tab /* per CTA, behind the warps' regions */;
  uint32_t rec_cap, vcap /*doubles*/, warps, agg;
  uint32_t tsr;                        // irregular timestamps: int32 row times relative to the chunk's first (0: the table has const-DDV timestamps only)
};
constexpr uint32_t WP_OFF_DROPS = WP_OFF_J + 64 * 8;          // TileDrops[WP_MAXC] behind the 64-word XOR prefix table
static_assert(WP_OFF_DROPS + WP_MAXC * sizeof(TileDrops) <= WP_OFF_REC, "drops fit in front of the record");
FILO_HD inline WpCtrSmem wp_ctr_layout(uint32_t max_rec_bytes, uint32_t max_rows, uint32_t max_chunks, uint32_t T, bool agg, bool irr = false) {
  WpCtrSmem L;
  if (max_chunks > (uint32_t)WP_MAXC) max_chunks = WP_MAXC;
  L.rec_cap = align_up(max_rec_bytes + 16, 16);
  const uint32_t P = max_rows + 8 * max_chunks + 8;            // 8 spare rows behind every chunk: the group decode runs up to 7 rows past it
  L.vcap = align_up(P + P / 8 + 2, 2);
  uint32_t o = WP_OFF_REC + L.rec_cap;
  L.vals = o; o += L.vcap * 8;
  L.kc = o; o += (uint32_t)(WP_MAXC * sizeof(WpCtrChunk));
  L.tsr = 0;
  if (irr) { L.tsr = o; o += align_up(P * 4, 16); }
  L.acc = o; L.nbad = o;
  if (agg) { o += T * 8; L.nbad = o; o += align_up(T * 2, 16); }
  L.per_warp = align_up(o, 16);
  L.tab = 0; L.warps = 0; L.agg = agg ? 1u : 0u;
  return L;
}

} // namespace filo
