// GPU-side synthetic chunk generator + encoder (bench/test data producer; SURVEY.md §8(f)3 "GPU-side encode").
// Restates, per series and per chunk, what FiloDB's ingestion would have produced for the reference's own test generator:
//   values     gateway/src/main/scala/filodb/timeseries/TestTimeseriesProducer.scala:147 (gauge), :196-198 (counter)
//   timestamps TimestampAppendingVector.optimize  -> DeltaDeltaVector.fromLongVector(approxConst = true)
//              core/src/main/scala/filodb.memory/format/vectors/LongBinaryVector.scala:333-340, DeltaDeltaVector.scala:63-135
//   doubles    DoubleVector.optimize (DDV-as-long if all integral, else raw f64) + DoubleCounterAppender drop flag
//              vectors/DoubleVector.scala:86-96, 456-473
//   XOR        NibblePack.packDoubles / pack8 / packUniversal, NibblePack.scala:73-183 (payload of this repo's container)
// The generator is deterministic (hash of seed, series id, row), so tests rebuild the same rows on the CPU, encode them with
// the oracle's restatement of the appenders and compare the arena bytes.  One thread per series; this is setup code, not
// the measured hot path.
#include "../../include/filo_b200.h"
#include "kernels.h"
#include <cub/cub.cuh>
#include <string>
#include <vector>

struct filo_ctx; struct filo_table;
filo_table* filo_internal_new_table();
void filo_internal_set_arena(filo_table* t, uint8_t* d_arena, int64_t* d_rec_off, int64_t n_series, int64_t n_chunks, int64_t n_samples,
                             int64_t arena_bytes, int64_t algorithmic_bytes, int32_t max_rows, int32_t max_chunks, int32_t schema_flags);
int32_t filo_internal_finish_table(filo_ctx* ctx, filo_table* t, const int32_t* d_group_ids, int32_t n_groups);
void filo_internal_set_layout(filo_table* t, uint32_t max_rec_bytes, bool any_nonconst_ts, bool any_drop);
cudaStream_t filo_internal_stream(filo_ctx* ctx);
int filo_internal_device(filo_ctx* ctx);
int32_t filo_internal_fail(filo_ctx* ctx, int32_t code, const char* msg);
int32_t filo_internal_set_hist(filo_ctx* ctx, filo_table* t, const uint8_t* hist_vector_header);

namespace filo {

struct SynthParams {
  int64_t n_series; int32_t rows, rows_per_chunk; int64_t t0; int32_t interval, jitter;
  int32_t value_kind, value_enc, reset_period, nan_ppm, n_groups, cumulative;
  uint64_t seed; int64_t gid_base; const double* sin_table; double noise_scale;
  // external samples (filo_encode_table): row-major [n_series][rows]; when set, `key` is the series ordinal and nothing is generated
  const int64_t* ext_ts = nullptr; const double* ext_vals = nullptr;
};

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t series_key(uint64_t seed, uint64_t gid) { return splitmix64(seed ^ (gid * 0xD1342543DE82EF95ull)); }
__device__ __forceinline__ uint64_t row_hash(uint64_t key, int row, int salt) { return splitmix64(key + ((uint64_t)(uint32_t)row << 3) + (uint64_t)salt); }

struct Gen { double v; };     // counter running value

// gauge-like sample: (15 + sin(n+1)) + noise, explicit rn ops so that no FMA contraction can change the bits
__device__ __forceinline__ double gauge_sample(const SynthParams& P, uint64_t key, int row) {
  const uint64_t h = row_hash(key, row, 0);
  const int x = (int)(h & 0xffff) + (int)((h >> 16) & 0xffff) + (int)((h >> 32) & 0xffff) + (int)(h >> 48);
  const double noise = __dmul_rn((double)(x - 131070), P.noise_scale);
  return __dadd_rn(__dadd_rn(15.0, P.sin_table[row]), noise);
}
__device__ __forceinline__ double gen_value(const SynthParams& P, uint64_t key, int row, bool last_in_chunk, Gen& g) {
  if (P.ext_vals) return P.ext_vals[(size_t)key * (size_t)P.rows + (size_t)row];
  if (last_in_chunk && P.nan_ppm > 0 && (int)(row_hash(key, row, 1) % 1000000ull) < P.nan_ppm)
    return __longlong_as_double(0x7ff8000000000000LL);
  const double s = gauge_sample(P, key, row);
  if (P.value_kind == 0) return s;
  double inc = s > 0.0 ? s : 0.0;
  if (P.value_kind == 2) inc = rint(inc);
  if (P.reset_period > 0 && row > 0 && (row_hash(key, row, 2) % (uint64_t)P.reset_period) == 0) g.v = inc;
  else g.v = __dadd_rn(g.v, inc);
  return g.v;
}
__device__ __forceinline__ int64_t gen_ts(const SynthParams& P, uint64_t key, int row) {
  if (P.ext_ts) return P.ext_ts[(size_t)key * (size_t)P.rows + (size_t)row];
  int64_t t = P.t0 + (int64_t)row * P.interval;
  if (P.jitter > 0) t += (int64_t)(row_hash(key, row, 3) % (uint64_t)(2 * P.jitter + 1)) - P.jitter;
  return t;
}

// ---- sequential little-endian bit writer with aligned 8-byte stores
struct Writer {
  uint64_t* p; uint64_t acc; int nb; uint64_t* base;
  __device__ void init(uint8_t* dst) { p = base = reinterpret_cast<uint64_t*>(dst); acc = 0; nb = 0; }
  __device__ __forceinline__ void put(uint64_t v, int bits) {        // bits in [1,64]; v has no bits above `bits`
    acc |= nb ? (v << nb) : v;
    int tot = nb + bits;
    if (tot >= 64) { *p++ = acc; acc = nb ? (v >> (64 - nb)) : 0; tot -= 64; }
    nb = tot;
  }
  __device__ __forceinline__ void pad8() { if (nb) { *p++ = acc; acc = 0; nb = 0; } }          // to 8-byte boundary
  __device__ __forceinline__ void pad_byte() { int r = nb & 7; if (r) put(0, 8 - r); }
  __device__ __forceinline__ uint32_t pos() const { return (uint32_t)((p - base) * 8 + (nb >> 3)); }
};

// IntBinaryVector.minMaxToNbitsSigned, IntBinaryVector.scala:161-177
__device__ __forceinline__ void minmax_to_nbits(int32_t mn, int32_t mx, int& nbits, bool& sgn) {
  if (mn >= 0 && mx < 4) { nbits = 2; sgn = false; }
  else if (mn >= 0 && mx < 16) { nbits = 4; sgn = false; }
  else if (mn >= -128 && mx <= 127) { nbits = 8; sgn = true; }
  else if (mn >= 0 && mx < 256) { nbits = 8; sgn = false; }
  else if (mn >= -32768 && mx <= 32767) { nbits = 16; sgn = true; }
  else if (mn >= 0 && mx < 65536) { nbits = 16; sgn = false; }
  else { nbits = 32; sgn = true; }
}

// Plan of a long vector per DeltaDeltaVector.fromLongVector: kind 0 raw, 1 const, 2 ddv
struct LongPlan { int kind; int64_t first; int32_t slope; int nbits; bool sgn; uint32_t total; };

template <class ValueAt>      // ValueAt(i) -> int64 value of element i (replayable)
__device__ __forceinline__ LongPlan plan_longs(int n, bool approxConst, ValueAt at_seq) {
  LongPlan pl; pl.kind = 0; pl.total = 8 + 8 * (uint32_t)n; pl.first = 0; pl.slope = 0; pl.nbits = 0; pl.sgn = false;
  if (n <= 2) return pl;
  int64_t first = 0, last = 0;
  at_seq([&](int i, int64_t v) { if (i == 0) first = v; last = v; });
  const int64_t slopeL = (last - first) / (int64_t)(n - 1);
  if (!(slopeL < 2147483647LL && slopeL > -2147483648LL)) return pl;
  const int32_t slope = (int32_t)slopeL;
  int32_t mx = INT32_MIN, mn = INT32_MAX; bool ok = true;
  int64_t base = first;
  at_seq([&](int i, int64_t v) {
    if (i == 0) return;
    base += slope;
    const int64_t d = v - base;
    if (d > 2147483647LL || d < -2147483648LL) ok = false;
    else { mx = max(mx, (int32_t)d); mn = min(mn, (int32_t)d); }
  });
  if (!ok) return pl;
  pl.first = first; pl.slope = slope;
  minmax_to_nbits(mn, mx, pl.nbits, pl.sgn);
  if ((mn == 0 && mx == 0) || (approxConst && mn >= -250 && mx <= 250)) { pl.kind = 1; pl.total = 24; return pl; }
  pl.kind = 2;
  pl.total = 20 + 8 + ((uint32_t)n * pl.nbits + 7) / 8;
  return pl;
}

template <class ValueAt>
__device__ __forceinline__ void write_longs(Writer& w, const LongPlan& pl, int n, bool drop, bool is_double_src, ValueAt at_seq) {
  const uint32_t dropbit = drop ? 0x80000000u : 0u;
  if (pl.kind == 1) {                                   // DeltaDeltaVector.const, DeltaDeltaVector.scala:98-106
    w.put(20, 32); w.put((uint32_t)WIRE_DDV_CONST | dropbit, 32); w.put((uint32_t)n, 32);
    w.put((uint64_t)pl.first, 64); w.put((uint32_t)pl.slope, 32);
  } else if (pl.kind == 2) {                            // DeltaDeltaAppendingVector, :293-340
    const uint32_t data = ((uint32_t)n * pl.nbits + 7) / 8;
    w.put(pl.total - 4, 32); w.put((uint32_t)WIRE_DDV | dropbit, 32); w.put((uint64_t)pl.first, 64); w.put((uint32_t)pl.slope, 32);
    const uint32_t bitshift = pl.nbits < 8 ? (((uint32_t)n * pl.nbits) % 8) : 0;
    w.put(4 + data, 32);
    w.put(0x0806u | ((uint32_t)((pl.nbits & 0x7f) | (pl.sgn ? 0x80 : 0)) << 16) | (bitshift << 24), 32);
    int64_t expected = pl.first;
    const uint64_t m = pl.nbits >= 32 ? 0xffffffffull : ((1ull << pl.nbits) - 1);
    at_seq([&](int i, int64_t v) { w.put(((uint64_t)(int64_t)(int32_t)(v - expected)) & m, pl.nbits); expected += pl.slope; });
    w.pad_byte();
  } else {                                              // raw 64-bit (Long/DoubleAppendingVector frozen)
    w.put(4 + 8 * (uint32_t)n, 32); w.put((uint32_t)WIRE_RAW64 | (0xC0u << 16) | dropbit, 32);
    at_seq([&](int i, int64_t v) { w.put((uint64_t)v, 64); });
  }
}

// NibblePack.pack8 size / emit
__device__ __forceinline__ uint32_t pack8_size(const uint64_t in[8]) {
  uint32_t mask = 0; uint64_t orv = 0; int mintz = 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) if (in[i]) { mask |= 1u << i; orv |= in[i]; mintz = min(mintz, __ffsll((long long)in[i]) - 1); }
  if (!mask) return 1;
  const int lz = __clzll((long long)orv);
  const int trailing = mintz / 4, numNibbles = 16 - lz / 4 - trailing;
  return 2 + ((uint32_t)numNibbles * 4 * __popc(mask) + 7) / 8;
}
__device__ __forceinline__ void pack8_emit(Writer& w, const uint64_t in[8]) {
  uint32_t mask = 0; uint64_t orv = 0; int mintz = 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) if (in[i]) { mask |= 1u << i; orv |= in[i]; mintz = min(mintz, __ffsll((long long)in[i]) - 1); }
  w.put(mask, 8);
  if (!mask) return;
  const int lz = __clzll((long long)orv);
  const int trailing = mintz / 4, numNibbles = 16 - lz / 4 - trailing, numBits = numNibbles * 4;
  w.put((uint32_t)(((numNibbles - 1) << 4) | trailing), 8);
  const uint64_t m = numBits >= 64 ? ~0ull : ((1ull << numBits) - 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) if (in[i]) w.put((in[i] >> (trailing * 4)) & m, numBits);
  w.pad_byte();
}

struct ChunkPlan { uint32_t ts_total, val_total; int val_kind; /*0 raw,1 xor,2 longwrap*/ LongPlan tsp, vlp; bool drop; int64_t start_t, end_t; };

// Plans one chunk (sizes + encodings).  g is the generator state at chunk start (restored by the caller for replays).
__device__ void plan_chunk(const SynthParams& P, uint64_t key, int r0, int n, const Gen& g0, ChunkPlan& cp) {
  auto ts_seq = [&](auto&& f) { for (int i = 0; i < n; ++i) f(i, gen_ts(P, key, r0 + i)); };
  cp.tsp = plan_longs(n, true, ts_seq);
  cp.ts_total = cp.tsp.total;
  cp.start_t = gen_ts(P, key, r0); cp.end_t = gen_ts(P, key, r0 + n - 1);
  // drop flag (DoubleCounterAppender.addData, DoubleVector.scala:456-466) + integrality (LongDoubleWrapper :521-533)
  bool drop = false, integral = true; double last = -1.7976931348623157e308;
  { Gen g = g0;
    for (int i = 0; i < n; ++i) {
      const double v = gen_value(P, key, r0 + i, i == n - 1, g);
      if (v != v || v < last) drop = true;
      if (v == v) last = v;
      if (v > 9.2233720368547758e18 || rint(v) != v) integral = false;
    } }
  cp.drop = drop && P.cumulative;
  cp.val_kind = 0; cp.val_total = 8 + 8 * (uint32_t)n;
  if (P.value_enc == 1) {
    cp.val_kind = 1;
    const int ng = (n - 1 + 7) / 8;
    const uint32_t payloadOff = align_up(16 + 2 * (uint32_t)ng, 8);
    uint32_t payload = 8; uint64_t arr[8]; uint64_t lastb = 0;
    Gen g = g0;
    for (int i = 0; i < n; ++i) {
      const uint64_t b = (uint64_t)__double_as_longlong(gen_value(P, key, r0 + i, i == n - 1, g));
      if (i > 0) { arr[(i - 1) & 7] = b ^ lastb; if (((i - 1) & 7) == 7) payload += pack8_size(arr); }
      lastb = b;
    }
    if ((n - 1) & 7) { for (int j = (n - 1) & 7; j < 8; ++j) arr[j] = 0; payload += pack8_size(arr); }
    cp.val_total = align_up(payloadOff + payload, 8);
  } else if (P.value_enc == 2 && integral) {
    auto v_seq = [&](auto&& f) { Gen g = g0; for (int i = 0; i < n; ++i) f(i, (int64_t)gen_value(P, key, r0 + i, i == n - 1, g)); };
    cp.vlp = plan_longs(n, false, v_seq);
    if (cp.vlp.kind != 0) { cp.val_kind = 2; cp.val_total = cp.vlp.total; }
  }
}

__global__ void synth_size_kernel(SynthParams P, uint32_t* rec_bytes, int32_t* group_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_series) return;
  const uint64_t gid = (uint64_t)(P.gid_base + i);
  const uint64_t key = P.ext_vals ? (uint64_t)i : series_key(P.seed, gid);
  const int nch = (P.rows + P.rows_per_chunk - 1) / P.rows_per_chunk;
  uint32_t bytes = sizeof(RecordHeader) + (uint32_t)nch * sizeof(ChunkEntry);
  Gen g; g.v = 0.0;
  for (int c = 0; c < nch; ++c) {
    const int r0 = c * P.rows_per_chunk, n = min(P.rows_per_chunk, P.rows - r0);
    ChunkPlan cp; plan_chunk(P, key, r0, n, g, cp);
    bytes += align_up(cp.ts_total, 8) + align_up(cp.val_total, 8);
    for (int k = 0; k < n; ++k) gen_value(P, key, r0 + k, k == n - 1, g);      // advance state
  }
  rec_bytes[i] = align_up(bytes, 16);
  if (group_ids) group_ids[i] = P.n_groups > 0 ? (int32_t)(splitmix64(P.seed ^ 0xA5A5A5A5ull ^ (gid * 0x9E3779B97F4A7C15ull)) % (uint64_t)P.n_groups) : 0;
}

__global__ void synth_fill_kernel(SynthParams P, const int64_t* rec_off, uint8_t* arena, unsigned long long* alg_bytes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_series) return;
  const uint64_t gid = (uint64_t)(P.gid_base + i);
  const uint64_t key = P.ext_vals ? (uint64_t)i : series_key(P.seed, gid);
  const int nch = (P.rows + P.rows_per_chunk - 1) / P.rows_per_chunk;
  uint8_t* rec = arena + rec_off[i];
  const uint32_t rec_bytes = (uint32_t)(rec_off[i + 1] - rec_off[i]);
  Writer w; w.init(rec);
  // header + entries: plan all chunks first (offsets), then emit vectors
  uint32_t flags = REC_ALL_TS_CONST; uint32_t off = sizeof(RecordHeader) + (uint32_t)nch * sizeof(ChunkEntry);
  w.put(rec_bytes, 32); w.put((uint32_t)nch, 32); w.put((uint32_t)P.rows, 32);
  // flags need the plans: compute in a first sweep, write entries in the same sweep into a second writer position
  Writer we; we.init(rec + sizeof(RecordHeader));
  unsigned long long alg = 0;
  { Gen g; g.v = 0.0; uint32_t row_base = 0;
    for (int c = 0; c < nch; ++c) {
      const int r0 = c * P.rows_per_chunk, n = min(P.rows_per_chunk, P.rows - r0);
      ChunkPlan cp; plan_chunk(P, key, r0, n, g, cp);
      if (cp.tsp.kind != 1) flags &= ~REC_ALL_TS_CONST;
      if (cp.drop) flags |= REC_ANY_DROP;
      if (cp.val_kind != 0) flags |= REC_ANY_DECODE;
      we.put((uint64_t)cp.start_t, 64); we.put((uint64_t)cp.end_t, 64); we.put((uint32_t)n, 32); we.put(off, 32);
      off += align_up(cp.ts_total, 8);
      we.put(off, 32); we.put(row_base, 32);
      off += align_up(cp.val_total, 8); row_base += (uint32_t)n;
      alg += 28 + 16 + cp.ts_total + cp.val_total;
      for (int k = 0; k < n; ++k) gen_value(P, key, r0 + k, k == n - 1, g);
    } }
  w.put(flags, 32);
  // vectors
  Writer wv; wv.init(rec + sizeof(RecordHeader) + (size_t)nch * sizeof(ChunkEntry));
  Gen g; g.v = 0.0;
  for (int c = 0; c < nch; ++c) {
    const int r0 = c * P.rows_per_chunk, n = min(P.rows_per_chunk, P.rows - r0);
    const Gen g0 = g;
    ChunkPlan cp; plan_chunk(P, key, r0, n, g0, cp);
    auto ts_seq = [&](auto&& f) { for (int k = 0; k < n; ++k) f(k, gen_ts(P, key, r0 + k)); };
    write_longs(wv, cp.tsp, n, false, false, ts_seq);
    wv.pad8();
    if (cp.val_kind == 2) {
      auto v_seq = [&](auto&& f) { Gen gg = g0; for (int k = 0; k < n; ++k) f(k, (int64_t)gen_value(P, key, r0 + k, k == n - 1, gg)); };
      write_longs(wv, cp.vlp, n, cp.drop, true, v_seq);
    } else if (cp.val_kind == 0) {
      auto v_seq = [&](auto&& f) { Gen gg = g0; for (int k = 0; k < n; ++k) f(k, __double_as_longlong(gen_value(P, key, r0 + k, k == n - 1, gg))); };
      LongPlan raw; raw.kind = 0; raw.total = cp.val_total;
      write_longs(wv, raw, n, cp.drop, true, v_seq);
    } else {
      const int ng = (n - 1 + 7) / 8;
      const uint32_t payloadOff = align_up(16 + 2 * (uint32_t)ng, 8);
      wv.put(cp.val_total - 4, 32); wv.put((uint32_t)WIRE_XOR | (cp.drop ? 0x80000000u : 0u), 32);
      wv.put((uint32_t)n, 32); wv.put((uint32_t)ng | (payloadOff << 16), 32);
      { // group offset table
        uint32_t goff = 0; uint64_t arr[8]; uint64_t lastb = 0; Gen gg = g0;
        for (int k = 0; k < n; ++k) {
          const uint64_t b = (uint64_t)__double_as_longlong(gen_value(P, key, r0 + k, k == n - 1, gg));
          if (k > 0) { arr[(k - 1) & 7] = b ^ lastb; if (((k - 1) & 7) == 7) { wv.put(goff, 16); goff += pack8_size(arr); } }
          lastb = b;
        }
        if ((n - 1) & 7) wv.put(goff, 16);
        wv.pad8();
      }
      { // payload
        uint64_t arr[8]; uint64_t lastb = 0; Gen gg = g0;
        for (int k = 0; k < n; ++k) {
          const uint64_t b = (uint64_t)__double_as_longlong(gen_value(P, key, r0 + k, k == n - 1, gg));
          if (k == 0) wv.put(b, 64);
          else { arr[(k - 1) & 7] = b ^ lastb; if (((k - 1) & 7) == 7) pack8_emit(wv, arr); }
          lastb = b;
        }
        if ((n - 1) & 7) { for (int j = (n - 1) & 7; j < 8; ++j) arr[j] = 0; pack8_emit(wv, arr); }
      }
    }
    wv.pad8();
    for (int k = 0; k < n; ++k) gen_value(P, key, r0 + k, k == n - 1, g);
  }
  // zero the tail padding (records are multiples of 16)
  while (wv.pos() < rec_bytes - (uint32_t)(sizeof(RecordHeader) + (size_t)nch * sizeof(ChunkEntry))) wv.put(0, 64);
  atomicAdd(alg_bytes, alg);
}

// ---------------------------------------------------------------------------------------------------------------------
// Histogram columns: SectDelta HistogramVectors written on the device (AppendableSectDeltaHistVector.appendHist, HistogramVector.scala:
// 489-545; SectionWriter, Section.scala:91-145; DeltaSectDiffPackSink, NibblePack.scala:296-345; BinaryHistogram delta formats :84-200).
// Vector = [+0 i32 numBytes][+4 u16 wire H_SECTDELTA][+6 u16 numHistograms][+8 u8 formatCode][+9 u16 bucketDefBytes][+11 bucket def body]
// sections...; section = [u16 bytes][u8 elems][u8 type (1 = drop)] records...; record = [u16 len] NibblePack groups.  The first record of a
// section holds the histogram's bucket deltas (NibblePack.packDelta), the others the difference to that record's deltas; a histogram with a
// bucket delta below the previous histogram's opens a Drop section.  Thread per series, sequential (generator / encoder, not a hot path).
// ---------------------------------------------------------------------------------------------------------------------
struct HistSynthParams {
  int64_t n_series; int32_t rows, rows_per_chunk; int64_t t0; int32_t interval;
  int32_t nb, format_code, def_bytes; const uint8_t* def;          // bucket definition (u16 length prefix + body) as BinaryHistogram carries it
  int32_t reset_period, n_groups; uint64_t seed; int64_t gid_base;
  const int64_t* ext_ts; const int64_t* ext_buckets;               // external samples: [n_series][rows] and [n_series][rows][nb] cumulative counts
};
constexpr int HS_MAXNB = 64;
// cumulative (over rows) per-bucket counts of the generator: row r adds 1 + (hash % 3) observations to bucket (r + series) % nb
// (gateway/src/main/scala/filodb/timeseries/TestTimeseriesProducer.scala:229-248); series with gid % reset_period == 0 restart at 5/8 of the rows
struct HistGen { int64_t cnt[HS_MAXNB]; };
__device__ __forceinline__ void hist_row(const HistSynthParams& P, int64_t si, uint64_t key, uint64_t gid, int row, HistGen& g, int64_t v[HS_MAXNB]) {
  if (P.ext_buckets) { const int64_t* src = P.ext_buckets + ((size_t)si * (size_t)P.rows + (size_t)row) * (size_t)P.nb; for (int b = 0; b < P.nb; ++b) v[b] = src[b]; return; }
  if (P.reset_period > 0 && (gid % (uint64_t)P.reset_period) == 0 && row == (P.rows * 5) / 8) for (int b = 0; b < P.nb; ++b) g.cnt[b] = 0;
  g.cnt[(int)(((uint64_t)row + gid) % (uint64_t)P.nb)] += 1 + (int64_t)(row_hash(key, row, 7) % 3ull);
  int64_t run = 0;
  for (int b = 0; b < P.nb; ++b) { run += g.cnt[b]; v[b] = run; }
}
__device__ __forceinline__ int64_t hist_ts(const HistSynthParams& P, int64_t si, int row) {
  return P.ext_ts ? P.ext_ts[(size_t)si * (size_t)P.rows + (size_t)row] : P.t0 + (int64_t)row * P.interval;
}
// NibblePack.pack8 of the nb values of `in` (groups of 8, zero padded) into out; returns the byte count
__device__ __forceinline__ int hs_pack_groups(const uint64_t* in, int nb, uint8_t* out, bool emit) {
  int len = 0;
  for (int i = 0; i < nb; i += 8) {
    uint64_t arr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) arr[j] = (i + j < nb) ? in[i + j] : 0ull;
    if (!emit) { len += (int)pack8_size(arr); continue; }
    uint64_t tmp[10]; Writer w; w.init(reinterpret_cast<uint8_t*>(tmp));
    pack8_emit(w, arr);
    const int n = (int)w.pos(); w.pad8();
    const uint8_t* tb = reinterpret_cast<const uint8_t*>(tmp);
    for (int k = 0; k < n; ++k) out[len + k] = tb[k];
    len += n;
  }
  return len;
}
// one chunk's histogram vector at dst (EMIT) or its size only; g is advanced over the chunk's rows
template <bool EMIT>
__device__ uint32_t hist_encode_chunk(const HistSynthParams& P, int64_t si, uint64_t key, uint64_t gid, int r0, int n, HistGen& g, uint8_t* dst) {
  const int nb = P.nb;
  auto put16 = [&](uint32_t off, uint32_t v) { if (EMIT) { dst[off] = (uint8_t)v; dst[off + 1] = (uint8_t)(v >> 8); } };
  auto put8 = [&](uint32_t off, uint32_t v) { if (EMIT) dst[off] = (uint8_t)v; };
  if (EMIT) {
    put16(4, (uint32_t)WIRE_H_SECTDELTA); put16(6, (uint32_t)n); put8(8, (uint32_t)P.format_code);
    for (int k = 0; k < P.def_bytes; ++k) dst[9 + k] = P.def[k];          // the serialized definition's u16 length IS the header's bucketDefBytes at +9, its body sits at +11
  }
  uint32_t cur = 9u + (uint32_t)P.def_bytes, secBytes = 0, secElems = 0;       // SectionWriter state
  put16(cur, 0); put8(cur + 2, 0); put8(cur + 3, 0);
  int64_t orig[HS_MAXNB], last[HS_MAXNB];
  for (int b = 0; b < nb; ++b) { orig[b] = 0; last[b] = 0; }
  uint8_t ob[8 * 66], rb[8 * 66];
  auto need_new = [&](int bytes) { return secElems >= 16u || secBytes + (uint32_t)bytes >= 65536u; };
  auto new_section = [&](int type) { cur = cur + 4 + secBytes; secBytes = 0; secElems = 0; put16(cur, 0); put8(cur + 2, 0); put8(cur + 3, (uint32_t)type); };
  auto add_blob = [&](const uint8_t* blob, int len) {
    const uint32_t w = cur + 4 + secBytes;
    put16(w, (uint32_t)len);
    if (EMIT) for (int k = 0; k < len; ++k) dst[w + 2 + k] = blob[k];
    secBytes += (uint32_t)len + 2; secElems += 1;
    put16(cur, secBytes); put8(cur + 2, secElems);
  };
  for (int i = 0; i < n; ++i) {
    int64_t v[HS_MAXNB]; uint64_t d[HS_MAXNB], pk[HS_MAXNB];
    hist_row(P, si, key, gid, r0 + i, g, v);
    bool dropped = false; int64_t prev = 0;
    for (int b = 0; b < nb; ++b) {                          // NibblePack.packDelta: delta to the previous bucket, 0 when it decreases
      const int64_t dl = v[b] >= prev ? v[b] - prev : 0; prev = v[b];
      d[b] = (uint64_t)dl; if (dl < last[b]) dropped = true;
      pk[b] = (uint64_t)dl - (uint64_t)orig[b];
    }
    for (int b = 0; b < nb; ++b) last[b] = (int64_t)d[b];
    const int olen = hs_pack_groups(d, nb, ob, EMIT);
    if (dropped) { for (int b = 0; b < nb; ++b) orig[b] = last[b]; new_section(1); add_blob(ob, olen); }
    else if (i == 0 || need_new(olen)) { for (int b = 0; b < nb; ++b) orig[b] = last[b]; if (need_new(olen)) new_section(0); add_blob(ob, olen); }
    else { const int rlen = hs_pack_groups(pk, nb, rb, EMIT); if (need_new(rlen)) new_section(0); add_blob(rb, rlen); }
  }
  const uint32_t total = cur + 4 + secBytes;
  if (EMIT) { const uint32_t nbytes = total - 4; dst[0] = (uint8_t)nbytes; dst[1] = (uint8_t)(nbytes >> 8); dst[2] = (uint8_t)(nbytes >> 16); dst[3] = (uint8_t)(nbytes >> 24); }
  return total;
}
__device__ __forceinline__ LongPlan hist_ts_plan(const HistSynthParams& P, int64_t si, int r0, int n) {
  auto ts_seq = [&](auto&& f) { for (int i = 0; i < n; ++i) f(i, hist_ts(P, si, r0 + i)); };
  return plan_longs(n, true, ts_seq);
}
__global__ void hist_synth_size_kernel(HistSynthParams P, uint32_t* rec_bytes, int32_t* group_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_series) return;
  const uint64_t gid = (uint64_t)(P.gid_base + i), key = series_key(P.seed, gid);
  const int nch = (P.rows + P.rows_per_chunk - 1) / P.rows_per_chunk;
  uint32_t bytes = sizeof(RecordHeader) + (uint32_t)nch * sizeof(ChunkEntry);
  HistGen g; for (int b = 0; b < HS_MAXNB; ++b) g.cnt[b] = 0;
  for (int c = 0; c < nch; ++c) {
    const int r0 = c * P.rows_per_chunk, n = min(P.rows_per_chunk, P.rows - r0);
    bytes += align_up(hist_ts_plan(P, i, r0, n).total, 8) + align_up(hist_encode_chunk<false>(P, i, key, gid, r0, n, g, nullptr), 8);
  }
  rec_bytes[i] = align_up(bytes, 16);
  if (group_ids) group_ids[i] = P.n_groups > 0 ? (int32_t)(splitmix64(P.seed ^ 0xA5A5A5A5ull ^ (gid * 0x9E3779B97F4A7C15ull)) % (uint64_t)P.n_groups) : 0;
}
__global__ void hist_synth_fill_kernel(HistSynthParams P, const int64_t* rec_off, uint8_t* arena, unsigned long long* alg_bytes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_series) return;
  const uint64_t gid = (uint64_t)(P.gid_base + i), key = series_key(P.seed, gid);
  const int nch = (P.rows + P.rows_per_chunk - 1) / P.rows_per_chunk;
  uint8_t* rec = arena + rec_off[i];
  const uint32_t rec_bytes = (uint32_t)(rec_off[i + 1] - rec_off[i]);
  for (uint32_t k = 0; k < rec_bytes; k += 8) *reinterpret_cast<uint64_t*>(rec + k) = 0ull;      // padding and gaps are zero
  uint32_t off = sizeof(RecordHeader) + (uint32_t)nch * sizeof(ChunkEntry), flags = REC_ALL_TS_CONST | REC_HIST | REC_ANY_DECODE, row_base = 0;
  unsigned long long alg = 0;
  HistGen g; for (int b = 0; b < HS_MAXNB; ++b) g.cnt[b] = 0;
  for (int c = 0; c < nch; ++c) {
    const int r0 = c * P.rows_per_chunk, n = min(P.rows_per_chunk, P.rows - r0);
    const LongPlan tsp = hist_ts_plan(P, i, r0, n);
    if (tsp.kind != 1) flags &= ~REC_ALL_TS_CONST;
    ChunkEntry e; e.start_time = hist_ts(P, i, r0); e.end_time = hist_ts(P, i, r0 + n - 1); e.num_rows = n; e.ts_off = off; e.row_base = row_base;
    { Writer w; w.init(rec + off); auto ts_seq = [&](auto&& f) { for (int k = 0; k < n; ++k) f(k, hist_ts(P, i, r0 + k)); }; write_longs(w, tsp, n, false, false, ts_seq); w.pad8(); }
    off += align_up(tsp.total, 8);
    e.val_off = off;
    const uint32_t vt = hist_encode_chunk<true>(P, i, key, gid, r0, n, g, rec + off);
    off += align_up(vt, 8); row_base += (uint32_t)n;
    reinterpret_cast<ChunkEntry*>(rec + sizeof(RecordHeader))[c] = e;
    alg += 28 + 16 + tsp.total + vt;
  }
  RecordHeader h; h.rec_bytes = rec_bytes; h.n_chunks = (uint32_t)nch; h.n_rows = (uint32_t)P.rows; h.flags = flags;
  *reinterpret_cast<RecordHeader*>(rec) = h;
  atomicAdd(alg_bytes, alg);
}

__global__ void widen_kernel(const uint32_t* in, int64_t* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = in[i];
}

} // namespace filo

using namespace filo;
#define S_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e); \
  return filo_internal_fail(ctx, _e == cudaErrorMemoryAllocation ? FILO_ERR_OOM : FILO_ERR_CUDA, m.c_str()); } } while (0)

static int32_t synth_build(filo_ctx* ctx, const filo_synth_spec* sp, const int64_t* d_ext_ts, const double* d_ext_vals, const int32_t* h_group_ids, bool any_nonconst_ts, filo_table** out);
extern "C" int32_t filo_synth_table(filo_ctx* ctx, const filo_synth_spec* sp, filo_table** out) {
  if (!ctx || !sp || !out) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_synth_table: null argument");
  if (sp->n_series < 0 || sp->rows_per_series <= 0 || sp->rows_per_chunk <= 0 || sp->interval_ms <= 0 || !sp->sin_table ||
      sp->ts_jitter_ms < 0 || 2 * (int64_t)sp->ts_jitter_ms >= sp->interval_ms || sp->value_kind < 0 || sp->value_kind > 2 ||
      sp->value_enc < 0 || sp->value_enc > 2 || sp->rows_per_chunk > 4096)
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_synth_table: bad spec");
  return synth_build(ctx, sp, nullptr, nullptr, nullptr, sp->ts_jitter_ms > 250, out);
}

// GPU-side encode of an ingest batch: raw samples (row-major [n_series][rows] timestamps and values in HOST memory) are copied to the
// device and encoded there into the chunk vectors FiloDB's appenders + optimize() would write -- timestamps through
// DeltaDeltaVector.fromLongVector with the +-250 ms approximate-const rule (DeltaDeltaVector.scala:20-80), values as raw doubles /
// this repo's XOR-NibblePack container / DoubleVector.optimize (integral values -> DDV longs, DoubleVector.scala:86-96), the counter
// drop flag from DoubleCounterAppender (DoubleVector.scala:456-466) -- chunked every rows_per_chunk rows, straight into a resident table.
extern "C" int32_t filo_encode_table(filo_ctx* ctx, const int64_t* timestamps, const double* values, int64_t n_series, int32_t rows_per_series,
                                     int32_t rows_per_chunk, int32_t value_enc, int32_t schema_flags, const int32_t* group_ids, int32_t n_groups,
                                     filo_table** out) {
  if (!ctx || !timestamps || !values || !out) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_table: null argument");
  if (n_series < 0 || rows_per_series <= 0 || rows_per_chunk <= 0 || rows_per_chunk > 4096 || value_enc < 0 || value_enc > 2 || (group_ids && n_groups <= 0))
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_table: bad arguments");
  for (int64_t i = 0; group_ids && i < n_series; ++i) if (group_ids[i] < 0 || group_ids[i] >= n_groups) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "group id out of range");
  S_TRY(cudaSetDevice(filo_internal_device(ctx)));
  cudaStream_t s = filo_internal_stream(ctx);
  const size_t n = (size_t)n_series * (size_t)rows_per_series;
  int64_t* d_ts = nullptr; double* d_v = nullptr;
  S_TRY(cudaMalloc(&d_ts, std::max<size_t>(n, 1) * 8));
  if (cudaMalloc(&d_v, std::max<size_t>(n, 1) * 8) != cudaSuccess) { cudaFree(d_ts); return filo_internal_fail(ctx, FILO_ERR_OOM, "filo_encode_table: staging"); }
  struct FreeIn { void* a; void* b; ~FreeIn() { cudaFree(a); cudaFree(b); } } guard{d_ts, d_v};
  S_TRY(cudaMemcpyAsync(d_ts, timestamps, n * 8, cudaMemcpyHostToDevice, s));
  S_TRY(cudaMemcpyAsync(d_v, values, n * 8, cudaMemcpyHostToDevice, s));
  // timestamps must increase inside a series (TimeSeriesPartition.ingest drops out-of-order samples, TimeSeriesPartition.scala:135-136)
  for (int64_t i = 0; i < n_series; ++i)
    for (int32_t r = 1; r < rows_per_series; ++r)
      if (!(timestamps[(size_t)i * rows_per_series + r] > timestamps[(size_t)i * rows_per_series + r - 1]))
        return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_table: timestamps of a series must be strictly increasing");
  filo_synth_spec sp{};
  sp.n_series = n_series; sp.rows_per_series = rows_per_series; sp.rows_per_chunk = rows_per_chunk; sp.t0_ms = 0; sp.interval_ms = 1; sp.ts_jitter_ms = 0;
  sp.value_kind = 0; sp.value_enc = value_enc; sp.reset_period = 0; sp.nan_per_million = 0; sp.n_groups = group_ids ? n_groups : 0; sp.schema_flags = schema_flags;
  sp.seed = 0; sp.series_id_base = 0; sp.sin_table = nullptr;
  return synth_build(ctx, &sp, d_ts, d_v, group_ids, true, out);
}

static int32_t synth_build(filo_ctx* ctx, const filo_synth_spec* sp, const int64_t* d_ext_ts, const double* d_ext_vals, const int32_t* h_group_ids, bool any_nonconst_ts, filo_table** out) {
  S_TRY(cudaSetDevice(filo_internal_device(ctx)));
  cudaStream_t s = filo_internal_stream(ctx);
  const int64_t S = sp->n_series;
  double* d_sin = nullptr; uint32_t* d_bytes = nullptr; int32_t* d_gid = nullptr; int64_t* d_off = nullptr; int64_t* d_wide = nullptr;
  unsigned long long* d_alg = nullptr; uint8_t* d_arena = nullptr;
  S_TRY(cudaMalloc(&d_sin, (size_t)sp->rows_per_series * 8));
  if (sp->sin_table) S_TRY(cudaMemcpyAsync(d_sin, sp->sin_table, (size_t)sp->rows_per_series * 8, cudaMemcpyHostToDevice, s));
  S_TRY(cudaMalloc(&d_bytes, (size_t)(S + 1) * 4)); S_TRY(cudaMalloc(&d_wide, (size_t)(S + 1) * 8)); S_TRY(cudaMalloc(&d_off, (size_t)(S + 1) * 8));
  S_TRY(cudaMemsetAsync(d_bytes, 0, (size_t)(S + 1) * 4, s));
  if (sp->n_groups > 0) S_TRY(cudaMalloc(&d_gid, (size_t)std::max<int64_t>(S, 1) * 4));
  S_TRY(cudaMalloc(&d_alg, 8)); S_TRY(cudaMemsetAsync(d_alg, 0, 8, s));
  SynthParams P{S, sp->rows_per_series, sp->rows_per_chunk, sp->t0_ms, sp->interval_ms, sp->ts_jitter_ms, sp->value_kind, sp->value_enc,
                sp->reset_period, sp->nan_per_million, sp->n_groups, (sp->schema_flags & FILO_SCHEMA_CUMULATIVE) ? 1 : 0,
                sp->seed, sp->series_id_base, d_sin, 1.0 / 37837.22772881784};
  P.ext_ts = d_ext_ts; P.ext_vals = d_ext_vals;
  const unsigned blocks = (unsigned)((S + 127) / 128);
  if (S > 0) { synth_size_kernel<<<blocks, 128, 0, s>>>(P, d_bytes, d_gid); S_TRY(cudaGetLastError()); }
  if (h_group_ids && d_gid && S > 0) S_TRY(cudaMemcpyAsync(d_gid, h_group_ids, (size_t)S * 4, cudaMemcpyHostToDevice, s));      // the caller's grouping replaces the generator's
  widen_kernel<<<(unsigned)((S + 1 + 255) / 256), 256, 0, s>>>(d_bytes, d_wide, S + 1); S_TRY(cudaGetLastError());
  size_t tmpb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tmpb, d_wide, d_off, (int)(S + 1), s);
  void* tmp = nullptr; S_TRY(cudaMalloc(&tmp, tmpb + 16));
  S_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmpb, d_wide, d_off, (int)(S + 1), s));
  int64_t arena_bytes = 0;
  S_TRY(cudaMemcpyAsync(&arena_bytes, d_off + S, 8, cudaMemcpyDeviceToHost, s));
  S_TRY(cudaStreamSynchronize(s));
  S_TRY(cudaMalloc(&d_arena, (size_t)arena_bytes + 64));
  S_TRY(cudaMemsetAsync(d_arena + arena_bytes, 0, 64, s));
  if (S > 0) { synth_fill_kernel<<<blocks, 128, 0, s>>>(P, d_off, d_arena, d_alg); S_TRY(cudaGetLastError()); }
  unsigned long long alg = 0;
  S_TRY(cudaMemcpyAsync(&alg, d_alg, 8, cudaMemcpyDeviceToHost, s));
  S_TRY(cudaStreamSynchronize(s));
  uint32_t max_rec = 0;
  if (S > 0) {
    uint32_t* d_max = nullptr; S_TRY(cudaMalloc(&d_max, 4));
    size_t tb = 0; cub::DeviceReduce::Max(nullptr, tb, d_bytes, d_max, (int)S, s);
    void* t2 = nullptr; S_TRY(cudaMalloc(&t2, tb + 16));
    S_TRY(cub::DeviceReduce::Max(t2, tb, d_bytes, d_max, (int)S, s));
    S_TRY(cudaMemcpyAsync(&max_rec, d_max, 4, cudaMemcpyDeviceToHost, s));
    S_TRY(cudaStreamSynchronize(s));
    cudaFree(t2); cudaFree(d_max);
  }
  cudaFree(tmp); cudaFree(d_bytes); cudaFree(d_wide); cudaFree(d_sin); cudaFree(d_alg);
  const int nch = (sp->rows_per_series + sp->rows_per_chunk - 1) / sp->rows_per_chunk;
  filo_table* t = filo_internal_new_table();
  filo_internal_set_arena(t, d_arena, d_off, S, S * nch, S * (int64_t)sp->rows_per_series, arena_bytes + (S + 1) * 8, (int64_t)alg,
                          sp->rows_per_series, nch, sp->schema_flags);
  filo_internal_set_layout(t, max_rec, any_nonconst_ts, (sp->schema_flags & FILO_SCHEMA_CUMULATIVE) != 0);
  int32_t rc = filo_internal_finish_table(ctx, t, d_gid, sp->n_groups > 0 ? sp->n_groups : 1);
  cudaFree(d_gid);
  if (rc) { filo_table_free(ctx, t); return rc; }
  *out = t;
  return FILO_OK;
}


// ---- histogram tables: generator (bench / tests) and encoder of raw bucket counts (ingest batches)
static int32_t hist_build(filo_ctx* ctx, HistSynthParams P, const uint8_t* h_def, int32_t schema_flags, const int32_t* h_group_ids, filo_table** out) {
  S_TRY(cudaSetDevice(filo_internal_device(ctx)));
  cudaStream_t s = filo_internal_stream(ctx);
  const int64_t S = P.n_series;
  uint8_t* d_def = nullptr; uint32_t* d_bytes = nullptr; int32_t* d_gid = nullptr; int64_t* d_off = nullptr; int64_t* d_wide = nullptr;
  unsigned long long* d_alg = nullptr; uint8_t* d_arena = nullptr;
  S_TRY(cudaMalloc(&d_def, (size_t)P.def_bytes + 16));
  S_TRY(cudaMemcpyAsync(d_def, h_def, (size_t)P.def_bytes, cudaMemcpyHostToDevice, s));
  P.def = d_def;
  S_TRY(cudaMalloc(&d_bytes, (size_t)(S + 1) * 4)); S_TRY(cudaMalloc(&d_wide, (size_t)(S + 1) * 8)); S_TRY(cudaMalloc(&d_off, (size_t)(S + 1) * 8));
  S_TRY(cudaMemsetAsync(d_bytes, 0, (size_t)(S + 1) * 4, s));
  if (P.n_groups > 0) S_TRY(cudaMalloc(&d_gid, (size_t)std::max<int64_t>(S, 1) * 4));
  S_TRY(cudaMalloc(&d_alg, 8)); S_TRY(cudaMemsetAsync(d_alg, 0, 8, s));
  const unsigned blocks = (unsigned)((S + 63) / 64);
  if (S > 0) { hist_synth_size_kernel<<<blocks, 64, 0, s>>>(P, d_bytes, d_gid); S_TRY(cudaGetLastError()); }
  if (h_group_ids && d_gid && S > 0) S_TRY(cudaMemcpyAsync(d_gid, h_group_ids, (size_t)S * 4, cudaMemcpyHostToDevice, s));
  widen_kernel<<<(unsigned)((S + 1 + 255) / 256), 256, 0, s>>>(d_bytes, d_wide, S + 1); S_TRY(cudaGetLastError());
  size_t tmpb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tmpb, d_wide, d_off, (int)(S + 1), s);
  void* tmp = nullptr; S_TRY(cudaMalloc(&tmp, tmpb + 16));
  S_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmpb, d_wide, d_off, (int)(S + 1), s));
  int64_t arena_bytes = 0;
  S_TRY(cudaMemcpyAsync(&arena_bytes, d_off + S, 8, cudaMemcpyDeviceToHost, s));
  S_TRY(cudaStreamSynchronize(s));
  S_TRY(cudaMalloc(&d_arena, (size_t)arena_bytes + 64));
  S_TRY(cudaMemsetAsync(d_arena + arena_bytes, 0, 64, s));
  if (S > 0) { hist_synth_fill_kernel<<<blocks, 64, 0, s>>>(P, d_off, d_arena, d_alg); S_TRY(cudaGetLastError()); }
  unsigned long long alg = 0;
  S_TRY(cudaMemcpyAsync(&alg, d_alg, 8, cudaMemcpyDeviceToHost, s));
  S_TRY(cudaStreamSynchronize(s));
  uint32_t max_rec = 0;
  if (S > 0) {
    uint32_t* d_max = nullptr; S_TRY(cudaMalloc(&d_max, 4));
    size_t tb = 0; cub::DeviceReduce::Max(nullptr, tb, d_bytes, d_max, (int)S, s);
    void* t2 = nullptr; S_TRY(cudaMalloc(&t2, tb + 16));
    S_TRY(cub::DeviceReduce::Max(t2, tb, d_bytes, d_max, (int)S, s));
    S_TRY(cudaMemcpyAsync(&max_rec, d_max, 4, cudaMemcpyDeviceToHost, s));
    S_TRY(cudaStreamSynchronize(s));
    cudaFree(t2); cudaFree(d_max);
  }
  cudaFree(tmp); cudaFree(d_bytes); cudaFree(d_wide); cudaFree(d_alg); cudaFree(d_def);
  const int nch = (P.rows + P.rows_per_chunk - 1) / P.rows_per_chunk;
  filo_table* t = filo_internal_new_table();
  filo_internal_set_arena(t, d_arena, d_off, S, S * nch, S * (int64_t)P.rows, arena_bytes + (S + 1) * 8, (int64_t)alg, P.rows, nch, schema_flags);
  filo_internal_set_layout(t, max_rec, P.ext_ts != nullptr, false);
  std::vector<uint8_t> hd((size_t)11 + (size_t)P.def_bytes + 32, 0);       // a HistogramVector header for the table's bucket scheme
  hd[8] = (uint8_t)P.format_code;
  std::memcpy(hd.data() + 9, h_def, (size_t)P.def_bytes);
  int32_t rc = filo_internal_set_hist(ctx, t, hd.data());
  if (rc == FILO_OK) rc = filo_internal_finish_table(ctx, t, d_gid, P.n_groups > 0 ? P.n_groups : 1);
  cudaFree(d_gid);
  if (rc) { filo_table_free(ctx, t); return rc; }
  *out = t;
  return FILO_OK;
}
static bool hist_def_ok(int32_t n_buckets, int32_t format_code, const uint8_t* def, int32_t def_bytes) {
  if (!def || n_buckets <= 0 || n_buckets > HS_MAXNB || def_bytes < 4 || def_bytes > 1024) return false;
  if (!(format_code == 3 || format_code == 4 || format_code == 5 || format_code == 9)) return false;   // 9: otel exponential, an 18-byte definition
  const int len = def[0] | (def[1] << 8), n = def[2] | (def[3] << 8);
  return len + 2 == def_bytes && n == n_buckets;
}
extern "C" int32_t filo_synth_hist_table(filo_ctx* ctx, int64_t n_series, int32_t rows_per_series, int32_t rows_per_chunk, int64_t t0_ms, int32_t interval_ms,
                                         int32_t n_buckets, int32_t format_code, const uint8_t* bucket_def, int32_t bucket_def_bytes,
                                         int32_t reset_period, int32_t n_groups, uint64_t seed, int64_t series_id_base, filo_table** out) {
  if (!ctx || !out) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_synth_hist_table: null argument");
  if (n_series < 0 || rows_per_series <= 0 || rows_per_chunk <= 0 || rows_per_chunk > 4096 || interval_ms <= 0 || !hist_def_ok(n_buckets, format_code, bucket_def, bucket_def_bytes))
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_synth_hist_table: bad arguments (1..64 geometric or custom buckets)");
  HistSynthParams P{n_series, rows_per_series, rows_per_chunk, t0_ms, interval_ms, n_buckets, format_code, bucket_def_bytes, nullptr, reset_period, n_groups, seed, series_id_base, nullptr, nullptr};
  return hist_build(ctx, P, bucket_def, FILO_SCHEMA_CUMULATIVE, nullptr, out);
}
// GPU-side encode of a histogram ingest batch: cumulative bucket counts [n_series][rows][n_buckets] and timestamps [n_series][rows] in HOST memory
extern "C" int32_t filo_encode_hist_table(filo_ctx* ctx, const int64_t* timestamps, const int64_t* bucket_counts, int64_t n_series, int32_t rows_per_series,
                                          int32_t rows_per_chunk, int32_t n_buckets, int32_t format_code, const uint8_t* bucket_def, int32_t bucket_def_bytes,
                                          int32_t schema_flags, const int32_t* group_ids, int32_t n_groups, filo_table** out) {
  if (!ctx || !timestamps || !bucket_counts || !out) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_hist_table: null argument");
  if (n_series < 0 || rows_per_series <= 0 || rows_per_chunk <= 0 || rows_per_chunk > 4096 || (group_ids && n_groups <= 0) || !hist_def_ok(n_buckets, format_code, bucket_def, bucket_def_bytes))
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_hist_table: bad arguments (1..64 geometric or custom buckets)");
  for (int64_t i = 0; group_ids && i < n_series; ++i) if (group_ids[i] < 0 || group_ids[i] >= n_groups) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "group id out of range");
  for (int64_t i = 0; i < n_series; ++i)
    for (int32_t r = 1; r < rows_per_series; ++r)
      if (!(timestamps[(size_t)i * rows_per_series + r] > timestamps[(size_t)i * rows_per_series + r - 1]))
        return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_hist_table: timestamps of a series must be strictly increasing");
  S_TRY(cudaSetDevice(filo_internal_device(ctx)));
  cudaStream_t s = filo_internal_stream(ctx);
  const size_t n = (size_t)n_series * (size_t)rows_per_series;
  int64_t* d_ts = nullptr; int64_t* d_b = nullptr;
  S_TRY(cudaMalloc(&d_ts, std::max<size_t>(n, 1) * 8));
  if (cudaMalloc(&d_b, std::max<size_t>(n * (size_t)n_buckets, 1) * 8) != cudaSuccess) { cudaFree(d_ts); return filo_internal_fail(ctx, FILO_ERR_OOM, "filo_encode_hist_table: staging"); }
  struct FreeIn { void* a; void* b; ~FreeIn() { cudaFree(a); cudaFree(b); } } guard{d_ts, d_b};
  S_TRY(cudaMemcpyAsync(d_ts, timestamps, n * 8, cudaMemcpyHostToDevice, s));
  S_TRY(cudaMemcpyAsync(d_b, bucket_counts, n * (size_t)n_buckets * 8, cudaMemcpyHostToDevice, s));
  HistSynthParams P{n_series, rows_per_series, rows_per_chunk, 0, 1, n_buckets, format_code, bucket_def_bytes, nullptr, 0, group_ids ? n_groups : 0, 0, 0, d_ts, d_b};
  return hist_build(ctx, P, bucket_def, schema_flags, group_ids, out);
}
