// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over filo_format.hpp / filo_query.hpp so that
// tests/ (ctypes), __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can drive it.
// Nothing under filodb_b200/ may link or load this library.
#include "filo_format.hpp"
#include "filo_query.hpp"
#include "filo_result.hpp"
#include <thread>
#include <atomic>
#include <chrono>

using namespace fo;

namespace {
struct Chunk {
  std::vector<uint8_t> ts, val;
  std::vector<uint8_t> info;   // ChunkSetInfo, 28 + 8*2 bytes, real pointers inside
};
struct StoreSeries { std::vector<std::unique_ptr<Chunk>> chunks; };
struct Store {
  std::vector<std::unique_ptr<StoreSeries>> series;
  std::string err;
};
thread_local std::string g_err;

void finishChunk(Chunk& c, int64_t startTime, int64_t endTime, int32_t numRows, int64_t ingestionTimeMs,
                 const uint8_t* tsExt = nullptr, const uint8_t* valExt = nullptr) {
  c.info.assign(csi::OffsetVectors + 16, 0);
  setLong(c.info.data() + csi::OffsetChunkID, csi::chunkID(startTime, ingestionTimeMs / 1000));
  setInt(c.info.data() + csi::OffsetNumRows, numRows);
  setLong(c.info.data() + csi::OffsetIngestionTime, ingestionTimeMs);
  setLong(c.info.data() + csi::OffsetEndTime, endTime);
  setLong(c.info.data() + csi::OffsetVectors, (int64_t)(uintptr_t)(tsExt ? tsExt : c.ts.data()));
  setLong(c.info.data() + csi::OffsetVectors + 8, (int64_t)(uintptr_t)(valExt ? valExt : c.val.data()));
}
}

extern "C" {

const char* fo_last_error() { return g_err.c_str(); }

// ---------------- NibblePack
int32_t fo_pack8(const uint64_t* in8, uint8_t* out, int32_t cap) {
  std::vector<uint8_t> buf; int end = nibble::pack8(in8, buf, 0);
  if (end > cap) return -1;
  std::memcpy(out, buf.data(), end); return end;
}
int32_t fo_unpack8(const uint8_t* in, int32_t cap, uint64_t* out8, int32_t* remaining) {
  Ptr p = in; int c = cap;
  int r = nibble::unpack8(p, c, out8);
  *remaining = c; return r;
}
int32_t fo_pack_doubles(const double* in, int32_t n, uint8_t* out, int32_t cap) {
  std::vector<uint8_t> buf; int end = nibble::packDoubles(in, n, buf, 0);
  if (end > cap) return -1;
  std::memcpy(out, buf.data(), end); return end;
}
int32_t fo_unpack_double_xor(const uint8_t* in, int32_t cap, double* out, int32_t n) { return nibble::unpackDoubleXOR(in, cap, out, n); }
int32_t fo_pack_delta(const int64_t* in, int32_t n, uint8_t* out, int32_t cap) {
  std::vector<uint8_t> buf; int end = nibble::packDelta(in, n, buf, 0);
  if (end > cap) return -1;
  std::memcpy(out, buf.data(), end); return end;
}
int32_t fo_unpack_delta(const uint8_t* in, int32_t cap, int64_t* out, int32_t n) { return nibble::unpackDelta(in, cap, out, n); }

// ---------------- encoders.  mode: 0 = reference optimize(), 1 = XOR container, 2 = raw (no optimize)
static int32_t emit(const enc::Bytes& b, uint8_t* out, int32_t cap) {
  if ((int)b.size() > cap) return -(int32_t)b.size();
  std::memcpy(out, b.data(), b.size()); return (int32_t)b.size();
}
int32_t fo_encode_timestamps(const int64_t* v, int32_t n, uint8_t* out, int32_t cap) { return emit(enc::timestamps(v, n), out, cap); }
int32_t fo_encode_longs(const int64_t* v, int32_t n, uint8_t* out, int32_t cap) { return emit(enc::longs(v, n), out, cap); }
int32_t fo_encode_doubles(const double* v, int32_t n, int32_t detectDrops, int32_t mode, uint8_t* out, int32_t cap) {
  try {
    if (mode == 1) return emit(enc::doublesXor(v, n, detectDrops != 0), out, cap);
    if (mode == 2) { enc::Bytes b = enc::rawDoubles(v, n); if (detectDrops && enc::counterDropFlag(v, n)) pv_markDrop(b.data()); return emit(b, out, cap); }
    return emit(enc::doubles(v, n, detectDrops != 0), out, cap);
  } catch (std::exception& e) { g_err = e.what(); return INT32_MIN; }
}
int32_t fo_encode_int_vector(const int32_t* v, int32_t n, int32_t nbits, int32_t sgn, uint8_t* out, int32_t cap) {
  return emit(enc::intVectorNoNA(v, n, nbits, sgn != 0), out, cap);
}
void fo_minmax_to_nbits(int32_t mn, int32_t mx, int32_t* nbits, int32_t* sgn) { int nb; bool s; enc::minMaxToNbitsSigned(mn, mx, nb, s); *nbits = nb; *sgn = s; }

// ---------------- readers (unit-level, for golden vector tests)
int32_t fo_long_length(const uint8_t* v) { try { return LongReader::of(v).length(v); } catch (std::exception& e) { g_err = e.what(); return INT32_MIN; } }
int64_t fo_long_apply(const uint8_t* v, int32_t n) { return LongReader::of(v).apply(v, n); }
int32_t fo_long_binary_search(const uint8_t* v, int64_t item) { return LongReader::of(v).binarySearch(v, item); }
int32_t fo_long_ceiling_index(const uint8_t* v, int64_t item) { return LongReader::of(v).ceilingIndex(v, item); }
double  fo_long_sum(const uint8_t* v, int32_t s, int32_t e) { return LongReader::of(v).sum(v, s, e); }
int32_t fo_int_length(const uint8_t* v) { return IntReader::simple(v).length(v); }
int32_t fo_int_apply(const uint8_t* v, int32_t n) { return IntReader::simple(v).apply(v, n); }
int64_t fo_int_sum(const uint8_t* v, int32_t s, int32_t e) { return IntReader::simple(v).sum(v, s, e); }
int32_t fo_double_length(const uint8_t* v) { try { return DoubleReader::of(v).length(); } catch (std::exception& e) { g_err = e.what(); return INT32_MIN; } }
double  fo_double_apply(const uint8_t* v, int32_t n) { return DoubleReader::of(v).apply(n); }
double  fo_double_sum(const uint8_t* v, int32_t s, int32_t e) { return DoubleReader::of(v).sum(s, e); }
int32_t fo_double_count(const uint8_t* v, int32_t s, int32_t e) { return DoubleReader::of(v).count(s, e); }
int32_t fo_double_dropped(const uint8_t* v) { return pv_dropped(v); }
int32_t fo_total_bytes(const uint8_t* v) { return totalBytes(v); }
int32_t fo_vector_type(const uint8_t* v) { return vectorType(v); }
// correction API.  meta_in/out = {some, lastValue, correction}
void fo_double_detect_drop(const uint8_t* v, const double* meta_in, double* meta_out) {
  DoubleCorrection m{meta_in[0] != 0, meta_in[1], meta_in[2]};
  DoubleCorrection r = DoubleReader::of(v).detectDropAndCorrection(m);
  meta_out[0] = r.some; meta_out[1] = r.lastValue; meta_out[2] = r.correction;
}
void fo_double_update_correction(const uint8_t* v, const double* meta_in, int32_t forceCorrected, double* meta_out) {
  DoubleCorrection m{meta_in[0] != 0, meta_in[1], meta_in[2]};
  DoubleReader rd = DoubleReader::of(v);
  if (forceCorrected && rd.correcting) rd.forceCorrected();
  DoubleCorrection r = rd.updateCorrection(m);
  meta_out[0] = r.some; meta_out[1] = r.lastValue; meta_out[2] = r.correction;
}
double fo_double_corrected_value(const uint8_t* v, int32_t n, const double* meta_in) {
  DoubleCorrection m{meta_in[0] != 0, meta_in[1], meta_in[2]};
  DoubleReader rd = DoubleReader::of(v);
  return rd.correctedValue(n, m);
}
int32_t fo_double_drop_positions(const uint8_t* v, int32_t* out, int32_t cap) {
  DoubleReader rd = DoubleReader::of(v);
  const auto& d = rd.dropPositions();
  for (size_t i = 0; i < d.size() && (int)i < cap; ++i) out[i] = d[i];
  return (int32_t)d.size();
}
double fo_extrapolated_rate(int64_t ws, int64_t we, int32_t n, int64_t t1, double v1, int64_t t2, double v2, int32_t isCounter, int32_t isRate) {
  return extrapolatedRate(ws, we, n, t1, v1, t2, v2, isCounter != 0, isRate != 0);
}
int64_t fo_chunk_id(int64_t startTime, int64_t ingestionSec) { return csi::chunkID(startTime, ingestionSec); }
int64_t fo_start_time_from_chunk_id(int64_t id) { return csi::startTimeFromChunkID(id); }

// ---------------- store of series/chunks with real ChunkSetInfo blocks
void* fo_store_new() { return new Store(); }
void fo_store_free(void* s) { delete (Store*)s; }
int64_t fo_store_num_series(void* s) { return (int64_t)((Store*)s)->series.size(); }
int64_t fo_store_add_series(void* sp) { Store* s = (Store*)sp; s->series.emplace_back(new StoreSeries()); return (int64_t)s->series.size() - 1; }

// Adds one chunk from row data, encoding with the reference appenders' optimize().
// valMode: 0 = DoubleVector.optimize, 1 = XOR container, 2 = raw f64.  tsMode: 0 = TimestampAppendingVector.optimize, 2 = raw.
int32_t fo_store_add_chunk(void* sp, int64_t series, const int64_t* ts, const double* vals, int32_t n,
                           int32_t valMode, int32_t detectDrops, int32_t tsMode) {
  Store* s = (Store*)sp;
  try {
    auto c = std::make_unique<Chunk>();
    c->ts = (tsMode == 2) ? enc::rawLongs(ts, n) : enc::timestamps(ts, n);
    if (valMode == 1) c->val = enc::doublesXor(vals, n, detectDrops != 0);
    else if (valMode == 2) { c->val = enc::rawDoubles(vals, n); if (detectDrops && enc::counterDropFlag(vals, n)) pv_markDrop(c->val.data()); }
    else c->val = enc::doubles(vals, n, detectDrops != 0);
    finishChunk(*c, ts[0], ts[n - 1], n, ts[n - 1] + 1000);
    s->series[series]->chunks.push_back(std::move(c));
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// Long-column chunk: the value vector comes from LongBinaryVector.appendingVectorNoNA(...).optimize() (enc::longs), masked != 0 wraps
// neither: masked vectors are injected with fo_store_add_chunk_raw.
int32_t fo_store_add_chunk_longs(void* sp, int64_t series, const int64_t* ts, const int64_t* vals, int32_t n, int32_t valRaw, int32_t tsMode) {
  Store* s = (Store*)sp;
  try {
    auto c = std::make_unique<Chunk>();
    c->ts = (tsMode == 2) ? enc::rawLongs(ts, n) : enc::timestamps(ts, n);
    c->val = valRaw ? enc::rawLongs(vals, n) : enc::longs(vals, n);
    finishChunk(*c, ts[0], ts[n - 1], n, ts[n - 1] + 1000);
    s->series[series]->chunks.push_back(std::move(c));
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// Adds a chunk from pre-encoded vector bytes (lets tests inject arbitrary/corrupt vectors and metadata).
int32_t fo_store_add_chunk_raw(void* sp, int64_t series, int64_t startTime, int64_t endTime, int32_t numRows,
                               const uint8_t* tsBytes, int32_t tsLen, const uint8_t* valBytes, int32_t valLen) {
  Store* s = (Store*)sp;
  auto c = std::make_unique<Chunk>();
  c->ts.assign(tsBytes, tsBytes + tsLen); c->val.assign(valBytes, valBytes + valLen);
  finishChunk(*c, startTime, endTime, numRows, endTime + 1000);
  s->series[series]->chunks.push_back(std::move(c));
  return 0;
}
// Bulk: one series from row arrays, split into chunks of chunkRows[i] rows.
int32_t fo_store_add_series_rows(void* sp, const int64_t* ts, const double* vals, int64_t n, const int32_t* chunkRows, int32_t nchunks,
                                 int32_t valMode, int32_t detectDrops, int32_t tsMode) {
  int64_t si = fo_store_add_series(sp);
  int64_t off = 0;
  for (int i = 0; i < nchunks; ++i) {
    int32_t r = chunkRows[i];
    if (r <= 0 || off + r > n) { g_err = "bad chunkRows"; return -1; }
    if (fo_store_add_chunk(sp, si, ts + off, vals + off, r, valMode, detectDrops, tsMode) != 0) return -1;
    off += r;
  }
  return 0;
}
// Builds series over a HOST copy of the product's device chunk arena (filodb_b200/csrc/filo_record.h layout: 16-byte record
// header {rec_bytes, n_chunks, n_rows, flags}, then 32-byte entries {start, end, numRows, tsOff, valOff, rowBase}).  Zero-copy:
// the ChunkSetInfo blocks point into `arena`, which must outlive the store.  Used by bench.py's cpu_baseline leg.
int64_t fo_store_add_from_arena(void* sp, const uint8_t* arena, const int64_t* rec_off, int64_t n_series) {
  Store* s = (Store*)sp;
  for (int64_t i = 0; i < n_series; ++i) {
    const uint8_t* rec = arena + rec_off[i];
    int32_t nch = getInt(rec + 4);
    int64_t si = fo_store_add_series(sp);
    for (int32_t c = 0; c < nch; ++c) {
      const uint8_t* e = rec + 16 + 32 * (int64_t)c;
      auto ch = std::make_unique<Chunk>();
      finishChunk(*ch, getLong(e), getLong(e + 8), getInt(e + 16), getLong(e + 8) + 1000,
                  rec + (uint32_t)getInt(e + 20), rec + (uint32_t)getInt(e + 24));
      s->series[si]->chunks.push_back(std::move(ch));
    }
  }
  return (int64_t)s->series.size();
}
// Deterministic synthetic series (same hash-based row generator as filodb_b200/csrc/synth_kernels.cu, restated here so the
// reference arm of bench.py needs nothing from the product), encoded with the appenders' optimize() restatement above.
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
int64_t fo_store_add_synth(void* sp, int64_t n_series, int32_t rows, int32_t rpc, int64_t t0, int32_t interval, int32_t jitter,
                           int32_t value_kind, int32_t value_enc, int32_t reset_period, int32_t nan_ppm, int32_t cumulative,
                           uint64_t seed, int64_t gid_base, const double* sin_table, int32_t nthreads) {
  Store* s = (Store*)sp;
  const int64_t first = (int64_t)s->series.size();
  for (int64_t i = 0; i < n_series; ++i) s->series.emplace_back(new StoreSeries());
  const double noise_scale = 1.0 / 37837.22772881784;
  std::atomic<int64_t> next{0};
  auto worker = [&]() {
    std::vector<int64_t> ts(rows); std::vector<double> vals(rows);
    for (;;) {
      int64_t i = next.fetch_add(1); if (i >= n_series) break;
      const uint64_t gid = (uint64_t)(gid_base + i);
      const uint64_t key = splitmix64(seed ^ (gid * 0xD1342543DE82EF95ull));
      auto rh = [&](int row, int salt) { return splitmix64(key + ((uint64_t)(uint32_t)row << 3) + (uint64_t)salt); };
      double v = 0.0;
      for (int r = 0; r < rows; ++r) {
        int64_t t = t0 + (int64_t)r * interval;
        if (jitter > 0) t += (int64_t)(rh(r, 3) % (uint64_t)(2 * jitter + 1)) - jitter;
        ts[r] = t;
        const bool lastInChunk = ((r % rpc) == rpc - 1) || r == rows - 1;
        if (lastInChunk && nan_ppm > 0 && (int)(rh(r, 1) % 1000000ull) < nan_ppm) { vals[r] = NaN; continue; }
        const uint64_t h = rh(r, 0);
        const int x = (int)(h & 0xffff) + (int)((h >> 16) & 0xffff) + (int)((h >> 32) & 0xffff) + (int)(h >> 48);
        const double noise = (double)(x - 131070) * noise_scale;
        const double sm = (15.0 + sin_table[r]) + noise;
        if (value_kind == 0) { vals[r] = sm; continue; }
        double inc = sm > 0.0 ? sm : 0.0;
        if (value_kind == 2) inc = std::rint(inc);
        if (reset_period > 0 && r > 0 && (rh(r, 2) % (uint64_t)reset_period) == 0) v = inc; else v = v + inc;
        vals[r] = v;
      }
      auto& ser = *s->series[first + i];
      for (int r0 = 0; r0 < rows; r0 += rpc) {
        const int n = std::min(rpc, rows - r0);
        auto c = std::make_unique<Chunk>();
        c->ts = enc::timestamps(ts.data() + r0, n);
        if (value_enc == 1) c->val = enc::doublesXor(vals.data() + r0, n, cumulative != 0);
        else if (value_enc == 0) { c->val = enc::rawDoubles(vals.data() + r0, n); if (cumulative && enc::counterDropFlag(vals.data() + r0, n)) pv_markDrop(c->val.data()); }
        else c->val = enc::doubles(vals.data() + r0, n, cumulative != 0);
        finishChunk(*c, ts[r0], ts[r0 + n - 1], n, ts[r0 + n - 1] + 1000);
        ser.chunks.push_back(std::move(c));
      }
    }
  };
  if (nthreads <= 1) worker();
  else { std::vector<std::thread> th; for (int i = 0; i < nthreads; ++i) th.emplace_back(worker); for (auto& t : th) t.join(); }
  return (int64_t)s->series.size();
}
int32_t fo_synth_group_ids(uint64_t seed, int64_t gid_base, int64_t n, int32_t n_groups, int32_t* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = n_groups > 0 ? (int32_t)(splitmix64(seed ^ 0xA5A5A5A5ull ^ ((uint64_t)(gid_base + i) * 0x9E3779B97F4A7C15ull)) % (uint64_t)n_groups) : 0;
  return 0;
}
int32_t fo_store_num_chunks(void* sp, int64_t series) { return (int32_t)((Store*)sp)->series[series]->chunks.size(); }
// ChunkSetInfo addresses of a series (what RawDataRangeVector.chunkInfos yields), for the product's filo_load_series.
void fo_store_info_addrs(void* sp, int64_t series, uint64_t* out) {
  auto& ch = ((Store*)sp)->series[series]->chunks;
  for (size_t i = 0; i < ch.size(); ++i) out[i] = (uint64_t)(uintptr_t)ch[i]->info.data();
}
int64_t fo_store_vector_bytes(void* sp, int64_t series, int32_t chunk, int32_t col, uint8_t* out, int64_t cap) {
  auto& c = *((Store*)sp)->series[series]->chunks[chunk];
  auto& b = col == 0 ? c.ts : c.val;
  if ((int64_t)b.size() > cap) return -(int64_t)b.size();
  std::memcpy(out, b.data(), b.size()); return (int64_t)b.size();
}
// Algorithmic bytes per SURVEY §8(d): Σ chunks (28 + 8*ncols + totalBytes(ts) + totalBytes(val))
int64_t fo_store_algorithmic_bytes(void* sp) {
  Store* s = (Store*)sp; int64_t b = 0;
  for (auto& se : s->series) for (auto& c : se->chunks)
    b += 28 + 16 + (int64_t)totalBytes(csi::vectorPtr(c->info.data(), 0)) + (int64_t)totalBytes(csi::vectorPtr(c->info.data(), 1));
  return b;
}

// ---------------- query: PeriodicSamplesMapper (+ optional AggregateMapReduce) over the store
// out_values: AGG_NONE -> [S*T]; else [G*T] (topk: [G*T*k]).  out_aux: avg counts [G*T] / topk ids [G*T*k] (may be null).
// stats: {samplesScanned, bytesScanned, elapsed_ns}
// schemaFlags: bit 0 = cumulative temporality (detectDrops column), bit 1 = the value column is a LongColumn.  p0, p1: static function
// arguments (quantile; sf, tf; duration).
int32_t fo_query2(void* sp, int32_t fn, int32_t schemaFlags, double p0, double p1, int64_t start, int64_t step, int64_t end, int64_t window,
                  int32_t inclusiveRange, int32_t aggrOp, int32_t k, const int32_t* groupIds, int32_t nGroups,
                  int32_t nThreads, int64_t seriesBegin, int64_t seriesEnd,
                  double* out_values, int64_t* out_aux, int64_t* stats);
int32_t fo_query(void* sp, int32_t fn, int32_t cumulative, int64_t start, int64_t step, int64_t end, int64_t window,
                 int32_t inclusiveRange, int32_t aggrOp, int32_t k, const int32_t* groupIds, int32_t nGroups,
                 int32_t nThreads, int64_t seriesBegin, int64_t seriesEnd,
                 double* out_values, int64_t* out_aux, int64_t* stats) {
  return fo_query2(sp, fn, cumulative ? 1 : 0, 0.0, 0.0, start, step, end, window, inclusiveRange, aggrOp, k, groupIds, nGroups, nThreads,
                   seriesBegin, seriesEnd, out_values, out_aux, stats);
}
int32_t fo_query2(void* sp, int32_t fn, int32_t schemaFlags, double p0, double p1, int64_t start, int64_t step, int64_t end, int64_t window,
                  int32_t inclusiveRange, int32_t aggrOp, int32_t k, const int32_t* groupIds, int32_t nGroups,
                  int32_t nThreads, int64_t seriesBegin, int64_t seriesEnd,
                  double* out_values, int64_t* out_aux, int64_t* stats) {
  Store* s = (Store*)sp;
  const int32_t cumulative = schemaFlags & 1;
  const bool longCol = (schemaFlags & 2) != 0;
  try {
    if (fn == FN_HOLT_WINTERS && !(p0 >= 0 && p0 <= 1 && p1 >= 0 && p1 <= 1)) throw std::invalid_argument("Sf/tf should be in between 0 and 1");
    if (seriesEnd < 0) seriesEnd = (int64_t)s->series.size();
    int64_t S = seriesEnd - seriesBegin;
    int T = numWindows(start, step, end);
    QueryConfig cfg; cfg.inclusiveRange = inclusiveRange != 0;
    std::vector<double> tmp;
    double* perSeries = out_values;
    if (aggrOp != AGG_NONE) { tmp.resize((size_t)S * T); perSeries = tmp.data(); }
    std::atomic<int64_t> next{0}; std::atomic<int64_t> samples{0}, bytes{0};
    std::string err; std::atomic<bool> failed{false};
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
      QueryStats qs;
      try {
        for (;;) {
          int64_t i0 = next.fetch_add(64);
          if (i0 >= S) break;
          int64_t i1 = std::min<int64_t>(S, i0 + 64);
          for (int64_t i = i0; i < i1; ++i) {
            Series se; for (auto& c : s->series[seriesBegin + i]->chunks) se.infos.push_back(c->info.data());
            se.longCol = longCol;
            periodicSamples(se, (RangeFn)fn, cumulative != 0, start, step, end, window, cfg, perSeries + (size_t)i * T, &qs, p0, p1);
          }
        }
      } catch (std::exception& e) { if (!failed.exchange(true)) err = e.what(); }
      samples += qs.samplesScanned; bytes += qs.bytesScanned;
    };
    if (nThreads <= 1) worker();
    else { std::vector<std::thread> th; for (int i = 0; i < nThreads; ++i) th.emplace_back(worker); for (auto& t : th) t.join(); }
    if (failed) { g_err = err; return -2; }
    if (aggrOp != AGG_NONE) {
      std::vector<const double*> rows((size_t)S);
      std::vector<int32_t> groups((size_t)S, 0);
      for (int64_t i = 0; i < S; ++i) { rows[i] = perSeries + (size_t)i * T; if (groupIds) groups[i] = groupIds[seriesBegin + i]; }
      AggResult r = aggregate((AggrOp)aggrOp, k, rows, groups, nGroups, T);
      std::memcpy(out_values, r.values.data(), r.values.size() * 8);
      if (out_aux && !r.aux.empty()) std::memcpy(out_aux, r.aux.data(), r.aux.size() * 8);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) { stats[0] = samples; stats[1] = bytes; stats[2] = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count(); }
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int32_t fo_num_windows(int64_t start, int64_t step, int64_t end) { return numWindows(start, step, end); }

// ---------------- result wire format (SerializedRangeVector over one shared RecordBuilder)
// values [nRows][T] -> containers (4096 bytes each, concatenated; returns their count or -needed when cap is too small) + per range vector
// numRowsSerialized / startRecordNo / firstContainer
int64_t fo_serialize_result(const double* values, int64_t nRows, int32_t T, int64_t start, int64_t step, int64_t end, int64_t nowMs,
                            uint8_t* outContainers, int64_t capContainers, int32_t* rowsSerialized, int32_t* startRecordNo, int64_t* firstContainer) {
  try {
    result::RecordBuilder b(result::MaxContainerSize, nowMs);
    for (int64_t i = 0; i < nRows; ++i) {
      const result::SerializedRangeVector srv = result::serialize(b, values + (size_t)i * T, T, start, step, end);
      rowsSerialized[i] = srv.numRowsSerialized; startRecordNo[i] = srv.startRecordNo; firstContainer[i] = srv.firstContainer;
    }
    const int64_t n = (int64_t)b.containers.size();
    if (n > capContainers) return -n;
    for (int64_t c = 0; c < n; ++c) std::memcpy(outContainers + c * result::MaxContainerSize, b.containers[(size_t)c]->bytes.data(), result::MaxContainerSize);
    return n;
  } catch (std::exception& e) { g_err = e.what(); return INT64_MIN; }
}
// SerializedRangeVector.rows of one range vector: fills ts / vals (cap entries), returns the row count
int64_t fo_result_rows(const uint8_t* containers, int64_t nContainers, int32_t rowsSerialized, int32_t startRecordNo, int64_t firstContainer,
                       int64_t start, int64_t step, int64_t end, int64_t* ts, double* vals, int64_t cap) {
  result::SerializedRangeVector srv; srv.numRowsSerialized = rowsSerialized; srv.startRecordNo = startRecordNo; srv.firstContainer = firstContainer;
  std::vector<int64_t> t; std::vector<double> v;
  result::rows(containers, nContainers, srv, start, step, end, t, v);
  const int64_t n = (int64_t)t.size();
  for (int64_t i = 0; i < n && i < cap; ++i) { ts[i] = t[(size_t)i]; vals[i] = v[(size_t)i]; }
  return n;
}

// Row-wise sliding cross-check for one series given raw rows.
void fo_sliding(const int64_t* ts, const double* vals, int64_t n, int32_t fn, int32_t cumulative,
                int64_t start, int64_t step, int64_t end, int64_t window, double* out) {
  std::vector<int64_t> t(ts, ts + n); std::vector<double> v(vals, vals + n);
  slidingSamples(t, v, (RangeFn)fn, cumulative != 0, start, step, end, window, out);
}

} // extern "C"
