// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over filo_hist.hpp (histogram column path).
#include "filo_hist.hpp"
#include <memory>
#include <string>

using namespace fo;
using namespace fo::hist;

namespace {
thread_local std::string h_err;
Buckets mkBuckets(int kind, double first, double mult, int minusOne, const double* les, int n) {
  if (kind == 1) return Buckets::geometric(first, mult, n, minusOne != 0);
  if (kind == 2) return Buckets::custom(les, n);
  if (kind == 3) return Buckets::exponential((int)first, (int)mult, n - 1);      // (scale, startIndexPositiveBuckets, numPositiveBuckets)
  return Buckets();
}
struct HChunk { std::vector<uint8_t> ts, hv, info; };
struct HSeries { std::vector<std::unique_ptr<HChunk>> chunks; };
struct HStore { std::vector<std::unique_ptr<HSeries>> series; };
}

extern "C" {
const char* fo_hist_last_error() { return h_err.c_str(); }

void fo_hist_bucket_tops(int kind, double first, double mult, int minusOne, const double* les, int n, double* out) {
  Buckets b = mkBuckets(kind, first, mult, minusOne, les, n);
  for (int i = 0; i < n; ++i) out[i] = b.bucketTop(i);
}
int32_t fo_hist_serialize_buckets(int kind, double first, double mult, int minusOne, const double* les, int n, uint8_t* out, int32_t cap) {
  try {
    std::vector<uint8_t> buf; const int end = mkBuckets(kind, first, mult, minusOne, les, n).serialize(buf, 0);
    if (end > cap) return -1;
    std::memcpy(out, buf.data(), (size_t)end); return end;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
// parse a serialized bucket definition (pointer to the u16 length prefix) back to tops
int32_t fo_hist_parse_buckets(const uint8_t* def, int32_t formatCode, double* tops, int32_t cap) {
  try {
    Buckets b = Buckets::parse(def, (uint8_t)formatCode);
    if (b.n > cap) return -1;
    for (int i = 0; i < b.n; ++i) tops[i] = b.bucketTop(i);
    return b.n;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_write_delta(int kind, double first, double mult, int minusOne, const double* les, int n, const int64_t* values, uint8_t* out, int32_t cap) {
  try {
    std::vector<uint8_t> b = bin::writeDelta(mkBuckets(kind, first, mult, minusOne, les, n), values, n);
    if ((int)b.size() > cap) return -1;
    std::memcpy(out, b.data(), b.size()); return (int32_t)b.size();
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_blob_to_values(const uint8_t* blob, int64_t* out, int32_t cap) {
  try { LongHist h = bin::toHistogram(blob); if (h.numBuckets() > cap) return -1; std::memcpy(out, h.values.data(), h.values.size() * 8); return h.numBuckets(); }
  catch (std::exception& e) { h_err = e.what(); return -2; }
}
double fo_hist_quantile(int kind, double first, double mult, int minusOne, const double* les, int n, const double* values, double q) {
  MutHist h; h.buckets = mkBuckets(kind, first, mult, minusOne, les, n); h.values.assign(values, values + n);
  return h.quantile(q);
}
void fo_hist_make_monotonic(double* values, int n) { MutHist h; h.buckets.n = n; h.values.assign(values, values + n); h.makeMonotonic(); std::memcpy(values, h.values.data(), (size_t)n * 8); }

// ---- appender
void* fo_hist_appender_new(int32_t sect, int32_t maxBytes) { return new HistAppender(sect == 1, maxBytes, sect == 2 /* exp vector */); }
void fo_hist_appender_free(void* a) { delete (HistAppender*)a; }
int32_t fo_hist_appender_add(void* a, const uint8_t* blob, int32_t len) {
  try { return ((HistAppender*)a)->addData(blob, len); } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_appender_length(void* a) { return ((HistAppender*)a)->length(); }
int32_t fo_hist_appender_bytes(void* a, uint8_t* out, int32_t cap) {
  std::vector<uint8_t> b = ((HistAppender*)a)->bytes();
  if ((int)b.size() > cap) return -1;
  std::memcpy(out, b.data(), b.size()); return (int32_t)b.size();
}
// ---- reader over vector bytes
int32_t fo_hist_vec_info(const uint8_t* vec, int32_t* length, int32_t* numBuckets, int32_t* sect) {
  try { HistReader r(vec); *length = r.length(); *numBuckets = r.nb; *sect = r.sect; return 0; } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_vec_apply(const uint8_t* vec, int32_t i, int64_t* out) {
  try { HistReader r(vec); LongHist h = r.apply(i); std::memcpy(out, h.values.data(), h.values.size() * 8); return 0; } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_vec_section_types(const uint8_t* vec, int32_t* out, int32_t cap) {
  try { HistReader r(vec); auto t = r.sectionTypes(); if ((int)t.size() > cap) return -1; for (size_t i = 0; i < t.size(); ++i) out[i] = t[i]; return (int32_t)t.size(); }
  catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_vec_sum(const uint8_t* vec, int32_t start, int32_t end, double* out) {
  try { HistReader r(vec); MutHist s = r.sum(start, end); std::memcpy(out, s.values.data(), s.values.size() * 8); return 0; } catch (std::exception& e) { h_err = e.what(); return -2; }
}
// ---- exponential schemes: scheme out = (scale, startIndexPositiveBuckets, numBuckets)
int32_t fo_hist_vec_apply_exp(const uint8_t* vec, int32_t i, int32_t* scheme, int64_t* out, int32_t cap) {
  try {
    HistReader r(vec); LongHist h = r.apply(i);
    if (h.numBuckets() > cap) return -1;
    scheme[0] = h.buckets.scale; scheme[1] = h.buckets.startIdx; scheme[2] = h.buckets.n;
    std::memcpy(out, h.values.data(), h.values.size() * 8); return h.numBuckets();
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_vec_sum_exp(const uint8_t* vec, int32_t start, int32_t end, int32_t* scheme, double* out, int32_t cap) {
  try {
    HistReader r(vec); MutHist s = r.sum(start, end);
    if (s.numBuckets() > cap) return -1;
    scheme[0] = s.buckets.scale; scheme[1] = s.buckets.startIdx; scheme[2] = s.buckets.n;
    std::memcpy(out, s.values.data(), s.values.size() * 8); return s.numBuckets();
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
// Base2ExpHistogramBuckets.add + canAccommodate; a/b/out = (scale, start, numBuckets)
int32_t fo_hist_exp_add_scheme(const int32_t* a, const int32_t* b, int32_t maxPos, int32_t* out) {
  try {
    Buckets x = Buckets::exponential(a[0], a[1], a[2] - 1), y = Buckets::exponential(b[0], b[1], b[2] - 1);
    Buckets r = x.expAdd(y, maxPos);
    out[0] = r.scale; out[1] = r.startIdx; out[2] = r.n;
    return (x.canAccommodate(y) ? 1 : 0) | (y.canAccommodate(x) ? 2 : 0);
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
// addValues: fold `other` (scheme b) into values laid out for scheme a
int32_t fo_hist_exp_add_values(const int32_t* a, double* ourValues, const int32_t* b, const double* otherValues) {
  try {
    Buckets x = Buckets::exponential(a[0], a[1], a[2] - 1);
    MutHist o; o.buckets = Buckets::exponential(b[0], b[1], b[2] - 1); o.values.assign(otherValues, otherValues + b[2]);
    std::vector<double> v(ourValues, ourValues + a[2]);
    x.expAddValues(v, o.buckets, o);
    std::memcpy(ourValues, v.data(), v.size() * 8); return 0;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
// MutableHistogram.addNoCorrection for two exponential histograms; result scheme/values in a / out (cap values)
int32_t fo_hist_exp_add_no_correction(int32_t* a, const double* aValues, const int32_t* b, const double* bValues, double* out, int32_t cap) {
  try {
    MutHist x; x.buckets = Buckets::exponential(a[0], a[1], a[2] - 1); x.values.assign(aValues, aValues + a[2]);
    MutHist y; y.buckets = Buckets::exponential(b[0], b[1], b[2] - 1); y.values.assign(bValues, bValues + b[2]);
    x.addNoCorrection(y);
    if (x.numBuckets() > cap) return -1;
    a[0] = x.buckets.scale; a[1] = x.buckets.startIdx; a[2] = x.buckets.n;
    std::memcpy(out, x.values.data(), x.values.size() * 8); return x.numBuckets();
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}

// correction API: meta = (has, lastValue[nb], correction[nb]); outputs written in place
int32_t fo_hist_vec_detect_drop(const uint8_t* vec, int32_t has, const int64_t* last, int64_t* corr_inout) {
  try {
    HistReader r(vec); HistCorrection m; m.some = has != 0;
    if (m.some) { m.lastValue.buckets = r.buckets; m.lastValue.values.assign(last, last + r.nb); m.correction.buckets = r.buckets; m.correction.values.assign(corr_inout, corr_inout + r.nb); }
    m = r.detectDropAndCorrection(m);
    if (m.some) std::memcpy(corr_inout, m.correction.values.data(), (size_t)r.nb * 8);
    return 0;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_vec_update_correction(const uint8_t* vec, int32_t has, const int64_t* corr_in, int64_t* out_last, int64_t* out_corr) {
  try {
    HistReader r(vec); HistCorrection m; m.some = has != 0;
    if (m.some) { m.lastValue = LongHist::empty(r.buckets); m.correction.buckets = r.buckets; m.correction.values.assign(corr_in, corr_in + r.nb); }
    HistCorrection o = r.updateCorrection(m);
    std::memcpy(out_last, o.lastValue.values.data(), (size_t)r.nb * 8); std::memcpy(out_corr, o.correction.values.data(), (size_t)r.nb * 8);
    return 0;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int32_t fo_hist_vec_corrected(const uint8_t* vec, int32_t n, int32_t has, const int64_t* corr_in, int64_t* out) {
  try {
    HistReader r(vec); HistCorrection m; m.some = has != 0;
    if (m.some) { m.lastValue = LongHist::empty(r.buckets); m.correction.buckets = r.buckets; m.correction.values.assign(corr_in, corr_in + r.nb); }
    LongHist h = r.correctedValue(n, m);
    std::memcpy(out, h.values.data(), (size_t)r.nb * 8);
    return 0;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}

// ---- store of histogram series + query
void* fo_hstore_new() { return new HStore(); }
void fo_hstore_free(void* s) { delete (HStore*)s; }
int64_t fo_hstore_add_series(void* sp) { HStore* s = (HStore*)sp; s->series.emplace_back(new HSeries()); return (int64_t)s->series.size() - 1; }
// one chunk: n rows, ts[n]; hist values row-major [n][nb] (cumulative bucket counts); encoded through BinaryHistogram.writeDelta +
// the (SectDelta | simple) appender exactly as ingestion does
int32_t fo_hstore_add_chunk(void* sp, int64_t series, const int64_t* ts, int32_t n, int kind, double first, double mult, int minusOne,
                            const double* les, int32_t nb, const int64_t* values, int32_t sect, int32_t maxBytes) {
  HStore* s = (HStore*)sp;
  try {
    auto c = std::make_unique<HChunk>();
    c->ts = enc::timestamps(ts, n);
    Buckets b = mkBuckets(kind, first, mult, minusOne, les, nb);
    HistAppender app(sect != 0, maxBytes);
    for (int r = 0; r < n; ++r) {
      std::vector<uint8_t> blob = bin::writeDelta(b, values + (size_t)r * nb, nb);
      const AddResponse res = app.addData(blob.data(), (int)blob.size());
      if (res != Ack) { h_err = "appender: response " + std::to_string((int)res) + " at row " + std::to_string(r); return -1; }
    }
    c->hv = app.bytes();
    c->info.assign(csi::OffsetVectors + 16, 0);
    setLong(c->info.data() + csi::OffsetChunkID, csi::chunkID(ts[0], (ts[n - 1] + 1000) / 1000));
    setInt(c->info.data() + csi::OffsetNumRows, n);
    setLong(c->info.data() + csi::OffsetIngestionTime, ts[n - 1] + 1000);
    setLong(c->info.data() + csi::OffsetEndTime, ts[n - 1]);
    setLong(c->info.data() + csi::OffsetVectors, (int64_t)(uintptr_t)c->ts.data());
    setLong(c->info.data() + csi::OffsetVectors + 8, (int64_t)(uintptr_t)c->hv.data());
    s->series[(size_t)series]->chunks.push_back(std::move(c));
    return 0;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
int64_t fo_hstore_num_chunks(void* sp, int64_t series) { return (int64_t)((HStore*)sp)->series[(size_t)series]->chunks.size(); }
void fo_hstore_info_addrs(void* sp, int64_t series, uint64_t* out) {
  auto& ch = ((HStore*)sp)->series[(size_t)series]->chunks;
  for (size_t i = 0; i < ch.size(); ++i) out[i] = (uint64_t)(uintptr_t)ch[i]->info.data();
}
int32_t fo_hstore_vector_bytes(void* sp, int64_t series, int64_t chunk, uint8_t* out, int32_t cap) {
  auto& hv = ((HStore*)sp)->series[(size_t)series]->chunks[(size_t)chunk]->hv;
  if ((int)hv.size() > cap) return -1;
  std::memcpy(out, hv.data(), hv.size()); return (int32_t)hv.size();
}
// PeriodicSamplesMapper over the histogram column (+ HistSumRowAggregator when aggr != 0, + histogram_quantile when q is not NaN).
//   aggr == 0: out_values [S][T][nb], out_empty [S][T] (1 = Histogram.empty)
//   aggr == 1: out_values [G][T][nb], out_empty [G][T]; out_quantile [G][T] when q is not NaN (NaN for empty histograms)
int32_t fo_hstore_query(void* sp, int32_t fn, int32_t cumulative, int64_t start, int64_t step, int64_t end, int64_t window, int32_t inclusive,
                        int32_t aggr, const int32_t* group_ids, int32_t n_groups, int32_t nb, double q,
                        double* out_values, uint8_t* out_empty, double* out_quantile) {
  HStore* s = (HStore*)sp;
  try {
    const int64_t adjStep = step > 0 ? step : step + 1;
    const int T = (int)((end - start) / adjStep) + 1;
    const int64_t S = (int64_t)s->series.size();
    std::vector<MutHist> acc;
    if (aggr) acc.assign((size_t)n_groups * T, MutHist());
    std::vector<MutHist> per;
    for (int64_t i = 0; i < S; ++i) {
      HistSeries hs; for (auto& c : s->series[(size_t)i]->chunks) hs.infos.push_back(c->info.data());
      periodicSamplesHist(hs, fn, cumulative != 0, start, step, end, window, inclusive != 0, per);
      for (int k = 0; k < T; ++k) {
        if (aggr) { histSumReduce(acc[(size_t)(group_ids ? group_ids[i] : 0) * T + k], per[(size_t)k]); continue; }
        const MutHist& h = per[(size_t)k];
        out_empty[(size_t)i * T + k] = h.numBuckets() == 0;
        for (int b = 0; b < nb; ++b) out_values[((size_t)i * T + k) * nb + b] = b < h.numBuckets() ? h.values[(size_t)b] : NaN;
      }
    }
    if (aggr) {
      for (size_t g = 0; g < (size_t)n_groups * T; ++g) {
        const MutHist& h = acc[g];
        out_empty[g] = h.numBuckets() == 0;
        for (int b = 0; b < nb; ++b) out_values[g * nb + b] = b < h.numBuckets() ? h.values[(size_t)b] : NaN;
        if (out_quantile) out_quantile[g] = (q == q && h.numBuckets() > 0) ? h.quantile(q) : NaN;
      }
    }
    return 0;
  } catch (std::exception& e) { h_err = e.what(); return -2; }
}
} // extern "C"
