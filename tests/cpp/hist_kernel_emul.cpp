// The histogram kernels on the CPU: hist_scan2_kernel + hist_merge2_kernel (filodb_b200/csrc/hist_kernels2.cu) and hist_scan_kernel +
// hist_merge_kernel (hist_kernels.cu, through tests/cpp/make_cusim_src.py) compiled for the host on the cusim emulator and checked
// against the oracle: fused sum + quantile on both kernels, per-series rate / increase, sum_over_time and delta-temporality rate.
// Test infrastructure: built and run by tests/test_abi.py.     hist_kernel_emul [schedule seed]
#define FILO_CUSIM 1
#include "cusim.h"
namespace filo { alignas(128) uint8_t smem[232448]; }
#include "../../filodb_b200/csrc/hist_kernels2.cu"
#include HIST_V1_SRC                                             // hist_kernels.cu with function-scope __shared__ turned into static
#include "../../oracle/filo_hist.hpp"
#include <memory>
#include <random>

namespace H = fo::hist;
struct Chunk { std::vector<uint8_t> ts, hv, info; };
struct Series { std::vector<std::unique_ptr<Chunk>> chunks; std::vector<uint8_t> record; };

static void build_series(Series& S, std::mt19937_64& rng, const H::Buckets& b, int rows, const std::vector<int>& chunk_rows, int64_t t0, int step_ms, int jitter,
                         int reset_every, bool sect, bool cumulative) {
  const int nb = b.n;
  std::vector<int64_t> ts((size_t)rows), vals((size_t)rows * nb), cur((size_t)nb, 0);
  std::vector<char> boundary((size_t)rows + 1, 0);
  { int r0 = 0; for (int n : chunk_rows) { r0 += n; if (r0 < rows) boundary[(size_t)r0] = 1; } }
  for (int r = 0; r < rows; ++r) {
    ts[(size_t)r] = t0 + (int64_t)r * step_ms + (jitter ? (int64_t)(rng() % (uint64_t)(2 * jitter + 1)) - jitter : 0);
    if (!cumulative) std::fill(cur.begin(), cur.end(), 0);        // delta temporality: every row stands alone
    else if (reset_every && r > 0 && (rng() % (uint64_t)reset_every == 0 || (boundary[(size_t)r] && rng() % 2))) std::fill(cur.begin(), cur.end(), 0);
    std::vector<int64_t> inc((size_t)nb, 0);
    const int k = 1 + (int)(rng() % 3);
    for (int j = 0; j < k; ++j) inc[(size_t)(rng() % (uint64_t)nb)] += 1 + (int64_t)(rng() % 5);
    int64_t acc = 0;
    for (int i = 0; i < nb; ++i) { acc += inc[(size_t)i]; cur[(size_t)i] += acc; vals[(size_t)r * nb + i] = cur[(size_t)i]; }
  }
  int r0 = 0;
  for (int n : chunk_rows) {
    auto c = std::make_unique<Chunk>();
    c->ts = fo::enc::timestamps(ts.data() + r0, n);
    H::HistAppender app(sect, 60000);
    for (int r = 0; r < n; ++r) {
      std::vector<uint8_t> blob = H::bin::writeDelta(b, vals.data() + (size_t)(r0 + r) * nb, nb);
      if (app.addData(blob.data(), (int)blob.size()) != H::Ack) { std::printf("appender failed\n"); std::exit(2); }
    }
    c->hv = app.bytes();
    c->info.assign(fo::csi::OffsetVectors + 16, 0);
    fo::setLong(c->info.data() + fo::csi::OffsetChunkID, fo::csi::chunkID(ts[(size_t)r0], (ts[(size_t)(r0 + n - 1)] + 1000) / 1000));
    fo::setInt(c->info.data() + fo::csi::OffsetNumRows, n);
    fo::setLong(c->info.data() + fo::csi::OffsetIngestionTime, ts[(size_t)(r0 + n - 1)] + 1000);
    fo::setLong(c->info.data() + fo::csi::OffsetEndTime, ts[(size_t)(r0 + n - 1)]);
    fo::setLong(c->info.data() + fo::csi::OffsetVectors, (int64_t)(uintptr_t)c->ts.data());
    fo::setLong(c->info.data() + fo::csi::OffsetVectors + 8, (int64_t)(uintptr_t)c->hv.data());
    S.chunks.push_back(std::move(c));
    r0 += n;
  }
  const size_t nch = S.chunks.size(), off = sizeof(filo::RecordHeader) + nch * sizeof(filo::ChunkEntry);
  std::vector<filo::ChunkEntry> E(nch); std::vector<uint8_t> body; uint32_t row_base = 0;
  for (size_t i = 0; i < nch; ++i) {
    Chunk& c = *S.chunks[i];
    E[i].start_time = fo::csi::startTime(c.info.data()); E[i].end_time = fo::csi::endTime(c.info.data()); E[i].num_rows = fo::csi::numRows(c.info.data());
    auto put = [&](const std::vector<uint8_t>& v) { while ((off + body.size()) % 8) body.push_back(0); const uint32_t o = (uint32_t)(off + body.size()); body.insert(body.end(), v.begin(), v.end()); return o; };
    E[i].ts_off = put(c.ts); E[i].val_off = put(c.hv); E[i].row_base = row_base; row_base += (uint32_t)E[i].num_rows;
  }
  size_t total = off + body.size(); total = (total + 15) & ~(size_t)15;
  S.record.assign(total, 0);
  filo::RecordHeader h; h.rec_bytes = (uint32_t)total; h.n_chunks = (uint32_t)nch; h.n_rows = row_base; h.flags = filo::REC_HIST;
  std::memcpy(S.record.data(), &h, sizeof h);
  std::memcpy(S.record.data() + sizeof h, E.data(), nch * sizeof(filo::ChunkEntry));
  std::memcpy(S.record.data() + off, body.data(), body.size());
}
static bool same_bits(double a, double b) { uint64_t x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8); return x == y || (a != a && b != b); }

int main(int argc, char** argv) {
  cusim::rng_state() = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 0;
  std::mt19937_64 rng(99);
  long checked = 0; int cases = 0;
  struct Cfg { int nb; bool geometric; std::vector<int> chunks; int jitter, reset_every; bool sect, cumulative; int fn; int64_t window, step; int nser, seg; int inclusive; int exp_scale = -100; /* > -100: otel exponential buckets (scale, start index -5) */ };
  const std::vector<Cfg> cfgs = {
    {20, false, {400, 80}, 0, 0, true, true, filo::FN_RATE, 300000, 15000, 7, 3, 1},          // C4 shape
    {8, true, {70, 50, 40}, 3000, 37, true, true, filo::FN_INCREASE, 120000, 15000, 9, 4, 1},  // resets inside chunks and at chunk starts, irregular scrapes
    {33, false, {60, 90}, 0, 53, true, true, filo::FN_RATE, 450000, 47000, 5, 2, 0},           // 33 buckets (5 NibblePack groups), exclusive range start
    {12, true, {80, 40}, 0, 0, false, false, filo::FN_SUM, 300000, 15000, 6, 3, 1},            // delta temporality, simple (row) vectors: sum_over_time
    {12, true, {80, 40}, 2000, 0, false, false, filo::FN_RATE, 200000, 30000, 6, 6, 1},        // ... and rate = window sum / window length
    {16, false, {100, 60}, 0, 0, true, true, filo::FN_SUM, 300000, 15000, 4, 2, 1},            // sum_over_time over cumulative SectDelta vectors
    {12, false, {90, 70}, 0, 41, true, true, filo::FN_RATE, 300000, 15000, 6, 2, 1, 3},        // otel exponential buckets (scale 3) in SectDelta vectors: log-space quantile
    {24, false, {50, 50, 30}, 1500, 0, true, true, filo::FN_INCREASE, 200000, 20000, 5, 5, 1, -1},  // ... scale -1 (base 4)
    {1, true, {40, 30}, 0, 23, true, true, filo::FN_RATE, 120000, 15000, 4, 2, 1},               // one bucket (no quantile: NaN)
    {2, false, {40, 30}, 0, 0, true, true, filo::FN_INCREASE, 120000, 15000, 4, 2, 1},           // two buckets
    {64, true, {30, 25}, 1000, 19, true, true, filo::FN_RATE, 150000, 15000, 3, 3, 1},           // the largest table the device path takes
  };
  for (size_t ci = 0; ci < cfgs.size(); ++ci) {
    const Cfg& c = cfgs[ci];
    std::vector<double> les; for (int i = 0; i < c.nb - 1; ++i) les.push_back(2.0 * std::pow(3.0, i)); les.push_back(INFINITY);
    const H::Buckets b = c.exp_scale > -100 ? H::Buckets::exponential(c.exp_scale, -5, c.nb - 1) : c.geometric ? H::Buckets::geometric(2.0, 2.0, c.nb) : H::Buckets::custom(les.data(), c.nb);
    int rows = 0; for (int n : c.chunks) rows += n;
    const int64_t t0 = 1700000000000LL;
    std::vector<Series> SS((size_t)c.nser); std::vector<int64_t> rec_off((size_t)c.nser + 1, 0); uint32_t max_rec = 0;
    for (int s = 0; s < c.nser; ++s) { build_series(SS[(size_t)s], rng, b, rows, c.chunks, t0, 15000, c.jitter, c.reset_every, c.sect, c.cumulative);
                                       rec_off[(size_t)s + 1] = rec_off[(size_t)s] + (int64_t)SS[(size_t)s].record.size(); max_rec = std::max<uint32_t>(max_rec, (uint32_t)SS[(size_t)s].record.size()); }
    std::vector<uint64_t> backing((size_t)rec_off.back() / 8 + 64, 0); uint8_t* arena = reinterpret_cast<uint8_t*>(backing.data());
    for (int s = 0; s < c.nser; ++s) std::memcpy(arena + rec_off[(size_t)s], SS[(size_t)s].record.data(), SS[(size_t)s].record.size());
    filo::QueryParams q{};
    q.start = t0 - 30000; q.step = c.step; q.end = t0 + (int64_t)rows * 15000 + 45000; q.window = c.window; q.T = (int)((q.end - q.start) / q.step) + 1;
    q.fn = c.fn; q.cumulative = c.cumulative; q.inclusive = c.inclusive;
    const int T = q.T, nb = c.nb;
    // oracle per series
    const int ofn = c.fn == filo::FN_SUM ? fo::FN_SUM_OVER_TIME : c.fn;
    std::vector<std::vector<H::MutHist>> ref((size_t)c.nser);
    for (int s = 0; s < c.nser; ++s) { H::HistSeries hs; for (auto& ch : SS[(size_t)s].chunks) hs.infos.push_back(ch->info.data());
                                       H::periodicSamplesHist(hs, ofn, c.cumulative, q.start, q.step, q.end, q.window, c.inclusive != 0, ref[(size_t)s]); }
    // items: one group, runs of `seg` series in a shuffled order
    std::vector<int32_t> order((size_t)c.nser); for (int s = 0; s < c.nser; ++s) order[(size_t)s] = s; std::shuffle(order.begin(), order.end(), rng);
    std::vector<int64_t> item_begin; for (int64_t p = 0; p < c.nser; p += c.seg) item_begin.push_back(p); item_begin.push_back(c.nser);
    const int64_t n_items = (int64_t)item_begin.size() - 1; const int64_t gis[2] = {0, n_items};
    std::vector<double> tops((size_t)nb); for (int i = 0; i < nb; ++i) tops[(size_t)i] = b.bucketTop(i);
    // expected fused result: HistSumRowAggregator.reduceAggregate at both levels (series inside an item, items inside the group): the first
    // histogram is copied, every further one goes through MutableHistogram.add (sum, then makeMonotonic); then the quantile
    std::vector<double> exp_vals((size_t)T * nb, 0.0), exp_q((size_t)T, 0.0); std::vector<char> exp_any((size_t)T, 0);
    for (int k = 0; k < T; ++k) {
      H::MutHist tot; bool any = false;
      for (int64_t it = 0; it < n_items; ++it) {
        H::MutHist part; bool iany = false;
        for (int64_t p = item_begin[(size_t)it]; p < item_begin[(size_t)it + 1]; ++p) {
          const H::MutHist& h = ref[(size_t)order[(size_t)p]][(size_t)k];
          if (h.numBuckets() == 0) continue;
          if (!iany) { part = h; iany = true; }
          else { for (int i = 0; i < nb; ++i) part.values[(size_t)i] += h.values[(size_t)i]; part.makeMonotonic(); }
        }
        if (!iany) continue;
        if (!any) { tot = part; any = true; }
        else { for (int i = 0; i < nb; ++i) tot.values[(size_t)i] += part.values[(size_t)i]; tot.makeMonotonic(); }
      }
      exp_any[(size_t)k] = any;
      if (any) { exp_q[(size_t)k] = tot.quantile(0.9); for (int i = 0; i < nb; ++i) exp_vals[(size_t)k * nb + i] = tot.values[(size_t)i]; }
    }
    const bool counter_mode = c.cumulative && (c.fn == filo::FN_RATE || c.fn == filo::FN_INCREASE);
    unsigned long long counters[2]; int derr[4];
    auto check_fused = [&](const char* what, const std::vector<double>& ov, const std::vector<double>& oq) -> bool {
      for (int k = 0; k < T; ++k) {
        for (int i = 0; i < nb; ++i) { const double e = exp_any[(size_t)k] ? exp_vals[(size_t)k * nb + i] : std::nan(""); if (!same_bits(ov[(size_t)k * nb + i], e)) { std::printf("FAIL cfg %zu %s window %d bucket %d: %.17g vs %.17g\n", ci, what, k, i, ov[(size_t)k * nb + i], e); return false; } ++checked; }
        const double eq = exp_any[(size_t)k] ? exp_q[(size_t)k] : std::nan("");
        if (!same_bits(oq[(size_t)k], eq)) { std::printf("FAIL cfg %zu %s window %d quantile: %.17g vs %.17g\n", ci, what, k, oq[(size_t)k], eq); return false; }
      }
      return true;
    };
    // ---- second kernel (fused sum of rate / increase over cumulative histograms)
    if (counter_mode) {
      std::vector<double> pval((size_t)n_items * T * nb, -1.0), ov((size_t)T * nb, -1.0), oq((size_t)T, -1.0); std::vector<uint8_t> pany((size_t)n_items * T + 16, 7);
      counters[0] = counters[1] = 0; std::memset(derr, 0, sizeof derr);
      cusim::launch(dim3(2), dim3(filo::H2_THREADS), [&] { filo::hist_scan2_kernel(arena, rec_off.data(), q, nb, rows, max_rec, order.data(), item_begin.data(), n_items, pval.data(), pany.data(), counters, derr); });
      if (derr[0]) { std::printf("FAIL cfg %zu: v2 device error %d\n", ci, derr[0]); return 1; }
      cusim::launch(dim3((unsigned)((T + 127) / 128)), dim3(128), [&] { filo::hist_merge2_kernel(pval.data(), pany.data(), gis, 1, T, nb, b.kind == H::Buckets::EXP ? 1 : 0, tops.data(), 0.9, ov.data(), oq.data()); });
      if (!check_fused("v2", ov, oq)) return 1;
      if ((int64_t)counters[0] != (int64_t)c.nser * rows) { std::printf("FAIL cfg %zu: v2 samples_scanned %llu\n", ci, counters[0]); return 1; }
    }
    // ---- first kernel, fused
    {
      std::vector<double> pval((size_t)n_items * T * nb, -1.0), ov((size_t)T * nb, -1.0), oq((size_t)T, -1.0); std::vector<uint8_t> pany((size_t)n_items * T + 16, 7);
      counters[0] = counters[1] = 0; std::memset(derr, 0, sizeof derr);
      cusim::launch(dim3(2), dim3(filo::HIST_THREADS), [&] { filo::hist_scan_kernel(arena, rec_off.data(), c.nser, q, nb, rows, max_rec, order.data(), item_begin.data(), n_items, 1, nullptr, pval.data(), pany.data(), counters, derr); }, 128 * 1024);
      if (derr[0]) { std::printf("FAIL cfg %zu: v1 device error %d\n", ci, derr[0]); return 1; }
      cusim::launch(dim3((unsigned)((T + 127) / 128)), dim3(128), [&] { filo::hist_merge_kernel(pval.data(), pany.data(), gis, 1, T, nb, b.kind == H::Buckets::EXP ? 1 : 0, tops.data(), 0.9, ov.data(), oq.data()); });
      if (!check_fused("v1", ov, oq)) return 1;
    }
    // ---- first kernel, per series (NaN buckets = empty histogram)
    {
      std::vector<double> out((size_t)c.nser * T * nb, -1.0);
      counters[0] = counters[1] = 0; std::memset(derr, 0, sizeof derr);
      cusim::launch(dim3(3), dim3(filo::HIST_THREADS), [&] { filo::hist_scan_kernel(arena, rec_off.data(), c.nser, q, nb, rows, max_rec, nullptr, nullptr, 0, 0, out.data(), nullptr, nullptr, counters, derr); }, 128 * 1024);
      if (derr[0]) { std::printf("FAIL cfg %zu: v1 per-series device error %d\n", ci, derr[0]); return 1; }
      for (int s = 0; s < c.nser; ++s) for (int k = 0; k < T; ++k) {
        const H::MutHist& h = ref[(size_t)s][(size_t)k];
        for (int i = 0; i < nb; ++i) {
          const double e = h.numBuckets() ? h.values[(size_t)i] : std::nan(""), a = out[((size_t)s * T + k) * nb + i];
          if (!same_bits(a, e)) { std::printf("FAIL cfg %zu per-series s %d window %d bucket %d: %.17g vs %.17g\n", ci, s, k, i, a, e); return 1; }
          ++checked;
        }
      }
    }
    std::printf("cfg %zu ok (%s)\n", ci, counter_mode ? "v2 + v1 fused, v1 per series" : "v1 fused, v1 per series");
    ++cases;
  }
  std::printf("OK %d cases, %ld values bit-exact\n", cases, checked);
  return 0;
}
