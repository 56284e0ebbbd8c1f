// CPU emulation of hist_scan2_kernel: runs the phase functions of filodb_b200/csrc/hist_phases.h (the kernel's own code, compiled for
// the host) thread id by thread id, with the oracle as the checker.  Test infrastructure: built and run by tests/test_abi.py.
//
// Series: cumulative SectDelta histogram chunks with counter resets inside chunks and at chunk boundaries, regular and jittered
// timestamps; queries: rate / increase over windows that span chunk boundaries.  A single-series item must equal the oracle's
// periodicSamplesHist bit for bit; a multi-series item must equal the oracle's values folded in series order.
#include "../../filodb_b200/csrc/hist_phases.h"
#include "../../oracle/filo_hist.hpp"
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>

namespace H = fo::hist;

struct Chunk { std::vector<uint8_t> ts, hv, info; };
struct Series { std::vector<std::unique_ptr<Chunk>> chunks; std::vector<uint8_t> record; };

static void build_series(Series& S, std::mt19937_64& rng, const H::Buckets& b, int rows, const std::vector<int>& chunk_rows, int64_t t0, int step_ms, int jitter,
                         int reset_every) {
  const int nb = b.n;
  std::vector<int64_t> ts((size_t)rows), vals((size_t)rows * nb);
  std::vector<int64_t> cur((size_t)nb, 0);
  std::vector<char> boundary((size_t)rows + 1, 0);
  { int r0 = 0; for (int n : chunk_rows) { r0 += n; if (r0 < rows) boundary[(size_t)r0] = 1; } }
  for (int r = 0; r < rows; ++r) {
    ts[(size_t)r] = t0 + (int64_t)r * step_ms + (jitter ? (int64_t)(rng() % (2 * jitter + 1)) - jitter : 0);
    if (reset_every && r > 0 && (rng() % reset_every == 0 || (boundary[(size_t)r] && rng() % 2))) std::fill(cur.begin(), cur.end(), 0);   // also at chunk starts
    int64_t acc = 0;
    std::vector<int64_t> inc((size_t)nb, 0);
    const int k = 1 + (int)(rng() % 3);
    for (int j = 0; j < k; ++j) inc[(size_t)(rng() % nb)] += 1 + (int64_t)(rng() % 5);
    for (int i = 0; i < nb; ++i) { acc += inc[(size_t)i]; cur[(size_t)i] += acc; }        // cumulative over buckets and over time
    for (int i = 0; i < nb; ++i) vals[(size_t)r * nb + i] = cur[(size_t)i];
  }
  int r0 = 0;
  for (int n : chunk_rows) {
    auto c = std::make_unique<Chunk>();
    c->ts = fo::enc::timestamps(ts.data() + r0, n);
    H::HistAppender app(true, 60000);
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
  // arena record (filo_record.h): header, chunk entries, vectors verbatim (8-byte aligned), padded to 16
  const size_t nch = S.chunks.size();
  size_t off = sizeof(filo::RecordHeader) + nch * sizeof(filo::ChunkEntry);
  std::vector<filo::ChunkEntry> E(nch);
  std::vector<uint8_t> body;
  uint32_t row_base = 0;
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

// one series through the phases, exactly as hist_scan2_kernel strings them together
static int run_series(const Series& S, const filo::H2Ctx& X, int max_rows, double* pv, std::vector<uint8_t>& any) {
  const int NT = filo::H2_THREADS;
  std::memcpy(X.smem + X.L.rec, S.record.data(), S.record.size());
  for (int t = 0; t < NT; ++t) filo::h2_tables(t, X, max_rows);
  if (X.ctl()->err) return X.ctl()->err;
  for (int t = 0; t < NT; ++t) filo::h2_decode_rows(t, NT, X);
  if (X.ctl()->bad) return 1;
  for (int t = 0; t < NT; ++t) filo::h2_add_base(t, NT, X);
  for (int t = 0; t < NT; ++t) filo::h2_chunk_corrections(t, NT, X);
  for (int t = 0; t < NT; ++t) filo::h2_chunk_less(t, X);
  for (int t = 0; t < NT; ++t) filo::h2_carried(t, NT, X);
  for (int t = 0; t < NT; ++t) for (int k = t; k < X.q.T; k += NT) if (filo::h2_window(k, X, pv, any[(size_t)k] == 0)) any[(size_t)k] = 1;
  return 0;
}

static bool same_bits(double a, double b) { uint64_t x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8); return x == y || (a != a && b != b); }

int main(int argc, char** argv) {
  std::mt19937_64 rng(argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 777);      // other seeds: other chunk layouts and histories
  long checked = 0, drops = 0, less = 0, empties = 0; int cases = 0;
  for (int cfg = 0; cfg < 27; ++cfg) {
    const int nb = cfg >= 24 ? (cfg == 24 ? 1 : cfg == 25 ? 2 : 64) : cfg % 3 == 0 ? 20 : (cfg % 3 == 1 ? 8 : 33);     // the last three: the smallest and the largest bucket counts
    std::vector<double> les; for (int i = 0; i < nb - 1; ++i) les.push_back(2.0 * std::pow(3.0, i)); les.push_back(INFINITY);
    const H::Buckets b = cfg % 2 ? H::Buckets::geometric(2.0, 2.0, nb) : H::Buckets::custom(les.data(), nb);
    const int rows = cfg < 4 ? 480 : 60 + (int)(rng() % 200);
    std::vector<int> chunk_rows;
    if (cfg < 4) chunk_rows = {400, 80};
    else { int left = rows; while (left > 0) { const int n = std::min(left, 1 + (int)(rng() % 90)); chunk_rows.push_back(n); left -= n; if (chunk_rows.size() == 7) { chunk_rows.back() += left; left = 0; } } }
    const int step_ms = 15000, jitter = (cfg % 4 == 3) ? 4000 : 0, reset_every = (cfg % 2) ? 37 : 0;
    const int64_t t0 = 1700000000000LL;
    const int nser = 1 + cfg % 3;
    std::vector<Series> SS((size_t)nser);
    uint32_t max_rec = 0;
    for (auto& S : SS) { build_series(S, rng, b, rows, chunk_rows, t0, step_ms, jitter, reset_every); max_rec = std::max<uint32_t>(max_rec, (uint32_t)S.record.size()); }
    for (int qi = 0; qi < 3; ++qi) {
      filo::QueryParams q{};
      q.window = qi == 0 ? 300000 : (qi == 1 ? 60000 : 1000000);
      q.step = qi == 2 ? 47000 : 15000;
      q.start = t0 + (qi == 1 ? 5 * 15000 : -30000); q.end = t0 + (int64_t)rows * step_ms + 60000;
      q.T = (int)((q.end - q.start) / q.step) + 1;
      q.fn = qi == 1 ? filo::FN_INCREASE : filo::FN_RATE; q.cumulative = 1; q.inclusive = qi != 2;
      const filo::H2Layout L = filo::h2_layout(rows, nb, max_rec);
      std::vector<uint64_t> backing(L.total / 8 + 4, 0);
      filo::H2Ctx X; filo::h2_ctx_init(X, reinterpret_cast<uint8_t*>(backing.data()), L, q, nb);
      std::vector<double> pv((size_t)q.T * nb, 0.0), ref((size_t)q.T * nb, 0.0); std::vector<uint8_t> any((size_t)q.T, 0), rany((size_t)q.T, 0);
      for (auto& S : SS) {
        const int err = run_series(S, X, rows, pv.data(), any);
        if (err) { std::printf("FAIL cfg %d q %d: phase error %d\n", cfg, qi, err); return 1; }
        for (int c = 0; c < X.ctl()->n; ++c) { drops += X.ctl()->ch[c].has_drop; less += X.ctl()->less[c]; }
        H::HistSeries hs; for (auto& c : S.chunks) hs.infos.push_back(c->info.data());
        std::vector<H::MutHist> out;
        H::periodicSamplesHist(hs, q.fn == filo::FN_RATE ? fo::FN_RATE : fo::FN_INCREASE, true, q.start, q.step, q.end, q.window, q.inclusive != 0, out);
        if ((int)out.size() != q.T) { std::printf("FAIL cfg %d q %d: T %d vs %zu\n", cfg, qi, q.T, out.size()); return 1; }
        for (int k = 0; k < q.T; ++k) {
          if (out[(size_t)k].numBuckets() == 0) continue;
          const bool firstm = !rany[(size_t)k];                    // HistSumRowAggregator: copy the first, MutableHistogram.add (sum + makeMonotonic) the others
          rany[(size_t)k] = 1;
          double mx = 0.0;
          for (int i = 0; i < nb; ++i) {
            double nv = ref[(size_t)i * q.T + k] + out[(size_t)k].values[(size_t)i];
            if (!firstm) { if (nv < mx || nv != nv) nv = mx; else if (nv > mx) mx = nv; }
            ref[(size_t)i * q.T + k] = nv;
          }
        }
        // scan counters of the series
        int64_t rows_in = 0; for (int n : chunk_rows) rows_in += n;
        if (X.ctl()->rows_scanned > rows_in) { std::printf("FAIL cfg %d: rows_scanned\n", cfg); return 1; }
      }
      for (int k = 0; k < q.T; ++k) {
        empties += !rany[(size_t)k];
        if (any[(size_t)k] != rany[(size_t)k]) { std::printf("FAIL cfg %d q %d window %d: any %d vs %d\n", cfg, qi, k, any[(size_t)k], rany[(size_t)k]); return 1; }
        for (int i = 0; i < nb; ++i) {
          const double a = pv[(size_t)i * q.T + k], r = ref[(size_t)i * q.T + k];
          if (!same_bits(a, r)) { std::printf("FAIL cfg %d q %d window %d bucket %d: %.17g vs %.17g\n", cfg, qi, k, i, a, r); return 1; }
          ++checked;
        }
      }
      ++cases;
    }
  }
  std::printf("OK %d cases, %ld values bit-exact (%ld chunks with drop sections, %ld chunk-boundary resets, %ld empty windows)\n", cases, checked, drops, less, empties);
  return 0;
}
