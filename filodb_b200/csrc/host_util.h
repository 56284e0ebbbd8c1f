// Small host helpers for the loader (no device code).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>
#include <algorithm>

namespace filo {

inline int host_threads() {
  if (const char* e = std::getenv("FILO_HOST_THREADS")) { int n = std::atoi(e); if (n > 0) return n; }
  unsigned hc = std::thread::hardware_concurrency();
  return (int)std::min<unsigned>(std::max<unsigned>(hc, 1), 32);
}

// fn(thread_id, begin, end) over [0, n) in contiguous blocks
template <class F>
inline void parallel_for(int64_t n, int nthreads, F&& fn) {
  if (n <= 0) return;
  nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(1, n / 64));
  if (nthreads <= 1) { fn(0, (int64_t)0, n); return; }
  std::vector<std::thread> th;
  const int64_t per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    const int64_t b = t * per, e = std::min<int64_t>(n, b + per);
    if (b >= e) break;
    th.emplace_back([&fn, t, b, e]() { fn(t, b, e); });
  }
  for (auto& x : th) x.join();
}

} // namespace filo
