// Small host helpers for the loader (no device code).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace filo {

inline int host_threads() {
  if (const char* e = std::getenv("FILO_HOST_THREADS")) { int n = std::atoi(e); if (n > 0) return n; }
  unsigned hc = std::thread::hardware_concurrency();
  return (int)std::min<unsigned>(std::max<unsigned>(hc, 1), 64);
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

// Persistent worker pool (the loader runs many short gather phases; spawning threads per phase costs more than the phase).
// run(n, fn): fn(worker, begin, end) over [0, n) in contiguous blocks, returns when all blocks are done.  One run at a time.
class HostPool {
 public:
  explicit HostPool(int n) : n_(std::max(1, n)) {
    for (int i = 1; i < n_; ++i) th_.emplace_back([this, i] { worker(i); });
  }
  ~HostPool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  template <class F>
  void run(int64_t n, F&& fn) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> one(run_mu_);
    const int used = (int)std::min<int64_t>(n_, std::max<int64_t>(1, n / 16));
    if (used <= 1) { fn(0, (int64_t)0, n); return; }
    const int64_t per = (n + used - 1) / used;
    std::function<void(int)> job = [&](int w) { const int64_t b = w * per, e = std::min<int64_t>(n, b + per); if (b < e) fn(w, b, e); };
    { std::lock_guard<std::mutex> g(m_); job_ = &job; used_ = used; pending_ = used - 1; ++gen_; }
    cv_.notify_all();
    job(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
    job_ = nullptr;
  }
 private:
  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (id >= used_) continue;
        job = job_;
      }
      (*job)(id);
      { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
    }
  }
  int n_; std::vector<std::thread> th_;
  std::mutex m_, run_mu_; std::condition_variable cv_, done_;
  std::function<void(int)>* job_ = nullptr; int used_ = 0, pending_ = 0; uint64_t gen_ = 0; bool stop_ = false;
};
inline HostPool& host_pool() { static HostPool pool(host_threads()); return pool; }

} // namespace filo
