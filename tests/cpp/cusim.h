// cusim — a small SIMT emulator for running this repo's CUDA kernels on the CPU (TEST INFRASTRUCTURE ONLY).
//
// Every CUDA thread of a CTA is a ucontext fiber on ONE OS thread; CTAs run one after the other.  A fiber runs until it reaches a
// synchronising operation (CTA / named barrier, warp shuffle / ballot / syncwarp, mbarrier wait) and then yields to the scheduler,
// which picks the next runnable fiber round-robin or in a seeded pseudo-random order.  Bulk (TMA) copies are deferred: a load lands
// and a store reads its source only some scheduling steps after it was issued (or when the kernel waits for it), so code that reads
// or overwrites the buffers early sees stale data and fails the comparison with the oracle.  When every live fiber is blocked and
// nothing can complete, the run aborts with the reason each fiber waits for (a barrier some threads never reach, a ballot inside a
// short-circuit, an mbarrier phase nobody completes).
//
// What is emulated: threadIdx / blockIdx / blockDim / gridDim, dynamic shared memory (`extern __shared__ ... smem[]`), __syncthreads,
// bar.sync id,count, __syncwarp, __shfl*_sync, __ballot_sync, __any/__all_sync, mbarrier init / arrive / expect_tx / try_wait.parity,
// cp.async.bulk global<->shared with complete_tx / bulk groups, atomics, and the integer / floating-point intrinsics the kernels use.
// What is not: memory-model weak ordering, bank conflicts, timing.  It checks logic and synchronisation structure, not speed.
#pragma once
#include <cuda_runtime.h>      // vector types and host-side qualifier macros only
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <vector>

namespace cusim {

struct Fiber { ucontext_t ctx; std::vector<char> stack; uint3 tid{0, 0, 0}; bool done = false; const char* waiting = nullptr; int site = 0; };
struct NamedBar { int arrived = 0; unsigned gen = 0; };
struct WarpState { uint64_t slot[32]; int arrived = 0; unsigned gen = 0; unsigned mask = 0; int site = 0, first_lane = 0; };
struct Deferred { int countdown; bool is_load; void* dst; const void* src; size_t bytes; uint64_t* bar; int owner; bool done; };

struct Cta {
  std::vector<Fiber> fibers; dim3 bdim{1, 1, 1}, gdim{1, 1, 1}; uint3 bidx{0, 0, 0};
  NamedBar bars[16]; std::vector<WarpState> warps;
  std::vector<Deferred> ops;
  int cur = -1; ucontext_t sched; uint64_t progress = 0;
  std::function<void()> body;
};
inline Cta*& g() { static Cta* p = nullptr; return p; }
inline uint64_t& rng_state() { static uint64_t s = 0; return s; }        // 0: round-robin; else seeded xorshift schedule
inline int& tma_delay() { static int d = 1 << 30; return d; }            // fiber switches before a deferred bulk copy is performed; the default is adversarial:
                                                                         // a copy happens as late as the program allows (when its issuer waits for it, or when
                                                                         // every fiber is blocked)
inline Fiber& me() { return g()->fibers[(size_t)g()->cur]; }
inline int linear_tid() { const Fiber& f = me(); return (int)(f.tid.x + f.tid.y * g()->bdim.x + f.tid.z * g()->bdim.x * g()->bdim.y); }
inline void progress() { ++g()->progress; }
inline void yield(const char* why) { Fiber& f = me(); f.waiting = why; swapcontext(&f.ctx, &g()->sched); f.waiting = nullptr; }

// ---- mbarrier state packed into the kernel's own 8-byte word: [63] phase, [62:48] expected arrivals, [47:32] pending arrivals, [31:0] pending tx bytes
struct MbarView {
  uint64_t* w;
  unsigned phase() const { return (unsigned)(*w >> 63); }
  unsigned expected() const { return (unsigned)((*w >> 48) & 0x7fff); }
  unsigned pending() const { return (unsigned)((*w >> 32) & 0xffff); }
  int32_t tx() const { return (int32_t)(uint32_t)*w; }
  void set(unsigned ph, unsigned ex, unsigned pe, int32_t t) { *w = ((uint64_t)ph << 63) | ((uint64_t)ex << 48) | ((uint64_t)pe << 32) | (uint32_t)t; }
  void check() { if (pending() == 0 && tx() == 0) { set(phase() ^ 1u, expected(), expected(), 0); progress(); } }
};
inline void mbar_init(uint64_t* bar, uint32_t count) { MbarView{bar}.set(0, count, count, 0); }
inline void mbar_arrive(uint64_t* bar) { MbarView m{bar}; if (m.pending() == 0) { std::fprintf(stderr, "cusim: mbarrier over-arrival\n"); std::abort(); } m.set(m.phase(), m.expected(), m.pending() - 1, m.tx()); progress(); m.check(); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { MbarView m{bar}; m.set(m.phase(), m.expected(), m.pending(), m.tx() + (int32_t)bytes); mbar_arrive(bar); }
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) { MbarView m{bar}; m.set(m.phase(), m.expected(), m.pending(), m.tx() - (int32_t)bytes); progress(); m.check(); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { while (MbarView{bar}.phase() == (parity & 1u)) yield("mbarrier wait"); }

// ---- deferred bulk copies
inline void run_op(Deferred& o) {
  if (o.done) return;
  std::memcpy(o.dst, o.src, o.bytes); o.done = true; progress();
  if (o.is_load && o.bar) mbar_complete_tx(o.bar, (uint32_t)o.bytes);
}
inline void tick_ops() { for (auto& o : g()->ops) if (!o.done && --o.countdown <= 0) run_op(o); }
inline void tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) { g()->ops.push_back(Deferred{tma_delay(), true, smem_dst, gsrc, bytes, bar, g()->cur, false}); }
inline void tma_store(void* gdst, const void* smem_src, uint32_t bytes) { g()->ops.push_back(Deferred{tma_delay(), false, gdst, smem_src, bytes, nullptr, g()->cur, false}); }
// cp.async (LDGSTS) global -> shared: lands when the issuing thread waits for its groups (as late as the program allows)
inline void cp_async(void* smem_dst, const void* gsrc, uint32_t bytes) { g()->ops.push_back(Deferred{tma_delay(), false, smem_dst, gsrc, bytes, nullptr, g()->cur, false}); }
inline void cp_async_wait_all() { for (auto& o : g()->ops) if (!o.is_load && o.owner == g()->cur) run_op(o); }
inline void tma_store_wait_read() { for (auto& o : g()->ops) if (!o.is_load && o.owner == g()->cur) run_op(o); }     // the fiber's bulk stores have read their source

// ---- CTA-level barriers
inline void bar_sync(int id, int count) {
  NamedBar& b = g()->bars[id];
  const unsigned gen = b.gen;
  progress();
  if (++b.arrived == count) { b.arrived = 0; ++b.gen; } else while (b.gen == gen) yield(id == 0 ? "__syncthreads" : "named barrier");
}
// ---- warp collectives (one mask in flight per warp; lanes of a warp must agree on it)
inline WarpState& my_warp() { return g()->warps[(size_t)(linear_tid() >> 5)]; }
inline void warp_barrier(unsigned mask) {
  WarpState& W = my_warp();
  const int lane = linear_tid() & 31;
  if (!((mask >> lane) & 1u)) { std::fprintf(stderr, "cusim: lane %d calls a warp collective whose mask %08x excludes it\n", lane, mask); std::abort(); }
  if (W.arrived == 0) { W.mask = mask; W.site = me().site; W.first_lane = lane; }
  else if (W.mask != mask) { std::fprintf(stderr, "cusim: warp collective with differing masks %08x vs %08x\n", W.mask, mask); std::abort(); }
  else if (W.site != me().site) {          // the lanes named in the mask must execute the SAME collective (convergence requirement of *_sync)
    std::fprintf(stderr, "cusim: DIVERGENT warp collective in block %u warp %d: lane %d is at source line %d while lane %d is at line %d\n",
                 g()->bidx.x, linear_tid() >> 5, W.first_lane, W.site, lane, me().site);
    std::abort();
  }
  const unsigned gen = W.gen;
  progress();
  if (++W.arrived == __builtin_popcount(mask)) { W.arrived = 0; ++W.gen; } else while (W.gen == gen) yield("warp collective");
}
template <class T> inline T warp_exchange(unsigned mask, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  WarpState& W = my_warp(); const int lane = linear_tid() & 31;
  uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T)); W.slot[lane] = raw;
  warp_barrier(mask);
  T r = v;
  if (src >= 0 && src < 32 && ((mask >> src) & 1u)) std::memcpy(&r, &W.slot[src], sizeof(T));
  warp_barrier(mask);
  return r;
}
inline unsigned warp_ballot(unsigned mask, int pred) {
  WarpState& W = my_warp(); const int lane = linear_tid() & 31;
  W.slot[lane] = pred ? 1u : 0u;
  warp_barrier(mask);
  unsigned r = 0; for (int l = 0; l < 32; ++l) if (((mask >> l) & 1u) && W.slot[l]) r |= 1u << l;
  warp_barrier(mask);
  return r;
}

// ---- running a grid
inline void trampoline() { g()->body(); me().done = true; progress(); swapcontext(&me().ctx, &g()->sched); }
inline void report_deadlock() {
  std::fprintf(stderr, "cusim: DEADLOCK in block (%u,%u,%u): no fiber can make progress\n", g()->bidx.x, g()->bidx.y, g()->bidx.z);
  for (size_t w = 0; w < g()->warps.size(); ++w) {            // per warp: how many lanes wait for what
    const char* reasons[8]; int counts[8]; int n = 0, done = 0;
    for (size_t l = 0; l < 32 && w * 32 + l < g()->fibers.size(); ++l) {
      const Fiber& f = g()->fibers[w * 32 + l];
      if (f.done) { ++done; continue; }
      const char* r = f.waiting ? f.waiting : "?";
      int k = 0; while (k < n && reasons[k] != r) ++k;
      if (k == n && n < 8) { reasons[n] = r; counts[n] = 0; ++n; }
      if (k < 8) ++counts[k];
    }
    std::fprintf(stderr, "  warp %zu:", w);
    if (done) std::fprintf(stderr, " %d lanes finished;", done);
    for (int k = 0; k < n; ++k) std::fprintf(stderr, " %d lanes in %s;", counts[k], reasons[k]);
    for (size_t l = 0; l < 32 && w * 32 + l < g()->fibers.size(); ++l) {
      const Fiber& f = g()->fibers[w * 32 + l];
      if (!f.done && f.waiting && !std::strcmp(f.waiting, "warp collective")) std::fprintf(stderr, " [lane %zu: collective at source line %d]", l, f.site);
    }
    std::fprintf(stderr, "\n");
  }
  std::abort();
}
// launch(grid, block, body): body is the kernel call; it runs once per CUDA thread
inline void launch(dim3 grid, dim3 block, const std::function<void()>& body, size_t stack_bytes = 256 * 1024) {
  const int nthreads = (int)(block.x * block.y * block.z);
  for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
    Cta cta; cta.bdim = block; cta.gdim = grid; cta.bidx = uint3{bx, by, bz}; cta.body = body;
    cta.fibers.resize((size_t)nthreads); cta.warps.resize((size_t)(nthreads + 31) / 32);
    g() = &cta;
    for (int t = 0; t < nthreads; ++t) {
      Fiber& f = cta.fibers[(size_t)t];
      f.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
      f.stack.resize(stack_bytes);
      getcontext(&f.ctx); f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = f.stack.size(); f.ctx.uc_link = &cta.sched;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    // rounds: every live fiber gets one turn per round, in index order (seed 0) or in a fresh pseudo-random permutation
    int live = nthreads, idle_rounds = 0; uint64_t last = cta.progress;
    uint64_t& rs = rng_state();
    std::vector<int> perm((size_t)nthreads); for (int t = 0; t < nthreads; ++t) perm[(size_t)t] = t;
    while (live > 0) {
      if (rs) for (int t = nthreads - 1; t > 0; --t) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; std::swap(perm[(size_t)t], perm[(size_t)(rs % (uint64_t)(t + 1))]); }
      for (int t = 0; t < nthreads; ++t) {
        const int pick = perm[(size_t)t];
        if (cta.fibers[(size_t)pick].done) continue;
        cta.cur = pick;
        swapcontext(&cta.sched, &cta.fibers[(size_t)pick].ctx);
        if (cta.fibers[(size_t)pick].done) --live;
        tick_ops();
      }
      if (cta.progress != last) { last = cta.progress; idle_rounds = 0; }
      else if (++idle_rounds > 1) {
        bool pending = false; for (auto& o : cta.ops) pending |= !o.done && o.is_load;
        if (pending) { for (auto& o : cta.ops) if (o.is_load) run_op(o); idle_rounds = 0; } else report_deadlock();     // stores stay pending: nobody can be blocked on them
      }
    }
    for (auto& o : cta.ops) run_op(o);                   // bulk stores still in flight at kernel end complete
    g() = nullptr;
  }
}

} // namespace cusim

// ------------------------------------------------------------------------------------------------ CUDA surface for the kernels
#define threadIdx (cusim::me().tid)
#define blockIdx (cusim::g()->bidx)
#define blockDim (cusim::g()->bdim)
#define gridDim (cusim::g()->gdim)
// dynamic shared memory: the kernels' block-scope `extern __shared__ ... smem[]` names <enclosing namespace>::smem; the harness defines it
// (e.g. `namespace filo { alignas(128) uint8_t smem[232448]; }`) — one CTA runs at a time, so one buffer serves them all

inline void __syncthreads() { cusim::bar_sync(0, (int)(blockDim.x * blockDim.y * blockDim.z)); }
namespace cusim {
inline void at(int line) { me().site = line; }
inline void syncwarp(unsigned mask = 0xffffffffu) { warp_barrier(mask); }
template <class T> inline T shfl(unsigned mask, T v, int src, int width = 32) { const int lane = linear_tid() & 31; return warp_exchange(mask, v, (lane & ~(width - 1)) + (src & (width - 1))); }
template <class T> inline T shfl_up(unsigned mask, T v, unsigned d, int width = 32) { const int lane = linear_tid() & 31; const int src = lane - (int)d; return warp_exchange(mask, v, src < (lane & ~(width - 1)) ? -1 : src); }
template <class T> inline T shfl_down(unsigned mask, T v, unsigned d, int width = 32) { const int lane = linear_tid() & 31; const int src = lane + (int)d; return warp_exchange(mask, v, src > (lane | (width - 1)) ? -1 : src); }
template <class T> inline T shfl_xor(unsigned mask, T v, int x, int width = 32) { const int lane = linear_tid() & 31; (void)width; return warp_exchange(mask, v, lane ^ x); }
inline unsigned ballot(unsigned mask, int pred) { return warp_ballot(mask, pred); }
inline int any(unsigned mask, int pred) { return warp_ballot(mask, pred) != 0; }
inline int all(unsigned mask, int pred) { return (warp_ballot(mask, pred) & mask) == mask; }
}
// the *_sync collectives record their source line: lanes of one collective must come from the same call site
#define __syncwarp(...) (cusim::at(__LINE__), cusim::syncwarp(__VA_ARGS__))
#define __shfl_sync(...) (cusim::at(__LINE__), cusim::shfl(__VA_ARGS__))
#define __shfl_up_sync(...) (cusim::at(__LINE__), cusim::shfl_up(__VA_ARGS__))
#define __shfl_down_sync(...) (cusim::at(__LINE__), cusim::shfl_down(__VA_ARGS__))
#define __shfl_xor_sync(...) (cusim::at(__LINE__), cusim::shfl_xor(__VA_ARGS__))
#define __ballot_sync(...) (cusim::at(__LINE__), cusim::ballot(__VA_ARGS__))
#define __any_sync(...) (cusim::at(__LINE__), cusim::any(__VA_ARGS__))
#define __all_sync(...) (cusim::at(__LINE__), cusim::all(__VA_ARGS__))
inline unsigned __activemask() { return 0xffffffffu; }

#ifdef __launch_bounds__
#undef __launch_bounds__
#endif
#define __launch_bounds__(...)
inline unsigned __double2uint_rz(double d) { return d != d || d <= 0.0 ? 0u : d >= 4294967295.0 ? 0xffffffffu : (unsigned)d; }
inline int __double2int_rz(double d) { return d != d ? 0 : d >= 2147483647.0 ? INT32_MAX : d <= -2147483648.0 ? INT32_MIN : (int)d; }
inline long long __double2ll_rz(double d) { return d != d ? 0 : d >= 9.2233720368547758e18 ? INT64_MAX : d <= -9.2233720368547758e18 ? INT64_MIN : (long long)d; }
inline long long __double2ll_rd(double d) { return __double2ll_rz(std::floor(d)); }
inline long long __double2ll_ru(double d) { return __double2ll_rz(std::ceil(d)); }
inline long long __double2ll_rn(double d) { return __double2ll_rz(std::nearbyint(d)); }
inline int __double2int_rn(double d) { return __double2int_rz(std::nearbyint(d)); }
using std::isinf; using std::isnan;
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { const uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { const uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline long long __mul64hi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
inline int __double2hiint(double d) { uint64_t b; std::memcpy(&b, &d, 8); return (int)(b >> 32); }
inline int __double2loint(double d) { uint64_t b; std::memcpy(&b, &d, 8); return (int)(uint32_t)b; }
inline double __hiloint2double(int hi, int lo) { const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; std::memcpy(&d, &b, 8); return d; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __drcp_rn(double a) { return 1.0 / a; }
inline double __int2double_rn(int v) { return (double)v; }
inline double __ll2double_rn(long long v) { return (double)v; }
inline double __uint2double_rn(unsigned v) { return (double)v; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T* p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { const T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { const T o = *p; *p = o | v; return o; }
template <class T> inline T atomicExch(T* p, T v) { const T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { const T o = *p; if (o == cmp) *p = v; return o; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) { cusim::yield("nanosleep"); }
