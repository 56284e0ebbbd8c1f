// Histogram column scan, second version (fused sum of hist rate / increase over cumulative SectDelta histograms): the kernel that
// strings the phases of hist_phases.h together.  filo_query_hist selects it for the shapes it serves (capi.cu; FILO_HIST_V2=0 turns it
// off for A/B runs); the first version (hist_kernels.cu) serves every other shape.
#include "kernels.h"
#include "hist_phases.h"

namespace filo {

// Per-phase cycle counters for profiling builds (-DFILO_HIST_PROF; scratch/hist_prof.py): thread 0 of every CTA reads clock64() after each
// phase barrier.  Compiled out of the product build.
#if defined(FILO_HIST_PROF) && !defined(FILO_CUSIM)
__device__ unsigned long long g_hist2_prof[16];
#define H2PROF_DECL long long hp_t0 = clock64(), hp_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define H2PROF(i) { const long long hp_t1 = clock64(); hp_acc[i] += hp_t1 - hp_t0; hp_t0 = hp_t1; }
#define H2PROF_FLUSH if (threadIdx.x == 0) { for (int hp_i = 0; hp_i < 12; ++hp_i) atomicAdd(&g_hist2_prof[hp_i], (unsigned long long)hp_acc[hp_i]); atomicAdd(&g_hist2_prof[15], 1ull); }
#else
#define H2PROF_DECL
#define H2PROF(i)
#define H2PROF_FLUSH
#endif

__device__ __forceinline__ void h2_report(int* d_err, int code, int64_t sid) {
  if (atomicCAS(&d_err[0], 0, code) == 0) { d_err[1] = (int)(sid & 0x7fffffff); d_err[2] = (int)(sid >> 31); }
}

// record bytes global -> shared with 16-byte cp.async (LDGSTS); the caller waits with cp.async.wait_group + a CTA barrier
#ifdef FILO_CUSIM          // host emulation build (tests/cpp/cusim.h)
inline void h2_stage_async(uint8_t* dst, const uint8_t* src, uint32_t bytes, int tid) {
  for (uint32_t i = (uint32_t)tid * 16; i < bytes; i += H2_THREADS * 16) cusim::cp_async(dst + i, src + i, 16);
}
inline void h2_stage_wait() { cusim::cp_async_wait_all(); }
#else
__device__ __forceinline__ void h2_stage_async(uint8_t* dst, const uint8_t* __restrict__ src, uint32_t bytes, int tid) {
  const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
  for (uint32_t i = (uint32_t)tid * 16; i < bytes; i += H2_THREADS * 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + i), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void h2_stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// One CTA folds work items (runs of series of one group, positions index `order`) into the item's partial row
// pval[it][bucket][window] (bucket-major) and pany[it][window].
__global__ void __launch_bounds__(H2_THREADS, 2)
hist_scan2_kernel(const uint8_t* __restrict__ arena, const int64_t* __restrict__ rec_off, QueryParams q, int nb, int max_rows, uint32_t max_rec,
                  const int32_t* __restrict__ order, const int64_t* __restrict__ item_begin, int64_t n_items,
                  double* __restrict__ pval, uint8_t* __restrict__ pany, unsigned long long* d_counters, int* d_err) {
  extern __shared__ __align__(16) uint8_t smem[];
  H2Ctx X; h2_ctx_init(X, smem, h2_layout(max_rows, nb, max_rec), q, nb);
  const int tid = threadIdx.x;
  int64_t rows_scanned = 0, bytes_scanned = 0;
  H2PROF_DECL
  for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int64_t pb = item_begin[it], pe = item_begin[it + 1];
    double* pv = pval + (size_t)it * q.T * nb; uint8_t* pa = pany + (size_t)it * q.T;
    for (int i = tid; i < q.T * nb; i += H2_THREADS) pv[i] = 0.0;
    uint32_t anyb = 0;                                   // bit j: window tid + j * H2_THREADS has a histogram
    __syncthreads();                                     // the zero fill is visible to the owners of the windows
    bool prefetched = false;                             // the record of `pos` is already on its way (cp.async issued after the previous decode)
    for (int64_t pos = pb; pos < pe; ++pos) {
      const int64_t sid = order ? (int64_t)order[pos] : pos;
      if (!prefetched) {                                 // stage the record (16-byte aligned, size a multiple of 16)
        const int64_t ro = rec_off[sid];
        h2_stage_async(smem + X.L.rec, arena + ro, (uint32_t)(rec_off[sid + 1] - ro), tid);
      }
      h2_stage_wait();
      __syncthreads();
      H2PROF(0)                                           // record staged (prefetched behind the previous series)
      h2_tables(tid, X, max_rows);
      __syncthreads();
      H2PROF(1)                                           // chunk range + section table (thread 0)
      if (tid == 0) {
        const H2Ctl* C = X.ctl();
        rows_scanned += C->rows_scanned; bytes_scanned += C->bytes_scanned;
        if (C->err) h2_report(d_err, C->err, sid);
      }
      h2_decode_rows(tid, H2_THREADS, X);
      __syncthreads();
      H2PROF(2)                                           // timestamps + rows decoded
      // the staged record is dead from here on: fetch the next series' record behind the remaining phases
      prefetched = pos + 1 < pe;
      if (prefetched) {
        const int64_t nsid = order ? (int64_t)order[pos + 1] : pos + 1;
        const int64_t ro = rec_off[nsid];
        h2_stage_async(smem + X.L.rec, arena + ro, (uint32_t)(rec_off[nsid + 1] - ro), tid);
      }
      if (tid == 0 && X.ctl()->bad) h2_report(d_err, 1, sid);
      h2_add_base(tid, H2_THREADS, X);
      __syncthreads();
      H2PROF(3)                                           // next record issued, SectDelta bases added
      h2_chunk_corrections(tid, H2_THREADS, X);
      __syncthreads();
      h2_chunk_less(tid, X);
      __syncthreads();
      h2_carried(tid, H2_THREADS, X);
      __syncthreads();
      H2PROF(4)                                           // corrections inside and across chunks
      int j = 0;
      for (int k = tid; k < q.T; k += H2_THREADS, ++j) if (h2_window(k, X, pv, !((anyb >> j) & 1u))) anyb |= 1u << j;
      __syncthreads();                                   // the series' rows and record are dead
      H2PROF(5)                                           // windows: descriptors + rates + partial-row update
    }
    { int j = 0; for (int k = tid; k < q.T; k += H2_THREADS, ++j) pa[k] = (uint8_t)((anyb >> j) & 1u); }
  }
  H2PROF_FLUSH
  if (tid == 0 && (rows_scanned | bytes_scanned)) { atomicAdd(&d_counters[0], (unsigned long long)rows_scanned); atomicAdd(&d_counters[1], (unsigned long long)bytes_scanned); }
}

// Fold the partial rows of each group in item order (deterministic), MutableHistogram.add per item (Histogram.scala:428-449),
// then Histogram.quantile (:65-108, hist_quantile in hist_phases.h).  Thread per (group, window); partial rows are bucket-major.
__global__ void hist_merge2_kernel(const double* __restrict__ pval, const uint8_t* __restrict__ pany, const int64_t* __restrict__ gis,
                                   int n_groups, int T, int nb, int exp_buckets, const double* __restrict__ tops, double qtl,
                                   double* __restrict__ out_values /* [G][T][nb] or null */, double* __restrict__ out_q /* [G][T] or null */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_groups * T) return;
  const int g = (int)(i / T), k = (int)(i - (int64_t)g * T);
  const double NaNv = __longlong_as_double(0x7ff8000000000000LL);
  double v[64]; bool any = false;
  for (int b = 0; b < nb; ++b) v[b] = 0.0;
  // ReduceAggregateExec over the items' partial aggregates with the same reduceAggregate: the first one is copied, every further one
  // is added and the sum made monotonic (HistSumRowAggregator.scala:25-36, Histogram.scala:428-449)
  for (int64_t it = gis[g]; it < gis[g + 1]; ++it) {
    if (!pany[(size_t)it * T + k]) continue;
    const double* pv = pval + (size_t)it * T * nb + k;
    if (!any) { for (int b = 0; b < nb; ++b) v[b] = pv[(size_t)b * T]; any = true; continue; }
    double mx = 0.0;
    for (int b = 0; b < nb; ++b) { double nv = v[b] + pv[(size_t)b * T]; if (nv < mx || nv != nv) nv = mx; else if (nv > mx) mx = nv; v[b] = nv; }
  }
  const double qv = (any && qtl == qtl) ? hist_quantile(v, nb, tops, qtl, exp_buckets != 0) : NaNv;
  if (out_values) for (int b = 0; b < nb; ++b) out_values[(size_t)i * nb + b] = any ? v[b] : NaNv;
  if (out_q) out_q[i] = qv;
}

#ifndef FILO_CUSIM      // launchers need nvcc
#ifdef FILO_HIST_PROF
extern "C" int filo_debug_hist2_prof(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_hist2_prof, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset) { unsigned long long z[16] = {}; e = cudaMemcpyToSymbol(g_hist2_prof, z, sizeof z); }
  return (int)e;
}
#endif
size_t hist2_smem_bytes(int max_rows, int nb, uint32_t max_rec) { return h2_layout(max_rows, nb, max_rec).total; }
cudaError_t launch_hist_scan2(const ScanLaunch& L, int nb, int max_rows, uint32_t max_rec, const int32_t* order, const int64_t* item_begin, int64_t n_items,
                              double* pval, uint8_t* pany) {
  const size_t smem = h2_layout(max_rows, nb, max_rec).total;
  cudaError_t e = cudaFuncSetAttribute(hist_scan2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  hist_scan2_kernel<<<L.grid, H2_THREADS, smem, L.stream>>>(L.arena, L.rec_off, L.q, nb, max_rows, max_rec, order, item_begin, n_items, pval, pany, L.d_counters, L.d_err);
  return cudaGetLastError();
}
cudaError_t launch_hist_merge2(const double* pval, const uint8_t* pany, const int64_t* gis, int n_groups, int T, int nb, int exp_buckets, const double* tops, double q,
                               double* out_values, double* out_q, cudaStream_t s) {
  const int64_t n = (int64_t)n_groups * T;
  if (n <= 0) return cudaSuccess;
  hist_merge2_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pval, pany, gis, n_groups, T, nb, exp_buckets, tops, q, out_values, out_q);
  return cudaGetLastError();
}

#endif // FILO_CUSIM

} // namespace filo
