// Host-visible declarations of the CUDA launchers (scan_kernels.cu, synth_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "filo_record.h"

namespace filo {

constexpr int SCAN_WARPS = 4;          // warps (= series in flight) per CTA, v1 kernels
constexpr int FAST_WARPS = 4;          // v2 kernels
constexpr int FAST_MIN_CTAS = 4;       // register budget of the v2 kernels: 4 CTAs x 128 threads per SM (<= 128 regs/thread)
constexpr int CHUNK_DESC_BYTES = 144;  // sizeof(ChunkDesc), scan_device.cuh
constexpr int FILO_MAX_TOPK = 32;
enum { AGG_NONE = 0, AGG_SUM = 1, AGG_AVG = 2, AGG_MIN = 3, AGG_MAX = 4, AGG_COUNT = 5, AGG_TOPK = 6, AGG_BOTTOMK = 7 };

struct QueryParams;

} // namespace filo
#include "scan_params.h"
#include "scan_tile_layout.h"
namespace filo {

struct ScanLaunch {
  const uint8_t* arena; const int64_t* rec_off; int64_t n_series;
  QueryParams q;
  uint8_t* gscratch; uint32_t scratch_bytes; int use_smem;
  unsigned long long* d_counters; int* d_err;
  int grid; cudaStream_t stream;
  const int64_t* list = nullptr; const unsigned long long* list_count = nullptr;   // optional: process only these series
};

cudaError_t launch_scan_series(const ScanLaunch& L, double* out);
cudaError_t launch_scan_agg(const ScanLaunch& L, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                            double* pval, uint32_t* pcnt, uint32_t acc_bytes);
cudaError_t launch_scan_series_v2(const ScanLaunch& L, double* out, uint32_t rec_cap);
cudaError_t launch_scan_agg_v2(const ScanLaunch& L, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                               double* pval, uint32_t* pcnt, uint32_t acc_bytes, uint32_t rec_cap);
size_t v2_smem_per_warp(uint32_t rec_cap, uint32_t scratch_bytes, uint32_t acc_bytes);
cudaError_t launch_scan_tile(const ScanLaunch& L, double* out, const TileSmem& T, int64_t* fallback_list, unsigned long long* fallback_count);
cudaError_t launch_scan_tile_agg(const ScanLaunch& L, const TileSmem& T, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                                 double* pval, uint32_t* pcnt, int64_t* fallback_list, unsigned long long* fallback_count);
cudaError_t launch_merge_partials(const double* pval, const uint32_t* pcnt, const int64_t* gis, int n_groups, int T, int agg_op,
                                  int partial_out, double* out_val, int64_t* out_cnt, cudaStream_t s);
cudaError_t launch_present(int agg_op, int64_t n, const double* vals, const int64_t* cnts, double* out, cudaStream_t s);
cudaError_t launch_topk(const double* per_series, const int32_t* order, const int64_t* group_start, int n_groups, int T, int k, int bottom,
                        double* out_val, int64_t* out_id, cudaStream_t s);
struct WpSmem;
cudaError_t launch_scan_wp(const ScanLaunch& L, double* out, const WpSmem& W, int64_t* fallback_list, unsigned long long* fallback_count);
struct WpCtrSmem;
cudaError_t launch_scan_wp_ctr(const ScanLaunch& L, double* out, const WpCtrSmem& W, int64_t* fallback_list, unsigned long long* fallback_count);
cudaError_t launch_scan_wp_ctr_agg(const ScanLaunch& L, const WpCtrSmem& W, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg_op,
                                   double* pval, uint32_t* pcnt, int64_t* fallback_list, unsigned long long* fallback_count);
size_t hist_smem_bytes(int max_rows, int nb, int T, bool agg, uint32_t max_rec);
cudaError_t launch_hist_scan(const ScanLaunch& L, int nb, int max_rows, uint32_t max_rec, const int32_t* order, const int64_t* item_begin, int64_t n_items, int agg,
                             double* out, double* pval, uint8_t* pany);
cudaError_t launch_hist_merge(const double* pval, const uint8_t* pany, const int64_t* gis, int n_groups, int T, int nb, int exp_buckets, const double* tops, double q,
                              double* out_values, double* out_q, cudaStream_t s);
// second version of the histogram scan (hist_kernels2.cu): fused sum of rate / increase over cumulative SectDelta histograms
size_t hist2_smem_bytes(int max_rows, int nb, uint32_t max_rec);
cudaError_t launch_hist_scan2(const ScanLaunch& L, int nb, int max_rows, uint32_t max_rec, const int32_t* order, const int64_t* item_begin, int64_t n_items,
                              double* pval, uint8_t* pany);
cudaError_t launch_hist_merge2(const double* pval, const uint8_t* pany, const int64_t* gis, int n_groups, int T, int nb, int exp_buckets, const double* tops, double q,
                               double* out_values, double* out_q, cudaStream_t s);
cudaError_t launch_iota(int32_t* a, int64_t n, cudaStream_t s);
cudaError_t launch_group_bounds(const int32_t* sorted_keys, int64_t n, int n_groups, int64_t* group_start, cudaStream_t s);
cudaError_t launch_group_item_count(const int64_t* group_start, int n_groups, int seg, int64_t* cnt, cudaStream_t s);
cudaError_t launch_fill_items(const int64_t* group_start, const int64_t* gis, int n_groups, int seg, int64_t n_items, int64_t n_series,
                              int64_t* item_begin, cudaStream_t s);

} // namespace filo
