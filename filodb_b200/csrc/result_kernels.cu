// Result wire format on the device: the (timestamp, double) rows of a query result encoded as BinaryRecord v2 records inside
// RecordContainers, the bytes SerializedRangeVector.apply produces with one shared RecordBuilder for all range vectors of a result.
//
// Reference (paths under /root/reference):
//   SerializedRangeVector.apply / canRemoveEmptyRows   core/src/main/scala/filodb.core/query/RangeVector.scala:511-575  (NaN rows are not
//                                                      encoded unless the query is an instant query, start == end)
//   RecordBuilder.startNewRecord / addLong / addDouble / endRecord / requireBytes / newContainer
//                                                      core/src/main/scala/filodb.core/binaryrecord2/RecordBuilder.scala:109-175,461-480,589-621
//   RecordContainer header                             core/src/main/scala/filodb.core/binaryrecord2/RecordContainer.scala:13-57
//   RecordSchema offsets                               core/src/main/scala/filodb.core/binaryrecord2/RecordSchema.scala:65-71
// Layout: container = [+0 i32 numBytes (bytes after this word)] [+4 i32 version word = 1 << 24] [+8 i64 server timestamp] records...;
// record of the (Timestamp, Double) schema = [+0 i32 16] [+4 i64 timestamp] [+12 f64 value]: 20 bytes, 4-byte aligned.  A container of
// MaxContainerSize = 4096 bytes takes 204 records (a record that does not fit opens the next container, RecordBuilder.scala:589-606).
// Range vector i of the result is described by (numRowsSerialized, startRecordNo, first container): its records are records
// [startRecordNo, startRecordNo + numRowsSerialized) of the concatenation of the containers from its first container on
// (RangeVector.scala:427-476); startRecordNo is the record count of the builder's current container when the vector started.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cub/device/device_scan.cuh>
#include "../../include/filo_b200.h"

struct filo_ctx;
cudaStream_t filo_internal_stream(filo_ctx* ctx);
int filo_internal_device(filo_ctx* ctx);
int32_t filo_internal_fail(filo_ctx* ctx, int32_t code, const char* msg);

namespace filo {

constexpr int RC_CONTAINER_BYTES = 4096;      // SerializedRangeVector.MaxContainerSize
constexpr int RC_HEADER = 16;                 // RecordBuilder.ContainerHeaderLen
constexpr int RC_RECORD = 20;                 // length word + Long + Double
constexpr int RC_PER_CONTAINER = (RC_CONTAINER_BYTES - RC_HEADER) / RC_RECORD;      // 204

// rows kept per range vector: warp per row of the result matrix
__global__ void rc_count_kernel(const double* __restrict__ vals, int64_t n_rows, int T, int keep_nan, int64_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const double* v = vals + (size_t)row * T;
  int c = 0;
  for (int k = lane; k < T; k += 32) { const double x = v[k]; c += (keep_nan || x == x) ? 1 : 0; }
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[row] = c;
}

__device__ __forceinline__ uint8_t* rc_record_ptr(uint8_t* containers, int64_t r) {
  const int64_t c = r / RC_PER_CONTAINER; const int i = (int)(r - c * RC_PER_CONTAINER);
  return containers + c * RC_CONTAINER_BYTES + RC_HEADER + i * RC_RECORD;
}

// records: warp per range vector, ballot compaction keeps the rows in window order
__global__ void rc_write_kernel(const double* __restrict__ vals, int64_t n_rows, int T, int keep_nan, int64_t start, int64_t step,
                                const int64_t* __restrict__ first_rec /* exclusive scan of cnt, [n_rows + 1] */, uint8_t* __restrict__ containers,
                                int32_t* __restrict__ rows_serialized, int32_t* __restrict__ start_record_no, int64_t* __restrict__ first_container) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const double* v = vals + (size_t)row * T;
  const int64_t r0 = first_rec[row];
  int64_t r = r0;
  for (int k0 = 0; k0 < T; k0 += 32) {
    const int k = k0 + lane;
    const double x = k < T ? v[k] : 0.0;
    const bool keep = k < T && (keep_nan || x == x);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      uint32_t* p = reinterpret_cast<uint32_t*>(rc_record_ptr(containers, r + __popc(m & ((1u << lane) - 1u))));      // 4-byte aligned
      const int64_t ts = start + (int64_t)k * step;
      const uint64_t xb = (uint64_t)__double_as_longlong(x);
      p[0] = 16u;                                            // schema.variableAreaStart - 4
      p[1] = (uint32_t)(uint64_t)ts; p[2] = (uint32_t)((uint64_t)ts >> 32);
      p[3] = (uint32_t)xb; p[4] = (uint32_t)(xb >> 32);
    }
    r += __popc(m);
  }
  if (lane == 0) {
    rows_serialized[row] = (int32_t)(r - r0);
    // builder.currentContainer at the start of this vector: none before the first record of the result, else the container of record r0 - 1
    const int64_t cur = r0 == 0 ? 0 : (r0 - 1) / RC_PER_CONTAINER;
    first_container[row] = cur;
    start_record_no[row] = (int32_t)(r0 - cur * RC_PER_CONTAINER);
  }
}

__global__ void rc_header_kernel(uint8_t* __restrict__ containers, int64_t n_containers, int64_t n_records, int64_t ts_ms) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_containers) return;
  int64_t n = n_records - c * RC_PER_CONTAINER; if (n > RC_PER_CONTAINER) n = RC_PER_CONTAINER; if (n < 0) n = 0;
  uint32_t* h = reinterpret_cast<uint32_t*>(containers + c * RC_CONTAINER_BYTES);
  h[0] = (uint32_t)(RC_HEADER - 4 + n * RC_RECORD);          // RecordContainer.updateLengthWithOffset
  h[1] = 1u << 24;                                           // writeVersionWord: RecordBuilder.Version << 24
  h[2] = (uint32_t)(uint64_t)ts_ms; h[3] = (uint32_t)((uint64_t)ts_ms >> 32);
}

} // namespace filo

#define R_TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { std::string m = std::string(#call) + ": " + cudaGetErrorString(_e); \
  return filo_internal_fail(ctx, _e == cudaErrorMemoryAllocation ? FILO_ERR_OOM : FILO_ERR_CUDA, m.c_str()); } } while (0)

extern "C" int64_t filo_result_max_containers(int64_t n_rows, int32_t n_windows) {
  if (n_rows <= 0 || n_windows <= 0) return 0;
  const int64_t recs = n_rows * (int64_t)n_windows;
  return (recs + filo::RC_PER_CONTAINER - 1) / filo::RC_PER_CONTAINER;
}

extern "C" int32_t filo_encode_result_device(filo_ctx* ctx, const void* d_values, int64_t n_rows, int64_t start_ms, int64_t step_ms, int64_t end_ms,
                                             int64_t container_ts_ms, void* d_containers, int64_t containers_cap_bytes,
                                             void* d_rows_serialized, void* d_start_record_no, void* d_first_container,
                                             int64_t* n_containers_out, int64_t* n_records_out, void* cuda_stream) {
  if (!ctx || !d_values || !d_containers || !d_rows_serialized || !d_start_record_no || !d_first_container || !n_containers_out)
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_result_device: null argument");
  if (n_rows < 0 || start_ms > end_ms || !(start_ms == end_ms || step_ms > 0)) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_result_device: bad range");
  R_TRY(cudaSetDevice(filo_internal_device(ctx)));
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : filo_internal_stream(ctx);
  const int T = filo_num_windows(start_ms, step_ms > 0 ? step_ms : step_ms + 1, end_ms);
  const int keep_nan = start_ms == end_ms ? 1 : 0;           // canRemoveEmptyRows: instant queries keep every row
  *n_containers_out = 0; if (n_records_out) *n_records_out = 0;
  if (n_rows == 0) return FILO_OK;
  int64_t* d_cnt = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
  R_TRY(cudaMallocAsync((void**)&d_cnt, (size_t)(n_rows + 1) * 8 * 2, s));
  int64_t* d_first = d_cnt + (n_rows + 1);
  struct Free { cudaStream_t s; void* a; void** b; ~Free() { cudaFreeAsync(a, s); if (*b) cudaFreeAsync(*b, s); } } guard{s, d_cnt, &d_tmp};
  R_TRY(cudaMemsetAsync(d_cnt + n_rows, 0, 8, s));
  const unsigned blocks = (unsigned)((n_rows * 32 + 255) / 256);
  filo::rc_count_kernel<<<blocks, 256, 0, s>>>((const double*)d_values, n_rows, T, keep_nan, d_cnt);
  R_TRY(cudaGetLastError());
  R_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_first, n_rows + 1, s));
  R_TRY(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 16, s));
  R_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_first, n_rows + 1, s));
  int64_t n_records = 0;
  R_TRY(cudaMemcpyAsync(&n_records, d_first + n_rows, 8, cudaMemcpyDeviceToHost, s));
  R_TRY(cudaStreamSynchronize(s));
  const int64_t n_containers = (n_records + filo::RC_PER_CONTAINER - 1) / filo::RC_PER_CONTAINER;
  if (n_containers * filo::RC_CONTAINER_BYTES > containers_cap_bytes)
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_result_device: container buffer too small (filo_result_max_containers * 4096 bytes always suffice)");
  filo::rc_write_kernel<<<blocks, 256, 0, s>>>((const double*)d_values, n_rows, T, keep_nan, start_ms, step_ms > 0 ? step_ms : step_ms + 1, d_first,
                                               (uint8_t*)d_containers, (int32_t*)d_rows_serialized, (int32_t*)d_start_record_no, (int64_t*)d_first_container);
  R_TRY(cudaGetLastError());
  if (n_containers) {
    filo::rc_header_kernel<<<(unsigned)((n_containers + 127) / 128), 128, 0, s>>>((uint8_t*)d_containers, n_containers, n_records, container_ts_ms);
    R_TRY(cudaGetLastError());
  }
  *n_containers_out = n_containers; if (n_records_out) *n_records_out = n_records;
  return FILO_OK;
}

// host convenience: values in, container bytes and the per-vector descriptors out (tests, callers without their own device buffers)
extern "C" int32_t filo_encode_result(filo_ctx* ctx, const double* values, int64_t n_rows, int64_t start_ms, int64_t step_ms, int64_t end_ms, int64_t container_ts_ms,
                                      uint8_t* out_containers, int64_t containers_cap_bytes, int32_t* rows_serialized, int32_t* start_record_no,
                                      int64_t* first_container, int64_t* n_containers_out, int64_t* n_records_out) {
  if (!ctx || !values || !out_containers || !rows_serialized || !start_record_no || !first_container || !n_containers_out)
    return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_result: null argument");
  if (n_rows < 0 || start_ms > end_ms || !(start_ms == end_ms || step_ms > 0)) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_result: bad range");
  R_TRY(cudaSetDevice(filo_internal_device(ctx)));
  cudaStream_t s = filo_internal_stream(ctx);
  const int T = filo_num_windows(start_ms, step_ms > 0 ? step_ms : step_ms + 1, end_ms);
  *n_containers_out = 0; if (n_records_out) *n_records_out = 0;
  if (n_rows == 0) return FILO_OK;
  const size_t nv = (size_t)n_rows * T;
  const int64_t maxc = filo_result_max_containers(n_rows, T);
  uint8_t* d = nullptr;
  const size_t off_c = (nv * 8 + 255) & ~(size_t)255, off_rs = off_c + (size_t)maxc * filo::RC_CONTAINER_BYTES, off_sr = off_rs + (((size_t)n_rows * 4 + 255) & ~(size_t)255),
               off_fc = off_sr + (((size_t)n_rows * 4 + 255) & ~(size_t)255), total = off_fc + (size_t)n_rows * 8;
  R_TRY(cudaMallocAsync((void**)&d, total, s));
  struct Free { cudaStream_t s; void* p; ~Free() { cudaFreeAsync(p, s); } } guard{s, d};
  R_TRY(cudaMemcpyAsync(d, values, nv * 8, cudaMemcpyHostToDevice, s));
  int64_t nc = 0, nr = 0;
  const int32_t rc = filo_encode_result_device(ctx, d, n_rows, start_ms, step_ms, end_ms, container_ts_ms, d + off_c, maxc * filo::RC_CONTAINER_BYTES,
                                               d + off_rs, d + off_sr, d + off_fc, &nc, &nr, s);
  if (rc != FILO_OK) return rc;
  if (nc * filo::RC_CONTAINER_BYTES > containers_cap_bytes) return filo_internal_fail(ctx, FILO_ERR_INVALID_ARG, "filo_encode_result: container buffer too small");
  // the tail of the last container past its records is not part of the wire format: zero it for reproducible bytes
  if (nc) {
    const int64_t used = (nc - 1) * filo::RC_CONTAINER_BYTES + filo::RC_HEADER + (nr - (nc - 1) * filo::RC_PER_CONTAINER) * filo::RC_RECORD;
    R_TRY(cudaMemsetAsync(d + off_c + used, 0, (size_t)(nc * filo::RC_CONTAINER_BYTES - used), s));
    R_TRY(cudaMemcpyAsync(out_containers, d + off_c, (size_t)nc * filo::RC_CONTAINER_BYTES, cudaMemcpyDeviceToHost, s));
  }
  R_TRY(cudaMemcpyAsync(rows_serialized, d + off_rs, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, s));
  R_TRY(cudaMemcpyAsync(start_record_no, d + off_sr, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, s));
  R_TRY(cudaMemcpyAsync(first_container, d + off_fc, (size_t)n_rows * 8, cudaMemcpyDeviceToHost, s));
  R_TRY(cudaStreamSynchronize(s));
  *n_containers_out = nc; if (n_records_out) *n_records_out = nr;
  return FILO_OK;
}
