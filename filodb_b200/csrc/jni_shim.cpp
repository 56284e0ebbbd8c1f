// JNI shim over the C-ABI of include/filo_b200.h: the native half of `object filodb.gpu.FiloB200NativeMethods`.
//
// Conventions are the reference's own for its native code: Scala `object` methods export as Java_<pkg>_<Class>_00024_<method>
// (core/src/rust/filodb_core/src/simd_vectors.rs:164,186), arguments are primitives, primitive arrays and raw addresses
// (SimdNativeMethods.scala:15-95), a failure raises java.lang.RuntimeException and the call returns the type's default
// (jni_exec, core/src/rust/filodb_core/src/exec.rs:15-61, errors.rs:36-47).  No state lives here: handles are the C-ABI's
// opaque pointers carried as jlong.
//
// Built into filodb_b200/libfilo_b200_jni.so by filodb_b200/build.py against the minimal jni.h of csrc/jni_stub (no JDK in this
// image); with a JDK: g++ -DFILO_USE_SYSTEM_JNI -I$JAVA_HOME/include -I$JAVA_HOME/include/linux ...
#ifdef FILO_USE_SYSTEM_JNI
#include <jni.h>
#else
#include "jni_stub/jni.h"
#endif
#include <cstdio>
#include <cstdint>
#include "../../include/filo_b200.h"

namespace {
void throw_filo(JNIEnv* env, filo_ctx* ctx, int32_t rc) {
  char msg[512]; msg[0] = 0;
  filo_last_error(ctx, msg, (int32_t)sizeof msg);
  char buf[600];
  std::snprintf(buf, sizeof buf, "filo_b200 error %d: %s", (int)rc, msg);
  jclass cls = env->FindClass("java/lang/RuntimeException");
  if (cls) env->ThrowNew(cls, buf);
}
inline filo_ctx* C(jlong h) { return reinterpret_cast<filo_ctx*>((uintptr_t)h); }
inline filo_table* T(jlong h) { return reinterpret_cast<filo_table*>((uintptr_t)h); }
// primitive arrays pinned for the duration of one C-ABI call (no JNI call may be made in between)
struct Critical {
  JNIEnv* env; jarray a; void* p;
  Critical(JNIEnv* e, jarray arr) : env(e), a(arr), p(arr ? e->GetPrimitiveArrayCritical(arr, nullptr) : nullptr) {}
  ~Critical() { if (p) env->ReleasePrimitiveArrayCritical(a, p, JNI_ABORT); }
};
void put_stats(JNIEnv* env, jlongArray stats, const filo_stats& st) {
  if (!stats) return;
  const jlong s[6] = {(jlong)st.samples_scanned, (jlong)st.bytes_scanned, (jlong)st.kernel_ns, (jlong)st.kernel_launches, (jlong)st.h2d_bytes, (jlong)st.d2h_bytes};
  jsize n = env->GetArrayLength(stats); if (n > 6) n = 6;
  env->SetLongArrayRegion(stats, 0, n, s);
}
}  // namespace

extern "C" {

JNIEXPORT jlong JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_ctxCreate(JNIEnv* env, jobject, jint device, jboolean inclusiveRange, jlong minStepMs,
                                                                             jlong maxBytesPerQuery, jint maxGroups) {
  filo_cfg cfg{};
  cfg.inclusive_range = inclusiveRange ? 1 : 0; cfg.min_step_ms = minStepMs; cfg.max_data_per_shard_query = maxBytesPerQuery;
  cfg.group_by_cardinality_limit = maxGroups;
  filo_ctx* ctx = nullptr;
  const int32_t rc = filo_ctx_create(device, &cfg, &ctx);
  if (rc != FILO_OK) { throw_filo(env, nullptr, rc); return 0; }
  return (jlong)(uintptr_t)ctx;
}

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_ctxDestroy(JNIEnv*, jobject, jlong ctx) { filo_ctx_destroy(C(ctx)); }

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_ctxSetFnArgs(JNIEnv* env, jobject, jlong ctx, jdouble arg0, jdouble arg1) {
  const int32_t rc = filo_ctx_set_fn_args(C(ctx), arg0, arg1);
  if (rc != FILO_OK) throw_filo(env, C(ctx), rc);
}

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_ctxCheck(JNIEnv* env, jobject, jlong ctx) {
  const int32_t rc = filo_ctx_check(C(ctx));
  if (rc != FILO_OK) throw_filo(env, C(ctx), rc);
}

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_hostRegister(JNIEnv* env, jobject, jlong ctx, jlong base, jlong bytes) {
  const int32_t rc = filo_host_register(C(ctx), reinterpret_cast<const void*>((uintptr_t)base), bytes);
  if (rc != FILO_OK) throw_filo(env, C(ctx), rc);
}

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_hostUnregister(JNIEnv* env, jobject, jlong ctx, jlong base) {
  const int32_t rc = filo_host_unregister(C(ctx), reinterpret_cast<const void*>((uintptr_t)base));
  if (rc != FILO_OK) throw_filo(env, C(ctx), rc);
}

// chunkInfoAddrs: ChunkSetInfo.infoAddr values grouped per series by nChunks; the caller holds the chunk locks for the call
JNIEXPORT jlong JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_loadSeries(JNIEnv* env, jobject, jlong ctx, jlong nSeries, jintArray nChunks, jlongArray chunkInfoAddrs,
                                                                              jint tsCol, jint valCol, jintArray groupIds, jint nGroups, jint schemaFlags) {
  filo_table* t = nullptr; int32_t rc;
  {
    Critical nc(env, nChunks), ia(env, chunkInfoAddrs), gi(env, groupIds);
    rc = filo_load_series(C(ctx), nSeries, static_cast<const int32_t*>(nc.p), static_cast<const uint64_t*>(ia.p), tsCol, valCol,
                          static_cast<const int32_t*>(gi.p), nGroups, schemaFlags, &t);
  }
  if (rc != FILO_OK) { throw_filo(env, C(ctx), rc); return 0; }
  return (jlong)(uintptr_t)t;
}

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_tableFree(JNIEnv*, jobject, jlong ctx, jlong table) { filo_table_free(C(ctx), T(table)); }

JNIEXPORT jint JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_numWindows(JNIEnv*, jobject, jlong startMs, jlong stepMs, jlong endMs) {
  return filo_num_windows(startMs, stepMs, endMs);
}

// outValuesAddr / outAuxAddr: off-heap addresses (direct ByteBuffer address or UnsafeUtils allocation); stats: long[6]
JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_query(JNIEnv* env, jobject, jlong ctx, jlong table, jint rangeFn, jlong startMs, jlong stepMs, jlong endMs,
                                                                        jlong windowMs, jint aggrOp, jint k, jint flags, jlong outValuesAddr, jlong outAuxAddr, jlongArray stats) {
  filo_stats st{};
  const int32_t rc = filo_query(C(ctx), T(table), rangeFn, startMs, stepMs, endMs, windowMs, aggrOp, k, flags,
                                reinterpret_cast<double*>((uintptr_t)outValuesAddr), reinterpret_cast<int64_t*>((uintptr_t)outAuxAddr), &st);
  if (rc != FILO_OK) { throw_filo(env, C(ctx), rc); return; }
  put_stats(env, stats, st);
}

JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_queryHist(JNIEnv* env, jobject, jlong ctx, jlong table, jint rangeFn, jlong startMs, jlong stepMs, jlong endMs,
                                                                            jlong windowMs, jint aggrOp, jdouble quantile, jlong outValuesAddr, jlong outQuantileAddr, jlongArray stats) {
  filo_stats st{};
  const int32_t rc = filo_query_hist(C(ctx), T(table), rangeFn, startMs, stepMs, endMs, windowMs, aggrOp, quantile,
                                     reinterpret_cast<double*>((uintptr_t)outValuesAddr), reinterpret_cast<double*>((uintptr_t)outQuantileAddr), &st);
  if (rc != FILO_OK) { throw_filo(env, C(ctx), rc); return; }
  put_stats(env, stats, st);
}

// avg_over_time over a downsample schema: the sum and the count column as two tables (AvgWithSumAndCountOverTimeFuncD / FuncL)
JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_queryAvgSumCount(JNIEnv* env, jobject, jlong ctx, jlong sumTable, jlong countTable, jlong startMs, jlong stepMs,
                                                                                   jlong endMs, jlong windowMs, jlong outValuesAddr, jlongArray stats) {
  filo_stats st{};
  const int32_t rc = filo_query_avg_sum_count(C(ctx), T(sumTable), T(countTable), startMs, stepMs, endMs, windowMs, reinterpret_cast<double*>((uintptr_t)outValuesAddr), &st);
  if (rc != FILO_OK) { throw_filo(env, C(ctx), rc); return; }
  put_stats(env, stats, st);
}

// the one-call form for a bare PeriodicSamplesMapper: gather + H2D + kernels + D2H, pipelined in batches
JNIEXPORT void JNICALL Java_filodb_gpu_FiloB200NativeMethods_00024_scanSeries(JNIEnv* env, jobject, jlong ctx, jlong nSeries, jintArray nChunks, jlongArray chunkInfoAddrs,
                                                                             jint tsCol, jint valCol, jint schemaFlags, jint rangeFn, jlong startMs, jlong stepMs, jlong endMs,
                                                                             jlong windowMs, jlong outValuesAddr, jlongArray stats) {
  filo_stats st{}; int32_t rc;
  {
    Critical nc(env, nChunks), ia(env, chunkInfoAddrs);
    rc = filo_scan_series(C(ctx), nSeries, static_cast<const int32_t*>(nc.p), static_cast<const uint64_t*>(ia.p), tsCol, valCol, schemaFlags,
                          rangeFn, startMs, stepMs, endMs, windowMs, reinterpret_cast<double*>((uintptr_t)outValuesAddr), &st);
  }
  if (rc != FILO_OK) { throw_filo(env, C(ctx), rc); return; }
  put_stats(env, stats, st);
}

}  // extern "C"
