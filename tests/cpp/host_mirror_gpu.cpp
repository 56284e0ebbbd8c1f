// GPU driver for the two host-side layers above the C-ABI (test infrastructure; built and run by tests/test_gpu_parity.py):
//   1. include/filo_b200.hpp  -- FusedGpuExec::execute over the reference's transformer chain (PeriodicSamplesMapper [+ AggregateMapReduce]),
//   2. filodb_b200/csrc/jni_shim.cpp -- the exported Java_filodb_gpu_FiloB200NativeMethods_00024_* functions, called through a JNIEnv whose
//      interface function table is filled with host functions (arrays are plain structs, ThrowNew records the message).
// Both are checked against the oracle (oracle/filo_query.hpp) on chunks built with the oracle's encoders: per-series results bit for
// bit, across-series sums to 1e-9.
#include "../../include/filo_b200.hpp"
#include "../../filodb_b200/csrc/jni_stub/jni.h"
#include "../../oracle/filo_query.hpp"
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>

// ---- the shim's exports (libfilo_b200_jni.so)
extern "C" {
jlong Java_filodb_gpu_FiloB200NativeMethods_00024_ctxCreate(JNIEnv*, jobject, jint, jboolean, jlong, jlong, jint);
void Java_filodb_gpu_FiloB200NativeMethods_00024_ctxDestroy(JNIEnv*, jobject, jlong);
void Java_filodb_gpu_FiloB200NativeMethods_00024_ctxSetFnArgs(JNIEnv*, jobject, jlong, jdouble, jdouble);
jlong Java_filodb_gpu_FiloB200NativeMethods_00024_loadSeries(JNIEnv*, jobject, jlong, jlong, jintArray, jlongArray, jint, jint, jintArray, jint, jint);
void Java_filodb_gpu_FiloB200NativeMethods_00024_tableFree(JNIEnv*, jobject, jlong, jlong);
jint Java_filodb_gpu_FiloB200NativeMethods_00024_numWindows(JNIEnv*, jobject, jlong, jlong, jlong);
void Java_filodb_gpu_FiloB200NativeMethods_00024_query(JNIEnv*, jobject, jlong, jlong, jint, jlong, jlong, jlong, jlong, jint, jint, jint, jlong, jlong, jlongArray);
void Java_filodb_gpu_FiloB200NativeMethods_00024_scanSeries(JNIEnv*, jobject, jlong, jlong, jintArray, jlongArray, jint, jint, jint, jint, jlong, jlong, jlong, jlong, jlong, jlongArray);
}

// ---- a host JNIEnv: arrays are FakeArray objects, exceptions are recorded
struct FakeArray { void* data; jsize len; };
static std::string g_thrown; static int g_critical = 0;
static jclass f_FindClass(JNIEnv*, const char* name) { static int cls; (void)name; return reinterpret_cast<jclass>(&cls); }
static jint f_ThrowNew(JNIEnv*, jclass, const char* msg) { g_thrown = msg; return 0; }
static jsize f_GetArrayLength(JNIEnv*, jarray a) { return reinterpret_cast<FakeArray*>(a)->len; }
static void f_SetLongArrayRegion(JNIEnv*, jlongArray a, jsize start, jsize len, const jlong* buf) { std::memcpy(static_cast<jlong*>(reinterpret_cast<FakeArray*>(a)->data) + start, buf, (size_t)len * 8); }
static void* f_GetCritical(JNIEnv*, jarray a, jboolean*) { ++g_critical; return reinterpret_cast<FakeArray*>(a)->data; }
static void f_ReleaseCritical(JNIEnv*, jarray, void*, jint) { --g_critical; }
static jboolean f_ExceptionCheck(JNIEnv*) { return g_thrown.empty() ? 0 : 1; }

struct Chunk { std::vector<uint8_t> ts, vv, info; };
static bool same_bits(double a, double b) { uint64_t x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8); return x == y || (a != a && b != b); }

int main() {
  // ---- data: counters with resets, XOR values, two chunks per series
  const int S = 64, rows = 300, G = 5; const int64_t t0 = 1700000000000LL, step = 15000;
  std::mt19937_64 rng(99); std::normal_distribution<double> N(0.0, 1.0);
  std::vector<std::vector<std::unique_ptr<Chunk>>> series((size_t)S);
  std::vector<filo::RawDataRangeVector> source((size_t)S);
  for (int s = 0; s < S; ++s) {
    std::vector<int64_t> ts((size_t)rows); std::vector<double> v((size_t)rows); double acc = 0;
    for (int r = 0; r < rows; ++r) { ts[(size_t)r] = t0 + r * step; if (r && rng() % 97 == 0) acc = 0; acc += std::max(0.0, 15 + std::sin(r + 1.0) + N(rng)); v[(size_t)r] = acc; }
    int r0 = 0;
    for (int n : {180, 120}) {
      auto c = std::make_unique<Chunk>();
      c->ts = fo::enc::timestamps(ts.data() + r0, n);
      c->vv = fo::enc::doublesXor(v.data() + r0, n, true);
      c->info.assign(fo::csi::OffsetVectors + 16, 0);
      fo::setLong(c->info.data() + fo::csi::OffsetChunkID, fo::csi::chunkID(ts[(size_t)r0], (ts[(size_t)(r0 + n - 1)] + 1000) / 1000));
      fo::setInt(c->info.data() + fo::csi::OffsetNumRows, n);
      fo::setLong(c->info.data() + fo::csi::OffsetIngestionTime, ts[(size_t)(r0 + n - 1)] + 1000);
      fo::setLong(c->info.data() + fo::csi::OffsetEndTime, ts[(size_t)(r0 + n - 1)]);
      fo::setLong(c->info.data() + fo::csi::OffsetVectors, (int64_t)(uintptr_t)c->ts.data());
      fo::setLong(c->info.data() + fo::csi::OffsetVectors + 8, (int64_t)(uintptr_t)c->vv.data());
      source[(size_t)s].chunkInfoAddrs.push_back((uint64_t)(uintptr_t)c->info.data());
      series[(size_t)s].push_back(std::move(c));
      r0 += n;
    }
    source[(size_t)s].group = s % G;
  }
  const int64_t qstart = t0 + 300000, qend = t0 + (rows - 1) * step, window = 300000;
  const int T = fo::numWindows(qstart, step, qend);
  // ---- oracle
  std::vector<double> ref((size_t)S * T);
  for (int s = 0; s < S; ++s) {
    fo::Series os; for (auto& c : series[(size_t)s]) os.infos.push_back(c->info.data());
    fo::periodicSamples(os, fo::FN_RATE, true, qstart, step, qend, window, fo::QueryConfig{true}, ref.data() + (size_t)s * T);
  }
  std::vector<const double*> rowsp; std::vector<int32_t> groups;
  for (int s = 0; s < S; ++s) { rowsp.push_back(ref.data() + (size_t)s * T); groups.push_back(s % G); }
  const fo::AggResult sumref = fo::aggregate(fo::AGG_SUM, 0, rowsp, groups, G, T);
  const fo::AggResult topref = fo::aggregate(fo::AGG_TOPK, 2, rowsp, groups, G, T);
  long checked = 0;

  // ---- 1. the C++ operator mirror
  {
    filo::FusedGpuExec ex(0);
    filo::PeriodicSamplesMapper psm(qstart, step, qend, window, filo::InternalRangeFunction::Rate);
    filo::QueryResult r = ex.execute(source, psm, nullptr, nullptr, 1, /*cumulative=*/true);
    if (r.rows != S || r.windows != T) { std::printf("FAIL execute: shape %d x %d\n", r.rows, r.windows); return 1; }
    for (size_t i = 0; i < ref.size(); ++i, ++checked) if (!same_bits(r.values[i], ref[i])) { std::printf("FAIL execute per series at %zu: %.17g vs %.17g\n", i, r.values[i], ref[i]); return 1; }
    if (r.stats.samples_scanned != (int64_t)S * rows) { std::printf("FAIL execute: samples_scanned %lld\n", (long long)r.stats.samples_scanned); return 1; }
    filo::AggregateMapReduce sum(filo::AggregationOperator::Sum, {}, G);
    filo::QueryResult a = ex.execute(source, psm, &sum, nullptr, 1, true);
    if (a.rows != G) { std::printf("FAIL execute sum: rows %d\n", a.rows); return 1; }
    for (size_t i = 0; i < sumref.values.size(); ++i, ++checked) {
      const double e = sumref.values[i], g = a.values[i];
      if ((e != e) != (g != g) || (e == e && std::fabs(g - e) > 1e-9 * std::fabs(e))) { std::printf("FAIL execute sum at %zu: %.17g vs %.17g\n", i, g, e); return 1; }
    }
    filo::AggregateMapReduce topk(filo::AggregationOperator::TopK, {2.0}, G);
    filo::QueryResult tk = ex.execute(source, psm, &topk, nullptr, 1, true);
    for (size_t i = 0; i < topref.values.size(); ++i, ++checked)
      if (topref.aux[i] >= 0 && !same_bits(tk.values[i], topref.values[i])) { std::printf("FAIL execute topk at %zu: %.17g vs %.17g\n", i, tk.values[i], topref.values[i]); return 1; }
    // requirement failures and engine errors surface as the reference's exceptions
    bool threw = false;
    try { filo::PeriodicSamplesMapper bad(qstart, step, qend, std::nullopt, filo::InternalRangeFunction::Rate); } catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { std::printf("FAIL: missing window accepted\n"); return 1; }
    threw = false;
    std::vector<filo::RawDataRangeVector> rev = source; std::swap(rev[0].chunkInfoAddrs[0], rev[0].chunkInfoAddrs[1]);      // chunks out of time order
    try { ex.execute(rev, psm, nullptr, nullptr, 1, true); } catch (const filo::QueryError& e) { threw = e.status == FILO_ERR_UNSUPPORTED; }
    if (!threw) { std::printf("FAIL: out-of-order chunks did not raise QueryError(UNSUPPORTED)\n"); return 1; }
  }

  // ---- 2. the JNI shim through a host JNIEnv
  {
    JNINativeInterface_ table{};
    table.slot[6] = (void*)f_FindClass; table.slot[14] = (void*)f_ThrowNew; table.slot[171] = (void*)f_GetArrayLength; table.slot[212] = (void*)f_SetLongArrayRegion;
    table.slot[222] = (void*)f_GetCritical; table.slot[223] = (void*)f_ReleaseCritical; table.slot[228] = (void*)f_ExceptionCheck;
    JNIEnv env{&table};
    std::vector<jint> nch((size_t)S, 2), gids; std::vector<jlong> addrs;
    for (int s = 0; s < S; ++s) { gids.push_back(s % G); for (uint64_t a : source[(size_t)s].chunkInfoAddrs) addrs.push_back((jlong)a); }
    FakeArray a_nch{nch.data(), (jsize)nch.size()}, a_addrs{addrs.data(), (jsize)addrs.size()}, a_gids{gids.data(), (jsize)gids.size()};
    jlong st[6] = {0, 0, 0, 0, 0, 0}; FakeArray a_st{st, 6};
    const jlong ctx = Java_filodb_gpu_FiloB200NativeMethods_00024_ctxCreate(&env, nullptr, 0, 1, 0, 0, 0);
    if (!ctx || !g_thrown.empty()) { std::printf("FAIL jni ctxCreate: %s\n", g_thrown.c_str()); return 1; }
    if (Java_filodb_gpu_FiloB200NativeMethods_00024_numWindows(&env, nullptr, qstart, step, qend) != T) { std::printf("FAIL jni numWindows\n"); return 1; }
    const jlong tab = Java_filodb_gpu_FiloB200NativeMethods_00024_loadSeries(&env, nullptr, ctx, S, (jintArray)&a_nch, (jlongArray)&a_addrs, 0, 1, (jintArray)&a_gids, G, FILO_SCHEMA_CUMULATIVE);
    if (!tab || !g_thrown.empty() || g_critical != 0) { std::printf("FAIL jni loadSeries: %s (critical %d)\n", g_thrown.c_str(), g_critical); return 1; }
    std::vector<double> out((size_t)S * T, -1.0);
    Java_filodb_gpu_FiloB200NativeMethods_00024_query(&env, nullptr, ctx, tab, FILO_FN_RATE, qstart, step, qend, window, FILO_AGG_NONE, 0, 0, (jlong)(uintptr_t)out.data(), 0, (jlongArray)&a_st);
    if (!g_thrown.empty()) { std::printf("FAIL jni query: %s\n", g_thrown.c_str()); return 1; }
    for (size_t i = 0; i < ref.size(); ++i, ++checked) if (!same_bits(out[i], ref[i])) { std::printf("FAIL jni query at %zu: %.17g vs %.17g\n", i, out[i], ref[i]); return 1; }
    if (st[0] != (jlong)S * rows || st[3] < 1) { std::printf("FAIL jni stats: samples %lld launches %lld\n", (long long)st[0], (long long)st[3]); return 1; }
    std::vector<double> gout((size_t)G * T, -1.0);
    Java_filodb_gpu_FiloB200NativeMethods_00024_query(&env, nullptr, ctx, tab, FILO_FN_RATE, qstart, step, qend, window, FILO_AGG_SUM, 0, 0, (jlong)(uintptr_t)gout.data(), 0, (jlongArray)&a_st);
    for (size_t i = 0; i < sumref.values.size(); ++i, ++checked) {
      const double e = sumref.values[i], g = gout[i];
      if ((e != e) != (g != g) || (e == e && std::fabs(g - e) > 1e-9 * std::fabs(e))) { std::printf("FAIL jni sum at %zu: %.17g vs %.17g\n", i, g, e); return 1; }
    }
    std::fill(out.begin(), out.end(), -1.0);
    Java_filodb_gpu_FiloB200NativeMethods_00024_scanSeries(&env, nullptr, ctx, S, (jintArray)&a_nch, (jlongArray)&a_addrs, 0, 1, FILO_SCHEMA_CUMULATIVE, FILO_FN_RATE, qstart, step, qend, window,
                                                           (jlong)(uintptr_t)out.data(), (jlongArray)&a_st);
    if (!g_thrown.empty()) { std::printf("FAIL jni scanSeries: %s\n", g_thrown.c_str()); return 1; }
    for (size_t i = 0; i < ref.size(); ++i, ++checked) if (!same_bits(out[i], ref[i])) { std::printf("FAIL jni scanSeries at %zu\n", i); return 1; }
    // a failing call raises RuntimeException("filo_b200 error <code>: <message>") and returns
    Java_filodb_gpu_FiloB200NativeMethods_00024_query(&env, nullptr, ctx, tab, 99, qstart, step, qend, window, FILO_AGG_NONE, 0, 0, (jlong)(uintptr_t)out.data(), 0, (jlongArray)&a_st);
    if (g_thrown.find("filo_b200 error -1") != 0) { std::printf("FAIL jni error path: '%s'\n", g_thrown.c_str()); return 1; }
    g_thrown.clear();
    Java_filodb_gpu_FiloB200NativeMethods_00024_tableFree(&env, nullptr, ctx, tab);
    Java_filodb_gpu_FiloB200NativeMethods_00024_ctxDestroy(&env, nullptr, ctx);
  }
  std::printf("OK host mirror + JNI shim: %ld values checked against the oracle\n", checked);
  return 0;
}
