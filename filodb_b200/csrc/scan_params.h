#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define FILO_HD __host__ __device__
#else
#define FILO_HD
#endif
namespace filo {
enum { FN_LAST = 0, FN_RATE, FN_INCREASE, FN_DELTA, FN_SUM, FN_AVG, FN_COUNT, FN_MIN, FN_MAX, FN_TIMESTAMP,
       // the other chunked range functions (RangeFunction.scala:341-375); all of them run window by window (CLASS_POINT)
       FN_STDDEV, FN_STDVAR, FN_CHANGES, FN_QUANTILE, FN_ZSCORE, FN_HOLT_WINTERS, FN_PREDICT_LINEAR, FN_MAD, FN_PRESENT, FN_COUNT_ };
// functions RangeFunction.longChunkedFunction (RangeFunction.scala:319-339) implements for a Long value column
FILO_HD inline bool fn_long_column_ok(int fn) {
  switch (fn) {
    case FN_LAST: case FN_COUNT: case FN_SUM: case FN_AVG: case FN_MIN: case FN_MAX: case FN_STDDEV: case FN_STDVAR: case FN_CHANGES:
    case FN_QUANTILE: case FN_PREDICT_LINEAR: case FN_MAD: return true;
    default: return false;
  }
}
enum { CLASS_SUM = 0, CLASS_MINMAX = 1, CLASS_POINT = 2, CLASS_COUNTER = 3 };
FILO_HD inline int fn_class_of(int fn, int cumulative, int long_values = 0) {
  if (long_values) return CLASS_POINT;      // Long value columns: every function runs window by window (the *L variants in eval_window)
  switch (fn) {
    case FN_SUM: case FN_AVG: case FN_COUNT: return CLASS_SUM;
    case FN_RATE: case FN_INCREASE: return cumulative ? CLASS_COUNTER : CLASS_SUM;
    case FN_DELTA: return CLASS_COUNTER;
    case FN_MIN: case FN_MAX: return CLASS_MINMAX;
    default: return CLASS_POINT;
  }
}
struct QueryParams {
  int64_t start, step, end, window;
  int32_t T;            // number of windows
  int32_t fn;
  int32_t cumulative;   // schema.hasCumulativeTemporalityColumn
  int32_t inclusive;    // filodb.query.inclusive-range
  int32_t long_values;  // the value column is a LongColumn: *L function variants, raw 64-bit value vectors hold longs
  int32_t pad_;
  double p0, p1;        // static function arguments (quantile; sf, tf; duration)
};
}
