#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define FILO_HD __host__ __device__
#else
#define FILO_HD
#endif
namespace filo {
enum { FN_LAST = 0, FN_RATE, FN_INCREASE, FN_DELTA, FN_SUM, FN_AVG, FN_COUNT, FN_MIN, FN_MAX, FN_TIMESTAMP };
enum { CLASS_SUM = 0, CLASS_MINMAX = 1, CLASS_POINT = 2, CLASS_COUNTER = 3 };
FILO_HD inline int fn_class_of(int fn, int cumulative) {
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
};
}
