#pragma once
#include <stdint.h>
namespace filo {
struct QueryParams {
  int64_t start, step, end, window;
  int32_t T;            // number of windows
  int32_t fn;
  int32_t cumulative;   // schema.hasCumulativeTemporalityColumn
  int32_t inclusive;    // filodb.query.inclusive-range
};
}
