// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the per-partition windowing loop and the chunked range functions:
//   WindowedChunkIterator          core/src/main/scala/filodb.core/store/ChunkSetInfo.scala:445-529
//   ChunkedWindowIterator.doNext   query/src/main/scala/filodb/query/exec/PeriodicSamplesMapper.scala:256-347
//   ChunkedRangeFunction family    query/.../exec/rangefn/{RangeFunction,RateFunctions,AggrOverTimeFunctions}.scala
//   RangeVectorAggregator + RowAggregators  query/.../exec/AggrOverRangeVectors.scala:214-378, exec/aggregator/*
// Same operation order as the Scala source (sequential folds), -ffp-contract=off.
#pragma once
#include "filo_format.hpp"
#include <map>
#include <queue>
#include <functional>

namespace fo {

constexpr double NaN = std::numeric_limits<double>::quiet_NaN();

// InternalRangeFunction subset on the hot path (query/.../exec/InternalRangeFunction.scala:11-70).
// Numbering is this repo's C-ABI numbering (include/filo_b200.h), not a reference ordinal.
enum RangeFn : int32_t {
  FN_LAST = 0,            // None / Last  -> LastSampleChunkedFunctionD
  FN_RATE = 1, FN_INCREASE = 2, FN_DELTA = 3,
  FN_SUM_OVER_TIME = 4, FN_AVG_OVER_TIME = 5, FN_COUNT_OVER_TIME = 6,
  FN_MIN_OVER_TIME = 7, FN_MAX_OVER_TIME = 8,
  FN_TIMESTAMP = 9,
};
enum AggrOp : int32_t { AGG_NONE = 0, AGG_SUM = 1, AGG_AVG = 2, AGG_MIN = 3, AGG_MAX = 4, AGG_COUNT = 5, AGG_TOPK = 6, AGG_BOTTOMK = 7 };

// QueryUtils.scala:109-123
inline double maxIgnoreNaN(double a, double b) { if (a != a) return b; if (b != b) return a; return a > b ? a : b; }
inline double minIgnoreNaN(double a, double b) { if (a != a) return b; if (b != b) return a; return a < b ? a : b; }

// RateFunctions.extrapolatedRate, RateFunctions.scala:72-111
inline double extrapolatedRate(int64_t windowStart, int64_t windowEnd, int32_t numSamples,
                               int64_t sample1Time, double sample1Value,
                               int64_t sample2Time, double sample2Value,
                               bool isCounter, bool isRate) {
  double durationToStart = (double)(sample1Time - windowStart) / 1000;
  double durationToEnd = (double)(windowEnd - sample2Time) / 1000;
  double sampledInterval = (double)(sample2Time - sample1Time) / 1000;
  double averageDurationBetweenSamples = sampledInterval / ((double)numSamples - 1);
  double delta = sample2Value - sample1Value;
  if (isCounter && delta > 0 && sample1Value >= 0) {
    double durationToZero = sampledInterval * (sample1Value / delta);
    if (durationToZero < durationToStart) durationToStart = durationToZero;
  }
  double extrapolationThreshold = averageDurationBetweenSamples * 1.1;
  double extrapolateToInterval = sampledInterval;
  extrapolateToInterval += (durationToStart < extrapolationThreshold) ? durationToStart : averageDurationBetweenSamples / 2;
  extrapolateToInterval += (durationToEnd < extrapolationThreshold) ? durationToEnd : averageDurationBetweenSamples / 2;
  double scaledDelta = delta * (extrapolateToInterval / sampledInterval);
  return isRate ? (scaledDelta / (double)(windowEnd - windowStart) * 1000) : scaledDelta;
}

// ChunkSetInfoReader with resolved readers (store/ChunkSetInfoReader.scala:14-73; resolved in
// WindowedChunkIterator.nextWindow, ChunkSetInfo.scala:495-502)
struct InfoReader {
  Ptr info = nullptr;
  Ptr tsVec = nullptr; Ptr valVec = nullptr;
  LongReader tsReader{LongReader::RAW64};
  DoubleReader valReader;
  int32_t numRows() const { return csi::numRows(info); }
  int64_t startTime() const { return csi::startTime(info); }
  int64_t endTime() const { return csi::endTime(info); }
};

// One time series = ordered list of ChunkSetInfo addresses (increasing chunkID), cf. RawDataRangeVector.
struct Series {
  std::vector<Ptr> infos;
  int tsCol = 0, valCol = 1;
};

struct QueryConfig { bool inclusiveRange = true; };   // filodb.query.inclusive-range, filodb-defaults.conf:590

// WindowedChunkIterator, ChunkSetInfo.scala:445-529
struct WindowedChunkIterator {
  const Series& rv; int64_t start, step, end, window; bool isInclusiveRange;
  int64_t curWindowEnd = -1, curWindowStart = -1;
  size_t readIndex = 0;
  std::vector<std::shared_ptr<InfoReader>> windowInfos;
  std::vector<Ptr> infos; size_t infoPos = 0;          // rv.chunkInfos(start - window, end)
  int64_t samplesScanned = 0, bytesScanned = 0;        // CountingChunkInfoIterator, ChunkSetInfo.scala:336-380

  WindowedChunkIterator(const Series& s, int64_t st, int64_t sp, int64_t en, int64_t w, bool incl)
      : rv(s), start(st), step(sp), end(en), window(w), isInclusiveRange(incl) {
    if (!(step > 0)) throw std::invalid_argument("Adjusted step not > 0");
    for (Ptr i : s.infos) if (csi::intersects(i, start - window, end)) infos.push_back(i);   // TimeSeriesPartition.scala:365-366
  }
  bool hasMoreWindows() const { return (curWindowEnd < 0) || (curWindowEnd + step <= end); }
  void nextWindow() {
    if (curWindowEnd == -1) {
      curWindowEnd = start;
      int64_t windowDuration = isInclusiveRange ? window : window - 1;
      curWindowStart = start - std::max<int64_t>(windowDuration, 0);
    } else { curWindowEnd += step; curWindowStart += step; }
    readIndex = 0;
    while (!windowInfos.empty() && windowInfos[0]->endTime() < curWindowStart) windowInfos.erase(windowInfos.begin());
    int64_t lastEndTime = windowInfos.empty() ? -1 : windowInfos.back()->endTime();
    while (curWindowEnd > lastEndTime && infoPos < infos.size()) {
      Ptr nextInfo = infos[infoPos++];
      bytesScanned += totalBytes(csi::vectorPtr(nextInfo, rv.tsCol)) + totalBytes(csi::vectorPtr(nextInfo, rv.valCol));
      samplesScanned += csi::numRows(nextInfo);
      if (curWindowStart <= csi::endTime(nextInfo) && csi::numRows(nextInfo) > 0) {
        auto r = std::make_shared<InfoReader>();
        r->info = nextInfo;
        r->tsVec = csi::vectorPtr(nextInfo, rv.tsCol);
        r->tsReader = LongReader::of(r->tsVec);
        r->valVec = csi::vectorPtr(nextInfo, rv.valCol);
        r->valReader = DoubleReader::of(r->valVec);
        windowInfos.push_back(r);
        lastEndTime = std::max(r->endTime(), lastEndTime);
      }
    }
  }
  bool hasNext() const { return readIndex < windowInfos.size(); }
  InfoReader& next() { return *windowInfos[readIndex++]; }
};

// ---- chunked range functions (state machine per window)
struct ChunkedFn {
  RangeFn fn; bool cumulative; bool inclusiveRange;
  // SumOverTimeChunkedFunctionD (AggrOverTimeFunctions.scala:553-572) / Avg (:992-1015) / Count (:940-958) / Min,Max (:40-97)
  double sum = NaN; int32_t count = 0; double countD = NaN; double mn = NaN, mx = NaN;
  // LastSampleChunkedFunctionD (RangeFunction.scala:595-627, 684-693)
  int64_t lastTs = -1; double lastVal = NaN;
  // ChunkedRateFunctionBase (RateFunctions.scala:230-285)
  int32_t numSamples = 0; int64_t lowestTime = INT64_MAX; double lowestValue = NaN; int64_t highestTime = 0; double highestValue = NaN;
  DoubleCorrection correctionMeta;      // CounterChunkedRangeFunction (RangeFunction.scala:131-136)
  double tsVal = NaN;                   // TimestampChunkedFunction (RangeFunction.scala:705-724)

  void reset() {
    sum = NaN; count = 0; countD = NaN; mn = NaN; mx = NaN; lastTs = -1; lastVal = NaN;
    numSamples = 0; lowestTime = INT64_MAX; lowestValue = NaN; highestTime = 0; highestValue = NaN;
    correctionMeta = DoubleCorrection{}; tsVal = NaN;
  }
  bool isCounterPath() const { return (fn == FN_RATE || fn == FN_INCREASE) ? cumulative : (fn == FN_DELTA); }

  void addSum(DoubleReader& r, int s, int e) {          // AggrOverTimeFunctions.scala:560-571
    double chunkSum = r.sum(s, e);
    if (!std::isnan(chunkSum) && std::isnan(sum)) sum = 0;
    sum += chunkSum;
  }
  void addChunks(InfoReader& ir, int64_t startTime, int64_t endTime) {
    LongReader& ts = ir.tsReader; DoubleReader& val = ir.valReader;
    if (fn == FN_LAST || fn == FN_TIMESTAMP) {            // RangeFunction.scala:603-613 / :708-716
      int32_t endRowNum = std::min(ts.ceilingIndex(ir.tsVec, endTime), ir.numRows() - 1);
      if (endRowNum >= 0) {
        int64_t t = ts.apply(ir.tsVec, endRowNum);
        if (fn == FN_TIMESTAMP) { tsVal = (double)t / (double)1000.0f; }
        else if (t >= startTime && t > lastTs) { lastTs = t; lastVal = val.apply(endRowNum); }
      }
      return;
    }
    int32_t startRowNum = ts.binarySearch(ir.tsVec, startTime) & 0x7fffffff;       // RangeFunction.scala:185-190 / :141-142
    int32_t endRowNum = std::min(ts.ceilingIndex(ir.tsVec, endTime), ir.numRows() - 1);
    if (isCounterPath()) {                                 // CounterChunkedRangeFunction.addChunks, RangeFunction.scala:138-163
      correctionMeta = val.detectDropAndCorrection(correctionMeta);
      if (startRowNum <= endRowNum)
        addTimeChunksRate(val, startRowNum, endRowNum, ts.apply(ir.tsVec, startRowNum), ts.apply(ir.tsVec, endRowNum));
      correctionMeta = val.updateCorrection(correctionMeta);
      return;
    }
    if (!(startRowNum <= endRowNum)) return;
    switch (fn) {
      case FN_RATE: case FN_INCREASE: case FN_SUM_OVER_TIME: addSum(val, startRowNum, endRowNum); break;   // delta branch: RateFunctions.scala:424-445
      case FN_AVG_OVER_TIME: addSum(val, startRowNum, endRowNum); count += val.count(startRowNum, endRowNum); break;
      case FN_COUNT_OVER_TIME: if (std::isnan(countD)) countD = 0; countD += val.count(startRowNum, endRowNum); break;
      case FN_MIN_OVER_TIME: for (int r = startRowNum; r <= endRowNum; ++r) mn = minIgnoreNaN(mn, val.apply(r)); break;
      case FN_MAX_OVER_TIME: for (int r = startRowNum; r <= endRowNum; ++r) mx = maxIgnoreNaN(mx, val.apply(r)); break;
      default: break;
    }
  }
  void addTimeChunksRate(DoubleReader& r, int s, int e, int64_t startTime, int64_t endTime) {
    if (fn == FN_DELTA) {                                  // ChunkedDeltaFunction, RateFunctions.scala:303-320
      if (startTime < lowestTime || endTime > highestTime) {
        numSamples += e - s + 1;
        if (startTime < lowestTime) { lowestTime = startTime; lowestValue = r.apply(s); }
        if (endTime > highestTime) { highestTime = endTime; highestValue = r.apply(e); }
      }
      return;
    }
    if (s == 0 && e == 0 && std::isnan(r.apply(s))) return;         // RateFunctions.scala:255-256
    if (startTime < lowestTime || endTime > highestTime) {            // :257-267
      numSamples += e - s + 1;
      if (startTime < lowestTime) { lowestTime = startTime; lowestValue = r.correctedValue(s, correctionMeta); }
      if (endTime > highestTime) { highestTime = endTime; highestValue = r.correctedValue(e, correctionMeta); }
    }
  }
  double apply(int64_t windowStart, int64_t windowEnd) const {
    switch (fn) {
      case FN_LAST: return lastVal;
      case FN_TIMESTAMP: return tsVal;
      case FN_SUM_OVER_TIME: return sum;
      case FN_AVG_OVER_TIME: return count > 0 ? sum / count : (std::isnan(sum) ? sum : 0.0);    // AggrOverTimeFunctions.scala:1000
      case FN_COUNT_OVER_TIME: return countD;
      case FN_MIN_OVER_TIME: return mn;
      case FN_MAX_OVER_TIME: return mx;
      case FN_RATE: case FN_INCREASE: case FN_DELTA: {
        int64_t curWindowStart = inclusiveRange ? windowStart : windowStart - 1;
        if (isCounterPath()) {                              // RateFunctions.scala:270-285
          if (highestTime > lowestTime)
            return extrapolatedRate(curWindowStart, windowEnd, numSamples, lowestTime, lowestValue, highestTime, highestValue,
                                    fn != FN_DELTA, fn == FN_RATE);
          return NaN;
        }
        if (fn == FN_RATE) return sum / (double)(windowEnd - curWindowStart) * 1000;     // RateFunctions.scala:436-442
        return sum;                                          // increase on delta schema = SumOverTimeChunkedFunctionD
      }
    }
    return NaN;
  }
};

struct QueryStats { int64_t samplesScanned = 0, bytesScanned = 0; };

inline int numWindows(int64_t start, int64_t step, int64_t end) {
  // hasMoreWindows/nextWindow: first window always, then while curWindowEnd + step <= end
  if (end < start) return 1;
  return (int)((end - start) / step) + 1;
}

// ChunkedWindowIteratorD, PeriodicSamplesMapper.scala:293-330 — all windows of one series
inline void periodicSamples(const Series& s, RangeFn fn, bool cumulative, int64_t start, int64_t step, int64_t end, int64_t window,
                            const QueryConfig& cfg, double* out /*[T]*/, QueryStats* stats = nullptr) {
  WindowedChunkIterator wit(s, start, step, end, window, cfg.inclusiveRange);
  ChunkedFn f{fn, cumulative, cfg.inclusiveRange};
  int k = 0;
  while (wit.hasMoreWindows()) {
    f.reset();
    wit.nextWindow();
    while (wit.hasNext()) { InfoReader& ir = wit.next(); f.addChunks(ir, wit.curWindowStart, wit.curWindowEnd); }
    out[k++] = f.apply(wit.curWindowStart, wit.curWindowEnd);
  }
  if (stats) { stats->samplesScanned += wit.samplesScanned; stats->bytesScanned += wit.bytesScanned; }
}

// ---- across-series aggregation: fold in series (arrival) order, per group, per window.
// Output layout: values[G*T] (+ aux per op).  RowAggregators: Sum :22-29, Min/Max :22-26, Count :31-43, Avg :31-46.
struct AggResult {
  std::vector<double> values;     // [G*T] ; topk: [G*T*k] values ascending (topk) / descending (bottomk), padded
  std::vector<int64_t> aux;       // avg: counts [G*T]; topk: series ids [G*T*k] (-1 pad)
};

inline AggResult aggregate(AggrOp op, int k, const std::vector<const double*>& rows /*[S] -> [T]*/,
                           const std::vector<int32_t>& groups, int G, int T) {
  AggResult r;
  size_t S = rows.size();
  if (op == AGG_TOPK || op == AGG_BOTTOMK) {
    // TopBottomKRowAggregator.scala:67-95: heap of (key,value), NaN skipped, keep k.  Presented (toRowReader :45-60)
    // in dequeue order: topk ascending, bottomk descending.  Ties: unspecified in the reference (PriorityQueue);
    // here: the element that arrived later is evicted first among equals.
    bool bottom = (op == AGG_BOTTOMK);
    r.values.assign((size_t)G * T * k, bottom ? std::numeric_limits<double>::max() : std::numeric_limits<double>::lowest());
    r.aux.assign((size_t)G * T * k, -1);
    for (int g = 0; g < G; ++g) for (int t = 0; t < T; ++t) {
      std::vector<std::pair<double, int64_t>> heap;    // kept sorted: best candidates
      for (size_t s = 0; s < S; ++s) if (groups[s] == g) {
        double v = rows[s][t];
        if (std::isnan(v)) continue;
        heap.emplace_back(v, (int64_t)s);
        // evict the "worst" (smallest for topk, largest for bottomk) when size > k
        if ((int)heap.size() > k) {
          size_t worst = 0;
          for (size_t i = 1; i < heap.size(); ++i) {
            bool worse = bottom ? (heap[i].first > heap[worst].first) : (heap[i].first < heap[worst].first);
            bool tie = heap[i].first == heap[worst].first;
            if (worse || (tie && heap[i].second > heap[worst].second)) worst = i;
          }
          heap.erase(heap.begin() + worst);
        }
      }
      std::sort(heap.begin(), heap.end(), [&](auto& a, auto& b) {
        if (a.first != b.first) return bottom ? a.first > b.first : a.first < b.first;
        return a.second > b.second;
      });
      for (size_t i = 0; i < heap.size(); ++i) {
        r.values[((size_t)g * T + t) * k + i] = heap[i].first;
        r.aux[((size_t)g * T + t) * k + i] = heap[i].second;
      }
    }
    return r;
  }
  r.values.assign((size_t)G * T, NaN);
  if (op == AGG_AVG) r.aux.assign((size_t)G * T, 0);
  for (size_t s = 0; s < S; ++s) {
    int g = groups[s];
    for (int t = 0; t < T; ++t) {
      double v = rows[s][t];
      double& acc = r.values[(size_t)g * T + t];
      switch (op) {
        case AGG_SUM: if (!std::isnan(v)) { if (std::isnan(acc)) acc = 0; acc += v; } break;
        case AGG_MIN: acc = minIgnoreNaN(acc, v); break;
        case AGG_MAX: acc = maxIgnoreNaN(acc, v); break;
        case AGG_COUNT: {                                  // CountRowAggregator.scala:31-43
          double x = std::isnan(v) ? 0.0 : 1.0;
          if (std::isnan(acc) && x > 0) acc = 0;
          if (!std::isnan(x)) acc += x;
          break;
        }
        case AGG_AVG: {                                    // AvgRowAggregator.scala:31-46
          int64_t& cnt = r.aux[(size_t)g * T + t];
          int64_t c = std::isnan(v) ? 0 : 1;
          if (!std::isnan(v)) {
            if (std::isnan(acc)) acc = 0;
            acc = (acc * (double)cnt + v * (double)c) / (double)(cnt + c);
            cnt += c;
          }
          break;
        }
        default: break;
      }
    }
  }
  return r;
}

// ---- sliding (row-wise) cross-check path: SlidingWindowIterator + RateFunction/etc.
// (PeriodicSamplesMapper.scala:362-456, RateFunctions.scala:18-32, AggrOverTimeFunctions.scala sliding variants)
// The reference's own tests use "chunked == sliding" as their oracle (RateFunctionsSpec.scala:181-212).
// Here: window = samples with wStart <= ts <= wEnd (inclusive-range=true), plain row-wise evaluation with
// Prometheus counter correction (BufferableCounterCorrectionIterator :551-574 — correction accumulates from
// the first sample of the series, NaN skipped by DropNaN for counter functions).
inline void slidingSamples(const std::vector<int64_t>& ts, const std::vector<double>& vals, RangeFn fn, bool cumulative,
                           int64_t start, int64_t step, int64_t end, int64_t window, double* out) {
  int T = numWindows(start, step, end);
  bool counter = (fn == FN_RATE || fn == FN_INCREASE) && cumulative;
  std::vector<int64_t> t2; std::vector<double> v2;
  if (counter) {                       // BufferableCounterCorrectionIterator, PeriodicSamplesMapper.scala:551-574
    double corr = 0, prevVal = 0;
    for (size_t i = 0; i < ts.size(); ++i) {
      double v = vals[i];
      if (std::isnan(v)) v = 0;        // explicit counter reset due to end of time series marker
      if (v < prevVal) corr += prevVal;
      t2.push_back(ts[i]); v2.push_back(v + corr);
      prevVal = v;
    }
  } else { t2 = ts; v2 = vals; }
  for (int k = 0; k < T; ++k) {
    int64_t wEnd = start + (int64_t)k * step, wStart = wEnd - window;
    size_t lo = std::lower_bound(t2.begin(), t2.end(), wStart) - t2.begin();
    size_t hi = std::upper_bound(t2.begin(), t2.end(), wEnd) - t2.begin();   // [lo, hi)
    double res = NaN;
    switch (fn) {
      case FN_RATE: case FN_INCREASE: case FN_DELTA:
        if (counter || fn == FN_DELTA) {
          // non-counter delta keeps NaN samples out too
          if (hi - lo >= 2) res = extrapolatedRate(wStart, wEnd, (int)(hi - lo), t2[lo], v2[lo], t2[hi - 1], v2[hi - 1], fn != FN_DELTA, fn == FN_RATE);
        } else {
          double s = NaN; for (size_t i = lo; i < hi; ++i) if (!std::isnan(v2[i])) { if (std::isnan(s)) s = 0; s += v2[i]; }
          res = (fn == FN_RATE) ? s / (double)(wEnd - wStart) * 1000 : s;
        }
        break;
      case FN_SUM_OVER_TIME: { double s = NaN; for (size_t i = lo; i < hi; ++i) if (!std::isnan(v2[i])) { if (std::isnan(s)) s = 0; s += v2[i]; } res = s; break; }
      case FN_COUNT_OVER_TIME: { if (hi > lo) { double c = 0; for (size_t i = lo; i < hi; ++i) if (!std::isnan(v2[i])) c += 1; res = c; } break; }
      case FN_MIN_OVER_TIME: for (size_t i = lo; i < hi; ++i) res = minIgnoreNaN(res, v2[i]); break;
      case FN_MAX_OVER_TIME: for (size_t i = lo; i < hi; ++i) res = maxIgnoreNaN(res, v2[i]); break;
      case FN_LAST: if (hi > lo) res = v2[hi - 1]; break;
      default: break;
    }
    out[k] = res;
  }
}

} // namespace fo
