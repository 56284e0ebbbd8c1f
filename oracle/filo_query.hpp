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
  // the remaining chunked range functions of RangeFunction.doubleChunkedFunction / longChunkedFunction (RangeFunction.scala:319-375)
  FN_STDDEV_OVER_TIME = 10, FN_STDVAR_OVER_TIME = 11, FN_CHANGES = 12,
  FN_QUANTILE_OVER_TIME = 13,     // param0 = q
  FN_ZSCORE = 14,
  FN_HOLT_WINTERS = 15,           // param0 = sf, param1 = tf
  FN_PREDICT_LINEAR = 16,         // param0 = duration (s)
  FN_MAD_OVER_TIME = 17,          // MedianAbsoluteDeviationOverTime
  FN_PRESENT_OVER_TIME = 18,
};
// which functions RangeFunction.longChunkedFunction (RangeFunction.scala:319-339) has a chunked implementation for
inline bool longColumnSupports(int fn) {
  switch (fn) {
    case FN_LAST: case FN_COUNT_OVER_TIME: case FN_SUM_OVER_TIME: case FN_AVG_OVER_TIME: case FN_MIN_OVER_TIME: case FN_MAX_OVER_TIME:
    case FN_STDDEV_OVER_TIME: case FN_STDVAR_OVER_TIME: case FN_CHANGES: case FN_QUANTILE_OVER_TIME: case FN_PREDICT_LINEAR:
    case FN_MAD_OVER_TIME: return true;
    default: return false;
  }
}
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
  LongReader valLong{LongReader::RAW64};   // value reader of a Long column (valueReader.asLongReader)
  int32_t numRows() const { return csi::numRows(info); }
  int64_t startTime() const { return csi::startTime(info); }
  int64_t endTime() const { return csi::endTime(info); }
};

// One time series = ordered list of ChunkSetInfo addresses (increasing chunkID), cf. RawDataRangeVector.
struct Series {
  std::vector<Ptr> infos;
  int tsCol = 0, valCol = 1;
  bool longCol = false;          // the value column is a LongColumn (RangeFunction.scala:300-304 picks the L variants)
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
        if (rv.longCol) r->valLong = LongReader::of(r->valVec);
        else r->valReader = DoubleReader::of(r->valVec);
        windowInfos.push_back(r);
        lastEndTime = std::max(r->endTime(), lastEndTime);
      }
    }
  }
  bool hasNext() const { return readIndex < windowInfos.size(); }
  InfoReader& next() { return *windowInfos[readIndex++]; }
};

// ---- chunked range functions (state machine per window)
// java.util.Arrays.sort(double[]) order: Double.compare (-0.0 < 0.0, NaN last)
inline bool javaDoubleLess(double a, double b) {
  if (a < b) return true;
  if (a > b) return false;
  if (a == b) { if (a != 0.0) return false; return std::signbit(a) && !std::signbit(b); }
  return !std::isnan(a) && std::isnan(b);      // at least one NaN
}
// QuantileOverTimeFunction.calculateRank, AggrOverTimeFunctions.scala:399-407
inline void calculateRank(double q, int counter, double& weight, int& upperIndex, int& lowerIndex) {
  double rank = q * (double)(counter - 1);
  double lower = std::max(0.0, std::floor(rank));
  double upper = std::min((double)(counter - 1), lower + 1);
  weight = rank - std::floor(rank);
  upperIndex = (int)upper; lowerIndex = (int)lower;
}

struct ChunkedFn {
  RangeFn fn; bool cumulative; bool inclusiveRange;
  bool longCol = false;                 // value column is a LongColumn: the *L function variants
  double p0 = 0, p1 = 0;                // static function arguments (funcParams: Seq[StaticFuncArgs])
  // SumOverTimeChunkedFunctionD (AggrOverTimeFunctions.scala:553-572) / Avg (:992-1015) / Count (:940-958) / Min,Max (:40-97)
  double sum = NaN; int32_t count = 0; double countD = NaN; double mn = NaN, mx = NaN;
  int64_t mnL = INT64_MAX, mxL = INT64_MIN;    // Min/MaxOverTimeChunkedFunctionL :60-116
  // LastSampleChunkedFunctionD (RangeFunction.scala:595-627, 684-693)
  int64_t lastTs = -1; double lastVal = NaN;
  // ChunkedRateFunctionBase (RateFunctions.scala:230-285)
  int32_t numSamples = 0; int64_t lowestTime = INT64_MAX; double lowestValue = NaN; int64_t highestTime = 0; double highestValue = NaN;
  DoubleCorrection correctionMeta;      // CounterChunkedRangeFunction (RangeFunction.scala:131-136)
  double tsVal = NaN;                   // TimestampChunkedFunction (RangeFunction.scala:705-724)
  // VarOverTimeChunkedFunctionD/L (:1082-1183), ZScoreChunkedFunctionD (:1592-1604): lastSample is NOT cleared by reset()
  double squaredSum = NaN, lastSample = NaN;
  // ChangesChunkedFunction (:1185-1225)
  double changes = NaN, prev = NaN;
  // Quantile / MAD (:1227-1359)
  double quantileResult = NaN; std::vector<double> values;
  // HoltWintersChunkedFunctionD (:1361-1453)
  double b0 = NaN, s0 = NaN, nextvalue = NaN, smoothedResult = NaN;
  // PredictLinearChunkedFunction (:1496-1590)
  double sumX = NaN, sumY = NaN, sumXY = NaN, sumX2 = NaN; int32_t counter = 0;

  void reset() {
    sum = NaN; count = 0; countD = NaN; mn = NaN; mx = NaN; lastTs = -1; lastVal = NaN;
    numSamples = 0; lowestTime = INT64_MAX; lowestValue = NaN; highestTime = 0; highestValue = NaN;
    correctionMeta = DoubleCorrection{}; tsVal = NaN;
    mnL = INT64_MAX; mxL = INT64_MIN;
    squaredSum = NaN;
    if (longCol && (fn == FN_STDDEV_OVER_TIME || fn == FN_STDVAR_OVER_TIME)) { sum = 0; squaredSum = 0; }    // VarOverTimeChunkedFunctionL.reset :1147
    changes = NaN; prev = NaN; quantileResult = NaN; values.clear();
    b0 = NaN; s0 = NaN; nextvalue = NaN; smoothedResult = NaN;
    sumX = NaN; sumY = NaN; sumXY = NaN; sumX2 = NaN; counter = 0;
  }
  bool isCounterPath() const { return (fn == FN_RATE || fn == FN_INCREASE) ? cumulative : (fn == FN_DELTA); }

  void addSum(DoubleReader& r, int s, int e) {          // AggrOverTimeFunctions.scala:560-571
    double chunkSum = r.sum(s, e);
    if (!std::isnan(chunkSum) && std::isnan(sum)) sum = 0;
    sum += chunkSum;
  }
  // ---- Long-column variants (ChunkedLongRangeFunction, RangeFunction.scala:224-242)
  void addTimeLongChunks(InfoReader& ir, int s, int e) {
    const LongReader& lr = ir.valLong; Ptr v = ir.valVec;
    switch (fn) {
      case FN_SUM_OVER_TIME: if (std::isnan(sum)) sum = 0; sum += lr.sum(v, s, e); break;                  // :574-585
      case FN_AVG_OVER_TIME: sum += lr.sum(v, s, e); count += (e - s + 1); break;                          // :1019-1028 (sum starts as NaN and stays NaN)
      case FN_COUNT_OVER_TIME: count += (e - s + 1); break;                                                // CountOverTimeChunkedFunction :924-938
      case FN_MIN_OVER_TIME: for (int r = s; r <= e; ++r) mnL = std::min(mnL, lr.apply(v, r)); break;      // :60-77
      case FN_MAX_OVER_TIME: for (int r = s; r <= e; ++r) mxL = std::max(mxL, lr.apply(v, r)); break;      // :99-116
      case FN_STDDEV_OVER_TIME: case FN_STDVAR_OVER_TIME: {                                                // :1144-1167
        double _sum = 0, _sqSum = 0;
        for (int r = s; r <= e; ++r) { double nextValue = (double)lr.apply(v, r); _sum += nextValue; _sqSum += nextValue * nextValue; }
        count += (e - s + 1); sum += _sum; squaredSum += _sqSum;
        break;
      }
      case FN_CHANGES: {                                                                                   // :1211-1225
        if (std::isnan(changes)) changes = 0;
        auto c = lr.changes(v, s, e, d2l(prev));
        changes += (double)c.first; prev = (double)c.second;
        break;
      }
      case FN_QUANTILE_OVER_TIME:                                                                          // :1322-1344
        if (p0 < 0) quantileResult = -std::numeric_limits<double>::infinity();
        else if (p0 > 1) quantileResult = std::numeric_limits<double>::infinity();
        else for (int r = s; r <= e; ++r) values.push_back((double)lr.apply(v, r));
        break;
      case FN_MAD_OVER_TIME: for (int r = s; r <= e; ++r) values.push_back((double)lr.apply(v, r)); break; // :1346-1359
      default: break;
    }
  }
  void addVar(DoubleReader& r, int s, int e) {           // VarOverTimeChunkedFunctionD.addTimeDoubleChunks :1087-1115
    double chunkSum = NaN, chunkSquaredSum = NaN; int chunkCount = 0;
    for (int elemNo = s; elemNo <= e; ++elemNo) {
      double nextValue = r.apply(elemNo);
      if (!std::isnan(nextValue)) {
        if (std::isnan(chunkSum)) chunkSum = 0;
        if (std::isnan(chunkSquaredSum)) chunkSquaredSum = 0;
        if (elemNo == e) lastSample = nextValue;
        chunkSum += nextValue;
        chunkSquaredSum += nextValue * nextValue;
        chunkCount += 1;
      }
    }
    if (!std::isnan(chunkSum) && std::isnan(sum)) sum = 0;
    sum += chunkSum;
    if (!std::isnan(chunkSquaredSum) && std::isnan(squaredSum)) squaredSum = 0;
    squaredSum += chunkSquaredSum;
    count += chunkCount;
  }
  // HoltWintersChunkedFunctionD.addTimeDoubleChunks :1413-1452.  The reference reads its value iterator one element past
  // endRowNum on every chunk and, on a continuation chunk, drops that chunk's first row in favour of the stale over-read
  // (undefined memory past the previous vector).  That is not reproducible; restated here with the evident intent: every
  // row of the window is consumed once, in order.  Single-chunk windows (all of the reference's tests) are literal.
  void addHoltWinters(DoubleReader& r, int s, int e) {
    const double sf = p0, tf = p1;
    int pos = s;                                           // iterator position
    auto getNextValue = [&](int startRowNum, double& res) -> int {     // :1399-1411
      res = NaN; int currRowNum = startRowNum;
      while (currRowNum <= e && std::isnan(res)) { double nv = r.apply(pos++); if (!std::isnan(nv)) res = nv; currRowNum += 1; }
      return currRowNum;
    };
    int rowNum = s;
    if (std::isnan(s0) && std::isnan(b0)) {
      double _s0, _b0; int firstrow = getNextValue(s, _s0); int currRow = getNextValue(firstrow, _b0);
      nextvalue = _b0; b0 = _b0 - _s0; rowNum = currRow - 1; s0 = _s0;
    } else if (std::isnan(b0)) {
      double _b0; int currRow = getNextValue(s, _b0);
      nextvalue = _b0; b0 = _b0 - s0; rowNum = currRow - 1;
    } else {
      nextvalue = r.apply(pos++);                          // (intent) first row of the continuation chunk
    }
    if (!std::isnan(b0)) {
      while (rowNum <= e) {
        if (!std::isnan(nextvalue)) {
          double _s0 = sf * nextvalue + (1 - sf) * (s0 + b0);
          b0 = tf * (_s0 - s0) + (1 - tf) * b0;
          s0 = _s0;
        }
        nextvalue = (pos <= e) ? r.apply(pos) : NaN; pos++;      // the reference's over-read past endRowNum is never used
        rowNum += 1;
      }
      smoothedResult = s0;
    }
  }
  void addChunks(InfoReader& ir, int64_t startTime, int64_t endTime) {
    LongReader& ts = ir.tsReader; DoubleReader& val = ir.valReader;
    if (fn == FN_LAST || fn == FN_TIMESTAMP || fn == FN_PRESENT_OVER_TIME) {            // RangeFunction.scala:603-613 / :708-716 / :725-748
      int32_t endRowNum = std::min(ts.ceilingIndex(ir.tsVec, endTime), ir.numRows() - 1);
      if (endRowNum >= 0) {
        int64_t t = ts.apply(ir.tsVec, endRowNum);
        if (fn == FN_TIMESTAMP) { tsVal = (double)t / (double)1000.0f; }
        else if (t >= startTime && t > lastTs) {
          if (fn == FN_LAST) { lastTs = t; lastVal = longCol ? (double)ir.valLong.apply(ir.valVec, endRowNum) : val.apply(endRowNum); }
          else {                                             // PresentOverTimeChunkedFunctionD.updateValue
            double doubleVal = val.apply(endRowNum);
            if (std::isnan(doubleVal)) {
              if (endRowNum > 0) { lastTs = t; double lv = val.apply(endRowNum - 1); lastVal = std::isnan(lv) ? NaN : 1.0; }
            } else { lastTs = t; lastVal = 1.0; }
          }
        }
      }
      return;
    }
    int32_t startRowNum = ts.binarySearch(ir.tsVec, startTime) & 0x7fffffff;       // RangeFunction.scala:185-190 / :141-142
    int32_t endRowNum = std::min(ts.ceilingIndex(ir.tsVec, endTime), ir.numRows() - 1);
    if (fn == FN_PREDICT_LINEAR) {                          // PredictLinearChunkedFunctionD/L.addChunks :1524-1553 / :1560-1588
      for (int r = startRowNum; r <= endRowNum; ++r) {
        double nextvalue_ = longCol ? (double)ir.valLong.apply(ir.valVec, r) : val.apply(r);
        int64_t nexttime = ts.apply(ir.tsVec, r);
        if (longCol || !std::isnan(nextvalue_)) {
          double x = (double)(nexttime - endTime) / 1000.0;
          if (std::isnan(sumY)) { sumY = nextvalue_; sumX = x; sumXY = x * nextvalue_; sumX2 = x * x; }
          else { sumY += nextvalue_; sumX += x; sumXY += x * nextvalue_; sumX2 += x * x; }
          counter += 1;
        }
      }
      return;
    }
    if (isCounterPath()) {                                 // CounterChunkedRangeFunction.addChunks, RangeFunction.scala:138-163
      correctionMeta = val.detectDropAndCorrection(correctionMeta);
      if (startRowNum <= endRowNum)
        addTimeChunksRate(val, startRowNum, endRowNum, ts.apply(ir.tsVec, startRowNum), ts.apply(ir.tsVec, endRowNum));
      correctionMeta = val.updateCorrection(correctionMeta);
      return;
    }
    if (!(startRowNum <= endRowNum)) return;
    if (longCol) { addTimeLongChunks(ir, startRowNum, endRowNum); return; }
    switch (fn) {
      case FN_RATE: case FN_INCREASE: case FN_SUM_OVER_TIME: addSum(val, startRowNum, endRowNum); break;   // delta branch: RateFunctions.scala:424-445
      case FN_AVG_OVER_TIME: addSum(val, startRowNum, endRowNum); count += val.count(startRowNum, endRowNum); break;
      case FN_COUNT_OVER_TIME: if (std::isnan(countD)) countD = 0; countD += val.count(startRowNum, endRowNum); break;
      case FN_MIN_OVER_TIME: for (int r = startRowNum; r <= endRowNum; ++r) mn = minIgnoreNaN(mn, val.apply(r)); break;
      case FN_MAX_OVER_TIME: for (int r = startRowNum; r <= endRowNum; ++r) mx = maxIgnoreNaN(mx, val.apply(r)); break;
      case FN_STDDEV_OVER_TIME: case FN_STDVAR_OVER_TIME: case FN_ZSCORE: addVar(val, startRowNum, endRowNum); break;
      case FN_CHANGES: {                                     // ChangesChunkedFunctionD :1193-1208
        if (std::isnan(changes)) changes = 0;
        auto c = val.changes(startRowNum, endRowNum, prev);
        changes += c.first; prev = c.second;
        break;
      }
      case FN_QUANTILE_OVER_TIME:                            // QuantileOverTimeChunkedFunctionD :1272-1300
        if (p0 < 0) quantileResult = -std::numeric_limits<double>::infinity();
        else if (p0 > 1) quantileResult = std::numeric_limits<double>::infinity();
        else for (int r = startRowNum; r <= endRowNum; ++r) { double nv = val.apply(r); if (!std::isnan(nv)) values.push_back(nv); }
        break;
      case FN_MAD_OVER_TIME:                                 // :1302-1320
        for (int r = startRowNum; r <= endRowNum; ++r) { double nv = val.apply(r); if (!std::isnan(nv)) values.push_back(nv); }
        break;
      case FN_HOLT_WINTERS: addHoltWinters(val, startRowNum, endRowNum); break;
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
  double apply(int64_t windowStart, int64_t windowEnd) {
    switch (fn) {
      case FN_LAST: case FN_PRESENT_OVER_TIME: return lastVal;
      case FN_TIMESTAMP: return tsVal;
      case FN_SUM_OVER_TIME: return sum;
      case FN_AVG_OVER_TIME: return count > 0 ? sum / count : (std::isnan(sum) ? sum : 0.0);    // AggrOverTimeFunctions.scala:1000
      case FN_COUNT_OVER_TIME: return longCol ? (double)count : countD;
      case FN_MIN_OVER_TIME: return longCol ? (double)mnL : mn;
      case FN_MAX_OVER_TIME: return longCol ? (double)mxL : mx;
      case FN_STDDEV_OVER_TIME: case FN_STDVAR_OVER_TIME: {
        if (longCol) {                                      // StdDev/StdVarOverTimeChunkedFunctionL :1169-1183
          double avg = count > 0 ? sum / count : 0.0;
          double var = squaredSum / count - avg * avg;
          return fn == FN_STDDEV_OVER_TIME ? std::sqrt(var) : var;
        }
        if (count > 0) {                                    // :1118-1142
          double avg = sum / count;
          double var = (squaredSum / count) - (avg * avg);
          return fn == FN_STDDEV_OVER_TIME ? std::sqrt(var) : var;
        }
        return std::isnan(sum) ? sum : 0.0;
      }
      case FN_ZSCORE: {                                     // :1592-1604
        if (count > 0) { double avg = sum / count; double stdDev = std::sqrt(squaredSum / count - avg * avg); return (lastSample - avg) / stdDev; }
        return std::isnan(sum) ? sum : 0.0;
      }
      case FN_CHANGES: return changes;
      case FN_QUANTILE_OVER_TIME: {                         // QuantileOverTimeChunkedFunction.apply :1232-1244
        int cnt = (int)values.size();
        std::sort(values.begin(), values.end(), javaDoubleLess);
        double weight; int upperIndex, lowerIndex; calculateRank(p0, cnt, weight, upperIndex, lowerIndex);
        if (cnt > 0) quantileResult = values[lowerIndex] * (1 - weight) + values[upperIndex] * weight;
        return quantileResult;
      }
      case FN_MAD_OVER_TIME: {                              // MedianAbsoluteDeviationOverTimeChunkedFunction.apply :1252-1268
        int size = (int)values.size();
        double weight; int upperIndex, lowerIndex; calculateRank(0.5, size, weight, upperIndex, lowerIndex);
        std::sort(values.begin(), values.end(), javaDoubleLess);
        double res = NaN;
        if (size > 0) {
          double median = values[lowerIndex] * (1 - weight) + values[upperIndex] * weight;
          std::vector<double> diff; diff.reserve(values.size());
          for (double v : values) diff.push_back(std::fabs(median - v));
          std::sort(diff.begin(), diff.end(), javaDoubleLess);
          res = diff[lowerIndex] * (1 - weight) + diff[upperIndex] * weight;
        }
        return res;
      }
      case FN_HOLT_WINTERS: return smoothedResult;
      case FN_PREDICT_LINEAR: {                             // PredictLinearChunkedFunction.apply :1507-1517
        double covXY = sumXY - sumX * sumY / counter;
        double varX = sumX2 - sumX * sumX / counter;
        double slope = covXY / varX;
        double intercept = sumY / counter - slope * sumX / counter;
        return counter >= 2 ? slope * p0 + intercept : NaN;
      }
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
                            const QueryConfig& cfg, double* out /*[T]*/, QueryStats* stats = nullptr, double p0 = 0, double p1 = 0) {
  if (s.longCol && !longColumnSupports(fn)) throw std::invalid_argument("no chunked function for this range function on a Long column");
  WindowedChunkIterator wit(s, start, step, end, window, cfg.inclusiveRange);
  ChunkedFn f{fn, cumulative, cfg.inclusiveRange};
  f.longCol = s.longCol; f.p0 = p0; f.p1 = p1;
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
