// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the result wire format of a (timestamp, double) query result:
//   RecordBuilder (startNewRecord / addLong / addDouble / endRecord / requireBytes / newContainer)
//       core/src/main/scala/filodb.core/binaryrecord2/RecordBuilder.scala:109-175,461-480,589-621,645-649
//   RecordContainer (length word, version word, timestamp; iterate)   core/src/main/scala/filodb.core/binaryrecord2/RecordContainer.scala:13-88
//   RecordSchema field offsets                                        core/src/main/scala/filodb.core/binaryrecord2/RecordSchema.scala:65-71
//   SerializedRangeVector.apply / rows / canRemoveEmptyRows           core/src/main/scala/filodb.core/query/RangeVector.scala:427-476,511-575
#pragma once
#include "filo_format.hpp"

namespace fo {
namespace result {

constexpr int ContainerHeaderLen = 16;         // RecordBuilder.scala:648
constexpr int Version = 1;                     // :647
constexpr int MaxContainerSize = 4096;         // SerializedRangeVector.MaxContainerSize, RangeVector.scala:586

// RecordSchema(Seq(TimestampColumn, DoubleColumn)) without partition fields: fixedStart = 4, offsets = [4, 12, 20], variableAreaStart = 20
struct TsDoubleSchema { static constexpr int fieldOffset(int i) { return i == 0 ? 4 : (i == 1 ? 12 : 20); } static constexpr int variableAreaStart = 20; static constexpr int numFields = 2; };

struct RecordContainer { std::vector<uint8_t> bytes; int numRecords = 0; int numBytes() const { return getInt(bytes.data()); } };

struct RecordBuilder {
  int containerSize; int64_t now;
  std::vector<std::unique_ptr<RecordContainer>> containers;
  int64_t curRecordOffset = -1, curRecEndOffset = -1, maxOffset = -1;      // offsets inside containers.back()
  int fieldNo = -1;
  explicit RecordBuilder(int size = MaxContainerSize, int64_t nowMs = 0) : containerSize(size), now(nowMs) {}
  RecordContainer* currentContainer() { return containers.empty() ? nullptr : containers.back().get(); }
  void updateLength(int64_t endOffset) { setInt(containers.back()->bytes.data(), (int32_t)(endOffset - 4)); }
  void newContainer() {                                                       // :609-621
    auto c = std::make_unique<RecordContainer>(); c->bytes.assign((size_t)containerSize, 0);
    containers.push_back(std::move(c));
    curRecordOffset = ContainerHeaderLen; curRecEndOffset = curRecordOffset;
    updateLength(curRecordOffset);
    setInt(containers.back()->bytes.data() + 4, Version << 24);               // writeVersionWord
    setLong(containers.back()->bytes.data() + 8, now);                         // updateTimestamp
    maxOffset = containerSize;
  }
  void requireBytes(int numBytes) {                                            // :589-607 (a started record never spills here: it is requested whole)
    if (containers.empty()) newContainer();
    else if (curRecEndOffset + numBytes > maxOffset) {
      if (!((containerSize - ContainerHeaderLen) > numBytes)) throw std::invalid_argument("The intermediate or final result is too big");
      newContainer();
    }
  }
  void startNewRecord() {                                                      // :109-125
    requireBytes(TsDoubleSchema::variableAreaStart);
    setInt(containers.back()->bytes.data() + curRecordOffset, TsDoubleSchema::variableAreaStart - 4);
    curRecEndOffset = curRecordOffset + TsDoubleSchema::variableAreaStart;
    fieldNo = 0;
  }
  void addLong(int64_t v) { setLong(containers.back()->bytes.data() + curRecordOffset + TsDoubleSchema::fieldOffset(fieldNo), v); fieldNo += 1; }      // :156-160
  void addDouble(double v) { setDouble(containers.back()->bytes.data() + curRecordOffset + TsDoubleSchema::fieldOffset(fieldNo), v); fieldNo += 1; }   // :167-171
  void endRecord() {                                                           // :461-478
    curRecEndOffset = (curRecEndOffset + 3) & ~(int64_t)3;
    curRecordOffset = curRecEndOffset; fieldNo = -1;
    updateLength(curRecEndOffset);
    containers.back()->numRecords += 1;
  }
};

struct SerializedRangeVector { int32_t numRowsSerialized = 0; int32_t startRecordNo = 0; int64_t firstContainer = 0; };

inline bool canRemoveEmptyRows(int64_t startMs, int64_t endMs) { return startMs != endMs; }      // (time series schema, 2 columns, double values)

// SerializedRangeVector.apply for one range vector whose rows are (start + k * step, values[k]), sharing `builder`
inline SerializedRangeVector serialize(RecordBuilder& builder, const double* values, int T, int64_t startMs, int64_t stepMs, int64_t endMs) {
  SerializedRangeVector srv;
  RecordContainer* old = builder.currentContainer();
  srv.startRecordNo = old ? old->numRecords : 0;
  srv.firstContainer = old ? (int64_t)builder.containers.size() - 1 : 0;
  for (int k = 0; k < T; ++k) {
    const double v = values[k];
    if (!canRemoveEmptyRows(startMs, endMs) || !std::isnan(v)) {
      srv.numRowsSerialized += 1;
      builder.startNewRecord(); builder.addLong(startMs + (int64_t)k * stepMs); builder.addDouble(v); builder.endRecord();
    }
  }
  return srv;
}

// SerializedRangeVector.rows: records [startRecordNo, startRecordNo + numRowsSerialized) of the containers from the first one on, with the
// NaN rows put back on the step grid (RangeVector.scala:445-476).  containers: concatenated 4096-byte containers.
inline void rows(const uint8_t* containers, int64_t nContainers, const SerializedRangeVector& srv, int64_t startMs, int64_t stepMs, int64_t endMs,
                 std::vector<int64_t>& ts, std::vector<double>& vals) {
  std::vector<std::pair<int64_t, double>> recs;
  int64_t idx = 0;
  for (int64_t c = srv.firstContainer; c < nContainers; ++c) {
    const uint8_t* base = containers + c * MaxContainerSize;
    const int64_t endOffset = 4 + getInt(base);
    int64_t cur = ContainerHeaderLen;
    while (cur < endOffset) {                                                  // RecordContainer.iterate :75-88
      const int32_t recordLen = getInt(base + cur);
      if (idx >= srv.startRecordNo && idx < (int64_t)srv.startRecordNo + srv.numRowsSerialized) recs.emplace_back(getLong(base + cur + 4), getDouble(base + cur + 12));
      ++idx;
      cur += (recordLen + 7) & ~3;
    }
  }
  ts.clear(); vals.clear();
  if (!canRemoveEmptyRows(startMs, endMs)) { for (auto& r : recs) { ts.push_back(r.first); vals.push_back(r.second); } return; }
  size_t p = 0;
  for (int64_t t = startMs; t <= endMs; t += stepMs) {
    if (p < recs.size() && recs[p].first == t) { ts.push_back(t); vals.push_back(recs[p].second); ++p; }
    else { ts.push_back(t); vals.push_back(std::numeric_limits<double>::quiet_NaN()); }
  }
}

} // namespace result
} // namespace fo
