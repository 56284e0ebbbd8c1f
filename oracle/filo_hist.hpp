// TEST INFRASTRUCTURE ONLY (see oracle/filo_format.hpp): CPU restatement of FiloDB's histogram column path.
// Histogram buckets + BinaryHistogram blobs, section-based HistogramVectors (simple and SectDelta), their readers with
// counter correction, the histogram range functions, HistSum row aggregation and histogram_quantile.
// Every function cites the reference file:line it restates (paths under core/src/main/scala/filodb.memory/format/ and
// query/src/main/scala/filodb/query/exec/).  Otel exponential buckets (Base2ExpHistogramBuckets, ExpHistogramVector.scala) are restated
// for the value/bucket arithmetic, the row vector and sum(); the XOR (double-valued) blob formats are not.
#pragma once
#include "filo_format.hpp"
#include "filo_query.hpp"
#include <cmath>

namespace fo {
namespace hist {

// BinaryHistogram format codes, vectors/HistogramVector.scala:136-143
enum : uint8_t { FMT_NULL = 0x00, FMT_GEO_DELTA = 0x03, FMT_GEO1_DELTA = 0x04, FMT_CUSTOM_DELTA = 0x05, FMT_OTEL_DELTA = 0x09,
                 FMT_GEO_XOR = 0x08, FMT_CUSTOM_XOR = 0x0a, FMT_OTEL_XOR = 0x10 };
constexpr int WIRE_H_SIMPLE = (0x10 << 8) | 0x09, WIRE_H_SECTDELTA = (0x12 << 8) | 0x09, WIRE_H_EXP_SIMPLE = (0x13 << 8) | 0x09;     // WireFormat.scala:17,35-38
constexpr int OffsetNumHistograms = 6, OffsetFormatCode = 8, OffsetBucketDefSize = 9, OffsetBucketDef = 11;   // HistogramVector.scala:239-244

// ---------------------------------------------------------------------------------------------------------------------
// HistogramBuckets: GeometricBuckets (Histogram.scala:601-626), CustomBuckets (:874-899)
// ---------------------------------------------------------------------------------------------------------------------
struct Buckets {
  enum Kind { EMPTY = 0, GEOMETRIC = 1, CUSTOM = 2, EXP = 3 } kind = EMPTY;
  double first = 0, mult = 0; bool minusOne = false; int n = 0;
  std::vector<double> les;
  int scale = 0, startIdx = 0;            // EXP: Base2ExpHistogramBuckets(scale, startIndexPositiveBuckets, numPositiveBuckets = n - 1)
  static constexpr int maxPositiveBuckets = 180, maxAbsScale = 20;          // Histogram.scala:639,645
  // Base2ExpHistogramBuckets.base / logBase, Histogram.scala:647-658 (tables of Math.pow(2, Math.pow(2, -scale)) and its Math.log)
  static double expBase(int sc) {
    if (sc < -maxAbsScale || sc > maxAbsScale) throw std::invalid_argument("requirement failed: Invalid scale");
    return std::pow(2.0, std::pow(2.0, (double)-sc));
  }
  static double expLogBase(int sc) { return std::log(expBase(sc)); }
  int numBuckets() const { return n; }
  int numPositive() const { return n - 1; }
  double bucketTop(int no) const {
    if (kind == CUSTOM) return les[(size_t)no];
    if (kind == EXP) {                                                       // Histogram.scala:716-727
      if (no == 0) return 0.0;
      const int index = startIdx + no - 1;
      return std::exp((double)(index + 1) * expLogBase(scale));
    }
    return (first * std::pow(mult, (double)no)) + (minusOne ? -1.0 : 0.0);     // Histogram.scala:606
  }
  double startBucketTop() const { return bucketTop(1); }                     // :695-696
  double endBucketTop() const { return bucketTop(n - 1); }
  int bucketIndexToArrayIndex(int index) const { return index - startIdx + 1; }   // :803
  bool canAccommodate(const Buckets& o) const { return endBucketTop() >= o.endBucketTop() && startBucketTop() <= o.startBucketTop(); }   // :767-770
  // Base2ExpHistogramBuckets.add, :772-795: the scheme that covers both ranges, scale reduced until it fits maxPosBuckets
  Buckets expAdd(const Buckets& o, int maxPosBuckets = maxPositiveBuckets) const {
    if (canAccommodate(o)) return *this;
    const double minTop = std::min(startBucketTop(), o.startBucketTop()), maxTop = std::max(endBucketTop(), o.endBucketTop());
    int newScale = std::min(scale, o.scale);
    double newBase = std::max(expBase(scale), expBase(o.scale));
    auto toInt = [](double d) { return std::isnan(d) ? 0 : d >= 2147483647.0 ? INT32_MAX : d <= -2147483648.0 ? INT32_MIN : (int)d; };   // Double.toInt
    int idxEnd = toInt(std::ceil(std::log(maxTop) / std::log(newBase))) - 1;
    int idxStart = toInt(std::floor(std::log(minTop) / std::log(newBase))) - 1;
    while (idxEnd - idxStart + 1 > maxPosBuckets) {
      newScale -= 1;
      newBase = expBase(newScale);
      idxEnd = toInt(std::ceil(std::log(maxTop) / std::log(newBase))) - 1;
      idxStart = toInt(std::floor(std::log(minTop) / std::log(newBase))) - 1;
    }
    return exponential(newScale, idxStart, idxEnd - idxStart + 1);
  }
  // addValues, :810-866: fold a histogram of a finer (or equal) scheme into values laid out for this scheme
  template <class H> void expAddValues(std::vector<double>& ourValues, const Buckets& otherBuckets, const H& other) const {
    if ((int)ourValues.size() != n || other.numBuckets() != otherBuckets.n || !canAccommodate(otherBuckets)) throw std::invalid_argument("requirement failed");
    const int scaleIncrease = otherBuckets.scale - scale;
    const int fac = (int)std::pow(2.0, (double)scaleIncrease);
    ourValues[0] += other.bucketValue(0);
    for (int ourBucketIndex = startIdx; ourBucketIndex < startIdx + numPositive(); ++ourBucketIndex) {
      const int ourPlus1 = ourBucketIndex + 1;
      const int otherPlus1 = (int)((int64_t)ourPlus1 * fac);               // Int multiply (wraps like the JVM's)
      const int ourArrayIndex = bucketIndexToArrayIndex(ourPlus1 - 1);
      const int otherArrayIndex = otherBuckets.bucketIndexToArrayIndex(otherPlus1 - 1);
      if (otherArrayIndex > 0 && otherArrayIndex < otherBuckets.n) ourValues[(size_t)ourArrayIndex] += other.bucketValue(otherArrayIndex);
      else if (otherArrayIndex >= otherBuckets.n) {
        if (ourArrayIndex == 0) throw std::invalid_argument("requirement failed: double counting zero bucket");
        ourValues[(size_t)ourArrayIndex] += other.bucketValue(otherBuckets.n - 1);
      }
    }
  }
  bool operator==(const Buckets& o) const {
    if (kind != o.kind) return false;
    if (kind == GEOMETRIC) return first == o.first && mult == o.mult && n == o.n && minusOne == o.minusOne;
    if (kind == CUSTOM) return les == o.les;
    if (kind == EXP) return scale == o.scale && startIdx == o.startIdx && n == o.n;
    return true;
  }
  bool operator!=(const Buckets& o) const { return !(*this == o); }
  // similarForMath, Histogram.scala:620-625, 893-898: equal, or one geometric and one custom with the same tops
  bool similarForMath(const Buckets& o) const {
    if (kind == EXP || o.kind == EXP) return kind == o.kind && *this == o;   // Histogram.scala:758-765
    if (kind != o.kind && kind != EMPTY && o.kind != EMPTY) {
      if (n != o.n) return false;
      for (int i = 0; i < n; ++i) if (bucketTop(i) != o.bucketTop(i)) return false;
      return true;
    }
    return *this == o;
  }
  uint8_t deltaFormat() const { return kind == GEOMETRIC ? (minusOne ? FMT_GEO1_DELTA : FMT_GEO_DELTA) : kind == CUSTOM ? FMT_CUSTOM_DELTA : kind == EXP ? FMT_OTEL_DELTA : FMT_NULL; }   // HistogramVector.scala:192-198
  // serialize at pos; returns the position after the definition.  Geometric: Histogram.scala:609-617; Custom: :878-884
  int serialize(std::vector<uint8_t>& buf, int pos) const {
    if (kind == GEOMETRIC) {
      if ((int)buf.size() < pos + 20) buf.resize((size_t)pos + 20);
      setShort(&buf[(size_t)pos], (int16_t)(2 + 8 + 8));
      setShort(&buf[(size_t)pos + 2], (int16_t)n);
      setDouble(&buf[(size_t)pos + 4], first); setDouble(&buf[(size_t)pos + 12], mult);
      return pos + 2 + 2 + 8 + 8;
    }
    if (kind == CUSTOM) {
      if ((int)buf.size() < pos + 4) buf.resize((size_t)pos + 4);
      setShort(&buf[(size_t)pos + 2], (int16_t)les.size());
      const int finalPos = nibble::packDoubles(les.data(), (int)les.size(), buf, pos + 4);
      setShort(&buf[(size_t)pos], (int16_t)(finalPos - pos - 2));
      return finalPos;
    }
    if (kind == EXP) {                                                       // Histogram.scala:729-752
      if ((int)buf.size() < pos + 18) buf.resize((size_t)pos + 18);
      setShort(&buf[(size_t)pos], (int16_t)(2 + 2 + 4 + 2 + 4 + 2));
      setShort(&buf[(size_t)pos + 2], (int16_t)n);
      setShort(&buf[(size_t)pos + 4], (int16_t)scale);
      setInt(&buf[(size_t)pos + 6], startIdx);
      setShort(&buf[(size_t)pos + 10], (int16_t)numPositive());
      setInt(&buf[(size_t)pos + 12], 0);                                     // startIndexNegativeBuckets
      setShort(&buf[(size_t)pos + 16], 0);                                   // numNegativeBuckets
      return pos + 18;
    }
    throw std::invalid_argument("serialize: empty buckets");
  }
  // HistogramBuckets.apply(acc, bucketsDef, formatCode), Histogram.scala:541-547: `def` points at the u16 length prefix
  static Buckets parse(Ptr def, uint8_t formatCode) {
    Buckets b;
    if (formatCode == FMT_GEO_DELTA || formatCode == FMT_GEO1_DELTA || formatCode == FMT_GEO_XOR) {
      b.kind = GEOMETRIC; b.minusOne = formatCode == FMT_GEO1_DELTA;
      b.n = getShort(def + 2); b.first = getDouble(def + 4); b.mult = getDouble(def + 12);       // :551-555
    } else if (formatCode == FMT_CUSTOM_DELTA || formatCode == FMT_CUSTOM_XOR) {
      b.kind = CUSTOM; b.n = getShort(def + 2) & 0xffff;                                          // :583-592
      b.les.assign((size_t)b.n, 0.0);
      const int cap = (getShort(def) & 0xffff) - 2;
      if (b.n > 0 && nibble::unpackDoubleXOR(def + 4, cap, b.les.data(), b.n) != nibble::Ok) throw CorruptVector("custom buckets: input too short");
    } else if (formatCode == FMT_OTEL_DELTA || formatCode == FMT_OTEL_XOR) {                      // :557-565
      b = exponential(getShort(def + 4), getInt(def + 6), getShort(def + 10));
    }
    return b;
  }
  static Buckets geometric(double first, double mult, int n, bool minusOne = false) { Buckets b; b.kind = GEOMETRIC; b.first = first; b.mult = mult; b.n = n; b.minusOne = minusOne; return b; }
  static Buckets exponential(int scale, int startIdx, int numPos) {           // constructor requirements, Histogram.scala:688-692
    if (!(numPos <= maxPositiveBuckets && numPos >= 0)) throw std::invalid_argument("requirement failed: Invalid buckets: numPositiveBuckets");
    if (scale < -maxAbsScale || scale > maxAbsScale) throw std::invalid_argument("requirement failed: Invalid scale");
    Buckets b; b.kind = EXP; b.scale = scale; b.startIdx = startIdx; b.n = numPos + 1; return b;
  }
  static Buckets custom(const double* les, int n) { Buckets b; b.kind = CUSTOM; b.n = n; b.les.assign(les, les + n); return b; }
};

// ---------------------------------------------------------------------------------------------------------------------
// Histogram values.  LongHistogram (Histogram.scala:262-318) and MutableHistogram (:319-470) share this shape; `empty()`
// (numBuckets == 0) is Histogram.empty / HistogramWithBuckets.empty.
// ---------------------------------------------------------------------------------------------------------------------
struct LongHist {
  Buckets buckets; std::vector<int64_t> values;
  int numBuckets() const { return buckets.n; }
  double bucketValue(int no) const { return (double)values[(size_t)no]; }
  void add(const LongHist& o) {                                             // LongHistogram.add, :271-283
    if (o.buckets != buckets) throw std::invalid_argument("Cannot add histograms with different bucket configurations.");
    for (int b = 0; b < numBuckets(); ++b) values[(size_t)b] = ladd(values[(size_t)b], o.values[(size_t)b]);
  }
  static LongHist empty(const Buckets& b) { LongHist h; h.buckets = b; h.values.assign((size_t)b.n, 0); return h; }
};

struct MutHist {
  Buckets buckets; std::vector<double> values;
  int numBuckets() const { return buckets.n; }
  bool isEmpty() const { return buckets.n == 0; }
  double bucketTop(int no) const { return buckets.bucketTop(no); }
  double bucketValue(int no) const { return values[(size_t)no]; }
  static MutHist emptyNaN(const Buckets& b) { MutHist h; h.buckets = b; h.values.assign((size_t)b.n, NaN); return h; }   // MutableHistogram.empty, :454-455
  static MutHist from(const LongHist& l) { MutHist h; h.buckets = l.buckets; h.values.resize(l.values.size()); for (size_t i = 0; i < l.values.size(); ++i) h.values[i] = (double)l.values[i]; return h; }
  // addNoCorrection, :367-421 (same-scheme branch, the exponential-scheme branch and the mismatch branch)
  template <class H> bool addNoCorrection(const H& o) {
    if (buckets.similarForMath(o.buckets)) {
      if (numBuckets() > 0 && std::isnan(values[0])) std::fill(values.begin(), values.end(), 0.0);
      for (int b = 0; b < numBuckets(); ++b) values[(size_t)b] += o.bucketValue(b);
      return true;
    }
    if (buckets.kind == Buckets::EXP && o.buckets.kind == Buckets::EXP) {    // :376-404
      if (std::isnan(values[0])) std::fill(values.begin(), values.end(), 0.0);
      if (buckets.canAccommodate(o.buckets)) {
        buckets.expAddValues(values, o.buckets, o);
      } else if (buckets.numPositive() == 0) {                               // we are zero-only: take the other scheme
        const double zero = values[0];
        buckets = o.buckets;
        values.resize((size_t)o.numBuckets());
        for (int b = 0; b < o.numBuckets(); ++b) values[(size_t)b] = o.bucketValue(b);
        values[0] += zero;
      } else if (o.buckets.numPositive() == 0) {
        values[0] += o.bucketValue(0);
      } else {
        const Buckets nb = buckets.expAdd(o.buckets);
        std::vector<double> nv((size_t)nb.n, 0.0);
        nb.expAddValues(nv, buckets, *this);
        nb.expAddValues(nv, o.buckets, o);
        buckets = nb; values = std::move(nv);
      }
      return false;
    }
    for (int b = 0; b < numBuckets(); ++b) values[(size_t)b] = NaN;
    return false;
  }
  template <class H> void add(const H& o) { if (addNoCorrection(o)) makeMonotonic(); }              // :428-432
  void makeMonotonic() {                                                                            // :440-449
    double mx = 0.0;
    for (size_t b = 0; b < values.size(); ++b) {
      if (values[b] < mx || std::isnan(values[b])) values[b] = mx;
      else if (values[b] > mx) mx = values[b];
    }
  }
  double topBucketValue() const { return numBuckets() <= 0 ? NaN : bucketValue(numBuckets() - 1); }  // Histogram.scala:50-51
  int firstBucketGTE(double rank) const { int b = 0; while (bucketValue(b) < rank) ++b; return b; } // :44-48
  // Histogram.quantile, :65-108 (min = 0, max = +Inf, evenDistribution = false)
  double quantile(double q) const {
    if (q < 0) return -std::numeric_limits<double>::infinity();
    if (q > 1) return std::numeric_limits<double>::infinity();
    if (numBuckets() < 2 || !(topBucketValue() > 0)) return NaN;            // `topBucketValue <= 0` is false for NaN in Scala too
    double rank = q * topBucketValue();
    const int bucket = firstBucketGTE(rank);
    double bucketStart = bucket == 0 ? 0.0 : bucketTop(bucket - 1);
    double bucketEnd = bucketTop(bucket);
    const double mn = 0.0, mxv = std::numeric_limits<double>::infinity();
    if (mn > bucketStart && mn <= bucketEnd) bucketStart = mn;
    if (mxv > bucketStart && mxv <= bucketEnd) bucketEnd = mxv;
    if (bucket == numBuckets() - 1 && std::isinf(bucketEnd) && bucketEnd > 0) return bucketTop(numBuckets() - 2);
    if (bucket == 0 && bucketTop(0) <= 0) return bucketTop(0);
    const double count = bucket == 0 ? bucketValue(bucket) : bucketValue(bucket) - bucketValue(bucket - 1);
    rank -= (bucket == 0 ? 0.0 : bucketValue(bucket - 1));
    const double fraction = rank / count;
    if (buckets.kind != Buckets::EXP || bucketStart == 0) return bucketStart + (bucketEnd - bucketStart) * fraction;
    auto log2j = [](double v) { return std::log(v) / std::log(2.0); };       // :111
    const double logEnd = log2j(bucketEnd), logStart = log2j(bucketStart);
    return std::pow(2.0, logStart + (logEnd - logStart) * fraction);
  }
};
// `topBucketValue <= 0` with NaN: Scala `NaN <= 0` is false, so a NaN top goes on to firstBucketGTE (NaN comparisons false
// -> bucket 0) and produces NaN through the arithmetic; the `!(top > 0)` shortcut above returns NaN directly for it.

// Histogram.compare, Histogram.scala:197-208 (Ordered: `a < b` is compare(a, b) < 0)
template <class A, class B> inline int compareHist(const A& a, const B& b) {
  auto cmp = [](double x, double y) { return x < y ? -1 : (x > y ? 1 : (x == y ? 0 : (std::isnan(x) ? (std::isnan(y) ? 0 : 1) : -1))); };   // java.lang.Double.compare
  const double ta = a.numBuckets() <= 0 ? NaN : a.bucketValue(a.numBuckets() - 1), tb = b.numBuckets() <= 0 ? NaN : b.bucketValue(b.numBuckets() - 1);
  if (a.numBuckets() != b.numBuckets()) return cmp(ta, tb);
  for (int i = 0; i < a.numBuckets(); ++i) if (a.buckets.bucketTop(i) != b.buckets.bucketTop(i)) return cmp(ta, tb);
  for (int i = a.numBuckets() - 1; i >= 0; --i) { const int c = cmp(a.bucketValue(i), b.bucketValue(i)); if (c != 0) return c; }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// BinaryHistogram blob: +0 u16 length (excl.), +2 u8 format, +3 u16 bucket-def bytes, +5 bucket def, values
// (HistogramVector.scala:18-52, writeDelta :171-198)
// ---------------------------------------------------------------------------------------------------------------------
namespace bin {
inline int totalLength(Ptr b) { return (getShort(b) & 0xffff) + 2; }
inline uint8_t formatCode(Ptr b) { return (uint8_t)b[2]; }
inline int bucketDefNumBytes(Ptr b) { return getShort(b + 3) & 0xffff; }
inline int numBuckets(Ptr b) { return getShort(b + 5); }
inline int valuesIndex(Ptr b) { return 2 + 3 + bucketDefNumBytes(b); }
inline int valuesNumBytes(Ptr b) { return totalLength(b) - valuesIndex(b); }
inline std::vector<uint8_t> writeDelta(const Buckets& buckets, const int64_t* values, int nvalues) {
  if (buckets.n > nvalues) throw std::invalid_argument("Values array size < numBuckets");
  std::vector<uint8_t> buf(64);
  const uint8_t fmt = buckets.n == 0 ? (uint8_t)FMT_NULL : buckets.deltaFormat();
  buf[2] = fmt;
  int finalPos = 3;
  if (fmt != FMT_NULL) { const int valuesIndex = buckets.serialize(buf, 3); finalPos = nibble::packDelta(values, buckets.n, buf, valuesIndex); }
  if (finalPos > 65535) throw std::invalid_argument("Histogram data is too large");
  if ((int)buf.size() < finalPos) buf.resize((size_t)finalPos);
  setShort(&buf[0], (int16_t)(finalPos - 2));
  buf.resize((size_t)finalPos);
  return buf;
}
inline bool isValidFormatCode(uint8_t c) { return c == FMT_NULL || c == FMT_GEO1_DELTA || c == FMT_GEO_DELTA || c == FMT_CUSTOM_DELTA || c == FMT_OTEL_DELTA || c == FMT_OTEL_XOR; }   // :145-149
// toHistogram for the delta formats, :84-96
inline LongHist toHistogram(Ptr b) {
  LongHist h; h.buckets = Buckets::parse(b + 3, formatCode(b));
  h.values.assign((size_t)h.buckets.n, 0);
  if (h.buckets.n > 0 && nibble::unpackDelta(b + valuesIndex(b), valuesNumBytes(b), h.values.data(), h.buckets.n) != nibble::Ok) throw CorruptVector("BinHistogram: input too short");
  return h;
}
} // namespace bin

// ---------------------------------------------------------------------------------------------------------------------
// Appenders: AppendableHistogramVector (HistogramVector.scala:326-436) and AppendableSectDeltaHistVector (:489-545), with
// the SectionWriter of Section.scala:91-145.  The vector is built in a byte array of maxBytes; bytes() returns the used
// prefix (numBytes + 4), which is what optimize()/freeze keep.
// ---------------------------------------------------------------------------------------------------------------------
enum AddResponse { Ack = 0, InvalidHistogram = 1, BucketSchemaMismatch = 2, VectorTooSmall = 3 };

struct HistAppender {
  bool sect; bool exp = false; int maxBytes; std::vector<uint8_t> v;
  int curSection = -1, bytesLeft = 0;                     // SectionWriter state (offsets into v)
  // DeltaSectDiffPackSink state, NibblePack.scala:296-345
  bool sinkInit = false; std::vector<int64_t> originalDeltas, lastHistDeltas;
  // expVector: AppendableExpHistogramVector (ExpHistogramVector.scala:37-121): every record is a whole BinaryHistogram blob
  HistAppender(bool sectDelta, int maxBytes_, bool expVector = false) : sect(sectDelta && !expVector), exp(expVector), maxBytes(maxBytes_), v((size_t)maxBytes_, 0) {
    setShort(&v[4], (int16_t)(exp ? WIRE_H_EXP_SIMPLE : sect ? WIRE_H_SECTDELTA : WIRE_H_SIMPLE));
    setShort(&v[OffsetNumHistograms], 0);
    setInt(&v[0], OffsetBucketDef + 2);                  // reset(): setNumBytes(OffsetNumBuckets + 2), :426-429
  }
  int maxElementsPerSection() const { return sect ? 16 : 64; }             // :340, :497; ExpHistogramVector.scala:51
  int length() const { return getShort(&v[OffsetNumHistograms]) & 0xffff; }
  int secNumBytes(int s) const { return getShort(&v[(size_t)s]) & 0xffff; }
  int secNumElems(int s) const { return v[(size_t)s + 2]; }
  int secEnd(int s) const { return s + 4 + secNumBytes(s); }
  void secInit(int s, int type) { v[(size_t)s + 2] = 0; v[(size_t)s + 3] = (uint8_t)type; setShort(&v[(size_t)s], 0); }
  bool needNewSection(int numBytes) const { return secNumElems(curSection) >= maxElementsPerSection() || secNumBytes(curSection) + numBytes >= 65536; }
  AddResponse addBlobInner(Ptr blob, int numBytes) {                        // Section.scala:132-143
    if (bytesLeft < numBytes + 2) return VectorTooSmall;
    const int w = secEnd(curSection);
    setShort(&v[(size_t)w], (int16_t)numBytes);
    std::memcpy(&v[(size_t)w + 2], blob, (size_t)numBytes);
    bytesLeft -= numBytes + 2;
    const int newBytes = secNumBytes(curSection) + numBytes + 2, newElems = secNumElems(curSection) + 1;
    setShort(&v[(size_t)curSection], (int16_t)newBytes); v[(size_t)curSection + 2] = (uint8_t)newElems;
    return Ack;
  }
  AddResponse appendBlob(Ptr blob, int numBytes) {                          // Section.scala:107-116
    if (needNewSection(numBytes)) {
      if (bytesLeft >= 4 + numBytes) { const int s = secEnd(curSection); secInit(s, 0); curSection = s; bytesLeft -= 4; }
      else return VectorTooSmall;
    }
    return addBlobInner(blob, numBytes);
  }
  AddResponse newSectionWithBlob(Ptr blob, int numBytes, int type) {        // Section.scala:119-129
    if (bytesLeft >= 4 + numBytes) { const int s = secEnd(curSection); secInit(s, type); curSection = s; bytesLeft -= 4; }
    else return VectorTooSmall;
    return addBlobInner(blob, numBytes);
  }
  AddResponse addData(Ptr buf, int cap) {                                   // HistogramVector.scala:364-409
    if (cap < 5 || !bin::isValidFormatCode(bin::formatCode(buf)) || bin::formatCode(buf) == FMT_NULL) return InvalidHistogram;
    if (bin::bucketDefNumBytes(buf) > bin::totalLength(buf)) return InvalidHistogram;
    const int numItems = length();
    const int defBytes = bin::bucketDefNumBytes(buf);
    if (exp) {                                                              // ExpHistogramVector.scala:73-97
      if (numItems == 0) { const int firstSect = OffsetNumHistograms + 2; secInit(firstSect, 0); curSection = firstSect; bytesLeft = (maxBytes - firstSect) - 4; }
      const AddResponse r = appendBlob(buf, bin::totalLength(buf));
      if (r == Ack) { setInt(&v[0], maxBytes - bytesLeft - 4); setShort(&v[OffsetNumHistograms], (int16_t)(numItems + 1)); }
      return r;
    }
    if (numItems == 0) {
      std::memcpy(&v[OffsetBucketDef], buf + 5, (size_t)defBytes);
      setShort(&v[OffsetBucketDefSize], (int16_t)defBytes);
      v[OffsetFormatCode] = bin::formatCode(buf);
      const int firstSect = OffsetBucketDef + defBytes;
      secInit(firstSect, 0); curSection = firstSect; bytesLeft = (maxBytes - firstSect) - 4;         // initSectionWriter
    } else {
      if (!(bin::formatCode(buf) == v[OffsetFormatCode] && defBytes == (getShort(&v[OffsetBucketDefSize]) & 0xffff) &&
            std::memcmp(&v[OffsetBucketDef], buf + 5, (size_t)defBytes) == 0)) return BucketSchemaMismatch;
    }
    const AddResponse res = sect ? appendSect(buf, numItems) : appendBlob(buf + bin::valuesIndex(buf), bin::valuesNumBytes(buf));
    if (res == Ack) { setInt(&v[0], maxBytes - bytesLeft - 4); setShort(&v[OffsetNumHistograms], (int16_t)(numItems + 1)); }
    return res;
  }
  AddResponse appendSect(Ptr buf, int numItems) {                           // AppendableSectDeltaHistVector.appendHist, :503-535
    const int numBuckets = bin::numBuckets(buf);
    if (!sinkInit) { sinkInit = true; originalDeltas.assign((size_t)numBuckets, 0); lastHistDeltas.assign((size_t)numBuckets, 0); }
    // DeltaSectDiffPackSink.process over all groups of the histogram
    std::vector<uint8_t> repacked(16); int writePos = 0; bool valueDropped = false; int i = 0;
    Ptr p = buf + bin::valuesIndex(buf); int cap = bin::valuesNumBytes(buf);
    int valuesLeft = numBuckets;
    while (valuesLeft > 0 && cap > 0) {
      uint64_t data[8];
      if (nibble::unpack8(p, cap, data) != nibble::Ok) throw CorruptVector("RepackError: input too short");
      const int numElems = std::min(numBuckets - i, 8);
      uint64_t packArray[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int n = 0; n < numElems; ++n) {
        if ((int64_t)data[n] < lastHistDeltas[(size_t)(i + n)]) valueDropped = true;
        packArray[n] = (uint64_t)lsub((int64_t)data[n], originalDeltas[(size_t)(i + n)]);
      }
      for (int n = 0; n < numElems; ++n) lastHistDeltas[(size_t)(i + n)] = (int64_t)data[n];
      writePos = nibble::pack8(packArray, repacked, writePos);
      i += 8; valuesLeft -= 8;
    }
    Ptr orig = buf + bin::valuesIndex(buf); const int origLen = bin::valuesNumBytes(buf);
    if (valueDropped) { originalDeltas = lastHistDeltas; return newSectionWithBlob(orig, origLen, 1 /* Section.TypeDrop */); }
    if (numItems == 0 || needNewSection(origLen)) { originalDeltas = lastHistDeltas; return appendBlob(orig, origLen); }
    return appendBlob(repacked.data(), writePos);
  }
  std::vector<uint8_t> bytes() const { return std::vector<uint8_t>(v.begin(), v.begin() + (getInt(&v[0]) + 4)); }
};

// ---------------------------------------------------------------------------------------------------------------------
// Readers: RowHistogramReader (HistogramVector.scala:557-616) and SectDeltaHistogramReader (:628-738)
// ---------------------------------------------------------------------------------------------------------------------
struct HistCorrection { bool some = false; LongHist lastValue, correction; };    // NoCorrection == !some; HistogramCorrection :618

struct HistReader {
  Ptr vec = nullptr; bool sect = false, exp = false; int len = 0, nb = 0; Buckets buckets;
  bool corrInit = false; std::vector<std::pair<int, LongHist>> corrections;
  explicit HistReader(Ptr v) : vec(v) {
    const int w = vectorType(v);
    if (w == WIRE_H_SECTDELTA) sect = true; else if (w == WIRE_H_EXP_SIMPLE) exp = true; else if (w != WIRE_H_SIMPLE) throw CorruptVector("not a histogram vector");
    len = getShort(v + OffsetNumHistograms) & 0xffff;
    if (exp) { buckets = Buckets::exponential(20, 0, 0); nb = 1; return; }   // RowExpHistogramReader.buckets = emptyExpBuckets, ExpHistogramVector.scala:134
    nb = len > 0 ? getShort(v + OffsetBucketDef) & 0xffff : 0;
    buckets = len > 0 ? Buckets::parse(v + OffsetBucketDefSize, (uint8_t)v[OffsetFormatCode]) : Buckets();
  }
  int length() const { return len; }
  Ptr endAddr() const { return vec + getInt(vec) + 4; }
  Ptr firstSection() const { return exp ? vec + OffsetNumHistograms + 2 : vec + OffsetBucketDef + (getShort(vec + OffsetBucketDefSize) & 0xffff); }
  // SectionReader.locate, Section.scala:176-203: section holding elemNo, its starting element, pointer to the record
  void locate(int elemNo, Ptr& sectOut, int& sectStart, Ptr& rec) const {
    if (elemNo < 0 || elemNo >= len) throw std::out_of_range("is out of vector bounds");
    Ptr s = firstSection(); int start = 0;
    while (elemNo >= start + s[2] && s + 4 + (getShort(s) & 0xffff) < endAddr()) { start += s[2]; s = s + 4 + (getShort(s) & 0xffff); }
    Ptr p = s + 4;
    for (int togo = elemNo - start; togo > 0; --togo) p += (getShort(p) & 0xffff) + 2;
    sectOut = s; sectStart = start; rec = p;
  }
  void unpackRecord(Ptr rec, std::vector<int64_t>& out) const {
    out.assign((size_t)nb, 0);
    if (nb > 0 && nibble::unpackDelta(rec + 2, getShort(rec) & 0xffff, out.data(), nb) != nibble::Ok) throw CorruptVector("hist record: input too short");
  }
  LongHist apply(int index) const {                                          // :601-609 / :646-666
    if (len <= 0) throw std::invalid_argument("EmptyHistogramException");
    Ptr s, rec; int start; locate(index, s, start, rec);
    if (exp) return bin::toHistogram(rec + 2);                               // ExpHistogramVector.scala:172-189: the record is a blob with its own scheme
    LongHist h; h.buckets = buckets;
    unpackRecord(rec, h.values);
    if (sect && index != start) {
      std::vector<int64_t> base; unpackRecord(s + 4, base);
      for (int b = 0; b < nb; ++b) h.values[(size_t)b] = ladd(base[(size_t)b], h.values[(size_t)b]);   // summedHist = base + delta
    }
    return h;
  }
  MutHist sum(int start, int end) const {                                    // :613-621
    if (!(len > 0 && start >= 0 && end < len)) throw std::invalid_argument("requirement failed");
    MutHist s = MutHist::emptyNaN(buckets);
    for (int i = start; i <= end; ++i) s.addNoCorrection(apply(i));
    return s;
  }
  // ---- CounterHistogramReader (SectDelta only)
  void forceCorrections() {                                                  // lazy val corrections, :690-707
    if (corrInit) return;
    corrInit = true;
    int index = 0;
    forEachSection([&](Ptr s) {
      if (index > 0 && s[3] == 1) corrections.emplace_back(index, apply(index - 1));     // Section.TypeDrop
      index += s[2];
    });
  }
  // SectionReader.iterateSections, Section.scala:217-225 (sections tile the vector up to endAddr)
  template <class F> void forEachSection(F&& f) const {
    if (len <= 0) return;
    for (Ptr s = firstSection(); s + 4 <= endAddr() && s + 4 + (getShort(s) & 0xffff) <= endAddr(); s = s + 4 + (getShort(s) & 0xffff)) f(s);
  }
  HistCorrection detectDropAndCorrection(HistCorrection meta) {              // :673-686
    if (!meta.some) return meta;
    const LongHist firstValue = apply(0);
    if (compareHist(firstValue, meta.lastValue) < 0) meta.correction.add(meta.lastValue);
    return meta;
  }
  HistCorrection updateCorrection(const HistCorrection& meta) {              // :717-728
    forceCorrections();
    HistCorrection out; out.some = true;
    out.correction = meta.some ? meta.correction : LongHist::empty(buckets);
    for (auto& c : corrections) out.correction.add(c.second);
    out.lastValue = apply(len - 1);
    return out;
  }
  LongHist correctedValue(int n, const HistCorrection& meta) {               // :730-746
    forceCorrections();
    LongHist h = apply(n);
    for (auto& c : corrections) if (c.first <= n) h.add(c.second);
    if (meta.some) h.add(meta.correction);
    return h;
  }
  std::vector<int> dropPositions() { forceCorrections(); std::vector<int> r; for (auto& c : corrections) r.push_back(c.first); return r; }
  std::vector<int> sectionTypes() const { std::vector<int> r; forEachSection([&](Ptr s) { r.push_back(s[3]); }); return r; }
};

// ---------------------------------------------------------------------------------------------------------------------
// Range functions over a histogram column, one window: HistogramRateFunctionBase / HistRateFunction / HistIncreaseFunction
// (RateFunctions.scala:330-418), SumOverTimeChunkedFunctionH (AggrOverTimeFunctions.scala:587-606),
// RateOverDeltaChunkedFunctionH (RateFunctions.scala:470-494).  Driven by the same WindowedChunkIterator as the double path.
// ---------------------------------------------------------------------------------------------------------------------
struct HistSeries {
  std::vector<Ptr> infos;           // ChunkSetInfo addresses, column 0 = timestamps, column `hcol` = histogram vector
  int hcol = 1;
};

// periodic samples of one series: out[k] = histogram of window k (isEmpty() when the reference emits Histogram.empty)
inline void periodicSamplesHist(const HistSeries& S, int fn, bool cumulative, int64_t start, int64_t step, int64_t end, int64_t window,
                                bool inclusiveRange, std::vector<MutHist>& out) {
  const int64_t adjStep = step > 0 ? step : step + 1;
  const int T = (int)((end - start) / adjStep) + 1;
  out.assign((size_t)T, MutHist());
  // resolve readers once per chunk (ChunkSetInfo.scala:495-502)
  struct Ch { Ptr info; LongReader ts{LongReader::RAW64}; Ptr tsVec; HistReader hr; Ch(Ptr i, Ptr tv, Ptr hv) : info(i), tsVec(tv), hr(hv) {} };
  std::vector<Ch> chunks;
  for (Ptr info : S.infos) {
    if (csi::numRows(info) <= 0) continue;
    Ptr tv = csi::vectorPtr(info, 0), hv = csi::vectorPtr(info, S.hcol);
    chunks.emplace_back(info, tv, hv);
    chunks.back().ts = LongReader::of(tv);
  }
  const int64_t winDur = inclusiveRange ? window : window - 1;
  for (int k = 0; k < T; ++k) {
    const int64_t wEnd = start + (int64_t)k * adjStep, wStart = wEnd - (winDur < 0 ? 0 : winDur);
    // chunk set of the window, closed form for time-ordered chunks (ChunkSetInfo.scala:481-510; see filo_query.hpp)
    const bool counterFn = cumulative && (fn == FN_RATE || fn == FN_INCREASE);
    int32_t numSamples = 0; int64_t lowestTime = INT64_MAX, highestTime = 0; LongHist lowestValue, highestValue;
    HistCorrection meta;
    MutHist h;                                           // SumOverTimeChunkedFunctionH.h (Histogram.empty)
    for (size_t c = 0; c < chunks.size(); ++c) {
      Ch& ch = chunks[c];
      if (csi::endTime(ch.info) < wStart) continue;
      if (c > 0 && !(csi::endTime(chunks[c - 1].info) < wEnd)) continue;
      const int32_t startRowNum = ch.ts.binarySearch(ch.tsVec, wStart) & 0x7fffffff;
      const int32_t endRowNum = std::min(ch.ts.ceilingIndex(ch.tsVec, wEnd), csi::numRows(ch.info) - 1);
      if (counterFn) {                                   // CounterChunkedRangeFunction.addChunks, RangeFunction.scala:138-163
        meta = ch.hr.detectDropAndCorrection(meta);
        if (startRowNum <= endRowNum) {
          const int64_t tS = ch.ts.apply(ch.tsVec, startRowNum), tE = ch.ts.apply(ch.tsVec, endRowNum);
          if (tS < lowestTime || tE > highestTime) {     // HistogramRateFunctionBase.addTimeChunks, RateFunctions.scala:349-364
            numSamples += endRowNum - startRowNum + 1;
            if (tS < lowestTime) { lowestTime = tS; lowestValue = ch.hr.correctedValue(startRowNum, meta); }
            if (tE > highestTime) { highestTime = tE; highestValue = ch.hr.correctedValue(endRowNum, meta); }
          }
        }
        meta = ch.hr.updateCorrection(meta);
      } else {                                           // TimeRangeFunction.addChunks + SumOverTimeChunkedFunctionH.addTimeChunks
        if (!(startRowNum <= endRowNum)) continue;
        MutHist sum = ch.hr.sum(startRowNum, endRowNum);
        if (h.numBuckets() == 0) h = sum; else h.add(sum);
      }
    }
    MutHist& o = out[(size_t)k];
    if (counterFn) {                                     // HistogramRateFunctionBase.apply, RateFunctions.scala:366-407
      if (highestTime > lowestTime && highestValue.buckets == lowestValue.buckets) {
        const int64_t cws = inclusiveRange ? wStart : wStart - 1;
        o.buckets = lowestValue.buckets; o.values.resize((size_t)lowestValue.numBuckets());
        for (int b = 0; b < lowestValue.numBuckets(); ++b)
          o.values[(size_t)b] = extrapolatedRate(cws, wEnd, numSamples, lowestTime, lowestValue.bucketValue(b), highestTime, highestValue.bucketValue(b),
                                                 true, fn == FN_RATE);
      }
    } else if (fn == FN_RATE) {                          // RateOverDeltaChunkedFunctionH.apply, RateFunctions.scala:476-485 (raw windowStart)
      o.buckets = h.buckets; o.values.resize((size_t)h.numBuckets());
      for (int b = 0; b < h.numBuckets(); ++b) o.values[(size_t)b] = h.bucketValue(b) / (double)(wEnd - wStart) * 1000;
    } else {
      o = h;                                             // sum_over_time / increase on a delta schema
    }
  }
}

// HistSumRowAggregator.reduceAggregate, aggregator/HistSumRowAggregator.scala:25-36: fold one series' window into the group sum
inline void histSumReduce(MutHist& acc, const MutHist& newHist) {
  if (acc.numBuckets() == 0) acc = newHist;
  else if (newHist.numBuckets() > 0) acc.add(newHist);
}

} // namespace hist
} // namespace fo
