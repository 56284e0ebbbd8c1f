// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of FiloDB's BinaryVector wire format: readers (decode side) and the
// appenders' optimize() (encode side, used only to generate byte-identical test chunks).
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Operation order follows the Scala source; compile with -O2 -ffp-contract=off (the JVM never
// fuses a*b+c).  Parity status: pinned against the reference's own golden vectors in
// tests/test_oracle_golden.py (NibblePackTest, DoubleVectorTest, LongVectorTest,
// IntBinaryVectorTest, WindowIteratorSpec, RateFunctionsSpec, AggrOverTimeFunctionsSpec ...).
// The XOR-double *container header* (FiloXorDoubleVector) has no reference counterpart:
// "parity unpinned" for that header, pinned for its NibblePack payload.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <limits>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <memory>

namespace fo {

using Ptr = const uint8_t*;

// ---- little-endian unaligned loads/stores (UnsafeUtils / MemoryReader, MemoryReader.scala:39-68)
inline int8_t   getByte(Ptr p)   { return (int8_t)p[0]; }
inline int16_t  getShort(Ptr p)  { int16_t v; std::memcpy(&v, p, 2); return v; }
inline int32_t  getInt(Ptr p)    { int32_t v; std::memcpy(&v, p, 4); return v; }
inline int64_t  getLong(Ptr p)   { int64_t v; std::memcpy(&v, p, 8); return v; }
inline double   getDouble(Ptr p) { double v;  std::memcpy(&v, p, 8); return v; }
inline void setByte(uint8_t* p, int8_t v)    { p[0] = (uint8_t)v; }
inline void setShort(uint8_t* p, int16_t v)  { std::memcpy(p, &v, 2); }
inline void setInt(uint8_t* p, int32_t v)    { std::memcpy(p, &v, 4); }
inline void setLong(uint8_t* p, int64_t v)   { std::memcpy(p, &v, 8); }
inline void setDouble(uint8_t* p, double v)  { std::memcpy(p, &v, 8); }

// Scala Int arithmetic wraps; C++ signed overflow is UB, so go through uint32.
inline int32_t imul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
inline int32_t iadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t isub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int64_t ladd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
inline int64_t lsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
inline int64_t lmul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }

struct CorruptVector : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- WireFormat.scala:7-53
namespace wire {
constexpr int VECTORTYPE_BINSIMPLE = 0x06;
constexpr int VECTORTYPE_DELTA2    = 0x08;
constexpr int VECTORTYPE_HISTOGRAM = 0x09;
constexpr int SUBTYPE_PRIMITIVE        = 0x00;
constexpr int SUBTYPE_PRIMITIVE_NOMASK = 0x05;
constexpr int SUBTYPE_REPEATED         = 0x06;
constexpr int SUBTYPE_INT              = 0x07;
constexpr int SUBTYPE_INT_NOMASK       = 0x08;
constexpr int SUBTYPE_H_SIMPLE         = 0x10;
constexpr int SUBTYPE_H_SECTDELTA      = 0x12;
// NOT in the reference: our XOR-NibblePack double container (SURVEY "Read this first" #1).
constexpr int SUBTYPE_XOR_NIBBLE       = 0x21;
constexpr int make(int major, int sub) { return ((sub & 0xff) << 8) | (major & 0xff); }
}

// BinaryVector.scala:31-42
inline int vectorType(Ptr v)      { return getShort(v + 4) & 0xffff; }
inline int majorVectorType(Ptr v) { return getShort(v + 4) & 0x00ff; }
inline int totalBytes(Ptr v)      { return getInt(v) + 4; }
inline int numBytes(Ptr v)        { return getInt(v); }

// PrimitiveVectorReader, BinaryVector.scala:511-536
inline int  pv_nbits(Ptr v)    { return getByte(v + 6) & 0x7f; }
inline int  pv_bitShift(Ptr v) { return getByte(v + 7) & 0x3f; }
inline bool pv_signed(Ptr v)   { return (getByte(v + 6) & 0x80) != 0; }
inline bool pv_dropped(Ptr v)  { return (getShort(v + 6) & 0x8000) != 0; }
inline void pv_markDrop(uint8_t* v) { setShort(v + 6, (int16_t)(getShort(v + 6) | 0x8000)); }

// =====================================================================================
// IntBinaryVector readers  (IntBinaryVector.scala:120-137 dispatch, :306-457 readers)
// =====================================================================================
struct IntReader {
  enum Kind { S32, S16, S8, U16, U8, U4, U2 } kind;
  // IntBinaryVector.simple, :120-137
  static IntReader simple(Ptr v) {
    int nb = pv_nbits(v);
    if (pv_signed(v)) {
      switch (nb) { case 32: return {S32}; case 16: return {S16}; case 8: return {S8}; }
    } else {
      switch (nb) { case 32: return {S32}; case 16: return {U16}; case 8: return {U8};
                    case 4: return {U4}; case 2: return {U2}; }
    }
    throw CorruptVector("IntBinaryVector.simple: MatchError nbits=" + std::to_string(nb));
  }
  // IntVectorDataReader.length, :248-250
  int length(Ptr v) const {
    int bs = pv_bitShift(v);
    return ((numBytes(v) - 4) * 8 + (bs != 0 ? bs - 8 : 0)) / pv_nbits(v);
  }
  int apply(Ptr v, int n) const {
    switch (kind) {
      case S32: return getInt(v + 8 + (int64_t)n * 4);                         // :307-308
      case S16: return (int)getShort(v + 8 + (int64_t)n * 2);                  // :324-325
      case S8:  return (int)getByte(v + 8 + n);                                // :341-342
      case U16: return getShort(v + 8 + (int64_t)n * 2) & 0xffff;              // :358-359
      case U8:  return getByte(v + 8 + n) & 0xff;                              // :382-383
      case U4:  return ((int)getByte(v + 8 + n / 2) >> ((n & 1) * 4)) & 0x0f;  // :408-409
      case U2:  return ((int)getByte(v + 8 + n / 4) >> ((n & 3) * 2)) & 0x03;  // :434-435
    }
    return 0;
  }
  // All sum() variants are exact Long sums (:310-456), so any evaluation order is identical.
  int64_t sum(Ptr v, int start, int end) const {
    if (!(start >= 0 && end < length(v))) throw std::invalid_argument("int sum out of bounds");
    int64_t s = 0;
    for (int r = start; r <= end; ++r) s += apply(v, r);
    return s;
  }
};

// =====================================================================================
// Long (timestamp) readers  (LongBinaryVector.scala:60-67 dispatch)
// =====================================================================================
inline double slopeSum(int64_t initVal, int32_t slope, int start, int end) {   // DeltaDeltaVector.scala:265-268
  int32_t len = iadd(isub(end, start), 1);
  int64_t a = ladd(initVal, lmul((int64_t)start, (int64_t)slope));            // start * slope.toLong
  int32_t half = imul(isub(end, start), len) / 2;                              // Int arithmetic
  return (double)len * (double)a + (double)lmul((int64_t)half, (int64_t)slope);
}

struct LongReader {
  enum Kind { DDV, DDV_CONST, MASKED, RAW64 } kind;
  static LongReader of(Ptr v) {
    int t = vectorType(v);
    if (t == wire::make(wire::VECTORTYPE_DELTA2, wire::SUBTYPE_INT_NOMASK)) return {DDV};
    if (t == wire::make(wire::VECTORTYPE_DELTA2, wire::SUBTYPE_REPEATED)) return {DDV_CONST};
    if (t == wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_PRIMITIVE)) return {MASKED};
    if (t == wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_PRIMITIVE_NOMASK)) return {RAW64};
    throw CorruptVector("LongBinaryVector: MatchError wire=" + std::to_string(t));
  }
  static Ptr subvect(Ptr v) { return v + getInt(v + 8); }                      // BinaryVector.scala:197-198

  int length(Ptr v) const {
    switch (kind) {
      case DDV: { Ptr in = v + 20; return IntReader::simple(in).length(in); } // DeltaDeltaVector.scala:149-150
      case DDV_CONST: return getInt(v + 8);                                     // :238
      case MASKED: { Ptr s = subvect(v); return of(s).length(s); }              // LongBinaryVector.scala:278-279
      case RAW64: return (numBytes(v) - 4) / 8;                                 // :116-117
    }
    return 0;
  }
  int64_t apply(Ptr v, int n) const {
    switch (kind) {
      case DDV: {                                                               // DeltaDeltaVector.scala:153-156
        Ptr in = v + 20;
        return ladd(ladd(getLong(v + 8), lmul((int64_t)getInt(v + 16), (int64_t)n)),
                    (int64_t)IntReader::simple(in).apply(in, n));
      }
      case DDV_CONST:                                                           // :241-242  (Int * Int, then widened)
        return ladd(getLong(v + 12), (int64_t)imul(getInt(v + 20), n));
      case MASKED: { Ptr s = subvect(v); return of(s).apply(s, n); }
      case RAW64: return getLong(v + 8 + (int64_t)n * 8);                       // LongBinaryVector.scala:205-206
    }
    return 0;
  }
  // bits 0-30: row; bit 31: inexact.  LongBinaryVector.scala:145-152
  int32_t binarySearch(Ptr v, int64_t item) const {
    switch (kind) {
      case DDV: {                                                               // DeltaDeltaVector.scala:159-188
        int64_t slope = (int64_t)getInt(v + 16);
        int64_t init = getLong(v + 8);
        int32_t len = length(v);
        int32_t elemNo;
        if (slope == 0) elemNo = (item <= init) ? 0 : len;
        else elemNo = (int32_t)(ladd(lsub(item, init), slope - 1) / slope);
        if (elemNo < 0) elemNo = 0;
        if (elemNo >= len) elemNo = len - 1;
        int64_t curBase = ladd(init, lmul(slope, (int64_t)elemNo));
        Ptr in = v + 20;
        IntReader ir = IntReader::simple(in);
        while (elemNo >= 0 && item < ladd(curBase, (int64_t)ir.apply(in, elemNo))) {
          elemNo -= 1; curBase = lsub(curBase, slope);
        }
        if (elemNo >= 0 && item == ladd(curBase, (int64_t)ir.apply(in, elemNo))) return elemNo;
        elemNo += 1; curBase = ladd(curBase, slope);
        while (elemNo < len && item > ladd(curBase, (int64_t)ir.apply(in, elemNo))) {
          elemNo += 1; curBase = ladd(curBase, slope);
        }
        if (elemNo < len && item == ladd(curBase, (int64_t)ir.apply(in, elemNo))) return elemNo;
        return (int32_t)((uint32_t)elemNo | 0x80000000u);
      }
      case DDV_CONST: {                                                         // DeltaDeltaVector.scala:245-253
        int64_t slope = (int64_t)getInt(v + 20);
        int64_t init = getLong(v + 12);
        int32_t len = length(v);
        int32_t guess;
        if (slope == 0) guess = (item <= init) ? 0 : len;
        else guess = (int32_t)(ladd(lsub(item, init), slope - 1) / slope);
        if (guess < 0) return (int32_t)0x80000000u;
        if (guess >= len) return (int32_t)(0x80000000u | (uint32_t)len);
        if (item != apply(v, guess)) return (int32_t)(0x80000000u | (uint32_t)guess);
        return guess;
      }
      case MASKED: { Ptr s = subvect(v); return of(s).binarySearch(s, item); }
      case RAW64: {                                                             // LongBinaryVector.scala:227-246
        int32_t len = length(v);
        if (len == 0) return (int32_t)0x80000000u;
        int32_t first = 0; int64_t element = 0;
        while (len > 0) {
          int32_t half = (int32_t)((uint32_t)len >> 1);
          int32_t middle = first + half;
          element = getLong(v + 8 + (int64_t)middle * 8);
          if (element == item) return middle;
          else if (element < item) { first = middle + 1; len = len - half - 1; }
          else len = half;
        }
        return element == item ? first : (int32_t)((uint32_t)first | 0x80000000u);
      }
    }
    return 0;
  }
  // last row with ts <= item, -1 if none.  LongBinaryVector.scala:162-169
  int32_t ceilingIndex(Ptr v, int64_t item) const {
    int32_t row = binarySearch(v, item);
    if (row < 0) return (row & 0x7fffffff) - 1;
    return row;
  }
  double sum(Ptr v, int start, int end) const {
    switch (kind) {
      case DDV: {                                                               // DeltaDeltaVector.scala:190-194
        Ptr in = v + 20;
        return slopeSum(getLong(v + 8), getInt(v + 16), start, end)
             + (double)IntReader::simple(in).sum(in, start, end);
      }
      case DDV_CONST:                                                           // :259-263
        if (!(start >= 0 && end < length(v))) throw std::invalid_argument("ddv const sum out of bounds");
        return slopeSum(getLong(v + 12), getInt(v + 20), start, end);
      case MASKED: { Ptr s = subvect(v); return of(s).sum(s, start, end); }
      case RAW64: {                                                             // LongBinaryVector.scala:211-222
        if (!(start >= 0 && end < length(v))) throw std::invalid_argument("long sum out of bounds");
        double s = 0;
        for (int r = start; r <= end; ++r) s += (double)getLong(v + 8 + (int64_t)r * 8);
        return s;
      }
    }
    return 0;
  }
  // changes(): number of value changes in [start, end] and the last value seen.  Every reader has its own treatment of `prev`
  // (the last value of the previous chunk) -- kept as they are in the reference.
  std::pair<int64_t, int64_t> changes(Ptr v, int start, int end, int64_t prev, bool ignorePrev = false) const {
    if (!(start >= 0 && end < length(v))) throw std::invalid_argument("long changes out of bounds");
    switch (kind) {
      case DDV: {                                                               // DeltaDeltaVector.scala:212-227
        int64_t prevVector = prev, ch = 0;
        for (int i = start; i <= end; ++i) {
          int64_t cur = apply(v, i);
          if (i == start && ignorePrev) prevVector = cur;
          if (prevVector != cur) ch += 1;
          prevVector = cur;
        }
        return {ch, prevVector};
      }
      case DDV_CONST: {                                                         // :280-288
        int64_t firstValue = apply(v, start), lastValue = apply(v, end);
        int64_t ch = (!ignorePrev && prev != firstValue) ? 1 : 0;
        if (getInt(v + 20) == 0) return {ch, lastValue};
        return {(int64_t)(end - start) + ch, lastValue};
      }
      case MASKED: { Ptr s = subvect(v); return of(s).changes(s, start, end, prev); }   // LongBinaryVector.scala:290-292 (ignorePrev not forwarded)
      case RAW64: {                                                             // :248-265: the first element always re-seeds prev
        int64_t prevVector = prev, ch = 0;
        for (int i = start; i <= end; ++i) {
          int64_t cur = getLong(v + 8 + (int64_t)i * 8);
          if (i == start) prevVector = cur;
          if (prevVector != cur) ch += 1;
          prevVector = cur;
        }
        return {ch, prevVector};
      }
    }
    return {0, prev};
  }
};

// JVM d2l: NaN -> 0, saturating
inline int64_t d2l(double d) {
  if (d != d) return 0;
  if (d >= 9223372036854775807.0) return INT64_MAX;
  if (d <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)d;
}

// =====================================================================================
// NibblePack  (NibblePack.scala)
// =====================================================================================
namespace nibble {
inline int nlz(uint64_t x) { return x == 0 ? 64 : __builtin_clzll(x); }
inline int ntz(uint64_t x) { return x == 0 ? 64 : __builtin_ctzll(x); }

// packUniversal, NibblePack.scala:147-183.  out must have >= 8*8+8 spare bytes.
inline int packUniversal(const uint64_t in[8], std::vector<uint8_t>& buf, int bufindex, int numNibbles, int trailingZeroNibbles) {
  int bufpos = bufindex;
  int trailingShift = trailingZeroNibbles * 4;
  int numBits = numNibbles * 4;
  uint64_t outWord = 0; int bitCursor = 0;
  auto putLong = [&](int pos, uint64_t w) {
    if ((int)buf.size() < pos + 8) buf.resize(pos + 8);
    std::memcpy(buf.data() + pos, &w, 8);
  };
  for (int i = 0; i < 8; ++i) {
    uint64_t input = in[i];
    if (input != 0) {
      int remaining = 64 - bitCursor;
      uint64_t shiftedInput = input >> trailingShift;
      outWord |= shiftedInput << bitCursor;
      if (remaining <= numBits) {
        putLong(bufpos, outWord); bufpos += 8;
        if (remaining < numBits) outWord = shiftedInput >> remaining;
        else outWord = 0;
      }
      bitCursor = (bitCursor + numBits) % 64;
    }
  }
  if (bitCursor > 0) { putLong(bufpos, outWord); bufpos += (bitCursor + 7) / 8; }
  return bufpos;
}

// pack8, NibblePack.scala:108-144.  Returns ending position; buf grows as needed
// (trailing bytes past the returned position are scratch, like ExpandableArrayBuffer).
inline int pack8(const uint64_t in[8], std::vector<uint8_t>& buf, int bufindex) {
  int bufpos = bufindex;
  if ((int)buf.size() < bufpos + 2) buf.resize(bufpos + 2);
  int bitmask = 0;
  for (int i = 0; i < 8; ++i) if (in[i] != 0) bitmask |= 1 << i;
  buf[bufpos++] = (uint8_t)bitmask;
  if (bitmask != 0) {
    int minLeadingZeros = 64, minTrailingZeros = 64;
    for (int i = 0; i < 8; ++i) {
      minLeadingZeros = std::min(minLeadingZeros, nlz(in[i]));
      minTrailingZeros = std::min(minTrailingZeros, ntz(in[i]));
    }
    int trailingNibbles = minTrailingZeros / 4;
    int numNibbles = 16 - (minLeadingZeros / 4) - trailingNibbles;
    int nibbleWord = ((numNibbles - 1) << 4) | trailingNibbles;
    buf[bufpos++] = (uint8_t)nibbleWord;
    bufpos = packUniversal(in, buf, bufpos, numNibbles, trailingNibbles);
  }
  return bufpos;
}

// packDoubles, NibblePack.scala:73-98
inline int packDoubles(const double* inputs, int n, std::vector<uint8_t>& buf, int bufindex) {
  if (n <= 0) throw std::invalid_argument("packDoubles: empty");
  if ((int)buf.size() < bufindex + 8) buf.resize(bufindex + 8);
  std::memcpy(buf.data() + bufindex, &inputs[0], 8);
  int pos = bufindex + 8;
  uint64_t arr[8];
  uint64_t last; std::memcpy(&last, &inputs[0], 8);
  int i = 0;
  while (i < n - 1) {
    uint64_t bits; std::memcpy(&bits, &inputs[i + 1], 8);
    arr[i % 8] = bits ^ last;
    last = bits;
    i += 1;
    if (i % 8 == 0) pos = pack8(arr, buf, pos);
  }
  if (i % 8 != 0) {
    for (int j = i % 8; j < 8; ++j) arr[j] = 0;
    pos = pack8(arr, buf, pos);
  }
  return pos;
}

// packDelta, NibblePack.scala:37-54 (+ packRemainder :56-63)
inline int packDelta(const int64_t* input, int n, std::vector<uint8_t>& buf, int bufindex) {
  uint64_t arr[8]; int64_t last = 0; int i = 0; int pos = bufindex;
  while (i < n) {
    int64_t delta = (input[i] >= last) ? lsub(input[i], last) : 0;
    last = input[i];
    arr[i % 8] = (uint64_t)delta;
    i += 1;
    if (i % 8 == 0) pos = pack8(arr, buf, pos);
  }
  if (i % 8 != 0) { for (int j = i % 8; j < 8; ++j) arr[j] = 0; pos = pack8(arr, buf, pos); }
  return pos;
}

// packNonIncreasing, NibblePack.scala:16-30
inline int packNonIncreasing(const int64_t* input, int n, std::vector<uint8_t>& buf, int bufindex) {
  uint64_t arr[8]; int i = 0; int pos = bufindex;
  while (i < n) {
    arr[i % 8] = (uint64_t)input[i];
    i += 1;
    if (i % 8 == 0) pos = pack8(arr, buf, pos);
  }
  if (i % 8 != 0) { for (int j = i % 8; j < 8; ++j) arr[j] = 0; pos = pack8(arr, buf, pos); }
  return pos;
}

// readLong with bounds, NibblePack.scala:457-469
inline uint64_t readLong(Ptr buf, int cap, int index) {
  if (index + 8 <= cap) { uint64_t w; std::memcpy(&w, buf + index, 8); return w; }
  uint64_t out = 0; int i = 0;
  while (index + i < cap) { out |= ((uint64_t)buf[index + i]) << (8 * i); i++; }
  return out;
}

enum UnpackResult { Ok = 0, InputTooShort = 1 };

// unpack8, NibblePack.scala:395-447.  Consumes from (buf, cap) and advances them (subslice :449-454).
inline UnpackResult unpack8(Ptr& buf, int& cap, uint64_t out[8]) {
  auto subslice = [&](int start) { if (cap > start) { buf += start; cap -= start; } else { cap = 0; } };
  uint8_t nonzeroMask = buf[0];
  if (nonzeroMask == 0) {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    subslice(1);
    return Ok;
  }
  int numNibblesU8 = buf[1] & 0xff;
  int numBits = ((numNibblesU8 >> 4) + 1) * 4;
  int trailingZeroes = (numNibblesU8 & 0x0f) * 4;
  int total = 2 + (numBits * __builtin_popcount(nonzeroMask) + 7) / 8;
  uint64_t mask = numBits >= 64 ? ~0ull : ((1ull << numBits) - 1);
  int bufIndex = 2; int bitCursor = 0;
  uint64_t inWord = readLong(buf, cap, bufIndex); bufIndex += 8;
  for (int bit = 0; bit < 8; ++bit) {
    if (nonzeroMask & (1 << bit)) {
      int remaining = 64 - bitCursor;
      uint64_t shiftedIn = inWord >> bitCursor;
      uint64_t outWord = shiftedIn & mask;
      if (remaining <= numBits && bufIndex < total) {
        if (bufIndex < cap) {
          inWord = readLong(buf, cap, bufIndex); bufIndex += 8;
          if (remaining < numBits) outWord |= (inWord << remaining) & mask;
        } else return InputTooShort;
      }
      out[bit] = outWord << trailingZeroes;
      bitCursor = (bitCursor + numBits) % 64;
    } else out[bit] = 0;
  }
  subslice(total);
  return Ok;
}

// unpackDoubleXOR + DoubleXORSink, NibblePack.scala:374-384, 232-244
inline UnpackResult unpackDoubleXOR(Ptr buf, int cap, double* outArray, int outLen) {
  if (cap < 8) return InputTooShort;
  uint64_t lastBits = readLong(buf, cap, 0);
  std::memcpy(&outArray[0], &lastBits, 8);
  if (cap > 8) { buf += 8; cap -= 8; } else cap = 0;
  int pos = 1; int valuesLeft = outLen - 1; UnpackResult res = Ok;
  uint64_t data[8];
  while (valuesLeft > 0 && res == Ok && cap > 0) {                      // unpackToSink :359-368
    res = unpack8(buf, cap, data);
    if (res == Ok) {
      int numElems = std::min(outLen - pos, 8);
      for (int n = 0; n < numElems; ++n) {
        uint64_t nextBits = lastBits ^ data[n];
        std::memcpy(&outArray[pos + n], &nextBits, 8);
        lastBits = nextBits;
      }
      pos += 8;
    }
    valuesLeft -= 8;
  }
  return res;
}

// unpackToSink with DeltaSink, NibblePack.scala:208-230, 359-368
inline UnpackResult unpackDelta(Ptr buf, int cap, int64_t* outArray, int outLen) {
  int i = 0; int64_t current = 0; int valuesLeft = outLen; UnpackResult res = Ok; uint64_t data[8];
  while (valuesLeft > 0 && res == Ok && cap > 0) {
    res = unpack8(buf, cap, data);
    if (res == Ok) {
      int numElems = std::min(outLen - i, 8);
      for (int n = 0; n < numElems; ++n) { current = ladd(current, (int64_t)data[n]); outArray[i + n] = current; }
      i += 8;
    }
    valuesLeft -= 8;
  }
  return res;
}
} // namespace nibble

// =====================================================================================
// XOR-NibblePack double container (NOT in the reference; payload == NibblePack.packDoubles)
//   +0  i32 numBytes (after this word)         +4 u16 wire = (0x21<<8)|0x06
//   +6  u16 flags (bit 15 = drop flag, same bit PrimitiveVectorReader.dropped tests)
//   +8  i32 numValues n                        +12 u16 numGroups = ceil((n-1)/8)   +14 u16 payloadOff
//   +16 u16 groupOff[numGroups]  byte offset of group g from (payload + 8)
//   +payloadOff  payload = f64 first, then NibblePack groups     (payloadOff is 8-byte aligned)
//   total padded to a multiple of 8 bytes with zeros
// =====================================================================================
namespace xorvec {
inline int numValues(Ptr v) { return getInt(v + 8); }
inline int numGroups(Ptr v) { return getShort(v + 12) & 0xffff; }
inline int payloadOff(Ptr v) { return getShort(v + 14) & 0xffff; }
inline void decode(Ptr v, std::vector<double>& out) {
  int n = numValues(v);
  out.assign(n, 0.0);
  if (n == 0) return;
  int po = payloadOff(v);
  int cap = totalBytes(v) - po;
  if (nibble::unpackDoubleXOR(v + po, cap, out.data(), n) != nibble::Ok) throw CorruptVector("xor: input too short");
}
}

// =====================================================================================
// Double readers  (DoubleVector.scala:62-71 dispatch)
// =====================================================================================
struct DoubleCorrection { bool some = false; double lastValue = 0; double correction = 0; };  // NoCorrection == !some

struct DoubleReader {
  enum Kind { LONGWRAP, MASKED, RAW64, XOR } kind = RAW64;
  bool correcting = false;          // CorrectingDoubleVectorReader wrapper (DoubleVector.scala:69, 308-392)
  Ptr vect = nullptr;               // bound vector (the reference's Correcting reader is bound to one chunk)
  // lazy state of CorrectingDoubleVectorReader
  bool correctedInit = false;
  std::vector<double> corrected;
  std::vector<int> drops;
  double _correction = 0.0;
  // decoded cache for XOR container
  mutable std::vector<double> xorDecoded; mutable bool xorInit = false;

  static DoubleReader of(Ptr v) {
    DoubleReader r; r.vect = v;
    int t = vectorType(v);
    if (t == wire::make(wire::VECTORTYPE_DELTA2, wire::SUBTYPE_INT_NOMASK)) r.kind = LONGWRAP;
    else if (t == wire::make(wire::VECTORTYPE_DELTA2, wire::SUBTYPE_REPEATED)) r.kind = LONGWRAP;
    else if (t == wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_PRIMITIVE)) r.kind = MASKED;
    else if (t == wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_PRIMITIVE_NOMASK)) r.kind = RAW64;
    else if (t == wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_XOR_NIBBLE)) r.kind = XOR;
    else throw CorruptVector("DoubleVector: MatchError wire=" + std::to_string(t));
    r.correcting = pv_dropped(v);
    return r;
  }
  const std::vector<double>& xorVals() const {
    if (!xorInit) { xorvec::decode(vect, xorDecoded); xorInit = true; }
    return xorDecoded;
  }
  // ---- inner reader methods
  int length() const {
    switch (kind) {
      case LONGWRAP: return LongReader::of(vect).length(vect);                  // DoubleVector.scala:549-550
      case MASKED: { Ptr s = vect + getInt(vect + 8); return (numBytes(s) - 4) / 8; }   // :404 (super.length on subvector)
      case RAW64: return (numBytes(vect) - 4) / 8;                              // :124-125
      case XOR: return xorvec::numValues(vect);
    }
    return 0;
  }
  double apply(int n) const {
    switch (kind) {
      case LONGWRAP: return (double)LongReader::of(vect).apply(vect, n);        // :551-552
      case MASKED: { Ptr s = vect + getInt(vect + 8); return DoubleReader::of(s).apply(n); }  // :399-402
      case RAW64: return getDouble(vect + 8 + (int64_t)n * 8);                  // :228-229
      case XOR: return xorVals()[n];
    }
    return 0;
  }
  // DoubleVectorDataReader64.sum, DoubleVector.scala:234-262 (ignoreNaN = true)
  static double sum64(const double* data, int start, int end) {
    double sum = std::numeric_limits<double>::quiet_NaN();
    for (int r = start; r <= end; ++r) {
      double nextDbl = data[r];
      if (!std::isnan(nextDbl)) { if (std::isnan(sum)) sum = 0; sum += nextDbl; }
    }
    return sum;
  }
  double sum(int start, int end) const {
    switch (kind) {
      case LONGWRAP: return LongReader::of(vect).sum(vect, start, end);         // :553-554
      case MASKED: { Ptr s = vect + getInt(vect + 8); return DoubleReader::of(s).sum(start, end); }
      case RAW64: {
        if (!(start >= 0 && end < length())) throw std::invalid_argument("double sum out of bounds");
        double sum = std::numeric_limits<double>::quiet_NaN();
        for (int r = start; r <= end; ++r) {
          double nextDbl = getDouble(vect + 8 + (int64_t)r * 8);
          if (!std::isnan(nextDbl)) { if (std::isnan(sum)) sum = 0; sum += nextDbl; }
        }
        return sum;
      }
      case XOR: {
        if (!(start >= 0 && end < length())) throw std::invalid_argument("double sum out of bounds");
        return sum64(xorVals().data(), start, end);
      }
    }
    return 0;
  }
  // DoubleVectorDataReader64.count, :264-281
  int count(int start, int end) const {
    switch (kind) {
      case LONGWRAP: return end - start + 1;                                     // :555
      case MASKED: { Ptr s = vect + getInt(vect + 8); return DoubleReader::of(s).count(start, end); }
      case RAW64: case XOR: {
        if (!(start >= 0 && end < length())) throw std::invalid_argument("double count out of bounds");
        int c = 0;
        for (int r = start; r <= end; ++r) if (!std::isnan(apply(r))) c++;
        return c;
      }
    }
    return 0;
  }
  // DoubleVectorDataReader64.changes :283-303; masked :414-416; DoubleLongWrapDataReader.changes :559-566
  std::pair<double, double> changes(int start, int end, double prev) const {
    switch (kind) {
      case LONGWRAP: {
        bool ignorePrev = std::isnan(prev);
        auto c = LongReader::of(vect).changes(vect, start, end, d2l(prev), ignorePrev);
        return {(double)c.first, (double)c.second};
      }
      case MASKED: { Ptr s = vect + getInt(vect + 8); return DoubleReader::of(s).changes(start, end, prev); }
      case RAW64: case XOR: {
        if (!(start >= 0 && end < length())) throw std::invalid_argument("double changes out of bounds");
        double ch = 0, prevVector = prev;
        for (int r = start; r <= end; ++r) {
          double nextDbl = apply(r);
          if (!std::isnan(nextDbl) && prevVector != nextDbl && !std::isnan(prevVector)) ch += 1;
          prevVector = nextDbl;
        }
        return {ch, prevVector};
      }
    }
    return {0, prev};
  }
  // ---- counter-correction API
  // detectDropAndCorrection, DoubleVector.scala:177-187 (same for Correcting wrapper: not overridden)
  DoubleCorrection detectDropAndCorrection(const DoubleCorrection& meta) const {
    if (!meta.some) return meta;
    double firstValue = apply(0);
    if (std::isnan(firstValue) || firstValue < meta.lastValue)
      return DoubleCorrection{true, meta.lastValue, meta.correction + meta.lastValue};
    return meta;
  }
  // lazy val corrected, DoubleVector.scala:325-342
  void forceCorrected() {
    if (correctedInit) return;
    correctedInit = true;
    int len = length();
    corrected.assign(len, 0.0);
    double last = std::numeric_limits<double>::lowest();   // Double.MinValue
    for (int pos = 0; pos < len; ++pos) {
      double nextVal = apply(pos);
      if (std::isnan(nextVal)) nextVal = 0;
      if (nextVal < last) { _correction += last; drops.push_back(pos); }
      corrected[pos] = nextVal + _correction;
      last = nextVal;
    }
  }
  // updateCorrection: plain :190-195 ; Correcting :375-391
  DoubleCorrection updateCorrection(const DoubleCorrection& meta) const {
    if (!correcting) {
      double last = apply(length() - 1);
      return DoubleCorrection{true, last, meta.some ? meta.correction : 0.0};
    }
    int index = length() - 1;
    double lastValue = 0.0;
    do { lastValue = apply(index); index -= 1; } while (std::isnan(lastValue) && index >= 0);
    if (std::isnan(lastValue)) lastValue = 0.0;
    // NOTE: uses _correction WITHOUT forcing `corrected` — reference behaviour, kept.
    return DoubleCorrection{true, lastValue, (meta.some ? meta.correction : 0.0) + _correction};
  }
  // correctedValue: plain :203-207 ; Correcting :350-365
  double correctedValue(int n, const DoubleCorrection& meta) {
    if (!correcting) return meta.some ? apply(n) + meta.correction : apply(n);
    forceCorrected();
    double c = (n >= (int)corrected.size()) ? corrected.back() : corrected[n];
    return meta.some ? c + meta.correction : c;
  }
  const std::vector<int>& dropPositions() { if (correcting) forceCorrected(); return drops; }
};

// =====================================================================================
// Encoders (appenders' optimize()) — used to generate byte-identical synthetic chunks
// =====================================================================================
namespace enc {
using Bytes = std::vector<uint8_t>;

// IntBinaryVector.minMaxToNbitsSigned, IntBinaryVector.scala:161-177
inline void minMaxToNbitsSigned(int32_t mn, int32_t mx, int& nbits, bool& sgn) {
  if (mn >= 0 && mx < 4) { nbits = 2; sgn = false; }
  else if (mn >= 0 && mx < 16) { nbits = 4; sgn = false; }
  else if (mn >= -128 && mx <= 127) { nbits = 8; sgn = true; }
  else if (mn >= 0 && mx < 256) { nbits = 8; sgn = false; }
  else if (mn >= -32768 && mx <= 32767) { nbits = 16; sgn = true; }
  else if (mn >= 0 && mx < 65536) { nbits = 16; sgn = false; }
  else { nbits = 32; sgn = true; }
}

// Frozen IntAppendingVector (PrimitiveAppendableVector, BinaryVector.scala:550-615; addData variants
// IntBinaryVector.scala:55-110).  Produces header + packed data.
inline Bytes intVectorNoNA(const int32_t* vals, int n, int nbits, bool sgn) {
  Bytes b(8, 0);
  auto flush_hdr = [&](int dataBytes, int bitShift) {
    setInt(b.data(), 4 + dataBytes);
    setShort(b.data() + 4, (int16_t)wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_INT_NOMASK));
    setShort(b.data() + 6, (int16_t)((nbits & 0x7f) | (sgn ? 0x80 : 0)));
    b[7] = (uint8_t)bitShift;
  };
  if (nbits >= 8) {
    int w = nbits / 8;
    b.resize(8 + (size_t)n * w);
    for (int i = 0; i < n; ++i) {
      if (w == 4) setInt(b.data() + 8 + i * 4, vals[i]);
      else if (w == 2) setShort(b.data() + 8 + i * 2, (int16_t)vals[i]);
      else setByte(b.data() + 8 + i, (int8_t)vals[i]);
    }
    flush_hdr(n * w, 0);
  } else {
    int dataBytes = 0, bitShift = 0; size_t writeOffset = 8;
    for (int i = 0; i < n; ++i) {
      if (b.size() <= writeOffset) b.resize(writeOffset + 1, 0);
      int orig = (bitShift == 0) ? 0 : (int8_t)b[writeOffset];
      b[writeOffset] = (uint8_t)(int8_t)(orig | (vals[i] << bitShift));
      if (bitShift == 0) dataBytes += 1;                     // bumpBitShift, BinaryVector.scala:577-582
      bitShift = (bitShift + nbits) % 8;
      if (bitShift == 0) writeOffset += 1;
    }
    b.resize(8 + dataBytes);
    flush_hdr(dataBytes, bitShift);
  }
  return b;
}

// DeltaDeltaVector.const, DeltaDeltaVector.scala:98-106
inline Bytes ddvConst(int n, int64_t init, int32_t slope) {
  Bytes b(24, 0);
  setInt(b.data(), 20);
  setInt(b.data() + 4, wire::make(wire::VECTORTYPE_DELTA2, wire::SUBTYPE_REPEATED));
  setInt(b.data() + 8, n);
  setLong(b.data() + 12, init);
  setInt(b.data() + 20, slope);
  return b;
}

// DeltaDeltaVector.fromLongVector, DeltaDeltaVector.scala:63-85 (+ getSlope :112-115, getDeltasMinMax :118-130,
// getNbitsSignedFromMinMax :132-135).  Returns empty if not eligible.
inline Bytes ddvFromLongs(const int64_t* v, int n, bool approxConst, int maxNBits = 32) {
  if (!(n > 2)) return {};
  int64_t slopeL = lsub(v[n - 1], v[0]) / (int64_t)(n - 1);
  if (!(slopeL < (int64_t)INT32_MAX && slopeL > (int64_t)INT32_MIN)) return {};
  int32_t slope = (int32_t)slopeL;
  int64_t baseValue = v[0];
  int32_t mx = INT32_MIN, mn = INT32_MAX;
  for (int i = 1; i < n; ++i) {
    baseValue = ladd(baseValue, slope);
    int64_t delta = lsub(v[i], baseValue);
    if (delta > (int64_t)INT32_MAX || delta < (int64_t)INT32_MIN) return {};
    mx = std::max(mx, (int32_t)delta); mn = std::min(mn, (int32_t)delta);
  }
  int nbits; bool sgn; minMaxToNbitsSigned(mn, mx, nbits, sgn);
  if (!(nbits <= maxNBits)) return {};
  if (mn == 0 && mx == 0) return ddvConst(n, v[0], slope);
  if (approxConst && mn >= -250 && mx <= 250) return ddvConst(n, v[0], slope);
  // DeltaDeltaAppendingVector, :293-340
  std::vector<int32_t> deltas(n);
  int64_t expected = v[0];
  for (int i = 0; i < n; ++i) { deltas[i] = (int32_t)lsub(v[i], expected); expected = ladd(expected, slope); }
  Bytes inner = intVectorNoNA(deltas.data(), n, nbits, sgn);
  Bytes b(20, 0);
  setInt(b.data() + 4, wire::make(wire::VECTORTYPE_DELTA2, wire::SUBTYPE_INT_NOMASK));
  setLong(b.data() + 8, v[0]);
  setInt(b.data() + 16, slope);
  b.insert(b.end(), inner.begin(), inner.end());
  setInt(b.data(), (int32_t)b.size() - 4);
  return b;
}

// LongAppendingVector frozen as-is (PRIMITIVE_NOMASK)
inline Bytes rawLongs(const int64_t* v, int n) {
  Bytes b(8 + (size_t)n * 8, 0);
  setInt(b.data(), 4 + n * 8);
  setShort(b.data() + 4, (int16_t)wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_PRIMITIVE_NOMASK));
  setShort(b.data() + 6, (int16_t)((64 & 0x7f) | 0x80));
  for (int i = 0; i < n; ++i) setLong(b.data() + 8 + (size_t)i * 8, v[i]);
  return b;
}
// TimestampAppendingVector.optimize, LongBinaryVector.scala:333-340
inline Bytes timestamps(const int64_t* v, int n) {
  Bytes d = ddvFromLongs(v, n, /*approxConst*/true);
  return d.empty() ? rawLongs(v, n) : d;
}
// LongAppendingVector.optimize → LongBinaryVector.optimize, :76-87
inline Bytes longs(const int64_t* v, int n) {
  Bytes d = ddvFromLongs(v, n, false);
  return d.empty() ? rawLongs(v, n) : d;
}

inline Bytes rawDoubles(const double* v, int n) {
  Bytes b(8 + (size_t)n * 8, 0);
  setInt(b.data(), 4 + n * 8);
  setShort(b.data() + 4, (int16_t)wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_PRIMITIVE_NOMASK));
  setShort(b.data() + 6, (int16_t)((64 & 0x7f) | 0x80));
  for (int i = 0; i < n; ++i) setDouble(b.data() + 8 + (size_t)i * 8, v[i]);
  return b;
}

// DoubleCounterAppender.addData drop detection, DoubleVector.scala:456-466
inline bool counterDropFlag(const double* v, int n) {
  double last = std::numeric_limits<double>::lowest();
  bool drop = false;
  for (int i = 0; i < n; ++i) {
    if (std::isnan(v[i]) || v[i] < last) drop = true;
    if (!std::isnan(v[i])) last = v[i];
  }
  return drop;
}

// DoubleVector.optimize (DoubleVector.scala:86-96) + LongDoubleWrapper.allIntegrals (:521-533) +
// DoubleCounterAppender.optimize (:468-473)
inline Bytes doubles(const double* v, int n, bool detectDrops) {
  const double MaxLongDouble = (double)INT64_MAX;
  int nonInts = 0;
  for (int i = 0; i < n; ++i) if (v[i] > MaxLongDouble || std::rint(v[i]) != v[i]) nonInts++;
  Bytes out;
  if (nonInts == 0) {
    std::vector<int64_t> lv(n);
    for (int i = 0; i < n; ++i) {           // Double.toLong saturates
      double d = v[i];
      lv[i] = d >= 9.2233720368547758e18 ? INT64_MAX : (d <= -9.2233720368547758e18 ? INT64_MIN : (int64_t)d);
    }
    out = ddvFromLongs(lv.data(), n, false);
  }
  if (out.empty()) out = rawDoubles(v, n);
  if (detectDrops && counterDropFlag(v, n)) pv_markDrop(out.data());
  return out;
}

// Our XOR container; payload = NibblePack.packDoubles.  dropFlag semantics as DoubleCounterAppender.
inline Bytes doublesXor(const double* v, int n, bool detectDrops) {
  if (n <= 0) throw std::invalid_argument("doublesXor: empty");
  int ngroups = (n - 1 + 7) / 8;
  int hdr = 16 + 2 * ngroups;
  int payloadOff = (hdr + 7) & ~7;
  Bytes payload;
  int end = nibble::packDoubles(v, n, payload, 0);
  payload.resize(end);
  // group offsets: walk the stream (each group's size from its own header, NibblePack.scala:405)
  std::vector<uint16_t> offs(ngroups);
  int pos = 8;
  for (int g = 0; g < ngroups; ++g) {
    if (pos - 8 > 0xffff) throw std::invalid_argument("doublesXor: payload too large for u16 offsets");
    offs[g] = (uint16_t)(pos - 8);
    uint8_t mask = payload[pos];
    if (mask == 0) pos += 1;
    else { int numBits = (((payload[pos + 1] & 0xff) >> 4) + 1) * 4; pos += 2 + (numBits * __builtin_popcount(mask) + 7) / 8; }
  }
  if (pos != end) throw std::logic_error("doublesXor: group walk mismatch");
  int total = (payloadOff + end + 7) & ~7;
  Bytes b(total, 0);
  setInt(b.data(), total - 4);
  setShort(b.data() + 4, (int16_t)wire::make(wire::VECTORTYPE_BINSIMPLE, wire::SUBTYPE_XOR_NIBBLE));
  setShort(b.data() + 6, 0);
  setInt(b.data() + 8, n);
  setShort(b.data() + 12, (int16_t)ngroups);
  setShort(b.data() + 14, (int16_t)payloadOff);
  for (int g = 0; g < ngroups; ++g) setShort(b.data() + 16 + 2 * g, (int16_t)offs[g]);
  std::memcpy(b.data() + payloadOff, payload.data(), end);
  if (detectDrops && counterDropFlag(v, n)) pv_markDrop(b.data());
  return b;
}
} // namespace enc

// =====================================================================================
// ChunkSetInfo  (store/ChunkSetInfo.scala:133-154; chunkID store/package.scala:112-123)
// =====================================================================================
namespace csi {
constexpr int OffsetChunkID = 0, OffsetNumRows = 8, OffsetIngestionTime = 12, OffsetEndTime = 20, OffsetVectors = 28;
inline int64_t chunkID(int64_t startTime, int64_t ingestionTimeSec) {
  int64_t m = ingestionTimeSec % (48LL * 24 * 60 * 60);
  if (m < 0) m += 48LL * 24 * 60 * 60;                                          // Math.floorMod
  return (int64_t)((1ull << 63) ^ ((uint64_t)startTime << 22)) | m;
}
inline int64_t startTimeFromChunkID(int64_t id) { return (int64_t)(((1ull << 63) ^ (uint64_t)id) >> 22); }
inline int64_t startTime(Ptr info) { return startTimeFromChunkID(getLong(info + OffsetChunkID)); }
inline int32_t numRows(Ptr info)   { return getInt(info + OffsetNumRows); }
inline int64_t endTime(Ptr info)   { return getLong(info + OffsetEndTime); }
inline Ptr vectorPtr(Ptr info, int col) { return (Ptr)(uintptr_t)getLong(info + OffsetVectors + 8 * col); }
// ChunkSetInfo.intersection(time1, time2).isDefined, ChunkSetInfo.scala:99-108
inline bool intersects(Ptr info, int64_t t1, int64_t t2) {
  if (t1 > t2) return false;
  return t1 <= endTime(info) && t2 >= startTime(info);
}
}

} // namespace fo
