// Device chunk-arena record layout (host + device).  One record per time series:
//
//   RecordHeader (16 B) | ChunkEntry[n_chunks] (32 B each) | BinaryVectors (verbatim bytes, each 8-byte aligned) | pad
//
// Records start 16-byte aligned and their size is a multiple of 16 so that a whole record can be moved with one
// TMA bulk copy (cp.async.bulk) or with coalesced 16-byte loads.  Vector bytes are FiloDB's wire format verbatim
// (SURVEY.md Appendix A); the one sanctioned change is pointers -> offsets: a ChunkEntry is ChunkSetInfo
// (core/src/main/scala/filodb.core/store/ChunkSetInfo.scala:133-154) with startTime decoded from the chunkID and the
// vector pointers replaced by byte offsets from the record start.
#pragma once
#include <stdint.h>

namespace filo {

struct RecordHeader {
  uint32_t rec_bytes;     // total record bytes (multiple of 16)
  uint32_t n_chunks;
  uint32_t n_rows;        // Σ value-vector lengths (decode scratch sizing)
  uint32_t flags;         // REC_* below
};
enum : uint32_t {
  REC_ALL_TS_CONST = 1u,      // every chunk's timestamp vector is a const DDV (closed-form row search)
  REC_ANY_DROP     = 2u,      // some value vector has the counter drop flag
  REC_ANY_DECODE   = 4u,      // some value vector needs decoding (XOR container / DDV-long)
  REC_HIST         = 8u,      // the value column holds histogram vectors (HistogramVector.scala)
};

struct ChunkEntry {
  int64_t  start_time;    // startTimeFromChunkID, store/package.scala:112-123
  int64_t  end_time;      // ChunkSetInfo +20
  int32_t  num_rows;      // ChunkSetInfo +8
  uint32_t ts_off;        // byte offset of the timestamp BinaryVector from record start
  uint32_t val_off;       // byte offset of the value BinaryVector
  uint32_t row_base;      // Σ value-vector lengths of previous chunks
};
static_assert(sizeof(RecordHeader) == 16, "RecordHeader");
static_assert(sizeof(ChunkEntry) == 32, "ChunkEntry");

// wire words (WireFormat.scala:7-53): (subtype << 8) | major
constexpr int WIRE_DDV        = (0x08 << 8) | 0x08;   // DELTA2 / INT_NOMASK
constexpr int WIRE_DDV_CONST  = (0x06 << 8) | 0x08;   // DELTA2 / REPEATED
constexpr int WIRE_MASKED     = (0x00 << 8) | 0x06;   // BINSIMPLE / PRIMITIVE
constexpr int WIRE_RAW64      = (0x05 << 8) | 0x06;   // BINSIMPLE / PRIMITIVE_NOMASK
constexpr int WIRE_XOR        = (0x21 << 8) | 0x06;
constexpr int WIRE_H_SIMPLE   = (0x10 << 8) | 0x09;   // HISTOGRAM / H_SIMPLE   (WireFormat.scala:17,35)
constexpr int WIRE_H_EXP_SIMPLE = (0x13 << 8) | 0x09; // HISTOGRAM / H_EXP_SIMPLE (WireFormat.scala:38): a blob with its own scheme per row; declined
constexpr int WIRE_H_SECTDELTA = (0x12 << 8) | 0x09;  // HISTOGRAM / H_SECTDELTA (WireFormat.scala:37)   // BINSIMPLE / XOR_NIBBLE (this repo's container)

// XOR container header (oracle/filo_format.hpp documents the same layout):
//  +0 i32 numBytes  +4 u16 wire  +6 u16 flags(bit15 drop)  +8 i32 n  +12 u16 numGroups  +14 u16 payloadOff
//  +16 u16 groupOff[numGroups]   +payloadOff: f64 first, NibblePack groups
constexpr int XOR_OFF_N = 8, XOR_OFF_NGROUPS = 12, XOR_OFF_PAYLOAD = 14, XOR_OFF_GROUPTAB = 16;

inline
#ifdef __CUDACC__
__host__ __device__
#endif
uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) & ~(a - 1); }

} // namespace filo
