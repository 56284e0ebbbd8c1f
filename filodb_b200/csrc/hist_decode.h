// NibblePack group decode for the histogram kernel, written so that the same code compiles for the device and for the host
// (tests/cpp/hist_decode_check.cpp runs it on the CPU against the oracle's restatement of NibblePack.unpack8).
//
// A group (NibblePack.scala:395-447; doc/compression.md:36-87): u8 nonzero mask; when non-zero, u8 = (numNibbles-1)<<4 | trailing
// zero nibbles, then numNibbles*4-bit fields for the set bits only, packed LSB first.  Group bytes = 2 + ceil(numBits*popc/8).
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define FILO_HDI __host__ __device__ __forceinline__
#else
#define FILO_HDI inline
#endif

namespace filo {

FILO_HDI uint64_t hd_ld64_aligned(const uint8_t* p) {       // p is 8-byte aligned
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const uint64_t*>(p);
#else
  uint64_t v; std::memcpy(&v, p, 8); return v;
#endif
}
FILO_HDI int hd_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// Bytes the group at p occupies (needs p[0], p[1] readable).
FILO_HDI int nibble_group_bytes(const uint8_t* p) {
  const uint32_t mask = p[0];
  if (mask == 0) return 1;
  const int numBits = ((p[1] >> 4) + 1) * 4;
  return 2 + ((numBits * hd_popc(mask) + 7) >> 3);
}

// Decodes one group whose bytes all lie inside the input (the caller checked nibble_group_bytes(p) <= cap).  Reads aligned
// 8-byte words around the fields: up to 15 bytes past the group may be touched (never interpreted), so the buffer needs that
// much slack.  out[] is indexed with compile-time constants only (stays in registers).  Returns the bytes consumed.
FILO_HDI int nibble_unpack8_inbounds(const uint8_t* p, uint64_t out[8]) {
  const uint32_t mask = p[0];
  if (mask == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = 0;
    return 1;
  }
  const uint32_t hdr = p[1];
  const int numBits = (int)((hdr >> 4) + 1) * 4, trailing = (int)(hdr & 0x0f) * 4;
  const uint64_t fmask = numBits >= 64 ? ~0ull : ((1ull << numBits) - 1);
  const uint8_t* d = p + 2;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(d);
  const uint8_t* base = d - (a0 & 7);
  const int lead = (int)(a0 & 7) * 8;            // bit offset of the first field inside the aligned word stream
  int bit = lead;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint64_t v = 0;
    if (mask & (1u << i)) {
      const uint8_t* w = base + ((bit >> 6) << 3);
      const int s = bit & 63;
      const uint64_t w0 = hd_ld64_aligned(w);
      uint64_t f = w0 >> s;
      if (s + numBits > 64) f |= hd_ld64_aligned(w + 8) << (64 - s);      // s > 0 here
      v = (f & fmask) << trailing;
      bit += numBits;
    }
    out[i] = v;
  }
  return 2 + ((numBits * hd_popc(mask) + 7) >> 3);
}

} // namespace filo
