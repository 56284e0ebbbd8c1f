// CPU check of filodb_b200/csrc/hist_decode.h (the in-bounds NibblePack group decoder of the histogram kernel) against the
// oracle's restatement of NibblePack.pack8 / unpack8.  Test infrastructure: built and run by tests/test_abi.py.
#include "../../filodb_b200/csrc/hist_decode.h"
#include "../../oracle/filo_format.hpp"
#include <cstdio>
#include <random>
#include <vector>

int main() {
  std::mt19937_64 rng(12345);
  long groups = 0;
  for (int iter = 0; iter < 200000; ++iter) {
    uint64_t in[8];
    const int width = (int)(rng() % 65), tz = (int)(rng() % 16);
    for (int i = 0; i < 8; ++i) {
      uint64_t v = width == 0 ? 0 : (rng() >> (64 - width));
      if (rng() % 3 == 0) v = 0;
      if (tz * 4 < 64) v = (v >> (tz * 4)) << (tz * 4); else v = 0;
      in[i] = v;
    }
    std::vector<uint8_t> buf;
    const int lead = (int)(rng() % 9);                     // every alignment of the group start
    buf.assign((size_t)lead, 0xAB);
    const int end = fo::nibble::pack8(in, buf, lead);
    buf.resize((size_t)end);
    const int used_bytes = end - lead;
    for (int k = 0; k < 24; ++k) buf.push_back((uint8_t)rng());     // slack (garbage: must not leak into the fields)
    // place the buffer so that its start is 8-byte aligned + lead gives all phases
    std::vector<uint64_t> backing((buf.size() + 15) / 8 + 1);
    uint8_t* base = reinterpret_cast<uint8_t*>(backing.data());
    std::memcpy(base, buf.data(), buf.size());
    const uint8_t* p = base + lead;
    if (filo::nibble_group_bytes(p) != used_bytes) { std::printf("FAIL bytes iter %d: %d vs %d\n", iter, filo::nibble_group_bytes(p), used_bytes); return 1; }
    uint64_t out[8];
    const int used = filo::nibble_unpack8_inbounds(p, out);
    uint64_t ref[8];
    fo::Ptr qp = p; int qcap = used_bytes;
    if (fo::nibble::unpack8(qp, qcap, ref) != fo::nibble::Ok) { std::printf("FAIL oracle unpack iter %d\n", iter); return 1; }
    if (used != used_bytes) { std::printf("FAIL used iter %d\n", iter); return 1; }
    for (int i = 0; i < 8; ++i) if (out[i] != in[i] || ref[i] != in[i]) { std::printf("FAIL value iter %d i %d: %llx vs %llx (oracle %llx)\n", iter, i, (unsigned long long)out[i], (unsigned long long)in[i], (unsigned long long)ref[i]); return 1; }
    ++groups;
  }
  std::printf("OK %ld groups\n", groups);
  return 0;
}
