// normalize.cuh -- optional device pre-pass: the reference's normalize_line (reference csrc/bpe/normalize.cpp:24-59)
// applied to every line of a text. The reference function has no caller in the reference; here it is
// an opt-in step in front of the trainer / encoder (swb_normalize).
//
// Per line (lines end at '\n', which is kept as the separator):
//   * ASCII upper case -> lower case (tolower in the C locale; every other byte unchanged)        normalize.cpp:47
//   * a run of ' ' '\t' '\r' between two other bytes -> U+2581 (E2 96 81), once per run             normalize.cpp:34-45
//   * runs at the start and at the end of the line disappear                                       normalize.cpp:32, 52-55
//   * quirk kept: the trailing-marker removal looks at the OUTPUT BYTES, so a line that ends with the literal bytes
//     E2 96 81 loses them (once)                                                                   normalize.cpp:52-55
// The reference stops at the caller's output_size (no caller, no fixed value); this pass has no such limit.
//
// What one input byte contributes is a local function: a byte that is not whitespace emits itself (lower-cased),
// preceded by the marker if the byte before it is whitespace and the run of whitespace before it does not reach back to
// the start of the line (walked backwards: runs are short, and each run is walked by exactly one thread); whitespace emits
// nothing. So this is a map with a variable output length (0, 1, 3 or 4 bytes per input byte): one pass of stream_map.cuh --
// tile + halo staged in shared memory by a bulk copy, decoupled look-back for the output offset. Traffic: the text once, the
// output once.
#pragma once

#include "normalize_rules.hpp"
#include "stream_map.cuh"

namespace swb {

struct NormEmit {  // code = what norm_emit returns: 0 nothing, 1 the byte, 3 the marker, 4 marker + byte
  static constexpr int MAX_OUT = 4;
  __device__ __forceinline__ bool fast(const TextWin &t, uint64_t n, uint64_t i, uint32_t &code) const { return norm_fast(t, n, i, &code); }
  __device__ __forceinline__ uint32_t slow(const TextWin &t, uint64_t n, uint64_t i) const {
    bool mk;
    return norm_emit(t, n, i, mk);
  }
  __device__ __forceinline__ uint32_t expand(const TextWin &t, uint64_t i, uint32_t code, uint8_t *out) const {
    uint32_t k = 0;
    if (code >= 3u) { out[0] = 0xE2; out[1] = 0x96; out[2] = 0x81; k = 3; }
    if (code == 1u || code == 4u) {
      const uint8_t c = tx_near(t, i);
      out[k++] = (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c;
    }
    return k;
  }
};

// text (device, n bytes) -> out (device, capacity cap). Returns the length of the full result (may exceed cap: then only the
// first cap bytes were written). `launches` is incremented per kernel.
inline uint64_t normalize_device(const uint8_t *d_text, uint64_t n, uint8_t *d_out, uint64_t cap, cudaStream_t stream, int sms, uint64_t *launches) {
  if (n == 0) return 0;
  return stream_map_run(d_text, n, NormEmit{}, d_out, cap, stream, sms, launches);
}

}  // namespace swb
