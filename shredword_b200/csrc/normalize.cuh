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

#include "stream_map.cuh"

namespace swb {

__device__ __forceinline__ bool norm_is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }  // normalize.cpp:13-15

// bytes that text[i] contributes (0, 1 or 4); *marker = the 3 marker bytes come first
template <class T>
__device__ __forceinline__ uint32_t norm_emit(const T &text, uint64_t n, uint64_t i, bool &marker) {
  marker = false;
  const uint8_t c = text[i];
  if (c == '\n') return 1;     // line separator, kept
  if (norm_is_ws(c)) return 0;
  // the literal bytes E2 96 81 at the very end of a line (which then ends with a non-whitespace byte) are dropped
  {
    // k = position of this byte among the last three bytes of the line, if it is one of them
    for (int k = 0; k < 3; k++) {
      const uint64_t e = i + (uint64_t)(3 - k);  // candidate line end (index one past the last byte) if this is byte k of the triple
      if (e > n || (e < n && text[e] != '\n')) continue;
      if (e < 3) continue;
      const uint64_t s = e - 3;
      if (text[s] == 0xE2 && text[s + 1] == 0x96 && text[s + 2] == 0x81) {
        // the three bytes must belong to this line (no newline among them: none of them is one)
        if (k == 0) {  // the marker in front of the triple, if any, is still emitted
          bool mk = false;
          if (i > 0 && norm_is_ws(text[i - 1]) && text[i - 1] != '\n') {
            uint64_t j = i - 1;
            while (j > 0 && norm_is_ws(text[j]) && text[j] != '\n') --j;
            mk = !norm_is_ws(text[j]);
          }
          marker = mk;
          return mk ? 3u : 0u;
        }
        return 0;
      }
    }
  }
  uint32_t len = 1;
  if (i > 0 && norm_is_ws(text[i - 1]) && text[i - 1] != '\n') {
    uint64_t j = i - 1;
    while (j > 0 && norm_is_ws(text[j]) && text[j] != '\n') --j;
    if (!norm_is_ws(text[j])) { marker = true; len = 4; }  // (j == 0 and whitespace, or a newline: the run starts the line)
  }
  return len;
}

struct NormEmit {
  static constexpr int MAX_OUT = 4;
  __device__ __forceinline__ uint32_t operator()(const TextWin &t, uint64_t n, uint64_t i, uint8_t *out) const {
    bool mk;
    const uint32_t e = norm_emit(t, n, i, mk);
    uint32_t k = 0;
    if (mk) { out[0] = 0xE2; out[1] = 0x96; out[2] = 0x81; k = 3; }
    if (e == 1 || e == 4) {
      const uint8_t c = t[i];
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
