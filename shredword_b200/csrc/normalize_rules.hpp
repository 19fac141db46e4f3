// normalize_rules.hpp -- what one input byte contributes to the normalised text (reference csrc/bpe/normalize.cpp:24-59 applied
// per line; see normalize.cuh). Compiles for the device and for the host: tests/pretok_host_check.cpp runs these functions on the
// CPU against the oracle -- test infrastructure only; the library exports the device path alone.
#pragma once

#include <cstdint>

#ifndef SWB_HD
#ifdef __CUDACC__
#define SWB_HD __host__ __device__ __forceinline__
#else
#define SWB_HD inline
#endif
#endif

namespace swb {

#ifndef SWB_TX_NEAR
#define SWB_TX_NEAR
// text[i] for an i within 128 bytes of the byte being decided (a windowed text answers it from its staged copy, unchecked)
template <class T>
SWB_HD uint8_t tx_near(const T &text, uint64_t i) { return text[i]; }
#endif

SWB_HD bool norm_is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }  // normalize.cpp:13-15

// bytes that text[i] contributes (0, 1 or 4); *marker = the 3 marker bytes come first
template <class T>
SWB_HD uint32_t norm_emit(const T &text, uint64_t n, uint64_t i, bool &marker) {
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

// cheap cases: the separator itself, whitespace, and a byte that neither follows whitespace nor sits within three bytes of
// its line's end (so neither the marker nor the trailing-triple rule can concern it). false = norm_emit is needed.
template <class T>
SWB_HD bool norm_fast(const T &text, uint64_t n, uint64_t i, uint32_t *code) {
  const uint8_t c = tx_near(text, i);
  if (c == '\n') { *code = 1; return true; }
  if (norm_is_ws(c)) { *code = 0; return true; }
  if (i + 3 >= n || tx_near(text, i + 1) == '\n' || tx_near(text, i + 2) == '\n' || tx_near(text, i + 3) == '\n') return false;
  if (i > 0) { const uint8_t p = tx_near(text, i - 1); if (norm_is_ws(p) && p != '\n') return false; }
  *code = 1;
  return true;
}

}  // namespace swb
