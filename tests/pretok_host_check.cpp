// TEST INFRASTRUCTURE: runs the per-byte rules of shredword_b200/csrc/pretok_rules.hpp and normalize_rules.hpp (the functions the
// device kernels call) on the CPU, in the order the kernel uses them -- cheap test first, general rule for the rest, then the
// expansion of the code to bytes -- so that they can be checked against the reference's own outputs without a GPU
// (tests/test_pretok.py and tests/test_normalize.py build this with g++ into a temporary directory). Every byte the cheap test
// settles is also put through the general rule: a disagreement returns -2. Not part of the library.
#include "../shredword_b200/csrc/normalize_rules.hpp"
#include "../shredword_b200/csrc/pretok_rules.hpp"

#include <vector>

extern "C" long long pretok_host_check(const unsigned char *text, unsigned long long n, unsigned char *out, unsigned long long cap,
                                       unsigned long long *n_fast) {
  static std::vector<uint8_t> tab;
  if (tab.empty()) { tab.resize(swb::PT_TABLE_BYTES); swb::pretok_build_table(tab.data()); }
  unsigned long long o = 0, fast = 0;
  for (unsigned long long i = 0; i < n; i++) {
    uint8_t b;
    const uint32_t general = swb::pt_emit(text, n, tab.data(), i, &b) - 1u;
    uint32_t code = 0;
    if (swb::pt_fast(text, n, i, &code)) {
      ++fast;
      if (code != general) return -2;
    } else {
      code = general;
    }
    if (o < cap) out[o] = b;
    ++o;
    if (code) { if (o < cap) out[o] = ' '; ++o; }
  }
  if (n_fast) *n_fast = fast;
  return (long long)o;
}

extern "C" long long normalize_host_check(const unsigned char *text, unsigned long long n, unsigned char *out, unsigned long long cap,
                                          unsigned long long *n_fast) {
  unsigned long long o = 0, fast = 0;
  auto put = [&](uint8_t b) { if (o < cap) out[o] = b; ++o; };
  for (unsigned long long i = 0; i < n; i++) {
    bool mk;
    const uint32_t general = swb::norm_emit(text, n, i, mk);
    uint32_t code = 0;
    if (swb::norm_fast(text, n, i, &code)) {
      ++fast;
      if (code != general) return -2;
    } else {
      code = general;
    }
    if (code >= 3u) { put(0xE2); put(0x96); put(0x81); }
    if (code == 1u || code == 4u) { const uint8_t c = text[i]; put((c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c); }
  }
  if (n_fast) *n_fast = fast;
  return (long long)o;
}
