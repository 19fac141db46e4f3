// TEST INFRASTRUCTURE: runs the piece-end rules of shredword_b200/csrc/pretok_rules.hpp (the functions the device kernels
// call) on the CPU, so that they can be checked against the reference's own apply_regex outputs without a GPU
// (tests/test_pretok.py builds this with g++ into a temporary directory). Not part of the library.
#include "../shredword_b200/csrc/pretok_rules.hpp"

#include <vector>

extern "C" long long pretok_host_check(const unsigned char *text, unsigned long long n, unsigned char *out, unsigned long long cap) {
  static std::vector<uint8_t> tab;
  if (tab.empty()) { tab.resize(swb::PT_TABLE_BYTES); swb::pretok_build_table(tab.data()); }
  unsigned long long o = 0;
  for (unsigned long long i = 0; i < n; i++) {
    uint8_t b;
    const uint32_t e = swb::pt_emit(text, n, tab.data(), i, &b);
    if (o < cap) out[o] = b;
    ++o;
    if (e == 2) { if (o < cap) out[o] = ' '; ++o; }
  }
  return (long long)o;
}
