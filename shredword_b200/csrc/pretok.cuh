// pretok.cuh -- opt-in regex pre-tokenisation on the device (reference shredword/base.py:38-58).
//
// Output = the text with one ' ' behind every piece of the reference's split pattern, and the trainer's four delimiter bytes
// (' ' \t \n \r) INSIDE pieces mapped to 0x1C-0x1F, so that the whitespace-splitting trainer / encoder sees exactly the
// reference's pieces as its words. Where a piece ends is a local predicate of the character (pretok_rules.hpp), so this is a
// map with a variable output length (1 or 2 bytes per input byte): one pass of stream_map.cuh -- tile + halo staged in shared
// memory by a bulk copy, decoupled look-back for the output offset. Traffic: the text once, the output (<= 2 x the text) once.
#pragma once

#include <mutex>
#include <vector>

#include "pretok_rules.hpp"
#include "stream_map.cuh"

namespace swb {

struct PretokEmit {  // code: 1 = a piece ends behind this byte
  static constexpr int MAX_OUT = 2;
  const uint8_t *__restrict__ tab;  // 2-bit class per code point
  __device__ __forceinline__ bool fast(const TextWin &t, uint64_t n, uint64_t i, uint32_t &code) const { return pt_fast(t, n, i, &code); }
  __device__ __forceinline__ uint32_t slow(const TextWin &t, uint64_t n, uint64_t i) const {
    uint8_t b0;
    return pt_emit(t, n, tab, i, &b0) - 1u;
  }
  __device__ __forceinline__ uint32_t expand(const TextWin &t, uint64_t i, uint32_t code, uint8_t *out) const {
    const uint8_t b = tx_near(t, i);
    out[0] = b == ' ' ? 0x1C : b == '\t' ? 0x1D : b == '\n' ? 0x1E : b == '\r' ? 0x1F : b;
    out[1] = ' ';
    return 1u + code;
  }
};

// the 2-bit class table, uploaded once per device
inline const uint8_t *pretok_table(int dev) {
  static std::mutex mu;
  static std::vector<uint8_t *> per_dev(64, nullptr);
  std::lock_guard<std::mutex> g(mu);
  if (dev < 0 || dev >= (int)per_dev.size()) throw Error("pretok: device index out of range");
  if (!per_dev[dev]) {
    std::vector<uint8_t> host(PT_TABLE_BYTES);
    pretok_build_table(host.data());
    uint8_t *d = nullptr;
    SWB_CUDA(cudaMalloc(&d, PT_TABLE_BYTES));
    SWB_CUDA(cudaMemcpy(d, host.data(), PT_TABLE_BYTES, cudaMemcpyHostToDevice));
    per_dev[dev] = d;
  }
  return per_dev[dev];
}

// text (device, n bytes) -> out (device, capacity cap). Returns the length of the full result (may exceed cap: then only the
// first cap bytes were written; 2 * n always suffices).
inline uint64_t pretokenize_device(const uint8_t *d_text, uint64_t n, uint8_t *d_out, uint64_t cap, cudaStream_t stream, int dev, int sms,
                                   uint64_t *launches) {
  if (n == 0) return 0;
  PretokEmit f;
  f.tab = pretok_table(dev);
  return stream_map_run(d_text, n, f, d_out, cap, stream, sms, launches);
}

}  // namespace swb
