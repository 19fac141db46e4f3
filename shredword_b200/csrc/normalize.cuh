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
// nothing. So: count per 4 KB tile -> exclusive scan (CUB) -> the same code again, staged in shared memory and copied out
// with coalesced stores. Traffic: the text twice, the output once.
#pragma once

#include <cub/device/device_scan.cuh>

#include "device_util.cuh"

namespace swb {

constexpr int NORM_THREADS = 256;
constexpr int NORM_TILE = NORM_THREADS * 16;

__device__ __forceinline__ bool norm_is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }  // normalize.cpp:13-15

// bytes that text[i] contributes (0, 1 or 4); *marker = the 3 marker bytes come first
__device__ __forceinline__ uint32_t norm_emit(const uint8_t *__restrict__ text, uint64_t n, uint64_t i, bool &marker) {
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

__global__ void __launch_bounds__(NORM_THREADS)
norm_count(const uint8_t *__restrict__ text, uint64_t n, unsigned long long *__restrict__ tile_bytes) {
  __shared__ unsigned int s_sum;
  const uint64_t n_tiles = (n + NORM_TILE - 1) / NORM_TILE;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const uint64_t b0 = tile * NORM_TILE + (uint64_t)threadIdx.x * 16;
    unsigned int mine = 0;
    for (int k = 0; k < 16; k++) {
      bool mk;
      if (b0 + k < n) mine += norm_emit(text, n, b0 + k, mk);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_sum, mine);
    __syncthreads();
    if (threadIdx.x == 0) tile_bytes[tile] = s_sum;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NORM_THREADS)
norm_write(const uint8_t *__restrict__ text, uint64_t n, const unsigned long long *__restrict__ tile_off, uint8_t *__restrict__ out, uint64_t cap) {
  __shared__ uint8_t stage[NORM_TILE * 4];
  __shared__ uint32_t wt[NORM_THREADS / 32];
  const uint64_t n_tiles = (n + NORM_TILE - 1) / NORM_TILE;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t b0 = tile * NORM_TILE + (uint64_t)threadIdx.x * 16;
    uint8_t buf[64];
    uint32_t mine = 0;
    for (int k = 0; k < 16; k++) {
      if (b0 + k >= n) break;
      bool mk;
      const uint32_t e = norm_emit(text, n, b0 + k, mk);
      if (!e) continue;
      if (mk) { buf[mine++] = 0xE2; buf[mine++] = 0x96; buf[mine++] = 0x81; }
      if (e == 1 || e == 4) {
        const uint8_t c = text[b0 + k];
        buf[mine++] = (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c;
      }
    }
    // block exclusive scan of the byte counts
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) wt[w] = inc;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < NORM_THREADS / 32; i++) { if (i < w) base += wt[i]; total += wt[i]; }
    const uint32_t at = base + inc - mine;
    for (uint32_t k = 0; k < mine; k++) stage[at + k] = buf[k];
    __syncthreads();
    const uint64_t o0 = tile_off[tile];
    for (uint32_t k = threadIdx.x; k < total; k += NORM_THREADS)
      if (o0 + k < cap) out[o0 + k] = stage[k];
    __syncthreads();
  }
}

// text (device, n bytes) -> out (device, capacity cap). Returns the length of the full result (may exceed cap: then only the
// first cap bytes were written). `launches` is incremented per kernel.
inline uint64_t normalize_device(const uint8_t *d_text, uint64_t n, uint8_t *d_out, uint64_t cap, cudaStream_t stream, int sms, uint64_t *launches) {
  if (n == 0) return 0;
  const uint64_t n_tiles = (n + NORM_TILE - 1) / NORM_TILE;
  DevBuf<unsigned long long> tile_bytes(n_tiles + 1), tile_off(n_tiles + 1);
  SWB_CUDA(cudaMemsetAsync(tile_bytes.get() + n_tiles, 0, 8, stream));
  const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)sms * 8);
  norm_count<<<grid, NORM_THREADS, 0, stream>>>(d_text, n, tile_bytes.get());
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, tile_bytes.get(), tile_off.get(), (int64_t)(n_tiles + 1), stream);
  DevBuf<uint8_t> tmp(tb);
  SWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tb, tile_bytes.get(), tile_off.get(), (int64_t)(n_tiles + 1), stream));
  unsigned long long total = 0;
  SWB_CUDA(cudaMemcpyAsync(&total, tile_off.get() + n_tiles, 8, cudaMemcpyDeviceToHost, stream));
  norm_write<<<grid, NORM_THREADS, 0, stream>>>(d_text, n, tile_off.get(), d_out, cap);
  SWB_CUDA(cudaGetLastError());
  SWB_CUDA(cudaStreamSynchronize(stream));
  if (launches) *launches += 4;
  return total;
}

}  // namespace swb
