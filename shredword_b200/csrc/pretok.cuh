// pretok.cuh -- opt-in regex pre-tokenisation on the device (reference shredword/base.py:38-58).
//
// Output = the text with one ' ' behind every piece of the reference's split pattern, and the trainer's four delimiter bytes
// (' ' \t \n \r) INSIDE pieces mapped to 0x1C-0x1F, so that the whitespace-splitting trainer / encoder sees exactly the
// reference's pieces as its words. Where a piece ends is a local predicate of the character (pretok_rules.hpp), so this is a
// map with a variable output length: count per 4 KB tile -> exclusive scan (CUB) -> the same code again, staged in shared
// memory and copied out with coalesced stores. Traffic: the text twice, the output (<= 2 x the text) once.
#pragma once

#include <cub/device/device_scan.cuh>

#include <mutex>
#include <vector>

#include "device_util.cuh"
#include "pretok_rules.hpp"

namespace swb {

constexpr int PT_THREADS = 256;
constexpr int PT_PER_THREAD = 16;
constexpr int PT_TILE = PT_THREADS * PT_PER_THREAD;

__global__ void __launch_bounds__(PT_THREADS)
pretok_count(const uint8_t *__restrict__ text, uint64_t n, const uint8_t *__restrict__ tab, unsigned long long *__restrict__ tile_bytes) {
  __shared__ unsigned int s_sum;
  const uint64_t n_tiles = (n + PT_TILE - 1) / PT_TILE;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const uint64_t b0 = tile * PT_TILE + (uint64_t)threadIdx.x * PT_PER_THREAD;
    unsigned int mine = 0;
    for (int k = 0; k < PT_PER_THREAD; k++) {
      uint8_t b;
      if (b0 + k < n) mine += pt_emit(text, n, tab, b0 + k, &b);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_sum, mine);
    __syncthreads();
    if (threadIdx.x == 0) tile_bytes[tile] = s_sum;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PT_THREADS)
pretok_write(const uint8_t *__restrict__ text, uint64_t n, const uint8_t *__restrict__ tab, const unsigned long long *__restrict__ tile_off,
             uint8_t *__restrict__ out, uint64_t cap) {
  __shared__ uint8_t stage[PT_TILE * 2];
  __shared__ uint32_t wt[PT_THREADS / 32];
  const uint64_t n_tiles = (n + PT_TILE - 1) / PT_TILE;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t b0 = tile * PT_TILE + (uint64_t)threadIdx.x * PT_PER_THREAD;
    uint8_t buf[2 * PT_PER_THREAD];
    uint32_t mine = 0;
    for (int k = 0; k < PT_PER_THREAD; k++) {
      if (b0 + k >= n) break;
      uint8_t b;
      const uint32_t e = pt_emit(text, n, tab, b0 + k, &b);
      buf[mine++] = b;
      if (e == 2) buf[mine++] = ' ';
    }
    uint32_t inc = mine;  // block exclusive scan of the byte counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) wt[w] = inc;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < PT_THREADS / 32; i++) { if (i < w) base += wt[i]; total += wt[i]; }
    const uint32_t at = base + inc - mine;
    for (uint32_t k = 0; k < mine; k++) stage[at + k] = buf[k];
    __syncthreads();
    const uint64_t o0 = tile_off[tile];
    for (uint32_t k = threadIdx.x; k < total; k += PT_THREADS)
      if (o0 + k < cap) out[o0 + k] = stage[k];
    __syncthreads();
  }
}

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
  const uint8_t *tab = pretok_table(dev);
  const uint64_t n_tiles = (n + PT_TILE - 1) / PT_TILE;
  DevBuf<unsigned long long> tile_bytes(n_tiles + 1), tile_off(n_tiles + 1);
  SWB_CUDA(cudaMemsetAsync(tile_bytes.get() + n_tiles, 0, 8, stream));
  const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)sms * 8);
  pretok_count<<<grid, PT_THREADS, 0, stream>>>(d_text, n, tab, tile_bytes.get());
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, tile_bytes.get(), tile_off.get(), (int64_t)(n_tiles + 1), stream);
  DevBuf<uint8_t> tmp(tb);
  SWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tb, tile_bytes.get(), tile_off.get(), (int64_t)(n_tiles + 1), stream));
  unsigned long long total = 0;
  SWB_CUDA(cudaMemcpyAsync(&total, tile_off.get() + n_tiles, 8, cudaMemcpyDeviceToHost, stream));
  pretok_write<<<grid, PT_THREADS, 0, stream>>>(d_text, n, tab, tile_off.get(), d_out, cap);
  SWB_CUDA(cudaGetLastError());
  SWB_CUDA(cudaStreamSynchronize(stream));
  if (launches) *launches += 4;
  return total;
}

}  // namespace swb
