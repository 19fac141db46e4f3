// stream_map.cuh -- single-pass "byte -> 0..K bytes" map over a text, the frame of the two optional pre-passes
// (normalize.cuh, pretok.cuh): what an input byte contributes is a local function of the text around it.
//
//   * Tiles of SM_TILE bytes are handed out in text order (atomic counter). The tile and a halo of SM_HALO bytes on either
//     side are staged in shared memory by ONE bulk copy (cp.async.bulk global -> shared, completion on an mbarrier), double
//     buffered: the copy of a block's next tile runs while it works on the current one. The per-byte functions read their
//     neighbourhood through TextWin (shared memory inside the window, global memory beyond it -- long runs only).
//   * A warp owns a 512-byte segment; lane l handles bytes l, l+32, ... (conflict-free shared-memory reads). What a byte
//     contributes is first a small CODE: most bytes are settled by a cheap test (F::fast); the ones that need the general rule
//     (F::slow: long, data dependent) are listed and then dealt out evenly over the lanes, so the warp pays for the general
//     rule once per 32 such bytes, not once per group of 32 input bytes that happens to contain one. The codes are then
//     expanded to bytes, placed by a warp scan and staged in shared memory.
//   * Where the tile's output goes is learnt from the tiles before it with a decoupled look-back over one 64-bit
//     descriptor per tile (status << 62 | bytes: 1 = this tile's own count, 2 = count of all tiles up to it), then every
//     warp copies its staged bytes out with coalesced stores.
//   Traffic: the text once (+ halos) and the output once. The grid must be co-resident (blocks wait for earlier tiles).
#pragma once

#include <algorithm>

#include "device_util.cuh"

namespace swb {

constexpr int SM_THREADS = 256;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int SM_TILE = 4096;
constexpr int SM_SEG = SM_TILE / SM_WARPS;  // bytes per warp
constexpr int SM_HALO = 128;
constexpr int SM_WIN = SM_TILE + 2 * SM_HALO;
constexpr unsigned int SM_SPIN_LIMIT = 1u << 28;  // a wait that long is a bug: trap instead of hanging the GPU

// The text as the per-byte functions see it: bytes [w0, w0 + len) come from shared memory, the rest from global memory.
struct TextWin {
  const uint8_t *__restrict__ g;
  const uint8_t *s;    // s[i - w0] for w0 <= i < w0 + len
  uint64_t w0;
  uint32_t len;
  __device__ __forceinline__ uint8_t operator[](uint64_t i) const {
    const uint64_t r = i - w0;  // (wraps for i < w0: then r >= len)
    return r < (uint64_t)len ? s[(uint32_t)r] : __ldg(g + i);
  }
};
// a byte within SM_HALO of the byte being decided (which lies in the tile): always inside the staged window, no range check
__device__ __forceinline__ uint8_t tx_near(const TextWin &t, uint64_t i) { return t.s[(uint32_t)(i - t.w0)]; }

__device__ __forceinline__ uint32_t sm_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sm_mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sm_mbar_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sm_bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_smem_addr(dst)), "l"(src), "r"(bytes),
               "r"(sm_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void sm_mbar_wait(void *bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned int spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sm_smem_addr(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > SM_SPIN_LIMIT) __trap();
  }
}

// exclusive prefix of `agg` over the tiles before `tile` (called by warp 0; every lane returns the same value)
__device__ __forceinline__ unsigned long long sm_look_back(unsigned long long *__restrict__ desc, uint64_t tile, unsigned long long agg) {
  const unsigned int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) *(volatile unsigned long long *)&desc[0] = (2ull << 62) | agg;
    return 0;
  }
  if (lane == 0) *(volatile unsigned long long *)&desc[tile] = (1ull << 62) | agg;
  unsigned long long excl = 0;
  long long idx = (long long)tile - 1;
  for (;;) {
    const long long my = idx - (long long)lane;
    unsigned long long d = my >= 0 ? *(volatile unsigned long long *)&desc[my] : (2ull << 62);  // (before the first tile: an empty prefix)
    unsigned int spins = 0;
    while ((d >> 62) == 0ull) {  // that tile has not published its own count yet
      d = *(volatile unsigned long long *)&desc[my];
      if (++spins > SM_SPIN_LIMIT) __trap();
    }
    const unsigned int pm = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
    const int first = __ffs(pm) - 1;  // nearest tile (lowest lane) that already knows its inclusive prefix
    unsigned long long v = ((int)lane <= first || first < 0) ? (d & 0x3FFFFFFFFFFFFFFFull) : 0ull;
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_down_sync(0xffffffffu, v, dd);
    excl += __shfl_sync(0xffffffffu, v, 0);
    if (first >= 0) break;
    idx -= 32;
  }
  if (lane == 0) *(volatile unsigned long long *)&desc[tile] = (2ull << 62) | (excl + agg);
  return excl;
}

// F: struct with
//   static constexpr int MAX_OUT                                   most bytes one input byte can produce
//   bool     fast(const TextWin &t, n, i, uint32_t &code) const    the cheap cases: true = code is set
//   uint32_t slow(const TextWin &t, n, i) const                    the general rule -> code (< 256)
//   uint32_t expand(const TextWin &t, i, code, uint8_t *out) const writes the bytes of input byte i, returns how many
template <class F>
__global__ void __launch_bounds__(SM_THREADS)
stream_map(const uint8_t *__restrict__ text, uint64_t n, F f, unsigned long long *__restrict__ desc, unsigned int *__restrict__ tile_counter,
           uint8_t *__restrict__ out, uint64_t cap, int bulk_ok) {
  __shared__ __align__(128) uint8_t win[2][SM_WIN];
  __shared__ __align__(16) uint8_t stage[SM_WARPS][SM_SEG * F::MAX_OUT];
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ uint8_t code[SM_WARPS][SM_SEG];
  __shared__ uint16_t slow_list[SM_WARPS][SM_SEG];
  __shared__ uint32_t wtot[SM_WARPS];
  __shared__ unsigned int s_first, s_next;
  __shared__ unsigned long long s_excl;
  const uint64_t n_tiles = (n + SM_TILE - 1) / SM_TILE;
  const unsigned int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t n16 = n & ~15ull;

  // fills win[b] for `tile`: the bulk copy covers the 16-byte chunks that lie wholly inside the text, the last partial chunk
  // is fetched with plain loads by the caller after the wait. Thread 0 only.
  auto issue = [&](uint64_t tile, int b) {
    const long long w0 = (long long)(tile * SM_TILE) - SM_HALO;
    const long long a0 = w0 < 0 ? 0 : w0;
    long long a1 = w0 + SM_WIN;
    if (a1 > (long long)n16) a1 = (long long)n16;
    if (a1 > a0) {
      const uint32_t bytes = (uint32_t)(a1 - a0);
      sm_mbar_expect_tx(&bar[b], bytes);
      sm_bulk_g2s(&win[b][a0 - w0], text + a0, bytes, &bar[b]);
    } else {
      sm_mbar_expect_tx(&bar[b], 0);
    }
  };

  if (threadIdx.x == 0) {
    sm_mbar_init(&bar[0], 1);
    sm_mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_first = atomicAdd(tile_counter, 1u);
  }
  __syncthreads();
  uint64_t tile = s_first;
  if (tile >= n_tiles) return;
  if (bulk_ok && threadIdx.x == 0) issue(tile, 0);
  uint32_t phases = 0;  // bit b = parity the next wait on bar[b] uses
  int b = 0;
  for (;;) {
    // claim the tile after this one and start its copy into the other buffer (every thread is past its reads of that buffer:
    // they happened before the barrier that ended the previous iteration)
    if (threadIdx.x == 0) {
      const unsigned int nx = atomicAdd(tile_counter, 1u);
      s_next = nx;
      if (bulk_ok && nx < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(nx, b ^ 1);
      }
    }
    const long long w0 = (long long)(tile * SM_TILE) - SM_HALO;
    long long w1 = w0 + SM_WIN;
    if (w1 > (long long)n) w1 = (long long)n;
    const long long lo = w0 < 0 ? 0 : w0;
    if (bulk_ok) {
      sm_mbar_wait(&bar[b], (phases >> b) & 1u);
      phases ^= 1u << b;
      // the bytes of the last, partial 16-byte chunk of the text
      for (long long i = max(lo, (long long)n16) + threadIdx.x; i < w1; i += SM_THREADS) win[b][i - w0] = __ldg(text + i);
    } else {
      for (long long i = lo + threadIdx.x; i < w1; i += SM_THREADS) win[b][i - w0] = __ldg(text + i);
    }
    __syncthreads();
    TextWin t;
    t.g = text; t.w0 = (uint64_t)lo; t.len = (uint32_t)(w1 - lo);
    t.s = win[b] + (lo - w0);  // s[i - lo]
    // ---- this warp's segment: 16 groups of 32 consecutive bytes
    const uint64_t seg0 = tile * SM_TILE + (uint64_t)w * SM_SEG;
    uint32_t n_slow = 0;
    for (int k = 0; k < SM_SEG / 32; k++) {  // codes of the cheap cases, list of the others
      const uint64_t i = seg0 + 32ull * k + lane;
      uint32_t c = 0;
      const bool need = i < n && !f.fast(t, n, i, c);
      if (i < n && !need) code[w][32 * k + lane] = (uint8_t)c;
      const unsigned int mask = __ballot_sync(0xffffffffu, need);
      if (need) slow_list[w][n_slow + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)(32 * k + lane);
      n_slow += __popc(mask);
    }
    __syncwarp();
    for (uint32_t q = lane; q < n_slow; q += 32) {  // the general rule, every lane busy
      const uint32_t j = slow_list[w][q];
      code[w][j] = (uint8_t)f.slow(t, n, seg0 + j);
    }
    __syncwarp();
    uint32_t run = 0;  // bytes staged by this warp so far
    for (int k = 0; k < SM_SEG / 32; k++) {
      const uint64_t i = seg0 + 32ull * k + lane;
      uint8_t ob[F::MAX_OUT];
      uint32_t cnt = 0;
      if (i < n) cnt = f.expand(t, i, code[w][32 * k + lane], ob);
      uint32_t inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += u; }
      const uint32_t at = run + inc - cnt;
#pragma unroll
      for (int q = 0; q < F::MAX_OUT; q++) if ((uint32_t)q < cnt) stage[w][at + q] = ob[q];
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) wtot[w] = run;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int q = 0; q < SM_WARPS; q++) { if ((unsigned)q < w) base += wtot[q]; total += wtot[q]; }
    if (w == 0) {
      const unsigned long long excl = sm_look_back(desc, tile, total);
      if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    {  // copy out: the warp's bytes are contiguous in the output
      const uint64_t o0 = s_excl + base;
      const uint8_t *src = stage[w];
      // head up to a 4-byte boundary of the output, then 4 bytes per lane, then the tail
      const uint32_t head = min(run, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(out + o0) & 3)) & 3));
      if (lane < head && o0 + lane < cap) out[o0 + lane] = src[lane];
      const uint32_t words = (run - head) / 4;
      for (uint32_t q = lane; q < words; q += 32) {
        const uint32_t p = head + 4 * q;
        const uint32_t v = (uint32_t)src[p] | ((uint32_t)src[p + 1] << 8) | ((uint32_t)src[p + 2] << 16) | ((uint32_t)src[p + 3] << 24);
        if (o0 + p + 4 <= cap) *reinterpret_cast<uint32_t *>(out + o0 + p) = v;
        else for (int z = 0; z < 4; z++) if (o0 + p + z < cap) out[o0 + p + z] = src[p + z];
      }
      const uint32_t done = head + 4 * words;
      if (lane < run - done && o0 + done + lane < cap) out[o0 + done + lane] = src[done + lane];
    }
    const uint64_t nx = s_next;
    __syncthreads();  // (stage, wtot, s_next and win[b] are free again)
    if (nx >= n_tiles) return;
    tile = nx;
    b ^= 1;
  }
}

// Runs F over text[0, n) on `stream`; returns the length of the full result (may exceed cap: then only the first cap bytes
// were written). One kernel launch + the 8-byte total read back.
template <class F>
inline uint64_t stream_map_run(const uint8_t *d_text, uint64_t n, const F &f, uint8_t *d_out, uint64_t cap, cudaStream_t stream, int sms, uint64_t *launches) {
  if (n == 0) return 0;
  const uint64_t n_tiles = (n + SM_TILE - 1) / SM_TILE;
  DevBuf<unsigned long long> desc(n_tiles + 2);
  SWB_CUDA(cudaMemsetAsync(desc.get(), 0, (n_tiles + 2) * 8, stream));
  unsigned int *counter = reinterpret_cast<unsigned int *>(desc.get() + n_tiles);
  int per_sm = 0;
  SWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream_map<F>, SM_THREADS, 0));
  if (per_sm < 1) throw Error("stream_map: the kernel does not fit an SM");
  const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)sms * (uint64_t)per_sm);  // co-resident: blocks wait for earlier tiles
  const int bulk_ok = (reinterpret_cast<uintptr_t>(d_text) & 15u) == 0 ? 1 : 0;        // cp.async.bulk needs 16-byte aligned addresses
  stream_map<F><<<grid, SM_THREADS, 0, stream>>>(d_text, n, f, desc.get(), counter, d_out, cap, bulk_ok);
  SWB_CUDA(cudaGetLastError());
  unsigned long long last = 0;
  SWB_CUDA(cudaMemcpyAsync(&last, desc.get() + (n_tiles - 1), 8, cudaMemcpyDeviceToHost, stream));
  SWB_CUDA(cudaStreamSynchronize(stream));
  if (launches) *launches += 1;
  if ((last >> 62) != 2ull) throw Error("stream_map: the last tile never published its prefix (internal error)");
  return last & 0x3FFFFFFFFFFFFFFFull;
}

}  // namespace swb
