// grid_probe.cu -- development probe: how long does a grid of short blocks take from first start to last done?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int SMEM>
__global__ void __launch_bounds__(256) probe(unsigned int *done, unsigned long long *stamps, const uint4 *data, size_t n16, int work, unsigned int *sink) {
  __shared__ char pad[SMEM];
  __shared__ bool last;
  if (blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = gtime();
  if (SMEM > 64 && threadIdx.x == 0) pad[SMEM - 1] = 1;
  unsigned int acc = 0;
  if (work) {  // each warp reads `work` rounds of 512 B, dependent rounds (like sig -> rows -> rows)
    size_t idx = ((size_t)blockIdx.x * 256 + threadIdx.x);
    for (int r = 0; r < work; r++) {
      uint4 v = data[idx % n16];
      acc += v.x + v.y + v.z + v.w;
      idx = idx * 2654435761ull + (acc & 1) + 12345;
    }
  }
  if (acc == 0xdeadbeef) *sink = acc;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) { stamps[1] = gtime(); *done = 0; }
}

template <int SMEM>
void run(const char *name, int grid, int work, unsigned int *done, unsigned long long *stamps, const uint4 *data, size_t n16, unsigned int *sink) {
  std::vector<double> v;
  unsigned long long h[2];
  for (int i = 0; i < 300; i++) {
    probe<SMEM><<<grid, 256>>>(done, stamps, data, n16, work, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, stamps, 16, cudaMemcpyDeviceToHost));
    if (i >= 50) v.push_back((double)(h[1] - h[0]) * 1e-3);
  }
  std::sort(v.begin(), v.end());
  printf("%-28s grid=%4d work=%d smem=%5d : first-start -> last-done p50 %.2f us p90 %.2f us\n", name, grid, work, SMEM, v[v.size() / 2], v[v.size() * 9 / 10]);
}

int main() {
  unsigned int *done, *sink; unsigned long long *stamps; uint4 *data;
  const size_t n16 = (48ull << 20) / 16;
  CK(cudaMalloc(&done, 4)); CK(cudaMemset(done, 0, 4)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&stamps, 16)); CK(cudaMalloc(&data, n16 * 16)); CK(cudaMemset(data, 1, n16 * 16));
  for (int grid : {148, 296, 356, 592}) {
    run<64>("empty blocks", grid, 0, done, stamps, data, n16, sink);
    run<64>("3 dependent loads", grid, 3, done, stamps, data, n16, sink);
    run<42000>("3 loads, 41 KB smem", grid, 3, done, stamps, data, n16, sink);
  }
  return 0;
}
