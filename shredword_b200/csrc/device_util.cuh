// device_util.cuh -- error handling, RAII device/pinned buffers, shared constants.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace swb {

struct Error : std::runtime_error {
  explicit Error(const std::string &m) : std::runtime_error(m) {}
};

#define SWB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      throw ::swb::Error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " +       \
                         __FILE__ + ":" + std::to_string(__LINE__) + " (" #call ")");           \
  } while (0)

// Process-wide cache of device and pinned allocations. cudaMalloc / cudaFree / cudaHostAlloc / cudaFreeHost
// synchronise the device and cost from 0.1 to 100+ ms each on a busy driver; a trainer handle needs ~25
// buffers, so a handle created after another one was destroyed reuses its blocks instead. Sizes are
// rounded up to one of four classes per power of two (<= 25 % slack). swb_release_cached_memory() gives
// everything back.
class BlockCache {
 public:
  static BlockCache &device() { static BlockCache c(false); return c; }
  static BlockCache &pinned() { static BlockCache c(true); return c; }
  static size_t size_class(size_t bytes) {
    if (bytes <= 4096) return 4096;
    size_t p = 4096;
    while (p < bytes) p <<= 1;           // p/2 < bytes <= p
    const size_t q = p >> 3;             // classes: 5/8, 6/8, 7/8, 8/8 of p
    for (size_t c = (p >> 1) + q; c < p; c += q)
      if (bytes <= c) return c;
    return p;
  }
  void *get(size_t bytes, size_t *cls_out) {
    const size_t cls = size_class(bytes);
    *cls_out = cls;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.find(cls);
      if (it != free_.end() && !it->second.empty()) {
        void *p = it->second.back();
        it->second.pop_back();
        cached_bytes_ -= cls;
        return p;
      }
    }
    void *p = nullptr;
    cudaError_t e = pinned_ ? cudaHostAlloc(&p, cls, cudaHostAllocMapped) : cudaMalloc(&p, cls);
    if (e != cudaSuccess) {  // out of memory: give the cache back and retry once
      cudaGetLastError();
      release_all();
      e = pinned_ ? cudaHostAlloc(&p, cls, cudaHostAllocMapped) : cudaMalloc(&p, cls);
    }
    if (e != cudaSuccess)
      throw Error(std::string(pinned_ ? "cudaHostAlloc" : "cudaMalloc") + " of " + std::to_string(cls) + " bytes failed: " + cudaGetErrorString(e));
    return p;
  }
  void put(void *p, size_t cls) {
    std::lock_guard<std::mutex> g(mu_);
    free_[cls].push_back(p);
    cached_bytes_ += cls;
  }
  void release_all() {
    std::lock_guard<std::mutex> g(mu_);
    for (auto &kv : free_)
      for (void *p : kv.second) { if (pinned_) cudaFreeHost(p); else cudaFree(p); }
    free_.clear();
    cached_bytes_ = 0;
  }
  size_t cached_bytes() const { return cached_bytes_; }

 private:
  explicit BlockCache(bool pinned) : pinned_(pinned) {}
  bool pinned_;
  std::mutex mu_;
  std::map<size_t, std::vector<void *>> free_;
  size_t cached_bytes_ = 0;
};

template <typename T>
class DevBuf {
 public:
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p_(o.p_), n_(o.n_), cls_(o.cls_) { o.p_ = nullptr; o.n_ = 0; o.cls_ = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p_ = o.p_; n_ = o.n_; cls_ = o.cls_; o.p_ = nullptr; o.n_ = 0; o.cls_ = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n) p_ = (T *)BlockCache::device().get(n * sizeof(T), &cls_);
    n_ = n;
  }
  void release() {
    if (p_) BlockCache::device().put(p_, cls_);
    p_ = nullptr; n_ = 0; cls_ = 0;
  }
  T *get() const { return p_; }
  size_t size() const { return n_; }
  size_t bytes() const { return n_ * sizeof(T); }

 private:
  T *p_ = nullptr;
  size_t n_ = 0, cls_ = 0;
};

// Pinned, device-mapped host memory: kernels write results straight into it, the host reads them
// without a separate D2H copy on the per-merge critical path.
template <typename T>
class PinnedBuf {
 public:
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n) {
      h_ = (T *)BlockCache::pinned().get(n * sizeof(T), &cls_);
      void *d = nullptr;
      SWB_CUDA(cudaHostGetDevicePointer(&d, h_, 0));
      d_ = (T *)d;
    }
    n_ = n;
  }
  void release() {
    if (h_) BlockCache::pinned().put(h_, cls_);
    h_ = nullptr; d_ = nullptr; n_ = 0; cls_ = 0;
  }
  T *host() const { return h_; }
  T *dev() const { return d_; }
  size_t size() const { return n_; }

 private:
  T *h_ = nullptr, *d_ = nullptr;
  size_t n_ = 0, cls_ = 0;
};

// ---- symbol stream constants (see DESIGN.md "Data layout in HBM")
constexpr int ROW = 128;                       // symbols per row; a word never straddles a row
constexpr int32_t PAD = (int32_t)0x80000000;   // filler after the last live symbol of a word / row
constexpr int32_t UNK_CODE = 0x7FFFFFFE;       // device-side stand-in for a NEGATIVE unk_id
constexpr int32_t NEG1_CODE = 0x7FFFFFFD;      // device-side stand-in for the id -1 a sign-extended delta key produces
// header of word with local index w is ~w (negative, never PAD)

__host__ __device__ __forceinline__ uint64_t dmix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ bool is_delim(uint8_t c) {  // reference bpe.cpp:247 "\t\r\n "
  return c == ' ' || c == '\n' || c == '\t' || c == '\r';
}
// first-touch key: word index major, then position, then slot (order of the 4 delta adds of a match)
// Debug build (SWB_DEBUG_BOUNDS=1 python -m shredword_b200.build): every index the merge kernels derive from data they read
// (log entries, header locations, candidate lists) is checked before use; the first violations are recorded here (code, three
// values) and the access is skipped, so that a corrupted index shows up as a report instead of an illegal address. The pool
// this was developed on does not allow compute-sanitizer.
#ifdef SWB_DEBUG_BOUNDS
__device__ unsigned long long g_dbg[1 + 4 * 16];  // [0] = number of violations, then 16 x {code, a, b, c}
#define SWB_DBG_OK(cond, code, a, b, c) swb_dbg_ok((cond), (code), (unsigned long long)(a), (unsigned long long)(b), (unsigned long long)(c))
__device__ __forceinline__ bool swb_dbg_ok(bool cond, unsigned long long code, unsigned long long a, unsigned long long b, unsigned long long c) {
  if (cond) return true;
  const unsigned long long i = atomicAdd(&g_dbg[0], 1ull);
  if (i < 16) { g_dbg[1 + 4 * i] = code; g_dbg[2 + 4 * i] = a; g_dbg[3 + 4 * i] = b; g_dbg[4 + 4 * i] = c; }
  return false;
}
#else
#define SWB_DBG_OK(cond, code, a, b, c) (true)
#endif

__host__ __device__ __forceinline__ uint64_t touch_key(uint64_t wi, uint32_t pos, uint32_t slot) {
  return (wi << 30) | ((uint64_t)pos << 2) | slot;
}

// Row signature: SIG_WORDS x 32 bits per row, one bit per symbol id present (conservative: bits are only
// ever added between rebuilds). A merge of (a, b) can only touch rows whose signature has both bits.
constexpr int SIG_WORDS = 8;
__host__ __device__ __forceinline__ uint32_t sig_hash(int32_t id) { return ((uint32_t)id * 0x9E3779B1u) >> 24; }  // 0..255

struct CastU32ToU64 {
  __host__ __device__ unsigned long long operator()(const uint32_t &x) const { return (unsigned long long)x; }
};

}  // namespace swb
