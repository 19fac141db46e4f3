// word_table.cuh -- corpus bytes -> unique-word table -> row-packed symbol stream (all on device).
//
// Replaces reference bpe_load_corpus' tokeniser + StrMap + Symbol chains
// (reference csrc/bpe/bpe.cpp:229-292, hash.cpp:29-53, histogram.cpp:7-36).
//
//   wt_tokenize      : 16 bytes per thread, delimiter bit-mask by SIMD byte compares; every word start
//                      is hashed and counted in a block-private shared-memory table which is flushed
//                      into the global open-addressing table (one global atomic per distinct word per
//                      flush instead of one per occurrence: the Zipf head would serialise otherwise).
//   wt_compact       : occupied slots -> (sort key = djb2 bucket << 40 | first offset, count)
//   (radix sort)     : reference word order = (bucket ascending, first occurrence ascending)
//   wt_word_info     : per unique word: length, byte histogram (unweighted, as the reference)
//   wt_pack_*        : greedy packing of words into 128-symbol rows (a word never straddles a row)
#pragma once

#include "device_util.cuh"

namespace swb {

constexpr uint64_t WT_EMPTY = ~0ull;
constexpr int WT_TAG_BITS = 24;
constexpr uint64_t WT_TAG_MASK = (1ull << WT_TAG_BITS) - 1;
constexpr int LT_SLOTS = 4096;   // block-private table (48 KB: 8 B key + 4 B count per slot)
constexpr int LT_PROBES = 16;
constexpr int WT_THREADS = 256;
constexpr int WT_SMEM_BYTES = LT_SLOTS * 12 + 16;

struct WordTableDev {
  unsigned long long *keys;    // (first offset seen << 24) | tag24 ; tag low 12 bits = djb2 & 4095
  unsigned long long *counts;
  uint64_t mask;               // capacity - 1
  unsigned int *n_unique;
  unsigned int *flags;         // bit 0: table overflow, bit 1: NUL byte seen
  uint64_t limit;              // max unique words before overflow is declared
};

// Scans the word starting at `off`: returns its length, fills both hashes.
__device__ __forceinline__ uint32_t wt_scan_word(const uint8_t *__restrict__ p, uint64_t off, uint64_t n,
                                                uint64_t &h64, uint32_t &djb) {
  uint64_t h = 0xcbf29ce484222325ull;
  uint32_t d = 5381u;  // reference hash.cpp:35-38 (only the low 12 bits are ever used)
  uint64_t i = off;
  while (i < n) {
    const uint8_t c = p[i];
    if (is_delim(c)) break;
    h = (h ^ c) * 0x100000001b3ull;
    d = d * 33u + c;
    ++i;
  }
  h64 = dmix64(h);
  djb = d;
  return (uint32_t)(i - off);
}

// true iff the word at offB is byte-identical to the word [offA, offA+len)
__device__ __forceinline__ bool wt_same_word(const uint8_t *__restrict__ p, uint64_t n, uint64_t offA, uint64_t offB,
                                            uint32_t len) {
  if (offA == offB) return true;
  if (offB + len > n) return false;
  for (uint32_t i = 0; i < len; i++)
    if (p[offA + i] != p[offB + i]) return false;
  return offB + len == n || is_delim(p[offB + len]);
}

__device__ __forceinline__ void wt_global_insert(const WordTableDev &t, const uint8_t *__restrict__ p, uint64_t n,
                                                 uint64_t off, uint32_t len, uint64_t h64, uint64_t tag,
                                                 unsigned long long cnt) {
  const unsigned long long key = (off << WT_TAG_BITS) | tag;
  uint64_t slot = h64 & t.mask;
  for (uint64_t probe = 0; probe <= t.mask; probe++) {
    unsigned long long cur = t.keys[slot];
    if (cur == WT_EMPTY) {
      const unsigned long long prev = atomicCAS(&t.keys[slot], WT_EMPTY, key);
      if (prev == WT_EMPTY) {
        atomicAdd(&t.counts[slot], cnt);
        const unsigned int u = atomicAdd(t.n_unique, 1u);
        if (u >= t.limit) atomicOr(t.flags, 1u);
        return;
      }
      cur = prev;
    }
    if ((cur & WT_TAG_MASK) == tag && wt_same_word(p, n, off, cur >> WT_TAG_BITS, len)) {
      atomicAdd(&t.counts[slot], cnt);
      if (key < cur) atomicMin(&t.keys[slot], key);  // same tag => orders by offset: keeps the first occurrence
      return;
    }
    slot = (slot + 1) & t.mask;
  }
  atomicOr(t.flags, 1u);
}

__device__ __forceinline__ uint32_t wt_delim_bits(uint32_t w, uint32_t &nul) {
  const uint32_t m = __vcmpeq4(w, 0x20202020u) | __vcmpeq4(w, 0x0a0a0a0au) | __vcmpeq4(w, 0x09090909u) |
                     __vcmpeq4(w, 0x0d0d0d0du);
  nul |= __vcmpeq4(w, 0u);
  return ((m & 0x01010101u) * 0x01020408u) >> 24;  // one bit per byte, byte 0 -> bit 0
}

// corpus must be 16-byte aligned and padded with >= 16 delimiter bytes after n.
// Counts the words that START in [lo, hi) (the whole corpus: lo = 0, hi = n). A range lets the host-buffer load
// tokenise a piece of the corpus while the next piece is still crossing PCIe: `hi` is then cut right after a
// delimiter, so every word that starts before it also ends before it, and the bytes of the 16-byte segments the
// range touches have all arrived (the copy granularity is a multiple of 16).
__global__ void __launch_bounds__(WT_THREADS)
wt_tokenize(const uint8_t *__restrict__ corpus, uint64_t n, WordTableDev tbl, uint64_t lo, uint64_t hi) {
  extern __shared__ __align__(16) unsigned long long wt_dyn_smem[];  // WT_SMEM_BYTES, opt-in above 48 KB
  unsigned long long *lkeys = wt_dyn_smem;
  unsigned int *lcnt = reinterpret_cast<unsigned int *>(wt_dyn_smem + LT_SLOTS);
  unsigned int &lused = lcnt[LT_SLOTS];
  for (int i = threadIdx.x; i < LT_SLOTS; i += WT_THREADS) { lkeys[i] = WT_EMPTY; lcnt[i] = 0; }
  if (threadIdx.x == 0) lused = 0;
  __syncthreads();

  const uint64_t seg_lo = lo / 16, seg_hi = (hi + 15) / 16, nseg = seg_hi - seg_lo;
  // each block owns a contiguous span of segments so that its private table sees a long stretch of text
  const uint64_t per_block = ((nseg + gridDim.x - 1) / gridDim.x + WT_THREADS - 1) / WT_THREADS * WT_THREADS;
  const uint64_t seg_begin = seg_lo + (uint64_t)blockIdx.x * per_block;
  const uint64_t seg_end = min(seg_hi, seg_begin + per_block);
  uint32_t nul = 0;

  for (uint64_t base = seg_begin; base < seg_end; base += WT_THREADS) {
    const uint64_t seg = base + threadIdx.x;
    if (seg < seg_end) {
      const uint4 v = *reinterpret_cast<const uint4 *>(corpus + seg * 16);
      uint32_t dm = wt_delim_bits(v.x, nul) | (wt_delim_bits(v.y, nul) << 4) | (wt_delim_bits(v.z, nul) << 8) |
                    (wt_delim_bits(v.w, nul) << 12);
      const uint32_t prev_delim = (seg == 0) ? 1u : (is_delim(corpus[seg * 16 - 1]) ? 1u : 0u);
      uint32_t starts = ~dm & ((dm << 1) | prev_delim) & 0xFFFFu;
      while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        const uint64_t off = seg * 16 + s;
        if (off >= hi) break;
        if (off < lo) continue;
        uint64_t h64; uint32_t djb;
        const uint32_t len = wt_scan_word(corpus, off, n, h64, djb);
        const uint64_t tag = (djb & 0xFFFu) | (((h64 >> 40) & 0xFFFu) << 12);
        const unsigned long long key = (off << WT_TAG_BITS) | tag;
        // block-private table first
        uint32_t slot = (uint32_t)h64 & (LT_SLOTS - 1);
        bool done = false;
        for (int probe = 0; probe < LT_PROBES && !done; probe++) {
          unsigned long long cur = lkeys[slot];
          if (cur == WT_EMPTY) {
            const unsigned long long prev = atomicCAS(&lkeys[slot], WT_EMPTY, key);
            if (prev == WT_EMPTY) { atomicAdd(&lcnt[slot], 1u); atomicAdd(&lused, 1u); done = true; break; }
            cur = prev;
          }
          if ((cur & WT_TAG_MASK) == tag && wt_same_word(corpus, n, off, cur >> WT_TAG_BITS, len)) {
            atomicAdd(&lcnt[slot], 1u);
            if (key < cur) atomicMin(&lkeys[slot], key);
            done = true;
            break;
          }
          slot = (slot + 1) & (LT_SLOTS - 1);
        }
        if (!done) wt_global_insert(tbl, corpus, n, off, len, h64, tag, 1ull);
      }
    }
    __syncthreads();
    const bool last = base + WT_THREADS >= seg_end;
    // block-uniform decision: everybody reads `lused` before anybody starts the next tile's inserts
    if (__syncthreads_or(lused > LT_SLOTS / 2 || last)) {
      for (int i = threadIdx.x; i < LT_SLOTS; i += WT_THREADS) {
        const unsigned long long k = lkeys[i];
        if (k != WT_EMPTY) {
          const uint64_t off = k >> WT_TAG_BITS;
          uint64_t h64; uint32_t djb;
          const uint32_t len = wt_scan_word(corpus, off, n, h64, djb);
          wt_global_insert(tbl, corpus, n, off, len, h64, k & WT_TAG_MASK, (unsigned long long)lcnt[i]);
          lkeys[i] = WT_EMPTY; lcnt[i] = 0;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) lused = 0;
      __syncthreads();
    }
  }
  if (nul) atomicOr(tbl.flags, 2u);
}

__global__ void wt_fill(unsigned long long *keys, unsigned long long *counts, uint64_t cap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    keys[i] = WT_EMPTY; counts[i] = 0;
  }
}

// occupied slots -> (bucket << 40 | first offset, count); order is fixed by the sort that follows
__global__ void wt_compact(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ counts,
                           uint64_t cap, unsigned long long *__restrict__ sort_keys,
                           unsigned long long *__restrict__ out_counts, unsigned int *cursor) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    if (k != WT_EMPTY) {
      const unsigned int j = atomicAdd(cursor, 1u);
      sort_keys[j] = ((k & 0xFFFull) << 40) | (k >> WT_TAG_BITS);
      out_counts[j] = counts[i];
    }
  }
}

// per unique word (reference order): offset, length, unweighted byte histogram (reference
// histogram.cpp:30-36 counts every byte of every UNIQUE word once); words longer than ROW-1
// symbols are registered in the long-word list.
__global__ void __launch_bounds__(256)
wt_word_info(const uint8_t *__restrict__ corpus, uint64_t n, const uint64_t *__restrict__ woff,
             uint64_t W, uint32_t *__restrict__ wlen,
             unsigned long long *__restrict__ hist256, unsigned int *n_long, unsigned long long *long_syms,
             uint32_t *__restrict__ long_index /* [W]: index into the long list or ~0 */,
             unsigned long long *__restrict__ byte_total) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned long long my_bytes = 0;
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < W; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t off = woff[w];
    uint64_t i = off;
    while (i < n) {
      const uint8_t c = corpus[i];
      if (is_delim(c)) break;
      atomicAdd(&sh[c], 1u);
      ++i;
    }
    const uint32_t len = (uint32_t)(i - off);
    wlen[w] = len;
    my_bytes += len;
    if (len > ROW - 1) {
      const unsigned int j = atomicAdd(n_long, 1u);
      long_index[w] = j;
      atomicAdd(long_syms, (unsigned long long)len);
    } else {
      long_index[w] = 0xFFFFFFFFu;
    }
  }
  atomicAdd(byte_total, my_bytes);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist256[i], (unsigned long long)sh[i]);
}

// sorted (bucket << 40 | first offset) keys -> byte offsets of the words, in reference order
__global__ void wt_unpack_sorted(const unsigned long long *__restrict__ sorted_keys, uint64_t W, uint64_t *__restrict__ woff) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < W; w += (uint64_t)gridDim.x * blockDim.x)
    woff[w] = sorted_keys[w] & ((1ull << 40) - 1);
}

// ---- range-split load (multi-GPU): every rank tokenises its own byte range, the per-rank unique-word
// tables are all-gathered (word bytes + count + GLOBAL first offset) and merged on every rank.
struct __align__(16) WordMeta { unsigned long long aoff, count, first; unsigned int len, pad; };  // 32 bytes

__global__ void wt_local_lens(const uint8_t *__restrict__ corpus, uint64_t n, const unsigned long long *__restrict__ skeys,
                              uint64_t Wl, unsigned long long *__restrict__ len1 /* len + 1 (separator) */) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < Wl; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t off = skeys[w] & ((1ull << 40) - 1);
    uint64_t i = off;
    while (i < n && !is_delim(corpus[i])) ++i;
    len1[w] = (i - off) + 1;
  }
}
// local unique word j -> its bytes + ' ' at arena[aoff[j]] and its meta record
__global__ void wt_export(const uint8_t *__restrict__ corpus, const unsigned long long *__restrict__ skeys,
                          const unsigned long long *__restrict__ scnt, const unsigned long long *__restrict__ len1,
                          const unsigned long long *__restrict__ aoff, uint64_t Wl, uint64_t global_offset,
                          uint8_t *__restrict__ arena, WordMeta *__restrict__ meta) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < Wl; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t off = skeys[w] & ((1ull << 40) - 1);
    const uint32_t len = (uint32_t)(len1[w] - 1);
    uint8_t *dst = arena + aoff[w];
    for (uint32_t k = 0; k < len; k++) dst[k] = corpus[off + k];
    dst[len] = ' ';
    WordMeta m;
    m.aoff = aoff[w]; m.count = scnt[w]; m.first = global_offset + off; m.len = len; m.pad = 0;
    meta[w] = m;
  }
}
// every rank's exported words -> one global table. The key only references bytes (claimed once); the
// global first offset is min-reduced in `gfirst`, the counts are summed.
__global__ void __launch_bounds__(256)
wt_insert_words(const uint8_t *__restrict__ base, uint64_t base_n, const WordMeta *__restrict__ meta_all,
                const unsigned long long *__restrict__ sizes /* [2*R]: Wl, arena bytes */, int R, uint64_t maxW, uint64_t maxA,
                WordTableDev tbl, unsigned long long *__restrict__ gfirst) {
  for (int r = 0; r < R; r++) {
    const uint64_t Wl = sizes[2 * r];
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < Wl; j += (uint64_t)gridDim.x * blockDim.x) {
      const WordMeta m = meta_all[(uint64_t)r * maxW + j];
      const uint64_t off = (uint64_t)r * maxA + m.aoff;
      uint64_t h64; uint32_t djb;
      const uint32_t len = wt_scan_word(base, off, base_n, h64, djb);
      const uint64_t tag = (djb & 0xFFFu) | (((h64 >> 40) & 0xFFFu) << 12);
      const unsigned long long key = (off << WT_TAG_BITS) | tag;
      uint64_t slot = h64 & tbl.mask;
      for (uint64_t probe = 0; probe <= tbl.mask; probe++) {
        unsigned long long cur = tbl.keys[slot];
        if (cur == WT_EMPTY) {
          const unsigned long long prev = atomicCAS(&tbl.keys[slot], WT_EMPTY, key);
          if (prev == WT_EMPTY) {
            if (atomicAdd(tbl.n_unique, 1u) >= tbl.limit) atomicOr(tbl.flags, 1u);
            cur = key;
          } else cur = prev;
        }
        if (cur == key || ((cur & WT_TAG_MASK) == tag && wt_same_word(base, base_n, off, cur >> WT_TAG_BITS, len))) {
          atomicAdd(&tbl.counts[slot], m.count);
          atomicMin(&gfirst[slot], m.first);
          break;
        }
        slot = (slot + 1) & tbl.mask;
      }
    }
  }
}
__global__ void wt_fill3(unsigned long long *keys, unsigned long long *counts, unsigned long long *gfirst, uint64_t cap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    keys[i] = WT_EMPTY; counts[i] = 0; gfirst[i] = ~0ull;
  }
}
__global__ void wt_compact_dist(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ gfirst,
                                uint64_t cap, unsigned long long *__restrict__ sort_keys, unsigned long long *__restrict__ slots,
                                unsigned int *cursor) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    if (k != WT_EMPTY) {
      const unsigned int j = atomicAdd(cursor, 1u);
      sort_keys[j] = ((k & 0xFFFull) << 40) | gfirst[i];
      slots[j] = i;
    }
  }
}
__global__ void wt_after_sort_dist(const unsigned long long *__restrict__ sorted_slots, const unsigned long long *__restrict__ keys,
                                   const unsigned long long *__restrict__ counts, uint64_t W, uint64_t *__restrict__ woff,
                                   unsigned long long *__restrict__ cnt) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < W; w += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long s = sorted_slots[w];
    woff[w] = keys[s] >> WT_TAG_BITS;
    cnt[w] = counts[s];
  }
}

// ---- row packing. A batch of PACK_BATCH consecutive words (one warp) starts on a fresh row and is
// packed greedily: a row is closed when the next word does not fit (or after 32 words).
constexpr int PACK_BATCH = 1024;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// need(w) = len+1 for words stored in this rank's rows; 0 for long words and for other ranks' words
// (word w belongs to rank w % nranks). The header of word w is ~w (global, reference-order index).
template <bool WRITE>
__global__ void __launch_bounds__(256)
wt_pack(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ woff, const uint32_t *__restrict__ wlen,
        const uint32_t *__restrict__ long_index /* [W] index into the long list or ~0 */, uint64_t W, int rank,
        int nranks, const int32_t *__restrict__ byte_map /* [256] */,
        uint32_t *__restrict__ batch_rows /* [n_batches] in: exclusive scan when WRITE; out: counts otherwise */,
        int4 *__restrict__ rows, uint64_t *__restrict__ wloc /* [W] row*ROW+pos of the header */,
        uint32_t *__restrict__ sig /* [n_rows * SIG_WORDS] */) {
  __shared__ __align__(16) int srow[8][ROW];
  __shared__ uint32_t ssig[8][SIG_WORDS];
  __shared__ int32_t bmap[256];
  if (WRITE) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) bmap[i] = byte_map[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t n_batches = (W + PACK_BATCH - 1) / PACK_BATCH;
  for (uint64_t batch = blockIdx.x * 8ull + wib; batch < n_batches; batch += gridDim.x * 8ull) {
    const uint64_t w0 = batch * PACK_BATCH, w1 = min(W, w0 + PACK_BATCH);
    uint64_t row = WRITE ? batch_rows[batch] : 0;
    uint32_t nrows = 0;
    uint64_t w = w0;
    while (w < w1) {
      const uint64_t mine = w + lane;
      uint32_t need = 0, len = 0;
      if (mine < w1 && (int)(mine % (uint64_t)nranks) == rank && long_index[mine] == 0xFFFFFFFFu) {
        len = wlen[mine];
        need = len + 1;
      }
      const uint32_t ps = warp_incl_scan(need, lane);
      const bool fits = (mine < w1) && ps <= ROW;
      const uint32_t fit_mask = __ballot_sync(0xffffffffu, fits);
      const int nfit = __popc(fit_mask);  // a prefix of the lanes; >= 1 because need <= ROW
      const uint32_t used = __shfl_sync(0xffffffffu, ps, nfit - 1);
      if (used > 0) {
        if (WRITE) {
          for (int i = lane; i < ROW; i += 32) srow[wib][i] = PAD;
          if (lane < SIG_WORDS) ssig[wib][lane] = 0;
          __syncwarp();
          if (lane < nfit && need) {
            const uint32_t pos = ps - need;
            srow[wib][pos] = ~(int32_t)(uint32_t)mine;
            const uint8_t *src = corpus + woff[mine];
            for (uint32_t k = 0; k < len; k++) {
              const int32_t id = bmap[src[k]];
              srow[wib][pos + 1 + k] = id;
              const uint32_t h = sig_hash(id);
              atomicOr(&ssig[wib][h >> 5], 1u << (h & 31));
            }
            wloc[mine] = row * ROW + pos;
          }
          __syncwarp();
          rows[row * (ROW / 4) + lane] = *reinterpret_cast<const int4 *>(&srow[wib][lane * 4]);
          if (lane < SIG_WORDS) sig[row * SIG_WORDS + lane] = ssig[wib][lane];
          __syncwarp();
        }
        row++; nrows++;
      }
      w += nfit;
    }
    if (!WRITE && lane == 0) batch_rows[batch] = nrows;
  }
}

// long words: CSR symbols (word bytes mapped through byte_map); one thread per long word.
// Long words of other ranks get length 0 here.
__global__ void wt_fill_long(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ woff,
                             const uint32_t *__restrict__ wlen, const uint32_t *__restrict__ long_index, uint64_t W,
                             int rank, int nranks, const int32_t *__restrict__ byte_map,
                             const uint64_t *__restrict__ long_off, int32_t *__restrict__ long_syms,
                             uint32_t *__restrict__ long_len, uint32_t *__restrict__ long_word) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < W; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t j = long_index[w];
    if (j == 0xFFFFFFFFu) continue;
    long_word[j] = (uint32_t)w;
    if ((int)(w % (uint64_t)nranks) != rank) { long_len[j] = 0; continue; }
    const uint8_t *src = corpus + woff[w];
    int32_t *dst = long_syms + long_off[j];
    const uint32_t len = wlen[w];
    for (uint32_t k = 0; k < len; k++) dst[k] = byte_map[src[k]];
    long_len[j] = len;
  }
}

// long list offsets: long_off[j] = exclusive position of long word j (any order; claimed atomically)
__global__ void wt_long_offsets(const uint32_t *__restrict__ wlen, const uint32_t *__restrict__ long_index, uint64_t W,
                                unsigned long long *cursor, uint64_t *__restrict__ long_off) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < W; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t j = long_index[w];
    if (j != 0xFFFFFFFFu) long_off[j] = atomicAdd(cursor, (unsigned long long)wlen[w]);
  }
}

// gathers the bytes of the unique words into one arena (so the corpus buffer can be released)
__global__ void wt_gather_bytes(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ woff,
                                const uint32_t *__restrict__ wlen, const uint64_t *__restrict__ boff, uint64_t W,
                                uint8_t *__restrict__ arena) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < W; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t *src = corpus + woff[w];
    uint8_t *dst = arena + boff[w];
    const uint32_t len = wlen[w];
    for (uint32_t k = 0; k < len; k++) dst[k] = src[k];
  }
}

}  // namespace swb
