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
constexpr int LT_SLOTS = 4096;   // block-private table: tag 8 B + word bytes 16 B + count 4 B + first offset 4 B per slot = 128 KB
constexpr int LT_PROBES = 4;
constexpr int WT_THREADS = 1024; // one block per SM
constexpr int WT_SMEM_BYTES = LT_SLOTS * 32 + 16;
constexpr int WT_SHORT_MAX = 15; // words of up to 15 bytes are identified by their bytes (two 64-bit words, zero padded)

// One slot of the global word table, one 32-byte sector:
//   key   = (offset of the first occurrence seen << 24) | tag24 ; tag low 12 bits = djb2 & 4095 (the reference's StrMap bucket)
//   count = occurrences
//   lo,hi = the word's bytes, zero padded, for words of up to WT_SHORT_MAX bytes (a corpus holds no NUL byte, so the padded
//           bytes identify the word); written once by whoever claims the slot, as ONE 16-byte store, read as one 16-byte
//           load: (0, 0) = not there (yet) -- then, and for longer words, the bytes are compared through the offset in `key`.
struct __align__(32) WSlot { unsigned long long key, count, lo, hi; };
struct WordTableDev {
  WSlot *slots;
  uint64_t mask;               // capacity - 1
  unsigned int *n_unique;
  unsigned int *flags;         // bit 0: table overflow, bit 1: NUL byte seen
  uint64_t limit;              // max unique words before overflow is declared
};
// the table the ranks' exported words are merged into (multi-GPU load): keys / counts / global first offsets side by side
struct WordTableSoA {
  unsigned long long *keys;
  unsigned long long *counts;
  uint64_t mask;
  unsigned int *n_unique;
  unsigned int *flags;
  uint64_t limit;
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

// one 16-byte access to shared memory (the word bytes of a block-private slot are written and read as a unit)
__device__ __forceinline__ ulonglong2 wt_lds16(const ulonglong2 *p) {
  ulonglong2 r;
  asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "r"((unsigned int)__cvta_generic_to_shared(p)) : "memory");
  return r;
}
__device__ __forceinline__ void wt_sts16(ulonglong2 *p, unsigned long long x, unsigned long long y) {
  asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"((unsigned int)__cvta_generic_to_shared(p)), "l"(x), "l"(y) : "memory");
}
// hash of a short word's padded bytes (never 0: 0 marks a free slot of the block-private table)
__device__ __forceinline__ uint64_t wt_short_hash(uint64_t lo, uint64_t hi) {
  const uint64_t h = dmix64(lo ^ dmix64(hi + 0x632BE59BD9B4E019ull));
  return h ? h : 0x9E3779B97F4A7C15ull;
}
// length and djb2 (reference hash.cpp:35-38) of a short word from its padded bytes
__device__ __forceinline__ uint32_t wt_short_len_djb(uint64_t lo, uint64_t hi, uint32_t &djb) {
  uint32_t d = 5381u, len = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const uint32_t c = (uint32_t)((k < 8 ? lo >> (8 * k) : hi >> (8 * (k - 8))) & 0xFFu);
    if (c) { d = d * 33u + c; len = k + 1; }  // (bytes are contiguous from 0: a zero byte ends the word)
  }
  djb = d;
  return len;
}

// Adds `cnt` occurrences of the word at `off` to the global table. short_word: (lo, hi) are its padded bytes.
__device__ __forceinline__ void wt_global_insert(const WordTableDev &t, const uint8_t *__restrict__ p, uint64_t n,
                                                 uint64_t off, uint32_t len, uint64_t h64, uint64_t tag,
                                                 unsigned long long cnt, bool short_word, uint64_t lo, uint64_t hi) {
  const unsigned long long key = (off << WT_TAG_BITS) | tag;
  uint64_t slot = h64 & t.mask;
  for (uint64_t probe = 0; probe <= t.mask; probe++) {
    // Once the table has been declared too small the host throws this run away and tokenises again into a larger one: every
    // further insert would only walk ever longer probe sequences through a table that is filling up (a 32 MB corpus with 350 k
    // unique words spent 650 ms that way in a table of 131 072 slots). Looked at only on long probe sequences: the flag is
    // one word that every SM would otherwise read on every insert.
    if ((probe & 31u) == 31u && (__ldcg(t.flags) & 1u)) return;
    WSlot *sl = t.slots + slot;
    unsigned long long cur = __ldcg(&sl->key);
    if (cur == WT_EMPTY) {
      const unsigned long long prev = atomicCAS(&sl->key, WT_EMPTY, key);
      if (prev == WT_EMPTY) {
        if (short_word) *reinterpret_cast<ulonglong2 *>(&sl->lo) = make_ulonglong2(lo, hi);  // one 16-byte store
        atomicAdd(&sl->count, cnt);
        const unsigned int u = atomicAdd(t.n_unique, 1u);
        if (u >= t.limit) atomicOr(t.flags, 1u);
        return;
      }
      cur = prev;
    }
    if ((cur & WT_TAG_MASK) == tag) {
      bool same;
      const ulonglong2 pay = short_word ? __ldcg(reinterpret_cast<const ulonglong2 *>(&sl->lo)) : make_ulonglong2(0ull, 0ull);
      if (pay.x != 0ull) same = pay.x == lo && pay.y == hi;          // both short: the bytes themselves
      else same = wt_same_word(p, n, off, cur >> WT_TAG_BITS, len);   // a long word, or bytes not published yet: through the corpus
      if (same) {
        atomicAdd(&sl->count, cnt);
        if (key < cur) atomicMin(&sl->key, key);  // same tag => orders by offset: keeps the first occurrence
        return;
      }
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

// corpus must be 16-byte aligned and padded with >= 48 delimiter bytes after n.
// Counts the words that START in [lo, hi) (the whole corpus: lo = 0, hi = n). A range lets the host-buffer load
// tokenise a piece of the corpus while the next piece is still crossing PCIe: `hi` is then cut right after a
// delimiter, so every word that starts before it also ends before it, and the bytes of the 16-byte segments the
// range touches have all arrived (the copy granularity is a multiple of 16).
//
// One block per SM, each over one contiguous span of the range. A thread takes 16 bytes (one streaming load), gets the
// next 16 from its neighbour lane, and handles the words that start in its segment entirely in registers: the 16 bytes
// from the word start are cut out of that 32-byte window with funnel shifts, the delimiter mask gives the length, and a
// word of up to 15 bytes IS its key (two zero-padded 64-bit words; no byte loop, no compare against an earlier occurrence).
// Occurrences are summed in a block-private table in shared memory that keeps the first LT_SLOTS distinct words the
// block meets -- in Zipfian text those contain the head of the distribution -- and is flushed once, at the end of the span;
// words that find no place there, and words of 16+ bytes, go to the global table directly (one 32-byte sector per probe).
__global__ void __launch_bounds__(WT_THREADS, 1)
wt_tokenize(const uint8_t *__restrict__ corpus, uint64_t n, WordTableDev tbl, uint64_t lo_b, uint64_t hi_b) {
  extern __shared__ __align__(16) unsigned long long wt_dyn_smem[];  // WT_SMEM_BYTES, opt-in above 48 KB
  ulonglong2 *lpay = reinterpret_cast<ulonglong2 *>(wt_dyn_smem);                 // [LT_SLOTS] word bytes
  unsigned long long *ltag = wt_dyn_smem + 2 * LT_SLOTS;                          // [LT_SLOTS] hash, 0 = free
  unsigned int *lcnt = reinterpret_cast<unsigned int *>(wt_dyn_smem + 3 * LT_SLOTS);  // [LT_SLOTS]
  unsigned int *lmin = lcnt + LT_SLOTS;                                           // [LT_SLOTS] first offset, relative to the span
  for (int i = threadIdx.x; i < LT_SLOTS; i += WT_THREADS) { lpay[i] = make_ulonglong2(0ull, 0ull); ltag[i] = 0ull; lcnt[i] = 0u; lmin[i] = 0xFFFFFFFFu; }
  __syncthreads();

  const uint64_t seg_lo = lo_b / 16, seg_hi = (hi_b + 15) / 16, nseg = seg_hi - seg_lo;
  const uint64_t per_block = ((nseg + gridDim.x - 1) / gridDim.x + WT_THREADS - 1) / WT_THREADS * WT_THREADS;
  const uint64_t seg_begin = seg_lo + (uint64_t)blockIdx.x * per_block;
  const uint64_t seg_end = min(seg_hi, seg_begin + per_block);
  const uint64_t span0 = seg_begin * 16;  // (spans are < 4 GB: the host sizes the grid accordingly)
  const int lane = threadIdx.x & 31;
  uint32_t nul = 0;

  for (uint64_t base = seg_begin; base < seg_end; base += WT_THREADS) {
    const uint64_t seg = base + threadIdx.x;
    const bool live = seg < seg_end;
    uint4 v = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
    if (live) v = __ldcs(reinterpret_cast<const uint4 *>(corpus + seg * 16));
    // the 16 bytes after this segment: the neighbour lane's, or (last lane of the warp) one more load
    uint4 nx;
    nx.x = __shfl_down_sync(0xffffffffu, v.x, 1); nx.y = __shfl_down_sync(0xffffffffu, v.y, 1);
    nx.z = __shfl_down_sync(0xffffffffu, v.z, 1); nx.w = __shfl_down_sync(0xffffffffu, v.w, 1);
    uint32_t prev_last = __shfl_up_sync(0xffffffffu, v.w, 1) >> 24;
    if (!live) continue;  // (whole warps leave together except in the last tile, where the shuffles above have already happened)
    if (lane == 31 || seg + 1 >= seg_end) nx = __ldg(reinterpret_cast<const uint4 *>(corpus + seg * 16 + 16));  // (padding makes this safe)
    if (lane == 0) prev_last = seg == 0 ? 0x20u : (uint32_t)corpus[seg * 16 - 1];
    const uint32_t dm = wt_delim_bits(v.x, nul) | (wt_delim_bits(v.y, nul) << 4) | (wt_delim_bits(v.z, nul) << 8) | (wt_delim_bits(v.w, nul) << 12);
    const uint32_t prev_delim = is_delim((uint8_t)prev_last) ? 1u : 0u;
    uint32_t starts = ~dm & ((dm << 1) | prev_delim) & 0xFFFFu;
    while (starts) {
      const int s = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint64_t off = seg * 16 + s;
      if (off >= hi_b) break;
      if (off < lo_b) continue;
      // the 16 bytes from the word start on
      uint32_t w0, w1, w2, w3, w4;
      switch (s >> 2) {
        case 0: w0 = v.x; w1 = v.y; w2 = v.z; w3 = v.w; w4 = nx.x; break;
        case 1: w0 = v.y; w1 = v.z; w2 = v.w; w3 = nx.x; w4 = nx.y; break;
        case 2: w0 = v.z; w1 = v.w; w2 = nx.x; w3 = nx.y; w4 = nx.z; break;
        default: w0 = v.w; w1 = nx.x; w2 = nx.y; w3 = nx.z; w4 = nx.w; break;
      }
      const uint32_t bs = (uint32_t)(s & 3) * 8u;
      uint32_t b0 = __funnelshift_r(w0, w1, bs), b1 = __funnelshift_r(w1, w2, bs), b2 = __funnelshift_r(w2, w3, bs), b3 = __funnelshift_r(w3, w4, bs);
      uint32_t dummy = 0;
      const uint32_t wm = wt_delim_bits(b0, dummy) | (wt_delim_bits(b1, dummy) << 4) | (wt_delim_bits(b2, dummy) << 8) | (wt_delim_bits(b3, dummy) << 12);
      const uint32_t L = wm ? (uint32_t)(__ffs(wm) - 1) : 16u;
      if (L > WT_SHORT_MAX) {  // 16 bytes or more: byte loop + comparison through the corpus (rare)
        uint64_t h64; uint32_t djb;
        const uint32_t len = wt_scan_word(corpus, off, n, h64, djb);
        const uint64_t tag = (djb & 0xFFFu) | (((h64 >> 40) & 0xFFFu) << 12);
        wt_global_insert(tbl, corpus, n, off, len, h64, tag, 1ull, false, 0ull, 0ull);
        continue;
      }
      // zero the bytes from L on: (lo, hi) is the word
      {
        const uint32_t k0 = L >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 - L))) & (L ? 0xFFFFFFFFu : 0u);
        const uint32_t k1 = L >= 8 ? 0xFFFFFFFFu : (L <= 4 ? 0u : 0xFFFFFFFFu >> (8 * (8 - L)));
        const uint32_t k2 = L >= 12 ? 0xFFFFFFFFu : (L <= 8 ? 0u : 0xFFFFFFFFu >> (8 * (12 - L)));
        const uint32_t k3 = L <= 12 ? 0u : 0xFFFFFFFFu >> (8 * (16 - L));
        b0 &= k0; b1 &= k1; b2 &= k2; b3 &= k3;
      }
      const uint64_t wlo = ((uint64_t)b1 << 32) | b0, whi = ((uint64_t)b3 << 32) | b2;
      const uint64_t h = wt_short_hash(wlo, whi);
      const uint32_t rel = (uint32_t)(off - span0);
      uint32_t slot = (uint32_t)h & (LT_SLOTS - 1);
      bool done = false;
#pragma unroll 1
      for (int probe = 0; probe < LT_PROBES; probe++) {
        unsigned long long cur = *(volatile unsigned long long *)&ltag[slot];
        if (cur == 0ull) {
          cur = atomicCAS(&ltag[slot], 0ull, (unsigned long long)h);
          if (cur == 0ull) { wt_sts16(&lpay[slot], wlo, whi); cur = h; atomicAdd(&lcnt[slot], 1u); atomicMin(&lmin[slot], rel); done = true; break; }
        }
        if (cur == h) {
          const ulonglong2 pay = wt_lds16(&lpay[slot]);
          if (pay.x == wlo && pay.y == whi) { atomicAdd(&lcnt[slot], 1u); atomicMin(&lmin[slot], rel); done = true; break; }
          // (bytes not written yet, or another word with the same hash: try the next slot; a word may end up in two
          //  slots or go to the global table directly -- the global table is the one that decides identity)
        }
        slot = (slot + 1) & (LT_SLOTS - 1);
      }
      if (!done) {
        uint32_t djb;
        wt_short_len_djb(wlo, whi, djb);
        const uint64_t tag = (djb & 0xFFFu) | (((h >> 40) & 0xFFFu) << 12);
        wt_global_insert(tbl, corpus, n, off, L, h, tag, 1ull, true, wlo, whi);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < LT_SLOTS; i += WT_THREADS) {
    const unsigned long long h = ltag[i];
    if (h == 0ull) continue;
    const ulonglong2 pay = lpay[i];
    uint32_t djb;
    const uint32_t len = wt_short_len_djb(pay.x, pay.y, djb);
    const uint64_t tag = (djb & 0xFFFu) | (((h >> 40) & 0xFFFu) << 12);
    wt_global_insert(tbl, corpus, n, span0 + lmin[i], len, h, tag, (unsigned long long)lcnt[i], true, pay.x, pay.y);
  }
  if (nul) atomicOr(tbl.flags, 2u);
}

__global__ void wt_fill(WSlot *slots, uint64_t cap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
    slots[i] = WSlot{WT_EMPTY, 0ull, 0ull, 0ull};
}

// occupied slots -> (bucket << 40 | first offset, count); order is fixed by the sort that follows
__global__ void wt_compact(const WSlot *__restrict__ slots, uint64_t cap, unsigned long long *__restrict__ sort_keys,
                           unsigned long long *__restrict__ out_counts, unsigned int *cursor) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = slots[i].key;
    if (k != WT_EMPTY) {
      const unsigned int j = atomicAdd(cursor, 1u);
      sort_keys[j] = ((k & 0xFFFull) << 40) | (k >> WT_TAG_BITS);
      out_counts[j] = slots[i].count;
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
                WordTableSoA tbl, unsigned long long *__restrict__ gfirst) {
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
