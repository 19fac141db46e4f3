// encoder_impl.cuh -- rank-ordered BPE encoder on the device.
//
// The reference has no encoder (reference shredword/base.py:107-109 raises NotImplementedError);
// the step is defined by its helpers get_stats (base.py:10-20) and merge (base.py:22-36): per
// whitespace-delimited word, repeatedly take the adjacent pair with the lowest merge rank and replace
// all its non-overlapping occurrences left to right. Applied to a training word this reproduces the
// trainer's final segmentation of that word (a merge of rank r only creates pairs containing id
// 256+r, which can only match ranks > r), which is how the tests pin it.
//
// enc_fused (below, "single-pass encoder") is the path every text takes: one kernel, the text read once, the ids written once.
// A text that holds a word longer than ENC_SHORT bytes falls back, piece by piece, to three kernels:
//   enc_words  : 16 bytes per thread -> delimiter bit-mask -> word list in shared memory -> one thread
//                encodes one word in shared memory (merge-rank table: 16-byte slots, L2 resident) and
//                parks the tokens in tmp[word start ...] (a word of L bytes yields <= L tokens, so the
//                regions never collide); tmp[start+L-1] = -(ntok)-1 when ntok < L.
//   (scan)     : per-tile token / word totals -> exclusive offsets
//   enc_gather : re-derives the word starts, block-scans the per-word token counts and writes the
//                compact id stream plus the optional per-word token counts.
#pragma once

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <vector>

#include "device_util.cuh"

namespace swb {

constexpr int ENC_THREADS = 256;
constexpr int ENC_TILE = ENC_THREADS * 16;  // bytes per block iteration
constexpr int ENC_SHORT = 64;               // words up to this many bytes are encoded in shared memory
constexpr uint32_t RANK_NONE = 0xFFFFFFFFu;

// Word memo: text is Zipfian, so most word occurrences repeat a word that was already encoded. A 64-byte
// slot caches word bytes -> tokens; it is filled by whoever encodes the word first (one CAS on the tag) and
// read with L2-coherent loads. No busy state and no ordering between the tag and the payload: a reader
// accepts a slot only if tag, length, BYTES and the token check word all match, and encodes the word
// itself otherwise (torn or foreign slots are simply misses).
constexpr int MEMO_MAX_LEN = 14, MEMO_MAX_TOK = 9;
struct __align__(16) MemoSlot {
  unsigned long long tag;             // word hash (never 0); 0 = free
  uint8_t len, ntok, bytes[MEMO_MAX_LEN];
  int32_t tok[MEMO_MAX_TOK];
  uint32_t check;                     // tag ^ ntok ^ xor of the tokens, folded to 32 bits
};
static_assert(sizeof(MemoSlot) == 64, "memo slot is one 64-byte block");
struct MemoDev { MemoSlot *slots; uint32_t mask; };

struct RankSlot { unsigned long long key; unsigned long long val; };  // val = rank << 32 | new_id
struct RankTableDev { const RankSlot *slots; uint32_t mask; };

__device__ __forceinline__ unsigned long long enc_lookup(const RankTableDev &t, int a, int b) {
  const unsigned long long k = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b;
  uint32_t h = (uint32_t)dmix64(k) & t.mask;
  for (;;) {
    const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2 *>(t.slots + h));
    if (s.x == k) return s.y;
    if (s.x == ~0ull) return ~0ull;
    h = (h + 1) & t.mask;
  }
}

// In-place rank-ordered merge of ids[0..L). Returns the new length.
__device__ __forceinline__ uint32_t enc_word(int *ids, uint32_t L, const RankTableDev &t) {
  while (L >= 2) {
    unsigned long long best = ~0ull;
    uint32_t kb = 0;
    for (uint32_t k = 0; k + 1 < L; k++) {
      const unsigned long long v = enc_lookup(t, ids[k], ids[k + 1]);
      if (v < best) { best = v; kb = k; }  // rank is in the high half: lowest rank wins, leftmost on ties
    }
    if (best == ~0ull) break;
    const int a = ids[kb], b = ids[kb + 1], nid = (int)(uint32_t)(best & 0xFFFFFFFFu);
    uint32_t w = 0, r = 0;
    while (r < L) {
      if (r + 1 < L && ids[r] == a && ids[r + 1] == b) { ids[w++] = nid; r += 2; }
      else ids[w++] = ids[r++];
    }
    L = w;
  }
  return L;
}

__device__ __forceinline__ uint32_t enc_delim_bits(uint32_t w) {
  const uint32_t m = __vcmpeq4(w, 0x20202020u) | __vcmpeq4(w, 0x0a0a0a0au) | __vcmpeq4(w, 0x09090909u) |
                     __vcmpeq4(w, 0x0d0d0d0du);
  return ((m & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t enc_starts(const uint8_t *__restrict__ text, uint64_t seg) {
  const uint4 v = *reinterpret_cast<const uint4 *>(text + seg * 16);
  const uint32_t dm = enc_delim_bits(v.x) | (enc_delim_bits(v.y) << 4) | (enc_delim_bits(v.z) << 8) | (enc_delim_bits(v.w) << 12);
  const uint32_t prev_delim = (seg == 0) ? 1u : (is_delim(text[seg * 16 - 1]) ? 1u : 0u);
  return ~dm & ((dm << 1) | prev_delim) & 0xFFFFu;
}
// 16 bytes starting at an arbitrary byte offset (two aligned 16-byte loads + a funnel shift); the buffers
// are padded, so reading past the end of a word or of the text is safe
__device__ __forceinline__ uint4 enc_load16(const uint8_t *__restrict__ text, uint64_t off) {
  const uint4 *p = reinterpret_cast<const uint4 *>(text + (off & ~15ull));
  const uint4 a = p[0], b = p[1];
  const uint32_t sh = (uint32_t)(off & 15), bs = (sh & 3) * 8;
  uint32_t v0, v1, v2, v3, v4;
  switch (sh >> 2) {
    case 0: v0 = a.x; v1 = a.y; v2 = a.z; v3 = a.w; v4 = b.x; break;
    case 1: v0 = a.y; v1 = a.z; v2 = a.w; v3 = b.x; v4 = b.y; break;
    case 2: v0 = a.z; v1 = a.w; v2 = b.x; v3 = b.y; v4 = b.z; break;
    default: v0 = a.w; v1 = b.x; v2 = b.y; v3 = b.z; v4 = b.w; break;
  }
  return make_uint4(__funnelshift_r(v0, v1, bs), __funnelshift_r(v1, v2, bs), __funnelshift_r(v2, v3, bs), __funnelshift_r(v3, v4, bs));
}
__device__ __forceinline__ uint32_t enc_word_len(const uint8_t *__restrict__ text, uint64_t off, uint64_t n) {
  uint64_t i = off;
  while (i < n && !is_delim(text[i])) ++i;
  return (uint32_t)(i - off);
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *warp_tot /* [ENC_THREADS/32] shared */, uint32_t &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < ENC_THREADS / 32; i++) { if (i < w) base += warp_tot[i]; tot += warp_tot[i]; }
  __syncthreads();
  total = tot;
  return base + inc - v;
}


// ---- memo helpers. A word of up to MEMO_MAX_LEN bytes has two candidate slots (two-way: a word whose first slot belongs to
// another word still gets memoised); whoever encodes it first fills the first free one.
struct MemoKey { unsigned long long h; uint32_t e0, e1, e2, e3; uint32_t w4[4]; };
__device__ __forceinline__ MemoKey memo_key(uint32_t L, uint4 wv) {
  MemoKey k;
  k.w4[0] = wv.x; k.w4[1] = wv.y; k.w4[2] = wv.z; k.w4[3] = wv.w;  // zero the bytes from L on
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int keep = (int)L - 4 * q;
    k.w4[q] = keep >= 4 ? k.w4[q] : (keep <= 0 ? 0u : (k.w4[q] & (0xFFFFFFFFu >> (8 * (4 - keep)))));
  }
  unsigned long long h = dmix64(((unsigned long long)k.w4[1] << 32 | k.w4[0]) ^ 0x9E3779B97F4A7C15ull) ^
                         dmix64(((unsigned long long)k.w4[3] << 32 | k.w4[2]) + ((unsigned long long)L << 56) + 0x632BE59BD9B4E019ull);
  k.h = dmix64(h) | 1ull;
  // expected image of {len, ntok, bytes}: len | ntok << 8 | bytes << 16 over words 2..5 of the slot (ntok byte masked out)
  k.e0 = L | (k.w4[0] << 16);
  k.e1 = (k.w4[0] >> 16) | (k.w4[1] << 16);
  k.e2 = (k.w4[1] >> 16) | (k.w4[2] << 16);
  k.e3 = (k.w4[2] >> 16) | (k.w4[3] << 16);
  return k;
}
__device__ __forceinline__ MemoSlot *memo_way(const MemoDev &memo, unsigned long long h, int way) {
  return memo.slots + ((uint32_t)(way == 0 ? h >> 20 : h >> 42) & memo.mask);
}
// tokens of the word if `slot` holds them (nt >= 1, ids[0 .. nt) filled), else 0; *tag = the slot's tag (0 = free)
__device__ __forceinline__ uint32_t memo_read(const MemoSlot *slot, const MemoKey &k, uint32_t L, int *ids, unsigned long long *tag) {
  const uint4 q0 = __ldcg(reinterpret_cast<const uint4 *>(slot));       // tag, len, ntok, bytes[0..5]
  const uint4 q1 = __ldcg(reinterpret_cast<const uint4 *>(slot) + 1);   // bytes[6..13], tok[0..1]
  const uint4 q2 = __ldcg(reinterpret_cast<const uint4 *>(slot) + 2);   // tok[2..5]
  const uint4 q3 = __ldcg(reinterpret_cast<const uint4 *>(slot) + 3);   // tok[6..8], check
  *tag = ((unsigned long long)q0.y << 32) | q0.x;
  if (!(*tag == k.h && (q0.z & 0xFFFF00FFu) == k.e0 && q0.w == k.e1 && q1.x == k.e2 && q1.y == k.e3)) return 0;
  const uint32_t mt = (q0.z >> 8) & 0xFFu;
  const int tk[MEMO_MAX_TOK] = {(int)q1.z, (int)q1.w, (int)q2.x, (int)q2.y, (int)q2.z, (int)q2.w, (int)q3.x, (int)q3.y, (int)q3.z};
  uint32_t ck = (uint32_t)k.h ^ (uint32_t)(k.h >> 32) ^ mt;
#pragma unroll
  for (int q = 0; q < MEMO_MAX_TOK; q++) ck ^= (uint32_t)tk[q] * (2u * q + 3u);
  if (!(ck == q3.w && mt >= 1 && mt <= MEMO_MAX_TOK && mt <= L)) return 0;
#pragma unroll
  for (int q = 0; q < MEMO_MAX_TOK; q++) if (q < (int)mt) ids[q] = tk[q];
  return mt;
}
// Memo probe only: the tokens of a word of L <= MEMO_MAX_LEN bytes if the memo has them (returns nt >= 1 and fills
// ids[0 .. nt)), 0 if not. Cheap and uniform -- the single-pass kernel runs it for every word first and hands the
// misses, whose full encoding is long and data dependent, to a second phase where they are dealt out evenly.
__device__ __forceinline__ uint32_t enc_memo_probe(uint32_t L, uint4 wv, int *ids, const MemoDev &memo) {
  const MemoKey k = memo_key(L, wv);
  unsigned long long tag;
  uint32_t nt = memo_read(memo_way(memo, k.h, 0), k, L, ids, &tag);
  if (!nt && tag != 0ull && tag != k.h) nt = memo_read(memo_way(memo, k.h, 1), k, L, ids, &tag);  // first slot is another word's
  return nt;
}

// Encodes one word of L <= ENC_SHORT bytes starting at text[off] into ids[0 .. nt) (shared memory, room for L ints);
// wv = the word's first 16 bytes. Words of up to MEMO_MAX_LEN bytes go through the memo. Returns nt.
__device__ __forceinline__ uint32_t enc_short_word(const uint8_t *__restrict__ text, uint64_t off, uint32_t L, uint4 wv, int *ids,
                                                   const RankTableDev &tbl, const MemoDev &memo, const int32_t *bmap) {
  if (L > MEMO_MAX_LEN) {
    for (uint32_t k = 0; k < L; k++) ids[k] = bmap[text[off + k]];
    return enc_word(ids, L, tbl);
  }
  const MemoKey k = memo_key(L, wv);
  MemoSlot *slot = memo_way(memo, k.h, 0);
  unsigned long long tag;
  uint32_t nt = memo_read(slot, k, L, ids, &tag);
  if (nt) return nt;
  if (tag != 0ull && tag != k.h) {
    slot = memo_way(memo, k.h, 1);
    nt = memo_read(slot, k, L, ids, &tag);
    if (nt) return nt;
  }
  for (uint32_t q = 0; q < L; q++) ids[q] = bmap[(k.w4[q >> 2] >> (8 * (q & 3))) & 0xFFu];
  nt = enc_word(ids, L, tbl);
  if (tag == 0ull && nt <= MEMO_MAX_TOK && atomicCAS(&slot->tag, 0ull, k.h) == 0ull) {  // first encoder fills the slot
    int tk[MEMO_MAX_TOK];
    uint32_t ck = (uint32_t)k.h ^ (uint32_t)(k.h >> 32) ^ nt;
#pragma unroll
    for (int q = 0; q < MEMO_MAX_TOK; q++) { tk[q] = q < (int)nt ? ids[q] : 0; ck ^= (uint32_t)tk[q] * (2u * q + 3u); }
    uint4 *dst = reinterpret_cast<uint4 *>(slot);
    // (the tag, in words 0..1, is already there: written by the CAS)
    reinterpret_cast<uint32_t *>(slot)[2] = k.e0 | (nt << 8);
    reinterpret_cast<uint32_t *>(slot)[3] = k.e1;
    dst[1] = make_uint4(k.e2, k.e3, (uint32_t)tk[0], (uint32_t)tk[1]);
    dst[2] = make_uint4((uint32_t)tk[2], (uint32_t)tk[3], (uint32_t)tk[4], (uint32_t)tk[5]);
    dst[3] = make_uint4((uint32_t)tk[6], (uint32_t)tk[7], (uint32_t)tk[8], ck);
  }
  return nt;
}

// ---------------------------------------------------------------- single-pass encoder
// One kernel, one pass over the text, no per-byte scratch: tiles are handed out in text order; a block finds the word
// starts of its tile, encodes the words in shared memory (words dealt out evenly over the threads), scans the token counts,
// learns where its tokens go from the tiles before it (decoupled look-back over one 64-bit descriptor per tile) and writes
// them, compacted, with coalesced stores. Traffic = the text once + the ids once.
//   descriptor = status << 62 | tokens << 31 | words   (status 1 = this tile's own totals, 2 = totals of all tiles up to it)
// A word longer than ENC_SHORT bytes does not fit the tile's shared memory: it raises `fallback` and the host encodes the
// whole piece with the three-kernel path above (enc_words / scan / enc_gather) instead.
constexpr int ENCF_MAX_WORDS = ENC_TILE / 2;
__global__ void __launch_bounds__(ENC_THREADS)
enc_fused(const uint8_t *__restrict__ text, uint64_t n, uint64_t lo_b, uint64_t hi_b, RankTableDev tbl, MemoDev memo,
          const int32_t *__restrict__ byte_map, unsigned long long *__restrict__ desc, unsigned int *__restrict__ tile_counter,
          unsigned int *__restrict__ fallback, unsigned long long tok_base, unsigned long long word_base, int32_t *__restrict__ out,
          uint64_t cap_ids, uint32_t *__restrict__ word_ntok, uint64_t cap_words, int32_t neg_id) {
  __shared__ int stok[ENC_TILE + ENC_SHORT];       // tokens of the word that starts at byte r of the tile: stok[r ...]
  __shared__ uint16_t wstart[ENCF_MAX_WORDS];      // ordered word starts (byte position in the tile)
  __shared__ uint16_t wnt[ENCF_MAX_WORDS];         // tokens per word
  __shared__ uint16_t woff[ENCF_MAX_WORDS];        // exclusive token offset of each word inside the tile
  __shared__ int32_t bmap[256];
  __shared__ uint32_t wt[ENC_THREADS / 32];
  __shared__ unsigned int s_tile, s_nmiss;
  __shared__ unsigned long long s_excl;
  for (int i = threadIdx.x; i < 256; i += ENC_THREADS) bmap[i] = byte_map[i];
  const uint64_t seg_lo = lo_b / 16, seg_hi = (hi_b + 15) / 16;
  const uint64_t n_tiles = (seg_hi - seg_lo + ENC_THREADS - 1) / ENC_THREADS;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint64_t tile = s_tile;
    if (tile >= n_tiles) return;
    const uint64_t seg = seg_lo + tile * ENC_THREADS + threadIdx.x;
    const uint64_t tile_byte0 = (seg_lo + tile * ENC_THREADS) * 16;
    // ---- word starts of this thread's 16 bytes, in text order
    uint32_t starts = 0;
    if (seg < seg_hi) {
      starts = enc_starts(text, seg);
      uint32_t st = starts;
      while (st) {  // only the words that start inside [lo_b, hi_b) and before the end of the text
        const int sb = __ffs(st) - 1;
        st &= st - 1;
        const uint64_t off = seg * 16 + sb;
        if (off < lo_b || off >= hi_b || off >= n) starts &= ~(1u << sb);
      }
    }
    uint32_t nw;
    uint32_t wbase = block_excl_scan(__popc(starts), wt, nw);
    {
      uint32_t st = starts;
      while (st) { const int sb = __ffs(st) - 1; st &= st - 1; wstart[wbase++] = (uint16_t)(threadIdx.x * 16 + sb); }
    }
    __syncthreads();
    // ---- encode, phase 1: word j -> thread j % ENC_THREADS looks it up in the memo (uniform work); what the memo does not
    // have is listed (the ordered list of starts is no longer needed as a whole: woff doubles as the miss list until the scan)
    if (threadIdx.x == 0) s_nmiss = 0;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < nw; j += ENC_THREADS) {
      const uint32_t rel = wstart[j];
      const uint64_t off = tile_byte0 + rel;
      const uint4 wv = enc_load16(text, off);
      const uint32_t dmask = enc_delim_bits(wv.x) | (enc_delim_bits(wv.y) << 4) | (enc_delim_bits(wv.z) << 8) | (enc_delim_bits(wv.w) << 12);
      uint32_t L = dmask ? (uint32_t)(__ffs(dmask) - 1) : 16u;
      if (off + L > n) L = (uint32_t)(n - off);
      uint32_t nt = 0;
      if (L <= MEMO_MAX_LEN) nt = enc_memo_probe(L, wv, stok + rel, memo);
      if (!nt) woff[atomicAdd(&s_nmiss, 1u)] = (uint16_t)j;
      wnt[j] = (uint16_t)nt;
    }
    __syncthreads();
    // ---- phase 2: the misses, dealt out evenly (the long, data-dependent path runs with as many lanes busy as there are misses)
    {
      const uint32_t nmiss = s_nmiss;
      for (uint32_t q = threadIdx.x; q < nmiss; q += ENC_THREADS) {
        const uint32_t j = woff[q];
        const uint32_t rel = wstart[j];
        const uint64_t off = tile_byte0 + rel;
        const uint4 wv = enc_load16(text, off);
        const uint32_t dmask = enc_delim_bits(wv.x) | (enc_delim_bits(wv.y) << 4) | (enc_delim_bits(wv.z) << 8) | (enc_delim_bits(wv.w) << 12);
        uint32_t L = dmask ? (uint32_t)(__ffs(dmask) - 1) : 16u;
        if (off + L > n) L = (uint32_t)(n - off);
        if (L >= 15) L = enc_word_len(text, off, n);
        uint32_t nt = 0;
        if (L <= ENC_SHORT) nt = enc_short_word(text, off, L, wv, stok + rel, tbl, memo, bmap);
        else atomicOr(fallback, 1u);
        wnt[j] = (uint16_t)nt;
      }
    }
    __syncthreads();
    // ---- token offsets of the words (thread t scans words 8t .. 8t+7)
    uint32_t mine[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { const uint32_t j = threadIdx.x * 8 + q; mine[q] = j < nw ? wnt[j] : 0u; sum += mine[q]; }
    uint32_t ntok;
    uint32_t tb = block_excl_scan(sum, wt, ntok);
#pragma unroll
    for (int q = 0; q < 8; q++) { const uint32_t j = threadIdx.x * 8 + q; if (j < nw) woff[j] = (uint16_t)tb; tb += mine[q]; }
    // ---- where the tile's tokens and words go: look back over the tiles before this one (warp 0)
    if (threadIdx.x < 32) {
      const unsigned long long agg = ((unsigned long long)ntok << 31) | nw;
      unsigned long long excl = 0;
      if (tile == 0) {
        if (threadIdx.x == 0) *(volatile unsigned long long *)&desc[0] = (2ull << 62) | agg;
      } else {
        if (threadIdx.x == 0) *(volatile unsigned long long *)&desc[tile] = (1ull << 62) | agg;
        long long idx = (long long)tile - 1;
        for (;;) {
          // four windows of 32 descriptors are requested together (one L2 round trip), then looked at nearest first
          unsigned long long dq[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const long long my = idx - 32 * u - (long long)threadIdx.x;
            dq[u] = my >= 0 ? *(volatile unsigned long long *)&desc[my] : (2ull << 62);  // (before the first tile: an empty prefix)
          }
          bool found = false;
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (found) continue;
            const long long my = idx - 32 * u - (long long)threadIdx.x;
            unsigned long long d = dq[u];
            while ((d >> 62) == 0ull) d = *(volatile unsigned long long *)&desc[my];  // that tile has not even published its own totals yet
            const unsigned int pm = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
            const int first = __ffs(pm) - 1;  // nearest tile (lowest lane) that already knows its inclusive prefix
            unsigned long long v = ((int)threadIdx.x <= first || first < 0) ? (d & 0x3FFFFFFFFFFFFFFFull) : 0ull;
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) v += __shfl_down_sync(0xffffffffu, v, dd);
            v = __shfl_sync(0xffffffffu, v, 0);
            excl += v;
            found = first >= 0;
          }
          if (found) break;
          idx -= 128;
        }
        if (threadIdx.x == 0) *(volatile unsigned long long *)&desc[tile] = (2ull << 62) | (excl + agg);
      }
      if (threadIdx.x == 0) s_excl = excl;
    }
    __syncthreads();
    // ---- write: word j's tokens go to out[to + woff[j] ...]; consecutive words are consecutive threads of a warp, so a
    // warp's stores cover one contiguous stretch of the output
    const unsigned long long excl = s_excl;
    const uint64_t to = tok_base + (excl >> 31), wo = word_base + (excl & 0x7FFFFFFFull);
    for (uint32_t j = threadIdx.x; j < nw; j += ENC_THREADS) {
      const uint32_t rel = wstart[j], nt = wnt[j];
      const uint64_t o = to + woff[j];
      for (uint32_t k = 0; k < nt; k++)
        if (o + k < cap_ids) { const int v = stok[rel + k]; out[o + k] = (v == UNK_CODE) ? neg_id : v; }
      if (word_ntok && wo + j < cap_words) word_ntok[wo + j] = nt;
    }
  }
}

// text: n bytes, 16-byte aligned, followed by >= 16 bytes of ' '. One tile per block iteration.
__global__ void __launch_bounds__(ENC_THREADS)
enc_words(const uint8_t *__restrict__ text, uint64_t n, RankTableDev tbl, MemoDev memo, const int32_t *__restrict__ byte_map,
          int32_t *__restrict__ tmp, uint32_t *__restrict__ tile_ntok, uint32_t *__restrict__ tile_nwords) {
  __shared__ int stok[ENC_TILE + ENC_SHORT];
  __shared__ uint16_t wstart[ENC_TILE / 2];
  __shared__ int32_t bmap[256];
  __shared__ unsigned int s_nwords, s_ntok;
  for (int i = threadIdx.x; i < 256; i += ENC_THREADS) bmap[i] = byte_map[i];
  const uint64_t n_tiles = (n + ENC_TILE - 1) / ENC_TILE;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) { s_nwords = 0; s_ntok = 0; }
    __syncthreads();
    const uint64_t seg = tile * ENC_THREADS + threadIdx.x;
    if (seg * 16 < n) {
      uint32_t starts = enc_starts(text, seg);
      if (starts) {
        const int c = __popc(starts);
        unsigned int base = atomicAdd(&s_nwords, (unsigned int)c);  // order inside the tile is irrelevant here
        while (starts) {
          const int s = __ffs(starts) - 1;
          starts &= starts - 1;
          if (seg * 16 + s < n) wstart[base++] = (uint16_t)(threadIdx.x * 16 + s);
          else atomicSub(&s_nwords, 1u);
        }
      }
    }
    __syncthreads();
    const unsigned int nw = s_nwords;
    unsigned int my_tok = 0;
    for (unsigned int j = threadIdx.x; j < nw; j += ENC_THREADS) {
      const uint32_t rel = wstart[j];
      const uint64_t off = tile * ENC_TILE + rel;
      // the word's first 16 bytes in registers: its length comes from the delimiter mask, not from a byte loop
      const uint4 wv = enc_load16(text, off);
      const uint32_t dmask = enc_delim_bits(wv.x) | (enc_delim_bits(wv.y) << 4) | (enc_delim_bits(wv.z) << 8) | (enc_delim_bits(wv.w) << 12);
      uint32_t L = dmask ? (uint32_t)(__ffs(dmask) - 1) : 16u;
      if (off + L > n) L = (uint32_t)(n - off);
      if (L >= 15) L = enc_word_len(text, off, n);
      uint32_t nt;
      if (L <= ENC_SHORT) {
        int *ids = stok + rel;
        nt = enc_short_word(text, off, L, wv, ids, tbl, memo, bmap);
        for (uint32_t k = 0; k < nt; k++) tmp[off + k] = ids[k];
      } else {  // long word: encode in place in global memory
        int *ids = tmp + off;
        for (uint32_t k = 0; k < L; k++) ids[k] = bmap[text[off + k]];
        nt = enc_word(ids, L, tbl);
      }
      if (nt < L) tmp[off + L - 1] = -(int)nt - 1;
      my_tok += nt;
    }
    if (my_tok) atomicAdd(&s_ntok, my_tok);
    __syncthreads();
    if (threadIdx.x == 0) { tile_ntok[tile] = s_ntok; tile_nwords[tile] = nw; }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(ENC_THREADS)
enc_gather(const uint8_t *__restrict__ text, uint64_t n, const int32_t *__restrict__ tmp,
           const unsigned long long *__restrict__ tile_tok_off, const unsigned long long *__restrict__ tile_word_off,
           unsigned long long tok_base, unsigned long long word_base, int32_t *__restrict__ out, uint64_t cap_ids,
           uint32_t *__restrict__ word_ntok, uint64_t cap_words, int32_t neg_id) {
  __shared__ uint32_t wt[ENC_THREADS / 32];
  const uint64_t n_tiles = (n + ENC_TILE - 1) / ENC_TILE;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t seg = tile * ENC_THREADS + threadIdx.x;
    uint32_t starts = (seg * 16 < n) ? enc_starts(text, seg) : 0u;
    uint32_t ntoks[8];  // at most 8 word starts in 16 bytes
    uint32_t nmine = 0, tok_sum = 0;
    uint32_t st = starts;
    while (st) {
      const int s = __ffs(st) - 1;
      st &= st - 1;
      const uint64_t off = seg * 16 + s;
      if (off >= n) { starts &= ~(1u << s); continue; }
      const uint32_t L = enc_word_len(text, off, n);
      const int last = tmp[off + L - 1];
      const uint32_t nt = last >= 0 ? L : (uint32_t)(-last - 1);
      ntoks[nmine] = nt; nmine++;
      tok_sum += nt;
    }
    uint32_t tot_t, tot_w;
    const uint32_t tbase = block_excl_scan(tok_sum, wt, tot_t);
    const uint32_t wbase = block_excl_scan(nmine, wt, tot_w);
    uint64_t to = tok_base + tile_tok_off[tile] + tbase;
    uint64_t wo = word_base + tile_word_off[tile] + wbase;
    uint32_t j = 0;
    st = starts;
    while (st) {
      const int s = __ffs(st) - 1;
      st &= st - 1;
      const uint64_t off = seg * 16 + s;
      const uint32_t nt = ntoks[j];
      for (uint32_t k = 0; k < nt; k++)
        if (to + k < cap_ids) { const int v = tmp[off + k]; out[to + k] = (v == UNK_CODE) ? neg_id : v; }
      if (word_ntok && wo < cap_words) word_ntok[wo] = nt;
      to += nt; wo++; j++;
    }
  }
}

class EncoderImpl {
 public:
  EncoderImpl(const int32_t *triples, size_t M, const int32_t *byte_map256, int32_t unk_for_negative) {
    merges_.assign(triples, triples + 3 * M);
    for (int i = 0; i < 256; i++) bmap_[i] = byte_map256 ? byte_map256[i] : i;
    // a negative id (only a negative unk_id can produce one) travels as UNK_CODE on the device
    neg_id_ = unk_for_negative;
    for (int i = 0; i < 256; i++) if (bmap_[i] < 0) { neg_id_ = bmap_[i]; dev_bmap_[i] = UNK_CODE; } else dev_bmap_[i] = bmap_[i];
    // host decode table
    tok_off_.assign(256 + M + 1, 0);
    std::vector<std::vector<uint8_t>> toks(256 + M);
    for (int i = 0; i < 256; i++) toks[i] = {(uint8_t)i};
    for (size_t m = 0; m < M; m++) {
      const int32_t a = triples[3 * m], b = triples[3 * m + 1];
      std::vector<uint8_t> v;
      if (a >= 0 && (size_t)a < 256 + m) v = toks[a];
      if (b >= 0 && (size_t)b < 256 + m) v.insert(v.end(), toks[b].begin(), toks[b].end());
      toks[256 + m] = std::move(v);
    }
    for (size_t i = 0; i < toks.size(); i++) {
      tok_off_[i] = tok_bytes_.size();
      tok_bytes_.insert(tok_bytes_.end(), toks[i].begin(), toks[i].end());
    }
    tok_off_[toks.size()] = tok_bytes_.size();
  }
  ~EncoderImpl() {
    if (s_in_) {
      cudaStreamDestroy(s_in_); cudaStreamDestroy(s_out_);
      for (int i = 0; i < 2; i++) { cudaEventDestroy(ev_in_[i]); cudaEventDestroy(ev_gather_[i]); cudaEventDestroy(ev_out_[i]); }
    }
    if (stream_) cudaStreamDestroy(stream_);
  }

  uint64_t launches = 0;

  void ensure_device() {
    if (stream_) return;
    int nd = 0;
    cudaError_t e = cudaGetDeviceCount(&nd);
    if (e != cudaSuccess || nd == 0)
      throw Error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                  " (this library has no CPU fallback)");
    int dev = 0;
    SWB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SWB_CUDA(cudaGetDeviceProperties(&prop, dev));
    sms_ = prop.multiProcessorCount;
    SWB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    // merge-rank table (a repeated pair keeps its lowest rank)
    const size_t M = merges_.size() / 3;
    uint64_t cap = 64;
    while (cap < 4 * M + 16) cap <<= 1;
    std::vector<RankSlot> h(cap, RankSlot{~0ull, ~0ull});
    for (size_t r = 0; r < M; r++) {
      const int32_t ma = merges_[3 * r] < 0 ? UNK_CODE : merges_[3 * r], mb = merges_[3 * r + 1] < 0 ? UNK_CODE : merges_[3 * r + 1];
      const unsigned long long k = ((unsigned long long)(uint32_t)ma << 32) | (uint32_t)mb;
      uint64_t hh = dmix64(k) & (cap - 1);
      bool dup = false;
      while (h[hh].key != ~0ull) { if (h[hh].key == k) { dup = true; break; } hh = (hh + 1) & (cap - 1); }
      if (!dup) h[hh] = RankSlot{k, ((unsigned long long)r << 32) | (uint32_t)merges_[3 * r + 2]};
    }
    slots_.alloc(cap);
    SWB_CUDA(cudaMemcpyAsync(slots_.get(), h.data(), cap * sizeof(RankSlot), cudaMemcpyHostToDevice, stream_));
    d_bmap_.alloc(256);
    SWB_CUDA(cudaMemcpyAsync(d_bmap_.get(), dev_bmap_, sizeof dev_bmap_, cudaMemcpyHostToDevice, stream_));
    SWB_CUDA(cudaStreamSynchronize(stream_));
    tbl_ = RankTableDev{slots_.get(), (uint32_t)(cap - 1)};
    const uint64_t memo_cap = 1ull << 22;  // 256 MB, two-way: room for millions of distinct words; the hot ones stay L2 resident
    memo_slots_.alloc(memo_cap);
    SWB_CUDA(cudaMemsetAsync(memo_slots_.get(), 0, memo_cap * sizeof(MemoSlot), stream_));
    SWB_CUDA(cudaStreamSynchronize(stream_));
    memo_ = MemoDev{memo_slots_.get(), (uint32_t)(memo_cap - 1)};
  }

  // The single-pass kernel (enc_fused) over [0, n), in ranges of at most 1 GiB (31-bit token / word counts per range).
  // false: the text holds a word longer than ENC_SHORT bytes -- the caller encodes the piece with the general path.
  uint64_t general_pieces = 0;  // pieces that went through the general three-kernel path
  static bool no_fused() {
    static const bool off = getenv("SWB_NO_FUSED_ENCODE") && atoi(getenv("SWB_NO_FUSED_ENCODE")) > 0;  // (tests: force the general path)
    return off;
  }
  bool encode_fused(const uint8_t *d_text, uint64_t n, int32_t *d_out, uint64_t cap_ids, uint64_t tok_base, uint32_t *d_word_ntok,
                    uint64_t cap_words, uint64_t word_base, uint64_t *ntok, uint64_t *nwords) {
    const uint64_t RANGE = 1ull << 30;
    uint64_t tt = 0, tw = 0;
    for (uint64_t lo = 0; lo < n; lo += RANGE) {
      const uint64_t hi = std::min(n, lo + RANGE);
      const uint64_t n_tiles = ((hi + 15) / 16 - lo / 16 + ENC_THREADS - 1) / ENC_THREADS;
      if (desc_.size() < n_tiles + 1) desc_.alloc(n_tiles + 1);  // [n_tiles] = {tile counter, fallback flag}
      SWB_CUDA(cudaMemsetAsync(desc_.get(), 0, (n_tiles + 1) * 8, stream_));
      unsigned int *ctr = reinterpret_cast<unsigned int *>(desc_.get() + n_tiles);
      const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)sms_ * 6);  // 30 KB of shared memory and 40 registers per thread: six blocks per SM
      enc_fused<<<grid, ENC_THREADS, 0, stream_>>>(d_text, n, lo, hi, tbl_, memo_, d_bmap_.get(), desc_.get(), ctr, ctr + 1, tok_base + tt,
                                                   word_base + tw, d_out, cap_ids, d_word_ntok, cap_words, neg_id_);
      launches++;
      SWB_CUDA(cudaGetLastError());
      unsigned long long h[2];
      SWB_CUDA(cudaMemcpyAsync(&h[0], desc_.get() + n_tiles - 1, 8, cudaMemcpyDeviceToHost, stream_));
      SWB_CUDA(cudaMemcpyAsync(&h[1], desc_.get() + n_tiles, 8, cudaMemcpyDeviceToHost, stream_));
      SWB_CUDA(cudaStreamSynchronize(stream_));
      if (h[1] >> 32) return false;  // a word longer than ENC_SHORT bytes
      tt += (h[0] >> 31) & 0x7FFFFFFFull; tw += h[0] & 0x7FFFFFFFull;
    }
    if (tok_base + tt > cap_ids) throw Error("swb_encode: output capacity too small");
    if (d_word_ntok && word_base + tw > cap_words) throw Error("swb_encode: word_ntok capacity too small");
    *ntok = tt; *nwords = tw;
    return true;
  }
  // Encodes one device-resident piece [d_text, d_text+n) that is 16-byte aligned and padded with 16 ' '.
  // Appends to d_out at tok_base / d_word_ntok at word_base. Returns (tokens, words) of the piece.
  void encode_piece(const uint8_t *d_text, uint64_t n, int32_t *d_out, uint64_t cap_ids, uint64_t tok_base,
                    uint32_t *d_word_ntok, uint64_t cap_words, uint64_t word_base, uint64_t *ntok, uint64_t *nwords) {
    *ntok = 0; *nwords = 0;
    if (n == 0) return;
    if (!no_fused() && encode_fused(d_text, n, d_out, cap_ids, tok_base, d_word_ntok, cap_words, word_base, ntok, nwords)) return;
    general_pieces++;
    const uint64_t n_tiles = (n + ENC_TILE - 1) / ENC_TILE;
    if (tmp_.size() < n + 16) tmp_.alloc(n + 16);
    if (tile_tok_.size() < n_tiles + 1) {
      tile_tok_.alloc(n_tiles + 1); tile_words_.alloc(n_tiles + 1);
      tile_tok_off_.alloc(n_tiles + 1); tile_word_off_.alloc(n_tiles + 1);
    }
    SWB_CUDA(cudaMemsetAsync(tile_tok_.get() + n_tiles, 0, 4, stream_));
    SWB_CUDA(cudaMemsetAsync(tile_words_.get() + n_tiles, 0, 4, stream_));
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)sms_ * 4);
    enc_words<<<grid, ENC_THREADS, 0, stream_>>>(d_text, n, tbl_, memo_, d_bmap_.get(), tmp_.get(), tile_tok_.get(), tile_words_.get());
    launches++;
    SWB_CUDA(cudaGetLastError());
    scan(tile_tok_.get(), tile_tok_off_.get(), n_tiles + 1);
    scan(tile_words_.get(), tile_word_off_.get(), n_tiles + 1);
    unsigned long long totals[2];
    SWB_CUDA(cudaMemcpyAsync(&totals[0], tile_tok_off_.get() + n_tiles, 8, cudaMemcpyDeviceToHost, stream_));
    SWB_CUDA(cudaMemcpyAsync(&totals[1], tile_word_off_.get() + n_tiles, 8, cudaMemcpyDeviceToHost, stream_));
    SWB_CUDA(cudaStreamSynchronize(stream_));
    *ntok = totals[0]; *nwords = totals[1];
    if (tok_base + totals[0] > cap_ids) throw Error("swb_encode: output capacity too small");
    if (d_word_ntok && word_base + totals[1] > cap_words) throw Error("swb_encode: word_ntok capacity too small");
    enc_gather<<<grid, ENC_THREADS, 0, stream_>>>(d_text, n, tmp_.get(), tile_tok_off_.get(), tile_word_off_.get(), tok_base,
                                                  word_base, d_out, cap_ids, d_word_ntok, cap_words, neg_id_);
    launches++;
    SWB_CUDA(cudaGetLastError());
  }

  // host text -> host ids, streamed through the device in pieces that end on a delimiter. Three streams, two
  // buffers each way: while piece i is encoded, piece i+1 crosses PCIe inbound and the ids of piece i-1 outbound
  // (pinned host memory makes the copies asynchronous; pageable memory still works, serialised by the driver).
  int64_t encode_host(const uint8_t *text, uint64_t n, int32_t *out, uint64_t cap_ids, uint32_t *word_ntok,
                      uint64_t cap_words, size_t *n_words) {
    ensure_device();
    const char *pe = getenv("SWB_ENCODE_PIECE");  // (test switch: bytes per piece)
    const uint64_t pv = pe ? strtoull(pe, nullptr, 10) : 0;
    const uint64_t PIECE = pv >= 4096 ? pv : (64ull << 20);
    if (!s_in_) {
      SWB_CUDA(cudaStreamCreateWithFlags(&s_in_, cudaStreamNonBlocking));
      SWB_CUDA(cudaStreamCreateWithFlags(&s_out_, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        SWB_CUDA(cudaEventCreateWithFlags(&ev_in_[i], cudaEventDisableTiming));
        SWB_CUDA(cudaEventCreateWithFlags(&ev_gather_[i], cudaEventDisableTiming));
        SWB_CUDA(cudaEventCreateWithFlags(&ev_out_[i], cudaEventDisableTiming));
      }
    }
    // piece boundaries (cut on a delimiter so that no word is split)
    std::vector<uint64_t> cut{0};
    uint64_t max_len = 0;
    while (cut.back() < n) {
      const uint64_t pos = cut.back();
      uint64_t end = std::min<uint64_t>(n, pos + PIECE);
      if (end < n) {
        uint64_t e = end;
        while (e > pos && !is_delim(text[e - 1])) --e;
        if (e == pos) { e = end; while (e < n && !is_delim(text[e])) ++e; }
        end = e;
      }
      max_len = std::max(max_len, end - pos);
      cut.push_back(end);
    }
    const size_t P = cut.size() - 1;
    for (int b = 0; b < 2; b++) {
      if (p_text_[b].size() < max_len + 64) p_text_[b].alloc(max_len + 64);
      if (p_out_[b].size() < max_len) p_out_[b].alloc(std::max<uint64_t>(max_len, 1));
      if (word_ntok && p_wn_[b].size() < max_len / 2 + 1) p_wn_[b].alloc(max_len / 2 + 1);
    }
    auto copy_in = [&](size_t i) {
      const int b = (int)(i & 1);
      const uint64_t len = cut[i + 1] - cut[i];
      if (i >= 2) SWB_CUDA(cudaStreamWaitEvent(s_in_, ev_gather_[b], 0));  // piece i-2 has been read out of this buffer
      SWB_CUDA(cudaMemcpyAsync(p_text_[b].get(), text + cut[i], len, cudaMemcpyHostToDevice, s_in_));
      SWB_CUDA(cudaMemsetAsync(p_text_[b].get() + len, ' ', 64, s_in_));
      SWB_CUDA(cudaEventRecord(ev_in_[b], s_in_));
    };
    uint64_t tok_total = 0, word_total = 0;
    SWB_CUDA(cudaStreamSynchronize(stream_));  // (buffers may have been (re)allocated from memory this stream used)
    if (P) copy_in(0);
    for (size_t i = 0; i < P; i++) {
      const int b = (int)(i & 1);
      const uint64_t len = cut[i + 1] - cut[i];
      if (i + 1 < P) copy_in(i + 1);
      SWB_CUDA(cudaStreamWaitEvent(stream_, ev_in_[b], 0));
      if (i >= 2) SWB_CUDA(cudaStreamWaitEvent(stream_, ev_out_[b], 0));  // the ids of piece i-2 have left this buffer
      uint64_t nt = 0, nw = 0;
      encode_piece(p_text_[b].get(), len, p_out_[b].get(), len, 0, word_ntok ? p_wn_[b].get() : nullptr, len / 2 + 1, 0, &nt, &nw);
      SWB_CUDA(cudaEventRecord(ev_gather_[b], stream_));
      if (tok_total + nt > cap_ids) { cudaDeviceSynchronize(); throw Error("swb_encode: output capacity too small"); }
      if (word_ntok && word_total + nw > cap_words) { cudaDeviceSynchronize(); throw Error("swb_encode: word_ntok capacity too small"); }
      SWB_CUDA(cudaStreamWaitEvent(s_out_, ev_gather_[b], 0));
      if (nt) SWB_CUDA(cudaMemcpyAsync(out + tok_total, p_out_[b].get(), nt * 4, cudaMemcpyDeviceToHost, s_out_));
      if (word_ntok && nw) SWB_CUDA(cudaMemcpyAsync(word_ntok + word_total, p_wn_[b].get(), nw * 4, cudaMemcpyDeviceToHost, s_out_));
      SWB_CUDA(cudaEventRecord(ev_out_[b], s_out_));
      tok_total += nt; word_total += nw;
    }
    SWB_CUDA(cudaStreamSynchronize(stream_));
    SWB_CUDA(cudaStreamSynchronize(s_out_));
    SWB_CUDA(cudaStreamSynchronize(s_in_));
    if (n_words) *n_words = word_total;
    return (int64_t)tok_total;
  }

  // device text -> device ids (the text is copied once into an aligned, padded buffer)
  int64_t encode_device(const uint8_t *d_text_in, uint64_t n, int32_t *d_out, uint64_t cap_ids, uint32_t *d_word_ntok,
                        uint64_t cap_words, size_t *n_words) {
    ensure_device();
    if (dtext_.size() < n + 64) dtext_.alloc(n + 64);
    if (n) SWB_CUDA(cudaMemcpyAsync(dtext_.get(), d_text_in, n, cudaMemcpyDeviceToDevice, stream_));
    SWB_CUDA(cudaMemsetAsync(dtext_.get() + n, ' ', 64, stream_));
    uint64_t nt = 0, nw = 0;
    encode_piece(dtext_.get(), n, d_out, cap_ids, 0, d_word_ntok, cap_words, 0, &nt, &nw);
    SWB_CUDA(cudaStreamSynchronize(stream_));
    if (n_words) *n_words = nw;
    return (int64_t)nt;
  }

  size_t decode(const int32_t *ids, size_t n, uint8_t *out, size_t cap) const {
    size_t pos = 0;
    const size_t T = tok_off_.size() - 1;
    for (size_t i = 0; i < n; i++) {
      if (ids[i] < 0 || (size_t)ids[i] >= T) continue;
      const size_t a = tok_off_[ids[i]], b = tok_off_[ids[i] + 1];
      for (size_t k = a; k < b; k++) { if (out && pos < cap) out[pos] = tok_bytes_[k]; pos++; }
    }
    return pos;
  }

 private:
  void scan(const uint32_t *in, unsigned long long *out, uint64_t count) {
    cub::TransformInputIterator<unsigned long long, CastU32ToU64, const uint32_t *> it(in, CastU32ToU64());
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int64_t)count, stream_);
    if (scan_tmp_.size() < bytes) scan_tmp_.alloc(bytes);
    SWB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp_.get(), bytes, it, out, (int64_t)count, stream_));
    launches += 2;
  }

  std::vector<int32_t> merges_;
  int32_t bmap_[256], dev_bmap_[256];
  int32_t neg_id_ = -1;
  std::vector<uint8_t> tok_bytes_;
  std::vector<size_t> tok_off_;
  cudaStream_t stream_ = nullptr;
  // encode_host pipeline: copy-in / copy-out streams, two device buffers per direction
  cudaStream_t s_in_ = nullptr, s_out_ = nullptr;
  cudaEvent_t ev_in_[2] = {}, ev_gather_[2] = {}, ev_out_[2] = {};
  DevBuf<uint8_t> p_text_[2];
  DevBuf<int32_t> p_out_[2];
  DevBuf<uint32_t> p_wn_[2];
  int sms_ = 148;
  DevBuf<RankSlot> slots_;
  DevBuf<MemoSlot> memo_slots_;
  MemoDev memo_{};
  DevBuf<int32_t> d_bmap_;
  RankTableDev tbl_{};
  DevBuf<int32_t> tmp_;
  DevBuf<uint32_t> tile_tok_, tile_words_;
  DevBuf<unsigned long long> tile_tok_off_, tile_word_off_;
  DevBuf<uint8_t> scan_tmp_, dtext_;
  DevBuf<unsigned long long> desc_;
};

}  // namespace swb
