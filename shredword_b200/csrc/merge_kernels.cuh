// merge_kernels.cuh -- the per-merge hot path over the row-packed symbol stream.
//
// Replaces reference bpe_count_bigrams' first pass (reference csrc/bpe/bpe.cpp:329-350), the scan +
// splice + delta bookkeeping of bpe_merge_batch (bpe.cpp:437-483, FreqChangeMap bpe.cpp:10-46) and
// bpe_save's token histogram (bpe.cpp:704-712).
//
// Stream layout: rows of ROW=128 int32. A word is [~local_index][sym ...][PAD ...]; it never
// straddles a row, so one warp owns whole words: lane l holds symbols 4l..4l+3 of a row (one 16-byte
// coalesced load), adjacency is a shuffle, and a row without the pair (a, b) -- almost all of them --
// costs one load, one shuffle, a few compares and one vote. Rows with a match take the slow path:
// the row is staged in shared memory and every word is rewritten left to right by the lane that holds
// its header, which reproduces the reference's sequential semantics exactly (already-merged left
// neighbour, not-yet-merged right neighbour, greedy self-pairs) and emits the four signed deltas of
// each match into a global open-addressing pair table together with a first-touch key.
#pragma once

#include "device_util.cuh"
#include "host_core.hpp"

namespace swb {

constexpr uint64_t PT_EMPTY = ~0ull;

struct PairTableDev {
  unsigned long long *keys;    // (first << 32) | second, PT_EMPTY when free
  unsigned long long *val;     // net delta (two's complement) or frequency
  unsigned long long *minkey;  // smallest first-touch key
  uint4 *touched;              // per claim: {slot, key lo, key hi, 0}: slots claimed since the last emit
  unsigned int *n_touched;
  unsigned int *flags;         // bit 0: table full
  unsigned int *done_blocks;   // emit kernel bookkeeping
  uint32_t mask;
  // device-table mode with a negative unk_id: the reference's delta-map key sign-extends a negative
  // `second` over `first` (bpe.cpp:456-457), i.e. (X, unk) collapses to (-1, unk) for every X.
  int32_t canon_on, canon_first;
  // device-table mode: the frequency-table slot of a newly claimed pair is prefetched into L2 here, so
  // that the emit tail's compare-and-swap on it does not pay an HBM miss on the critical path
  struct GSlotRef { void *slots; uint32_t mask; } gpf;
};

// Device-resident frequency table (single-GPU mode): pair -> current frequency, for every pair ever
// touched. With it the emit step applies the deltas itself (clamp at zero on the net delta, reference
// bpe.cpp:500-509) and only pairs that are, or were, at or above min_pair_freq travel to the host: the
// Zipf tail of rare pairs (most records) never crosses PCIe and never enters the host's table.
// stamp = creation order (operation index, delta-map bucket, inverted first-touch key), kept so that the
// reference's creation-ordered table can be rebuilt exactly on the host when it is needed.
struct __align__(32) GSlot { unsigned long long key, freq, stamp_hi, stamp_lo; };  // one 32-byte sector per pair
struct GlobalTableDev {
  GSlot *slots;
  unsigned int *n_used;
  unsigned int *flags;  // bit 0: above 50 % load, the host grows the table before the next operation
  uint32_t mask;
};
__device__ __forceinline__ uint32_t gt_home(const GlobalTableDev &g, unsigned long long k) {
  return (uint32_t)dmix64(k + 0x632BE59BD9B4E019ull) & g.mask;
}
__device__ __forceinline__ void gt_prefetch(const GlobalTableDev &g, unsigned long long k) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(g.slots + gt_home(g, k)));
}
// `inserted` is incremented when a new slot is claimed; the caller adds its total to n_used once
// (gt_account), instead of one same-address atomic per insertion.
__device__ __forceinline__ uint32_t gt_upsert(const GlobalTableDev &g, unsigned long long k, unsigned long long sh,
                                              unsigned long long sl, unsigned int &inserted) {
  uint32_t h = gt_home(g, k);
  for (;;) {
    const unsigned long long cur = atomicCAS(&g.slots[h].key, PT_EMPTY, k);
    if (cur == PT_EMPTY) {
      g.slots[h].stamp_hi = sh; g.slots[h].stamp_lo = sl;
      inserted++;
      return h;
    }
    if (cur == k) return h;
    h = (h + 1) & g.mask;
  }
}

// frequency of pair k, ~0 when the pair is not in the table (read through L2: other SMs update the table)
__device__ __forceinline__ unsigned long long gt_find_freq(const GlobalTableDev &g, unsigned long long k) {
  uint32_t h = gt_home(g, k);
#pragma unroll 1
  for (uint32_t probe = 0; probe <= g.mask; probe++) {
    const ulonglong2 gs = __ldcg(reinterpret_cast<const ulonglong2 *>(g.slots + h));
    if (gs.x == k) return gs.y;
    if (gs.x == PT_EMPTY) break;
    h = (h + 1) & g.mask;
  }
  return ~0ull;
}

__device__ __forceinline__ void gt_account(const GlobalTableDev &g, unsigned int inserted) {
  if (inserted && atomicAdd(g.n_used, inserted) + inserted >= (g.mask >> 1)) atomicOr(g.flags, 1u);
}

__global__ void gt_clear(GlobalTableDev g) {
  const uint64_t cap = (uint64_t)g.mask + 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
    g.slots[i] = GSlot{PT_EMPTY, 0, 0, 0};
  if (blockIdx.x == 0 && threadIdx.x == 0) { *g.n_used = 0; *g.flags = 0; }
}
__global__ void gt_rehash(GlobalTableDev src, GlobalTableDev dst) {
  const uint64_t cap = (uint64_t)src.mask + 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    const GSlot e = src.slots[i];
    if (e.key == PT_EMPTY) continue;
    unsigned int ins = 0;
    const uint32_t h = gt_upsert(dst, e.key, e.stamp_hi, e.stamp_lo, ins);
    dst.slots[h].freq = e.freq;
    gt_account(dst, ins);
  }
}
// all entries -> host (conversion to the host-resident table): 5 words per entry
__global__ void gt_dump(GlobalTableDev g, unsigned long long *__restrict__ out, unsigned int *cursor) {
  const uint64_t cap = (uint64_t)g.mask + 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    const GSlot e = g.slots[i];
    if (e.key == PT_EMPTY) continue;
    const unsigned int j = atomicAdd(cursor, 1u);
    out[5ull * j] = e.key; out[5ull * j + 1] = e.freq; out[5ull * j + 2] = e.stamp_hi; out[5ull * j + 3] = e.stamp_lo;
    out[5ull * j + 4] = 0;
  }
}

constexpr int HDR_WORDS = 32;

// ---- birth log: pair -> words that can contain it
// An adjacent pair (x, y) with max(x, y) >= 256 comes into existence only while the NEWER of its two
// tokens is being created: as (L, new) or (new, R) of a match of that merge (reference bpe.cpp:459-470).
// Every merge therefore appends, for each positive delta it emits, one entry {other symbol, side, word,
// location of the word's header} to its own contiguous log -- once per (pair, word). A later merge of
// (a, b) reads the log of the merge that created max(a, b), keeps the entries with the right neighbour
// and side, and visits exactly those words, one THREAD per word (a superset of the words that still hold
// the pair: occurrences are only ever destroyed afterwards, and a word never moves).
// Pairs of two initial symbols (< 256) have no log and use the row-signature scan.
struct BirthLogDev {
  uint4 *ent;            // {other symbol, side << 31 | word index, header location lo, hi}; side 0: (other, new), 1: (new, other)
  unsigned int *cursor;  // entries appended so far
  unsigned int *start;   // start[m] .. start[m+1]: the log of merge m (token 256 + m)
  unsigned int *flags;   // bit 0: capacity exceeded (internal sizing error)
  uint32_t cap;
  uint32_t m_cur;        // index of the merge being performed; logs 0 .. m_cur-1 are complete
  // Occurrence index of the pairs of two INITIAL symbols (both < 256), which have no birth log: built once after the
  // load, same entry format, one contiguous range per pair: ip_ent[ip_start[x * 256 + y] .. ip_start[x * 256 + y + 1])
  // lists every word that held (x, y) when it was loaded (once per word). Such a pair can only lose occurrences later
  // (a merge creates adjacencies around its new token only), so the list stays a superset. nullptr: not built.
  const uint4 *ip_ent;
  const unsigned int *ip_start;
  uint32_t ip_local_max;  // longest list the resident kernel's leader cluster takes alone (every entry is a candidate word); longer: whole-grid row scan
};
constexpr uint32_t IP_PAIRS = 256u * 256u;
// true: (a, b) is a pair of two initial symbols with an occurrence index
__device__ __forceinline__ bool ip_lookup(const BirthLogDev &lg, int32_t a, int32_t b) {
  return lg.ip_start != nullptr && (uint32_t)a < 256u && (uint32_t)b < 256u;
}
// Builds the index, one thread per (row) word. FILL = false: counts per pair into cnt[]; FILL = true: cnt[] restarts
// from zero as the per-pair cursor and the entries are written from start[] on.
template <bool FILL>
__global__ void __launch_bounds__(256)
ip_index_words(const int32_t *__restrict__ flat_rows, const uint64_t *__restrict__ wloc, const uint32_t *__restrict__ long_index,
               uint32_t W, unsigned int *__restrict__ cnt, const unsigned int *__restrict__ start, uint4 *__restrict__ ent) {
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
    if (long_index[w] != 0xFFFFFFFFu) continue;
    const uint64_t hloc = wloc[w];
    const int32_t *p = flat_rows + hloc + 1;
    const int room = ROW - 1 - (int)(hloc & (ROW - 1));  // slots between the header and the end of the row
    for (int i = 0; i + 1 < room; i++) {
      const int32_t x = p[i], y = p[i + 1];
      if (x < 0 || y < 0) break;  // next header / padding: the word ends here
      if ((uint32_t)x >= 256u || (uint32_t)y >= 256u) continue;
      bool seen = false;  // once per (pair, word): the visiting thread rewrites every occurrence in the word
      for (int j = 0; j < i && !seen; j++) seen = p[j] == x && p[j + 1] == y;
      if (seen) continue;
      const uint32_t k = (uint32_t)x * 256u + (uint32_t)y;
      const unsigned int at = atomicAdd(&cnt[k], 1u);
      if (FILL) ent[start[k] + at] = make_uint4((uint32_t)x, w, (uint32_t)hloc, (uint32_t)(hloc >> 32));  // {other, side 0 | word, header location}
    }
  }
}
__device__ __forceinline__ uint4 log_entry(uint32_t other, bool right_side, uint32_t wi, uint64_t hloc) {
  return make_uint4(other, (right_side ? 0x80000000u : 0u) | wi, (uint32_t)hloc, (uint32_t)(hloc >> 32));
}
// which log, neighbour and side hold the births of pair (a, b); false: no log (two initial symbols, or ids
// that are not tokens of this stream)
__device__ __forceinline__ bool log_lookup(const BirthLogDev &lg, int32_t a, int32_t b, uint32_t &merge, uint32_t &other, uint32_t &side) {
  const int32_t newer = a > b ? a : b;
  if (lg.ent == nullptr || newer < 256 || (uint32_t)(newer - 256) >= lg.m_cur) return false;
  merge = (uint32_t)(newer - 256);
  if (b == newer) { other = (uint32_t)a; side = 0u; }     // (a, b) was born as (L, new)
  else { other = (uint32_t)b; side = 0x80000000u; }       // ... or as (new, R)
  return true;
}

// parameters of an emit in device-table mode (mode 0 = host-resident table: plain records)
struct EmitMode {
  int mode;                       // 0 host records (delta), 1 device table: merge, 2 device table: count
  unsigned long long min_freq;
  unsigned long long merged_key;  // mode 1: the pair being merged (skipped, then zeroed)
  unsigned long long stamp_base;  // operation index << 10
  int32_t neg_unk_bucket;         // >= 0: a pair whose second is UNK_CODE lives in this delta-map bucket (negative unk_id)
  GlobalTableDev g;
  unsigned int fused_max;  // touched pairs the single-block tail takes; more -> flag 8, the host runs the full-grid pt_emit
  BirthLogDev log;         // ent == nullptr: no birth log
};
__device__ __forceinline__ unsigned long long delta_bucket(const EmitMode &em, unsigned long long k) {
  if (em.neg_unk_bucket >= 0 && (uint32_t)k == (uint32_t)UNK_CODE) return (unsigned long long)em.neg_unk_bucket;
  return k & 1023ull;
}

__device__ __forceinline__ void pt_add(const PairTableDev &t, int32_t a, int32_t b, long long delta, uint64_t key) {
  if (t.canon_on && b == UNK_CODE) a = t.canon_first;
  const unsigned long long k = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b;
  uint32_t h = (uint32_t)dmix64(k) & t.mask;
#pragma unroll 1
  for (uint32_t probe = 0; probe <= t.mask; probe++) {
    // CAS straight away (one L2 round trip): the table is empty at the start of every merge, so the
    // common case is a first touch
    unsigned long long cur = atomicCAS(&t.keys[h], PT_EMPTY, k);
    // claims of the lanes that are converged here share ONE atomic on the list cursor (same-address global
    // atomics serialise in L2 at tens of ns each; a merge makes hundreds of claims)
    const unsigned int conv = __activemask();
    const unsigned int claimers = __ballot_sync(conv, cur == PT_EMPTY);
    if (cur == PT_EMPTY) {
      const int leader = __ffs(claimers) - 1;
      const int lane_id = threadIdx.x & 31;
      unsigned int base = 0;
      if (lane_id == leader) base = atomicAdd(t.n_touched, (unsigned int)__popc(claimers));
      base = __shfl_sync(claimers, base, leader);
      const unsigned int idx = base + __popc(claimers & ((1u << lane_id) - 1u));
      t.touched[idx] = make_uint4(h, (uint32_t)k, (uint32_t)(k >> 32), 0u);
      if (t.gpf.slots)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<char *>(t.gpf.slots) +
                                                     32ull * ((uint32_t)dmix64(k + 0x632BE59BD9B4E019ull) & t.gpf.mask)));
      if (idx >= (t.mask >> 1)) atomicOr(t.flags, 1u);  // past 50 % load: the host grows the table and reruns
      cur = k;
    }
    if (cur == k) {
      atomicAdd(&t.val[h], (unsigned long long)delta);
      atomicMin(&t.minkey[h], (unsigned long long)key);
      return;
    }
    h = (h + 1) & t.mask;
    if (probe > 64 && *(volatile unsigned int *)t.flags) return;  // overflow already declared: do not crawl
  }
  atomicOr(t.flags, 1u);
}

__global__ void pt_clear(PairTableDev t) {
  const uint64_t cap = (uint64_t)t.mask + 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    t.keys[i] = PT_EMPTY; t.val[i] = 0; t.minkey[i] = ~0ull;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { *t.n_touched = 0; *t.flags = 0; *t.done_blocks = 0; }
}

// touched slots -> records in mapped host memory; frees the slots. The last block to finish publishes
// an 8-word header and resets the counters, so the next merge needs no extra memset on the critical path.
//   hdr[0] = seq, [1] = number of records, [2] = flags, [3] = removed symbols of this merge,
//   hdr[4] = XOR of all record words, [5] = SUM of all record words, [6] = header check word, [7] = seq
// The host polls the header instead of synchronising the stream. No system-scope fence sits between the
// record writes and the header (each one costs a PCIe round trip, ~2.3 us on this box): the header and
// the records validate themselves -- the host accepts them only when the check word and the XOR/SUM of
// the records it reads match, and keeps polling otherwise (the bytes are still in flight).
// flags: 1 = table full, 4 = more records than out_cap, 8 = records not emitted (fused tail only),
//        16 = device frequency table past 50 % load (the host grows it before the next merge).
constexpr unsigned long long HDR_MAGIC = 0x9E3779B97F4A7C15ull;
__host__ __device__ __forceinline__ unsigned long long hdr_check(unsigned long long seq, unsigned long long n,
                                                                 unsigned long long flags, unsigned long long removed,
                                                                 unsigned long long x, unsigned long long sm) {
  return (seq * HDR_MAGIC) ^ (n + 0x1234567ull) ^ (flags << 48) ^ ((flags >> 32) * 0x9E3779B1ull) ^ (removed * 31ull) ^ x ^ (sm << 1 | sm >> 63);
}

// `out` may be a shared-memory stage of `stage_cap` records backed by `direct` (the final destination):
// records that do not fit the stage are written straight to their final place.
__device__ __forceinline__ void rec_out(Rec *__restrict__ out, size_t out_cap, unsigned int j, unsigned long long k,
                                        long long v, unsigned long long mk, unsigned long long &cx, unsigned long long &cs,
                                        Rec *__restrict__ direct = nullptr, unsigned int stage_cap = 0xFFFFFFFFu) {
  if (j >= out_cap) return;
  if (j >= stage_cap) out = direct;
  Rec r;
  r.first = (int32_t)(k >> 32); r.second = (int32_t)(k & 0xFFFFFFFFu);
  r.delta = v; r.key = (long long)mk;
  out[j] = r;
  cx ^= (unsigned long long)r.first ^ (unsigned long long)r.second ^ (unsigned long long)r.delta ^ (unsigned long long)r.key;
  cs += (unsigned long long)r.first + 3ull * (unsigned long long)r.second + 5ull * (unsigned long long)r.delta +
        7ull * (unsigned long long)r.key;
}

// Emits this thread's share of the touched slots and frees them.
//   mode 0: every touched pair -> (pair, net delta, first-touch key), record i at out[i]
//   mode 1: deltas are applied to the device frequency table; a record (pair, NEW frequency, key) is
//           written (compacted through *out_count) only if the old or the new frequency reaches min_freq
//   mode 2: counts are added to the device frequency table; records only for pairs >= min_freq
__device__ __forceinline__ void pt_emit_range(const PairTableDev &t, const EmitMode &em, Rec *__restrict__ out, size_t out_cap,
                                              unsigned int n, unsigned int first, unsigned int stride,
                                              unsigned int *out_count, unsigned long long &cx, unsigned long long &cs,
                                              unsigned int &inserted, Rec *__restrict__ direct = nullptr,
                                              unsigned int stage_cap = 0xFFFFFFFFu, unsigned long long *maxpush = nullptr /* this thread's largest new frequency >= min_freq */) {
#pragma unroll 1
  for (unsigned int i = first; i < n; i += stride) {
    const uint4 tc = __ldcg(&t.touched[i]);
    const unsigned int h = tc.x;
    const unsigned long long k = ((unsigned long long)tc.z << 32) | tc.y;
    // the frequency-table slot and the delta-table slot are fetched together (one round trip); the slot was
    // prefetched into L2 when the pair was claimed
    uint32_t g = 0;
    ulonglong2 gs = make_ulonglong2(0, 0);  // {key, freq}
    const bool dev = em.mode != 0 && !(em.mode == 1 && k == em.merged_key);
    if (dev) { g = gt_home(em.g, k); gs = __ldcg(reinterpret_cast<const ulonglong2 *>(em.g.slots + g)); }
    const long long d = (long long)__ldcg(&t.val[h]);
    const unsigned long long mk = __ldcg(&t.minkey[h]);
    t.keys[h] = PT_EMPTY; t.val[h] = 0; t.minkey[h] = ~0ull;
    if (em.mode == 0) { rec_out(out, out_cap, i, k, d, mk, cx, cs, direct, stage_cap); continue; }
    if (!dev) continue;  // the merged pair itself: reference bpe.cpp:494-496
    unsigned long long old;
    if (gs.x == k) old = gs.y;  // common case: the pair is already in the table, at its home slot
    else {
      const unsigned long long bucket = em.mode == 1 ? delta_bucket(em, k) : 0ull;
      g = gt_upsert(em.g, k, em.stamp_base | bucket, em.mode == 1 ? ~mk : mk, inserted);
      old = em.g.slots[g].freq;
    }
    unsigned long long nw;
    if (d < 0) { const unsigned long long ad = (unsigned long long)(-d); nw = old >= ad ? old - ad : 0ull; }
    else nw = old + (unsigned long long)d;
    em.g.slots[g].freq = nw;
    if (maxpush && nw >= em.min_freq && nw > *maxpush) *maxpush = nw;  // the host pushes this pair (reference bpe.cpp:512-515)
    if (old >= em.min_freq || nw >= em.min_freq)
      rec_out(out, out_cap, atomicAdd(out_count, 1u), k, (long long)nw, mk, cx, cs, direct, stage_cap);
  }
}
// block-wide XOR / SUM of per-thread checksums (all threads of the block must call)
__device__ __forceinline__ void block_checksum(unsigned long long &cx, unsigned long long &cs, unsigned long long *sh /* [2*32] */) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { cx ^= __shfl_down_sync(0xffffffffu, cx, d); cs += __shfl_down_sync(0xffffffffu, cs, d); }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { sh[w] = cx; sh[32 + w] = cs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long x = 0, sm = 0;
    for (int i = 0; i < nw; i++) { x ^= sh[i]; sm += sh[32 + i]; }
    cx = x; cs = sm;
  }
}
__device__ __forceinline__ unsigned long long pt_publish(const PairTableDev &t, unsigned int n, size_t out_cap, unsigned int extra_flags,
                                           volatile unsigned long long *out_hdr, unsigned long long *removed,
                                           unsigned long long seq, unsigned long long cx, unsigned long long cs,
                                           const BirthLogDev *lg = nullptr, bool keep_state = false) {
  unsigned long long flags = __ldcg(t.flags) | (n > out_cap ? 4u : 0u) | extra_flags;
  const unsigned long long rem = removed ? __ldcg(removed) : 0ull;
  if (lg && lg->ent) {  // this merge's birth log ends here; the host keeps the ranges too (high half of the flags word)
    const unsigned int cur = __ldcg(lg->cursor);
    lg->start[lg->m_cur + 1] = cur;
    if (__ldcg(lg->flags)) flags |= 32u;
    flags |= (unsigned long long)cur << 32;
  }
  // The counters go back to zero BEFORE the header can be seen: whoever sees the header may start the next operation on this
  // table at once (the host answers within microseconds), and a reset that lands after the first blocks of the next
  // operation have counted themselves in would leave that operation without a last block -- and its result unpublished.
  if (removed && !(extra_flags & 8u) && !keep_state) *removed = 0;
  if (!(extra_flags & 8u) && !keep_state) { *t.n_touched = 0; *t.flags = 0; }
  *t.done_blocks = 0;
  __threadfence();
  out_hdr[1] = n; out_hdr[2] = flags; out_hdr[3] = rem; out_hdr[4] = cx; out_hdr[5] = cs;
  out_hdr[6] = hdr_check(seq, n, flags, rem, cx, cs);
  out_hdr[0] = seq; out_hdr[7] = seq;
  return flags;
}

// full-grid emit (more records than the fused tail takes, long words present, or the count pass): every
// block writes its share and its partial checksum; the last block to finish adds the partials up and publishes.
__global__ void __launch_bounds__(256)
pt_emit(PairTableDev t, EmitMode em, Rec *__restrict__ out, size_t out_cap, unsigned long long *__restrict__ out_hdr,
        unsigned long long *removed, unsigned long long seq, unsigned long long *__restrict__ partial /* [2*gridDim.x] */,
        unsigned int *out_count /* global, zero on entry */) {
  __shared__ bool is_last;
  __shared__ unsigned long long sh[64];
  const unsigned int n = __ldcg(t.n_touched);
  unsigned long long cx = 0, cs = 0;
  unsigned int inserted = 0;
  pt_emit_range(t, em, out, out_cap, n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, out_count, cx, cs, inserted);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) inserted += __shfl_down_sync(0xffffffffu, inserted, d);
  if ((threadIdx.x & 31) == 0 && em.mode != 0) gt_account(em.g, inserted);
  block_checksum(cx, cs, sh);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = cx; partial[2 * blockIdx.x + 1] = cs;
    __threadfence();
    is_last = atomicAdd(t.done_blocks, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    unsigned long long x = 0, sm = 0;
    for (unsigned int i = 0; i < gridDim.x; i++) { x ^= __ldcg(&partial[2 * i]); sm += __ldcg(&partial[2 * i + 1]); }
    unsigned int n_out = n;
    if (em.mode != 0) {
      n_out = __ldcg(out_count);
      *out_count = 0;
      if (em.mode == 1) {
        unsigned int ins = 0;
        em.g.slots[gt_upsert(em.g, em.merged_key, em.stamp_base | delta_bucket(em, em.merged_key), 0ull, ins)].freq = 0;  // bpe.cpp:523
        gt_account(em.g, ins);
      }
    }
    const unsigned int gflag = (em.mode != 0 && __ldcg(em.g.flags)) ? 16u : 0u;  // frequency table past 50 % load
    pt_publish(t, n_out, out_cap, gflag, out_hdr, removed, seq, x, sm, &em.log);
  }
}

constexpr unsigned int FUSED_EMIT_MAX = 1024;  // touched pairs the single-block tail takes (more: full-grid pt_emit)
constexpr unsigned int STAGE_RECS = 384;        // records staged in shared memory before they cross PCIe  // above this the records are emitted by a full-grid pt_emit

// Tail run by ALL threads of the last block of a kernel (blockDim.x == 256): emits the touched pairs of
// table `t` (staged in shared memory, copied out with lane-consecutive 16-byte stores: a store to mapped
// host memory becomes a PCIe write, and scattered per-thread writes cost ~16x more transactions than full
// lines), applies them to the device frequency table when em.mode != 0, and publishes the header.
struct TailSmem { Rec *stage; unsigned long long *csum; unsigned int *count; };
__device__ __forceinline__ void fused_tail(const PairTableDev &t, const EmitMode &em, const TailSmem &ts, Rec *__restrict__ out,
                                           size_t out_cap, unsigned long long *__restrict__ out_hdr, unsigned long long *removed,
                                           unsigned long long seq, unsigned int extra_flags, unsigned long long *trace = nullptr,
                                           unsigned long long *tail_max = nullptr /* global, zeroed by the caller: largest frequency the host will push */,
                                           unsigned long long *tail_flags = nullptr /* global: the header's flag word */) {
  const long long tc0 = clock64();
  const unsigned int n = __ldcg(t.n_touched);
  const bool small = n <= em.fused_max;
  unsigned long long cx = 0, cs = 0;
  if (threadIdx.x == 0) *ts.count = 0;
  __syncthreads();
  unsigned int inserted = 0;
  const long long tc1 = clock64();
  unsigned long long mp = 0;
  if (small) pt_emit_range(t, em, ts.stage, out_cap, n, threadIdx.x, blockDim.x, ts.count, cx, cs, inserted, out, STAGE_RECS, tail_max ? &mp : nullptr);
  if (tail_max) {  // (ahead of the block barrier inside block_checksum, which is ahead of the publisher's fence)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_down_sync(0xffffffffu, mp, d); mp = o > mp ? o : mp; }
    if ((threadIdx.x & 31) == 0 && mp) atomicMax(tail_max, mp);
  }
  const long long tc2 = clock64();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) inserted += __shfl_down_sync(0xffffffffu, inserted, d);
  if ((threadIdx.x & 31) == 0 && em.mode != 0) gt_account(em.g, inserted);
  block_checksum(cx, cs, ts.csum);  // (contains a __syncthreads: the stage is complete after it)
  const unsigned int n_out = !small ? n : (em.mode == 0 ? n : *ts.count);
  const long long tc3 = clock64();
  if (small) {
    const uint4 *src = reinterpret_cast<const uint4 *>(ts.stage);
    uint4 *dst = reinterpret_cast<uint4 *>(out);
    const unsigned int chunks = 2u * (unsigned int)min(min((size_t)n_out, out_cap), (size_t)STAGE_RECS);
    for (unsigned int i = threadIdx.x; i < chunks; i += blockDim.x) dst[i] = src[i];
  }
  const long long tc4 = clock64();
  if (threadIdx.x == 0 && small && em.mode == 1) {
    unsigned int ins = 0;
    em.g.slots[gt_upsert(em.g, em.merged_key, em.stamp_base | delta_bucket(em, em.merged_key), 0ull, ins)].freq = 0;  // bpe.cpp:523
    gt_account(em.g, ins);
  }
  const long long tc5 = clock64();
  if (threadIdx.x == 0) {
    const unsigned int gflag = (em.mode != 0 && __ldcg(em.g.flags)) ? 16u : 0u;  // frequency table past 50 % load
#ifdef SWB_KERNEL_TRACE
    unsigned long long tr_emit_done; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr_emit_done));
    out_hdr[9] = tr_emit_done; out_hdr[10] = n;
    if (em.log.ent) {
      out_hdr[20] = em.log.flags[1]; out_hdr[22] = em.log.flags[2]; em.log.flags[1] = 0; em.log.flags[2] = 0;
      const int32_t ka = (int32_t)(em.merged_key >> 32), kb = (int32_t)(em.merged_key & 0xFFFFFFFFu), nw_ = ka > kb ? ka : kb;
      out_hdr[21] = (nw_ >= 256 && (uint32_t)(nw_ - 256) < em.log.m_cur) ? em.log.start[nw_ - 256 + 1] - em.log.start[nw_ - 256] : ~0ull;
    }
#endif
    const unsigned long long fl = pt_publish(t, n_out, out_cap, (small ? 0u : 8u) | gflag | extra_flags, out_hdr, removed, seq, cx, cs, &em.log);
    if (tail_flags) *(volatile unsigned long long *)tail_flags = fl;
    if (trace) {
      const long long tc6 = clock64();
      trace[9] += (unsigned long long)(tc1 - tc0); trace[10] += (unsigned long long)(tc2 - tc1); trace[11] += (unsigned long long)(tc3 - tc2);
      trace[12] += (unsigned long long)(tc4 - tc3); trace[13] += (unsigned long long)(tc5 - tc4); trace[14] += (unsigned long long)(tc6 - tc5);
      trace[15] += n;
    }
  }
}

struct StreamDev {
  int4 *rows;                 // n_rows * ROW int32
  uint32_t *sig;              // n_rows * SIG_WORDS: which symbol ids (hashed to 256 bits) a row may contain
  uint64_t n_rows;
  const unsigned long long *cnt;   // [W] word count, indexed by the (global, reference-order) word index
  // long words (more than ROW-1 symbols): CSR, one warp per word
  int32_t *long_syms;
  const uint64_t *long_off;
  uint32_t *long_len;         // live length
  const uint32_t *long_word;  // word index of long word j
  uint32_t n_long;
};

__device__ __forceinline__ uint64_t word_gwi(const StreamDev &, uint32_t wi) { return (uint64_t)wi; }

// ---------------------------------------------------------------- merge (a, b) -> new_id
constexpr int MERGE_THREADS = 256;
constexpr int MERGE_WARPS = MERGE_THREADS / 32;
constexpr int MATCH_CAP = 96;  // matches of one warp iteration awaiting emission (a row holds at most 64)

// One match of the pair inside a row, recorded by the lane that rewrites the word; the four signed
// deltas it stands for (reference bpe.cpp:453-470) are emitted afterwards, one lane per delta.
struct Match { int32_t L, R; uint32_t wi; uint32_t pos; unsigned long long cnt; uint32_t row, hpos; };  // L / R = -1: no such neighbour

// Sequential rewrite of one word living in shared memory at sm[p+1 ...]; p = header position.
__device__ __forceinline__ uint32_t merge_word_smem(int *sm, int p, Match *ml, unsigned int *n_match,
                                                    const unsigned long long *__restrict__ cnt, int32_t a, int32_t b,
                                                    int32_t new_id, uint32_t row) {
  const uint32_t wi = (uint32_t)(~sm[p]);
  unsigned long long c = 0;
  int r = p + 1, w = p + 1;
  uint32_t nmatch = 0;
  while (r < ROW) {
    const int x = sm[r];
    if (x < 0) break;
    if (x == a && r + 1 < ROW && sm[r + 1] == b) {
      Match m;
      m.L = (w > p + 1) ? sm[w - 1] : -1;                      // left neighbour: the already rewritten symbol
      m.R = (r + 2 < ROW && sm[r + 2] >= 0) ? sm[r + 2] : -1;  // right neighbour: not yet rewritten
      m.wi = wi; m.pos = (uint32_t)r; m.row = row; m.hpos = (uint32_t)p;
      if (nmatch == 0) c = __ldg(&cnt[wi]);  // issued here so that its latency hides behind the rest of the rewrite
      m.cnt = c;
      ml[atomicAdd(n_match, 1u)] = m;
      sm[w++] = new_id;
      r += 2;
      nmatch++;
    } else {
      if (w != r) sm[w] = x;
      w++; r++;
    }
  }
  for (int q = w; q < r; q++) sm[q] = PAD;
  return nmatch;
}

// Emits the four signed deltas of every recorded match, one lane per delta (reference bpe.cpp:453-470):
// the chain of dependent global atomics is paid once per warp, not once per matched row. The positive
// deltas -- the pairs that come into existence here -- also go to the birth log, once per (pair, row).
__device__ __forceinline__ void emit_matches(const Match *ml, unsigned int nm, int lane, const PairTableDev &t,
                                             const BirthLogDev &lg, int32_t a, int32_t b, int32_t new_id) {
  for (unsigned int i0 = 0; i0 < 4 * nm; i0 += 32) {  // (warp-uniform bounds: the log append below votes)
    const unsigned int i = i0 + lane;
    bool born = false;
    int nb = -1;
    uint32_t row = 0, slot = i & 3, wi = 0;
    uint64_t hloc = 0;
    if (i < 4 * nm) {
      const Match m = ml[i >> 2];
      nb = slot < 2 ? m.L : m.R;
      row = m.row; wi = m.wi; hloc = (uint64_t)m.row * ROW + m.hpos;
      if (nb >= 0) {
        const long long c = (long long)m.cnt;
        const uint64_t key = touch_key(m.wi, m.pos, slot);
        {  // (one call site: pt_add is large)
          const int32_t px = slot == 0 ? nb : slot == 1 ? nb : slot == 2 ? b : new_id;
          const int32_t py = slot == 0 ? a : slot == 1 ? new_id : nb;
          pt_add(t, px, py, (slot & 1) ? c : -c, key);
        }
        if (lg.ent && (slot & 1)) {
          born = true;  // unless an earlier match of the same word already logged this neighbour on this side
          for (int j = (int)(i >> 2) - 1; j >= 0 && ml[j].row == row; j--)
            if (ml[j].wi == m.wi && (slot == 1 ? ml[j].L : ml[j].R) == nb) { born = false; break; }
        }
      }
    }
    const unsigned int bm = __ballot_sync(0xffffffffu, born);
    if (bm) {
      unsigned int base = 0;
      const int leader = __ffs(bm) - 1;
      if (lane == leader) base = atomicAdd(lg.cursor, (unsigned int)__popc(bm));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (born) {
        const unsigned int idx = base + __popc(bm & ((1u << lane) - 1u));
        if (idx < lg.cap) lg.ent[idx] = log_entry((uint32_t)nb, slot == 3, wi, hloc);
        else atomicOr(lg.flags, 1u);
      }
    }
  }
}

// Slow path of one row (warp-uniform): stage in shared memory, rewrite word by word (the lane that holds
// a header rewrites that word) and store back. The matches are appended to the warp's list `ml`; their
// deltas are emitted later for all rows of this warp iteration together (emit_matches).
__device__ __noinline__ uint32_t merge_row_slow(int *sm, Match *ml, unsigned int *n_match, int4 v, int lane, uint64_t row,
                                                const StreamDev &s, const PairTableDev &t, const BirthLogDev &lg, int32_t a,
                                                int32_t b, int32_t new_id) {
  if (*n_match > MATCH_CAP - ROW / 2) {  // not enough room left for a full row of matches: flush first
    emit_matches(ml, *n_match, lane, t, lg, a, b, new_id);
    __syncwarp();
    if (lane == 0) *n_match = 0;
  }
  *reinterpret_cast<int4 *>(&sm[lane * 4]) = v;
  __syncwarp();
  uint32_t removed = 0;
  const int h[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (h[k] < 0 && h[k] != PAD) removed += merge_word_smem(sm, lane * 4 + k, ml, n_match, s.cnt, a, b, new_id, (uint32_t)row);
  __syncwarp();
  s.rows[row * (ROW / 4) + lane] = *reinterpret_cast<const int4 *>(&sm[lane * 4]);
  if (lane == 0) {  // the row now contains new_id
    const uint32_t hh = sig_hash(new_id);
    atomicOr(&s.sig[row * SIG_WORDS + (hh >> 5)], 1u << (hh & 31));
  }
  __syncwarp();
  return removed;
}

// The row scan of one merge (pairs without a birth log): every warp tests 32 row signatures per iteration,
// loads the candidate rows (four in flight), rewrites those with a match and emits their deltas. Loads of
// mutable data bypass L1: the word-granular path rewrites rows behind the back of the warp that scans them
// here, and the resident kernel keeps its L1 across merges. Returns this thread's removed-symbol count.
__device__ __forceinline__ uint32_t scan_rows(const StreamDev &s, const PairTableDev &t, const BirthLogDev &lg, int32_t a, int32_t b,
                                              int32_t new_id, int (*sm)[ROW], Match (*ml)[MATCH_CAP], unsigned int *n_match) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t ha = sig_hash(a), hb = sig_hash(b);
  const uint32_t wa = ha >> 5, ba = 1u << (ha & 31), wb = hb >> 5, bb = 1u << (hb & 31);
  uint32_t removed = 0;
  for (uint64_t base = warp * 32; base < s.n_rows; base += n_warps * 32) {
    const uint64_t row = base + lane;
    bool cand = false;
    if (row < s.n_rows) {
      const uint32_t *sg = s.sig + row * SIG_WORDS;
      cand = (__ldcg(&sg[wa]) & ba) && (__ldcg(&sg[wb]) & bb);
    }
    uint32_t cmask = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) n_match[wib] = 0;
    __syncwarp();
    while (cmask) {  // candidate rows, four loads in flight at a time
      uint64_t rr[4];
      int4 vv[4];
      int nc = 0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (cmask) {
          rr[u] = base + (__ffs(cmask) - 1);
          cmask &= cmask - 1;
          vv[u] = __ldcg(&s.rows[rr[u] * (ROW / 4) + lane]);
          nc = u + 1;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (u < nc) {
          const int4 v = vv[u];
          int nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
          if (lane == 31) nxt = PAD;
          const bool m = (v.x == a && v.y == b) || (v.y == a && v.z == b) || (v.z == a && v.w == b) || (v.w == a && nxt == b);
          if (__any_sync(0xffffffffu, m))
            removed += merge_row_slow(sm[wib], ml[wib], &n_match[wib], v, lane, rr[u], s, t, lg, a, b, new_id);
        }
      }
    }
    __syncwarp();
    if (n_match[wib]) emit_matches(ml[wib], n_match[wib], lane, t, lg, a, b, new_id);
    __syncwarp();
  }
  return removed;
}

// ---------------------------------------------------------------- word-granular merge (pairs with a birth log)
// One thread rewrites one word: the symbols are copied to thread-local memory, rewritten left to right with
// the reference's sequential semantics (already-merged left neighbour, not-yet-merged right neighbour:
// reference bpe.cpp:437-483), and the changed tail is stored back. Every match hands its four signed deltas
// to `sink.add`, every newly born pair (once per word) to `sink.birth`.
// hloc = flat index of the word's header in the row stream. Returns the number of matches.
__device__ unsigned long long g_word_trace[8];  // development aid (SWB_KERNEL_TRACE builds): stage cycles of block 0 / thread 0
#ifdef SWB_KERNEL_TRACE
#define SWB_WT(i, t0) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t1_ = clock64(); g_word_trace[i] += (unsigned long long)(t1_ - t0); t0 = t1_; } } while (0)
#else
#define SWB_WT(i, t0) do { } while (0)
#endif

// The four signed deltas of each recorded match (reference bpe.cpp:453-470), handed to the sink together.
// Called when the lanes of the warp have reconverged: the matches of different words sit at different
// positions, and emitting from inside the rewrite loop would run the table updates once per lane instead of
// once per warp.
template <typename Sink>
__device__ __forceinline__ void emit_word_matches(uint32_t nmatch, const int *mL, const int *mR, const unsigned char *mpos, uint32_t wi,
                                                  long long cc, int32_t a, int32_t b, int32_t new_id, Sink &sink) {
#pragma unroll 1
  for (uint32_t mi = 0; mi < nmatch; mi++) {
    const int L = mL[mi], R = mR[mi];
    const uint32_t pos = mpos[mi];
    int32_t bx[4], by[4];
    long long bd[4];
    uint64_t bk[4];
    int nbt = 0;
    if (L >= 0) {
      bx[0] = L; by[0] = a; bd[0] = -cc; bk[0] = touch_key(wi, pos, 0);
      bx[1] = L; by[1] = new_id; bd[1] = cc; bk[1] = touch_key(wi, pos, 1);
      nbt = 2;
    }
    if (R >= 0) {
      bx[nbt] = b; by[nbt] = R; bd[nbt] = -cc; bk[nbt] = touch_key(wi, pos, 2);
      bx[nbt + 1] = new_id; by[nbt + 1] = R; bd[nbt + 1] = cc; bk[nbt + 1] = touch_key(wi, pos, 3);
      nbt += 2;
    }
    sink.add_batch(nbt, bx, by, bd, bk);
  }
}

// One thread rewrites one word (reference bpe.cpp:437-483: already-merged left neighbour, not-yet-merged right
// neighbour). Every match hands its four signed deltas to the sink, every newly born pair (once per word) to
// `sink.birth`. hloc = flat index of the word's header in the row stream. Returns the number of matches.
//
// Words of up to WORD_FAST symbols (nearly all of them) never leave the registers: the three 16-byte chunks that
// hold them are requested together (one memory round trip), realigned, and rewritten by fully unrolled code whose
// only memory traffic are the stores of the changed symbols. Longer words take the general path through
// thread-local memory.
constexpr int WORD_FAST = 8;
template <typename Sink>
__device__ __forceinline__ uint32_t merge_one_word(const StreamDev &s, uint64_t hloc, uint32_t wi, int32_t a, int32_t b, int32_t new_id,
                                                   Sink &sink) {
  int32_t *flat = reinterpret_cast<int32_t *>(s.rows);
  const int hpos = (int)(hloc & (ROW - 1));
  const uint64_t row_base = hloc - hpos;
  const unsigned long long c = __ldg(&s.cnt[wi]);  // in flight while the symbols arrive
  long long wt0 = clock64(); (void)wt0;
  const unsigned int entered = __activemask();
  const int4 *rowv = reinterpret_cast<const int4 *>(flat + row_base);
  const int q0 = (hpos + 1) >> 2;
  const int4 pad4 = make_int4(PAD, PAD, PAD, PAD);
  int4 pre[3];
#pragma unroll
  for (int u = 0; u < 3; u++) pre[u] = q0 + u < ROW / 4 ? __ldcg(rowv + q0 + u) : pad4;
  const long long cc = (long long)c;
  int32_t *wsym = flat + row_base + hpos + 1;  // the word's first symbol
  int mL[ROW / 2], mR[ROW / 2];
  unsigned char mpos[ROW / 2];
  uint32_t nmatch = 0;

  // ---- realign: y[j] = symbol j of the word (PAD past what was loaded)
  int y[12];
  {
    const int x[12] = {pre[0].x, pre[0].y, pre[0].z, pre[0].w, pre[1].x, pre[1].y, pre[1].z, pre[1].w, pre[2].x, pre[2].y, pre[2].z, pre[2].w};
    const int off = (hpos + 1) & 3;
#pragma unroll
    for (int j = 0; j < 12; j++) {
      int v = PAD;
      if (off == 0) v = x[j];
      else if (off == 1) v = j + 1 < 12 ? x[j + 1] : PAD;
      else if (off == 2) v = j + 2 < 12 ? x[j + 2] : PAD;
      else v = j + 3 < 12 ? x[j + 3] : PAD;
      y[j] = v;
    }
  }
  int n = 0;  // length, if the word ends inside the first WORD_FAST + 1 slots
  bool fast = false;
#pragma unroll
  for (int j = WORD_FAST; j >= 0; j--)
    if (y[j] < 0) { n = j; fast = true; }
  // where the pair sits (bit r: symbols r, r+1). The register path takes the words with exactly ONE match -- nearly all
  // candidates; overlapping or repeated matches go through the general path below.
  unsigned int mm = 0;
#pragma unroll
  for (int r = 0; r + 1 < WORD_FAST; r++)
    if (y[r] == a && y[r + 1] == b && r + 1 < n) mm |= 1u << r;
  SWB_WT(0, wt0);
  if (fast && mm != 0 && (mm & (mm - 1)) == 0) {
    const int r0 = __ffs(mm) - 1;
    int L = -1, R = -1;
#pragma unroll
    for (int j = 0; j < WORD_FAST; j++) {
      if (j == r0 - 1) L = y[j];
      if (j == r0 + 2 && j < n) R = y[j];
    }
    // symbols r0 .. n-1 change: new_id, then the tail moved up by one, then PAD
#pragma unroll
    for (int q = 0; q < WORD_FAST; q++)
      if (q >= r0 && q < n) wsym[q] = q == r0 ? new_id : (q + 1 < n ? y[q + 1] : PAD);
    {
      const uint32_t hh = sig_hash(new_id);
      atomicOr(&s.sig[(row_base / ROW) * SIG_WORDS + (hh >> 5)], 1u << (hh & 31));
    }
    SWB_WT(1, wt0);
    __syncwarp(entered);
    SWB_WT(2, wt0);
    mL[0] = L; mR[0] = R; mpos[0] = (unsigned char)(hpos + 1 + r0);
    emit_word_matches(1, mL, mR, mpos, wi, cc, a, b, new_id, sink);
    if (L >= 0) sink.birth((uint32_t)L, false, wi, hloc);  // the final neighbours of the one new_id
    if (R >= 0) sink.birth((uint32_t)R, true, wi, hloc);
    SWB_WT(3, wt0);
#ifdef SWB_KERNEL_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) g_word_trace[6] += 1;
#endif
    return 1;
  }
  if (fast && mm == 0) { __syncwarp(entered); return 0; }  // a candidate that no longer holds the pair

  // ---- general path: the word goes through thread-local memory
  int buf[ROW];
  n = 0;
  {
    bool open = true;
    const int have = 12 - ((hpos + 1) & 3);  // symbols of the word that the three chunks can hold
#pragma unroll
    for (int j = 0; j < 12; j++) {
      if (!open || j >= have) continue;
      if (y[j] < 0) { open = false; continue; }
      buf[n++] = y[j];
    }
    // (a word that stops exactly at the end of the loaded chunks or of the row is closed by the loop condition)
    for (int p = hpos + 1 + 12 - ((hpos + 1) & 3); open && p < ROW; p += 4) {
      const int4 v = __ldcg(rowv + (p >> 2));
      const int x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!open) continue;
        if (x[k] < 0) { open = false; continue; }
        buf[n++] = x[k];
      }
    }
  }
  int w = 0, r = 0, first_changed = -1;
  while (r < n) {
    const int x = buf[r];
    if (x == a && r + 1 < n && buf[r + 1] == b) {
      mL[nmatch] = w > 0 ? buf[w - 1] : -1;       // already rewritten
      mR[nmatch] = r + 2 < n ? buf[r + 2] : -1;   // not yet rewritten
      mpos[nmatch] = (unsigned char)(hpos + 1 + r);
      if (first_changed < 0) first_changed = w;
      buf[w++] = new_id;
      r += 2;
      nmatch++;
    } else {
      buf[w++] = x;
      r++;
    }
  }
  __syncwarp(entered);
  emit_word_matches(nmatch, mL, mR, mpos, wi, cc, a, b, new_id, sink);
  if (!nmatch) return 0;
  for (int i = first_changed; i < n; i++) wsym[i] = i < w ? buf[i] : PAD;
  {  // the row now contains new_id
    const uint32_t hh = sig_hash(new_id);
    atomicOr(&s.sig[(row_base / ROW) * SIG_WORDS + (hh >> 5)], 1u << (hh & 31));
  }
  // births: the pairs around every new_id of the rewritten word, each (side, neighbour) once per word
  for (int i = first_changed; i < w; i++) {
    if (buf[i] != new_id) continue;
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int j = side ? i + 1 : i - 1;
      if (j < 0 || j >= w) continue;
      const int nb = buf[j];
      if (side == 1 && nb == new_id) continue;  // (new, new) is looked up through its left-side entry
      bool dup = false;
      for (int q = first_changed; q < i && !dup; q++) {
        if (buf[q] != new_id) continue;
        const int jq = side ? q + 1 : q - 1;
        dup = jq >= 0 && jq < w && buf[jq] == nb;
      }
      if (!dup) sink.birth((uint32_t)nb, side == 1, wi, hloc);
    }
  }
  return nmatch;
}

// warp-aggregated append to the birth log from divergent code
__device__ __forceinline__ void log_append(const BirthLogDev &lg, uint4 e) {
  const unsigned int conv = __activemask();
  const int lane_id = threadIdx.x & 31;
  const int leader = __ffs(conv) - 1;
  unsigned int base = 0;
  if (lane_id == leader) base = atomicAdd(lg.cursor, (unsigned int)__popc(conv));
  base = __shfl_sync(conv, base, leader);
  const unsigned int idx = base + __popc(conv & ((1u << lane_id) - 1u));
  if (idx < lg.cap) lg.ent[idx] = e;
  else atomicOr(lg.flags, 1u);
}

// deltas -> the global per-merge pair table, births -> the global log
struct GlobalSink {
  const PairTableDev &t;
  const BirthLogDev &lg;
  __device__ __forceinline__ void add_batch(int n, const int32_t *x, const int32_t *y, const long long *delta, const uint64_t *key) {
#pragma unroll 1
    for (int i = 0; i < n; i++) pt_add(t, x[i], y[i], delta[i], key[i]);
  }
  __device__ __forceinline__ void birth(uint32_t other, bool right_side, uint32_t wi, uint64_t hloc) {
    log_append(lg, log_entry(other, right_side, wi, hloc));
  }
};

// The indexed scan of one merge over the whole grid: one thread per entry of a list of (neighbour, side, word) entries --
// the birth log of the merge that created max(a, b), or the occurrence index of a pair of two initial symbols; the entries
// with the wanted neighbour and side name the words to rewrite. Returns this thread's removed-symbol count.
__device__ __forceinline__ uint32_t scan_entry_words(const StreamDev &s, const PairTableDev &t, const BirthLogDev &lg, const uint4 *__restrict__ ent,
                                                     uint64_t lo, uint64_t n, uint64_t ent_cap, uint32_t want_other, uint32_t want_side, int32_t a, int32_t b,
                                                     int32_t new_id) {
  GlobalSink sink{t, lg};
  uint32_t removed = 0;
#pragma unroll 1
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (!SWB_DBG_OK(lo + i < ent_cap, 4, lo, i, n)) break;
    const uint4 e = __ldcg(&ent[lo + i]);
    if (e.x == want_other && (e.y & 0x80000000u) == want_side && SWB_DBG_OK((((uint64_t)e.w << 32) | e.z) < s.n_rows * ROW, 5, ((uint64_t)e.w << 32) | e.z, e.y, i))
      removed += merge_one_word(s, ((uint64_t)e.w << 32) | e.z, e.y & 0x7FFFFFFFu, a, b, new_id, sink);
  }
  return removed;
}

// One merge over the whole grid: the indexed word scan when the pair has a birth log or (two initial symbols) an occurrence
// list, the row-signature scan otherwise.
__device__ __forceinline__ uint32_t scan_merge(const StreamDev &s, const PairTableDev &t, const BirthLogDev &lg, int32_t a, int32_t b,
                                               int32_t new_id, int (*sm)[ROW], Match (*ml)[MATCH_CAP], unsigned int *n_match) {
  uint32_t merge, other, side;
  if (log_lookup(lg, a, b, merge, other, side)) {
    const uint64_t lo = __ldcg(&lg.start[merge]), hi = __ldcg(&lg.start[merge + 1]);
    return scan_entry_words(s, t, lg, lg.ent, lo, hi >= lo ? hi - lo : 0, lg.cap, other, side, a, b, new_id);
  }
  if (ip_lookup(lg, a, b)) {  // every word that held (a, b) when the corpus was loaded (a superset of those that still do)
    const uint32_t pk = (uint32_t)a * 256u + (uint32_t)b;
    const uint64_t lo = __ldcg(&lg.ip_start[pk]), hi = __ldcg(&lg.ip_start[pk + 1]);
    return scan_entry_words(s, t, lg, lg.ip_ent, lo, hi - lo, ~0ull, (uint32_t)a, 0u, a, b, new_id);
  }
  return scan_rows(s, t, lg, a, b, new_id, sm, ml, n_match);
}

// fused != 0: the last block to finish also emits the records and publishes the header (one launch per
// merge); used when there are no long words to process after this kernel.
__global__ void __launch_bounds__(MERGE_THREADS)
merge_rows(StreamDev s, PairTableDev t, int32_t a, int32_t b, int32_t new_id, unsigned long long *removed_total,
           int fused, EmitMode em, Rec *__restrict__ out, size_t out_cap, unsigned long long *__restrict__ out_hdr,
           unsigned long long seq) {
  __shared__ __align__(16) int sm[MERGE_WARPS][ROW];
  __shared__ Match ml[MERGE_WARPS][MATCH_CAP];
  __shared__ unsigned int n_match[MERGE_WARPS];
  __shared__ unsigned long long csum_sh[64];
  __shared__ __align__(16) Rec stage[STAGE_RECS];  // 12 KB: the fused tail's records before they cross PCIe
  __shared__ unsigned int tail_count;
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31;
#ifdef SWB_KERNEL_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0)); out_hdr[11] = t0; }
#endif
  uint32_t removed = scan_merge(s, t, em.log, a, b, new_id, sm, ml, n_match);
#ifdef SWB_KERNEL_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tt)); out_hdr[13] = tt; }
#endif
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) removed += __shfl_down_sync(0xffffffffu, removed, d);
  if (lane == 0 && removed) atomicAdd(removed_total, (unsigned long long)removed);
  if (!fused) return;
  __syncthreads();
#ifdef SWB_KERNEL_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tt)); out_hdr[14] = tt; }
#endif
  if (threadIdx.x == 0) {  // (a release by one thread after the barrier covers the whole block's writes)
    __threadfence();
    is_last = atomicAdd(t.done_blocks, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#ifdef SWB_KERNEL_TRACE
  unsigned long long tr_scan_done; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr_scan_done));
#endif
  TailSmem ts{stage, csum_sh, &tail_count};
#ifdef SWB_KERNEL_TRACE
  out_hdr[8] = tr_scan_done;
#endif
  fused_tail(t, em, ts, out, out_cap, out_hdr, removed_total, seq, 0u);
}

__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// long words: lanes stride over the word to detect; the rare word with a match is rewritten by lane 0
__global__ void __launch_bounds__(128)
merge_long(StreamDev s, PairTableDev t, int32_t a, int32_t b, int32_t new_id, unsigned long long *removed_total) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < s.n_long; j += n_warps) {
    int32_t *sym = s.long_syms + s.long_off[j];
    const uint32_t L = s.long_len[j];
    bool m = false;
    for (uint32_t i = lane; i + 1 < L; i += 32) m |= (sym[i] == a && sym[i + 1] == b);
    if (!__any_sync(0xffffffffu, m)) continue;
    if (lane == 0) {
      const uint32_t li = s.long_word[j];
      const long long c = (long long)s.cnt[li];
      const uint64_t g = word_gwi(s, li);
      uint32_t r = 0, w = 0, nmatch = 0;
      while (r < L) {
        const int x = sym[r];
        if (x == a && r + 1 < L && sym[r + 1] == b) {
          if (w > 0) {
            const int Lf = sym[w - 1];
            pt_add(t, Lf, a, -c, touch_key(g, r, 0));
            pt_add(t, Lf, new_id, c, touch_key(g, r, 1));
          }
          if (r + 2 < L) {
            const int R = sym[r + 2];
            pt_add(t, b, R, -c, touch_key(g, r, 2));
            pt_add(t, new_id, R, c, touch_key(g, r, 3));
          }
          sym[w++] = new_id; r += 2; nmatch++;
        } else {
          if (w != r) sym[w] = x;
          w++; r++;
        }
      }
      s.long_len[j] = w;
      atomicAdd(removed_total, (unsigned long long)nmatch);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- initial pair count
// reference bpe.cpp:329-350: every adjacent pair of every word, weighted by the word count, pairs
// containing unk_id skipped. key = (word, position) so the host can replay first-touch order.
__global__ void __launch_bounds__(MERGE_THREADS)
count_rows(StreamDev s, PairTableDev t, int32_t unk) {
  __shared__ __align__(16) int sm[MERGE_THREADS / 32][ROW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warp = (blockIdx.x * (uint64_t)MERGE_THREADS + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * MERGE_THREADS) >> 5;
  int *row = sm[wib];
  for (uint64_t r0 = warp; r0 < s.n_rows; r0 += n_warps) {
    const int4 v = s.rows[r0 * (ROW / 4) + lane];
    *reinterpret_cast<int4 *>(&row[lane * 4]) = v;
    __syncwarp();
    const int h[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (h[k] < 0 && h[k] != PAD) {
        const int p = lane * 4 + k;
        const uint32_t li = (uint32_t)(~h[k]);
        const long long c = (long long)s.cnt[li];
        const uint64_t g = word_gwi(s, li);
        for (int r = p + 1; r + 1 < ROW; r++) {
          const int x = row[r], y = row[r + 1];
          if (y < 0) break;
          if (x == unk || y == unk) continue;
          pt_add(t, x, y, c, touch_key(g, r, 0));
        }
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128)
count_long(StreamDev s, PairTableDev t, int32_t unk) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < s.n_long; j += n_warps) {
    const int32_t *sym = s.long_syms + s.long_off[j];
    const uint32_t L = s.long_len[j];
    if (L < 2) continue;
    const uint32_t li = s.long_word[j];
    const long long c = (long long)s.cnt[li];
    const uint64_t g = word_gwi(s, li);
    for (uint32_t i = lane; i + 1 < L; i += 32) {
      const int x = sym[i], y = sym[i + 1];
      if (x == unk || y == unk) continue;
      pt_add(t, x, y, c, touch_key(g, i, 0));
    }
  }
}

// ---------------------------------------------------------------- token histogram (bpe_save)
__global__ void __launch_bounds__(MERGE_THREADS)
tokfreq_rows(StreamDev s, unsigned long long *__restrict__ freq, uint32_t T) {
  __shared__ __align__(16) int sm[MERGE_THREADS / 32][ROW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warp = (blockIdx.x * (uint64_t)MERGE_THREADS + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * MERGE_THREADS) >> 5;
  int *row = sm[wib];
  for (uint64_t r0 = warp; r0 < s.n_rows; r0 += n_warps) {
    const int4 v = s.rows[r0 * (ROW / 4) + lane];
    *reinterpret_cast<int4 *>(&row[lane * 4]) = v;
    __syncwarp();
    const int h[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (h[k] < 0 && h[k] != PAD) {
        const unsigned long long c = s.cnt[(uint32_t)(~h[k])];
        for (int r = lane * 4 + k + 1; r < ROW; r++) {
          const int x = row[r];
          if (x < 0) break;
          if ((uint32_t)x < T) atomicAdd(&freq[x], c);
        }
      }
    }
    __syncwarp();
  }
}

__global__ void tokfreq_long(StreamDev s, unsigned long long *__restrict__ freq, uint32_t T) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < s.n_long; j += n_warps) {
    const int32_t *sym = s.long_syms + s.long_off[j];
    const uint32_t L = s.long_len[j];
    const unsigned long long c = s.cnt[s.long_word[j]];
    for (uint32_t i = lane; i < L; i += 32)
      if ((uint32_t)sym[i] < T) atomicAdd(&freq[sym[i]], c);
  }
}

// ---------------------------------------------------------------- multi-GPU exchange
// Per merge every rank copies its local (pair, net delta, first-touch key) records into its slot of an
// all-gather buffer [DIST_HDR_WORDS x u64 header][cap x Rec]; after the all-gather every rank reduces all
// slots by pair into a second table and runs the same tail as the single-GPU path on it (replicated
// frequency table -> identical records -> identical heap replica on every rank).
constexpr int DIST_HDR_WORDS = 16;

// local table -> this rank's slot, WITHOUT clearing it (a capacity overflow anywhere makes every rank grow
// its slots and copy again; the table is cleared by dist_reduce once the exchange has succeeded).
__global__ void __launch_bounds__(256)
dist_copy_out(PairTableDev t, unsigned long long *__restrict__ slot, size_t cap) {
  const unsigned int n = __ldcg(t.n_touched);
  Rec *out = reinterpret_cast<Rec *>(slot + DIST_HDR_WORDS);
  if (n <= cap) {
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const uint4 tc = __ldcg(&t.touched[i]);
      Rec r;
      r.first = (int32_t)tc.z; r.second = (int32_t)tc.y;
      r.delta = (long long)__ldcg(&t.val[tc.x]); r.key = (long long)__ldcg(&t.minkey[tc.x]);
      out[i] = r;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { slot[0] = n; slot[1] = (n > cap ? 4u : 0u) | (__ldcg(t.flags) & 1u); }
}

__global__ void __launch_bounds__(256)
dist_reduce(const unsigned long long *__restrict__ slots, int nranks, size_t slot_words, PairTableDev local, PairTableDev t,
            EmitMode em, Rec *__restrict__ out, size_t out_cap, unsigned long long *__restrict__ out_hdr,
            unsigned long long *removed, unsigned long long seq) {
  __shared__ __align__(16) Rec stage[STAGE_RECS];
  __shared__ unsigned long long csum_sh[64];
  __shared__ unsigned int tail_count;
  __shared__ bool is_last;
  unsigned int worst = 0, fl = 0;
  for (int r = 0; r < nranks; r++) {
    const unsigned long long *h = slots + (size_t)r * slot_words;
    worst = max(worst, (unsigned int)h[0]);
    fl |= (unsigned int)h[1];
  }
  const bool overflow = (fl & 5u) != 0;
  if (!overflow) {
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // the exchange succeeded: the local table can be cleared now
    const unsigned int nl = __ldcg(local.n_touched);
    for (unsigned int i = gtid; i < nl; i += gsz) {
      const unsigned int h = __ldcg(&local.touched[i]).x;
      local.keys[h] = PT_EMPTY; local.val[h] = 0; local.minkey[h] = ~0ull;
    }
    for (int r = 0; r < nranks; r++) {
      const unsigned long long *h = slots + (size_t)r * slot_words;
      const unsigned int n = (unsigned int)h[0];
      const Rec *recs = reinterpret_cast<const Rec *>(h + DIST_HDR_WORDS);
      for (unsigned int i = gtid; i < n; i += gsz) {
        const Rec rc = recs[i];
        pt_add(t, (int32_t)rc.first, (int32_t)rc.second, rc.delta, (uint64_t)rc.key);
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(t.done_blocks, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  if (overflow) {  // nothing was consumed: report the largest list so that every rank grows its slots alike
    if (threadIdx.x == 0) {
      out_hdr[1] = worst; out_hdr[2] = 4u | (fl & 1u); out_hdr[3] = 0; out_hdr[4] = 0; out_hdr[5] = 0;
      out_hdr[6] = hdr_check(seq, worst, 4u | (fl & 1u), 0, 0, 0);
      out_hdr[0] = seq; out_hdr[7] = seq;
      *t.done_blocks = 0;
    }
    return;
  }
  if (threadIdx.x == 0) { *local.n_touched = 0; *local.flags = 0; }
  TailSmem ts{stage, csum_sh, &tail_count};
  fused_tail(t, em, ts, out, out_cap, out_hdr, removed, seq, 0u);
}

// ---------------------------------------------------------------- word extraction (accessors)
// live length of every word of this rank (0 for other ranks' words)
__global__ void words_live_len(StreamDev s, const uint64_t *__restrict__ wloc, const uint32_t *__restrict__ long_index,
                               uint32_t W, int rank, int nranks, uint32_t *__restrict__ out_len) {
  const int32_t *flat = reinterpret_cast<const int32_t *>(s.rows);
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
    if ((int)(w % (uint32_t)nranks) != rank) { out_len[w] = 0; continue; }
    const uint32_t j = long_index[w];
    if (j != 0xFFFFFFFFu) { out_len[w] = s.long_len[j]; continue; }
    const uint64_t base = wloc[w];
    const int pos = (int)(base % ROW);
    uint32_t n = 0;
    for (int r = pos + 1; r < ROW && flat[base - pos + r] >= 0; r++) n++;
    out_len[w] = n;
  }
}

__global__ void words_copy_syms(StreamDev s, const uint64_t *__restrict__ wloc, const uint32_t *__restrict__ long_index,
                                uint32_t W, int rank, int nranks, const uint64_t *__restrict__ out_off,
                                int32_t *__restrict__ out) {
  const int32_t *flat = reinterpret_cast<const int32_t *>(s.rows);
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
    if ((int)(w % (uint32_t)nranks) != rank) continue;
    const uint32_t j = long_index[w];
    int32_t *dst = out + out_off[w];
    if (j != 0xFFFFFFFFu) {
      const int32_t *src = s.long_syms + s.long_off[j];
      for (uint32_t k = 0; k < s.long_len[j]; k++) dst[k] = src[k];
      continue;
    }
    const uint64_t base = wloc[w];
    const int pos = (int)(base % ROW);
    for (int r = pos + 1, k = 0; r < ROW && flat[base - pos + r] >= 0; r++, k++) dst[k] = flat[base - pos + r];
  }
}

}  // namespace swb
