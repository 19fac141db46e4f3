// host_core.hpp -- host half of the BPE merge loop: pair table, exact heap replica, delta application.
//
// The merge order of the reference is not a pure function of (freq, pair): equal-frequency ties
// are broken by the physical history of its binary heap and by the iteration orders of its three
// hash maps. This file replays exactly those orders from the compact
// per-merge records the GPU kernels produce: (pair, net delta, first-touch key). It is tiny and
// sequential by nature (a heap), so it lives on the host; the O(corpus) work lives in the kernels.
//
// Pure C++17, no CUDA, so the multi-rank logic (swb_dist_*) can be exercised without a GPU.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <sys/mman.h>

#include "../../include/shredword_b200.h"

namespace swb {

struct Rec { int64_t first, second, delta, key; };  // wire format of a record: 4 x int64
static_assert(sizeof(Rec) == 32, "record must be 4 x int64");

// The heap array and the pair table are tens of MB that every merge probes at random: on 4 KB pages nearly every
// probe also misses the TLB. Large blocks are 2 MB-aligned and offered to the kernel as transparent huge pages
// (a hint: without THP this is a plain allocation). Memory comes from / goes back to malloc's free().
static inline void *huge_alloc(size_t bytes) {
  constexpr size_t HP = 2u << 20;
  if (bytes < HP) return malloc(bytes ? bytes : 1);
  const size_t rounded = (bytes + HP - 1) / HP * HP;
  void *p = aligned_alloc(HP, rounded);
  if (p) madvise(p, rounded, MADV_HUGEPAGE);
  return p;
}
template <class T>
struct HugeAllocator {
  using value_type = T;
  HugeAllocator() = default;
  template <class U> HugeAllocator(const HugeAllocator<U> &) {}
  T *allocate(size_t n) {
    void *p = huge_alloc(n * sizeof(T));
    if (!p) throw std::bad_alloc();
    return static_cast<T *>(p);
  }
  void deallocate(T *p, size_t) { free(p); }
  template <class U> bool operator==(const HugeAllocator<U> &) const { return true; }
  template <class U> bool operator!=(const HugeAllocator<U> &) const { return false; }
};

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
static inline uint64_t pack_pair(int32_t a, int32_t b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }

// pair -> {freq, version}. One 24-byte open-addressing slot per pair (a lookup is one cache miss, and
// the misses of a whole delta list are overlapped by prefetching). `order` is the creation rank:
// the reference's BIMap iteration order (bucket ascending, chain in creation order; reference
// hash.cpp:104-130) is observable through bpe_count_bigrams' push pass (reference bpe.cpp:359-366).
struct PairInfo { int32_t first, second; uint64_t freq; uint32_t version; uint32_t order; /* 0 = free slot */ };
static_assert(sizeof(PairInfo) == 24, "PairInfo is one 24-byte slot");

class PairTable {
 public:
  PairTable() { clear(); }
  void clear() {
    slots_.assign(1u << 14, PairInfo{0, 0, 0, 0, 0});
    n_ = 0;
  }
  size_t size() const { return n_; }
  size_t home(int32_t a, int32_t b) const { return (size_t)mix64(pack_pair(a, b)) & (slots_.size() - 1); }
  void prefetch(int32_t a, int32_t b) const { __builtin_prefetch(&slots_[home(a, b)]); }
  // find-or-create (reference bimap_get: a missing pair reads as freq 0, version 0).
  // The reference is valid until the next get().
  PairInfo &get(int32_t a, int32_t b) {
    if ((n_ + 1) * 2 > slots_.size()) rehash(slots_.size() * 4);
    const size_t mask = slots_.size() - 1;
    size_t h = home(a, b);
    for (;;) {
      PairInfo &p = slots_[h];
      if (!p.order) {
        p.first = a; p.second = b; p.freq = 0; p.version = 0; p.order = (uint32_t)(++n_);
        return p;
      }
      if (p.first == a && p.second == b) return p;
      h = (h + 1) & mask;
    }
  }
  uint32_t find_version(int32_t a, int32_t b) const {
    const size_t mask = slots_.size() - 1;
    size_t h = home(a, b);
    for (;;) {
      const PairInfo &p = slots_[h];
      if (!p.order) return 0;
      if (p.first == a && p.second == b) return p.version;
      h = (h + 1) & mask;
    }
  }
  const PairInfo *find(int32_t a, int32_t b) const {
    const size_t mask = slots_.size() - 1;
    size_t h = home(a, b);
    for (;;) {
      const PairInfo &p = slots_[h];
      if (!p.order) return nullptr;
      if (p.first == a && p.second == b) return &p;
      h = (h + 1) & mask;
    }
  }
  // all entries in creation order
  void in_creation_order(std::vector<PairInfo> &out) const {
    out.resize(n_);
    for (const PairInfo &p : slots_)
      if (p.order) out[p.order - 1] = p;
  }

 private:
  void rehash(size_t n) {
    std::vector<PairInfo, HugeAllocator<PairInfo>> old;
    old.swap(slots_);
    slots_.assign(n, PairInfo{0, 0, 0, 0, 0});
    for (const PairInfo &p : old) {
      if (!p.order) continue;
      size_t h = home(p.first, p.second);
      while (slots_[h].order) h = (h + 1) & (n - 1);
      slots_[h] = p;
    }
  }
  std::vector<PairInfo, HugeAllocator<PairInfo>> slots_;
  size_t n_ = 0;
};

// FNV-1a over the 8 little-endian bytes of {first, second} (reference hash.cpp:7-16)
static inline uint32_t ref_pair_hash(int32_t first, int32_t second) {
  uint32_t h = 2166136261u;
  const uint32_t w[2] = {(uint32_t)first, (uint32_t)second};
  for (int k = 0; k < 2; k++)
    for (int i = 0; i < 4; i++) { h ^= (w[k] >> (8 * i)) & 0xffu; h *= 16777619u; }
  return h;
}

// Reduce records by pair: sum delta, min key. Order of the output is unspecified. Returns new count.
static inline size_t reduce_records(Rec *recs, size_t n) {
  if (n < 2) return n;
  size_t nslots = 16;
  while (nslots < 2 * n) nslots *= 2;
  std::vector<uint32_t> slots(nslots, 0);
  size_t out = 0;
  for (size_t i = 0; i < n; i++) {
    const Rec r = recs[i];
    size_t h = (size_t)mix64(pack_pair((int32_t)r.first, (int32_t)r.second)) & (nslots - 1);
    for (;;) {
      if (!slots[h]) { recs[out] = r; slots[h] = (uint32_t)(++out); break; }
      Rec &q = recs[slots[h] - 1];
      if (q.first == r.first && q.second == r.second) {
        q.delta += r.delta;
        if ((uint64_t)r.key < (uint64_t)q.key) q.key = r.key;
        break;
      }
      h = (h + 1) & (nslots - 1);
    }
  }
  return out;
}

class HostCore {
 public:
  explicit HostCore(Trainer *tr) : tr_(tr) {}
  ~HostCore() { free(tr_->heap.data); tr_->heap.data = nullptr; tr_->heap.size = tr_->heap.cap = 0; }

  int log_level = 0;
  uint64_t n_records = 0, n_pushes = 0, n_pops = 0, heap_peak = 0;

  // ---- heap: the exact array heap of reference heap.cpp:53-114, stored in the public Trainer fields
  void heap_reset() {  // heap_free + heap_init(4096), reference bpe.cpp:182-183
    MaxHeap &h = tr_->heap;
    if (!h.data) { h.cap = 4096; h.data = (HeapEntry *)huge_alloc(h.cap * sizeof(HeapEntry)); }
    h.size = 0;
  }
  void heap_push(int32_t a, int32_t b, uint64_t freq, uint32_t version) {
    MaxHeap &h = tr_->heap;
    if (!h.data) heap_reset();
    if (h.size == h.cap) {
      h.cap *= 2;
      HeapEntry *bigger = (HeapEntry *)huge_alloc(h.cap * sizeof(HeapEntry));
      if (!bigger) { fprintf(stderr, "[ERROR]\t heap reallocation failed\n"); abort(); }
      memcpy(bigger, h.data, h.size * sizeof(HeapEntry));
      free(h.data);
      h.data = bigger;
    }
    size_t i = h.size++;
    n_pushes++;
    if (h.size > heap_peak) heap_peak = h.size;
    h.data[i].key.first = a; h.data[i].key.second = b; h.data[i].freq = freq; h.data[i].version = version;
    while (i > 0) {  // stop as soon as parent.freq >= child.freq (reference heap.cpp:74-79)
      size_t p = (i - 1) >> 1;
      if (h.data[p].freq >= h.data[i].freq) break;
      std::swap(h.data[p], h.data[i]);
      i = p;
    }
  }
  HeapEntry heap_pop() {
    MaxHeap &h = tr_->heap;
    HeapEntry top = h.data[0];
    n_pops++;
    const HeapEntry last = h.data[--h.size];
    // the moved element sinks from the root: the hole moves down, only the entries that move up are written
    size_t i = 0;
    for (;;) {  // left child if strictly larger, then right if strictly larger than that (heap.cpp:97-111)
      const size_t l = 2 * i + 1, r = l + 1;
      {  // the four grandchildren are contiguous (96 bytes): fetch them while the children are compared
        const size_t g = 4 * i + 3;
        if (g < h.size) { __builtin_prefetch(&h.data[g]); __builtin_prefetch(reinterpret_cast<const char *>(&h.data[g]) + 64); }
      }
      size_t best = i;
      uint64_t bf = last.freq;
      if (l < h.size && h.data[l].freq > bf) { best = l; bf = h.data[l].freq; }
      if (r < h.size && h.data[r].freq > bf) best = r;
      if (best == i) break;
      h.data[i] = h.data[best];
      i = best;
    }
    if (h.size > 0) h.data[i] = last;
    return top;
  }
  bool heap_empty() const { return tr_->heap.size == 0; }

  // ---- reference bpe_init's table part (bpe.cpp:177-183)
  void reset_tables() {
    pairs_.clear();
    heap_reset();
  }

  // ---- reference bpe_count_bigrams (bpe.cpp:315-370), fed with the reduced count records.
  // Pass 1 creates/accumulates entries in first-touch order; pass 2 pushes every entry of the
  // table with freq >= min_pair_freq in (FNV bucket, creation) order.
  void seed_counts(const Rec *recs, size_t n) {
    std::vector<Rec> v(recs, recs + n);
    std::sort(v.begin(), v.end(), [](const Rec &x, const Rec &y) { return (uint64_t)x.key < (uint64_t)y.key; });
    for (const Rec &r : v) {
      PairInfo &p = pairs_.get((int32_t)r.first, (int32_t)r.second);
      if (p.freq == 0) p.version = 0;  // bpe.cpp:342-345
      p.freq += (uint64_t)r.delta;
    }
    const size_t P = pairs_.size();
    const size_t pushed = push_all_in_table_order();
    const uint64_t minf = tr_->config.min_pair_freq;
    if (log_level > 0) {
      printf("[INFO]\t Counted %zu unique pairs\n", P);
      printf("[INFO]\t Added %zu pairs to heap (freq >= %llu)\n", pushed, (unsigned long long)minf);
    }
  }

  // reference bpe.cpp:359-366: every entry of the table with freq >= min_pair_freq, FNV bucket ascending,
  // chain in creation order
  size_t push_all_in_table_order() {
    std::vector<PairInfo> all;
    pairs_.in_creation_order(all);
    const size_t P = all.size();
    std::vector<uint32_t> order(P), bucket(P);
    for (size_t i = 0; i < P; i++) { order[i] = (uint32_t)i; bucket[i] = ref_pair_hash(all[i].first, all[i].second) & 4095u; }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return bucket[x] < bucket[y]; });
    const uint64_t minf = tr_->config.min_pair_freq;
    size_t pushed = 0;
    for (uint32_t i : order) {
      const PairInfo &p = all[i];
      if (p.freq >= minf) { heap_push(p.first, p.second, p.freq, p.version); pushed++; }
    }
    return pushed;
  }

  // ---- reference bpe_merge_batch, decision part (bpe.cpp:405-429): pop until a live entry.
  bool next_merge(int32_t *a, int32_t *b, int32_t *new_id) {
    const uint64_t minf = tr_->config.min_pair_freq;
    while (!heap_empty()) {
      {  // the entries that can reach the root next: start the misses of their table slots while this pop sifts down
        const MaxHeap &h = tr_->heap;
        for (size_t c = 1; c < 7 && c < h.size; c++) pairs_.prefetch(h.data[c].key.first, h.data[c].key.second);
      }
      HeapEntry top = heap_pop();
      PairInfo &info = pairs_.get(top.key.first, top.key.second);
      if (top.version != info.version) continue;  // stale (bpe.cpp:412-415)
      if (info.freq < minf) continue;             // bpe.cpp:418-421
      cur_a_ = top.key.first; cur_b_ = top.key.second;
      cur_new_ = (int32_t)(256 + tr_->num_merges);  // bpe.cpp:424
      merges_.push_back(PairKey{cur_a_, cur_b_});
      tr_->merge_ops = merges_.data();
      pending_ = true;
      if (log_level > 0)
        printf("[MERGE]\t Merging (%d,%d) freq=%llu -> new_id=%d (merge %zu)\n", cur_a_, cur_b_,
               (unsigned long long)info.freq, cur_new_, tr_->num_merges + 1);
      *a = cur_a_; *b = cur_b_; *new_id = cur_new_;
      return true;
    }
    return false;
  }

  // ---- look-ahead for the resident kernel: which pairs will the next calls of next_merge() return?
  // Called between next_merge() and the apply of that merge (the "pending" merge, number j), i.e. before the
  // pushes of merge j are known. Fills up to `want` entries: entry i = (pair, F_i) is the (i+1)-th LIVE entry in
  // the order in which the heap would pop (reference heap.cpp:88-114) if nothing were pushed. Statement made
  // for entry i:  merge j+1+i is that pair, PROVIDED that when its turn comes
  //   (1) every entry pushed by the merges j .. j+i had a frequency < F_i, and
  //   (2) the pair's frequency is (still) F_i.
  // Why this is exact, ties included: entries with frequency >= F_i form a top-closed region of the heap. A push
  // below F_i sifts up only past entries smaller than itself, so it never moves an entry of the region; a
  // sift-down prefers any entry of the region over any entry outside it (strict comparisons, heap.cpp:97-111),
  // so the order in which the entries of the region reach the root depends on the region alone -- provided the
  // elements moved from the tail to the root are outside it (checked here). Entries popped before entry i are
  // stale (for ever: versions only grow) or earlier entries of the list; one of those that loses its place
  // (touched by a merge in between) makes the later ones move up, and then (2) fails for them once they have
  // been merged (frequency 0). So an accepted entry is always the true next pair; a wrong guess is only ever
  // turned down. The heap is not modified and no pop is simulated: inside that region a removed root is replaced by its
  // larger child, the LEFT one on ties (the element the real algorithm moves in from the tail is smaller than both and
  // sinks out of the region), and so on down -- i.e. the entries of a subtree reach the root as "the subtree's root, then
  // the two child streams merged, left first on ties". That makes the pop order a total order: frequency descending, and
  // among equal frequencies the PRE-ORDER position in the tree (an ancestor before its descendants, a left subtree before
  // the right one). A child never precedes its parent in that order, so a best-first walk from the root with a small
  // priority queue over the frontier produces exactly the pop order, visiting every entry once.
  struct Peek { int32_t a, b; uint64_t freq; };
  size_t peek_next(Peek *out, size_t want) {
    const MaxHeap &h = tr_->heap;
    const uint64_t minf = tr_->config.min_pair_freq;
    // the real pops would move the last MAX_POPS elements of the array to the root: the look-ahead makes no statement
    // as soon as it has to look at one of them (small heaps: the top region may reach the tail of the array)
    const size_t MAX_POPS = std::min<size_t>(640, h.size / 8);
    if (MAX_POPS < 4) return 0;
    const size_t n_eff = h.size - MAX_POPS;
    front_.clear();
    front_.push_back(Front{h.data[0].freq, 0});
    pairs_.prefetch(h.data[0].key.first, h.data[0].key.second);
    size_t got = 0, pops = 0;
    uint64_t tail_max = 0;  // largest frequency among the tail elements the real pops would have moved to the root
    while (got < want && pops < MAX_POPS && !front_.empty()) {
      std::pop_heap(front_.begin(), front_.end(), front_after);
      const size_t pos = front_.back().pos;
      front_.pop_back();
      if (pos >= n_eff) break;  // too close to the tail of the array: no statement
      const HeapEntry &e = h.data[pos];
      for (size_t c = 2 * pos + 1; c <= 2 * pos + 2 && c < h.size; c++) {
        pairs_.prefetch(h.data[c].key.first, h.data[c].key.second);  // (looked up when it leaves the queue: start the miss now)
        front_.push_back(Front{h.data[c].freq, c});
        std::push_heap(front_.begin(), front_.end(), front_after);
      }
      const uint64_t tf = h.data[h.size - 1 - pops].freq;
      if (tf > tail_max) tail_max = tf;
      pops++;
      const PairInfo *info = pairs_.find(e.key.first, e.key.second);
      if (info && info->version == e.version && info->freq >= minf) {
        if (info->freq != e.freq || tail_max >= e.freq) break;
        out[got++] = Peek{e.key.first, e.key.second, e.freq};
      }
    }
    return got;
  }
  // Is a look-ahead entry made earlier still what it was (its pair untouched since)? A merge in between that changed the
  // pair's frequency has made the heap entry stale: the list it came from has to be made again.
  bool peek_still_valid(const Peek &e) const {
    const PairInfo *info = pairs_.find(e.a, e.b);
    return info && info->freq == e.freq;
  }
  bool peek_next(int32_t *a, int32_t *b, uint64_t *freq) {  // the first entry only
    Peek p;
    if (peek_next(&p, 1) != 1) return false;
    *a = p.a; *b = p.b; *freq = p.freq;
    return true;
  }

  // ---- reference bpe_merge_batch, bookkeeping part (bpe.cpp:486-526) for the pending merge.
  // recs: (first, second) are the raw ids of the touched pair, delta its net change, key the
  // first-touch order of the pair inside this merge.
  void apply(const Rec *recs, size_t n) {
    if (!pending_) return;
    n_records += n;
    if (tr_->config.unk_id < 0 && n) {
      // The reference keys its per-merge delta map by ((uint64_t)first << 32) | (uint64_t)second with
      // int32 operands (bpe.cpp:456-457): a negative `second` sign-extends over `first`. Replayed here
      // so that negative unk ids behave identically (distinct raw pairs can collapse into one entry).
      scratch_.assign(recs, recs + n);
      for (Rec &r : scratch_) {
        const uint64_t ph = ((uint64_t)(int64_t)(int32_t)r.first << 32) | (uint64_t)(int64_t)(int32_t)r.second;
        r.first = (int32_t)(ph >> 32); r.second = (int32_t)(ph & 0xFFFFFFFFu);
      }
      n = reduce_records(scratch_.data(), n);
      recs = scratch_.data();
    }
    sort_delta_order(recs, n);
    const uint64_t minf = tr_->config.min_pair_freq;
    constexpr size_t PF = 12;  // slot prefetch distance: keeps ~a dozen cache misses in flight
    for (size_t k = 0; k < n && k < PF; k++) pairs_.prefetch((int32_t)recs[order_[k].idx].first, (int32_t)recs[order_[k].idx].second);
    for (size_t k = 0; k < n; k++) {
      if (k + PF < n) pairs_.prefetch((int32_t)recs[order_[k + PF].idx].first, (int32_t)recs[order_[k + PF].idx].second);
      const Rec &r = recs[order_[k].idx];
      const int32_t f = (int32_t)r.first, s = (int32_t)r.second;
      if (f == cur_a_ && s == cur_b_) continue;  // bpe.cpp:494-496
      PairInfo &p = pairs_.get(f, s);
      if (r.delta < 0) {  // clamp at zero on the NET delta (bpe.cpp:500-509)
        const uint64_t ad = (uint64_t)(-r.delta);
        p.freq = p.freq >= ad ? p.freq - ad : 0;
      } else {
        p.freq += (uint64_t)r.delta;
      }
      if (p.freq >= minf) { p.version++; heap_push(f, s, p.freq, p.version); }  // bpe.cpp:512-515
    }
    finish_merge();
  }

  // Delta-map iteration order (bpe.cpp:30, 41-45, 486-487): bucket = pair_hash % 1024 ascending; inside a
  // bucket entries were prepended, so the most recently first-touched pair comes first. Counting sort by
  // bucket, then each bucket by key descending ((L, new_id) for every L share one bucket: runs can be long).
  void sort_delta_order(const Rec *recs, size_t n) {
    if (n <= 48) {  // short lists (most merges): one insertion sort by (bucket ascending, key descending) beats the 1024-bucket pass
      order_.resize(n);
      for (size_t i = 0; i < n; i++) {
        const uint32_t bk = (uint32_t)recs[i].second & 1023u;
        const uint64_t key = (uint64_t)recs[i].key;
        size_t y = i;
        while (y > 0) {
          const uint32_t pb = (uint32_t)recs[order_[y - 1].idx].second & 1023u;
          if (pb < bk || (pb == bk && order_[y - 1].key >= key)) break;
          order_[y] = order_[y - 1];
          y--;
        }
        order_[y] = KeyIdx{key, (uint32_t)i};
      }
      return;
    }
    uint32_t start[1025];
    memset(start, 0, sizeof start);
    for (size_t i = 0; i < n; i++) start[((uint32_t)recs[i].second & 1023u) + 1]++;
    for (int b = 0; b < 1024; b++) start[b + 1] += start[b];
    order_.resize(n);
    {
      uint32_t fill[1024];
      memcpy(fill, start, sizeof fill);
      for (size_t i = 0; i < n; i++) order_[fill[(uint32_t)recs[i].second & 1023u]++] = KeyIdx{(uint64_t)recs[i].key, (uint32_t)i};
    }
    for (int b = 0; b < 1024; b++) {
      const uint32_t lo = start[b], hi = start[b + 1];
      if (hi - lo < 2) continue;
      if (hi - lo <= 8) {
        for (uint32_t x = lo + 1; x < hi; x++) {
          const KeyIdx v = order_[x];
          uint32_t y = x;
          while (y > lo && order_[y - 1].key < v.key) { order_[y] = order_[y - 1]; y--; }
          order_[y] = v;
        }
      } else {
        std::sort(order_.begin() + lo, order_.begin() + hi, [](const KeyIdx &x, const KeyIdx &y) { return x.key > y.key; });
      }
    }
  }
  void finish_merge() {
    PairInfo &info = pairs_.get(cur_a_, cur_b_);
    info.freq = 0; info.version++;  // bpe.cpp:523-524
    tr_->num_merges++;
    tr_->next_token = 256 + tr_->num_merges;
    pending_ = false;
  }

  // ---- device-table mode (single GPU): the device keeps every pair's frequency and has already applied
  // the deltas; a record carries the pair's NEW frequency and arrives only if the old or the new value
  // reaches min_pair_freq. Pairs below the threshold on both sides can neither be pushed nor invalidate a
  // heap entry, so the host's view (exact for every pair >= min_pair_freq, "< min" otherwise) takes the
  // same decisions as the reference's full table.
  void seed_absolute(const Rec *recs, size_t n) {  // fresh table: records are the pairs with count >= min
    std::vector<Rec> v(recs, recs + n);
    std::sort(v.begin(), v.end(), [](const Rec &x, const Rec &y) { return (uint64_t)x.key < (uint64_t)y.key; });
    for (const Rec &r : v) {
      PairInfo &p = pairs_.get((int32_t)r.first, (int32_t)r.second);
      p.version = 0;
      p.freq = (uint64_t)r.delta;
    }
    push_all_in_table_order();
  }
  void apply_absolute(const Rec *recs, size_t n) {
    if (!pending_) return;
    n_records += n;
    sort_delta_order(recs, n);
    const uint64_t minf = tr_->config.min_pair_freq;
    constexpr size_t PF = 12;
    for (size_t k = 0; k < n && k < PF; k++) pairs_.prefetch((int32_t)recs[order_[k].idx].first, (int32_t)recs[order_[k].idx].second);
    for (size_t k = 0; k < n; k++) {
      if (k + PF < n) pairs_.prefetch((int32_t)recs[order_[k + PF].idx].first, (int32_t)recs[order_[k + PF].idx].second);
      const Rec &r = recs[order_[k].idx];
      const int32_t f = (int32_t)r.first, s = (int32_t)r.second;
      PairInfo &p = pairs_.get(f, s);
      p.freq = (uint64_t)r.delta;
      if (p.freq >= minf) { p.version++; heap_push(f, s, p.freq, p.version); }  // bpe.cpp:512-515
    }
    finish_merge();
  }
  // Rebuilds the table from a dump of the device table (entries already in creation order), keeping the
  // versions the host knows; used when a caller leaves the fresh-count -> merge sequence.
  void rebuild_from_dump(const std::vector<PairInfo> &in_creation_order) {
    PairTable old;
    std::swap(old, pairs_);
    for (const PairInfo &e : in_creation_order) {
      PairInfo &p = pairs_.get(e.first, e.second);
      p.freq = e.freq;
      p.version = old.find_version(e.first, e.second);
    }
  }

  // every pair the host knows with frequency >= min_pair_freq (device-table mode: that is all of them)
  void pairs_at_or_above_min(std::vector<PairInfo> &out) const {
    std::vector<PairInfo> all;
    pairs_.in_creation_order(all);
    out.clear();
    const uint64_t minf = tr_->config.min_pair_freq;
    for (const PairInfo &p : all) if (p.freq >= minf) out.push_back(p);
  }

  const std::vector<PairKey> &merges() const { return merges_; }
  PairTable &pairs() { return pairs_; }
  bool pending() const { return pending_; }
  // The merge announced by next_merge() could not be carried out (device error): take it off the merge list again, so
  // that merge_ops[0 .. num_merges) stays exactly the merges that were applied.
  void abort_pending() {
    if (!pending_) return;
    merges_.pop_back();
    tr_->merge_ops = merges_.data();
    pending_ = false;
  }

 private:
  Trainer *tr_;
  PairTable pairs_;
  std::vector<PairKey> merges_;
  std::vector<Rec> scratch_;
  struct KeyIdx { uint64_t key; uint32_t idx; };
  // frontier of the look-ahead walk, ordered like the heap's pops: frequency descending, then pre-order position
  struct Front { uint64_t freq; size_t pos; };
  // true iff position p comes before position q in a pre-order walk of the implicit binary tree (0-based array indices)
  static bool preorder_before(size_t p, size_t q) {
    const size_t i = p + 1, j = q + 1;  // 1-based: the bits after the leading one are the path from the root, 0 = left
    const int di = 63 - __builtin_clzll((unsigned long long)i), dj = 63 - __builtin_clzll((unsigned long long)j);
    if (di <= dj) { const size_t ja = j >> (dj - di); return ja == i ? di < dj : i < ja; }  // i is an ancestor of j (or further left)
    const size_t ia = i >> (di - dj);
    return ia == j ? false : ia < j;  // j is an ancestor of i: j first
  }
  // std heap comparator ("less" = comes later): the top of the queue is the entry the real heap would pop next
  static bool front_after(const Front &x, const Front &y) {
    if (x.freq != y.freq) return x.freq < y.freq;
    return preorder_before(y.pos, x.pos);
  }
  std::vector<Front> front_;
  std::vector<KeyIdx> order_;
  int32_t cur_a_ = 0, cur_b_ = 0, cur_new_ = 0;
  bool pending_ = false;
};

}  // namespace swb
