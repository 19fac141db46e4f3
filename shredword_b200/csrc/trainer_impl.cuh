// trainer_impl.cuh -- device-side state of one Trainer handle and the orchestration of the kernels.
//
// Phases (reference file:line each one replaces):
//   load   : file/host/device bytes -> HBM -> wt_tokenize -> sort -> keep rule -> row packing
//            (reference csrc/bpe/bpe.cpp:208-297)
//   count  : count_rows/count_long -> pt_emit -> HostCore::seed_counts      (bpe.cpp:315-370)
//   merge  : HostCore::next_merge -> merge_rows/merge_long -> pt_emit -> HostCore::apply (bpe.cpp:391-535)
//   save   : tokfreq_rows -> host writes the two files                       (bpe.cpp:678-739)
#pragma once

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <atomic>
#include <exception>
#include <thread>

#include <chrono>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <vector>

#include "device_util.cuh"
#include "host_core.hpp"
#include "merge_kernels.cuh"
#include "cluster_kernel.cuh"
#include "nccl_dyn.hpp"
#include "word_table.cuh"

namespace swb {

static inline double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
static inline uint64_t pow2_ceil(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

class TrainerImpl {
 public:
  explicit TrainerImpl(Trainer *tr) : tr_(tr), core(tr) {
    for (int i = 0; i < 256; i++) { byte_map[i] = i; keep[i] = 1; }
    memset(&stats, 0, sizeof stats);
  }
  ~TrainerImpl() {
    if (trace_wait_ && !wait_trace_.empty()) {  // development aid: SWB_TRACE_WAIT=1 prints launch->result latency percentiles
      std::vector<float> v = wait_trace_;
      const size_t n = v.size();
      auto pct = [&](std::vector<float> &x, double p) { std::sort(x.begin(), x.end()); return x[(size_t)(p * (x.size() - 1))]; };
      std::vector<float> head(v.begin(), v.begin() + std::min<size_t>(n, 200)), tail(v.begin() + std::min<size_t>(n, 200), v.end());
      fprintf(stderr, "[trace] merges=%zu first200: p50=%.1f p90=%.1f max=%.1f us", n, pct(head, 0.5), pct(head, 0.9), pct(head, 1.0));
      if (!tail.empty()) fprintf(stderr, " | rest: p10=%.1f p50=%.1f p90=%.1f p99=%.1f max=%.1f us", pct(tail, 0.1), pct(tail, 0.5), pct(tail, 0.9), pct(tail, 0.99), pct(tail, 1.0));
      double sh = 0, st = 0; for (size_t i = 0; i < n; i++) (i < 200 ? sh : st) += wait_trace_[i];
      fprintf(stderr, " | sum first200=%.1f ms rest=%.1f ms\n", sh / 1e3, st / 1e3);
      if (tw_n_) fprintf(stderr, "[trace]   block0/warp0 on quiet merges: sig %.2f us, candidate rows %.2f us (%.1f candidates of 32), fence+sync %.2f us\n", tw_sig_ / tw_n_, tw_rows_ / tw_n_, tw_cand_ / tw_n_, tw_fence_ / tw_n_);
      if (trace_cand_.size() == n) {  // birth-log statistics: candidates per merge
        const uint32_t ce[] = {0, 1, 9, 33, 129, 513, 2049, 8193, 0xFFFFFFFFu};
        size_t base_pairs = 0;
        for (size_t i = 0; i < n; i++) base_pairs += trace_logn_[i] == ~0ull;
        fprintf(stderr, "[trace]   merges without a birth log (two initial symbols): %zu\n", base_pairs);
        for (int b = 0; b + 1 < 9; b++) {
          size_t c = 0; double lat = 0, logn = 0, mr = 0, sc = 0;
          for (size_t i = 0; i < n; i++) if (trace_logn_[i] != ~0ull && trace_cand_[i] >= ce[b] && trace_cand_[i] < ce[b + 1]) { c++; lat += wait_trace_[i]; logn += (double)trace_logn_[i]; mr += trace_mrows_[i]; sc += trace_scan_[i]; }
          if (c) fprintf(stderr, "[trace]   candidates in [%u, %u): %zu merges, mean log entries %.0f, matched rows %.1f, latency %.1f us (scan %.1f)\n", ce[b], ce[b + 1], c, logn / c, mr / c, lat / c, sc / c);
        }
      }
      if (trace_removed_.size() == n) {
        const uint32_t edges[] = {0, 10, 30, 100, 300, 1000, 3000, 10000, 100000, 0xFFFFFFFFu};
        for (int b = 0; b + 1 < 10; b++) {
          double sum = 0; size_t c = 0;
          for (size_t i = 200; i < n; i++) if (trace_removed_[i] >= edges[b] && trace_removed_[i] < edges[b + 1]) { sum += wait_trace_[i]; c++; }
          double ssc = 0, stl = 0, snr = 0;
          if (trace_scan_.size() == n)
            for (size_t i = 200; i < n; i++) if (trace_removed_[i] >= edges[b] && trace_removed_[i] < edges[b + 1]) { ssc += trace_scan_[i]; stl += trace_tail_[i]; snr += trace_nrec_[i]; }
          if (c) fprintf(stderr, "[trace]   matches in [%u, %u): %zu merges, mean latency %.1f us (kernel: scan %.1f us, tail %.1f us, touched pairs %.0f)\n", edges[b], edges[b + 1], c, sum / c, ssc / c, stl / c, snr / c);
        }
      }
    }
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (stream_) cudaStreamDestroy(stream_);
  }

  HostCore core;
  SwbStats stats;
  int32_t byte_map[256];  // byte -> initial id as the caller sees it (unk_id for dropped bytes)
  uint8_t keep[256];
  int rank = 0, nranks = 1;
  // Multi-GPU, replicated merge loop (default): the load is range-split and the word tables are exchanged (NCCL), then every
  // rank holds ALL unique words and runs the single-GPU resident merge loop on them -- no per-merge collective. The merge
  // loop is index-driven and latency-bound (~25 us per merge on one GPU); sharding it adds one NCCL all-gather per merge
  // (measured 62 us per merge at 2 GPUs) and shortens nothing. replicated_ = false keeps the sharded loop (word wi on rank
  // wi % nranks, per-merge all-gather of the delta records).
  bool replicated_ = true;
  int mrank() const { return replicated_ ? 0 : rank; }
  int mnranks() const { return replicated_ ? 1 : nranks; }
  NcclComm mcomm() const { return replicated_ ? nullptr : comm_; }
  bool timing = false;
  uint64_t W = 0;  // unique words (global)
  std::vector<uint64_t> h_counts;  // host mirror of the word counts (Corpus.word_counts)

  // ---------------------------------------------------------------- device bring-up
  // Per-process, per-device facts (cudaGetDeviceProperties, occupancy and attribute calls take the driver's
  // global lock and cost from 1 to 100+ ms each while other work is in flight): queried once, not per handle.
  struct DeviceFacts { int sms = 0; bool coop_ok = false; };
  static const DeviceFacts &device_facts(int device) {
    static std::mutex mu;
    static std::map<int, DeviceFacts> facts;
    std::lock_guard<std::mutex> g(mu);
    auto it = facts.find(device);
    if (it != facts.end()) return it->second;
    DeviceFacts f;
    SWB_CUDA(cudaDeviceGetAttribute(&f.sms, cudaDevAttrMultiProcessorCount, device));
    SWB_CUDA(cudaFuncSetAttribute(wt_tokenize, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES));
    int coop = 0;
    SWB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    f.coop_ok = coop != 0;
    return facts.emplace(device, f).first->second;
  }
  void ensure_device() {
    if (stream_) return;
    const double t_init0 = now_ms();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw Error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                  " (this library has no CPU fallback)");
    SWB_CUDA(cudaGetDevice(&device_));
    const DeviceFacts &f = device_facts(device_);
    sms_ = f.sms; coop_ok_ = f.coop_ok;
    SWB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    SWB_CUDA(cudaEventCreate(&ev0_));
    SWB_CUDA(cudaEventCreate(&ev1_));
    scalars_.alloc(16);
    SWB_CUDA(cudaMemsetAsync(scalars_.get(), 0, scalars_.bytes(), stream_));
    hdr_.alloc(HDR_WORDS);
    memset(hdr_.host(), 0, HDR_WORDS * sizeof(unsigned long long));
    if (getenv("SWB_TRACE_INIT")) fprintf(stderr, "[trace] ensure_device %.3f ms\n", now_ms() - t_init0);
  }
  void sync() { SWB_CUDA(cudaStreamSynchronize(stream_)); }
  void launched(uint64_t n = 1) { stats.kernel_launches += n; }
  int32_t unk_dev() const { return tr_->config.unk_id < 0 ? UNK_CODE : tr_->config.unk_id; }

  // ---------------------------------------------------------------- load
  // Copies host bytes into a padded device buffer (double-buffered through pinned staging when the
  // source is a file) and builds the word table.
  void load_file(const char *path) {
    ensure_device();
    const double t0 = now_ms();
    FILE *f = fopen(path, "rb");
    if (!f) throw Error(std::string("Couldn't open file: ") + path);
    fseeko(f, 0, SEEK_END);
    const uint64_t n = (uint64_t)ftello(f);
    fseeko(f, 0, SEEK_SET);
    DevBuf<uint8_t> corpus(n + 64);
    const size_t CH = 64ull << 20;
    PinnedBuf<uint8_t> stage[2];
    cudaEvent_t done[2];
    for (int i = 0; i < 2; i++) { stage[i].alloc(std::min<uint64_t>(CH, n ? n : 1)); SWB_CUDA(cudaEventCreate(&done[i])); }
    uint64_t pos = 0;
    int b = 0;
    bool ok = true;
    while (pos < n) {
      const size_t want = (size_t)std::min<uint64_t>(CH, n - pos);
      SWB_CUDA(cudaEventSynchronize(done[b]));  // staging buffer free again
      const size_t got = fread(stage[b].host(), 1, want, f);
      if (got != want) { ok = false; break; }
      SWB_CUDA(cudaMemcpyAsync(corpus.get() + pos, stage[b].host(), got, cudaMemcpyHostToDevice, stream_));
      SWB_CUDA(cudaEventRecord(done[b], stream_));
      pos += got;
      b ^= 1;
    }
    fclose(f);
    sync();
    for (int i = 0; i < 2; i++) cudaEventDestroy(done[i]);
    if (!ok) throw Error(std::string("short read on ") + path);
    SWB_CUDA(cudaMemsetAsync(corpus.get() + n, ' ', 64, stream_));
    build_word_table(corpus, n);
    stats.load_ms += now_ms() - t0;
  }
  void load_host(const void *data, uint64_t n) {
    ensure_device();
    const double t0 = now_ms();
    DevBuf<uint8_t> corpus(n + 64);
    // (development / test switches, read per load: SWB_NO_LOAD_PIPELINE=1, SWB_LOAD_PIECE=<bytes per piece>)
    const bool no_pipe = getenv("SWB_NO_LOAD_PIPELINE") && atoi(getenv("SWB_NO_LOAD_PIPELINE")) > 0;
    if (no_pipe || n < 4 * load_piece_bytes()) {
      if (n) SWB_CUDA(cudaMemcpyAsync(corpus.get(), data, n, cudaMemcpyHostToDevice, stream_));
      SWB_CUDA(cudaMemsetAsync(corpus.get() + n, ' ', 64, stream_));
      build_word_table(corpus, n);
    } else {  // large host buffers: the tokeniser follows the copy piece by piece (see build_word_table)
      SWB_CUDA(cudaMemsetAsync(corpus.get() + n, ' ', 64, stream_));
      build_word_table(corpus, n, false, 0, static_cast<const uint8_t *>(data));
    }
    stats.load_ms += now_ms() - t0;
  }
  static uint64_t load_piece_bytes() {
    const char *e = getenv("SWB_LOAD_PIECE");
    const uint64_t v = e ? strtoull(e, nullptr, 10) : 0;
    return v >= 4096 ? v / 16 * 16 : (64ull << 20);
  }
  void load_shard(const void *data, uint64_t n, uint64_t global_offset, bool on_device) {
    ensure_device();
    const double t0 = now_ms();
    DevBuf<uint8_t> corpus(n + 64);
    const bool no_pipe = getenv("SWB_NO_LOAD_PIPELINE") && atoi(getenv("SWB_NO_LOAD_PIPELINE")) > 0;
    const bool piped = !on_device && !no_pipe && n >= 4 * load_piece_bytes();  // (host buffers: copy and tokeniser overlap, see load_host)
    if (n && !piped) SWB_CUDA(cudaMemcpyAsync(corpus.get(), data, n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream_));
    SWB_CUDA(cudaMemsetAsync(corpus.get() + n, ' ', 64, stream_));
    build_word_table(corpus, n, nranks > 1, global_offset, piped ? static_cast<const uint8_t *>(data) : nullptr);
    stats.load_ms += now_ms() - t0;
  }
  void load_device(const void *d_data, uint64_t n) {
    ensure_device();
    const double t0 = now_ms();
    DevBuf<uint8_t> corpus(n + 64);
    if (n) SWB_CUDA(cudaMemcpyAsync(corpus.get(), d_data, n, cudaMemcpyDeviceToDevice, stream_));
    SWB_CUDA(cudaMemsetAsync(corpus.get() + n, ' ', 64, stream_));
    build_word_table(corpus, n);
    stats.load_ms += now_ms() - t0;
  }

  // corpus: n bytes + >= 16 bytes of ' ' padding, 16-byte aligned (cudaMalloc)
  // split: `corpus` is only this rank's byte range of the global corpus (starting at global_offset); the
  // per-rank tables are exchanged over NCCL and merged, so every rank ends with the global word table.
  // host_src != nullptr: the n corpus bytes are still in host memory; they are copied in pieces on a second stream and
  // every piece is tokenised as soon as it has landed (PCIe and the tokeniser run at about the same rate: overlapped,
  // the load takes the longer of the two instead of their sum).
  void build_word_table(DevBuf<uint8_t> &corpus, uint64_t n, bool split = false, uint64_t global_offset = 0,
                        const uint8_t *host_src = nullptr) {
    if (n >= (1ull << 40) || global_offset + n >= (1ull << 40)) throw Error("corpus larger than 1 TiB is not supported (40-bit offsets)");
    if (split && !comm_) throw Error("a range-split load needs swb_dist_init first");
    static const bool trace_load = getenv("SWB_TRACE_INIT") != nullptr;
    double t_tr = now_ms();
    auto tr = [&](const char *what) {  // (SWB_TRACE_INIT=1: where a load spends its host time)
      if (!trace_load) return;
      const double t = now_ms();
      fprintf(stderr, "[trace] load: %-28s %9.3f ms\n", what, t - t_tr);
      t_tr = t;
    };
    free_corpus_state();
    tr("free previous state");
    unsigned int *d_nuniq = scalars_.get() + 0, *d_flags = scalars_.get() + 1, *d_cursor = scalars_.get() + 2,
                 *d_nlong = scalars_.get() + 3;
    unsigned long long *d_u64 = reinterpret_cast<unsigned long long *>(scalars_.get() + 8);  // [0]=long syms [1]=bytes [2]=cursor
    // ---- 1. tokenise + dedupe
    // Unique words grow far slower than the corpus (Heaps' law): the table starts at one 32-byte slot per 256 corpus bytes, at
    // most 16 M slots (512 MB; only the occupied 32-byte sectors are ever touched again, and for a few million unique words those mostly stay in the 126 MB L2), and is rebuilt
    // four times as large if the corpus turns out to hold more unique words than 60 % of that.
    uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(1ull << 20, pow2_ceil(n / 256)), 1ull << 24);  // (small corpora hold far more unique words per byte: at least 1 M slots = 32 MB)
    {  // SWB_TEST_WT_CAP=<slots> (tests): start with a table that is too small, so that the grow-and-run-again path is taken
      static const uint64_t test_cap = getenv("SWB_TEST_WT_CAP") ? strtoull(getenv("SWB_TEST_WT_CAP"), nullptr, 10) : 0;
      if (test_cap) cap = std::max<uint64_t>(64, pow2_ceil(test_cap));
    }
    DevBuf<WSlot> wslots;
    unsigned int h_scal[4];
    const int tok_grid = (int)std::max<uint64_t>((uint64_t)sms_, (n >> 31) + 1);  // one block per SM; a block's span stays below 4 GB
    for (;;) {
      wslots.alloc(cap);
      wt_fill<<<sms_ * 8, 256, 0, stream_>>>(wslots.get(), cap); launched();
      SWB_CUDA(cudaMemsetAsync(scalars_.get(), 0, scalars_.bytes(), stream_));
      WordTableDev tbl{wslots.get(), cap - 1, d_nuniq, d_flags, (uint64_t)(cap * 0.6)};
      SWB_CUDA(cudaEventRecord(ev0_, stream_));
      if (n && host_src) {
        const uint64_t PIECE = load_piece_bytes();  // (a multiple of 16: see wt_tokenize)
        if (!copy_stream_) SWB_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        cudaEvent_t fill_done;
        SWB_CUDA(cudaEventCreateWithFlags(&fill_done, cudaEventDisableTiming));
        SWB_CUDA(cudaEventRecord(fill_done, stream_));
        SWB_CUDA(cudaStreamWaitEvent(copy_stream_, fill_done, 0));  // (also orders the copies after the allocation's previous users)
        std::vector<cudaEvent_t> landed;
        uint64_t tok_lo = 0;
        for (uint64_t c0 = 0; c0 < n; c0 += PIECE) {
          const uint64_t c1 = std::min(n, c0 + PIECE);
          SWB_CUDA(cudaMemcpyAsync(corpus.get() + c0, host_src + c0, c1 - c0, cudaMemcpyHostToDevice, copy_stream_));
          cudaEvent_t ev;
          SWB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          SWB_CUDA(cudaEventRecord(ev, copy_stream_));
          landed.push_back(ev);
          // words starting before tok_hi end before it: cut after the last delimiter of what has been copied
          uint64_t tok_hi = c1;
          if (c1 < n) while (tok_hi > tok_lo && !is_delim(host_src[tok_hi - 1])) tok_hi--;
          SWB_CUDA(cudaStreamWaitEvent(stream_, ev, 0));
          if (tok_hi > tok_lo) {
            wt_tokenize<<<tok_grid, WT_THREADS, WT_SMEM_BYTES, stream_>>>(corpus.get(), n, tbl, tok_lo, tok_hi); launched();
            tok_lo = tok_hi;
          }
        }
        SWB_CUDA(cudaGetLastError());
        sync();
        for (cudaEvent_t ev : landed) cudaEventDestroy(ev);
        cudaEventDestroy(fill_done);
        host_src = nullptr;  // (a rerun with a larger table finds the corpus on the device)
      } else if (n) { wt_tokenize<<<tok_grid, WT_THREADS, WT_SMEM_BYTES, stream_>>>(corpus.get(), n, tbl, 0, n); launched(); }
      SWB_CUDA(cudaGetLastError());
      SWB_CUDA(cudaEventRecord(ev1_, stream_));
      SWB_CUDA(cudaMemcpyAsync(h_scal, scalars_.get(), sizeof h_scal, cudaMemcpyDeviceToHost, stream_));
      sync();
      {
        float tms = 0;
        SWB_CUDA(cudaEventElapsedTime(&tms, ev0_, ev1_));
        stats.tokenize_ms = tms; stats.tokenize_bytes = n;
      }
      if (h_scal[1] & 2u) throw Error("NUL byte in corpus: outside the parity domain (the reference drops a libc-buffer-dependent span)");
      if (!(h_scal[1] & 1u)) break;
      cap *= 4;  // more unique words than expected: bigger table, run again
    }
    tr("tokenise (all tries)");
    W = h_scal[0];
    DevBuf<uint64_t> woff;
    if (!split) {
      if (W >= (1ull << 31)) throw Error("more than 2^31 unique words");
      // ---- 2. reference word order: (djb2 bucket, first occurrence)
      DevBuf<unsigned long long> skeys(W), scnt(W), skeys2(W);
      cnt_.alloc(W);
      woff.alloc(W);
      if (W) {
        wt_compact<<<sms_ * 8, 256, 0, stream_>>>(wslots.get(), cap, skeys.get(), scnt.get(), d_cursor); launched();
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, skeys.get(), skeys2.get(), scnt.get(), cnt_.get(), (int64_t)W, 0, 52, stream_);
        DevBuf<uint8_t> tmp(tmp_bytes);
        SWB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, skeys.get(), skeys2.get(), scnt.get(), cnt_.get(), (int64_t)W, 0, 52, stream_));
        launched(8);
        wt_unpack_sorted<<<sms_ * 4, 256, 0, stream_>>>(skeys2.get(), W, woff.get()); launched();
        sync();
      }
      wslots.release();
    } else {
      // ---- 2'. export this rank's unique words, all-gather, merge into the global table
      NcclApi &api = NcclApi::get();
      const uint64_t Wl = W;
      DevBuf<unsigned long long> skeys(Wl), scnt(Wl), len1(Wl + 1), aoff(Wl + 1);
      unsigned long long arena_bytes = 0;
      if (Wl) {
        wt_compact<<<sms_ * 8, 256, 0, stream_>>>(wslots.get(), cap, skeys.get(), scnt.get(), d_cursor); launched();
        wt_local_lens<<<sms_ * 8, 256, 0, stream_>>>(corpus.get(), n, skeys.get(), Wl, len1.get()); launched();
        SWB_CUDA(cudaMemsetAsync(len1.get() + Wl, 0, 8, stream_));
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, len1.get(), aoff.get(), (int64_t)(Wl + 1), stream_);
        DevBuf<uint8_t> tmp(tb);
        SWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tb, len1.get(), aoff.get(), (int64_t)(Wl + 1), stream_));
        launched(2);
        SWB_CUDA(cudaMemcpyAsync(&arena_bytes, aoff.get() + Wl, 8, cudaMemcpyDeviceToHost, stream_));
        sync();
      }
      wslots.release();
      const double t_ex0 = now_ms();  // (the stream is idle here: the exchange's host time is its device time)
      DevBuf<unsigned long long> d_sizes(2 * (size_t)nranks);
      const unsigned long long mine[2] = {Wl, arena_bytes};
      SWB_CUDA(cudaMemcpyAsync(d_sizes.get() + 2 * rank, mine, 16, cudaMemcpyHostToDevice, stream_));
      api.check(api.AllGather(d_sizes.get() + 2 * rank, d_sizes.get(), 16, NcclApi::kChar, comm_, stream_), "ncclAllGather(sizes)");
      std::vector<unsigned long long> h_sizes(2 * (size_t)nranks);
      SWB_CUDA(cudaMemcpyAsync(h_sizes.data(), d_sizes.get(), h_sizes.size() * 8, cudaMemcpyDeviceToHost, stream_));
      sync();
      uint64_t maxW = 1, maxA = 16, sumW = 0;
      for (int r = 0; r < nranks; r++) { maxW = std::max<uint64_t>(maxW, h_sizes[2 * r]); maxA = std::max<uint64_t>(maxA, h_sizes[2 * r + 1]); sumW += h_sizes[2 * r]; }
      maxA = (maxA + 63) / 64 * 64;
      const uint64_t base_n = maxA * (uint64_t)nranks;
      if (base_n >= (1ull << 40)) throw Error("exchanged word arenas exceed 1 TiB");
      DevBuf<uint8_t> arena_all(base_n + 64);
      DevBuf<WordMeta> meta_all(maxW * (uint64_t)nranks);
      SWB_CUDA(cudaMemsetAsync(arena_all.get(), ' ', base_n + 64, stream_));
      if (Wl) {
        wt_export<<<sms_ * 8, 256, 0, stream_>>>(corpus.get(), skeys.get(), scnt.get(), len1.get(), aoff.get(), Wl, global_offset,
                                                arena_all.get() + maxA * rank, meta_all.get() + maxW * rank);
        launched();
      }
      api.check(api.AllGather(arena_all.get() + maxA * rank, arena_all.get(), maxA, NcclApi::kChar, comm_, stream_), "ncclAllGather(arena)");
      api.check(api.AllGather(meta_all.get() + maxW * rank, meta_all.get(), maxW * sizeof(WordMeta), NcclApi::kChar, comm_, stream_), "ncclAllGather(meta)");
      stats.collectives += 3;
      stats.exchange_bytes += (maxA + maxW * sizeof(WordMeta)) * (uint64_t)nranks;
      skeys.release(); scnt.release(); len1.release(); aoff.release();
      // global table
      uint64_t gcap = std::max<uint64_t>(1ull << 16, pow2_ceil(2 * sumW + 16));
      DevBuf<unsigned long long> gkeys(gcap), gcounts(gcap), gfirst(gcap);
      wt_fill3<<<sms_ * 8, 256, 0, stream_>>>(gkeys.get(), gcounts.get(), gfirst.get(), gcap); launched();
      SWB_CUDA(cudaMemsetAsync(scalars_.get(), 0, scalars_.bytes(), stream_));
      WordTableSoA gtbl{gkeys.get(), gcounts.get(), gcap - 1, d_nuniq, d_flags, gcap};
      wt_insert_words<<<sms_ * 8, 256, 0, stream_>>>(arena_all.get(), base_n, meta_all.get(), d_sizes.get(), nranks, maxW, maxA, gtbl, gfirst.get());
      launched();
      SWB_CUDA(cudaMemcpyAsync(h_scal, scalars_.get(), sizeof h_scal, cudaMemcpyDeviceToHost, stream_));
      sync();
      stats.exchange_ns += (uint64_t)((now_ms() - t_ex0) * 1e6);
      W = h_scal[0];
      if (W >= (1ull << 31)) throw Error("more than 2^31 unique words");
      DevBuf<unsigned long long> sk(W), sv(W), sk2(W), sv2(W);
      cnt_.alloc(W);
      woff.alloc(W);
      if (W) {
        wt_compact_dist<<<sms_ * 8, 256, 0, stream_>>>(gkeys.get(), gfirst.get(), gcap, sk.get(), sv.get(), d_cursor); launched();
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, sk.get(), sk2.get(), sv.get(), sv2.get(), (int64_t)W, 0, 52, stream_);
        DevBuf<uint8_t> tmp(tmp_bytes);
        SWB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, sk.get(), sk2.get(), sv.get(), sv2.get(), (int64_t)W, 0, 52, stream_));
        launched(8);
        wt_after_sort_dist<<<sms_ * 4, 256, 0, stream_>>>(sv2.get(), gkeys.get(), gcounts.get(), W, woff.get(), cnt_.get()); launched();
        sync();
      }
      // from here on the "corpus" is the concatenation of the exchanged arenas
      corpus = std::move(arena_all);
      n = base_n;
    }
    tr("compact + sort");
    // ---- 3. lengths, byte histogram, long words
    DevBuf<uint32_t> wlen(W);
    long_index_.alloc(W);
    DevBuf<unsigned long long> hist(256);
    SWB_CUDA(cudaMemsetAsync(hist.get(), 0, hist.bytes(), stream_));
    if (W) {
      wt_word_info<<<std::min<uint64_t>(sms_ * 8, (W + 255) / 256), 256, 0, stream_>>>(
          corpus.get(), n, woff.get(), W, wlen.get(), hist.get(), d_nlong, d_u64 + 0, long_index_.get(), d_u64 + 1);
      launched();
    }
    unsigned long long h_hist[256], h_u64[3];
    SWB_CUDA(cudaMemcpyAsync(h_hist, hist.get(), sizeof h_hist, cudaMemcpyDeviceToHost, stream_));
    SWB_CUDA(cudaMemcpyAsync(h_scal, scalars_.get(), sizeof h_scal, cudaMemcpyDeviceToHost, stream_));
    SWB_CUDA(cudaMemcpyAsync(h_u64, d_u64, sizeof h_u64, cudaMemcpyDeviceToHost, stream_));
    sync();
    n_long_ = h_scal[3];
    const uint64_t long_total = h_u64[0];
    word_bytes_total_ = h_u64[1];
    tr("word info");
    // ---- 4. character coverage rule on the host (256 values; reference bpe.cpp:257-279)
    apply_keep_rule(h_hist);
    DevBuf<int32_t> d_bmap(256);
    int32_t dev_map[256];
    for (int i = 0; i < 256; i++) dev_map[i] = keep[i] ? i : unk_dev();
    SWB_CUDA(cudaMemcpyAsync(d_bmap.get(), dev_map, sizeof dev_map, cudaMemcpyHostToDevice, stream_));
    // ---- 5. pack this rank's words into rows
    const uint64_t n_batches = (W + PACK_BATCH - 1) / PACK_BATCH;
    wloc_.alloc(W);
    n_rows_ = 0;
    if (W) {
      DevBuf<uint32_t> batch_rows(n_batches);
      const int pgrid = (int)std::min<uint64_t>(sms_ * 4, (n_batches + 7) / 8);
      wt_pack<false><<<pgrid, 256, 0, stream_>>>(corpus.get(), woff.get(), wlen.get(), long_index_.get(), W, mrank(), mnranks(),
                                                 d_bmap.get(), batch_rows.get(), nullptr, nullptr, nullptr);
      launched();
      std::vector<uint32_t> h_rows(n_batches);
      SWB_CUDA(cudaMemcpyAsync(h_rows.data(), batch_rows.get(), n_batches * 4, cudaMemcpyDeviceToHost, stream_));
      sync();
      uint64_t acc = 0;
      for (uint64_t i = 0; i < n_batches; i++) { const uint32_t c = h_rows[i]; h_rows[i] = (uint32_t)acc; acc += c; }
      if (acc >= (1ull << 32)) throw Error("too many rows");
      n_rows_ = acc;
      rows_.alloc(n_rows_ * (ROW / 4));
      sig_.alloc(n_rows_ * SIG_WORDS);
      SWB_CUDA(cudaMemcpyAsync(batch_rows.get(), h_rows.data(), n_batches * 4, cudaMemcpyHostToDevice, stream_));
      wt_pack<true><<<pgrid, 256, 0, stream_>>>(corpus.get(), woff.get(), wlen.get(), long_index_.get(), W, mrank(), mnranks(),
                                                d_bmap.get(), batch_rows.get(), rows_.get(), wloc_.get(), sig_.get());
      launched();
      sync();
    }
    tr("pack rows");
    // ---- 6. long words
    long_off_.alloc(n_long_); long_len_.alloc(n_long_); long_word_.alloc(n_long_); long_syms_.alloc(long_total);
    if (n_long_) {
      wt_long_offsets<<<sms_ * 4, 256, 0, stream_>>>(wlen.get(), long_index_.get(), W, d_u64 + 2, long_off_.get());
      wt_fill_long<<<sms_ * 4, 256, 0, stream_>>>(corpus.get(), woff.get(), wlen.get(), long_index_.get(), W, mrank(), mnranks(),
                                                  d_bmap.get(), long_off_.get(), long_syms_.get(), long_len_.get(), long_word_.get());
      launched(2);
    }
    tr("long words");
    // ---- 7. keep the bytes of the unique words (accessors, encoder pin checks); release the corpus
    word_boff_.alloc(W + 1);
    word_bytes_.alloc(word_bytes_total_);
    if (W) {
      cub::TransformInputIterator<unsigned long long, CastU32ToU64, const uint32_t *> it(wlen.get(), CastU32ToU64());
      size_t tmp_bytes = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, (unsigned long long *)word_boff_.get(), (int64_t)W, stream_);
      DevBuf<uint8_t> tmp(tmp_bytes);
      SWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, it, (unsigned long long *)word_boff_.get(), (int64_t)W, stream_));
      launched(2);
      wt_gather_bytes<<<sms_ * 8, 256, 0, stream_>>>(corpus.get(), woff.get(), wlen.get(), word_boff_.get(), W, word_bytes_.get());
      launched();
      sync();
    }
    SWB_CUDA(cudaMemcpyAsync(word_boff_.get() + W, &word_bytes_total_, 8, cudaMemcpyHostToDevice, stream_));
    h_counts.resize(W);
    if (W) SWB_CUDA(cudaMemcpyAsync(h_counts.data(), cnt_.get(), W * 8, cudaMemcpyDeviceToHost, stream_));
    sync();
    // live symbols of this rank
    live_symbols_ = 0;
    if (mnranks() == 1) live_symbols_ = word_bytes_total_;
    else {
      std::vector<uint32_t> h_len(W);
      if (W) SWB_CUDA(cudaMemcpy(h_len.data(), wlen.get(), W * 4, cudaMemcpyDeviceToHost));
      for (uint64_t w = mrank(); w < W; w += mnranks()) live_symbols_ += h_len[w];
    }
    tr("word bytes + counts");
    corpus.release();
    SWB_CUDA(cudaMemsetAsync(scalars_.get(), 0, scalars_.bytes(), stream_));
    setup_birth_log();
    tr("birth log / occurrence index");
    // public mirror fields (reference Corpus)
    tr_->corpus.vocab_size = W;
    tr_->corpus.word_counts = h_counts.data();
    tr_->corpus.words = reinterpret_cast<void **>(this);  // opaque, non-NULL
    stats.words = W; stats.rows = n_rows_; stats.long_words = n_long_; stats.live_symbols = live_symbols_;
    loaded_ = true;
    // reference bpe.cpp:295: a fresh pair table after every load
    core.pairs().clear();
    reset_global_table();
    device_tables_ = false;
    tables_fresh_ = true;
    gt_used_estimate_ = 0;
    tr("host tables reset");
  }

  // reference bpe.cpp:262-279 + histogram.cpp:47-53: bytes listed in the bucket
  // order of a 256-bucket djb2 map of 1-byte strings, stable-sorted by count descending, the first
  // (size_t)(c * coverage) kept -- the product is taken in float, as in the reference.
  void apply_keep_rule(const unsigned long long *hist) {
    uint8_t order[256];
    size_t c = 0;
    for (int b = 0; b < 256; b++) {
      const uint8_t byte = (uint8_t)((b - 165) & 255);
      if (hist[byte]) order[c++] = byte;
    }
    std::stable_sort(order, order + c, [&](uint8_t x, uint8_t y) { return hist[x] > hist[y]; });
    const size_t nkeep = (size_t)((float)c * tr_->config.character_coverage);
    memset(keep, 0, sizeof keep);
    for (size_t i = 0; i < nkeep && i < c; i++) keep[order[i]] = 1;
    for (int i = 0; i < 256; i++) byte_map[i] = keep[i] ? i : tr_->config.unk_id;
    if (core.log_level > 0) printf("[DEBUG]\t Character histogram built with %zu unique characters.\n", c);
  }

  // ---------------------------------------------------------------- birth log
  // A match removes one symbol and logs at most two entries; a word of L symbols can match at most L-1
  // times over a whole training run, so 2 * (symbols - words) entries can never overflow.
  void setup_birth_log() {
    static const bool off = getenv("SWB_NO_BIRTH_LOG") && atoi(getenv("SWB_NO_BIRTH_LOG")) > 0;
    log_ok_ = false; stream_merges_ = 0; log_cap_ = 0;
    log_start_h_.assign(1, 0u);
    if (off || !n_rows_) return;
    const uint64_t cap = 2 * live_symbols_ + 64;
    if (cap >= (1ull << 31)) return;  // (entries are indexed with 32 bits; such a table falls back to the signature scan)
    log_ent_.alloc(cap);
    log_cap_ = (uint32_t)cap;
    log_scal_.alloc(4);
    SWB_CUDA(cudaMemsetAsync(log_scal_.get(), 0, log_scal_.bytes(), stream_));
    log_merge_cap_ = 0;
    ensure_log_merges(tr_->config.target_vocab_size > 256 ? (uint32_t)(tr_->config.target_vocab_size - 256) : 1024u);
    log_ok_ = true;
    build_initial_pair_index();
  }
  // Occurrence index of the pairs of two initial symbols (BirthLogDev::ip_*): count per pair, prefix sum, fill. Only where
  // the resident kernel can run (one GPU's worth of words in rows, no long words); a few ms once per load against ~30 us
  // saved on every later merge of such a pair (the whole-grid row scan is the alternative: the row signatures cannot
  // tell rows apart for common letters).
  void build_initial_pair_index() {
    ip_ent_.release(); ip_start_.release();
    static const bool off = getenv("SWB_NO_IP_INDEX") && atoi(getenv("SWB_NO_IP_INDEX")) > 0;
    if (off || !log_ok_ || n_long_ || mnranks() != 1 || !W || W >= (1ull << 31)) return;
    DevBuf<unsigned int> cnt(IP_PAIRS + 1);
    ip_start_.alloc(IP_PAIRS + 1);
    SWB_CUDA(cudaMemsetAsync(cnt.get(), 0, cnt.bytes(), stream_));
    const int grid = (int)std::min<uint64_t>((uint64_t)sms_ * 8, (W + 255) / 256);
    const int32_t *flat = reinterpret_cast<const int32_t *>(rows_.get());
    ip_index_words<false><<<grid, 256, 0, stream_>>>(flat, wloc_.get(), long_index_.get(), (uint32_t)W, cnt.get(), nullptr, nullptr); launched();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.get(), ip_start_.get(), (int)(IP_PAIRS + 1), stream_);
    DevBuf<uint8_t> tmp(tmp_bytes);
    SWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, cnt.get(), ip_start_.get(), (int)(IP_PAIRS + 1), stream_));
    launched(2);
    unsigned int total = 0;
    SWB_CUDA(cudaMemcpyAsync(&total, ip_start_.get() + IP_PAIRS, 4, cudaMemcpyDeviceToHost, stream_));
    sync();
    if (!total) { ip_start_.release(); return; }
    ip_ent_.alloc(total);
    SWB_CUDA(cudaMemsetAsync(cnt.get(), 0, cnt.bytes(), stream_));
    ip_index_words<true><<<grid, 256, 0, stream_>>>(flat, wloc_.get(), long_index_.get(), (uint32_t)W, cnt.get(), ip_start_.get(), ip_ent_.get()); launched();
    SWB_CUDA(cudaGetLastError());
    sync();  // (cnt and tmp go back to the block cache here)
  }
  void ensure_log_merges(uint32_t merges) {  // room for start[0 .. merges]
    if (merges + 2 <= log_merge_cap_) return;
    const uint32_t cap = (uint32_t)pow2_ceil((uint64_t)merges + 2);
    DevBuf<unsigned int> ns(cap);
    SWB_CUDA(cudaMemsetAsync(ns.get(), 0, ns.bytes(), stream_));
    if (log_merge_cap_) SWB_CUDA(cudaMemcpyAsync(ns.get(), log_start_.get(), (size_t)log_merge_cap_ * 4, cudaMemcpyDeviceToDevice, stream_));
    sync();
    log_start_ = std::move(ns);
    log_merge_cap_ = cap;
  }
  void merge_done_on_stream(unsigned long long hdr_flags) {
    stream_merges_++;
    if (log_ok_) log_start_h_.push_back((uint32_t)(hdr_flags >> 32));
  }
  unsigned long long log_cursor_h() const { return (log_ok_ && !log_start_h_.empty()) ? log_start_h_.back() : 0ull; }
  // lo << 32 | n of the log that holds the births of (a, b) (device ids), ~0 when the pair has none
  unsigned long long log_range_of(int32_t da, int32_t db) const {
    const int32_t newer = da > db ? da : db;
    if (!log_ok_ || newer < 256 || (uint32_t)(newer - 256) >= stream_merges_ || (size_t)(newer - 256) + 1 >= log_start_h_.size()) return ~0ull;
    const uint32_t lo = log_start_h_[newer - 256], hi = log_start_h_[newer - 256 + 1];
    return ((unsigned long long)lo << 32) | (unsigned long long)(hi - lo);
  }
  // the log as the next merge (which creates new_id) sees it; tokens must be numbered 256 + merges done on this stream
  BirthLogDev birth_log(int32_t new_id, uint32_t merges_ahead) {
    BirthLogDev lg;
    memset(&lg, 0, sizeof lg);
    if (log_ok_ && (int64_t)new_id != 256 + (int64_t)stream_merges_) log_ok_ = false;
    if (!log_ok_) return lg;
    ensure_log_merges(stream_merges_ + merges_ahead);
    lg.ent = log_ent_.get(); lg.cursor = log_scal_.get(); lg.flags = log_scal_.get() + 1; lg.start = log_start_.get();
    lg.cap = log_cap_; lg.m_cur = stream_merges_;
    lg.ip_ent = ip_ent_.size() ? ip_ent_.get() : nullptr; lg.ip_start = ip_ent_.size() ? ip_start_.get() : nullptr;
    {
      const char *e = getenv("SWB_IP_LOCAL_MAX");  // (tuning switch)
      lg.ip_local_max = e ? (uint32_t)strtoul(e, nullptr, 10) : CL_IP_LOCAL_MAX;
    }
    return lg;
  }

  // ---------------------------------------------------------------- pair table plumbing
  void ensure_pair_table(uint64_t min_cap) {
    ensure_device();
    uint64_t cap = std::max<uint64_t>(1ull << 20, pow2_ceil(min_cap));
    if (pt_cap_ >= cap) return;
    if (cap > (1ull << 31)) throw Error("pair table would exceed 2^31 slots");
    sync();
    pt_keys_.alloc(cap); pt_val_.alloc(cap); pt_min_.alloc(cap); pt_touched_.alloc(cap);
    pt_scal_.alloc(8);
    SWB_CUDA(cudaMemsetAsync(pt_scal_.get(), 0, pt_scal_.bytes(), stream_));
    recs_.alloc(cap / 2);
    removed_.alloc(1);
    emit_partial_.alloc(2 * 32);
    SWB_CUDA(cudaMemsetAsync(removed_.get(), 0, 8, stream_));
    pt_cap_ = cap;
    pt_ = PairTableDev{pt_keys_.get(), pt_val_.get(), pt_min_.get(), pt_touched_.get(), pt_scal_.get() + 0,
                       pt_scal_.get() + 1, pt_scal_.get() + 2, (uint32_t)(cap - 1), 0, 0, {nullptr, 0}};
    pt_clear<<<sms_ * 8, 256, 0, stream_>>>(pt_); launched();
    if (comm_) {  // second table: the cross-rank reduction target (the local one keeps the deltas until the exchange succeeded)
      pt2_keys_.alloc(cap); pt2_val_.alloc(cap); pt2_min_.alloc(cap); pt2_touched_.alloc(cap); pt2_scal_.alloc(8);
      SWB_CUDA(cudaMemsetAsync(pt2_scal_.get(), 0, pt2_scal_.bytes(), stream_));
      pt2_ = PairTableDev{pt2_keys_.get(), pt2_val_.get(), pt2_min_.get(), pt2_touched_.get(), pt2_scal_.get() + 0,
                          pt2_scal_.get() + 1, pt2_scal_.get() + 2, (uint32_t)(cap - 1), 0, 0, {nullptr, 0}};
      pt_clear<<<sms_ * 8, 256, 0, stream_>>>(pt2_); launched();
    }
  }
  // ---- device frequency table
  bool want_device_tables() const {
    static const bool forced_host = getenv("SWB_HOST_TABLE") && atoi(getenv("SWB_HOST_TABLE")) > 0;
    return (mnranks() == 1 || mcomm() != nullptr) && !forced_host;
  }
  // SWB_TEST_SMALL_GT=1 (tests): the frequency table starts just large enough for the counted pairs, so that it passes
  // 50 % load, stops the resident kernel, grows and is rehashed several times during a small training run.
  static bool gt_small_for_tests() {
    static const bool on = getenv("SWB_TEST_SMALL_GT") && atoi(getenv("SWB_TEST_SMALL_GT")) > 0;
    return on;
  }
  uint64_t gt_presize(uint64_t n_pairs) const {
    if (gt_small_for_tests()) return 2 * n_pairs + 16ull * 260;
    // sized up front so that growing (alloc + rehash, ~10 ms each) is rare: pairs ever touched ~ O(words)
    return std::max<uint64_t>(4ull * n_pairs + 16 * (258 + tr_->config.target_vocab_size), 4 * W);
  }
  void ensure_global_table(uint64_t min_cap) {
    ensure_device();
    const uint64_t cap = std::max<uint64_t>(gt_small_for_tests() ? 1ull << 10 : 1ull << 20, pow2_ceil(min_cap));
    if (gt_cap_ >= cap) return;
    if (cap > (1ull << 31)) throw Error("frequency table would exceed 2^31 slots");
    sync();
    DevBuf<GSlot> k(cap);
    DevBuf<unsigned int> scal(4);
    GlobalTableDev ng{k.get(), scal.get() + 0, scal.get() + 1, (uint32_t)(cap - 1)};
    gt_clear<<<sms_ * 8, 256, 0, stream_>>>(ng); launched();
    if (gt_cap_) { gt_rehash<<<sms_ * 8, 256, 0, stream_>>>(gt_, ng); launched(); }
    sync();
    gt_slots_ = std::move(k); gt_scal_ = std::move(scal);
    gt_ = ng;
    gt_cap_ = cap;
  }
  void reset_global_table() {
    if (!gt_cap_) return;
    gt_clear<<<sms_ * 8, 256, 0, stream_>>>(gt_); launched();
  }
  // One merge inserts at most 4 pairs per distinct neighbour symbol; the kernels raise flag 16 once the
  // table is past 50 % load, so growing then (x4) and keeping capacity >= 16 * symbols means it never fills.
  void maybe_grow_global_table(bool flagged) {
    const uint64_t need = 16ull * (258ull + tr_->num_merges + 2);
    if (flagged) ensure_global_table(std::max<uint64_t>(gt_cap_ * 4, need));
    else if (need > gt_cap_) ensure_global_table(need);
  }
  EmitMode emit_mode(int mode, int32_t da, int32_t db) {
    EmitMode em;
    memset(&em, 0, sizeof em);
    em.mode = device_tables_ ? mode : 0;
    em.min_freq = tr_->config.min_pair_freq;
    em.merged_key = ((unsigned long long)(uint32_t)da << 32) | (uint32_t)db;
    em.stamp_base = op_index_ << 10;
    em.neg_unk_bucket = tr_->config.unk_id < 0 ? (int32_t)((uint32_t)tr_->config.unk_id & 1023u) : -1;
    em.g = gt_;
    em.fused_max = 384;  // measured: beyond this a 32-block pt_emit (one more launch) beats the single-block tail
    return em;
  }
  // caller id -> device code (negative ids: unk_id travels as UNK_CODE, the -1 of a sign-extended key as NEG1_CODE)
  int32_t to_dev(int32_t id) const {
    if (id >= 0) return id;
    if (id == tr_->config.unk_id) return UNK_CODE;
    return NEG1_CODE;  // matches nothing in the stream
  }
  // Leaves device-table mode: the whole device table comes to the host in creation order (stamps), with
  // the versions the host already tracks. Needed for anything but the fresh-count -> merge sequence
  // (a second bpe_count_bigrams without bpe_init, the swb_dist_* / swb_shard_* building blocks).
  void convert_to_host_tables() {
    if (!device_tables_) return;
    device_tables_ = false;
    if (!gt_cap_) return;
    unsigned int used = 0;
    SWB_CUDA(cudaMemcpyAsync(&used, gt_.n_used, 4, cudaMemcpyDeviceToHost, stream_));
    sync();
    std::vector<PairInfo> entries;
    if (used) {
      DevBuf<unsigned long long> d(5ull * used);
      DevBuf<unsigned int> cur(1);
      SWB_CUDA(cudaMemsetAsync(cur.get(), 0, 4, stream_));
      gt_dump<<<sms_ * 8, 256, 0, stream_>>>(gt_, d.get(), cur.get()); launched();
      std::vector<unsigned long long> h(5ull * used);
      SWB_CUDA(cudaMemcpyAsync(h.data(), d.get(), h.size() * 8, cudaMemcpyDeviceToHost, stream_));
      sync();
      std::vector<uint32_t> idx(used);
      for (uint32_t i = 0; i < used; i++) idx[i] = i;
      std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) {
        if (h[5ull * x + 2] != h[5ull * y + 2]) return h[5ull * x + 2] < h[5ull * y + 2];
        return h[5ull * x + 3] < h[5ull * y + 3];
      });
      entries.reserve(used);
      for (uint32_t i : idx) {
        int32_t f = (int32_t)(h[5ull * i] >> 32), s = (int32_t)(h[5ull * i] & 0xFFFFFFFFu);
        if (tr_->config.unk_id < 0) {
          if (f == UNK_CODE) f = tr_->config.unk_id; else if (f == NEG1_CODE) f = -1;
          if (s == UNK_CODE) s = tr_->config.unk_id; else if (s == NEG1_CODE) s = -1;
        }
        entries.push_back(PairInfo{f, s, h[5ull * i + 1], 0, 0});
      }
    }
    core.rebuild_from_dump(entries);
  }

  StreamDev stream_dev() {
    return StreamDev{rows_.get(), sig_.get(), n_rows_, cnt_.get(), long_syms_.get(), long_off_.get(), long_len_.get(),
                     long_word_.get(), n_long_};
  }
  // Waits until the kernels of this merge have published a valid header with sequence `seq` in mapped host
  // memory, and until the n records it announces have all arrived (XOR/SUM check, see pt_emit_range).
  // Polling is cheaper than cudaStreamSynchronize on a loop that runs once per merge; the stream is
  // queried from time to time so that a failed launch cannot hang the host.
  void wait_seq(unsigned long long seq) { wait_seq_at(seq, hdr_.host(), recs_.host(), recs_.size()); }
  void wait_seq_at(unsigned long long seq, volatile unsigned long long *h, const Rec *recs, size_t recs_cap) {
    uint64_t spin = 0;
    double t_wait0 = 0;
    auto check_stream = [&]() {
      if ((++spin & 0x3FFF) != 0) return;
      // A result normally arrives within tens of microseconds: the driver is only asked once the wait is unusually long
      // (every driver call can block behind whatever else holds the driver's locks -- other ranks' NVML clock samplers,
      // for one -- and a blocked host thread stalls the whole merge loop).
      const double t = now_ms();
      if (t_wait0 == 0) { t_wait0 = t; return; }
      if (t - t_wait0 < 5.0) return;
      if (resident_wait_ && t - t_wait0 > 200.0) {
        // The resident kernel reports which command it accepted last (mailbox word 4). If it has moved on to waiting for the
        // command AFTER the merge whose result we are waiting for, that result is lost and neither side will ever move
        // again: fail now (with everything that can help to find out why) instead of when the kernel's watchdog fires.
        const volatile HostCmd2 *stw = hcmd2_.host() + 4;
        const unsigned int sx = stw->x, sy = stw->y, sz = stw->z, sw = stw->w;
        if (sx == (unsigned int)(seq + 1) && ((sy >> 16) & 0xFu) == 0u) {
          volatile unsigned long long *ho = (h == hdr_.host()) ? hdr_b_.host() : hdr_.host();
          char msg[768];
          snprintf(msg, sizeof msg, "lost result: the resident kernel has accepted command %u (mode %u, from a hint %u, GRID merges so far %u; it last published seq ..%u through path %u) "
                   "while the result of merge %llu never arrived (this header: seq %llu/%llu n %llu flags %llx check %s; other header: seq %llu/%llu n %llu flags %llx)",
                   sx, sy & 0xFFu, (sy >> 8) & 0xFFu, sy >> 20, sz, sw, seq, (unsigned long long)h[0], (unsigned long long)h[7], (unsigned long long)h[1],
                   (unsigned long long)h[2], h[6] == hdr_check(h[0], h[1], h[2], h[3], h[4], h[5]) ? "ok" : "BAD", (unsigned long long)ho[0], (unsigned long long)ho[7],
                   (unsigned long long)ho[1], (unsigned long long)ho[2]);
          std::string full(msg);
          {  // stop the kernel (it accepts a stop addressed to the merge before the one it is waiting for) and read its counters
            HostCmd2Sender stopper;
            stopper.c = hcmd2_.host(); stopper.next_seq = seq;
            stopper.send(0, 0, 1);
            unsigned long long ac[16] = {0};
            unsigned int sc[3] = {0, 0, 0};
            if (cudaStreamSynchronize(stream_) == cudaSuccess && cudaMemcpy(ac, cl_acct_.get(), sizeof ac, cudaMemcpyDeviceToHost) == cudaSuccess) {
              if (pt_scal_.size()) cudaMemcpy(sc, pt_scal_.get(), sizeof sc, cudaMemcpyDeviceToHost);
              snprintf(msg, sizeof msg, "; LOCAL %llu GRID %llu merges done in this launch, last GRID tail: seq %llu in block %llu; pair table: touched %u flags %u done_blocks %u",
                       ac[0], ac[2], ac[10], ac[11], sc[0], sc[1], sc[2]);
              full += msg;
            }
          }
          throw Error(full);
        }
      }
      if (resident_wait_ && !launch_returned_.load(std::memory_order_acquire)) return;  // the launch call itself has not returned (a serialising tool): the stream has nothing to say yet
      cudaError_t e = cudaStreamQuery(stream_);
      if (e == cudaSuccess) {
        if (++idle_polls_ > 64) {
          char msg[512];
          const unsigned long long n = h[1], fl = h[2], rem = h[3], cx = h[4], cs = h[5];
          unsigned long long x = 0, sm = 0;
          const volatile long long *r = reinterpret_cast<const volatile long long *>(recs);
          for (size_t i = 0; i < (size_t)std::min<unsigned long long>(n, recs_cap); i++) {
            const unsigned long long a = (unsigned long long)r[4 * i], b = (unsigned long long)r[4 * i + 1], c = (unsigned long long)r[4 * i + 2], d = (unsigned long long)r[4 * i + 3];
            x ^= a ^ b ^ c ^ d; sm += a + 3ull * b + 5ull * c + 7ull * d;
          }
          snprintf(msg, sizeof msg, "merge kernels finished without publishing a valid result (want seq %llu; header seq %llu/%llu n %llu flags %llx removed %llu; "
                   "header check %s; records xor %llx vs %llx, sum %llx vs %llx)", seq, (unsigned long long)h[0], (unsigned long long)h[7], n, fl, rem,
                   h[6] == hdr_check(h[0], n, fl, rem, cx, cs) ? "ok" : "BAD", x, cx, sm, cs);
          std::string full(msg);
          if (cl_acct_.size() >= 16) {  // did the resident kernel's watchdog give up on the host?
            unsigned long long ac[16] = {0};
            if (cudaMemcpy(ac, cl_acct_.get(), sizeof ac, cudaMemcpyDeviceToHost) == cudaSuccess && ac[8]) {
              unsigned int sc[3] = {0, 0, 0};
              if (pt_scal_.size()) cudaMemcpy(sc, pt_scal_.get(), sizeof sc, cudaMemcpyDeviceToHost);
              snprintf(msg, sizeof msg, "; the resident kernel's watchdog gave up waiting for command %llu after %.1f ms (SWB_HOST_TIMEOUT_MS); LOCAL %llu GRID %llu merges done, "
                       "last GRID tail: seq %llu in block %llu; pair table: touched %u flags %u done_blocks %u", ac[8], (double)ac[9] * 1e-6, ac[0], ac[2], ac[10], ac[11], sc[0], sc[1], sc[2]);
              full += msg;
            }
          }
          throw Error(full);
        }
      }
      else if (e != cudaErrorNotReady) SWB_CUDA(e);
    };
    idle_polls_ = 0;
    for (;;) {
      if (h[0] == seq && h[7] == seq) {
        const unsigned long long n = h[1], fl = h[2], rem = h[3], cx = h[4], cs = h[5];
        if (h[6] == hdr_check(seq, n, fl, rem, cx, cs)) {
          if (fl & 8u) return;  // records were not emitted by this kernel
          const size_t m = (size_t)std::min<unsigned long long>(n, recs_cap);
          for (;;) {
            const volatile long long *r = reinterpret_cast<const volatile long long *>(recs);
            unsigned long long x = 0, sm = 0;
            for (size_t i = 0; i < m; i++) {
              const unsigned long long a = (unsigned long long)r[4 * i], b = (unsigned long long)r[4 * i + 1],
                                       c = (unsigned long long)r[4 * i + 2], d = (unsigned long long)r[4 * i + 3];
              x ^= a ^ b ^ c ^ d;
              sm += a + 3ull * b + 5ull * c + 7ull * d;
            }
            if (x == cx && sm == cs) return;
            check_stream();
          }
        }
      }
      check_stream();
    }
  }
  // runs pt_emit, waits, returns the record count (records are in mapped host memory until the next emit)
  size_t emit_and_wait(const EmitMode &em, unsigned int *flags_out, uint64_t *removed_out) {
    return emit_and_wait(pt_, em, flags_out, removed_out);
  }
  size_t emit_and_wait(const PairTableDev &tbl, const EmitMode &em, unsigned int *flags_out, uint64_t *removed_out) {
    const unsigned long long seq = ++seq_;
    pt_emit<<<32, 256, 0, stream_>>>(tbl, em, recs_.dev(), recs_.size(), hdr_.dev(), removed_.get(), seq, emit_partial_.get(),
                                     (tbl.keys == pt_.keys ? pt_scal_.get() : pt2_scal_.get()) + 3);
    launched();
    SWB_CUDA(cudaGetLastError());
    const double tw0 = now_ms();
    stats.host_launch_ms += tw0 - t_launch0_;
    wait_seq(seq);
    stats.host_wait_ms += now_ms() - tw0;
    const size_t n = (size_t)hdr_.host()[1];
    *flags_out = (unsigned int)hdr_.host()[2];
    if (removed_out) *removed_out = hdr_.host()[3];
    return n;
  }
  // device ids -> caller ids (a negative unk_id travels as UNK_CODE on the device)
  void translate_out(Rec *r, size_t n) const {
    if (tr_->config.unk_id >= 0) return;
    for (size_t i = 0; i < n; i++) {
      if (r[i].first == UNK_CODE) r[i].first = tr_->config.unk_id; else if (r[i].first == NEG1_CODE) r[i].first = -1;
      if (r[i].second == UNK_CODE) r[i].second = tr_->config.unk_id; else if (r[i].second == NEG1_CODE) r[i].second = -1;
    }
  }

  // ---------------------------------------------------------------- multi-GPU exchange (NCCL)
  // The communicator is process-wide (creating one costs 0.1-1 s and its first collective as much again):
  // every handle of this process with the same (rank, nranks) shares it. `id` is only read when a new one
  // has to be created, which all ranks do at the same point of the same program.
  struct SharedComm { NcclComm comm = nullptr; int rank = -1, nranks = 0; };
  static SharedComm &shared_comm() { static SharedComm c; return c; }
  static bool have_shared_comm(int rank_, int nranks_) {
    const SharedComm &c = shared_comm();
    return c.comm && c.rank == rank_ && c.nranks == nranks_;
  }
  static void destroy_shared_comm() {
    SharedComm &c = shared_comm();
    if (c.comm) NcclApi::get().CommDestroy(c.comm);
    c = SharedComm();
  }
  void dist_init(int rank_, int nranks_, const NcclUniqueId *id) {
    ensure_device();
    NcclApi &api = NcclApi::get();
    if (!api.ok()) throw Error("NCCL is not available: " + api.error());
    rank = rank_; nranks = nranks_;
    comm_ = nullptr;
    if (nranks > 1) {
      if (!have_shared_comm(rank, nranks)) {
        if (!id) throw Error("swb_dist_init: a unique id is needed to create the communicator");
        destroy_shared_comm();
        SharedComm &c = shared_comm();
        api.check(api.CommInitRank(&c.comm, nranks, *id, rank), "ncclCommInitRank");
        c.rank = rank; c.nranks = nranks;
      }
      comm_ = shared_comm().comm;
    }
    pt_cap_ = 0;  // the pair tables are (re)built with the reduction table next time
  }
  void ensure_dist_buffers(size_t cap) {
    if (cap <= dist_cap_ && d_all_.size()) return;
    sync();
    dist_cap_ = std::max<size_t>(cap, 1024);
    dist_slot_words_ = DIST_HDR_WORDS + 4 * dist_cap_;
    d_all_.alloc(dist_slot_words_ * (size_t)nranks);
  }
  // local table -> all ranks -> reduced into pt2_ -> tail in mode `em`; returns the record count on the host
  size_t exchange_and_reduce(const EmitMode &em, unsigned int *flags_out, uint64_t *removed_out) {
    NcclApi &api = NcclApi::get();
    for (;;) {
      ensure_dist_buffers(dist_cap_ ? dist_cap_ : 8192);
      unsigned long long *self = d_all_.get() + dist_slot_words_ * (size_t)rank;
      dist_copy_out<<<32, 256, 0, stream_>>>(pt_, self, dist_cap_); launched();
      api.check(api.AllGather(self, d_all_.get(), dist_slot_words_ * 8, NcclApi::kChar, comm_, stream_), "ncclAllGather");
      stats.collectives++;
      stats.exchange_bytes += dist_slot_words_ * 8 * (uint64_t)nranks;
      const unsigned long long seq = ++seq_;
      pt2_.canon_on = pt_.canon_on; pt2_.canon_first = pt_.canon_first; pt2_.gpf = pt_.gpf;
      dist_reduce<<<sms_, 256, 0, stream_>>>(d_all_.get(), nranks, dist_slot_words_, pt_, pt2_, em, recs_.dev(), recs_.size(),
                                             hdr_.dev(), removed_.get(), seq);
      launched();
      SWB_CUDA(cudaGetLastError());
      const double tw0 = now_ms();
      stats.host_launch_ms += tw0 - t_launch0_;
      wait_seq(seq);
      stats.host_wait_ms += now_ms() - tw0;
      unsigned int flags = (unsigned int)hdr_.host()[2];
      if (flags & 4u) {  // some rank had more records than a slot holds (or a full table): every rank sees the same header
        if (flags & 1u) { *flags_out = flags; return 0; }
        t_launch0_ = now_ms();
        ensure_dist_buffers(pow2_ceil(2 * (size_t)hdr_.host()[1]));
        continue;
      }
      size_t n = (size_t)hdr_.host()[1];
      if (removed_out) *removed_out = hdr_.host()[3];
      if (flags & 8u) {
        t_launch0_ = now_ms();
        n = emit_and_wait(pt2_, em, &flags, removed_out);
      }
      *flags_out = flags;
      return n;
    }
  }

  // ---------------------------------------------------------------- count (this rank's words)
  // Launches the count kernels and the emit in the given mode; returns the records.
  //   mode 0: (pair, weighted frequency, first-touch key) for every distinct pair of this rank's share
  //   mode 2: the counts go to the device frequency table; records only for pairs >= min_pair_freq
  const Rec *run_count(int mode, size_t *n_out) {
    ensure_pair_table(1);
    const double t0 = now_ms();
    if (!loaded_) { *n_out = 0; return recs_.host(); }
    for (;;) {
      StreamDev s = stream_dev();
      t_launch0_ = now_ms();
      if (n_rows_) {
        const int grid = (int)std::min<uint64_t>((uint64_t)sms_ * 8, (n_rows_ + 7) / 8);
        count_rows<<<grid, MERGE_THREADS, 0, stream_>>>(s, pt_, unk_dev()); launched();
      }
      if (n_long_) { count_long<<<std::min<uint32_t>(sms_ * 4, (n_long_ + 3) / 4), 128, 0, stream_>>>(s, pt_, unk_dev()); launched(); }
      unsigned int flags = 0;
      size_t n = 0;
      if (mcomm()) {
        if (mode == 2) {  // sized for the union of all ranks' pairs: at most nranks x the largest local list
          unsigned int n_pairs = 0;
          SWB_CUDA(cudaMemcpyAsync(&n_pairs, pt_.n_touched, 4, cudaMemcpyDeviceToHost, stream_));
          sync();
          ensure_global_table(gt_presize((uint64_t)n_pairs * nranks));
        }
        EmitMode em = emit_mode(mode, 0, 0);
        n = exchange_and_reduce(em, &flags, nullptr);
        if (flags & 1u) {  // a rank's pair table was too small: every rank grows alike and recounts
          sync();
          const uint64_t want = pt_cap_ * 4;
          pt_cap_ = 0;
          ensure_pair_table(want);
          continue;
        }
      } else {
        unsigned int overflow = 0;  // read the flag BEFORE emitting: in mode 2 the emit writes into the frequency table
        SWB_CUDA(cudaMemcpyAsync(&overflow, pt_.flags, 4, cudaMemcpyDeviceToHost, stream_));
        sync();
        if (overflow & 1u) {  // more distinct pairs than the table holds: drain it, grow, recount
          emit_and_wait(emit_mode(0, 0, 0), &flags, nullptr);
          const uint64_t want = pt_cap_ * 4;
          pt_cap_ = 0;
          ensure_pair_table(want);
          continue;
        }
        if (mode == 2) {
          unsigned int n_pairs = 0;
          SWB_CUDA(cudaMemcpyAsync(&n_pairs, pt_.n_touched, 4, cudaMemcpyDeviceToHost, stream_));
          sync();
          ensure_global_table(gt_presize(n_pairs));
        }
        EmitMode em = emit_mode(mode, 0, 0);
        n = emit_and_wait(em, &flags, nullptr);
      }
      if (flags & 4u) throw Error("record buffer too small for the count pass (internal sizing error)");
      if (flags & 16u) gt_flagged_ = true;
      translate_out(recs_.host(), n);
      *n_out = n;
      stats.count_ms += now_ms() - t0;
      return recs_.host();
    }
  }
  const Rec *shard_count(size_t *n_out) {  // building block of the multi-rank loop: always plain records
    convert_to_host_tables();
    tables_fresh_ = false;
    return run_count(0, n_out);
  }

  // ---------------------------------------------------------------- merge (this rank's words)
  const Rec *run_merge(int32_t a, int32_t b, int32_t new_id, size_t *n_out) {
    // at most 4 distinct pairs per distinct neighbour symbol: size the table so that it cannot fill up
    ensure_pair_table(8ull * (258ull + tr_->num_merges + 2));
    *n_out = 0;
    if (!loaded_) return recs_.host();
    const int32_t unk = tr_->config.unk_id;
    const bool absent = (a < 0 && a != unk) || (b < 0 && b != unk);  // such ids exist in no word
    if (absent && !device_tables_) { log_ok_ = false; return recs_.host(); }  // (the token numbering moves on without this stream)
    const int32_t da = to_dev(a), db = to_dev(b);
    if (device_tables_) { maybe_grow_global_table(gt_flagged_); gt_flagged_ = false; }
    op_index_++;
    EmitMode em = emit_mode(1, da, db);
    em.log = birth_log(new_id, 1);
    pt_.canon_on = (device_tables_ && unk < 0) ? 1 : 0;
    pt_.canon_first = unk == -1 ? UNK_CODE : NEG1_CODE;
    pt_.gpf.slots = device_tables_ ? (void *)gt_.slots : nullptr;
    pt_.gpf.mask = gt_.mask;
    StreamDev s = stream_dev();
    t_launch0_ = now_ms();
    if (timing) SWB_CUDA(cudaEventRecord(ev0_, stream_));
    const bool fused = n_rows_ && !n_long_ && !mcomm();
    unsigned long long seq = 0;
    if (n_rows_) {
      const uint64_t warps_needed = (n_rows_ + 31) / 32;  // one warp tests 32 row signatures per iteration
      const int grid = (int)std::min<uint64_t>((uint64_t)sms_ * 4, (warps_needed + MERGE_WARPS - 1) / MERGE_WARPS);
      if (fused) seq = ++seq_;
      merge_rows<<<grid, MERGE_THREADS, 0, stream_>>>(s, pt_, da, db, new_id, removed_.get(), fused ? 1 : 0, em, recs_.dev(),
                                                      recs_.size(), hdr_.dev(), seq);
      launched();
      stats.merge_launches++;
    }
    if (timing) SWB_CUDA(cudaEventRecord(ev1_, stream_));
    if (n_long_) { merge_long<<<std::min<uint32_t>(sms_ * 4, (n_long_ + 3) / 4), 128, 0, stream_>>>(s, pt_, da, db, new_id, removed_.get()); launched(); }
    unsigned int flags = 0;
    uint64_t removed = 0;
    size_t n = 0;
    if (fused) {
      SWB_CUDA(cudaGetLastError());
      const double tw0 = now_ms();
      stats.host_launch_ms += tw0 - t_launch0_;
      wait_seq(seq);
      const double tw1 = now_ms();
      stats.host_wait_ms += tw1 - tw0;
      if (trace_wait_) {
        wait_trace_.push_back((float)((tw1 - t_launch0_) * 1e3));
#ifdef SWB_KERNEL_TRACE
        const volatile unsigned long long *hh = hdr_.host();
        trace_scan_.push_back((float)((double)(hh[8] - hh[11]) * 1e-3));
        trace_tail_.push_back((float)((double)(hh[9] - hh[8]) * 1e-3));
        trace_nrec_.push_back((uint32_t)hh[10]);
        trace_cand_.push_back((uint32_t)hh[20]); trace_logn_.push_back((uint64_t)hh[21]); trace_mrows_.push_back((uint32_t)hh[22]);
        if (trace_removed_.size() > 200 && hh[10] < 10) { tw_sig_ += (double)(hh[12] - hh[11]) * 1e-3; tw_rows_ += (double)(hh[13] - hh[12]) * 1e-3; tw_fence_ += (double)(hh[14] - hh[13]) * 1e-3; tw_cand_ += (double)hh[15]; tw_n_++; }
#endif
      }
      n = (size_t)hdr_.host()[1];
      flags = (unsigned int)hdr_.host()[2];
      removed = hdr_.host()[3];
      if (flags & 8u) {  // too many records for the fused tail: emit them with a full grid
        t_launch0_ = now_ms();
        n = emit_and_wait(em, &flags, &removed);
      }
    } else if (mcomm()) {
      n = exchange_and_reduce(em, &flags, &removed);
    } else {
      n = emit_and_wait(em, &flags, &removed);
    }
    if (flags & 16u) gt_flagged_ = true;
    if (flags & 5u) throw Error("pair table overflow during a merge (internal sizing error)");
    if (flags & 32u) throw Error("birth log overflow (internal sizing error)");
    merge_done_on_stream(hdr_.host()[2]);
    if (timing) {
      float ms = 0;
      SWB_CUDA(cudaEventSynchronize(ev1_));
      SWB_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
      stats.merge_kernel_ms += ms;
    }
    stats.merge_scan_bytes += n_rows_ * ROW * 4;
    stats.merge_alg_bytes += 4 * live_symbols_ + 8 * (mnranks() == 1 ? W : (W + mnranks() - 1 - mrank()) / mnranks());
    live_symbols_ -= removed;
    stats.live_symbols = live_symbols_;
    if (trace_wait_ && !wait_trace_.empty()) trace_removed_.push_back((uint32_t)removed);
    translate_out(recs_.host(), n);
    *n_out = n;
    return recs_.host();
  }
  const Rec *shard_merge(int32_t a, int32_t b, int32_t new_id, size_t *n_out) {  // multi-rank building block
    convert_to_host_tables();
    tables_fresh_ = false;
    return run_merge(a, b, new_id, n_out);
  }

  // ---------------------------------------------------------------- single-process drivers
  void count_bigrams() {  // reference bpe_count_bigrams
    check_not_failed();
    size_t n = 0;
    if (tables_fresh_ && want_device_tables() && loaded_) {
      // fresh pair table (right after a load or a reset): the device keeps the frequencies from here on
      device_tables_ = true;
      ensure_device();
      ensure_global_table(1);
      tables_fresh_ = false;
      op_index_++;
      const Rec *r = run_count(2, &n);
      core.seed_absolute(r, n);
      return;
    }
    // counting on top of a used table (the reference then ADDS to the existing frequencies and pushes every
    // entry again): replay on the host-resident table, which needs every pair -> leave device mode
    convert_to_host_tables();
    tables_fresh_ = false;
    const Rec *r = run_count(0, &n);
    core.seed_counts(r, n);
  }
  void reset_tables() {  // reference bpe.cpp:177-183
    core.reset_tables();
    if (stream_) reset_global_table();
    device_tables_ = false;
    tables_fresh_ = true;
    gt_used_estimate_ = 0;
  }
  void mark_used() { tables_fresh_ = false; }
  void init() {  // reference bpe_init
    reset_tables();
    count_bigrams();
  }
  // ---------------------------------------------------------------- when the resident kernel can serve the merges
  // (SWB_NO_PERSISTENT=1: one merge_rows launch per merge instead -- the path the sharded multi-GPU loop, long words
  // and kernel timing use anyway; tests/test_gpu_variants.py runs the parity suite over it)
  bool use_persistent() const {
    static const bool off = getenv("SWB_NO_PERSISTENT") && atoi(getenv("SWB_NO_PERSISTENT")) > 0;
    return !off && device_tables_ && !mcomm() && !timing && loaded_ && n_rows_ > 0 && n_long_ == 0 && coop_ok_;
  }
  // ---------------------------------------------------------------- resident cluster kernel (default single-GPU path)
  struct ClusterFacts { int clusters = 0; bool ok = false, cooperative = false; };
  static const ClusterFacts &cluster_facts(int device) {
    static std::mutex mu;
    static std::map<int, ClusterFacts> facts;
    std::lock_guard<std::mutex> g(mu);
    auto it = facts.find(device);
    if (it != facts.end()) return it->second;
    ClusterFacts f;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (cudaFuncSetAttribute(merge_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(CL_SIZE * 64); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = CL_SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CL_SIZE; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, merge_cluster, &cfg) == cudaSuccess && n >= 1) { f.clusters = n; f.ok = true; f.cooperative = coop != 0; }
    }
    cudaGetLastError();
    if (getenv("SWB_TRACE_INIT")) fprintf(stderr, "[trace] cluster kernel: %d resident clusters of %d CTAs, %zu bytes of shared memory each\n", f.clusters, CL_SIZE, CL_SMEM_BYTES);
    return facts.emplace(device, f).first->second;
  }
  // Device-side watchdog of the resident kernel: how long its leader waits for the next command before it gives up
  // (a host thread that died without telling it to stop). It is NOT a latency bound: a host thread that is descheduled
  // for seconds (eight busy ranks on one box, a debugger, paging) must not lose the kernel, so the default is generous;
  // an exception on the host side stops the kernel explicitly (HostCmd2Sender's destructor).
  static unsigned long long host_timeout_ns() {
    static const unsigned long long v = [] {
      const char *e = getenv("SWB_HOST_TIMEOUT_MS");
      const unsigned long long ms = e ? strtoull(e, nullptr, 10) : 0;
      return (ms ? ms : 60000ull) * 1000000ull;
    }();
    return v;
  }
  bool use_resident() const {
    return use_persistent() && cluster_facts(device_).ok;
  }
  struct HostCmd2Sender {  // makes sure the resident kernel is always told to stop, also when an exception unwinds
    volatile HostCmd2 *c = nullptr;
    unsigned long long next_seq = 0;
    bool running = false;
    // One mailbox word = 16 bytes {x, y, z, check(seq, x, y, z)}. On x86-64 it is written with ONE aligned 16-byte SSE store
    // (single-copy atomic on every CPU this has run on, not architecturally promised); elsewhere as two 8-byte stores, the
    // half that carries the 32-bit check word last, behind a release fence. Either way the device accepts a word only when
    // the check matches the sequence number it expects (cluster_kernel.cuh), so a torn read is read again, never acted on.
    static void store16(volatile HostCmd2 *dst, unsigned int x, unsigned int y, unsigned int z, unsigned int check) {
#if defined(__x86_64__) || defined(_M_X64)
      const __m128i v = _mm_set_epi32((int)check, (int)z, (int)y, (int)x);
      _mm_store_si128(reinterpret_cast<__m128i *>(const_cast<HostCmd2 *>(dst)), v);
#else
      volatile unsigned long long *w = reinterpret_cast<volatile unsigned long long *>(dst);
      __atomic_store_n(const_cast<unsigned long long *>(w), (unsigned long long)x | ((unsigned long long)y << 32), __ATOMIC_RELAXED);
      __atomic_thread_fence(__ATOMIC_RELEASE);
      __atomic_store_n(const_cast<unsigned long long *>(w + 1), (unsigned long long)z | ((unsigned long long)check << 32), __ATOMIC_RELEASE);
#endif
      __atomic_thread_fence(__ATOMIC_SEQ_CST);
    }
    void send(unsigned long long pair, unsigned int new_id, unsigned int op) {
      const unsigned int x = (unsigned int)pair, y = (unsigned int)(pair >> 32), z = (new_id & 0x0FFFFFFFu) | (op << 28);
      store16(c, x, y, z, cmd3_word(next_seq, x, y, z));
    }
    // hint for the merge with sequence number `seq`: mailbox word 1 + (seq & 1) (two hint words used in turn, so that the
    // hint for merge q+2 can be written while the device may still be reading the one for merge q+1)
    void hint(unsigned long long seq, unsigned long long pair, unsigned int freq) {
      const unsigned int x = (unsigned int)pair, y = (unsigned int)(pair >> 32), z = freq;
      store16(c + 1 + (seq & 1ull), x, y, z, cmd3_word(seq, x, y, z));
    }
    ~HostCmd2Sender() { if (running) send(0, 0, 1); }
  };
  // Debug builds (-DSWB_DEBUG_BOUNDS): index violations the kernels recorded instead of following them
  void check_debug_bounds() {
#ifdef SWB_DEBUG_BOUNDS
    unsigned long long d[1 + 4 * 16];
    SWB_CUDA(cudaMemcpyFromSymbol(d, g_dbg, sizeof d));
    if (!d[0]) return;
    std::string msg = "debug bounds: " + std::to_string(d[0]) + " violation(s):";
    for (unsigned long long i = 0; i < d[0] && i < 16; i++) {
      char b[160];
      snprintf(b, sizeof b, " [code %llu: %llu %llu %llu]", d[1 + 4 * i], d[2 + 4 * i], d[3 + 4 * i], d[4 + 4 * i]);
      msg += b;
    }
    throw Error(msg);
#endif
  }
  // Limits of the resident kernel's fast paths; the SWB_TEST_* variables only ever SHRINK them (tests/test_gpu_variants.py
  // drives the overflow paths on small corpora with them).
  static ClusterTune cluster_tune() {
    static const ClusterTune t = [] {
      ClusterTune c = cluster_tune_default();
      auto shrink = [](const char *name, unsigned int &v) {
        const char *e = getenv(name);
        if (e && *e) { const unsigned long x = strtoul(e, nullptr, 10); if (x < v) v = (unsigned int)x; }
      };
      shrink("SWB_TEST_LOCAL_MAX", c.local_max); shrink("SWB_TEST_CAND_CAP", c.cand_cap); shrink("SWB_TEST_INBOX", c.inbox_cap);
      shrink("SWB_TEST_REC_STAGE", c.rec_stage); shrink("SWB_TEST_BIRTH_STAGE", c.birth_stage); shrink("SWB_TEST_MAX_PROBES", c.max_probes); shrink("SWB_TEST_SOLO_MAX", c.solo_max);
      if (c.max_probes < 1) c.max_probes = 1;
      return c;
    }();
    return t;
  }
  // One launch of the resident kernel: clusters of CL_SIZE CTAs, co-resident (cooperative attribute where available: GRID
  // merges spin on a grid-wide counter), the symbol rows pinned in L2 for the duration of the launch.
  // Nsight Compute refuses a launch that carries both a cluster dimension and the cooperative attribute ("LaunchFailed", and
  // the process is gone). The attribute is a launch-time check only (this kernel never calls grid.sync(); its grid is sized
  // from the occupancy query), so it is dropped when the profiler's injection library is in the process.
  static bool profiler_attached() {
    static const bool attached = [] {
      if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) return true;
      FILE *f = fopen("/proc/self/maps", "r");
      if (!f) return false;
      char line[1024];
      bool hit = false;
      while (!hit && fgets(line, sizeof line, f))
        hit = strstr(line, "nsight-compute") || strstr(line, "libcuda-injection") || strstr(line, "libInterceptorInjectionTarget");
      fclose(f);
      return hit;
    }();
    return attached;
  }
  void launch_merge_cluster(const StreamDev &s, const EmitMode &em, unsigned long long *removed_p, Rec *out0, Rec *out1, size_t out_cap,
                            unsigned long long *out_hdr0, unsigned long long *out_hdr1, unsigned long long seq_base, unsigned long long op_base,
                            volatile HostCmd2 *hc, DevCmd2 *dc, unsigned long long timeout_ns, unsigned long long *trace_p,
                            const HostCmd2 *script, unsigned long long script_n) {
    const ClusterFacts &cf = cluster_facts(device_);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    // (profiling switches: SWB_CLUSTERS=n caps the number of clusters, SWB_NO_COOP=1 drops the cooperative attribute -- a profiler
    //  that takes SMs for itself can make the full cooperative grid "too large"; fewer clusters only slow GRID merges down)
    static const int cap_clusters = getenv("SWB_CLUSTERS") ? atoi(getenv("SWB_CLUSTERS")) : 0;
    static const bool no_coop = (getenv("SWB_NO_COOP") && atoi(getenv("SWB_NO_COOP")) > 0) || profiler_attached();
    const int n_clusters = cap_clusters > 0 ? std::min(cap_clusters, cf.clusters) : cf.clusters;
    cfg.gridDim = dim3((unsigned)(n_clusters * CL_SIZE)); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = CL_SMEM_BYTES; cfg.stream = stream_;
    cudaLaunchAttribute at[3];
    int n_at = 0;
    at[n_at].id = cudaLaunchAttributeClusterDimension;
    at[n_at].val.clusterDim.x = CL_SIZE; at[n_at].val.clusterDim.y = 1; at[n_at].val.clusterDim.z = 1;
    n_at++;
    if (cf.cooperative && !no_coop) { at[n_at].id = cudaLaunchAttributeCooperative; at[n_at].val.cooperative = 1; n_at++; }
    {  // keep the symbol rows resident in L2 across the merges of this launch (the birth log streams through it)
      static const bool no_persist = getenv("SWB_NO_L2_PERSIST") && atoi(getenv("SWB_NO_L2_PERSIST")) > 0;
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device_);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device_);
      const size_t want = n_rows_ * (size_t)ROW * 4;
      if (!no_persist && max_persist > 0 && max_window > 0 && want > 0) {
        const size_t persist = std::min<size_t>((size_t)max_persist, want);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
          const size_t win = std::min<size_t>(want, (size_t)max_window);
          at[n_at].id = cudaLaunchAttributeAccessPolicyWindow;
          at[n_at].val.accessPolicyWindow.base_ptr = (void *)rows_.get();
          at[n_at].val.accessPolicyWindow.num_bytes = win;
          at[n_at].val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)win);
          at[n_at].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          at[n_at].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          n_at++;
        }
        cudaGetLastError();
      }
    }
    cfg.attrs = at; cfg.numAttrs = n_at;
    SWB_CUDA(cudaLaunchKernelEx(&cfg, merge_cluster, s, pt_, em, removed_p, out0, out1, out_cap, out_hdr0, out_hdr1, seq_base, op_base, hc, dc, timeout_ns, trace_p, cl_ovf_.get(), script, script_n, cluster_tune(), cl_acct_.get()));
    SWB_CUDA(cudaGetLastError());
  }
  // Profiling aid (swb_profile_scripted_merges): runs the first n merges of `triples` -- a merge list this same corpus
  // produced before -- inside ONE launch of merge_cluster whose commands come from a script in device memory instead of the
  // mailbox. The kernel then never waits for the host, so (a) ncu's kernel replay can capture it (the product launch cannot be
  // replayed: its commands depend on a host that has moved on) and (b) its duration is the device-side floor of the merge loop.
  // Must follow init()/count_bigrams() on a freshly loaded corpus. The host's heap replica does not follow: the handle is
  // consumed (marked failed) afterwards. Returns the kernel's duration in ms (CUDA events on the launch stream).
  double profile_scripted(const int32_t *triples, size_t n) {
    check_not_failed();
    if (!use_resident() || tr_->num_merges != 0) throw Error("profile_scripted needs a freshly counted single-GPU corpus without long words");
    if (tr_->config.unk_id < 0) throw Error("profile_scripted: negative unk_id is not supported");
    ensure_pair_table(8ull * (258ull + n + 2));
    ensure_global_table(std::max<uint64_t>(gt_cap_ * 4, 16ull * (258ull + n + 2)));  // (the script cannot stop to let the table grow)
    if (!hcmd2_.size()) { hcmd2_.alloc(8); dcmd2_.alloc(1); cl_ovf_.alloc(CL_LOCAL_MAX); cl_acct_.alloc(32); }
    if (recs_b_.size() != recs_.size()) recs_b_.alloc(recs_.size());
    if (!hdr_b_.size()) hdr_b_.alloc(HDR_WORDS);
    SWB_CUDA(cudaMemsetAsync(cl_acct_.get(), 0, cl_acct_.bytes(), stream_));
    memset((void *)hcmd2_.host(), 0, 8 * sizeof(HostCmd2));
    SWB_CUDA(cudaMemsetAsync(dcmd2_.get(), 0, sizeof(DevCmd2), stream_));
    EmitMode em = emit_mode(1, 0, 0);
    em.log = birth_log(256, (uint32_t)n + 1);
    em.fused_max = 0xFFFFFFFFu;
    pt_.canon_on = 0; pt_.canon_first = NEG1_CODE;
    pt_.gpf.slots = (void *)gt_.slots; pt_.gpf.mask = gt_.mask;
    const unsigned long long seq_base = seq_, op_base = op_index_ + 1;
    std::vector<HostCmd2> h(n);
    for (size_t m = 0; m < n; m++) {
      HostCmd2 c;
      c.x = (unsigned int)triples[3 * m + 1]; c.y = (unsigned int)triples[3 * m]; c.z = (unsigned int)triples[3 * m + 2] & 0x0FFFFFFFu;
      c.w = cmd3_word(seq_base + m + 1, c.x, c.y, c.z);
      h[m] = c;
    }
    DevBuf<HostCmd2> script(n ? n : 1);
    if (n) SWB_CUDA(cudaMemcpyAsync(script.get(), h.data(), n * sizeof(HostCmd2), cudaMemcpyHostToDevice, stream_));
    failed_ = "consumed by swb_profile_scripted_merges (the heap replica did not follow the device)";
    SWB_CUDA(cudaEventRecord(ev0_, stream_));
    launch_merge_cluster(stream_dev(), em, removed_.get(), recs_.dev(), recs_b_.dev(), recs_.size(), hdr_.dev(), hdr_b_.dev(), seq_base, op_base,
                         hcmd2_.dev(), dcmd2_.get(), host_timeout_ns(), nullptr, script.get(), n);
    launched(); stats.merge_launches++;
    SWB_CUDA(cudaEventRecord(ev1_, stream_));
    sync();
    float ms = 0;
    SWB_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    unsigned long long ac[16];
    SWB_CUDA(cudaMemcpy(ac, cl_acct_.get(), sizeof ac, cudaMemcpyDeviceToHost));
    stats.resident_local_merges += ac[0]; stats.resident_local_ms += (double)ac[1] * 1e-6;
    stats.resident_grid_merges += ac[2]; stats.resident_grid_ms += (double)ac[3] * 1e-6;
    stats.merge_kernel_ms += (double)(ac[1] + ac[3]) * 1e-6;
    if (ac[0] + ac[2] != n) throw Error("profile_scripted: the kernel performed " + std::to_string(ac[0] + ac[2]) + " of " + std::to_string(n) + " merges");
    return (double)ms;
  }
  // Runs up to max_merges merges starting with (a, b) (already popped, pending in `core`) inside ONE launch of
  // merge_cluster. Returns the number performed; stops early when the heap is exhausted or a table has to grow.
  int run_resident(int32_t a, int32_t b, int32_t new_id, int max_merges) {
    ensure_pair_table(8ull * (258ull + tr_->num_merges + (uint64_t)max_merges + 2));
    maybe_grow_global_table(gt_flagged_);
    gt_flagged_ = false;
    if (!hcmd2_.size()) { hcmd2_.alloc(8); dcmd2_.alloc(1); cl_ovf_.alloc(CL_LOCAL_MAX); cl_acct_.alloc(32); }
    if (recs_b_.size() != recs_.size()) recs_b_.alloc(recs_.size());  // second record buffer + header: merges started from a hint
    if (!hdr_b_.size()) hdr_b_.alloc(HDR_WORDS);
    memset(hdr_b_.host(), 0, HDR_WORDS * sizeof(unsigned long long));
    SWB_CUDA(cudaMemsetAsync(cl_acct_.get(), 0, cl_acct_.bytes(), stream_));
    memset((void *)hcmd2_.host(), 0, 8 * sizeof(HostCmd2));
    SWB_CUDA(cudaMemsetAsync(dcmd2_.get(), 0, sizeof(DevCmd2), stream_));
    const int32_t unk = tr_->config.unk_id;
    EmitMode em = emit_mode(1, 0, 0);
    em.log = birth_log(new_id, (uint32_t)max_merges + 1);
    em.fused_max = 0xFFFFFFFFu;  // the resident kernel's tails take every merge, whatever its size
    pt_.canon_on = unk < 0 ? 1 : 0;
    pt_.canon_first = unk == -1 ? UNK_CODE : NEG1_CODE;
    pt_.gpf.slots = (void *)gt_.slots; pt_.gpf.mask = gt_.mask;
    StreamDev s = stream_dev();
    const unsigned long long seq_base = seq_, op_base = op_index_ + 1;
    unsigned long long *removed_p = removed_.get();
    // results of the merge with sequence number q travel through buffer q & 1
    Rec *out0 = recs_.dev(), *out1 = recs_b_.dev();
    size_t out_cap = recs_.size();
    unsigned long long *out_hdr0 = hdr_.dev(), *out_hdr1 = hdr_b_.dev();
    volatile HostCmd2 *hc = hcmd2_.dev();
    DevCmd2 *dc = dcmd2_.get();
    const unsigned long long timeout_ns = host_timeout_ns();
    static const bool no_hints = getenv("SWB_NO_HINTS") && atoi(getenv("SWB_NO_HINTS")) > 0;
    const bool hints = !no_hints && unk >= 0;  // (a negative unk_id canonicalises pairs on the device: no look-ahead there)
    constexpr unsigned long long NO_HINT = ~0ull;
    // Look-ahead (HostCore::peek_next): while merge q is pending, the exact heap names the pairs of merges q+1 and q+2,
    // each under the condition that the merges before it push nothing at or above its frequency F and leave its own
    // frequency at F. The device checks that against its frequency table when the merge before has finished and then
    // starts without waiting for the command (which still follows, and which the host checks against: flag 64).
    //   late hint  for q+1: sent right after the look-ahead (the device checks merge q)
    //   early hint for q+2: sent as soon as the records of merge q have arrived and show that q pushed nothing >= F
    //                       (the device checks merge q+1) -- it is on its way a whole merge before it is needed
    // The hints need two look-ahead entries. SWB_PEEK_DEPTH=n (3..8) asks for a longer list and reuses its tail for the following
    // merges instead of walking the heap again (entry i names merge q+1+i under conditions that do not depend on WHEN the host
    // looks, so one merge later the list minus its first entry is still valid if the merge applied in between pushed nothing
    // at or above the entries kept and left their frequencies alone). Exact (tests/test_dist_host_logic.py checks the claims
    // to depth 6) and it takes 46 ms of heap walking off a config-3 step -- but the step got 25-35 ms SLOWER in an alternating
    // A/B on one box (fewer hints accepted, the device waits longer for the host), so the default stays at 2: a fresh walk per merge.
    constexpr size_t PEEK_MAX = 8;
    static const size_t peek_depth = [] {
      const char *e = getenv("SWB_PEEK_DEPTH");
      const size_t v = e ? (size_t)strtoul(e, nullptr, 10) : 2;
      return std::min<size_t>(std::max<size_t>(v, 2), (size_t)8);
    }();
    static_assert(PEEK_MAX == 8, "SWB_PEEK_DEPTH is clamped to the list's capacity");
    HostCore::Peek pk[PEEK_MAX];
    size_t npk = 0;
    uint64_t last_mp = ~0ull;  // largest frequency the merge applied last made the heap push; ~0 = unknown
    unsigned long long n_early = 0;
    static const bool no_early = getenv("SWB_NO_EARLY_HINTS") && atoi(getenv("SWB_NO_EARLY_HINTS")) > 0;
    unsigned long long hint_ring[4][2];  // what was sent for sequence number q: [q & 3] = {early, late}
    for (auto &h2 : hint_ring) h2[0] = h2[1] = NO_HINT;
    auto send_hint = [&](unsigned long long hint_seq, const HostCore::Peek &e, int which, HostCmd2Sender &sd) {
      if (e.freq == 0 || e.freq >= (1ull << 32)) return;
      const unsigned long long key = ((unsigned long long)(uint32_t)to_dev(e.a) << 32) | (uint32_t)to_dev(e.b);
      if (which == 1 && hint_ring[hint_seq & 3][0] == key) return;  // the early hint already said so
      sd.hint(hint_seq, key, (unsigned int)e.freq);
      hint_ring[hint_seq & 3][which] = key;
      stats.hints_sent++;
    };
    auto look_ahead = [&](int32_t ca, int32_t cb) {  // (ca, cb) = the merge that has just been chosen (host ids)
      if (!hints) { npk = 0; return; }
      const double th0 = now_ms();
      while (npk > 0 && pk[npk - 1].freq <= last_mp) npk--;  // entries the merge applied last may have overtaken (or: nothing known about it)
      bool reuse = npk >= 3 && pk[0].a == ca && pk[0].b == cb;
      for (size_t i = 1; reuse && i < npk; i++) reuse = core.peek_still_valid(pk[i]);  // (a pair the last merge touched: its entry is stale now)
      if (reuse) {
        for (size_t i = 1; i < npk; i++) pk[i - 1] = pk[i];
        npk--;
      } else {
        npk = core.peek_next(pk, peek_depth);
      }
      stats.host_peek_ms += now_ms() - th0;
    };
    unsigned long long *trace_p = nullptr;
    if (trace_wait_) {
      if (!ptrace_.size()) { ptrace_.alloc(32); SWB_CUDA(cudaMemsetAsync(ptrace_.get(), 0, ptrace_.bytes(), stream_)); }
      trace_p = ptrace_.get();
    }
    // The launch call is made from a helper thread. Normally it returns within microseconds and the thread is gone before the
    // first result arrives. Under a tool that serialises kernels (ncu's launch list, compute-sanitizer) cudaLaunchKernel does
    // not return until the kernel has finished -- and this kernel only finishes once THIS thread has served its mailbox: a
    // launch from this thread would wait for itself until the watchdog fires. (Declared before `sender`: on an exception the
    // sender's destructor tells the kernel to stop first, then this one joins.)
    struct Launcher {
      std::thread th;
      std::exception_ptr err;
      std::atomic<int> *returned = nullptr;
      ~Launcher() { if (th.joinable()) th.join(); if (returned) returned->store(1); }
    } launcher;
    launcher.returned = &launch_returned_;
    HostCmd2Sender sender;
    sender.c = hcmd2_.host();
    sender.next_seq = seq_base + 1;  // the first merge travels through the mailbox like all the others
    int32_t da = to_dev(a), db = to_dev(b);
    sender.send(((unsigned long long)(uint32_t)da << 32) | (uint32_t)db, (unsigned int)new_id, 0);
    launch_returned_.store(0);
    sender.running = true;  // (from here on the kernel may be running: a stop must reach it whatever happens)
    launcher.th = std::thread([&]() {
      try {
        SWB_CUDA(cudaSetDevice(device_));
        launch_merge_cluster(s, em, removed_p, out0, out1, out_cap, out_hdr0, out_hdr1, seq_base, op_base, hc, dc, timeout_ns, trace_p, nullptr, 0);
      } catch (...) { launcher.err = std::current_exception(); }
      launch_returned_.store(1, std::memory_order_release);
    });
    {  // the normal case: wait the few microseconds the launch takes, so that a launch error surfaces here and not as a time-out
      const double tl0 = now_ms();
      while (!launch_returned_.load(std::memory_order_acquire) && now_ms() - tl0 < 20.0) std::this_thread::yield();
      if (launch_returned_.load(std::memory_order_acquire)) {
        launcher.th.join();
        if (launcher.err) { sender.running = false; std::rethrow_exception(launcher.err); }
      }
    }
    launched(); stats.merge_launches++;
    int done = 0;
    unsigned long long cur_key = ((unsigned long long)(uint32_t)da << 32) | (uint32_t)db;
    const uint64_t minf = tr_->config.min_pair_freq;
    look_ahead(a, b);
    if (npk >= 1 && max_merges > 1) send_hint(seq_base + 2, pk[0], 1, sender);
    for (;;) {
      const double tw0 = now_ms();
      const unsigned long long q = seq_base + done + 1;
      unsigned long long *hh = (q & 1ull) ? hdr_b_.host() : hdr_.host();
      Rec *hr = (q & 1ull) ? recs_b_.host() : recs_.host();
      resident_wait_ = true;
      wait_seq_at(q, hh, hr, recs_.size());
      resident_wait_ = false;
      const double tw1 = now_ms();
      stats.host_wait_ms += tw1 - tw0;
      const size_t n = (size_t)hh[1];
      const unsigned long long hflags = hh[2];
      const unsigned int flags = (unsigned int)hflags;
      const uint64_t removed = hh[3];
      if (flags & 64u) {  // the device started this merge from a hint: it must be the pair the exact heap chose
        if (hint_ring[q & 3][0] != cur_key && hint_ring[q & 3][1] != cur_key)
          throw Error("resident kernel followed a hint the heap replica did not confirm (internal error)");
        stats.hints_taken++;
      }
      hint_ring[(q + 2) & 3][0] = hint_ring[(q + 2) & 3][1] = NO_HINT;
      last_mp = ~0ull;
      if (hints && !(flags & ~64u) && n <= recs_.size()) {  // the largest frequency this merge makes the heap push (records carry the new frequency)
        uint64_t mp = 0;
        for (size_t i = 0; i < n; i++) { const uint64_t f = (uint64_t)hr[i].delta; if (f >= minf && f > mp) mp = f; }
        last_mp = mp;
      }
      if (npk >= 2 && done + 2 < max_merges && last_mp != ~0ull) {  // early hint for merge q+2
        if (last_mp < pk[1].freq && !no_early) {
          // With this hint the device may start merge q+2 -- whose results go into THIS record buffer -- before the records of
          // merge q have been applied below: they move to private memory first.
          rec_copy_.assign(hr, hr + n);
          hr = rec_copy_.data();
          send_hint(q + 2, pk[1], 0, sender);
          n_early++;
        }
      }
      if (flags & 5u) throw Error("pair table overflow during a merge (internal sizing error)");
      if (flags & 32u) throw Error("birth log overflow (internal sizing error)");
      if (flags & 16u) gt_flagged_ = true;
      op_index_++;
      merge_done_on_stream(hflags);
      seq_ = seq_base + done + 1;
      stats.merge_scan_bytes += n_rows_ * ROW * 4;
      stats.merge_alg_bytes += 4 * live_symbols_ + 8 * W;
      live_symbols_ -= removed;
      stats.live_symbols = live_symbols_;
      translate_out(hr, n);
      core.apply_absolute(hr, n);
      const double ta1 = now_ms();
      stats.host_apply_ms += ta1 - tw1;
      done++;
      bool go = done < max_merges && !gt_flagged_;
      int32_t na = 0, nb = 0, nn = 0;
      if (go) { go = core.next_merge(&na, &nb, &nn); stats.host_pop_ms += now_ms() - ta1; }
      sender.next_seq = seq_base + done + 1;
      if (!go) { sender.send(0, 0, 1); sender.running = false; break; }
      da = to_dev(na); db = to_dev(nb);
      cur_key = ((unsigned long long)(uint32_t)da << 32) | (uint32_t)db;
      sender.send(cur_key, (unsigned int)nn, 0);
      look_ahead(na, nb);
      if (npk >= 1 && done + 1 < max_merges) send_hint(seq_base + done + 2, pk[0], 1, sender);  // late hint for the merge after this command's
    }
    if (launcher.th.joinable()) launcher.th.join();  // (a serialising tool: the launch call returns now that the kernel has been told to stop)
    if (launcher.err) std::rethrow_exception(launcher.err);
    sync();
    {  // device time of this launch's merges (command seen -> result published), per mode
      check_debug_bounds();
      unsigned long long ac[32];
      SWB_CUDA(cudaMemcpy(ac, cl_acct_.get(), sizeof ac, cudaMemcpyDeviceToHost));
      for (int i = 0; i < 4; i++) { stats.local_by_log[i] += ac[16 + i]; stats.local_by_log_ms[i] += (double)ac[24 + i] * 1e-6; stats.local_by_log_recs[i] += ac[20 + i]; }
      stats.hints_rejected += ac[5]; stats.resident_spill_merges += ac[7];
      stats.resident_local_merges += ac[0]; stats.resident_local_ms += (double)ac[1] * 1e-6;
      stats.resident_grid_merges += ac[2]; stats.resident_grid_ms += (double)ac[3] * 1e-6;
      stats.merge_kernel_ms += (double)(ac[1] + ac[3]) * 1e-6;
    }
    if (trace_p) {
      unsigned long long ac[8];
      SWB_CUDA(cudaMemcpy(ac, cl_acct_.get(), sizeof ac, cudaMemcpyDeviceToHost));
      fprintf(stderr, "[trace] hints: %llu sent a merge ahead; device accepted %llu (%llu of them already loaded when needed), rejected %llu; %llu LOCAL merges spilled\n", n_early, ac[4], ac[6], ac[5], ac[7]);
      unsigned long long h32[32];
      SWB_CUDA(cudaMemcpy(h32, trace_p, sizeof h32, cudaMemcpyDeviceToHost));
      fprintf(stderr, "[trace] GRID merges by device time (<16, <24, <32, <48, <64, <128, <256, more us): count/ms");
      for (int i = 0; i < 8; i++) fprintf(stderr, " %llu/%.1f", h32[16 + i], (double)h32[24 + i] * 1e-6);
      fprintf(stderr, "\n");
    }
    if (trace_p) {
      unsigned long long h[16];
      SWB_CUDA(cudaMemcpy(h, trace_p, sizeof h, cudaMemcpyDeviceToHost));
      const double n = (double)std::max<unsigned long long>(h[0], 1), ghz = 1.965;
      fprintf(stderr, "[trace] cluster kernel, per LOCAL merge (%llu of %d merges): words+deltas %.2f us, apply+records %.2f us, publish %.2f us, "
              "waiting for the host %.2f us | words+deltas: thread 0 alone %.2f us, its CTA %.2f us; log entries %.0f\n", h[0], done, h[1] / n / ghz / 1e3, h[2] / n / ghz / 1e3, h[3] / n / ghz / 1e3, h[4] / n / ghz / 1e3,
              h[5] / n / ghz / 1e3, h[7] / n / ghz / 1e3, h[6] / n);
      fprintf(stderr, "[trace]   thread 0: entries arrived after %.2f us, own words %.2f us\n", h[8] / n / ghz / 1e3, h[9] / n / ghz / 1e3);
      unsigned long long wt[8];
      if (cudaMemcpyFromSymbol(wt, g_word_trace, sizeof wt) == cudaSuccess && wt[6]) {
        const double nw = (double)wt[6];
        fprintf(stderr, "[trace]   one word (block 0 / thread 0, %llu words): load+parse %.2f us, rewrite %.2f us, reconverge %.2f us, deltas %.2f us, write back+births %.2f us\n",
                wt[6], wt[0] / nw / ghz / 1e3, wt[1] / nw / ghz / 1e3, wt[2] / nw / ghz / 1e3, wt[3] / nw / ghz / 1e3, wt[4] / nw / ghz / 1e3);
      }
    }
    return done;
  }

  // A device-side failure in the middle of a merge leaves the symbol stream, the frequency table and the heap replica
  // out of step: the handle is marked failed (every later compute call reports the first error again), and the merge
  // that was announced but not applied is taken off the merge list, so num_merges / merge_ops / bpe_save stay consistent.
  std::string failed_;
  void check_not_failed() const {
    if (!failed_.empty()) throw Error("this trainer failed earlier and cannot continue (destroy it): " + failed_);
  }
  int merge_batch(int batch) {  // reference bpe_merge_batch
    check_not_failed();
    try {
      return merge_batch_unguarded(batch);
    } catch (const std::exception &e) {
      core.abort_pending();
      failed_ = e.what();
      throw;
    }
  }
  int merge_batch_unguarded(int batch) {
    const double t0 = now_ms();
    int done = 0;
    if (use_resident()) {
      while (done < batch && !core.heap_empty()) {
        int32_t a, b, nid;
        const double tp0 = now_ms();
        if (!core.next_merge(&a, &b, &nid)) break;
        stats.host_pop_ms += now_ms() - tp0;
        tables_fresh_ = false;
        done += run_resident(a, b, nid, batch - done);
      }
      stats.merge_ms += now_ms() - t0;
      return done;
    }
    while (done < batch && !core.heap_empty()) {
      int32_t a, b, nid;
      const double tp0 = now_ms();
      if (!core.next_merge(&a, &b, &nid)) break;
      stats.host_pop_ms += now_ms() - tp0;
      tables_fresh_ = false;
      size_t n = 0;
      const Rec *r = run_merge(a, b, nid, &n);
      const double ta0 = now_ms();
      if (device_tables_) core.apply_absolute(r, n);
      else core.apply(r, n);
      stats.host_apply_ms += now_ms() - ta0;
      done++;
    }
    stats.merge_ms += now_ms() - t0;
    return done;
  }

  // ---------------------------------------------------------------- results
  void token_freq(std::vector<uint64_t> &freq) {
    const size_t T = 256 + tr_->num_merges;
    freq.assign(T, 0);
    if (!loaded_) return;
    DevBuf<unsigned long long> d(T);
    SWB_CUDA(cudaMemsetAsync(d.get(), 0, d.bytes(), stream_));
    StreamDev s = stream_dev();
    if (n_rows_) { tokfreq_rows<<<(int)std::min<uint64_t>((uint64_t)sms_ * 8, (n_rows_ + 7) / 8), MERGE_THREADS, 0, stream_>>>(s, d.get(), (uint32_t)T); launched(); }
    if (n_long_) { tokfreq_long<<<std::min<uint32_t>(sms_ * 4, (n_long_ + 3) / 4), 128, 0, stream_>>>(s, d.get(), (uint32_t)T); launched(); }
    SWB_CUDA(cudaMemcpyAsync(freq.data(), d.get(), T * 8, cudaMemcpyDeviceToHost, stream_));
    sync();
  }
  uint64_t num_symbols() const { return live_symbols_; }
  uint64_t word_bytes_total() const { return word_bytes_total_; }
  bool loaded() const { return loaded_; }

  // word table -> host (this rank's words have their symbols; other ranks' words come back empty)
  void get_words(uint64_t *byte_off, uint8_t *bytes, uint64_t *sym_off, int32_t *syms, uint64_t *counts) {
    if (!loaded_ || W == 0) {
      if (byte_off) byte_off[0] = 0;
      if (sym_off) sym_off[0] = 0;
      return;
    }
    if (byte_off) SWB_CUDA(cudaMemcpyAsync(byte_off, word_boff_.get(), (W + 1) * 8, cudaMemcpyDeviceToHost, stream_));
    if (bytes && word_bytes_total_) SWB_CUDA(cudaMemcpyAsync(bytes, word_bytes_.get(), word_bytes_total_, cudaMemcpyDeviceToHost, stream_));
    if (counts) memcpy(counts, h_counts.data(), W * 8);
    if (sym_off || syms) {
      StreamDev s = stream_dev();
      DevBuf<uint32_t> len(W);
      words_live_len<<<sms_ * 4, 256, 0, stream_>>>(s, wloc_.get(), long_index_.get(), (uint32_t)W, mrank(), mnranks(), len.get()); launched();
      std::vector<uint32_t> h_len(W);
      SWB_CUDA(cudaMemcpyAsync(h_len.data(), len.get(), W * 4, cudaMemcpyDeviceToHost, stream_));
      sync();
      std::vector<uint64_t> off(W + 1);
      uint64_t acc = 0;
      for (uint64_t w = 0; w < W; w++) { off[w] = acc; acc += h_len[w]; }
      off[W] = acc;
      if (sym_off) memcpy(sym_off, off.data(), (W + 1) * 8);
      if (syms && acc) {
        DevBuf<uint64_t> d_off(W + 1);
        DevBuf<int32_t> d_out(acc);
        SWB_CUDA(cudaMemcpyAsync(d_off.get(), off.data(), (W + 1) * 8, cudaMemcpyHostToDevice, stream_));
        words_copy_syms<<<sms_ * 4, 256, 0, stream_>>>(s, wloc_.get(), long_index_.get(), (uint32_t)W, mrank(), mnranks(), d_off.get(), d_out.get()); launched();
        SWB_CUDA(cudaMemcpyAsync(syms, d_out.get(), acc * 4, cudaMemcpyDeviceToHost, stream_));
        sync();
        if (tr_->config.unk_id < 0)
          for (uint64_t i = 0; i < acc; i++) if (syms[i] == UNK_CODE) syms[i] = tr_->config.unk_id;
      }
    }
    sync();
  }

 private:
  void free_corpus_state() {
    rows_.release(); sig_.release(); cnt_.release(); wloc_.release(); long_index_.release(); long_syms_.release(); long_off_.release();
    long_len_.release(); long_word_.release(); word_bytes_.release(); word_boff_.release();
    log_ent_.release(); log_ok_ = false; stream_merges_ = 0; log_cap_ = 0;
    ip_ent_.release(); ip_start_.release();
    n_rows_ = 0; n_long_ = 0; W = 0; live_symbols_ = 0; word_bytes_total_ = 0; loaded_ = false;
    h_counts.clear();
  }

  Trainer *tr_;
  int device_ = 0, sms_ = 148;
  cudaStream_t stream_ = nullptr, copy_stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  bool loaded_ = false;
  double t_launch0_ = 0;
  unsigned long long seq_ = 0;
  DevBuf<unsigned int> scalars_;
  PinnedBuf<unsigned long long> hdr_, hdr_b_;
  // symbol stream
  DevBuf<int4> rows_;
  DevBuf<uint32_t> sig_;
  uint64_t n_rows_ = 0;
  DevBuf<unsigned long long> cnt_;
  DevBuf<uint64_t> wloc_;
  DevBuf<uint4> ip_ent_;            // occurrence index of the pairs of two initial symbols (see BirthLogDev)
  DevBuf<unsigned int> ip_start_;
  DevBuf<uint32_t> long_index_;
  DevBuf<int32_t> long_syms_;
  DevBuf<uint64_t> long_off_;
  DevBuf<uint32_t> long_len_, long_word_;
  uint32_t n_long_ = 0;
  uint64_t live_symbols_ = 0, word_bytes_total_ = 0;
  DevBuf<uint8_t> word_bytes_;
  DevBuf<uint64_t> word_boff_;
  // birth log (pair -> candidate rows), see BirthLogDev
  DevBuf<uint4> log_ent_;
  DevBuf<unsigned int> log_scal_, log_start_;
  uint32_t log_cap_ = 0, log_merge_cap_ = 0, stream_merges_ = 0;
  bool log_ok_ = false;
  std::vector<uint32_t> log_start_h_;  // host copy of the log ranges: [m] .. [m+1] = log of merge m (from the result headers)
  // pair table
  uint64_t pt_cap_ = 0;
  PairTableDev pt_{};
  DevBuf<unsigned long long> pt_keys_, pt_val_, pt_min_, removed_, emit_partial_;
  int idle_polls_ = 0;
  std::atomic<int> launch_returned_{1};  // 0 while the launch call of the resident kernel is still inside the driver (see run_resident)
  uint64_t gt_used_estimate_ = 0;
  bool gt_flagged_ = false;
  // persistent merge kernel
  bool coop_ok_ = false;
  PinnedBuf<HostCmd2> hcmd2_;
  DevBuf<DevCmd2> dcmd2_;
  DevBuf<uint4> cl_ovf_;
  DevBuf<unsigned long long> cl_acct_;
  DevBuf<unsigned long long> ptrace_;
  // device-resident loop
  bool trace_wait_ = getenv("SWB_TRACE_WAIT") != nullptr;
  bool resident_wait_ = false;  // wait_seq_at is waiting for a result of the resident kernel (its status word is meaningful)
  std::vector<float> wait_trace_;
  std::vector<uint32_t> trace_removed_, trace_nrec_, trace_cand_, trace_mrows_;
  std::vector<uint64_t> trace_logn_;
  std::vector<float> trace_scan_, trace_tail_;
  double tw_sig_ = 0, tw_rows_ = 0, tw_fence_ = 0, tw_cand_ = 0; size_t tw_n_ = 0;
  // device-resident frequency table (single-GPU mode)
  bool device_tables_ = false;   // true: the device owns the frequencies, the host only sees pairs >= min_pair_freq
  bool tables_fresh_ = true;     // no count / merge since the last load or reset
  uint64_t gt_cap_ = 0, op_index_ = 0;
  GlobalTableDev gt_{};
  DevBuf<GSlot> gt_slots_;
  DevBuf<unsigned int> gt_scal_;
  DevBuf<uint4> pt_touched_;
  DevBuf<unsigned int> pt_scal_;
  // multi-GPU
  NcclComm comm_ = nullptr;
  PairTableDev pt2_{};
  DevBuf<unsigned long long> pt2_keys_, pt2_val_, pt2_min_;
  DevBuf<uint4> pt2_touched_;
  DevBuf<unsigned int> pt2_scal_;
  DevBuf<unsigned long long> d_all_;
  size_t dist_cap_ = 0, dist_slot_words_ = 0;
  PinnedBuf<Rec> recs_, recs_b_;
  std::vector<Rec> rec_copy_;
};

}  // namespace swb
