// capi.cu -- extern "C" surface of libtrainer.so (declared in include/shredword_b200.h).
//
// Part 1 is the reference's own ABI (reference shredword/csrc/bpe/bpe.h:62-72, bound by
// shredword/cbase.py:44-59); Part 2 is additive. No entry point has a CPU fallback: without a CUDA
// device the compute calls fail with an error code and a message in swb_last_error().
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/shredword_b200.h"
#include "encoder_impl.cuh"
#include "normalize.cuh"
#include "pretok.cuh"
#include "trainer_impl.cuh"

using swb::EncoderImpl;
using swb::Rec;
using swb::TrainerImpl;

static_assert(sizeof(BPEConfig) == 24, "BPEConfig layout must match reference bpe.h:43-48");
static_assert(sizeof(HeapEntry) == 24, "HeapEntry layout must match reference heap.h:17-21");
static_assert(offsetof(Trainer, heap) == 24 && offsetof(Trainer, corpus) == 48 && offsetof(Trainer, bigram_map) == 72 &&
                  offsetof(Trainer, num_merges) == 96 && offsetof(Trainer, merge_ops) == 104 && offsetof(Trainer, impl) == 128,
              "Trainer prefix must match reference bpe.h:50-60");

static thread_local std::string g_err;
static int g_log_level = -1;

static int log_level() {
  if (g_log_level < 0) {
    const char *e = getenv("SHREDWORD_LOG");
    g_log_level = (e && atoi(e) > 0) ? atoi(e) : 0;
  }
  return g_log_level;
}
static void set_err(const std::string &m) {
  g_err = m;
  fprintf(stderr, "[ERROR]\t %s\n", m.c_str());
}
static TrainerImpl *impl_of(const Trainer *t) { return t ? static_cast<TrainerImpl *>(t->impl) : nullptr; }

#define SWB_TRY try {
#define SWB_CATCH(ret)                                     \
  }                                                        \
  catch (const std::exception &e) { set_err(e.what()); return ret; } \
  catch (...) { set_err("unknown error"); return ret; }

extern "C" {

// ------------------------------------------------------------------ Part 1: reference ABI

Trainer *create_trainer(const BPEConfig *config) {
  if (!config) { set_err("Config pointer is NULL"); return nullptr; }
  SWB_TRY
  Trainer *t = static_cast<Trainer *>(calloc(1, sizeof(Trainer)));
  if (!t) { set_err("Couldn't allocate Memory to Trainer"); return nullptr; }
  t->config = *config;
  // reference bpe.cpp:124-130
  if (t->config.character_coverage <= 0.0 || t->config.character_coverage >= 1.0) t->config.character_coverage = 0.995;
  if (t->config.min_pair_freq == 0) t->config.min_pair_freq = 2000;
  t->bigram_map.nbuckets = 4096;
  TrainerImpl *im = new TrainerImpl(t);
  im->core.log_level = log_level();
  im->core.heap_reset();  // reference bpe.cpp:133 heap_init(MIN_HEAP_SIZE)
  t->impl = im;
  if (log_level() > 0) printf("[INFO]\t BPE trainer initialized. Heap initialized successfully.\n");
  return t;
  SWB_CATCH(nullptr)
}

void bpe_trainer_destroy(Trainer *trainer) {
  if (!trainer) return;
  delete impl_of(trainer);
  free(trainer);
}

int bpe_load_corpus(Trainer *trainer, const char *input_path) {
  if (!trainer || !input_path) { set_err("NULL trainer or input path pointers"); return -1; }
  SWB_TRY
  impl_of(trainer)->load_file(input_path);
  return 0;
  SWB_CATCH(-1)
}

void bpe_count_bigrams(Trainer *trainer) {
  if (!trainer) { set_err("NULL trainer pointer"); return; }
  SWB_TRY
  impl_of(trainer)->count_bigrams();
  SWB_CATCH()
}

void bpe_init(Trainer *trainer) {
  if (!trainer) { set_err("NULL trainer pointer"); return; }
  SWB_TRY
  impl_of(trainer)->init();
  SWB_CATCH()
}

int bpe_merge_batch(Trainer *trainer, int batch_size) {
  if (!trainer) { set_err("Trainer pointer is NULL!"); return -1; }
  SWB_TRY
  return impl_of(trainer)->merge_batch(batch_size);
  SWB_CATCH(-1)
}

int bpe_train(Trainer *trainer) {
  if (!trainer) { set_err("Trainer pointer is NULL!"); return -1; }
  SWB_TRY
  TrainerImpl *im = impl_of(trainer);
  if (log_level() > 0) printf("[INFO]\t Starting BPE training (target vocab size: %zu)\n", trainer->config.target_vocab_size);
  im->init();
  // reference bpe.cpp:604-637. The reference picks a "batch size" per iteration from the heap top, but
  // bpe_merge_batch performs its merges strictly one after another, so the batching only paces its
  // logging; one call with the remaining budget is equivalent.
  int total = 0;
  const int target = (int)trainer->config.target_vocab_size - 256;
  while (total < target) {
    if (im->core.heap_empty()) break;
    const int got = im->merge_batch(target - total);
    if (got <= 0) break;
    total += got;
  }
  if (log_level() > 0) printf("[INFO]\t Training completed. Performed %d merges\n", total);
  return total;
  SWB_CATCH(-1)
}

// token strings exactly as the reference builds them (bpe.cpp:687-701): C strings, so a NUL byte
// (token 0, or any token built from it) cuts the piece short.
static void build_c_tokens(const Trainer *t, std::vector<std::string> &tok) {
  const size_t M = t->num_merges;
  tok.assign(256 + M, std::string());
  for (int i = 1; i < 256; i++) tok[i] = std::string(1, (char)i);
  for (size_t m = 0; m < M; m++) {
    const PairKey op = t->merge_ops[m];
    std::string s;
    if (op.first >= 0 && (size_t)op.first < 256 + m) s = tok[op.first];
    if (op.second >= 0 && (size_t)op.second < 256 + m) s += tok[op.second];
    tok[256 + m] = s;
  }
}

static void save_files(const Trainer *trainer, const char *model_path, const char *vocab_path, const uint64_t *freq) {
  const size_t M = trainer->num_merges, T = 256 + M;
  std::vector<std::string> tok;
  build_c_tokens(trainer, tok);
  FILE *vf = fopen(vocab_path, "w");
  if (!vf) throw swb::Error(std::string("cannot open ") + vocab_path);
  for (size_t i = 0; i < T; i++) fprintf(vf, "%s %llu\n", tok[i].c_str(), (unsigned long long)freq[i]);
  fclose(vf);
  FILE *mf = fopen(model_path, "wb");
  if (!mf) throw swb::Error(std::string("cannot open ") + model_path);
  for (size_t m = 0; m < M; m++) {
    const int32_t rec[3] = {trainer->merge_ops[m].first, trainer->merge_ops[m].second, (int32_t)(256 + m)};
    fwrite(rec, sizeof(int32_t), 3, mf);
  }
  fclose(mf);
  if (log_level() > 0) printf("[INFO]\tSaved %zu-token vocab to %s and %zu merges to %s\n", T, vocab_path, M, model_path);
}

void bpe_save(const Trainer *trainer, const char *model_path, const char *vocab_path) {
  if (!trainer || !model_path || !vocab_path) { set_err("Trainer pointer is NULL!"); return; }
  SWB_TRY
  std::vector<uint64_t> freq;
  impl_of(trainer)->token_freq(freq);
  save_files(trainer, model_path, vocab_path, freq.data());
  SWB_CATCH()
}

// ------------------------------------------------------------------ Part 2: additive

const char *swb_last_error(void) { return g_err.c_str(); }
void swb_set_log_level(int level) { g_log_level = level < 0 ? 0 : level; }

int swb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int swb_set_device(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); cudaGetLastError(); return -1; }
  return 0;
}

size_t swb_release_cached_memory(void) {
  const size_t n = swb::BlockCache::device().cached_bytes() + swb::BlockCache::pinned().cached_bytes();
  cudaDeviceSynchronize();
  swb::BlockCache::device().release_all();
  swb::BlockCache::pinned().release_all();
  return n;
}

int swb_load_corpus_buffer(Trainer *trainer, const void *data, size_t nbytes) {
  if (!trainer || (!data && nbytes)) { set_err("NULL trainer or data"); return -1; }
  SWB_TRY
  impl_of(trainer)->load_host(data, nbytes);
  return 0;
  SWB_CATCH(-1)
}
int swb_load_corpus_device(Trainer *trainer, const void *device_data, size_t nbytes) {
  if (!trainer || (!device_data && nbytes)) { set_err("NULL trainer or data"); return -1; }
  SWB_TRY
  impl_of(trainer)->load_device(device_data, nbytes);
  return 0;
  SWB_CATCH(-1)
}

size_t swb_num_merges(const Trainer *trainer) { return trainer ? trainer->num_merges : 0; }
size_t swb_get_merges(const Trainer *trainer, int32_t *out, size_t cap) {
  if (!trainer) return 0;
  const size_t M = trainer->num_merges;
  for (size_t m = 0; m < M && m < cap; m++) {
    out[3 * m] = trainer->merge_ops[m].first; out[3 * m + 1] = trainer->merge_ops[m].second; out[3 * m + 2] = (int32_t)(256 + m);
  }
  return M;
}
size_t swb_token_bytes(const Trainer *trainer, int32_t id, uint8_t *out, size_t cap) {
  if (!trainer || id < 0 || (size_t)id >= 256 + trainer->num_merges) return 0;
  // iterative expansion through the merge list
  std::vector<int32_t> stack{id};
  size_t n = 0;
  while (!stack.empty()) {
    const int32_t x = stack.back(); stack.pop_back();
    if (x >= 256 && (size_t)(x - 256) < trainer->num_merges) {
      stack.push_back(trainer->merge_ops[x - 256].second);
      stack.push_back(trainer->merge_ops[x - 256].first);
    } else if (x >= 0 && x < 256) {
      if (out && n < cap) out[n] = (uint8_t)x;
      n++;
    }
  }
  return n;
}
void swb_get_byte_map(const Trainer *trainer, int32_t *out256) {
  if (!trainer || !out256) return;
  memcpy(out256, impl_of(trainer)->byte_map, 256 * sizeof(int32_t));
}
int swb_token_freq(const Trainer *trainer, uint64_t *out, size_t cap) {
  if (!trainer || !out) return -1;
  SWB_TRY
  std::vector<uint64_t> f;
  impl_of(trainer)->token_freq(f);
  for (size_t i = 0; i < f.size() && i < cap; i++) out[i] = f[i];
  return 0;
  SWB_CATCH(-1)
}
size_t swb_num_words(const Trainer *trainer) { return trainer ? impl_of(trainer)->W : 0; }
size_t swb_num_symbols(const Trainer *trainer) { return trainer ? impl_of(trainer)->num_symbols() : 0; }
size_t swb_word_bytes_total(const Trainer *trainer) { return trainer ? impl_of(trainer)->word_bytes_total() : 0; }
int swb_get_words(const Trainer *trainer, uint64_t *byte_off, uint8_t *bytes, uint64_t *sym_off, int32_t *syms, uint64_t *counts) {
  if (!trainer) return -1;
  SWB_TRY
  impl_of(trainer)->get_words(byte_off, bytes, sym_off, syms, counts);
  return 0;
  SWB_CATCH(-1)
}
void swb_get_stats(const Trainer *trainer, SwbStats *out) {
  if (!trainer || !out) return;
  TrainerImpl *im = impl_of(trainer);
  im->stats.records = im->core.n_records; im->stats.heap_pushes = im->core.n_pushes;
  im->stats.heap_pops = im->core.n_pops; im->stats.heap_peak = im->core.heap_peak;
  *out = im->stats;
}
void swb_set_kernel_timing(Trainer *trainer, int enabled) {
  if (trainer) impl_of(trainer)->timing = enabled != 0;
}

int swb_profile_scripted_merges(Trainer *trainer, const int32_t *merge_triples, size_t n, double *kernel_ms) {
  if (!trainer || (!merge_triples && n)) { set_err("swb_profile_scripted_merges: NULL argument"); return -1; }
  SWB_TRY
  const double ms = impl_of(trainer)->profile_scripted(merge_triples, n);
  if (kernel_ms) *kernel_ms = ms;
  return 0;
  SWB_CATCH(-1)
}

// ---- encoder
struct SwbEncoder { EncoderImpl *impl; };

SwbEncoder *swb_encoder_create(const int32_t *merge_triples, size_t n_merges, const int32_t *byte_map256) {
  if (!merge_triples && n_merges) { set_err("NULL merges"); return nullptr; }
  SWB_TRY
  SwbEncoder *e = new SwbEncoder;
  e->impl = new EncoderImpl(merge_triples, n_merges, byte_map256, -1);
  return e;
  SWB_CATCH(nullptr)
}
SwbEncoder *swb_encoder_from_trainer(const Trainer *trainer) {
  if (!trainer) { set_err("NULL trainer"); return nullptr; }
  SWB_TRY
  std::vector<int32_t> m(3 * trainer->num_merges + 3);
  swb_get_merges(trainer, m.data(), trainer->num_merges);
  SwbEncoder *e = new SwbEncoder;
  e->impl = new EncoderImpl(m.data(), trainer->num_merges, impl_of(trainer)->byte_map, trainer->config.unk_id);
  return e;
  SWB_CATCH(nullptr)
}
void swb_encoder_destroy(SwbEncoder *enc) {
  if (!enc) return;
  delete enc->impl;
  delete enc;
}
int64_t swb_encode(SwbEncoder *enc, const void *text, size_t nbytes, int32_t *out_ids, size_t cap_ids, uint32_t *word_ntok,
                   size_t cap_words, size_t *n_words) {
  if (!enc || (!text && nbytes) || (!out_ids && cap_ids)) { set_err("swb_encode: NULL argument"); return -1; }
  SWB_TRY
  return enc->impl->encode_host(static_cast<const uint8_t *>(text), nbytes, out_ids, cap_ids, word_ntok, cap_words, n_words);
  SWB_CATCH(-1)
}
int64_t swb_encode_device(SwbEncoder *enc, const void *d_text, size_t nbytes, int32_t *d_out_ids, size_t cap_ids,
                          uint32_t *d_word_ntok, size_t cap_words, size_t *n_words) {
  if (!enc || (!d_text && nbytes)) { set_err("swb_encode_device: NULL argument"); return -1; }
  SWB_TRY
  return enc->impl->encode_device(static_cast<const uint8_t *>(d_text), nbytes, d_out_ids, cap_ids, d_word_ntok, cap_words, n_words);
  SWB_CATCH(-1)
}
size_t swb_decode(const SwbEncoder *enc, const int32_t *ids, size_t n, uint8_t *out, size_t cap) {
  if (!enc || (!ids && n)) return 0;
  return enc->impl->decode(ids, n, out, cap);
}
uint64_t swb_encoder_kernel_launches(const SwbEncoder *enc) { return enc ? enc->impl->launches : 0; }

// ---- optional pre-pass: the reference's normalize_line on every line (reference csrc/bpe/normalize.cpp:24-59)
int64_t swb_normalize(const void *text, size_t nbytes, void *out, size_t cap, int on_device) {
  if ((!text && nbytes) || (!out && cap)) { set_err("swb_normalize: NULL argument"); return -1; }
  SWB_TRY
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) { cudaGetLastError(); throw swb::Error("no CUDA device (this library has no CPU fallback)"); }
  int dev = 0, sms = 0;
  SWB_CUDA(cudaGetDevice(&dev));
  SWB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st;
  SWB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  uint64_t total = 0;
  try {
    if (on_device) {
      total = swb::normalize_device(static_cast<const uint8_t *>(text), nbytes, static_cast<uint8_t *>(out), cap, st, sms, nullptr);
    } else {
      swb::DevBuf<uint8_t> d_in(nbytes + 16), d_out(cap + 16);
      if (nbytes) SWB_CUDA(cudaMemcpyAsync(d_in.get(), text, nbytes, cudaMemcpyHostToDevice, st));
      total = swb::normalize_device(d_in.get(), nbytes, d_out.get(), cap, st, sms, nullptr);
      const size_t back = (size_t)std::min<uint64_t>(total, cap);
      if (back) SWB_CUDA(cudaMemcpyAsync(out, d_out.get(), back, cudaMemcpyDeviceToHost, st));
      SWB_CUDA(cudaStreamSynchronize(st));
    }
  } catch (...) { cudaStreamDestroy(st); throw; }
  cudaStreamDestroy(st);
  return (int64_t)total;
  SWB_CATCH(-1)
}

// ---- optional pre-pass: the reference's regex pre-tokenisation (reference shredword/base.py:38-58, apply_regex)
int64_t swb_pretokenize(const void *text, size_t nbytes, void *out, size_t cap, int on_device) {
  if ((!text && nbytes) || (!out && cap)) { set_err("swb_pretokenize: NULL argument"); return -1; }
  SWB_TRY
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) { cudaGetLastError(); throw swb::Error("no CUDA device (this library has no CPU fallback)"); }
  int dev = 0, sms = 0;
  SWB_CUDA(cudaGetDevice(&dev));
  SWB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st;
  SWB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  uint64_t total = 0;
  try {
    if (on_device) {
      total = swb::pretokenize_device(static_cast<const uint8_t *>(text), nbytes, static_cast<uint8_t *>(out), cap, st, dev, sms, nullptr);
    } else {
      swb::DevBuf<uint8_t> d_in(nbytes + 16), d_out(cap + 16);
      if (nbytes) SWB_CUDA(cudaMemcpyAsync(d_in.get(), text, nbytes, cudaMemcpyHostToDevice, st));
      total = swb::pretokenize_device(d_in.get(), nbytes, d_out.get(), cap, st, dev, sms, nullptr);
      const size_t back = (size_t)std::min<uint64_t>(total, cap);
      if (back) SWB_CUDA(cudaMemcpyAsync(out, d_out.get(), back, cudaMemcpyDeviceToHost, st));
      SWB_CUDA(cudaStreamSynchronize(st));
    }
  } catch (...) { cudaStreamDestroy(st); throw; }
  cudaStreamDestroy(st);
  return (int64_t)total;
  SWB_CATCH(-1)
}

// ---- multi-GPU building blocks
int swb_device_pci_bus_id(int device, char *out, size_t cap) {
  if (!out || cap < 16) { set_err("swb_device_pci_bus_id: buffer too small"); return -1; }
  if (cudaDeviceGetPCIBusId(out, (int)cap, device) != cudaSuccess) { cudaGetLastError(); set_err("swb_device_pci_bus_id: no such CUDA device"); return -1; }
  return 0;
}
int swb_set_shard(Trainer *trainer, int rank, int nranks) {
  if (!trainer || nranks < 1 || rank < 0 || rank >= nranks) { set_err("swb_set_shard: bad arguments"); return -1; }
  impl_of(trainer)->rank = rank;
  impl_of(trainer)->nranks = nranks;
  impl_of(trainer)->replicated_ = false;  // the caller drives the per-merge exchange itself: this rank keeps only its words
  return 0;
}
int swb_dist_set_sharded(Trainer *trainer, int sharded) {
  if (!trainer) { set_err("swb_dist_set_sharded: NULL trainer"); return -1; }
  impl_of(trainer)->replicated_ = sharded == 0;
  return 0;
}
size_t swb_dist_reduce_records(int64_t *recs, size_t n) { return swb::reduce_records(reinterpret_cast<Rec *>(recs), n); }
void swb_dist_seed(Trainer *trainer, const int64_t *recs, size_t n) {
  if (!trainer) return;
  TrainerImpl *im = impl_of(trainer);
  im->reset_tables();
  im->mark_used();
  im->core.seed_counts(reinterpret_cast<const Rec *>(recs), n);
}
int swb_dist_next_merge(Trainer *trainer, int32_t *a, int32_t *b, int32_t *new_id) {
  if (!trainer || !a || !b || !new_id) return 0;
  return impl_of(trainer)->core.next_merge(a, b, new_id) ? 1 : 0;
}
int swb_dist_peek_next(Trainer *trainer, int32_t *a, int32_t *b, uint64_t *freq) {
  if (!trainer || !a || !b || !freq) return 0;
  return impl_of(trainer)->core.peek_next(a, b, freq) ? 1 : 0;
}
size_t swb_dist_peek_list(Trainer *trainer, int64_t *out, size_t want) {
  if (!trainer || !out) return 0;
  swb::HostCore::Peek p[16];
  if (want > 16) want = 16;
  const size_t n = impl_of(trainer)->core.peek_next(p, want);
  for (size_t i = 0; i < n; i++) { out[3 * i] = p[i].a; out[3 * i + 1] = p[i].b; out[3 * i + 2] = (int64_t)p[i].freq; }
  return n;
}
void swb_dist_apply(Trainer *trainer, const int64_t *recs, size_t n) {
  if (!trainer) return;
  impl_of(trainer)->core.apply(reinterpret_cast<const Rec *>(recs), n);
}
int swb_dist_unique_id(void *out128) {
  if (!out128) return -1;
  SWB_TRY
  swb::NcclApi &api = swb::NcclApi::get();
  if (!api.ok()) throw swb::Error("NCCL is not available: " + api.error());
  swb::NcclUniqueId id;
  api.check(api.GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, &id, sizeof id);
  return 0;
  SWB_CATCH(-1)
}
int swb_dist_init(Trainer *trainer, int rank, int nranks, const void *unique_id128) {
  if (!trainer || nranks < 1 || rank < 0 || rank >= nranks) { set_err("swb_dist_init: bad arguments"); return -1; }
  SWB_TRY
  swb::NcclUniqueId id;
  if (unique_id128) memcpy(&id, unique_id128, sizeof id);
  impl_of(trainer)->dist_init(rank, nranks, unique_id128 ? &id : nullptr);
  return 0;
  SWB_CATCH(-1)
}
int swb_load_corpus_shard(Trainer *trainer, const void *data, size_t nbytes, uint64_t global_offset, int on_device) {
  if (!trainer || (!data && nbytes)) { set_err("NULL trainer or data"); return -1; }
  SWB_TRY
  impl_of(trainer)->load_shard(data, nbytes, global_offset, on_device != 0);
  return 0;
  SWB_CATCH(-1)
}
int swb_dist_has_comm(int rank, int nranks) { return TrainerImpl::have_shared_comm(rank, nranks) ? 1 : 0; }
void swb_dist_shutdown(void) {
  SWB_TRY
  cudaDeviceSynchronize();
  TrainerImpl::destroy_shared_comm();
  SWB_CATCH()
}
int64_t swb_shard_count(Trainer *trainer, int64_t *recs, size_t cap) {
  if (!trainer) return -1;
  SWB_TRY
  size_t n = 0;
  const Rec *r = impl_of(trainer)->shard_count(&n);
  if (n > cap) throw swb::Error("swb_shard_count: record buffer too small");
  if (n) memcpy(recs, r, n * sizeof(Rec));
  return (int64_t)n;
  SWB_CATCH(-1)
}
int64_t swb_shard_merge(Trainer *trainer, int32_t a, int32_t b, int32_t new_id, int64_t *recs, size_t cap) {
  if (!trainer) return -1;
  SWB_TRY
  size_t n = 0;
  const Rec *r = impl_of(trainer)->shard_merge(a, b, new_id, &n);
  if (n > cap) throw swb::Error("swb_shard_merge: record buffer too small");
  if (n) memcpy(recs, r, n * sizeof(Rec));
  return (int64_t)n;
  SWB_CATCH(-1)
}

int swb_save_with_freq(const Trainer *trainer, const char *model_path, const char *vocab_path, const uint64_t *freq,
                       size_t n_freq) {
  if (!trainer || !model_path || !vocab_path || !freq) { set_err("swb_save_with_freq: NULL argument"); return -1; }
  if (n_freq < 256 + trainer->num_merges) { set_err("swb_save_with_freq: freq too short"); return -1; }
  SWB_TRY
  save_files(trainer, model_path, vocab_path, freq);
  return 0;
  SWB_CATCH(-1)
}

}  // extern "C"
