/* shredword_b200.h -- C-ABI of shredword_b200/libtrainer.so
 *
 * A B200-native (sm_100a) replacement for the native core behind ShredWord's
 * `shredword.trainer.BPETrainer`. Part 1 below is, symbol for symbol, what the
 * reference's ctypes binding loads (reference shredword/cbase.py:44-59, declared in
 * reference shredword/csrc/bpe/bpe.h:62-72): a maintainer drops this library into the
 * reference package directory as `libtrainer.so` and nothing else changes
 * (INTEGRATION.md). Part 2 is additive (the reference has no such entry points):
 * host/device buffer loading, result accessors, the rank-ordered encoder, and the
 * building blocks of the multi-GPU merge loop.
 *
 * All signatures are plain C: pointers, sizes, fixed-width integers. No torch types.
 * Every compute entry point needs a CUDA device; without one it fails loudly
 * (error return + message from swb_last_error()), never by falling back to the CPU.
 */
#ifndef SHREDWORD_B200_H
#define SHREDWORD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ Part 1: reference ABI */

/* reference csrc/bpe/hash.h:27-29 */
typedef struct PairKey { int32_t first, second; } PairKey;

/* reference csrc/bpe/heap.h:17-21 (24 bytes) */
typedef struct HeapEntry { PairKey key; uint64_t freq; uint32_t version; } HeapEntry;

/* reference csrc/bpe/heap.h:23-27. Live: `data[0..size)` is the exact binary-heap array the
 * reference would hold at the same point (same order, same ties). */
typedef struct MaxHeap { HeapEntry *data; size_t size; size_t cap; } MaxHeap;

/* reference csrc/bpe/bpe.h:37-41. `vocab_size` (unique words) and `word_counts` (host mirror, in
 * reference word order) are live. `words` is an OPAQUE non-NULL handle once a corpus is loaded:
 * the word table lives in HBM, not in malloc'd Symbol chains; use swb_get_words(). */
typedef struct Corpus { void **words; uint64_t *word_counts; size_t vocab_size; } Corpus;

/* reference csrc/bpe/hash.h:41-44 (field-less in cbase.py:33). Not populated: the pair table
 * is an open-addressing map inside the handle. nbuckets is kept at 4096 for readers. */
typedef struct BIMap { void **buckets; size_t nbuckets; } BIMap;

/* reference csrc/bpe/bpe.h:43-48, cbase.py:40 (24 bytes: size_t @0, int32 @8, float @12, uint64 @16) */
typedef struct BPEConfig {
  size_t target_vocab_size;
  int32_t unk_id;            /* id given to bytes dropped by character_coverage */
  float character_coverage;  /* outside (0,1) -> 0.995 (reference bpe.cpp:124-126) */
  uint64_t min_pair_freq;    /* 0 -> 2000 (reference bpe.cpp:128-130) */
} BPEConfig;

/* reference csrc/bpe/bpe.h:50-60: same field order and offsets (config 0, heap 24, corpus 48,
 * bigram_map 72, next_token 88, num_merges 96, merge_ops 104, token_strs 112, token_freq 120),
 * so code that peeks at config / heap / num_merges / merge_ops (reference test/bpe_test.cpp)
 * keeps working. `impl` (offset 128) is private. */
typedef struct Trainer {
  BPEConfig config;
  MaxHeap heap;
  Corpus corpus;
  BIMap bigram_map;
  size_t next_token;
  size_t num_merges;
  PairKey *merge_ops;   /* merge_ops[m] = pair merged into id 256+m, m < num_merges */
  char **token_strs;    /* unused by the reference too (bpe.h:58) */
  uint64_t *token_freq; /* unused by the reference too (bpe.h:59) */
  void *impl;
} Trainer;

/* reference bpe.cpp:112-136. NULL config -> returns NULL (the reference exit()s; trainer.py:14
 * already turns NULL into RuntimeError). */
Trainer *create_trainer(const BPEConfig *config);
/* reference bpe.cpp:148-158. NULL is ignored (the reference exit()s). */
void bpe_trainer_destroy(Trainer *trainer);
/* reference bpe.cpp:208-297. 0 on success, -1 on error (missing file, no CUDA device,
 * NUL byte in the corpus). Reads the file, builds the unique-word table in HBM. */
int bpe_load_corpus(Trainer *trainer, const char *input_path);
/* reference bpe.cpp:171-185: reset pair table + heap, recount. */
void bpe_init(Trainer *trainer);
/* reference bpe.cpp:315-370: count adjacent pairs, push those with freq >= min_pair_freq. */
void bpe_count_bigrams(Trainer *trainer);
/* reference bpe.cpp:391-535: up to batch_size merges; returns merges done, -1 on NULL / device error. */
int bpe_merge_batch(Trainer *trainer, int batch_size);
/* reference bpe.cpp:597-655: bpe_init + merges until target_vocab_size-256 or heap exhausted. */
int bpe_train(Trainer *trainer);
/* reference bpe.cpp:678-739: vocab text file + binary merge file, byte-identical formats. */
void bpe_save(const Trainer *trainer, const char *model_path, const char *vocab_path);

/* ------------------------------------------------------------------ Part 2: additive */

/* Last error message of the calling thread ("" if none). */
const char *swb_last_error(void);
/* 0 = silent (default), 1 = the reference's [INFO]/[MERGE] lines on stdout. Env SHREDWORD_LOG=1 too. */
void swb_set_log_level(int level);
/* Number of visible CUDA devices (0 when there is none / no driver). Never throws. */
int swb_device_count(void);
/* Select the device used by handles created afterwards on this thread (default 0). */
int swb_set_device(int device);

/* The library keeps freed device / pinned blocks for reuse by later handles (allocation calls are slow
 * and synchronise the device). This returns them to the driver; returns the number of bytes released. */
size_t swb_release_cached_memory(void);

/* Same as bpe_load_corpus but from memory. `data` is a HOST pointer (pinned or pageable) for
 * _buffer and a DEVICE pointer for _device (the bytes are only read, never kept). */
int swb_load_corpus_buffer(Trainer *trainer, const void *data, size_t nbytes);
int swb_load_corpus_device(Trainer *trainer, const void *device_data, size_t nbytes);

/* Results. Merges are (a, b, new_id) triples in rank order (README.md:95 of the reference). */
size_t swb_num_merges(const Trainer *trainer);
size_t swb_get_merges(const Trainer *trainer, int32_t *out_triples, size_t cap_triples);
/* Bytes of token `id` (true bytes: unlike the .vocab file, not cut at NUL). Returns the length. */
size_t swb_token_bytes(const Trainer *trainer, int32_t id, uint8_t *out, size_t cap);
/* byte -> initial id map after character_coverage (identity for kept bytes, unk_id otherwise). */
void swb_get_byte_map(const Trainer *trainer, int32_t *out256);
/* Token histogram of the current segmentation (the .vocab frequency column), ids < 256+num_merges. */
int swb_token_freq(const Trainer *trainer, uint64_t *out, size_t cap);

/* Word table, in reference word order (device -> host copy; for tests and tools).
 * Pass NULL for parts you do not need. Sizes: swb_num_words / swb_num_symbols / swb_word_bytes_total. */
size_t swb_num_words(const Trainer *trainer);
size_t swb_num_symbols(const Trainer *trainer); /* live symbols */
size_t swb_word_bytes_total(const Trainer *trainer);
int swb_get_words(const Trainer *trainer, uint64_t *byte_off /*[W+1]*/, uint8_t *bytes,
                  uint64_t *sym_off /*[W+1]*/, int32_t *syms, uint64_t *counts /*[W]*/);

/* PCI bus id of CUDA device `device` ("0000:1b:00.0") into out; 0 on success. For NUMA placement of the calling
 * thread (the merge loop is a latency chain through mapped host memory). */
int swb_device_pci_bus_id(int device, char *out, size_t cap);

/* Counters of the device work done so far by this handle. */
typedef struct SwbStats {
  uint64_t kernel_launches;   /* kernels of this library launched */
  uint64_t merge_launches;    /* launches of the merge-scan kernel */
  double load_ms, count_ms, merge_ms;  /* host wall time per phase (includes syncs) */
  double merge_kernel_ms;     /* device time of the merge-scan kernel (CUDA events), if enabled */
  uint64_t merge_scan_bytes;  /* sum over merges of bytes the scan kernel was launched over */
  uint64_t merge_alg_bytes;   /* sum over merges of 4*S_live + 8*W */
  uint64_t rows, live_symbols, words, long_words;
  uint64_t repacks;
  double host_pop_ms, host_launch_ms, host_wait_ms, host_apply_ms; /* merge loop split on the host */
  uint64_t records, heap_pushes, heap_pops, heap_peak;             /* host replica counters */
  uint64_t collectives, exchange_bytes;                            /* multi-GPU: NCCL all-gathers and their bytes */
  uint64_t resident_spill_merges; /* LOCAL merges of the resident kernel whose deltas overflowed shared memory into the global pair table */
  uint64_t exchange_ns;                                             /* multi-GPU load: host time from the first all-gather to the merged global word table */
  uint64_t reserved_[3];
  /* resident cluster kernel: merges done by the leader cluster alone / by the whole grid, and the device time
   * they took (command seen -> result published, %globaltimer; also added to merge_kernel_ms) */
  uint64_t resident_local_merges, resident_grid_merges;
  double resident_local_ms, resident_grid_ms;
  /* look-ahead of the resident kernel: hints the host sent (HostCore::peek_next), merges the device started
   * from one without waiting for the command, hints it turned down; host time spent looking ahead */
  uint64_t hints_sent, hints_taken, hints_rejected;
  double host_peek_ms;
  /* device time of the tokenise + dedupe kernel(s) of the last load (CUDA events on the launch stream; for a host
   * buffer this spans the pipelined host-to-device copy as well) and the corpus bytes they covered */
  double tokenize_ms;
  uint64_t tokenize_bytes;
  /* LOCAL merges of the resident kernel by the length of the birth log they read (<= 512, <= 4096, <= 32768, more
   * entries): how many, their device time, the records they sent to the host */
  uint64_t local_by_log[4];
  double local_by_log_ms[4];
  uint64_t local_by_log_recs[4];
} SwbStats;
void swb_get_stats(const Trainer *trainer, SwbStats *out);
/* 1: bracket every merge-scan launch with CUDA events (adds a little latency); 0: off (default). */
void swb_set_kernel_timing(Trainer *trainer, int enabled);
/* Profiling aid. After bpe_init on a freshly loaded corpus (one GPU, no long words): runs the first n merges of
 * `merge_triples` -- a merge list this same corpus produced before -- inside ONE launch of the resident merge kernel
 * (merge_cluster) whose commands come from a script in device memory instead of the host mailbox. The launch never
 * waits for the host, so a profiler's kernel replay can capture it and its duration is the device-side floor of the
 * merge loop. The handle is consumed: only swb_get_stats / bpe_trainer_destroy are meaningful afterwards.
 * *kernel_ms receives the launch duration (CUDA events). 0 on success, -1 on error. */
int swb_profile_scripted_merges(Trainer *trainer, const int32_t *merge_triples, size_t n, double *kernel_ms);

/* ---- encoder (no reference entry point exists; semantics: reference base.py:10-36 applied
 * lowest merge rank first, per whitespace-delimited word; delimiters \t \r \n space) ---- */
typedef struct SwbEncoder SwbEncoder;
SwbEncoder *swb_encoder_create(const int32_t *merge_triples, size_t n_merges, const int32_t *byte_map256);
SwbEncoder *swb_encoder_from_trainer(const Trainer *trainer);
void swb_encoder_destroy(SwbEncoder *enc);
/* Encodes host text. Token ids of all words, in order, go to out_ids (capacity cap_ids);
 * if word_ntok != NULL it receives the token count of each word (capacity cap_words).
 * Returns the number of tokens, or -1 on error (then see swb_last_error; capacity too small
 * is an error: nothing partial is returned). *n_words receives the number of words. */
int64_t swb_encode(SwbEncoder *enc, const void *text, size_t nbytes, int32_t *out_ids, size_t cap_ids,
                   uint32_t *word_ntok, size_t cap_words, size_t *n_words);
/* Same with device pointers (text, out_ids, word_ntok all in device memory). */
int64_t swb_encode_device(SwbEncoder *enc, const void *d_text, size_t nbytes, int32_t *d_out_ids, size_t cap_ids,
                          uint32_t *d_word_ntok, size_t cap_words, size_t *n_words);
/* ids -> bytes on the host (table lookup; not a device path). Returns the byte count. */
size_t swb_decode(const SwbEncoder *enc, const int32_t *ids, size_t n, uint8_t *out, size_t cap);
uint64_t swb_encoder_kernel_launches(const SwbEncoder *enc);

/* ---- optional pre-pass: the reference's normalize_line (reference csrc/bpe/normalize.cpp:24-59; it has
 * no caller there) applied to every line of `text`: ASCII lower-casing, whitespace runs inside a line -> U+2581, runs at
 * the start / end of a line dropped, lines joined by their '\n' again. text / out are host pointers, or device pointers
 * when on_device != 0. Returns the length of the full result (when it exceeds cap only the first cap bytes were
 * written; 3 * nbytes always suffices), or -1 on error. */
int64_t swb_normalize(const void *text, size_t nbytes, void *out, size_t cap, int on_device);

/* ---- optional pre-pass: the reference's regex pre-tokenisation (reference shredword/base.py:38-58,
 * apply_regex: regex.findall of its GPT-4 style split pattern). The result is `text` with one ' ' behind every piece of the
 * pattern and the bytes ' ' \t \n \r inside pieces replaced by 0x1C 0x1D 0x1E 0x1F, so that the whitespace-splitting
 * trainer / encoder treats exactly the reference's pieces as its words; dropping the ' ' bytes and mapping 0x1C-0x1F back
 * restores the text. `text` is UTF-8 (bytes that are not well-formed UTF-8 count as characters outside \p{L} \p{N} \s).
 * Pointers and return value as for swb_normalize; 2 * nbytes always suffices. */
int64_t swb_pretokenize(const void *text, size_t nbytes, void *out, size_t cap, int on_device);

/* ---- building blocks of the multi-GPU merge loop ----
 * Unique words are sharded over ranks (word wi belongs to rank wi % nranks); every rank keeps an
 * identical replica of the pair table + heap. A record is 4 x int64: first, second, delta, key
 * (key = global first-touch order; for the count pass delta is the weighted frequency).
 * The swb_dist_* host functions need no GPU; the swb_shard_* functions run this rank's kernels. */
int swb_set_shard(Trainer *trainer, int rank, int nranks);     /* before loading the corpus */
/* Reduce concatenated per-rank record lists by pair: sum delta, min key. Returns the new count. */
size_t swb_dist_reduce_records(int64_t *recs, size_t n);
/* Replica step 1: reset pair table + heap and seed them from the reduced count records. */
void swb_dist_seed(Trainer *trainer, const int64_t *recs, size_t n);
/* Replica step 2: pop until a valid pair; 1 = (a, b, new_id) filled and recorded, 0 = heap exhausted. */
int swb_dist_next_merge(Trainer *trainer, int32_t *a, int32_t *b, int32_t *new_id);
/* Replica step 3: apply the reduced delta records of the merge returned by step 2. */
void swb_dist_apply(Trainer *trainer, const int64_t *recs, size_t n);
/* Look-ahead between steps 2 and 3 (what the resident merge kernel's hints are made of): 1 = the pair step 2
 * will return NEXT is (a, b), currently at frequency *freq, PROVIDED the pending merge pushes no entry with a
 * frequency >= *freq and leaves (a, b)'s frequency unchanged; 0 = no statement. The heap (reference
 * heap.cpp:53-114 array) is left bit for bit as it was. */
int swb_dist_peek_next(Trainer *trainer, int32_t *a, int32_t *b, uint64_t *freq);
/* The same for the next `want` pairs: out[3*i .. 3*i+2] = {a, b, freq} of the pair that step 2 returns i calls
 * after the next one, PROVIDED no merge from the pending one up to the one before it pushes an entry with a
 * frequency >= that freq, and the pair's frequency is still that freq when its turn comes. Returns the
 * number of entries filled (0 .. want). */
size_t swb_dist_peek_list(Trainer *trainer, int64_t *out, size_t want);
/* In-library multi-GPU training (one process per GPU): NCCL is loaded with dlopen and the per-merge
 * all-gather is issued from C++ on the handle's stream. Rank 0 calls swb_dist_unique_id, the caller
 * ships the 128 bytes to every rank (any transport), every rank calls swb_dist_init before loading the
 * corpus; afterwards bpe_init / bpe_merge_batch / bpe_train run the sharded loop (every rank must make
 * the same calls). 0 on success, -1 on error. */
int swb_dist_unique_id(void *out128);
int swb_dist_init(Trainer *trainer, int rank, int nranks, const void *unique_id128);
/* How the merge loop runs on more than one GPU (call after swb_dist_init, before loading):
 *   sharded == 0 (default): REPLICATED. The load is range-split and the unique-word tables are exchanged
 *     (swb_load_corpus_shard: the part of the path that scales with the corpus), then every rank holds all
 *     unique words and runs the single-GPU merge loop; no collective per merge. The merge loop touches only
 *     the words that hold the pair and is latency-bound, so splitting it buys nothing and costs one NCCL
 *     exchange per merge.
 *   sharded != 0: word wi lives on rank wi % nranks, every merge all-gathers the delta records (NCCL) and
 *     every rank applies them to its replica of the pair table + heap. */
int swb_dist_set_sharded(Trainer *trainer, int sharded);
/* Range-split load: `data` (host pointer, or device pointer when on_device != 0) is only THIS rank's
 * byte range of the corpus, starting at byte `global_offset` of the whole; ranges must be cut on
 * delimiters. Every rank tokenises its range, the unique-word tables are exchanged over NCCL and merged,
 * so all ranks end with the same global word table in reference order. Needs swb_dist_init. */
int swb_load_corpus_shard(Trainer *trainer, const void *data, size_t nbytes, uint64_t global_offset, int on_device);
/* The NCCL communicator is process-wide and shared by all handles with the same (rank, nranks): when
 * swb_dist_has_comm() is 1, swb_dist_init may be called with unique_id128 == NULL (no id exchange needed).
 * swb_dist_shutdown destroys it (call once, after the last handle is gone). */
int swb_dist_has_comm(int rank, int nranks);
void swb_dist_shutdown(void);
/* This rank's kernels: local pair count / local merge of (a,b)->new_id. Return the record count
 * (-1 on error); records are written to recs (capacity cap records). */
int64_t swb_shard_count(Trainer *trainer, int64_t *recs, size_t cap);
int64_t swb_shard_merge(Trainer *trainer, int32_t a, int32_t b, int32_t new_id, int64_t *recs, size_t cap);
/* bpe_save with a caller-supplied token histogram (the all-reduced one when words are sharded).
 * freq must hold 256+num_merges entries. 0 on success. */
int swb_save_with_freq(const Trainer *trainer, const char *model_path, const char *vocab_path, const uint64_t *freq,
                       size_t n_freq);

#ifdef __cplusplus
}
#endif
#endif /* SHREDWORD_B200_H */
