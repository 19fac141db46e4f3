/* bpe_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's BPE trainer hot path, plus the
 * rank-ordered encode step the reference only defines through helper functions.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker. The shipped library
 * (shredword_b200/libtrainer.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED for training (merge list, .model and .vocab bytes) against the
 * unmodified reference run under a zero-filling malloc (oracle/_ref, see
 * tests/golden/make_golden.py and tests/test_oracle_vs_reference.py).
 * Encode is "parity unpinned" by reference code (the reference has no encoder,
 * reference shredword/base.py:107-109); it is pinned indirectly: the token histogram of
 * the encoded training corpus must equal the frequency column of the reference's .vocab.
 *
 * Reference files cited below are relative to /root/reference/shredword/.
 */
#ifndef BPE_ORACLE_H
#define BPE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleTrainer OracleTrainer;

/* csrc/bpe/bpe.cpp:112-136 (defaults: coverage outside (0,1) -> 0.995, min_pair_freq 0 -> 2000) */
OracleTrainer *oracle_create(size_t target_vocab_size, int32_t unk_id, float character_coverage,
                             uint64_t min_pair_freq);
void oracle_destroy(OracleTrainer *t);

/* csrc/bpe/bpe.cpp:208-297. Returns 0, or -1 (unreadable file / NUL byte in corpus). */
int oracle_load_corpus(OracleTrainer *t, const char *path);
int oracle_load_corpus_buffer(OracleTrainer *t, const uint8_t *data, size_t n);
/* The same load fed chunk by chunk (chunks end on a delimiter), for corpora larger than host memory. */
typedef struct OracleStream OracleStream;
OracleStream *oracle_stream_begin(void);
int oracle_stream_feed(OracleStream *s, const uint8_t *data, size_t n);
int oracle_stream_finish(OracleStream *s, OracleTrainer *t); /* builds t's word table; frees s */

void oracle_init(OracleTrainer *t);           /* csrc/bpe/bpe.cpp:171-185 */
void oracle_count_bigrams(OracleTrainer *t);  /* csrc/bpe/bpe.cpp:315-370 */
int oracle_merge_batch(OracleTrainer *t, int batch_size); /* csrc/bpe/bpe.cpp:391-535 */
int oracle_train(OracleTrainer *t);           /* csrc/bpe/bpe.cpp:597-655 */
int oracle_save(const OracleTrainer *t, const char *model_path, const char *vocab_path); /* :678-739 */

/* ---- inspection (for parity tests) ---- */
size_t oracle_num_merges(const OracleTrainer *t);
void oracle_get_merges(const OracleTrainer *t, int32_t *out /* [3*num_merges]: a, b, new_id */);
size_t oracle_num_words(const OracleTrainer *t);
size_t oracle_num_symbols(const OracleTrainer *t);      /* live symbols over all words */
size_t oracle_word_bytes_total(const OracleTrainer *t);
/* word table in reference order: byte_off[W+1], bytes[], sym_off[W+1], syms[], counts[W] */
void oracle_get_words(const OracleTrainer *t, uint64_t *byte_off, uint8_t *bytes, uint64_t *sym_off,
                      int32_t *syms, uint64_t *counts);
void oracle_get_keep(const OracleTrainer *t, uint8_t *keep256);
size_t oracle_heap_size(const OracleTrainer *t);
void oracle_get_heap(const OracleTrainer *t, int32_t *first, int32_t *second, uint64_t *freq,
                     uint32_t *version);
size_t oracle_num_pairs(const OracleTrainer *t);
/* pair table in creation order */
void oracle_get_pairs(const OracleTrainer *t, int32_t *first, int32_t *second, uint64_t *freq,
                      uint32_t *version);
/* token histogram of the current segmentation, ids in [0, 256+num_merges) (bpe.cpp:704-712) */
void oracle_token_freq(const OracleTrainer *t, uint64_t *freq);

/* ---- sharded building blocks (CPU stand-in for one rank's kernels in world_size>1 tests) ----
 * A word wi belongs to rank (wi % nranks). Records are 4 x int64: first, second, delta, key;
 * for the count pass `delta` is the weighted frequency. key orders first touches
 * (word index major, then position, then slot). Returns the number of records written. */
size_t oracle_shard_count(const OracleTrainer *t, int rank, int nranks, int64_t *recs, size_t cap);
size_t oracle_shard_merge(OracleTrainer *t, int rank, int nranks, int32_t a, int32_t b, int32_t new_id,
                          int64_t *recs, size_t cap);

/* ---- encode (definition: base.py:10-36 get_stats/merge, applied lowest rank first per word) ----
 * byte_map[256]: byte -> initial id (identity for kept bytes, unk_id for dropped ones).
 * Text is split on \t \r \n and space exactly like csrc/bpe/bpe.cpp:247-251.
 * Writes token ids of all words in order to out (capacity cap) and, if word_ntok != NULL,
 * the token count of each word. Returns the number of tokens (may exceed cap: nothing is
 * written past cap), *n_words_out = number of words. */
size_t oracle_encode(const int32_t *merges /* [3*M] */, size_t M, const int32_t *byte_map,
                     const uint8_t *text, size_t n, int32_t *out, size_t cap, uint32_t *word_ntok,
                     size_t word_cap, size_t *n_words_out);
/* inverse: expands ids to bytes through the merge list (token 0..255 -> one byte). */
size_t oracle_decode(const int32_t *merges, size_t M, const int32_t *ids, size_t n, uint8_t *out,
                     size_t cap);

/* ---- optional pre-pass: csrc/bpe/normalize.cpp:24-59, normalize_line restated ----
 * One line (no NUL inside): ASCII upper case -> lower case (tolower in the C locale), every run of ' ' '\t' '\n' '\r'
 * between two other bytes -> U+2581 (E2 96 81), runs at the start and at the end of the line dropped.
 * Writes at most cap bytes (no terminator) and returns the length the full result has. PINNED: tests/golden/
 * normalize_cases.json holds outputs of the reference's own normalize_line (oracle/_ref) for the same inputs. */
size_t oracle_normalize_line(const uint8_t *line, size_t n, uint8_t *out, size_t cap);
/* A whole text, line by line: the lines (split at '\n') are normalised one by one and joined by '\n' again (a final
 * '\n' is kept). This is how the device pre-pass (swb_normalize) applies the reference function to a corpus. */
size_t oracle_normalize_text(const uint8_t *text, size_t n, uint8_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
