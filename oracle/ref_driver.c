/* ref_driver.c -- TEST INFRASTRUCTURE. Drives the UNMODIFIED reference library
 * (oracle/_ref/libtrainer_ref.so, built by oracle/Makefile from /root/reference) through
 * its own C-ABI (reference shredword/csrc/bpe/bpe.h:62-72) and times each phase.
 *
 * The reference reads Symbol.deleted without ever initialising it
 * (reference histogram.cpp:14-22 vs bpe.cpp:333), so its results depend on heap garbage.
 * Defining malloc() here makes every allocation in the process zero-filled, which is the
 * only configuration in which "the reference's output" is well defined.
 *
 * usage: ref_driver corpus vocab_size min_pair_freq unk_id coverage model_out vocab_out [max_merges]
 *   max_merges < 0 (default): bpe_train(); else bpe_init() + bpe_merge_batch(max_merges).
 * stdout of the library is discarded; one JSON line with the timings goes to stderr.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

extern void *__libc_calloc(size_t, size_t);
void *malloc(size_t n) { return __libc_calloc(1, n ? n : 1); }

/* layout of reference bpe.h:43-48 (24 bytes) */
typedef struct {
  size_t target_vocab_size;
  int32_t unk_id;
  float character_coverage;
  uint64_t min_pair_freq;
} RefConfig;

extern void *create_trainer(const RefConfig *);
extern void bpe_trainer_destroy(void *);
extern int bpe_load_corpus(void *, const char *);
extern void bpe_init(void *);
extern int bpe_merge_batch(void *, int);
extern int bpe_train(void *);
extern void bpe_save(const void *, const char *, const char *);

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
  if (argc < 8) {
    fprintf(stderr, "usage: %s corpus vocab min_freq unk cov model vocab [max_merges]\n", argv[0]);
    return 2;
  }
  RefConfig cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.target_vocab_size = (size_t)strtoull(argv[2], NULL, 10);
  cfg.min_pair_freq = strtoull(argv[3], NULL, 10);
  cfg.unk_id = (int32_t)strtol(argv[4], NULL, 10);
  cfg.character_coverage = strtof(argv[5], NULL);
  long max_merges = argc > 8 ? strtol(argv[8], NULL, 10) : -1;
  int save = strcmp(argv[6], "-") != 0;

  if (!freopen("/dev/null", "w", stdout)) return 3;
  void *t = create_trainer(&cfg);
  double t0 = now();
  if (bpe_load_corpus(t, argv[1]) != 0) {
    fprintf(stderr, "{\"error\": \"load failed\"}\n");
    return 4;
  }
  double t1 = now();
  int merges;
  double t2;
  if (max_merges < 0) {
    t2 = t1; /* bpe_train does init+merges in one call */
    merges = bpe_train(t);
  } else {
    bpe_init(t);
    t2 = now();
    merges = 0;
    while (merges < max_merges) { /* bpe_merge_batch may stop early on stale entries only at heap exhaustion */
      int got = bpe_merge_batch(t, (int)(max_merges - merges));
      if (got <= 0) break;
      merges += got;
    }
  }
  double t3 = now();
  if (save) bpe_save(t, argv[6], argv[7]);
  double t4 = now();
  fprintf(stderr,
          "{\"load_s\": %.6f, \"init_s\": %.6f, \"merge_s\": %.6f, \"save_s\": %.6f, \"merges\": %d}\n",
          t1 - t0, t2 - t1, t3 - t2, t4 - t3, merges);
  /* no bpe_trainer_destroy: the reference leaks its symbols anyway and exit is faster */
  return 0;
}
