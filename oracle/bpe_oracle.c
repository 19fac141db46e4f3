/* bpe_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE. See bpe_oracle.h for the rules.
 *
 * A CPU restatement of the reference's BPE trainer that reproduces its *observable
 * orders* (word order, pair-table iteration order, heap tie-breaks, delta-map order)
 * with flat arrays instead of linked lists and fixed-bucket chained maps, so that it is
 * fast enough to check 100 MB - 1 GB corpora. Every function cites the reference code
 * whose behaviour it restates (paths relative to /root/reference/shredword/).
 *
 * Semantics restated:
 *  - Symbol.deleted is treated as always false (it is uninitialised in
 *    the reference and functionally dead under a zero-filling malloc).
 *  - A NUL byte in the corpus is rejected (-1): the reference's fgets/strlen loop
 *    (csrc/bpe/bpe.cpp:230-246) drops a libc-buffer-dependent span after a NUL.
 */
#define _GNU_SOURCE
#include "bpe_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define INITIAL_VOCAB 256       /* csrc/bpe/bpe.h:20 */
#define WORD_BUCKETS 4096       /* csrc/bpe/bpe.h:21, bpe.cpp:214 */
#define PAIR_BUCKETS 4096       /* csrc/bpe/bpe.h:19, bpe.cpp:179,295 */
#define DELTA_BUCKETS 1024      /* csrc/bpe/bpe.cpp:17 */
#define DEFAULT_MIN_PAIR_FREQ 2000 /* csrc/bpe/bpe.h:23 */

/* ------------------------------------------------------------------ small utilities */
static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static void *xcalloc(size_t n, size_t s) {
  void *p = calloc(n ? n : 1, s ? s : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static void *xrealloc(void *q, size_t n) {
  void *p = realloc(q, n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static int is_delim(uint8_t c) { /* csrc/bpe/bpe.cpp:247 strtok(line, "\t\r\n ") */
  return c == '\t' || c == '\r' || c == '\n' || c == ' ';
}
static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

/* ------------------------------------------------------------------ pair table
 * Restates BIMap (csrc/bpe/hash.cpp:104-130): pair -> {freq, version}, entries created on
 * first lookup with freq 0 / version 0. Entries are kept in creation order because the
 * reference's iteration order (bucket ascending, chain = creation order) is observable. */
typedef struct { int32_t first, second; uint64_t freq; uint32_t version; } PairInfo;
typedef struct { PairInfo *e; size_t n, cap; uint32_t *slots; size_t nslots; } PairMap;

static void pm_init(PairMap *m) {
  m->n = 0; m->cap = 1024; m->e = xmalloc(m->cap * sizeof(PairInfo));
  m->nslots = 4096; m->slots = xcalloc(m->nslots, sizeof(uint32_t));
}
static void pm_free(PairMap *m) { free(m->e); free(m->slots); m->e = NULL; m->slots = NULL; m->n = 0; }
static size_t pm_hash(int32_t a, int32_t b) {
  return (size_t)mix64(((uint64_t)(uint32_t)a << 32) | (uint32_t)b);
}
static void pm_grow(PairMap *m) {
  free(m->slots);
  m->nslots *= 4;
  m->slots = xcalloc(m->nslots, sizeof(uint32_t));
  for (size_t i = 0; i < m->n; i++) {
    size_t h = pm_hash(m->e[i].first, m->e[i].second) & (m->nslots - 1);
    while (m->slots[h]) h = (h + 1) & (m->nslots - 1);
    m->slots[h] = (uint32_t)(i + 1);
  }
}
static PairInfo *pm_get(PairMap *m, int32_t a, int32_t b) {
  size_t h = pm_hash(a, b) & (m->nslots - 1);
  while (m->slots[h]) {
    PairInfo *p = &m->e[m->slots[h] - 1];
    if (p->first == a && p->second == b) return p;
    h = (h + 1) & (m->nslots - 1);
  }
  if (m->n == m->cap) { m->cap *= 2; m->e = xrealloc(m->e, m->cap * sizeof(PairInfo)); }
  PairInfo *p = &m->e[m->n++];
  p->first = a; p->second = b; p->freq = 0; p->version = 0;
  m->slots[h] = (uint32_t)m->n;
  if (m->n * 2 > m->nslots) { pm_grow(m); return &m->e[m->n - 1]; }
  return p;
}
/* csrc/bpe/hash.cpp:7-16: FNV-1a over the 8 bytes of {first, second} (little endian) */
static uint32_t ref_hash_pair(int32_t first, int32_t second) {
  uint32_t h = 2166136261u;
  uint32_t w[2] = {(uint32_t)first, (uint32_t)second};
  for (int k = 0; k < 2; k++)
    for (int i = 0; i < 4; i++) { h ^= (w[k] >> (8 * i)) & 0xffu; h *= 16777619u; }
  return h;
}

/* ------------------------------------------------------------------ heap
 * Restates csrc/bpe/heap.cpp:53-114 exactly (array binary heap ordered by freq only). */
typedef struct { int32_t first, second; uint64_t freq; uint32_t version; } HeapEnt;
typedef struct { HeapEnt *d; size_t n, cap; } Heap;

static void heap_reset(Heap *h) { h->n = 0; }
static void heap_push(Heap *h, int32_t a, int32_t b, uint64_t freq, uint32_t version) {
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 4096; h->d = xrealloc(h->d, h->cap * sizeof(HeapEnt)); }
  size_t i = h->n++;
  h->d[i].first = a; h->d[i].second = b; h->d[i].freq = freq; h->d[i].version = version;
  while (i > 0) { /* heap.cpp:74-79: stop when parent.freq >= child.freq */
    size_t p = (i - 1) >> 1;
    if (h->d[p].freq >= h->d[i].freq) break;
    HeapEnt tmp = h->d[p]; h->d[p] = h->d[i]; h->d[i] = tmp;
    i = p;
  }
}
static HeapEnt heap_pop(Heap *h) {
  HeapEnt top = h->d[0];
  h->d[0] = h->d[--h->n];
  size_t i = 0;
  for (;;) { /* heap.cpp:97-111: left if strictly greater, then right if strictly greater than best */
    size_t l = 2 * i + 1, r = l + 1, best = i;
    if (l < h->n && h->d[l].freq > h->d[best].freq) best = l;
    if (r < h->n && h->d[r].freq > h->d[best].freq) best = r;
    if (best == i) break;
    HeapEnt tmp = h->d[i]; h->d[i] = h->d[best]; h->d[best] = tmp;
    i = best;
  }
  return top;
}

/* ------------------------------------------------------------------ trainer state */
struct OracleTrainer {
  size_t target; int32_t unk; float cov; uint64_t minf;
  size_t W;            /* unique words, in reference order (wi) */
  uint8_t *wbytes; uint64_t *wboff; /* word bytes, [W+1] offsets */
  int32_t *syms; uint64_t *soff;    /* fixed slots: word wi owns syms[soff[wi] .. soff[wi+1]) */
  uint32_t *slen;                   /* live symbols of word wi (compacted to the slot's front) */
  uint64_t *cnt;
  uint8_t keep[256];
  PairMap pm; Heap heap;
  int32_t *merges; size_t nm, mcap; /* (a, b) pairs; new id = 256 + index */
};

OracleTrainer *oracle_create(size_t target_vocab_size, int32_t unk_id, float cov, uint64_t min_pair_freq) {
  OracleTrainer *t = xcalloc(1, sizeof *t);
  t->target = target_vocab_size; t->unk = unk_id;
  /* bpe.cpp:124-130 (comparison is done in double against a float field, same here) */
  if (cov <= 0.0 || cov >= 1.0) cov = 0.995;
  t->cov = cov;
  t->minf = min_pair_freq ? min_pair_freq : DEFAULT_MIN_PAIR_FREQ;
  pm_init(&t->pm);
  return t;
}
static void free_corpus(OracleTrainer *t) {
  free(t->wbytes); free(t->wboff); free(t->syms); free(t->soff); free(t->slen); free(t->cnt);
  t->wbytes = NULL; t->wboff = NULL; t->syms = NULL; t->soff = NULL; t->slen = NULL; t->cnt = NULL; t->W = 0;
}
void oracle_destroy(OracleTrainer *t) {
  if (!t) return;
  free_corpus(t); pm_free(&t->pm); free(t->heap.d); free(t->merges); free(t);
}

/* Unique words in first-occurrence order (bytes at data + uoff[u]) -> the word table in reference order, kept bytes, symbols.
 * Frees uoff / ulen / ucnt. */
static int build_table(OracleTrainer *t, const uint8_t *data, size_t un, uint64_t *uoff, uint32_t *ulen, uint64_t *ucnt) {
  /* pass 2: stable counting sort by djb2 bucket (hash.cpp:35-39; only the low 12 bits matter) */
  uint32_t *bucket = xmalloc(un * sizeof *bucket);
  size_t start[WORD_BUCKETS + 1]; memset(start, 0, sizeof start);
  for (size_t u = 0; u < un; u++) {
    size_t h = 5381;
    for (uint32_t k = 0; k < ulen[u]; k++) h = ((h << 5) + h) + data[uoff[u] + k];
    bucket[u] = (uint32_t)(h & (WORD_BUCKETS - 1));
    start[bucket[u] + 1]++;
  }
  for (size_t b = 0; b < WORD_BUCKETS; b++) start[b + 1] += start[b];
  size_t *order = xmalloc(un * sizeof *order);
  for (size_t u = 0; u < un; u++) order[start[bucket[u]]++] = u;
  free(bucket);
  /* character histogram over UNIQUE words, each once (histogram.cpp:30-36, bpe.cpp:257-259) */
  uint64_t ch[256]; memset(ch, 0, sizeof ch);
  for (size_t u = 0; u < un; u++)
    for (uint32_t k = 0; k < ulen[u]; k++) ch[data[uoff[u] + k]]++;
  /* char_map is a 256-bucket StrMap of 1-byte strings: bucket = (5381*33 + b) & 255 = (165 + b) & 255,
   * one entry per bucket, iterated ascending (bpe.cpp:267-268); then qsort descending by count
   * (bpe.cpp:270, histogram.cpp:47-53) -- glibc's qsort is a stable merge sort here, so ties keep
   * the bucket order. A stable insertion sort restates that. */
  uint8_t sym[256]; size_t c = 0;
  for (int b = 0; b < 256; b++) { uint8_t byte = (uint8_t)((b - 165) & 255); if (ch[byte]) sym[c++] = byte; }
  for (size_t a = 1; a < c; a++) {
    uint8_t v = sym[a]; size_t j = a;
    while (j > 0 && ch[sym[j - 1]] < ch[v]) { sym[j] = sym[j - 1]; j--; }
    sym[j] = v;
  }
  size_t keep = (size_t)((float)c * t->cov); /* bpe.cpp:274: size_t * float -> float -> size_t */
  memset(t->keep, 0, sizeof t->keep);
  for (size_t a = 0; a < keep && a < c; a++) t->keep[sym[a]] = 1;
  /* build the word table in reference order (bpe.cpp:283-292, histogram.cpp:7-27) */
  t->W = un;
  t->wboff = xmalloc((un + 1) * sizeof *t->wboff); t->soff = xmalloc((un + 1) * sizeof *t->soff);
  t->slen = xmalloc(un * sizeof *t->slen); t->cnt = xmalloc(un * sizeof *t->cnt);
  uint64_t total = 0;
  for (size_t w = 0; w < un; w++) { t->wboff[w] = total; t->soff[w] = total; total += ulen[order[w]]; }
  t->wboff[un] = total; t->soff[un] = total;
  t->wbytes = xmalloc(total); t->syms = xmalloc(total * sizeof *t->syms);
  for (size_t w = 0; w < un; w++) {
    size_t u = order[w];
    memcpy(t->wbytes + t->wboff[w], data + uoff[u], ulen[u]);
    for (uint32_t k = 0; k < ulen[u]; k++) {
      uint8_t b = data[uoff[u] + k];
      t->syms[t->soff[w] + k] = t->keep[b] ? (int32_t)b : t->unk; /* histogram.cpp:15 */
    }
    t->slen[w] = ulen[u]; t->cnt[w] = ucnt[u];
  }
  free(order); free(uoff); free(ulen); free(ucnt);
  /* bpe.cpp:295: fresh pair table */
  pm_free(&t->pm); pm_init(&t->pm);
  return 0;
}

/* ------------------------------------------------------------------ load (T1-T3)
 * csrc/bpe/bpe.cpp:229-252 + hash.cpp:29-53: whitespace split, word -> count.
 * Word order = StrMap iteration order (hash.cpp:67-71): bucket = djb2(word) & 4095 ascending,
 * within a bucket in order of first occurrence (new entries are appended at the chain tail). */
int oracle_load_corpus_buffer(OracleTrainer *t, const uint8_t *data, size_t n) {
  if (!t || (!data && n)) return -1;
  if (memchr(data, 0, n)) { fprintf(stderr, "oracle: NUL byte in corpus is outside the parity domain\n"); return -1; }
  free_corpus(t);
  /* pass 1: unique words in first-occurrence order */
  size_t ucap = 1 << 16, un = 0;
  uint64_t *uoff = xmalloc(ucap * sizeof *uoff);  /* offset of the first occurrence */
  uint32_t *ulen = xmalloc(ucap * sizeof *ulen);
  uint64_t *ucnt = xmalloc(ucap * sizeof *ucnt);
  uint64_t *uhash = xmalloc(ucap * sizeof *uhash);
  size_t nslots = 1 << 18; uint32_t *slots = xcalloc(nslots, sizeof *slots);
  size_t i = 0;
  while (i < n) {
    while (i < n && is_delim(data[i])) i++;
    if (i >= n) break;
    size_t s = i; uint64_t h = 1469598103934665603ULL;
    while (i < n && !is_delim(data[i])) { h = (h ^ data[i]) * 1099511628211ULL; i++; }
    size_t len = i - s;
    if (len > 0xffffffffu) { fprintf(stderr, "oracle: word longer than 4 GiB\n"); return -1; }
    h = mix64(h);
    size_t p = (size_t)h & (nslots - 1);
    for (;;) {
      uint32_t v = slots[p];
      if (!v) break;
      size_t u = v - 1;
      if (uhash[u] == h && ulen[u] == len && memcmp(data + uoff[u], data + s, len) == 0) { ucnt[u]++; goto next_word; }
      p = (p + 1) & (nslots - 1);
    }
    if (un == ucap) {
      ucap *= 2;
      uoff = xrealloc(uoff, ucap * sizeof *uoff); ulen = xrealloc(ulen, ucap * sizeof *ulen);
      ucnt = xrealloc(ucnt, ucap * sizeof *ucnt); uhash = xrealloc(uhash, ucap * sizeof *uhash);
    }
    uoff[un] = s; ulen[un] = (uint32_t)len; ucnt[un] = 1; uhash[un] = h; slots[p] = (uint32_t)(++un);
    if (un * 2 > nslots) {
      free(slots); nslots *= 4; slots = xcalloc(nslots, sizeof *slots);
      for (size_t u = 0; u < un; u++) {
        size_t q = (size_t)uhash[u] & (nslots - 1);
        while (slots[q]) q = (q + 1) & (nslots - 1);
        slots[q] = (uint32_t)(u + 1);
      }
    }
  next_word:;
  }
  free(slots); free(uhash);
  return build_table(t, data, un, uoff, ulen, ucnt);
}

int oracle_load_corpus(OracleTrainer *t, const char *path) {
  if (!t || !path) return -1;
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return -1; }
  long sz = ftell(f);
  if (sz < 0) { fclose(f); return -1; }
  rewind(f);
  uint8_t *buf = xmalloc((size_t)sz);
  size_t got = fread(buf, 1, (size_t)sz, f);
  fclose(f);
  int rc = got == (size_t)sz ? oracle_load_corpus_buffer(t, buf, got) : -1;
  free(buf);
  return rc;
}

/* ---- streaming load (test infrastructure for corpora larger than host memory: the 50 GB configuration) ----
 * The same pass 1 as oracle_load_corpus_buffer, fed chunk by chunk: a chunk must end on a delimiter (no word straddles two
 * chunks). The bytes of a word are copied into an arena when it is first seen, so the chunk can be dropped afterwards. */
struct OracleStream {
  uint8_t *arena; size_t an, acap;
  uint64_t *uoff; uint32_t *ulen; uint64_t *ucnt; uint64_t *uhash; size_t un, ucap;
  uint32_t *slots; size_t nslots;
  int bad;
};
OracleStream *oracle_stream_begin(void) {
  OracleStream *s = xcalloc(1, sizeof *s);
  s->acap = 1 << 20; s->arena = xmalloc(s->acap);
  s->ucap = 1 << 16;
  s->uoff = xmalloc(s->ucap * sizeof *s->uoff); s->ulen = xmalloc(s->ucap * sizeof *s->ulen);
  s->ucnt = xmalloc(s->ucap * sizeof *s->ucnt); s->uhash = xmalloc(s->ucap * sizeof *s->uhash);
  s->nslots = 1 << 18; s->slots = xcalloc(s->nslots, sizeof *s->slots);
  return s;
}
int oracle_stream_feed(OracleStream *s, const uint8_t *data, size_t n) {
  if (!s || (!data && n)) return -1;
  if (n && memchr(data, 0, n)) { s->bad = 1; return -1; }
  size_t i = 0;
  while (i < n) {
    while (i < n && is_delim(data[i])) i++;
    if (i >= n) break;
    size_t st = i; uint64_t h = 1469598103934665603ULL;
    while (i < n && !is_delim(data[i])) { h = (h ^ data[i]) * 1099511628211ULL; i++; }
    size_t len = i - st;
    h = mix64(h);
    size_t p = (size_t)h & (s->nslots - 1);
    int found = 0;
    for (;;) {
      uint32_t v = s->slots[p];
      if (!v) break;
      size_t u = v - 1;
      if (s->uhash[u] == h && s->ulen[u] == len && memcmp(s->arena + s->uoff[u], data + st, len) == 0) { s->ucnt[u]++; found = 1; break; }
      p = (p + 1) & (s->nslots - 1);
    }
    if (found) continue;
    if (s->un == s->ucap) {
      s->ucap *= 2;
      s->uoff = xrealloc(s->uoff, s->ucap * sizeof *s->uoff); s->ulen = xrealloc(s->ulen, s->ucap * sizeof *s->ulen);
      s->ucnt = xrealloc(s->ucnt, s->ucap * sizeof *s->ucnt); s->uhash = xrealloc(s->uhash, s->ucap * sizeof *s->uhash);
    }
    while (s->an + len > s->acap) { s->acap *= 2; s->arena = xrealloc(s->arena, s->acap); }
    memcpy(s->arena + s->an, data + st, len);
    s->uoff[s->un] = s->an; s->ulen[s->un] = (uint32_t)len; s->ucnt[s->un] = 1; s->uhash[s->un] = h; s->an += len;
    s->slots[p] = (uint32_t)(++s->un);
    if (s->un * 2 > s->nslots) {
      free(s->slots); s->nslots *= 4; s->slots = xcalloc(s->nslots, sizeof *s->slots);
      for (size_t u = 0; u < s->un; u++) {
        size_t q = (size_t)s->uhash[u] & (s->nslots - 1);
        while (s->slots[q]) q = (q + 1) & (s->nslots - 1);
        s->slots[q] = (uint32_t)(u + 1);
      }
    }
  }
  return 0;
}
int oracle_stream_finish(OracleStream *s, OracleTrainer *t) {
  if (!s || !t) return -1;
  int rc = -1;
  free(s->slots); free(s->uhash);
  if (!s->bad) {
    free_corpus(t);
    rc = build_table(t, s->arena, s->un, s->uoff, s->ulen, s->ucnt);  /* (frees uoff / ulen / ucnt) */
  } else { free(s->uoff); free(s->ulen); free(s->ucnt); }
  free(s->arena); free(s);
  return rc;
}

/* ------------------------------------------------------------------ count (T4)
 * csrc/bpe/bpe.cpp:315-370. key orders first touches: (wi, position). */
typedef struct { int32_t first, second; int64_t delta; uint64_t key; } Rec;
typedef struct { Rec *e; size_t n, cap; uint32_t *slots; size_t nslots; } RecMap;

static void rm_init(RecMap *m) {
  m->n = 0; m->cap = 256; m->e = xmalloc(m->cap * sizeof(Rec));
  m->nslots = 1024; m->slots = xcalloc(m->nslots, sizeof(uint32_t));
}
static void rm_free(RecMap *m) { free(m->e); free(m->slots); }
static void rm_add(RecMap *m, int32_t a, int32_t b, int64_t delta, uint64_t key) {
  size_t h = pm_hash(a, b) & (m->nslots - 1);
  while (m->slots[h]) {
    Rec *r = &m->e[m->slots[h] - 1];
    if (r->first == a && r->second == b) { r->delta += delta; if (key < r->key) r->key = key; return; }
    h = (h + 1) & (m->nslots - 1);
  }
  if (m->n == m->cap) { m->cap *= 2; m->e = xrealloc(m->e, m->cap * sizeof(Rec)); }
  Rec *r = &m->e[m->n++];
  r->first = a; r->second = b; r->delta = delta; r->key = key;
  m->slots[h] = (uint32_t)m->n;
  if (m->n * 2 > m->nslots) {
    free(m->slots); m->nslots *= 4; m->slots = xcalloc(m->nslots, sizeof(uint32_t));
    for (size_t i = 0; i < m->n; i++) {
      size_t q = pm_hash(m->e[i].first, m->e[i].second) & (m->nslots - 1);
      while (m->slots[q]) q = (q + 1) & (m->nslots - 1);
      m->slots[q] = (uint32_t)(i + 1);
    }
  }
}
#define KEY(wi, pos, slot) (((uint64_t)(wi) << 30) | ((uint64_t)(pos) << 2) | (uint64_t)(slot))

static void count_shard(const OracleTrainer *t, int rank, int nranks, RecMap *out) {
  for (size_t wi = (size_t)rank; wi < t->W; wi += (size_t)nranks) {
    const int32_t *s = t->syms + t->soff[wi];
    uint32_t L = t->slen[wi];
    for (uint32_t k = 0; k + 1 < L; k++) {
      if (s[k] == t->unk || s[k + 1] == t->unk) continue; /* bpe.cpp:333-337 */
      rm_add(out, s[k], s[k + 1], (int64_t)t->cnt[wi], KEY(wi, k, 0));
    }
  }
}

static int rec_cmp_key(const void *x, const void *y) {
  const Rec *a = x, *b = y;
  return a->key < b->key ? -1 : a->key > b->key;
}

void oracle_count_bigrams(OracleTrainer *t) {
  if (!t) return;
  /* first pass (bpe.cpp:329-350): entries are created in first-touch order */
  RecMap rm; rm_init(&rm);
  count_shard(t, 0, 1, &rm);
  qsort(rm.e, rm.n, sizeof(Rec), rec_cmp_key); /* keys are unique per pair -> total order */
  for (size_t i = 0; i < rm.n; i++) {
    PairInfo *p = pm_get(&t->pm, rm.e[i].first, rm.e[i].second);
    if (p->freq == 0) p->version = 0; /* bpe.cpp:342-345 */
    p->freq += (uint64_t)rm.e[i].delta;
  }
  rm_free(&rm);
  /* second pass (bpe.cpp:359-366): bucket ascending, chain in creation order, ALL entries of the map */
  size_t start[PAIR_BUCKETS + 1]; memset(start, 0, sizeof start);
  uint32_t *bk = xmalloc(t->pm.n * sizeof *bk);
  for (size_t i = 0; i < t->pm.n; i++) { bk[i] = ref_hash_pair(t->pm.e[i].first, t->pm.e[i].second) & (PAIR_BUCKETS - 1); start[bk[i] + 1]++; }
  for (size_t b = 0; b < PAIR_BUCKETS; b++) start[b + 1] += start[b];
  size_t *order = xmalloc(t->pm.n * sizeof *order);
  for (size_t i = 0; i < t->pm.n; i++) order[start[bk[i]]++] = i;
  for (size_t k = 0; k < t->pm.n; k++) {
    PairInfo *p = &t->pm.e[order[k]];
    if (p->freq >= t->minf) heap_push(&t->heap, p->first, p->second, p->freq, p->version);
  }
  free(order); free(bk);
}

void oracle_init(OracleTrainer *t) { /* bpe.cpp:171-185 */
  if (!t) return;
  pm_free(&t->pm); pm_init(&t->pm);
  heap_reset(&t->heap);
  oracle_count_bigrams(t);
}

/* ------------------------------------------------------------------ merge (T6, T7)
 * csrc/bpe/bpe.cpp:437-483 for the words of one shard. Emits the FreqChangeMap content
 * (bpe.cpp:10-46) as records keyed by the reference's 64-bit pair_hash, in first-touch order. */
static void merge_shard(OracleTrainer *t, int rank, int nranks, int32_t a, int32_t b, int32_t new_id, RecMap *fc) {
  for (size_t wi = (size_t)rank; wi < t->W; wi += (size_t)nranks) {
    int32_t *s = t->syms + t->soff[wi];
    uint32_t L = t->slen[wi];
    int64_t wc = (int64_t)t->cnt[wi];
    uint32_t r = 0, w = 0; /* read / write cursors: s[0..w) is the rewritten prefix */
    while (r < L) {
      if (r + 1 < L && s[r] == a && s[r + 1] == b) {
        if (w > 0) { /* left neighbour = already rewritten symbol (bpe.cpp:453-460) */
          int32_t Lft = s[w - 1];
          uint64_t oh = ((uint64_t)Lft << 32) | (uint64_t)a;       /* same C expressions as bpe.cpp:456-457 */
          uint64_t nh = ((uint64_t)Lft << 32) | (uint64_t)new_id;
          rm_add(fc, (int32_t)(oh >> 32), (int32_t)(oh & 0xFFFFFFFF), -wc, KEY(wi, r, 0));
          rm_add(fc, (int32_t)(nh >> 32), (int32_t)(nh & 0xFFFFFFFF), wc, KEY(wi, r, 1));
        }
        if (r + 2 < L) { /* right neighbour = not yet rewritten symbol (bpe.cpp:463-470) */
          int32_t R = s[r + 2];
          uint64_t oh = ((uint64_t)b << 32) | (uint64_t)R;
          uint64_t nh = ((uint64_t)new_id << 32) | (uint64_t)R;
          rm_add(fc, (int32_t)(oh >> 32), (int32_t)(oh & 0xFFFFFFFF), -wc, KEY(wi, r, 2));
          rm_add(fc, (int32_t)(nh >> 32), (int32_t)(nh & 0xFFFFFFFF), wc, KEY(wi, r, 3));
        }
        s[w++] = new_id; r += 2; /* bpe.cpp:473-479; the merged symbol cannot match again */
      } else {
        s[w++] = s[r++];
      }
    }
    t->slen[wi] = w;
  }
}

/* csrc/bpe/bpe.cpp:486-517: iterate delta buckets 0..1023 (bucket = pair_hash % 1024), each chain
 * from its head; entries are prepended (bpe.cpp:41-45), so a chain lists the most recently
 * first-touched pair first. recs must be in first-touch order. */
static void apply_deltas(OracleTrainer *t, const Rec *recs, size_t n, int32_t a, int32_t b) {
  size_t start[DELTA_BUCKETS + 1]; memset(start, 0, sizeof start);
  for (size_t i = 0; i < n; i++) {
    uint64_t ph = ((uint64_t)(uint32_t)recs[i].first << 32) | (uint32_t)recs[i].second;
    start[(ph % DELTA_BUCKETS) + 1]++;
  }
  for (size_t k = 0; k < DELTA_BUCKETS; k++) start[k + 1] += start[k];
  size_t *order = xmalloc(n * sizeof *order);
  for (size_t i = n; i-- > 0;) {
    uint64_t ph = ((uint64_t)(uint32_t)recs[i].first << 32) | (uint32_t)recs[i].second;
    order[start[ph % DELTA_BUCKETS]++] = i;
  }
  for (size_t k = 0; k < n; k++) {
    const Rec *r = &recs[order[k]];
    if (r->first == a && r->second == b) continue; /* bpe.cpp:494-496 */
    PairInfo *p = pm_get(&t->pm, r->first, r->second);
    if (r->delta < 0) { /* bpe.cpp:500-509 */
      uint64_t ad = (uint64_t)(-r->delta);
      p->freq = p->freq >= ad ? p->freq - ad : 0;
    } else {
      p->freq += (uint64_t)r->delta;
    }
    if (p->freq >= t->minf) { p->version++; heap_push(&t->heap, r->first, r->second, p->freq, p->version); } /* :512-515 */
  }
  free(order);
}

int oracle_merge_batch(OracleTrainer *t, int batch_size) {
  if (!t) return -1;
  int done = 0;
  while (done < batch_size && t->heap.n > 0) { /* bpe.cpp:405 */
    HeapEnt top = heap_pop(&t->heap);
    PairInfo *info = pm_get(&t->pm, top.first, top.second);
    if (top.version != info->version) continue;  /* bpe.cpp:412-415 */
    if (info->freq < t->minf) continue;          /* bpe.cpp:418-421 */
    int32_t new_id = (int32_t)(INITIAL_VOCAB + t->nm); /* bpe.cpp:424 */
    if (t->nm == t->mcap) { t->mcap = t->mcap ? t->mcap * 2 : 1024; t->merges = xrealloc(t->merges, t->mcap * 2 * sizeof(int32_t)); }
    t->merges[2 * t->nm] = top.first; t->merges[2 * t->nm + 1] = top.second;
    RecMap fc; rm_init(&fc);
    merge_shard(t, 0, 1, top.first, top.second, new_id, &fc);
    qsort(fc.e, fc.n, sizeof(Rec), rec_cmp_key); /* first-touch order */
    apply_deltas(t, fc.e, fc.n, top.first, top.second);
    rm_free(&fc);
    info = pm_get(&t->pm, top.first, top.second); /* table may have been reallocated */
    info->freq = 0; info->version++;              /* bpe.cpp:523-524 */
    t->nm++; done++;
  }
  return done;
}

int oracle_train(OracleTrainer *t) { /* bpe.cpp:597-655; the batch size there only paces logging */
  if (!t) return -1;
  oracle_init(t);
  int total = 0, target = (int)t->target - INITIAL_VOCAB;
  while (total < target) {
    if (t->heap.n == 0) break;
    int got = oracle_merge_batch(t, target - total);
    if (got <= 0) break;
    total += got;
  }
  return total;
}

/* ------------------------------------------------------------------ save (S1, S2)
 * csrc/bpe/bpe.cpp:678-739. Token strings are NUL-terminated concatenations, so token 0
 * (and any token built from it) contributes an empty piece. */
void oracle_token_freq(const OracleTrainer *t, uint64_t *freq) {
  size_t T = INITIAL_VOCAB + t->nm;
  memset(freq, 0, T * sizeof *freq);
  for (size_t w = 0; w < t->W; w++)
    for (uint32_t k = 0; k < t->slen[w]; k++) {
      int32_t id = t->syms[t->soff[w] + k];
      if (id >= 0 && (size_t)id < T) freq[id] += t->cnt[w]; /* out-of-range ids are UB in the reference */
    }
}

int oracle_save(const OracleTrainer *t, const char *model_path, const char *vocab_path) {
  if (!t || !model_path || !vocab_path) return -1;
  size_t M = t->nm, T = INITIAL_VOCAB + M;
  char **tok = xcalloc(T, sizeof *tok);
  size_t *tl = xcalloc(T, sizeof *tl);
  for (size_t i = 0; i < INITIAL_VOCAB; i++) { tok[i] = xmalloc(2); tok[i][0] = (char)i; tok[i][1] = 0; tl[i] = i ? 1 : 0; }
  for (size_t m = 0; m < M; m++) {
    int32_t A = t->merges[2 * m], B = t->merges[2 * m + 1];
    const char *sa = (A >= 0 && (size_t)A < INITIAL_VOCAB + m) ? tok[A] : "";
    const char *sb = (B >= 0 && (size_t)B < INITIAL_VOCAB + m) ? tok[B] : "";
    size_t la = strlen(sa), lb = strlen(sb);
    tok[INITIAL_VOCAB + m] = xmalloc(la + lb + 1);
    memcpy(tok[INITIAL_VOCAB + m], sa, la); memcpy(tok[INITIAL_VOCAB + m] + la, sb, lb + 1);
    tl[INITIAL_VOCAB + m] = la + lb;
  }
  uint64_t *freq = xcalloc(T, sizeof *freq);
  oracle_token_freq(t, freq);
  FILE *vf = fopen(vocab_path, "w");
  if (!vf) return -1;
  for (size_t i = 0; i < T; i++) fprintf(vf, "%s %llu\n", tok[i], (unsigned long long)freq[i]);
  fclose(vf);
  FILE *mf = fopen(model_path, "wb");
  if (!mf) return -1;
  for (size_t m = 0; m < M; m++) {
    int32_t rec[3] = {t->merges[2 * m], t->merges[2 * m + 1], (int32_t)(INITIAL_VOCAB + m)};
    fwrite(rec, sizeof(int32_t), 3, mf);
  }
  fclose(mf);
  for (size_t i = 0; i < T; i++) free(tok[i]);
  free(tok); free(tl); free(freq);
  return 0;
}

/* ------------------------------------------------------------------ inspection */
size_t oracle_num_merges(const OracleTrainer *t) { return t->nm; }
void oracle_get_merges(const OracleTrainer *t, int32_t *out) {
  for (size_t m = 0; m < t->nm; m++) { out[3 * m] = t->merges[2 * m]; out[3 * m + 1] = t->merges[2 * m + 1]; out[3 * m + 2] = (int32_t)(INITIAL_VOCAB + m); }
}
size_t oracle_num_words(const OracleTrainer *t) { return t->W; }
size_t oracle_num_symbols(const OracleTrainer *t) { size_t s = 0; for (size_t w = 0; w < t->W; w++) s += t->slen[w]; return s; }
size_t oracle_word_bytes_total(const OracleTrainer *t) { return t->W ? (size_t)t->wboff[t->W] : 0; }
void oracle_get_words(const OracleTrainer *t, uint64_t *byte_off, uint8_t *bytes, uint64_t *sym_off, int32_t *syms, uint64_t *counts) {
  uint64_t so = 0;
  for (size_t w = 0; w < t->W; w++) {
    if (byte_off) byte_off[w] = t->wboff[w];
    if (sym_off) sym_off[w] = so;
    if (syms) memcpy(syms + so, t->syms + t->soff[w], t->slen[w] * sizeof(int32_t));
    so += t->slen[w];
    if (counts) counts[w] = t->cnt[w];
  }
  if (byte_off) byte_off[t->W] = t->W ? t->wboff[t->W] : 0;
  if (sym_off) sym_off[t->W] = so;
  if (bytes && t->W) memcpy(bytes, t->wbytes, t->wboff[t->W]);
}
void oracle_get_keep(const OracleTrainer *t, uint8_t *keep256) { memcpy(keep256, t->keep, 256); }
size_t oracle_heap_size(const OracleTrainer *t) { return t->heap.n; }
void oracle_get_heap(const OracleTrainer *t, int32_t *first, int32_t *second, uint64_t *freq, uint32_t *version) {
  for (size_t i = 0; i < t->heap.n; i++) { first[i] = t->heap.d[i].first; second[i] = t->heap.d[i].second; freq[i] = t->heap.d[i].freq; version[i] = t->heap.d[i].version; }
}
size_t oracle_num_pairs(const OracleTrainer *t) { return t->pm.n; }
void oracle_get_pairs(const OracleTrainer *t, int32_t *first, int32_t *second, uint64_t *freq, uint32_t *version) {
  for (size_t i = 0; i < t->pm.n; i++) { first[i] = t->pm.e[i].first; second[i] = t->pm.e[i].second; freq[i] = t->pm.e[i].freq; version[i] = t->pm.e[i].version; }
}

/* ------------------------------------------------------------------ sharded building blocks */
static size_t emit_recs(const RecMap *rm, int64_t *recs, size_t cap) {
  for (size_t i = 0; i < rm->n && i < cap; i++) {
    recs[4 * i] = rm->e[i].first; recs[4 * i + 1] = rm->e[i].second;
    recs[4 * i + 2] = rm->e[i].delta; recs[4 * i + 3] = (int64_t)rm->e[i].key;
  }
  return rm->n;
}
size_t oracle_shard_count(const OracleTrainer *t, int rank, int nranks, int64_t *recs, size_t cap) {
  RecMap rm; rm_init(&rm);
  count_shard(t, rank, nranks, &rm);
  size_t n = emit_recs(&rm, recs, cap);
  rm_free(&rm);
  return n;
}
size_t oracle_shard_merge(OracleTrainer *t, int rank, int nranks, int32_t a, int32_t b, int32_t new_id, int64_t *recs, size_t cap) {
  RecMap rm; rm_init(&rm);
  merge_shard(t, rank, nranks, a, b, new_id, &rm);
  size_t n = emit_recs(&rm, recs, cap);
  rm_free(&rm);
  return n;
}

/* ------------------------------------------------------------------ encode / decode (E1)
 * Definition: base.py:10-20 (get_stats: the adjacent pairs of a sequence) and base.py:22-36
 * (merge: replace every non-overlapping occurrence, left to right). Per whitespace-delimited
 * word, repeatedly apply the applicable merge of lowest rank until none applies. */
typedef struct { uint64_t *key; uint32_t *rank; size_t nslots; } RankMap;
static void rk_build(RankMap *m, const int32_t *merges, size_t M) {
  m->nslots = 16; while (m->nslots < 4 * M + 16) m->nslots *= 2;
  m->key = xmalloc(m->nslots * sizeof *m->key); m->rank = xmalloc(m->nslots * sizeof *m->rank);
  memset(m->key, 0xff, m->nslots * sizeof *m->key);
  for (size_t r = 0; r < M; r++) {
    uint64_t k = ((uint64_t)(uint32_t)merges[3 * r] << 32) | (uint32_t)merges[3 * r + 1];
    size_t h = (size_t)mix64(k) & (m->nslots - 1);
    int dup = 0;
    while (m->key[h] != ~0ULL) { if (m->key[h] == k) { dup = 1; break; } h = (h + 1) & (m->nslots - 1); }
    if (!dup) { m->key[h] = k; m->rank[h] = (uint32_t)r; } /* a repeated pair keeps its lowest rank */
  }
}
static uint32_t rk_get(const RankMap *m, int32_t a, int32_t b) {
  uint64_t k = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
  if (k == ~0ULL) return 0xffffffffu;
  size_t h = (size_t)mix64(k) & (m->nslots - 1);
  while (m->key[h] != ~0ULL) { if (m->key[h] == k) return m->rank[h]; h = (h + 1) & (m->nslots - 1); }
  return 0xffffffffu;
}

size_t oracle_encode(const int32_t *merges, size_t M, const int32_t *byte_map, const uint8_t *text, size_t n,
                     int32_t *out, size_t cap, uint32_t *word_ntok, size_t word_cap, size_t *n_words_out) {
  RankMap rk; rk_build(&rk, merges, M);
  size_t ntok = 0, nwords = 0, bufcap = 64;
  int32_t *ids = xmalloc(bufcap * sizeof *ids);
  size_t i = 0;
  while (i < n) {
    while (i < n && is_delim(text[i])) i++;
    if (i >= n) break;
    size_t s = i;
    while (i < n && !is_delim(text[i])) i++;
    size_t L = i - s;
    if (L > bufcap) { bufcap = L * 2; ids = xrealloc(ids, bufcap * sizeof *ids); }
    for (size_t k = 0; k < L; k++) ids[k] = byte_map[text[s + k]];
    for (;;) {
      uint32_t best = 0xffffffffu;
      for (size_t k = 0; k + 1 < L; k++) { uint32_t r = rk_get(&rk, ids[k], ids[k + 1]); if (r < best) best = r; }
      if (best == 0xffffffffu) break;
      int32_t a = merges[3 * best], b = merges[3 * best + 1], nid = merges[3 * best + 2];
      size_t w = 0, r = 0;
      while (r < L) {
        if (r + 1 < L && ids[r] == a && ids[r + 1] == b) { ids[w++] = nid; r += 2; }
        else ids[w++] = ids[r++];
      }
      L = w;
    }
    for (size_t k = 0; k < L; k++) { if (ntok < cap && out) out[ntok] = ids[k]; ntok++; }
    if (word_ntok && nwords < word_cap) word_ntok[nwords] = (uint32_t)L;
    nwords++;
  }
  free(ids); free(rk.key); free(rk.rank);
  if (n_words_out) *n_words_out = nwords;
  return ntok;
}

size_t oracle_decode(const int32_t *merges, size_t M, const int32_t *ids, size_t n, uint8_t *out, size_t cap) {
  size_t pos = 0, sp_cap = 64;
  int32_t *stack = xmalloc(sp_cap * sizeof *stack);
  for (size_t i = 0; i < n; i++) {
    size_t sp = 0; stack[sp++] = ids[i];
    while (sp) {
      int32_t id = stack[--sp];
      if (id >= INITIAL_VOCAB && (size_t)(id - INITIAL_VOCAB) < M) {
        if (sp + 2 > sp_cap) { sp_cap *= 2; stack = xrealloc(stack, sp_cap * sizeof *stack); }
        stack[sp++] = merges[3 * (id - INITIAL_VOCAB) + 1];
        stack[sp++] = merges[3 * (id - INITIAL_VOCAB)];
      } else if (id >= 0 && id < INITIAL_VOCAB) {
        if (pos < cap && out) out[pos] = (uint8_t)id;
        pos++;
      }
    }
  }
  free(stack);
  return pos;
}

/* ---------------------------------------------------------------- normalize (csrc/bpe/normalize.cpp:24-59) */
static int orc_is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }  /* normalize.cpp:13-15 */

size_t oracle_normalize_line(const uint8_t *line, size_t n, uint8_t *out, size_t cap) {
  size_t o = 0;
  int in_space = 1;  /* normalize.cpp:32: the start of the line counts as a space */
  uint8_t t3[3] = {0, 0, 0};  /* the last three bytes written */
#define ORC_PUT(b) do { if (o < cap) out[o] = (b); o++; t3[0] = t3[1]; t3[1] = t3[2]; t3[2] = (b); } while (0)
  for (size_t i = 0; i < n; i++) {
    const uint8_t c = line[i];
    if (orc_is_ws(c)) {
      if (!in_space) {  /* normalize.cpp:36-45: one marker per run */
        ORC_PUT(0xE2); ORC_PUT(0x96); ORC_PUT(0x81);
        in_space = 1;
      }
    } else {
      const uint8_t lc = (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c;  /* normalize.cpp:47 tolower, C locale */
      ORC_PUT(lc);
      in_space = 0;
    }
  }
#undef ORC_PUT
  /* normalize.cpp:52-55: if the output ENDS WITH the three marker bytes they are removed, once -- whether they came from
   * a trailing whitespace run or were already in the input */
  if (o >= 3 && t3[0] == 0xE2 && t3[1] == 0x96 && t3[2] == 0x81) o -= 3;
  return o;
}

size_t oracle_normalize_text(const uint8_t *text, size_t n, uint8_t *out, size_t cap) {
  size_t o = 0, i = 0;
  while (i < n) {
    size_t e = i;
    while (e < n && text[e] != '\n') e++;
    o += oracle_normalize_line(text + i, e - i, o < cap ? out + o : out, o < cap ? cap - o : 0);
    if (e < n) { if (o < cap) out[o] = '\n'; o++; }
    i = e + 1;
  }
  return o;
}
