"""Seeded synthetic corpora for the BPE trainer/encoder hot path.

The reference ships no corpus; BASELINE.json's configs are all "synthetic Zipfian" text.
This module is the single generator used by tests/, bench.py and the golden-vector script,
so that a (config, seed) pair names exactly one byte string everywhere.

Shape of the text: lines of 12 space-separated words, '\\n'-terminated. Word types are drawn
once per corpus; a type's rank-k probability is proportional to k**-s (Zipf, s = 1.1).
  * alphabet "ascii":  letters a-z, letter i drawn with probability ~ 1/(i+1), length U[1,11].
  * alphabet "multi":  60 % ASCII types + 40 % types over Cyrillic / Greek / CJK code points
                       (2-3 byte UTF-8), for config 5's long low-frequency tail.
"""
from __future__ import annotations

import io
import os
from dataclasses import dataclass

import numpy as np

WORDS_PER_LINE = 12
MAX_LEN = 11


@dataclass(frozen=True)
class CorpusSpec:
  name: str
  nbytes: int
  n_types: int
  alphabet: str = "ascii"
  zipf_s: float = 1.1
  seed: int = 7


# BASELINE.json configs -> generator parameters.
CONFIGS = {
  "config1_10MB": CorpusSpec("config1_10MB", 10 * 1000 * 1000, 200_000, "ascii", 1.1, 7),
  "config2_1GB": CorpusSpec("config2_1GB", 1000 * 1000 * 1000, 2_000_000, "ascii", 1.1, 11),
  "config3_10GB": CorpusSpec("config3_10GB", 10 * 1000 * 1000 * 1000, 5_000_000, "ascii", 1.1, 13),
  "config5_50GB": CorpusSpec("config5_50GB", 50 * 1000 * 1000 * 1000, 20_000_000, "multi", 1.1, 17),
}


def _ascii_types(rng: np.random.Generator, n: int):
  """n word types over a-z; returns (padded uint8 matrix [n, W], lengths [n])."""
  lens = rng.integers(1, MAX_LEN + 1, size=n)
  p = 1.0 / np.arange(1, 27)
  p /= p.sum()
  letters = rng.choice(26, size=(n, MAX_LEN), p=p).astype(np.uint8) + ord("a")
  return letters, lens.astype(np.int64)


_MULTI_RANGES = [
  (0x0430, 0x044F),  # Cyrillic small letters (2-byte UTF-8)
  (0x03B1, 0x03C9),  # Greek small letters (2-byte)
  (0x4E00, 0x4FFF),  # CJK unified ideographs, first 512 (3-byte)
]


def _multi_types(rng: np.random.Generator, n: int):
  """n word types over non-ASCII scripts; each type is 1..5 code points of one script."""
  width = 15
  mat = np.zeros((n, width), dtype=np.uint8)
  lens = np.zeros(n, dtype=np.int64)
  script = rng.integers(0, len(_MULTI_RANGES), size=n)
  ncp = rng.integers(1, 6, size=n)
  for s, (lo, hi) in enumerate(_MULTI_RANGES):
    rows = np.nonzero(script == s)[0]
    if rows.size == 0:
      continue
    span = hi - lo + 1
    pr = 1.0 / np.arange(1, span + 1)
    pr /= pr.sum()
    cps = rng.choice(span, size=(rows.size, 5), p=pr) + lo
    if hi < 0x800:  # 2-byte sequences
      b0 = (0xC0 | (cps >> 6)).astype(np.uint8)
      b1 = (0x80 | (cps & 0x3F)).astype(np.uint8)
      enc = np.stack([b0, b1], axis=2).reshape(rows.size, 10)
      per = 2
    else:  # 3-byte sequences
      b0 = (0xE0 | (cps >> 12)).astype(np.uint8)
      b1 = (0x80 | ((cps >> 6) & 0x3F)).astype(np.uint8)
      b2 = (0x80 | (cps & 0x3F)).astype(np.uint8)
      enc = np.stack([b0, b1, b2], axis=2).reshape(rows.size, 15)
      per = 3
    mat[rows, : enc.shape[1]] = enc
    lens[rows] = ncp[rows] * per
  return mat, lens


def make_types(spec: CorpusSpec, rng: np.random.Generator):
  """Returns (matrix [n_types, W+1] with a separator slot after each word, lens)."""
  if spec.alphabet == "ascii":
    mat, lens = _ascii_types(rng, spec.n_types)
  elif spec.alphabet == "multi":
    n_multi = int(spec.n_types * 0.4)
    a_mat, a_lens = _ascii_types(rng, spec.n_types - n_multi)
    m_mat, m_lens = _multi_types(rng, n_multi)
    width = max(a_mat.shape[1], m_mat.shape[1])
    mat = np.zeros((spec.n_types, width), dtype=np.uint8)
    mat[: a_mat.shape[0], : a_mat.shape[1]] = a_mat
    mat[a_mat.shape[0]:, : m_mat.shape[1]] = m_mat
    lens = np.concatenate([a_lens, m_lens])
    perm = rng.permutation(spec.n_types)  # interleave scripts over the Zipf ranks
    mat, lens = mat[perm], lens[perm]
  else:
    raise ValueError(f"unknown alphabet {spec.alphabet!r}")
  out = np.zeros((mat.shape[0], mat.shape[1] + 1), dtype=np.uint8)
  out[:, : mat.shape[1]] = mat
  return out, lens


def _chunk(spec: CorpusSpec, mat, lens, cdf, chunk_no: int, chunk_words: int) -> np.ndarray:
  """Chunk `chunk_no` of the corpus: chunk_words words drawn with their own counter-based RNG stream,
  so chunks can be produced in any order / in parallel and the corpus is still one fixed byte string."""
  rng = np.random.default_rng([spec.seed, 1 + chunk_no])
  idx = np.searchsorted(cdf, rng.random(chunk_words), side="right")
  np.minimum(idx, spec.n_types - 1, out=idx)
  rows = mat[idx]  # [chunk, width]
  l = lens[idx]
  sep = np.full(chunk_words, ord(" "), dtype=np.uint8)
  sep[WORDS_PER_LINE - 1:: WORDS_PER_LINE] = ord("\n")  # chunk_words is a multiple of WORDS_PER_LINE
  rows[np.arange(chunk_words), l] = sep
  return rows[np.arange(mat.shape[1])[None, :] <= l[:, None]]


def generate(spec: CorpusSpec, chunk_words: int = 1_200_000, threads: int | None = None, first_chunk: int = 0):
  """Yields consecutive uint8 chunks of the corpus; total length == spec.nbytes exactly.

  The last line is cut at spec.nbytes and terminated with '\n' (a cut word is still a word).
  first_chunk > 0 starts the word stream at that chunk number: piece r of a multi-rank corpus is
  generate(spec, first_chunk=r * PIECE_STRIDE) -- same word types, a disjoint part of the sampling stream."""
  from concurrent.futures import ThreadPoolExecutor
  assert chunk_words % WORDS_PER_LINE == 0
  rng = np.random.default_rng(spec.seed)
  mat, lens = make_types(spec, rng)
  ranks = np.arange(1, spec.n_types + 1, dtype=np.float64)
  cdf = np.cumsum(ranks ** (-spec.zipf_s))
  cdf /= cdf[-1]
  threads = threads or min(8, os.cpu_count() or 1)
  produced = 0
  chunk_no = first_chunk
  with ThreadPoolExecutor(threads) as pool:
    while produced < spec.nbytes:
      futs = [pool.submit(_chunk, spec, mat, lens, cdf, chunk_no + k, chunk_words) for k in range(threads)]
      chunk_no += threads
      for f in futs:
        flat = f.result()
        if produced >= spec.nbytes:
          continue
        if produced + flat.size >= spec.nbytes:
          flat = flat[: spec.nbytes - produced].copy()
          flat[-1] = ord("\n")
        produced += flat.size
        yield flat


PIECE_STRIDE = 100_000  # chunk-number distance between the pieces of a multi-rank corpus


# ---- strong scaling: rank r of N holds a contiguous byte range of THE SAME corpus -------------------------------------
# The corpus is the concatenation of the chunks 0, 1, 2, ... cut at spec.nbytes. A chunk's content depends only on its
# number, so rank r can produce chunks [c_r, c_{r+1}) on its own; only the chunk SIZES of the ranks before it are needed to
# know where its range starts, and those are exchanged by the caller (one all-gather of N integers).
def mean_chunk_bytes(spec: CorpusSpec, chunk_words: int = 1_200_000) -> float:
  rng = np.random.default_rng(spec.seed)
  _, lens = make_types(spec, rng)
  ranks = np.arange(1, spec.n_types + 1, dtype=np.float64)
  w = ranks ** (-spec.zipf_s)
  w /= w.sum()
  return float(chunk_words * (np.dot(w, lens) + 1.0))


def piece_chunk_range(spec: CorpusSpec, rank: int, world: int, chunk_words: int = 1_200_000) -> tuple[int, int]:
  """Chunks [c0, c1) of piece `rank`. The total is a slight over-estimate of what spec.nbytes needs (the last piece is
  cut by `cut_piece`), so that the pieces of all ranks together always cover the corpus."""
  total = int(np.ceil(spec.nbytes / mean_chunk_bytes(spec, chunk_words) * 1.003)) + 1
  total = max(total, world)
  return rank * total // world, (rank + 1) * total // world


def generate_chunks(spec: CorpusSpec, c0: int, c1: int, chunk_words: int = 1_200_000, threads: int | None = None):
  """Yields chunks c0 .. c1-1 of the corpus stream, uncut."""
  from concurrent.futures import ThreadPoolExecutor
  rng = np.random.default_rng(spec.seed)
  mat, lens = make_types(spec, rng)
  ranks = np.arange(1, spec.n_types + 1, dtype=np.float64)
  cdf = np.cumsum(ranks ** (-spec.zipf_s))
  cdf /= cdf[-1]
  threads = threads or min(8, os.cpu_count() or 1)
  with ThreadPoolExecutor(threads) as pool:
    for b in range(c0, c1, threads):
      futs = [pool.submit(_chunk, spec, mat, lens, cdf, k, chunk_words) for k in range(b, min(c1, b + threads))]
      for f in futs:
        yield f.result()


def cut_piece(spec: CorpusSpec, piece_sizes: list[int], rank: int) -> tuple[int, int]:
  """(global byte offset, byte count) of piece `rank`, given the UNCUT sizes of all pieces: the corpus ends at
  spec.nbytes, so the piece that crosses that mark is cut there (its last byte becomes a newline, exactly as
  generate() cuts the single-rank corpus) and later pieces are empty."""
  start = int(sum(piece_sizes[:rank]))
  end = min(start + int(piece_sizes[rank]), spec.nbytes)
  assert sum(piece_sizes) >= spec.nbytes, "pieces do not cover the corpus (chunk estimate too low)"
  return min(start, spec.nbytes), max(0, end - start)


def corpus_bytes(spec: CorpusSpec, first_chunk: int = 0) -> np.ndarray:
  """Whole corpus as one uint8 array (use only for sizes that fit in host memory)."""
  out = np.empty(spec.nbytes, dtype=np.uint8)
  pos = 0
  for chunk in generate(spec, first_chunk=first_chunk):
    out[pos: pos + chunk.size] = chunk
    pos += chunk.size
  assert pos == spec.nbytes
  return out


def write_corpus(spec: CorpusSpec, path: str) -> str:
  """Writes the corpus to `path` unless a file of the right size is already there."""
  if os.path.exists(path) and os.path.getsize(path) == spec.nbytes:
    return path
  tmp = path + ".tmp"
  with open(tmp, "wb") as f:
    for chunk in generate(spec):
      f.write(chunk.tobytes())
  os.replace(tmp, path)
  return path


def small_spec(nbytes: int, n_types: int, seed: int, alphabet: str = "ascii", name: str | None = None) -> CorpusSpec:
  return CorpusSpec(name or f"synth_{alphabet}_{nbytes}_{n_types}_{seed}", nbytes, n_types, alphabet, 1.1, seed)


def reference_test_corpus() -> bytes:
  """The fixture of the reference's own test (reference test/bpe_test.cpp:31-56), restated."""
  buf = io.StringIO()
  for line in (
    "the quick brown fox jumps over the lazy dog",
    "the brown fox is quick and the dog is lazy",
    "quick brown foxes jump over lazy dogs",
    "the the the quick quick brown brown fox fox",
    "jumping foxes are quick brown animals",
    "lazy dogs sleep under the brown tree",
    "the quick fox and the lazy dog are friends",
    "brown and quick describe the fox perfectly",
    "the lazy dog watches the quick brown fox",
    "quick movements by the brown fox surprise the dog",
  ):
    buf.write(line + "\n")
  for _ in range(20):
    buf.write("hello world hello world programming programming\n")
    buf.write("testing testing the the quick quick brown brown\n")
    buf.write("algorithm algorithm implementation implementation\n")
  return buf.getvalue().encode("ascii")
