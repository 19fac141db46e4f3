"""ctypes front-end of oracle/liboracle.so and oracle/_ref/ref_driver -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker. The product (shredword_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")
REF_LIB = os.path.join(HERE, "_ref", "libtrainer_ref.so")


def build(ref: bool = False) -> None:
  """Compiles the C restatement (and, where /root/reference exists, the unmodified reference)."""
  subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"], check=True)
  if ref and os.path.isdir("/root/reference/shredword/csrc"):
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_lib = None


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      build()
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u64, f32 = C.c_void_p, C.c_size_t, C.c_int32, C.c_uint64, C.c_float
    L.oracle_create.argtypes = [sz, i32, f32, u64]; L.oracle_create.restype = vp
    L.oracle_destroy.argtypes = [vp]; L.oracle_destroy.restype = None
    L.oracle_load_corpus.argtypes = [vp, C.c_char_p]; L.oracle_load_corpus.restype = C.c_int
    L.oracle_load_corpus_buffer.argtypes = [vp, vp, sz]; L.oracle_load_corpus_buffer.restype = C.c_int
    L.oracle_stream_begin.argtypes = []; L.oracle_stream_begin.restype = vp
    L.oracle_stream_feed.argtypes = [vp, vp, sz]; L.oracle_stream_feed.restype = C.c_int
    L.oracle_stream_finish.argtypes = [vp, vp]; L.oracle_stream_finish.restype = C.c_int
    L.oracle_init.argtypes = [vp]; L.oracle_init.restype = None
    L.oracle_count_bigrams.argtypes = [vp]; L.oracle_count_bigrams.restype = None
    L.oracle_merge_batch.argtypes = [vp, C.c_int]; L.oracle_merge_batch.restype = C.c_int
    L.oracle_train.argtypes = [vp]; L.oracle_train.restype = C.c_int
    L.oracle_save.argtypes = [vp, C.c_char_p, C.c_char_p]; L.oracle_save.restype = C.c_int
    for name in ("oracle_num_merges", "oracle_num_words", "oracle_num_symbols", "oracle_word_bytes_total",
                 "oracle_heap_size", "oracle_num_pairs"):
      getattr(L, name).argtypes = [vp]; getattr(L, name).restype = sz
    L.oracle_get_merges.argtypes = [vp, vp]; L.oracle_get_merges.restype = None
    L.oracle_get_words.argtypes = [vp, vp, vp, vp, vp, vp]; L.oracle_get_words.restype = None
    L.oracle_get_keep.argtypes = [vp, vp]; L.oracle_get_keep.restype = None
    L.oracle_get_heap.argtypes = [vp, vp, vp, vp, vp]; L.oracle_get_heap.restype = None
    L.oracle_get_pairs.argtypes = [vp, vp, vp, vp, vp]; L.oracle_get_pairs.restype = None
    L.oracle_token_freq.argtypes = [vp, vp]; L.oracle_token_freq.restype = None
    L.oracle_shard_count.argtypes = [vp, C.c_int, C.c_int, vp, sz]; L.oracle_shard_count.restype = sz
    L.oracle_shard_merge.argtypes = [vp, C.c_int, C.c_int, i32, i32, i32, vp, sz]; L.oracle_shard_merge.restype = sz
    L.oracle_encode.argtypes = [vp, sz, vp, vp, sz, vp, sz, vp, sz, vp]; L.oracle_encode.restype = sz
    L.oracle_decode.argtypes = [vp, sz, vp, sz, vp, sz]; L.oracle_decode.restype = sz
    L.oracle_normalize_line.argtypes = [vp, sz, vp, sz]; L.oracle_normalize_line.restype = sz
    L.oracle_normalize_text.argtypes = [vp, sz, vp, sz]; L.oracle_normalize_text.restype = sz
    _lib = L
  return _lib


def _p(a):
  return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
  """Mirror of the reference BPETrainer API over the C restatement."""

  def __init__(self, target_vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000):
    self.h = lib().oracle_create(target_vocab_size, unk_id, character_coverage, min_pair_freq)

  def __del__(self):
    if getattr(self, "h", None):
      lib().oracle_destroy(self.h)
      self.h = None

  def load_corpus(self, path: str):
    if lib().oracle_load_corpus(self.h, path.encode()) != 0:
      raise IOError(f"Failed to load corpus from {path}")

  def load_buffer(self, data) -> int:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    a = np.ascontiguousarray(a)
    return lib().oracle_load_corpus_buffer(self.h, _p(a), a.size)

  def load_chunks(self, chunks) -> int:
    """Streaming load: `chunks` yields uint8 arrays that end on a delimiter (corpora larger than host memory)."""
    st = lib().oracle_stream_begin()
    for c in chunks:
      a = np.ascontiguousarray(c, dtype=np.uint8)
      if lib().oracle_stream_feed(st, _p(a), a.size) != 0:
        break
    return lib().oracle_stream_finish(st, self.h)

  def init(self):
    lib().oracle_init(self.h)

  def count_bigrams(self):
    lib().oracle_count_bigrams(self.h)

  def merge_batch(self, n: int) -> int:
    return lib().oracle_merge_batch(self.h, n)

  def train(self) -> int:
    return lib().oracle_train(self.h)

  def save(self, model_path: str, vocab_path: str):
    if lib().oracle_save(self.h, model_path.encode(), vocab_path.encode()) != 0:
      raise IOError("oracle_save failed")

  # ---- inspection
  @property
  def merges(self) -> np.ndarray:
    n = lib().oracle_num_merges(self.h)
    out = np.zeros((n, 3), dtype=np.int32)
    if n:
      lib().oracle_get_merges(self.h, _p(out))
    return out

  @property
  def num_words(self) -> int:
    return lib().oracle_num_words(self.h)

  @property
  def num_symbols(self) -> int:
    return lib().oracle_num_symbols(self.h)

  def words(self):
    """(byte_off[W+1], bytes, sym_off[W+1], syms, counts[W]) in reference word order."""
    W = self.num_words
    boff = np.zeros(W + 1, dtype=np.uint64)
    soff = np.zeros(W + 1, dtype=np.uint64)
    cnt = np.zeros(W, dtype=np.uint64)
    by = np.zeros(lib().oracle_word_bytes_total(self.h), dtype=np.uint8)
    sy = np.zeros(self.num_symbols, dtype=np.int32)
    lib().oracle_get_words(self.h, _p(boff), _p(by), _p(soff), _p(sy), _p(cnt))
    return boff, by, soff, sy, cnt

  def keep(self) -> np.ndarray:
    k = np.zeros(256, dtype=np.uint8)
    lib().oracle_get_keep(self.h, _p(k))
    return k

  def heap(self):
    n = lib().oracle_heap_size(self.h)
    f, s = np.zeros(n, np.int32), np.zeros(n, np.int32)
    fr, v = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    if n:
      lib().oracle_get_heap(self.h, _p(f), _p(s), _p(fr), _p(v))
    return f, s, fr, v

  def pairs(self):
    n = lib().oracle_num_pairs(self.h)
    f, s = np.zeros(n, np.int32), np.zeros(n, np.int32)
    fr, v = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    if n:
      lib().oracle_get_pairs(self.h, _p(f), _p(s), _p(fr), _p(v))
    return f, s, fr, v

  def token_freq(self) -> np.ndarray:
    out = np.zeros(256 + lib().oracle_num_merges(self.h), dtype=np.uint64)
    lib().oracle_token_freq(self.h, _p(out))
    return out

  def byte_map(self, unk_id: int) -> np.ndarray:
    k = self.keep()
    m = np.arange(256, dtype=np.int32)
    m[k == 0] = unk_id
    return m

  # ---- sharded building blocks (one rank's share of the work, for world_size>1 tests)
  def shard_count(self, rank: int, nranks: int) -> np.ndarray:
    cap = 1 << 16
    while True:
      recs = np.zeros((cap, 4), dtype=np.int64)
      n = lib().oracle_shard_count(self.h, rank, nranks, _p(recs), cap)
      if n <= cap:
        return recs[:n]
      cap = n

  def shard_merge(self, rank: int, nranks: int, a: int, b: int, new_id: int) -> np.ndarray:
    cap = 1 << 20  # must be large enough the first time: the call mutates the shard
    recs = np.zeros((cap, 4), dtype=np.int64)
    n = lib().oracle_shard_merge(self.h, rank, nranks, a, b, new_id, _p(recs), cap)
    if n > cap:
      raise RuntimeError("shard_merge record buffer too small")
    return recs[:n]


def encode(merges: np.ndarray, byte_map: np.ndarray, text, with_word_counts: bool = False):
  """Rank-ordered BPE encode of `text` (bytes / uint8 array). Returns ids (and per-word token counts)."""
  t = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text)
  m = np.ascontiguousarray(merges, dtype=np.int32).reshape(-1, 3)
  bm = np.ascontiguousarray(byte_map, dtype=np.int32)
  out = np.zeros(max(t.size, 1), dtype=np.int32)
  wn = np.zeros(max(t.size // 2 + 1, 1), dtype=np.uint32) if with_word_counts else None
  nw = C.c_size_t(0)
  n = lib().oracle_encode(_p(m), m.shape[0], _p(bm), _p(t), t.size, _p(out), out.size, _p(wn),
                          0 if wn is None else wn.size, C.byref(nw))
  if with_word_counts:
    return out[:n].copy(), wn[: nw.value].copy()
  return out[:n].copy()


def decode(merges: np.ndarray, ids: np.ndarray) -> bytes:
  m = np.ascontiguousarray(merges, dtype=np.int32).reshape(-1, 3)
  ids = np.ascontiguousarray(ids, dtype=np.int32)
  n = lib().oracle_decode(_p(m), m.shape[0], _p(ids), ids.size, None, 0)
  out = np.zeros(max(n, 1), dtype=np.uint8)
  lib().oracle_decode(_p(m), m.shape[0], _p(ids), ids.size, _p(out), out.size)
  return out[:n].tobytes()


# ---------------------------------------------------------------- the unmodified reference
def normalize_line(line: bytes) -> bytes:
  """normalize_line of reference csrc/bpe/normalize.cpp:24-59, restated (one line, no NUL)."""
  a = np.frombuffer(bytes(line), dtype=np.uint8)
  out = np.zeros(3 * a.size + 8, dtype=np.uint8)
  n = lib().oracle_normalize_line(_p(a) if a.size else None, a.size, _p(out), out.size)
  return out[:n].tobytes()


def normalize_text(text) -> bytes:
  """The restated normalize_line applied to every line of a text (lines joined by the newline again)."""
  a = np.ascontiguousarray(text if isinstance(text, np.ndarray) else np.frombuffer(bytes(text), dtype=np.uint8), dtype=np.uint8)
  out = np.zeros(3 * a.size + 8, dtype=np.uint8)
  n = lib().oracle_normalize_text(_p(a) if a.size else None, a.size, _p(out), out.size)
  return out[:n].tobytes()


def ref_normalize_line(line: bytes, cap: int = 1 << 20) -> bytes:
  """The REFERENCE's own normalize_line (oracle/_ref/libtrainer_ref.so), for generating / checking golden vectors."""
  L = C.CDLL(REF_LIB)
  L.normalize_line.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]; L.normalize_line.restype = C.c_int
  buf = C.create_string_buffer(cap)
  n = L.normalize_line(bytes(line), buf, cap)
  assert n >= 0
  return buf.raw[:n]


def ref_available() -> bool:
  return os.path.exists(REF_DRIVER) and os.path.exists(REF_LIB)


def run_reference(corpus_path: str, vocab_size: int, min_pair_freq: int, model_out: str, vocab_out: str,
                  unk_id: int = 0, coverage: float = 0.995, max_merges: int = -1, timeout: float | None = None) -> dict:
  """Runs the unmodified reference (zero-filling malloc) through its C-ABI; returns phase timings."""
  if not ref_available():
    raise FileNotFoundError("oracle/_ref is not built (needs /root/reference): make -C oracle ref")
  cmd = [REF_DRIVER, corpus_path, str(vocab_size), str(min_pair_freq), str(unk_id), repr(float(coverage)),
         model_out, vocab_out, str(max_merges)]
  r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=timeout, check=True)
  return json.loads(r.stderr.decode().strip().splitlines()[-1])


def read_model(path: str) -> np.ndarray:
  """The reference's binary .model: M x 3 little-endian int32 (a, b, 256+m), no header (bpe.cpp:722-732)."""
  return np.fromfile(path, dtype="<i4").reshape(-1, 3)
