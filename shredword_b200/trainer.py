"""BPETrainer -- same class, arguments and error behaviour as reference shredword/trainer.py:5-39,
backed by the B200 library. Additive: the `merges` / `vocab` / `special_tokens` properties the
reference README describes (README.md:64-69, 89-97) but its trainer never had, in-memory loading, and
an encoder bound to the trained merges."""
import ctypes
from typing import *

import numpy as np

from .cbase import lib, BPEConfig, SwbStats, last_error

def _ptr(a):
  return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

def bind_host_thread_to_gpu(device_index: int) -> list[int] | None:
  """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs). The merge loop is a
  latency chain through mapped host memory: a host thread (and its pinned pages, first touch) on the far socket adds a
  hop to every PCIe round trip. Returns the CPU list, or None when the topology cannot be read."""
  import os
  try:
    import ctypes as _c
    buf = _c.create_string_buffer(64)
    if lib.swb_device_pci_bus_id(device_index, buf, 64) != 0:
      return None
    bus = buf.value.decode().lower()
    with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
      spec = f.read().strip()
    cpus = []
    for part in spec.split(","):
      if "-" in part:
        lo, hi = part.split("-"); cpus.extend(range(int(lo), int(hi) + 1))
      elif part:
        cpus.append(int(part))
    allowed = sorted(set(cpus) & os.sched_getaffinity(0))
    if not allowed:
      return None
    os.sched_setaffinity(0, allowed)
    return allowed
  except Exception:
    return None


def normalize(text) -> bytes:
  """Optional pre-pass on the GPU: the reference's normalize_line (reference csrc/bpe/normalize.cpp:24-59) applied to every
  line of `text` (bytes / uint8 array): lower-cases ASCII, turns whitespace runs inside a line into U+2581, drops them at
  the ends of a line. Feed the result to load_buffer / encode for a SentencePiece-style pipeline."""
  a = text if isinstance(text, np.ndarray) else np.frombuffer(bytes(text), dtype=np.uint8)
  a = np.ascontiguousarray(a, dtype=np.uint8)
  out = np.empty(3 * a.size + 16, dtype=np.uint8)
  n = lib.swb_normalize(_ptr(a), a.size, _ptr(out), out.size, 0)
  if n < 0:
    raise RuntimeError(f"normalize failed: {last_error()}")
  return out[:n].tobytes()


_PT_UNMAP = bytes.maketrans(b"\x1c\x1d\x1e\x1f", b" \t\n\r")


def pretokenize(text) -> bytes:
  """Optional pre-pass on the GPU: the reference's regex pre-tokenisation (reference shredword/base.py:38-58). Returns the
  UTF-8 text with one ' ' behind every piece of the reference's split pattern and the bytes ' ' \\t \\n \\r inside pieces
  mapped to 0x1C-0x1F: feed it to load_buffer / encode and the trainer's words are exactly the reference's pieces.
  `undo_pretokenize` restores the text."""
  if isinstance(text, str):
    text = text.encode("utf-8")
  a = text if isinstance(text, np.ndarray) else np.frombuffer(bytes(text), dtype=np.uint8)
  a = np.ascontiguousarray(a, dtype=np.uint8)
  out = np.empty(2 * a.size + 16, dtype=np.uint8)
  n = lib.swb_pretokenize(_ptr(a), a.size, _ptr(out), out.size, 0)
  if n < 0:
    raise RuntimeError(f"pretokenize failed: {last_error()}")
  return out[:n].tobytes()


def undo_pretokenize(data) -> bytes:
  return bytes(data).replace(b" ", b"").translate(_PT_UNMAP)


def apply_regex(text: str) -> list:
  """Same result as the reference's apply_regex (shredword/base.py:38-58): the list of pieces, computed on the GPU."""
  out = pretokenize(text)
  return [p.translate(_PT_UNMAP).decode("utf-8") for p in out.split(b" ")[:-1]] if out else []


class BPETrainer:
  def __init__(self, target_vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000):
    self.config = BPEConfig(
      target_vocab_size=target_vocab_size,
      unk_id=unk_id,
      character_coverage=character_coverage,
      min_pair_freq=min_pair_freq
    )
    self._special_tokens: List[Tuple[str, int]] = []
    self.trainer = lib.create_trainer(ctypes.byref(self.config))
    if not self.trainer:
      raise RuntimeError("Failed to create BPE trainer")

  def load_corpus(self, path: str):
    result = lib.bpe_load_corpus(self.trainer, path.encode('utf-8'))
    if result != 0:
      raise IOError(f"Failed to load corpus from {path}")

  def train(self):
    merges = lib.bpe_train(self.trainer)
    if merges < 0:
      raise RuntimeError("Training failed")
    print(f"Training completed: {merges} merges performed.")

  def save(self, model_path: str, vocab_path: str):
    lib.bpe_save(self.trainer, model_path.encode('utf-8'), vocab_path.encode('utf-8'))
    print(f"Model saved to: {model_path}")
    print(f"Vocabulary saved to: {vocab_path}")

  def destroy(self):
    if getattr(self, "trainer", None):
      lib.bpe_trainer_destroy(self.trainer)
      self.trainer = None

  def __del__(self):
    self.destroy()

  # ---------------------------------------------------------------- additive
  def load_buffer(self, data) -> None:
    """load_corpus from host memory (bytes, bytearray, or a uint8 numpy array)."""
    a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if lib.swb_load_corpus_buffer(self.trainer, _ptr(a), a.size) != 0:
      raise IOError(f"Failed to load corpus from memory: {last_error()}")

  def load_device(self, device_ptr: int, nbytes: int) -> None:
    """load_corpus from a CUDA device pointer (e.g. torch_tensor.data_ptr())."""
    if lib.swb_load_corpus_device(self.trainer, ctypes.c_void_p(device_ptr), nbytes) != 0:
      raise IOError(f"Failed to load corpus from device memory: {last_error()}")

  def init(self):
    lib.bpe_init(self.trainer)

  def count_bigrams(self):
    lib.bpe_count_bigrams(self.trainer)

  def merge_batch(self, n: int) -> int:
    r = lib.bpe_merge_batch(self.trainer, n)
    if r < 0:
      raise RuntimeError(f"merge_batch failed: {last_error()}")
    return r

  def train_quiet(self) -> int:
    merges = lib.bpe_train(self.trainer)
    if merges < 0:
      raise RuntimeError(f"Training failed: {last_error()}")
    return merges

  @property
  def num_merges(self) -> int:
    return lib.swb_num_merges(self.trainer)

  @property
  def merges(self) -> List[Tuple[int, int, int]]:
    """Merge rules in rank order as (a, b, new_id) (reference README.md:95)."""
    return [tuple(int(x) for x in row) for row in self.merges_array()]

  def merges_array(self) -> np.ndarray:
    n = self.num_merges
    out = np.zeros((n, 3), dtype=np.int32)
    if n:
      lib.swb_get_merges(self.trainer, _ptr(out), n)
    return out

  @property
  def vocab(self) -> Dict[int, bytes]:
    """id -> bytes (reference README.md:67, base.py:74-79)."""
    out = {}
    buf = np.zeros(1 << 16, dtype=np.uint8)
    for i in range(256 + self.num_merges):
      n = lib.swb_token_bytes(self.trainer, i, _ptr(buf), buf.size)
      if n > buf.size:
        buf = np.zeros(n, dtype=np.uint8)
        n = lib.swb_token_bytes(self.trainer, i, _ptr(buf), buf.size)
      out[i] = buf[:n].tobytes()
    return out

  @property
  def special_tokens(self) -> List[Tuple[str, int]]:
    """(token, id) pairs (reference README.md:91). The trainer reserves exactly one: unk."""
    return [("<UNK>", int(self.config.unk_id))] + list(self._special_tokens)

  @special_tokens.setter
  def special_tokens(self, value: Iterable[Tuple[str, int]]):
    self._special_tokens = [(str(s), int(i)) for s, i in value if s != "<UNK>"]

  def byte_map(self) -> np.ndarray:
    m = np.zeros(256, dtype=np.int32)
    lib.swb_get_byte_map(self.trainer, _ptr(m))
    return m

  def token_freq(self) -> np.ndarray:
    out = np.zeros(256 + self.num_merges, dtype=np.uint64)
    if lib.swb_token_freq(self.trainer, _ptr(out), out.size) != 0:
      raise RuntimeError(last_error())
    return out

  def words(self):
    """(byte_off[W+1], bytes, sym_off[W+1], syms, counts[W]) in reference word order."""
    W = lib.swb_num_words(self.trainer)
    boff = np.zeros(W + 1, dtype=np.uint64); soff = np.zeros(W + 1, dtype=np.uint64)
    cnt = np.zeros(W, dtype=np.uint64)
    by = np.zeros(lib.swb_word_bytes_total(self.trainer), dtype=np.uint8)
    sy = np.zeros(lib.swb_num_symbols(self.trainer), dtype=np.int32)
    if lib.swb_get_words(self.trainer, _ptr(boff), _ptr(by), _ptr(soff), _ptr(sy), _ptr(cnt)) != 0:
      raise RuntimeError(last_error())
    return boff, by, soff, sy, cnt

  def stats(self) -> dict:
    s = SwbStats()
    lib.swb_get_stats(self.trainer, ctypes.byref(s))
    out = {}
    for k, _ in SwbStats._fields_:
      if k.endswith("_"):
        continue
      v = getattr(s, k)
      out[k] = list(v) if hasattr(v, "__len__") else v
    return out

  def profile_scripted_merges(self, merge_triples: np.ndarray) -> float:
    """Profiling aid (swb_profile_scripted_merges): after init() on a fresh load, runs the given merges inside one
    launch of the resident kernel with no host in the loop; returns the launch duration in ms. Consumes the handle."""
    m = np.ascontiguousarray(merge_triples, dtype=np.int32).reshape(-1, 3)
    ms = ctypes.c_double(0.0)
    if lib.swb_profile_scripted_merges(self.trainer, _ptr(m), m.shape[0], ctypes.byref(ms)) != 0:
      raise RuntimeError(f"profile_scripted_merges failed: {last_error()}")
    return float(ms.value)

  def set_kernel_timing(self, on: bool):
    lib.swb_set_kernel_timing(self.trainer, 1 if on else 0)

  def encoder(self) -> "BPEEncoder":
    return BPEEncoder(_handle=lib.swb_encoder_from_trainer(self.trainer))


class BPEEncoder:
  """Rank-ordered BPE encoder on the GPU (no reference counterpart: base.py:107-109 is unimplemented)."""

  def __init__(self, merges=None, byte_map=None, _handle=None):
    if _handle is None:
      m = np.ascontiguousarray(np.asarray(merges, dtype=np.int32).reshape(-1, 3))
      bm = None if byte_map is None else np.ascontiguousarray(byte_map, dtype=np.int32)
      _handle = lib.swb_encoder_create(_ptr(m), m.shape[0], _ptr(bm))
    if not _handle:
      raise RuntimeError(f"Failed to create encoder: {last_error()}")
    self.h = _handle

  @classmethod
  def from_model_file(cls, model_path: str, byte_map=None) -> "BPEEncoder":
    """Loads the reference's binary .model (M x 3 little-endian int32, reference bpe.cpp:722-732)."""
    return cls(np.fromfile(model_path, dtype="<i4").reshape(-1, 3), byte_map)

  def encode(self, text, with_word_counts: bool = False):
    a = text if isinstance(text, np.ndarray) else np.frombuffer(text.encode("utf-8") if isinstance(text, str) else text, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.empty(max(a.size, 1), dtype=np.int32)
    wn = np.empty(a.size // 2 + 1, dtype=np.uint32) if with_word_counts else None
    nw = ctypes.c_size_t(0)
    n = lib.swb_encode(self.h, _ptr(a), a.size, _ptr(out), out.size, _ptr(wn), 0 if wn is None else wn.size, ctypes.byref(nw))
    if n < 0:
      raise RuntimeError(f"encode failed: {last_error()}")
    if with_word_counts:
      return out[:n], wn[: nw.value]
    return out[:n]

  def encode_into(self, text: np.ndarray, out: np.ndarray) -> int:
    """Encodes a uint8 array into a preallocated int32 array (both may be pinned); returns the token count."""
    n = lib.swb_encode(self.h, _ptr(text), text.size, _ptr(out), out.size, None, 0, None)
    if n < 0:
      raise RuntimeError(f"encode failed: {last_error()}")
    return int(n)

  def encode_device(self, d_text: int, nbytes: int, d_out: int, cap_ids: int) -> int:
    n = lib.swb_encode_device(self.h, ctypes.c_void_p(d_text), nbytes, ctypes.c_void_p(d_out), cap_ids, None, 0, None)
    if n < 0:
      raise RuntimeError(f"encode failed: {last_error()}")
    return int(n)

  def decode(self, ids) -> bytes:
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    n = lib.swb_decode(self.h, _ptr(ids), ids.size, None, 0)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    lib.swb_decode(self.h, _ptr(ids), ids.size, _ptr(out), out.size)
    return out[:n].tobytes()

  @property
  def kernel_launches(self) -> int:
    return int(lib.swb_encoder_kernel_launches(self.h))

  def __del__(self):
    if getattr(self, "h", None):
      lib.swb_encoder_destroy(self.h)
      self.h = None
