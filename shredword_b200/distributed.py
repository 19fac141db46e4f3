"""Multi-GPU BPE training.

Default (replicated merge loop): the corpus is range-split over the ranks, every rank tokenises its range and the
unique-word tables are exchanged over NCCL (load_shard); every rank then holds all unique words and runs the
single-GPU resident merge loop -- identical results on every rank, no collective per merge.

sharded_merge=True: unique words sharded over ranks, pair table + heap replicated.
One process per GPU (torch.distributed). Word wi of the reference word order belongs to rank
wi % world_size; each rank's kernels count / merge only its own rows. Per merge every rank emits its
local (pair, net delta, first-touch key) records, the records are all-gathered, reduced by pair
(sum delta, min key) and applied to an identical replica of the pair table and the exact heap on
every rank -- so all ranks pop the same pair next without any further agreement step (SURVEY.md
8(e): the push ORDER has to be serialised somewhere; a replica does that with one small exchange).

The exchange is torch.distributed (NCCL between GPUs, gloo in the CPU tests). `local_ops` lets a test
substitute a CPU stand-in for this rank's kernels; the product path always uses the CUDA library.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from .cbase import lib, last_error
from .trainer import BPETrainer, _ptr


class _CudaLocalOps:
  """This rank's kernels through the C-ABI (swb_shard_count / swb_shard_merge)."""

  def __init__(self, trainer: BPETrainer):
    self.t = trainer
    self.buf = np.zeros((1 << 16, 4), dtype=np.int64)

  def _call(self, fn, *args):
    while True:
      n = fn(self.t.trainer, *args, _ptr(self.buf), self.buf.shape[0])
      if n >= 0:
        return self.buf[:n]
      if "too small" not in last_error():
        raise RuntimeError(last_error())
      self.buf = np.zeros((self.buf.shape[0] * 4, 4), dtype=np.int64)  # only the count pass can retry safely

  def count(self) -> np.ndarray:
    return self._call(lib.swb_shard_count)

  def merge(self, a: int, b: int, new_id: int) -> np.ndarray:
    n = lib.swb_shard_merge(self.t.trainer, a, b, new_id, _ptr(self.buf), self.buf.shape[0])
    if n < 0:
      raise RuntimeError(last_error())
    return self.buf[:n]


class DistributedBPETrainer(BPETrainer):
  """BPETrainer whose word table is sharded over the ranks of `group` (default: the world)."""

  def __init__(self, *args, group=None, device: torch.device | None = None, local_ops=None, native: bool | None = None,
               sharded_merge: bool = False, **kw):
    """native=True (default on CUDA devices): NCCL is driven from C++ inside the library (swb_dist_init).
    With sharded_merge=False (default) only the load is split -- range-split tokenising + NCCL word-table
    exchange in load_shard -- and every rank then runs the single-GPU merge loop on all unique words (no
    collective per merge: the loop is latency-bound); sharded_merge=True keeps each
    rank's words on that rank and all-gathers the delta records of every merge.
    native=False: the sharded loop with the exchange driven from Python through torch.distributed (any
    backend; what the gloo tests use)."""
    super().__init__(*args, **kw)
    self.group = group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    self.device = device if device is not None else torch.device("cpu")
    if native is None:
      native = local_ops is None and self.device.type == "cuda"
    self.native = bool(native)
    self.sharded = bool(sharded_merge) or not self.native
    self.exchange_bytes = 0
    self.collectives = 0
    if self.native:
      uid = None
      if not lib.swb_dist_has_comm(self.rank, self.world):  # the communicator is created once per process
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0 and lib.swb_dist_unique_id(_ptr(uid)) != 0:
          raise RuntimeError(last_error())
        box = [uid.tobytes()]
        dist.broadcast_object_list(box, src=0, group=group)
        uid = np.frombuffer(box[0], dtype=np.uint8).copy()
      if lib.swb_dist_init(self.trainer, self.rank, self.world, _ptr(uid)) != 0:
        raise RuntimeError(last_error())
      if lib.swb_dist_set_sharded(self.trainer, 1 if self.sharded else 0) != 0:
        raise RuntimeError(last_error())
      self.local = None
      return
    if lib.swb_set_shard(self.trainer, self.rank, self.world) != 0:
      raise RuntimeError(last_error())
    self.local = local_ops if local_ops is not None else _CudaLocalOps(self)

  def load_shard(self, data, global_offset: int) -> None:
    """Range-split load (native mode): `data` is this rank's byte range of the corpus (cut on delimiters),
    starting at byte `global_offset` of the whole. A host buffer/array, or a CUDA torch tensor."""
    if isinstance(data, torch.Tensor) and data.is_cuda:
      rc = lib.swb_load_corpus_shard(self.trainer, ctypes.c_void_p(data.data_ptr()), data.numel(), global_offset, 1)
    else:
      a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
      a = np.ascontiguousarray(a, dtype=np.uint8)
      rc = lib.swb_load_corpus_shard(self.trainer, _ptr(a), a.size, global_offset, 0)
    if rc != 0:
      raise IOError(f"Failed to load corpus shard: {last_error()}")

  # ---- the one exchange step: variable-length record lists -> every rank has all of them
  def _allgather_records(self, recs: np.ndarray) -> np.ndarray:
    n_local = torch.tensor([recs.shape[0]], dtype=torch.int64, device=self.device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
    dist.all_gather(sizes, n_local, group=self.group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    self.collectives += 1
    if mx == 0:
      return np.zeros((0, 4), dtype=np.int64)
    send = torch.zeros((mx, 4), dtype=torch.int64, device=self.device)
    if recs.shape[0]:
      send[: recs.shape[0]] = torch.from_numpy(np.ascontiguousarray(recs)).to(self.device)
    out = [torch.empty_like(send) for _ in range(self.world)]
    dist.all_gather(out, send, group=self.group)
    self.collectives += 1
    self.exchange_bytes += mx * 32 * self.world
    parts = [o[:s].cpu().numpy() for o, s in zip(out, sizes) if s]
    return np.ascontiguousarray(np.concatenate(parts, axis=0))

  def _reduce(self, allrecs: np.ndarray) -> np.ndarray:
    if allrecs.shape[0] == 0:
      return allrecs
    n = lib.swb_dist_reduce_records(_ptr(allrecs), allrecs.shape[0])
    return allrecs[:n]

  def init(self):
    if self.native:
      return super().init()
    allrecs = self._reduce(self._allgather_records(self.local.count()))
    lib.swb_dist_seed(self.trainer, _ptr(allrecs), allrecs.shape[0])

  def merge_batch(self, n: int) -> int:
    if self.native:
      return super().merge_batch(n)
    a, b, nid = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    done = 0
    while done < n:
      if not lib.swb_dist_next_merge(self.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid)):
        break
      allrecs = self._reduce(self._allgather_records(self.local.merge(a.value, b.value, nid.value)))
      lib.swb_dist_apply(self.trainer, _ptr(allrecs), allrecs.shape[0])
      done += 1
    return done

  def train_quiet(self) -> int:
    if self.native:
      return super().train_quiet()
    self.init()
    target = int(self.config.target_vocab_size) - 256
    total = 0
    while total < target:
      got = self.merge_batch(target - total)
      if got <= 0:
        break
      total += got
    return total

  def train(self):
    merges = self.train_quiet()
    if self.rank == 0:
      print(f"Training completed: {merges} merges performed.")

  def token_freq(self) -> np.ndarray:
    """Global token histogram: sum of the ranks' shard histograms."""
    if not self.sharded:  # every rank holds all words: its histogram is the global one
      return super().token_freq()
    local = torch.from_numpy(super().token_freq().astype(np.int64)).to(self.device)
    dist.all_reduce(local, group=self.group)
    return local.cpu().numpy().astype(np.uint64)

  def save(self, model_path: str, vocab_path: str):
    """Rank 0 writes the two files with the all-reduced token histogram; the other ranks only take part
    in the reduction."""
    freq = np.ascontiguousarray(self.token_freq(), dtype=np.uint64)
    if self.rank == 0:
      if lib.swb_save_with_freq(self.trainer, model_path.encode("utf-8"), vocab_path.encode("utf-8"), _ptr(freq), freq.size) != 0:
        raise IOError(last_error())
      print(f"Model saved to: {model_path}")
      print(f"Vocabulary saved to: {vocab_path}")
