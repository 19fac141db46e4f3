"""Multi-GPU BPE training and document-parallel encoding.

The corpus is range-split over the ranks, every rank tokenises its range and the unique-word tables are exchanged over
NCCL (load_shard) -- the part of the path whose work grows with the corpus. The merge loop itself is one latency chain
(each merge depends on the heap decision after the one before it, and touches only the few words that hold the pair), so
more GPUs cannot shorten it:
  merge_loop="rank0" (default): rank 0 runs the single-GPU resident loop on all unique words; the merge list and the
      byte map (a few hundred KB) are broadcast afterwards. One host thread, one GPU busy, nothing contending.
  merge_loop="replicated": every rank runs that same loop (identical results everywhere, no collective per merge).

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
    # a merge touches at most 4 pairs per distinct neighbour symbol: 8 x (initial symbols + merges) records can never
    # overflow, so swb_shard_merge (which has already rewritten the symbol stream when it reports) cannot fail on size
    self.buf = np.zeros((max(1 << 16, 8 * (258 + int(trainer.config.target_vocab_size))), 4), dtype=np.int64)

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
               sharded_merge: bool = False, merge_loop: str | None = None, **kw):
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
    if merge_loop is None:
      merge_loop = "sharded" if (sharded_merge or not self.native) else "rank0"
    if merge_loop not in ("rank0", "replicated", "sharded"):
      raise ValueError(f"merge_loop must be rank0, replicated or sharded, not {merge_loop!r}")
    if not self.native and merge_loop != "sharded":
      raise ValueError("the Python-driven exchange (native=False) only runs the sharded merge loop")
    self.merge_loop = merge_loop
    self.sharded = merge_loop == "sharded"
    self._bcast = None  # merge_loop == "rank0", ranks > 0: (merges [M, 3] int32, byte_map [256] int32, token_freq) from rank 0
    self.exchange_bytes = 0
    self.collectives = 0
    if self.native:
      uid = None
      if not lib.swb_dist_has_comm(self.rank, self.world):  # the communicator is created once per process
        uid = np.zeros(128, dtype=np.uint8)
        err = ""
        if self.rank == 0 and lib.swb_dist_unique_id(_ptr(uid)) != 0:
          err = last_error() or "swb_dist_unique_id failed"
        box = [uid.tobytes(), err]  # (the status travels with the id: a failure on rank 0 raises on every rank, nobody hangs)
        dist.broadcast_object_list(box, src=0, group=group)
        if box[1]:
          raise RuntimeError(f"rank 0 could not create the NCCL id: {box[1]}")
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

  def _idle_rank(self) -> bool:
    return self.merge_loop == "rank0" and self.rank != 0

  def _broadcast_result(self, failed: bool = False):
    """merge_loop == "rank0": rank 0's merge list, byte map and token histogram go to every rank (one broadcast of sizes,
    one of the payload; NCCL when the device is a GPU). failed (rank 0 only): training failed there -- every rank learns it
    (a negative merge count) and raises, instead of waiting for a payload that never comes."""
    if self.merge_loop != "rank0":
      return
    if self.rank == 0 and not failed:
      m = super().merges_array().astype(np.int64).reshape(-1)
      bm = super().byte_map().astype(np.int64)
      tf = super().token_freq().astype(np.int64)
      payload = np.concatenate([m, bm, tf])
      hdr = torch.tensor([m.size // 3, payload.size], dtype=torch.int64, device=self.device)
    elif self.rank == 0:
      hdr = torch.tensor([-1, 0], dtype=torch.int64, device=self.device)
    else:
      hdr = torch.zeros(2, dtype=torch.int64, device=self.device)
    dist.broadcast(hdr, src=0, group=self.group)
    n_merges, n = int(hdr[0].item()), int(hdr[1].item())
    if n_merges < 0:
      if self.rank != 0:
        raise RuntimeError("training failed on rank 0 (which runs the merge loop)")
      return
    buf = torch.from_numpy(payload).to(self.device) if self.rank == 0 else torch.empty(n, dtype=torch.int64, device=self.device)
    dist.broadcast(buf, src=0, group=self.group)
    self.collectives += 2
    self.exchange_bytes += 8 * n
    if self.rank != 0:
      a = buf.cpu().numpy()
      self._bcast = (a[: 3 * n_merges].astype(np.int32).reshape(-1, 3), a[3 * n_merges: 3 * n_merges + 256].astype(np.int32),
                     a[3 * n_merges + 256:].astype(np.uint64))

  def init(self):
    if self._idle_rank():
      return
    if self.native:
      return super().init()
    allrecs = self._reduce(self._allgather_records(self.local.count()))
    lib.swb_dist_seed(self.trainer, _ptr(allrecs), allrecs.shape[0])

  def merge_batch(self, n: int) -> int:
    if self.merge_loop == "rank0":
      raise RuntimeError('merge_loop="rank0" trains in one piece: use train() / train_quiet()')
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
    if self.merge_loop == "rank0":
      got, err = 0, None
      if self.rank == 0:
        try:
          got = super().train_quiet()
        except RuntimeError as e:  # the other ranks are waiting for the result: tell them before raising
          err = e
      self._broadcast_result(failed=err is not None)
      if err is not None:
        raise err
      return got if self.rank == 0 else self._bcast[0].shape[0]
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

  # ---- results on the ranks that did not run the loop (merge_loop == "rank0"): the broadcast copies
  @property
  def num_merges(self) -> int:
    return self._bcast[0].shape[0] if self._bcast is not None else BPETrainer.num_merges.fget(self)

  def merges_array(self) -> np.ndarray:
    return self._bcast[0].copy() if self._bcast is not None else super().merges_array()

  def byte_map(self) -> np.ndarray:
    return self._bcast[1].copy() if self._bcast is not None else super().byte_map()

  def encoder(self):
    if self._bcast is not None:
      from .trainer import BPEEncoder
      return BPEEncoder(self._bcast[0], self._bcast[1])
    return super().encoder()

  def token_freq(self) -> np.ndarray:
    """Global token histogram: sum of the ranks' shard histograms."""
    if self._bcast is not None:
      return self._bcast[2].copy()
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


# ------------------------------------------------------------------------------------------------ document-parallel encode
def split_on_newline(nbytes_total: int, rank: int, world: int, peek) -> tuple[int, int]:
  """Byte range [lo, hi) of rank `rank` when a text of nbytes_total bytes is split into `world` ranges cut AFTER a
  newline (documents = lines; no word straddles two ranks). `peek(lo, hi)` returns the bytes [lo, hi) of the text
  as a uint8 array; only a window around each nominal cut is read."""
  def cut(k: int) -> int:
    if k <= 0:
      return 0
    if k >= world:
      return nbytes_total
    pos = nbytes_total * k // world
    WIN = 1 << 16
    while pos < nbytes_total:
      w = peek(pos, min(nbytes_total, pos + WIN))
      hit = np.nonzero(w == 10)[0]
      if hit.size:
        return pos + int(hit[0]) + 1
      pos += w.size
    return nbytes_total
  return cut(rank), cut(rank + 1)


def gather_token_offsets(n_local_tokens: int, group=None, device: torch.device | None = None) -> tuple[int, int, list[int]]:
  """The one collective of the document-parallel encode (reference base.py:10-36 applied per document, no exchange of
  text or ids): every rank learns all token counts. Returns (this rank's offset in the global id stream, total tokens,
  the per-rank counts)."""
  device = device if device is not None else torch.device("cpu")
  world = dist.get_world_size(group)
  mine = torch.tensor([int(n_local_tokens)], dtype=torch.int64, device=device)
  out = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
  dist.all_gather(out, mine, group=group)
  counts = [int(x.item()) for x in out]
  r = dist.get_rank(group)
  return sum(counts[:r]), sum(counts), counts


class DistributedBPEEncoder:
  """Document-parallel encoding over the ranks of `group` (BASELINE config 4): rank r encodes its own byte range of
  the text (cut after a newline), the token counts are all-gathered so that every rank knows where its ids sit in the
  global stream. `encoder` is this rank's BPEEncoder (or any object with the same encode / encode_device methods:
  the CPU tests plug in the oracle's encoder)."""

  def __init__(self, encoder, group=None, device: torch.device | None = None):
    self.enc = encoder
    self.group = group
    self.device = device if device is not None else torch.device("cpu")
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)

  def encode_range(self, text: np.ndarray) -> tuple[np.ndarray, int, int]:
    """`text` = this rank's byte range (host). Returns (ids of the range, global token offset, total tokens)."""
    ids = self.enc.encode(text)
    off, total, _ = gather_token_offsets(len(ids), self.group, self.device)
    return ids, off, total

  def encode_whole(self, text: np.ndarray) -> tuple[np.ndarray, int, int]:
    """`text` = the WHOLE text, present on every rank; each rank encodes its newline-cut share."""
    a = np.ascontiguousarray(text, dtype=np.uint8)
    lo, hi = split_on_newline(a.size, self.rank, self.world, lambda x, y: a[x:y])
    return self.encode_range(a[lo:hi])
