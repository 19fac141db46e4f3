"""CPU: the multi-rank merge loop. The pair table + heap replica and the record reduction are host code of
the product (swb_dist_*), exercised here without a GPU: this rank's kernels are replaced by the oracle's
shard functions (a CPU stand-in that emits the same records), the exchange runs over gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleLocalOps:
  def __init__(self, o, rank, world):
    self.o, self.rank, self.world = o, rank, world

  def count(self):
    return np.ascontiguousarray(self.o.shard_count(self.rank, self.world))

  def merge(self, a, b, new_id):
    return np.ascontiguousarray(self.o.shard_merge(self.rank, self.world, a, b, new_id))


def test_single_process_replica_matches_oracle(product, oracle_mod):
  """world_size 1 through the same swb_dist_* path (no process group needed)."""
  import ctypes
  from shredword_b200.cbase import lib
  from shredword_b200.trainer import _ptr
  for name in ("ascii_ties", "multi_unk97", "self_pairs", "negative_unk"):
    kw = cases.kwargs(name)
    full = oracle_mod.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw["min_pair_freq"])
    full.load_buffer(cases.corpus(name)); n_full = full.train()
    sh = oracle_mod.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw["min_pair_freq"])
    sh.load_buffer(cases.corpus(name))
    t = product.BPETrainer(**kw)
    # three emulated ranks in one process: concatenate their records, reduce, apply
    R = 3
    recs = np.ascontiguousarray(np.concatenate([sh.shard_count(r, R) for r in range(R)]))
    n = lib.swb_dist_reduce_records(_ptr(recs), recs.shape[0])
    lib.swb_dist_seed(t.trainer, _ptr(recs), n)
    a, b, nid = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    done = 0
    while done < kw["target_vocab_size"] - 256 and lib.swb_dist_next_merge(t.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid)):
      parts = [sh.shard_merge(r, R, a.value, b.value, nid.value) for r in range(R)]
      recs = np.ascontiguousarray(np.concatenate(parts)) if sum(len(p) for p in parts) else np.zeros((0, 4), np.int64)
      n = lib.swb_dist_reduce_records(_ptr(recs), recs.shape[0]) if recs.shape[0] else 0
      lib.swb_dist_apply(t.trainer, _ptr(recs), n)
      done += 1
    assert done == n_full, name
    assert np.array_equal(full.merges, t.merges_array()), name


def _worker(rank, world, port, name, out_dir):
  sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch.distributed as dist
  import oracle as O
  from shredword_b200.distributed import DistributedBPETrainer
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  kw = cases.kwargs(name)
  o = O.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw["min_pair_freq"])
  o.load_buffer(cases.corpus(name))
  t = DistributedBPETrainer(**kw, local_ops=OracleLocalOps(o, rank, world))
  n = t.train_quiet()
  np.save(os.path.join(out_dir, f"merges_{rank}.npy"), t.merges_array())
  # every rank holds the same replica: same heap size, same merge count
  sizes = [None] * world
  dist.all_gather_object(sizes, (n, int(t.trainer.contents.heap.size), t.collectives))
  assert len(set(s[:2] for s in sizes)) == 1, sizes
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ascii_ties", "multi_ties"])
def test_gloo_world2_matches_oracle(name, product, oracle_mod, tmp_path):
  import socket
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  kw = cases.kwargs(name)
  full = oracle_mod.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw["min_pair_freq"])
  full.load_buffer(cases.corpus(name)); full.train()
  mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
  for r in range(2):
    assert np.array_equal(np.load(tmp_path / f"merges_{r}.npy"), full.merges), f"rank {r}"


@pytest.mark.parametrize("name", ["ascii_ties", "multi_ties", "self_pairs", "ref_fixture_f5"])
def test_look_ahead_is_exact_and_leaves_the_heap_alone(name, product, oracle_mod):
  """HostCore::peek_next (what the resident kernel's hints are made of): entry i of the look-ahead taken while merge
  j is pending names merge j+1+i. Whenever the merges j .. j+i then push nothing at or above the quoted frequency and
  the pair still has that frequency when its turn comes, the heap must pop exactly that pair -- ties included.
  The heap array must be bit-identical before and after the look-ahead."""
  import ctypes
  from shredword_b200.cbase import lib
  from shredword_b200.trainer import _ptr
  kw = cases.kwargs(name)
  minf = kw["min_pair_freq"]
  DEPTH = 6  # (the resident loop keeps a list of up to 8 and reuses its tail for the following merges)
  sh = oracle_mod.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), minf)
  sh.load_buffer(cases.corpus(name))
  t = product.BPETrainer(**kw)
  recs = np.ascontiguousarray(sh.shard_count(0, 1))
  n = lib.swb_dist_reduce_records(_ptr(recs), recs.shape[0])
  lib.swb_dist_seed(t.trainer, _ptr(recs), n)
  freq = {(int(r[0]), int(r[1])): int(r[2]) for r in recs[:n]}  # the test's own frequency table

  def heap_bytes():
    h = t.trainer.contents.heap
    raw = np.frombuffer(ctypes.string_at(h.data, h.size * 24), dtype=np.uint8).reshape(-1, 24)
    return raw[:, :20].tobytes()  # {first, second, freq, version}; bytes 20..23 are struct padding

  a, b, nid = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
  out = np.zeros(3 * DEPTH, dtype=np.int64)
  claims = []      # (merge number it names, pair, frequency, number of the merge that was pending)
  maxpush = {}     # merge number -> largest frequency it made the heap push
  expect = {}      # merge number -> pair, for claims whose conditions held
  made = [0] * DEPTH
  confirmed = [0] * DEPTH
  tie_claims = 0
  j = 0
  target = 2 * (kw["target_vocab_size"] - 256)
  while j < target and lib.swb_dist_next_merge(t.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid)):
    j += 1  # merge j is pending
    if j in expect:
      for depth, pair in expect[j]:
        assert (a.value, b.value) == pair, f"{name}: look-ahead (depth {depth}) said {pair} for merge {j}, the heap popped {(a.value, b.value)}"
        confirmed[depth] += 1
    check_heap = j % 8 == 0  # (copying the whole heap out every merge would dominate the test)
    before = heap_bytes() if check_heap else None
    got = lib.swb_dist_peek_list(t.trainer, _ptr(out), DEPTH)
    assert not check_heap or heap_bytes() == before, "look-ahead changed the heap"
    fs = [int(out[3 * i + 2]) for i in range(got)]
    assert fs == sorted(fs, reverse=True)
    tie_claims += len(set(fs)) < len(fs)
    for i in range(got):
      claims.append((j + 1 + i, (int(out[3 * i]), int(out[3 * i + 1])), fs[i], j, i))
      made[i] += 1
    recs = np.ascontiguousarray(sh.shard_merge(0, 1, a.value, b.value, nid.value))
    n = lib.swb_dist_reduce_records(_ptr(recs), recs.shape[0]) if recs.shape[0] else 0
    mp = 0
    for r in recs[:n]:
      k = (int(r[0]), int(r[1]))
      if k == (a.value, b.value):
        continue
      nw = max(freq.get(k, 0) + int(r[2]), 0)
      freq[k] = nw
      if nw >= minf:
        mp = max(mp, nw)
    freq[(a.value, b.value)] = 0
    maxpush[j] = mp
    rest = []
    for c in claims:
      tgt, pair, f, base, depth = c
      if tgt != j + 1:
        rest.append(c)
      elif max(maxpush[m] for m in range(base, j + 1)) < f and freq.get(pair, 0) == f:
        expect.setdefault(tgt, []).append((depth, pair))
    claims = rest
    lib.swb_dist_apply(t.trainer, _ptr(recs), n)
  print(f"\n[{name}] merges={j} claims per depth={made} confirmed per depth={confirmed} look-aheads with equal frequencies={tie_claims}")
  if name in ("ascii_ties", "multi_ties"):  # (the small cases have heaps too small for a look-ahead: nothing may be claimed wrongly, that is all)
    assert confirmed[0] > 1000 and confirmed[1] > 500 and confirmed[2] > 250 and tie_claims > 100, (made, confirmed, tie_claims)


@pytest.mark.parametrize("seed,span", [(1, 3), (2, 8), (3, 40), (4, 1000)])
def test_look_ahead_equals_pop_order_on_random_heaps(seed, span, product):
  """The claim behind the look-ahead, tested on its own: the reference heap's pop order is "frequency descending, then
  pre-order position in the tree". Random heaps with few distinct frequencies (span) and many dead entries; with nothing
  pushed in between, the next 8 pairs named by the look-ahead must be exactly the next 8 pops that survive the version test."""
  import ctypes
  from shredword_b200.cbase import lib
  from shredword_b200.trainer import _ptr
  rng = np.random.default_rng(seed)
  minf = 10
  P = 6000
  t = product.BPETrainer(target_vocab_size=100000, min_pair_freq=minf)
  # count records: pair (i, j) with frequency minf + random small offset; key = first-touch order
  first = rng.integers(0, 200, size=P).astype(np.int64)
  second = np.arange(P, dtype=np.int64) + 300  # distinct pairs
  freq = (minf + 1 + rng.integers(0, span, size=P)).astype(np.int64)
  freq[rng.random(P) < 0.7] = minf  # filler at the bottom, so that the tail of the array is usually below the entries at the top
  recs = np.ascontiguousarray(np.stack([first, second, freq, rng.permutation(P).astype(np.int64)], axis=1))
  lib.swb_dist_seed(t.trainer, _ptr(recs), P)
  a, b, nid = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
  empty = np.zeros((1, 4), dtype=np.int64)
  # phase B: kill entries -- a zero net delta re-pushes the pair with a new version (reference bpe.cpp:512-515), the old entry stays as a dead one
  for _ in range(40):
    assert lib.swb_dist_next_merge(t.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid))
    pick = rng.choice(P, size=60, replace=False)
    touch = np.ascontiguousarray(np.stack([first[pick], second[pick], np.zeros(60, np.int64), np.arange(60, dtype=np.int64)], axis=1))
    lib.swb_dist_apply(t.trainer, _ptr(touch), 60)
  # phase C: nothing is pushed any more; look-ahead vs the real pops
  out = np.zeros(3 * 8, dtype=np.int64)
  checked = 0
  for _ in range(60):
    if not lib.swb_dist_next_merge(t.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid)):
      break
    got = lib.swb_dist_peek_list(t.trainer, _ptr(out), 8)
    lib.swb_dist_apply(t.trainer, _ptr(empty), 0)
    # (fewer than 8 entries come back when an element at the tail of the array is as frequent as the entries at the top --
    # then the pops would move an element of the top region itself to the root and no statement is made; common for span = 3)
    for i in range(got):
      assert lib.swb_dist_next_merge(t.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid))
      assert (a.value, b.value) == (int(out[3 * i]), int(out[3 * i + 1])), f"entry {i} of the look-ahead is not the pair the heap popped"
      lib.swb_dist_apply(t.trainer, _ptr(empty), 0)
      checked += 1
  assert checked >= (400 if span >= 40 else 100), checked


@pytest.mark.parametrize("block", range(4))
def test_random_small_corpora_replica_matches_oracle(block, product, oracle_mod):
  """Differential fuzzing of the host replica (pair table, exact heap, delta-map order, clamping, look-ahead) against the
  oracle on tiny tie-heavy corpora: small alphabets, low min_pair_freq, unk ids inside and outside 0..255, emulated ranks."""
  import ctypes
  from shredword_b200.cbase import lib
  from shredword_b200.trainer import _ptr
  a, b, nid = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
  out = np.zeros(3 * 3, dtype=np.int64)
  for seed in range(block * 8, block * 8 + 8):
    rng = np.random.default_rng(1000 + seed)
    alphabet = np.frombuffer(b"abcdefgh"[: int(rng.integers(2, 9))], dtype=np.uint8)
    n_words = int(rng.integers(50, 1500))
    words = [bytes(rng.choice(alphabet, size=int(rng.integers(1, 8)))) for _ in range(n_words)]
    seps = [b" ", b"\n", b"\t", b"  ", b"\r\n"]
    data = b"".join(w + seps[int(rng.integers(0, len(seps)))] for w in words)
    kw = dict(target_vocab_size=int(rng.integers(257, 420)), min_pair_freq=int(rng.integers(1, 5)),
              unk_id=int(rng.choice([0, 0, 97, 98, -1, -7, 300])), character_coverage=float(rng.choice([0.995, 0.9, 0.6])))
    full = oracle_mod.Oracle(kw["target_vocab_size"], kw["unk_id"], kw["character_coverage"], kw["min_pair_freq"])
    full.load_buffer(data); n_full = full.train()
    sh = oracle_mod.Oracle(kw["target_vocab_size"], kw["unk_id"], kw["character_coverage"], kw["min_pair_freq"])
    sh.load_buffer(data)
    t = product.BPETrainer(**kw)
    R = int(rng.integers(1, 4))
    recs = np.ascontiguousarray(np.concatenate([sh.shard_count(r, R) for r in range(R)]))
    n = lib.swb_dist_reduce_records(_ptr(recs), recs.shape[0]) if recs.shape[0] else 0
    lib.swb_dist_seed(t.trainer, _ptr(recs), n)
    done = 0
    while done < kw["target_vocab_size"] - 256 and lib.swb_dist_next_merge(t.trainer, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nid)):
      lib.swb_dist_peek_list(t.trainer, _ptr(out), 3)  # (must not disturb anything)
      parts = [sh.shard_merge(r, R, a.value, b.value, nid.value) for r in range(R)]
      recs = np.ascontiguousarray(np.concatenate(parts)) if sum(len(p) for p in parts) else np.zeros((0, 4), np.int64)
      n = lib.swb_dist_reduce_records(_ptr(recs), recs.shape[0]) if recs.shape[0] else 0
      lib.swb_dist_apply(t.trainer, _ptr(recs) if n else _ptr(np.zeros((1, 4), np.int64)), n)
      done += 1
    assert done == n_full, (seed, kw, done, n_full)
    assert np.array_equal(full.merges, t.merges_array()), (seed, kw)


class OracleEncoder:
  """Stands in for this rank's BPEEncoder in the CPU test of the document-parallel encode (same `encode` method)."""

  def __init__(self, O, merges, byte_map):
    self.O, self.merges, self.byte_map = O, merges, byte_map

  def encode(self, text):
    return self.O.encode(self.merges, self.byte_map, bytes(text))


def _encode_worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch.distributed as dist
  import oracle as O
  from shredword_b200.distributed import DistributedBPEEncoder
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  merges = np.load(os.path.join(out_dir, "merges.npy"))
  byte_map = np.load(os.path.join(out_dir, "byte_map.npy"))
  text = np.frombuffer(open(os.path.join(out_dir, "text.bin"), "rb").read(), dtype=np.uint8)
  enc = DistributedBPEEncoder(OracleEncoder(O, merges, byte_map))
  ids, off, total = enc.encode_whole(text)
  np.save(os.path.join(out_dir, f"ids_{rank}.npy"), np.asarray(ids, dtype=np.int32))
  np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([off, total], dtype=np.int64))
  dist.barrier()
  dist.destroy_process_group()


def test_gloo_world2_document_parallel_encode(product, oracle_mod, tmp_path):
  """Document-parallel encode (BASELINE config 4) over gloo, world_size 2: every rank encodes its newline-cut share, the
  token counts are all-gathered; the shares placed at the gathered offsets are the encoding of the whole text."""
  import socket
  name = "multi_ties"
  kw = cases.kwargs(name)
  o = oracle_mod.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw["min_pair_freq"])
  data = cases.corpus(name)
  o.load_buffer(data); o.train()
  byte_map = np.asarray(o.byte_map(kw.get("unk_id", 0)), dtype=np.int32)
  np.save(tmp_path / "merges.npy", o.merges); np.save(tmp_path / "byte_map.npy", byte_map)
  (tmp_path / "text.bin").write_bytes(bytes(data))
  want = np.asarray(oracle_mod.encode(o.merges, byte_map, bytes(data)), dtype=np.int32)
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  mp.spawn(_encode_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  got = np.full(want.size, -7, dtype=np.int32)
  for r in range(2):
    ids = np.load(tmp_path / f"ids_{r}.npy"); off, total = np.load(tmp_path / f"meta_{r}.npy")
    assert total == want.size and len(ids) > 0
    got[off: off + len(ids)] = ids
  assert np.array_equal(got, want)
