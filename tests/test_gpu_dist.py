"""GPU, 2+ devices: the multi-GPU trainer -- merge loop on rank 0 + broadcast (default), replicated, the sharded merge loop with the NCCL
exchange issued from inside the library, and the Python-driven torch.distributed exchange -- must reproduce the
single-GPU / oracle merge list bit for bit.
Skipped on a single-GPU box (the round-end -m gpu run); exercised with `gpurun --gpus 2`."""
import os
import socket
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, mode, out_dir):
  native = mode != "python"
  sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(rank)
  from shredword_b200.cbase import lib
  lib.swb_set_device(rank)
  dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from shredword_b200.distributed import DistributedBPETrainer
  kw = cases.kwargs(name)
  t = DistributedBPETrainer(**kw, device=torch.device("cuda", rank), native=native, merge_loop=("sharded" if mode == "python" else mode))
  assert t.sharded == (mode in ("sharded", "python"))
  t.load_buffer(cases.corpus(name))
  n = t.train_quiet()
  np.save(os.path.join(out_dir, f"merges_{mode}_{rank}.npy"), t.merges_array())
  np.save(os.path.join(out_dir, f"freq_{mode}_{rank}.npy"), t.token_freq())
  t.save(os.path.join(out_dir, f"m_{mode}.model"), os.path.join(out_dir, f"m_{mode}.vocab"))
  st = t.stats()
  assert n == len(t.merges_array())
  if mode == "sharded":
    assert st["collectives"] >= n
  if mode == "replicated" or (mode == "rank0" and rank == 0):
    assert st["collectives"] == 0 and (n == 0 or st["resident_local_merges"] + st["resident_grid_merges"] > 0 or st["long_words"] > 0)
  if mode == "rank0" and rank != 0:
    assert st["merge_launches"] == 0 and st["resident_local_merges"] + st["resident_grid_merges"] == 0  # this rank ran no merge
    ids = t.encoder().encode(cases.corpus(name))  # ... and still has the result: an encoder from the broadcast merges
    np.save(os.path.join(out_dir, f"ids_{mode}_{rank}.npy"), ids)
  dist.barrier()
  t.destroy()
  lib.swb_dist_shutdown()
  dist.destroy_process_group()


@pytest.mark.parametrize("name,mode", [("ascii_ties", "rank0"), ("multi_unk97", "rank0"), ("long_words", "rank0"),
                                       ("ascii_ties", "replicated"), ("long_words", "replicated"),
                                       ("ascii_ties", "sharded"), ("multi_unk97", "sharded"), ("long_words", "sharded"), ("negative_unk", "sharded"),
                                       ("ascii_ties", "python"), ("multi_unk97", "python")])
def test_two_gpus_match_oracle(name, mode, product, oracle_mod, tmp_path):
  import torch
  import torch.multiprocessing as mp
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  kw = cases.kwargs(name)
  o = oracle_mod.Oracle(kw["target_vocab_size"], kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw["min_pair_freq"])
  o.load_buffer(cases.corpus(name)); o.train()
  o.save(str(tmp_path / "o.model"), str(tmp_path / "o.vocab"))
  mp.spawn(_worker, args=(2, port, name, mode, str(tmp_path)), nprocs=2, join=True)
  for r in range(2):
    assert np.array_equal(np.load(tmp_path / f"merges_{mode}_{r}.npy"), o.merges), f"rank {r}"
    assert np.array_equal(np.load(tmp_path / f"freq_{mode}_{r}.npy"), o.token_freq()), f"rank {r}"
  if mode == "rank0":
    ids = np.load(tmp_path / f"ids_{mode}_1.npy")
    assert np.array_equal(ids, oracle_mod.encode(o.merges, o.byte_map(kw.get("unk_id", 0)), cases.corpus(name)))
  assert (tmp_path / f"m_{mode}.model").read_bytes() == (tmp_path / "o.model").read_bytes()
  assert (tmp_path / f"m_{mode}.vocab").read_bytes() == (tmp_path / "o.vocab").read_bytes()


def _shard_worker(rank, world, port, sharded, out_dir):
  sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(rank)
  from shredword_b200.cbase import lib
  lib.swb_set_device(rank)
  dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from shredword_b200.distributed import DistributedBPETrainer
  data = np.load(os.path.join(out_dir, "corpus.npy"))
  cuts = np.load(os.path.join(out_dir, "cuts.npy"))
  if not sharded:
    os.environ["SWB_LOAD_PIECE"] = "65536"  # rank 0's host buffer goes through the pipelined copy + tokenise path
  t = DistributedBPETrainer(1500, min_pair_freq=5, device=torch.device("cuda", rank), merge_loop=("sharded" if sharded else "rank0"))
  piece = data[cuts[rank]:cuts[rank + 1]]
  if rank == 0:
    t.load_shard(piece, int(cuts[rank]))                                   # host buffer
  else:
    t.load_shard(torch.from_numpy(piece.copy()).cuda(), int(cuts[rank]))   # device tensor
  boff, by, _, _, cnt = t.words()
  np.save(os.path.join(out_dir, f"wbytes_{rank}.npy"), by); np.save(os.path.join(out_dir, f"wcnt_{rank}.npy"), cnt)
  n = t.train_quiet()
  st = t.stats()
  assert st["collectives"] == (3 if not sharded else st["collectives"]) and st["collectives"] >= 3  # the word-table exchange
  np.save(os.path.join(out_dir, f"smerges_{rank}.npy"), t.merges_array())
  np.save(os.path.join(out_dir, f"sfreq_{rank}.npy"), t.token_freq())
  dist.barrier()
  t.destroy()
  lib.swb_dist_shutdown()
  dist.destroy_process_group()


@pytest.mark.parametrize("sharded", [False, True])
def test_range_split_load_matches_oracle(sharded, product, oracle_mod, tmp_path):
  """Every rank tokenises only its byte range; the exchanged + merged word table must be the global one
  (same words, same order, summed counts, first occurrence = global minimum)."""
  import torch
  import torch.multiprocessing as mp
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  from shredword_b200 import synth
  data = synth.corpus_bytes(synth.small_spec(4_000_000, 60_000, 77, "multi"))
  mid = data.size // 2
  while data[mid] not in b" \n":
    mid += 1
  np.save(tmp_path / "corpus.npy", data); np.save(tmp_path / "cuts.npy", np.array([0, mid + 1, data.size]))
  o = oracle_mod.Oracle(1500, 0, 0.995, 5); o.load_buffer(data)
  _, oby, _, _, ocnt = o.words()
  o.train()
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  mp.spawn(_shard_worker, args=(2, port, sharded, str(tmp_path)), nprocs=2, join=True)
  for r in range(2):
    assert np.array_equal(np.load(tmp_path / f"wbytes_{r}.npy"), oby) and np.array_equal(np.load(tmp_path / f"wcnt_{r}.npy"), ocnt), f"word table, rank {r}"
    assert np.array_equal(np.load(tmp_path / f"smerges_{r}.npy"), o.merges), f"rank {r}"
    assert np.array_equal(np.load(tmp_path / f"sfreq_{r}.npy"), o.token_freq()), f"rank {r}"
