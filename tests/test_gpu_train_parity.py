"""GPU parity: the CUDA trainer (through the C-ABI) against the CPU oracle, bit for bit.

Checked per case: the unique-word table (bytes, counts, order), the kept-byte map, the exact heap
array after the initial count (the heap replica must match entry for entry, not just as a set), the
ordered merge list, the final segmentation of every word, and the bytes of the .model / .vocab files."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _mk(T, kw):
  return T.BPETrainer(**kw)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_train_matches_oracle(name, product, oracle_mod, tmp_path):
  data = cases.corpus(name)
  kw = cases.kwargs(name)
  path = tmp_path / "corpus.txt"
  path.write_bytes(data)

  o = oracle_mod.Oracle(kw.get("target_vocab_size", 8192), kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw.get("min_pair_freq", 2000))
  o.load_corpus(str(path))
  t = _mk(product, kw)
  t.load_corpus(str(path))

  # word table
  ob, oby, _, osy, oc = o.words()
  tb, tby, _, tsy, tc = t.words()
  assert np.array_equal(ob, tb) and np.array_equal(oby, tby), "unique words / reference word order differ"
  assert np.array_equal(oc, tc), "word counts differ"
  assert np.array_equal(osy, tsy), "initial symbols differ"
  assert np.array_equal(o.byte_map(kw.get("unk_id", 0)), t.byte_map())

  # initial count: exact heap array
  o.init(); t.init()
  of, os_, ofr, ov = o.heap()
  h = t.trainer.contents.heap
  tf = np.array([h.data[i].key.first for i in range(h.size)], dtype=np.int32)
  ts = np.array([h.data[i].key.second for i in range(h.size)], dtype=np.int32)
  tfr = np.array([h.data[i].freq for i in range(h.size)], dtype=np.uint64)
  assert h.size == len(of)
  assert np.array_equal(of, tf) and np.array_equal(os_, ts) and np.array_equal(ofr, tfr), "heap array after count differs"

  # full training through bpe_train (which re-inits)
  n_o = o.train()
  n_t = t.train_quiet()
  assert n_o == n_t
  assert np.array_equal(o.merges, t.merges_array()), "merge lists differ"
  _, _, oso, osy, _ = o.words()
  _, _, tso, tsy, _ = t.words()
  assert np.array_equal(oso, tso) and np.array_equal(osy, tsy), "final segmentation differs"
  assert np.array_equal(o.token_freq(), t.token_freq())

  o.save(str(tmp_path / "o.model"), str(tmp_path / "o.vocab"))
  t.save(str(tmp_path / "t.model"), str(tmp_path / "t.vocab"))
  assert (tmp_path / "o.model").read_bytes() == (tmp_path / "t.model").read_bytes()
  assert (tmp_path / "o.vocab").read_bytes() == (tmp_path / "t.vocab").read_bytes()
  assert os.path.getsize(tmp_path / "t.model") == 12 * n_t  # reference test/bpe_test.cpp:262-270
  # and against the bytes the UNMODIFIED reference wrote for this corpus (tests/golden/make_golden.py)
  gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
  if os.path.exists(gold + ".model"):
    assert (tmp_path / "t.model").read_bytes() == open(gold + ".model", "rb").read()
    assert (tmp_path / "t.vocab").read_bytes() == open(gold + ".vocab", "rb").read()


def test_merge_batch_stepwise_matches(product, oracle_mod):
  """bpe_init + bpe_merge_batch(k) in uneven steps == the same prefix of the merge list."""
  data = cases.corpus("ascii_ties")
  o = oracle_mod.Oracle(2000, 0, 0.995, 5); o.load_buffer(data); o.init()
  t = product.BPETrainer(2000, min_pair_freq=5); t.load_buffer(data); t.init()
  for step in (1, 2, 7, 50, 1, 140):
    assert o.merge_batch(step) == t.merge_batch(step)
    assert np.array_equal(o.merges, t.merges_array())
  # re-init on the partially merged corpus (recount with ids >= 256 present), then continue
  o.init(); t.init()
  assert o.merge_batch(100) == t.merge_batch(100)
  assert np.array_equal(o.merges, t.merges_array())


def test_load_buffer_and_device_equal_file(product, tmp_path):
  import torch
  data = cases.corpus("multi_ties")
  p = tmp_path / "c.txt"; p.write_bytes(data)
  a = product.BPETrainer(700, min_pair_freq=5); a.load_corpus(str(p)); a.train_quiet()
  b = product.BPETrainer(700, min_pair_freq=5); b.load_buffer(data); b.train_quiet()
  d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
  c = product.BPETrainer(700, min_pair_freq=5); c.load_device(d.data_ptr(), d.numel()); c.train_quiet()
  assert np.array_equal(a.merges_array(), b.merges_array()) and np.array_equal(a.merges_array(), c.merges_array())


def test_errors(product, tmp_path):
  t = product.BPETrainer(300, min_pair_freq=2)
  with pytest.raises(IOError):
    t.load_corpus(str(tmp_path / "missing.txt"))       # reference: -1 -> IOError (trainer.py:19-20)
  with pytest.raises(IOError):
    t.load_buffer(b"abc\x00def ghi")                    # NUL is outside the parity domain: rejected loudly
  assert t.train_quiet() == 0                          # nothing loaded -> no merges


def test_vocab_and_merges_properties(product):
  t = product.BPETrainer(300, min_pair_freq=2)
  t.load_buffer(cases.corpus("ref_fixture")); t.train_quiet()
  m = t.merges
  assert len(m) == t.num_merges and all(nid == 256 + i for i, (_, _, nid) in enumerate(m))
  v = t.vocab
  assert len(v) == 256 + len(m) and v[ord("a")] == b"a"
  for a, b, nid in m:
    assert v[nid] == v[a] + v[b]
  assert t.special_tokens[0] == ("<UNK>", 0)


@pytest.mark.parametrize("piece", [4096, 65536, 1 << 20])
def test_pipelined_host_load_equals_oracle(piece, product, oracle_mod, monkeypatch):
  """swb_load_corpus_buffer on large buffers copies the corpus in pieces and tokenises each piece as it lands
  (words that straddle a piece boundary, a word longer than a piece, pieces without any delimiter). The word
  table, its order and the merges must not depend on the piece size."""
  rng = np.random.default_rng(9)
  giant = bytes(rng.choice(np.frombuffer(b"xyz", dtype=np.uint8), size=3 * 4096 + 77))  # longer than the smallest piece
  data = cases.corpus("multi_ties")[:1_500_000]
  data = data[: len(data) // 2] + b" " + giant + b" " + giant + b"\n" + data[len(data) // 2:]
  kw = dict(target_vocab_size=600, min_pair_freq=5)
  o = oracle_mod.Oracle(600, 0, 0.995, 5)
  assert o.load_buffer(data) == 0
  monkeypatch.setenv("SWB_LOAD_PIECE", str(piece))
  t = product.BPETrainer(**kw)
  t.load_buffer(data)
  ob, oby, _, osy, oc = o.words()
  tb, tby, _, tsy, tc = t.words()
  assert np.array_equal(ob, tb) and np.array_equal(oby, tby), "unique words / reference word order differ"
  assert np.array_equal(oc, tc), "word counts differ"
  assert o.train() == t.train_quiet()
  assert np.array_equal(o.merges, t.merges_array())
  assert t.stats()["kernel_launches"] > len(data) // piece  # (one tokeniser launch per piece: the pipelined path ran)


def test_resident_loop_uses_hints_and_the_initial_pair_index(product, oracle_mod):
  """The single-GPU product path is the resident cluster kernel: on a tie-heavy corpus most merges must start from a
  look-ahead hint (checked by the device, cross-checked by the host) and pairs of two initial symbols must run on the
  leader cluster through their occurrence index -- with the merge list still equal to the oracle's."""
  data = cases.corpus("ascii_ties")
  kw = cases.kwargs("ascii_ties")
  o = oracle_mod.Oracle(kw["target_vocab_size"], 0, 0.995, kw["min_pair_freq"]); o.load_buffer(data); n_o = o.train()
  t = product.BPETrainer(**kw); t.load_buffer(data); n_t = t.train_quiet()
  assert n_o == n_t and np.array_equal(o.merges, t.merges_array())
  st = t.stats()
  assert st["resident_local_merges"] + st["resident_grid_merges"] == n_t, "the resident kernel did not serve every merge"
  assert st["hints_sent"] > n_t // 2 and st["hints_taken"] > n_t // 4, (st["hints_sent"], st["hints_taken"], n_t)
  assert st["resident_local_merges"] > st["resident_grid_merges"], "initial-symbol pairs should mostly run on the leader cluster"
