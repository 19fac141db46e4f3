"""GPU parity of the rank-ordered encoder (through the C-ABI) against the CPU oracle, bit for bit.

The reference has no encoder, so beyond oracle equality the encoder is pinned the way SURVEY.md 8(c)
prescribes: encoding the training corpus must reproduce the trainer's own final segmentation (token
histogram == the .vocab frequency column; every training word == its final chain)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ENC_CASES = ["ref_fixture", "ascii_ties", "multi_ties", "multi_unk97", "unk_enters_by_delta", "long_words", "ragged",
             "self_pairs", "single_word", "only_delims", "empty", "negative_unk", "config1_10MB"]


@pytest.mark.parametrize("name", ENC_CASES)
def test_encode_matches_oracle_and_trainer(name, product, oracle_mod):
  data = cases.corpus(name)
  kw = cases.kwargs(name)
  t = product.BPETrainer(**kw)
  t.load_buffer(data)
  t.train_quiet()
  merges, bmap = t.merges_array(), t.byte_map()
  enc = t.encoder()
  ids, wn = enc.encode(data, with_word_counts=True)
  oids, own = oracle_mod.encode(merges, bmap, data, with_word_counts=True)
  assert np.array_equal(own, wn), "per-word token counts differ"
  assert np.array_equal(oids, ids), "token ids differ"
  assert enc.kernel_launches > 0 or len(data.split()) == 0

  # pin 1: token histogram of the encoded training corpus == the trainer's .vocab frequency column
  T = 256 + len(merges)
  valid = ids[(ids >= 0) & (ids < T)]
  hist = np.bincount(valid, minlength=T).astype(np.uint64)
  assert np.array_equal(hist, t.token_freq())

  # pin 2: every unique training word encodes to the trainer's final chain of that word
  boff, by, soff, sy, _ = t.words()
  W = len(boff) - 1
  if W:
    pick = np.unique(np.linspace(0, W - 1, min(W, 5000)).astype(np.int64))
    text = b" ".join(by[boff[w]:boff[w + 1]].tobytes() for w in pick)
    wids, wcnt = enc.encode(text, with_word_counts=True)
    expect = np.concatenate([sy[soff[w]:soff[w + 1]] for w in pick])
    assert np.array_equal(wcnt, np.array([soff[w + 1] - soff[w] for w in pick], dtype=np.uint32))
    assert np.array_equal(wids, expect)


def test_encode_unseen_text_and_roundtrip(product, oracle_mod, tmp_path):
  from shredword_b200 import synth
  train = cases.corpus("ascii_ties")
  t = product.BPETrainer(1200, min_pair_freq=5)
  t.load_buffer(train); t.train_quiet()
  t.save(str(tmp_path / "m.model"), str(tmp_path / "m.vocab"))
  # identity byte map (what a bare .model file gives): decode(encode(x)) == x without its delimiters
  enc = product.BPEEncoder.from_model_file(str(tmp_path / "m.model"))
  other = bytes(synth.corpus_bytes(synth.small_spec(3_000_000, 80_000, 99, "multi")))
  ids = enc.encode(other)
  oids = oracle_mod.encode(t.merges_array(), np.arange(256, dtype=np.int32), other)
  assert np.array_equal(ids, oids)
  plain = other.translate(None, b" \t\r\n")
  assert enc.decode(ids) == plain
  assert oracle_mod.decode(t.merges_array(), ids) == plain
  # idempotence of the segmentation: encoding the decoded words one by one gives the same ids
  assert len(ids) < len(plain)


def test_encode_device_pointers(product):
  import torch
  data = cases.corpus("multi_ties")
  t = product.BPETrainer(800, min_pair_freq=5)
  t.load_buffer(data); t.train_quiet()
  enc = t.encoder()
  host_ids = enc.encode(data)
  d_text = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
  d_out = torch.empty(d_text.numel(), dtype=torch.int32, device="cuda")
  n = enc.encode_device(d_text.data_ptr(), d_text.numel(), d_out.data_ptr(), d_out.numel())
  assert n == len(host_ids)
  assert np.array_equal(d_out[:n].cpu().numpy(), host_ids)


def test_encode_capacity_error(product):
  import ctypes
  from shredword_b200.cbase import lib
  t = product.BPETrainer(300, min_pair_freq=2)
  t.load_buffer(cases.corpus("ref_fixture")); t.train_quiet()
  enc = t.encoder()
  text = np.frombuffer(cases.corpus("ref_fixture"), dtype=np.uint8)
  out = np.zeros(4, dtype=np.int32)
  n = lib.swb_encode(enc.h, text.ctypes.data_as(ctypes.c_void_p), text.size, out.ctypes.data_as(ctypes.c_void_p), out.size, None, 0, None)
  assert n == -1  # too small: an error, never a truncated result


@pytest.mark.parametrize("piece", [5000, 100_000])
def test_encode_host_pipeline_many_pieces(piece, product, oracle_mod, monkeypatch):
  """swb_encode streams host text through the device in pieces on three streams (copy in / encode / copy out, two
  buffers each way). Many small pieces -- including a word longer than a piece -- must give the ids of one piece."""
  data = cases.corpus("multi_ties")[:700_000]
  data = data[:300_000] + b" " + b"abcdefgh" * 2000 + b" " + data[300_000:]
  t = product.BPETrainer(target_vocab_size=700, min_pair_freq=5)
  t.load_buffer(data)
  t.train_quiet()
  enc = t.encoder()
  oids, own = oracle_mod.encode(t.merges_array(), t.byte_map(), data, with_word_counts=True)
  monkeypatch.setenv("SWB_ENCODE_PIECE", str(piece))
  l0 = enc.kernel_launches
  ids, wn = enc.encode(data, with_word_counts=True)
  assert enc.kernel_launches - l0 >= 2 * (len(data) // (piece + 16000) )  # (several pieces went through)
  assert np.array_equal(own, wn) and np.array_equal(oids, ids)
  out = np.empty(len(data), dtype=np.int32)
  assert enc.encode_into(np.frombuffer(data, dtype=np.uint8), out) == len(oids) and np.array_equal(out[: len(oids)], oids)
