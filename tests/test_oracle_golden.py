"""CPU: the oracle (C restatement) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py), and, where oracle/_ref is built, against the reference run live."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


def _oracle_for(oracle_mod, name):
  kw = cases.kwargs(name)
  o = oracle_mod.Oracle(kw.get("target_vocab_size", 8192), kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw.get("min_pair_freq", 2000))
  assert o.load_buffer(cases.corpus(name)) == 0
  return o


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_oracle_reproduces_reference_golden(name, oracle_mod, tmp_path):
  data = cases.corpus(name)
  assert hashlib.sha256(data).hexdigest() == MANIFEST[name]["corpus_sha256"], "corpus generator changed: regenerate tests/golden"
  o = _oracle_for(oracle_mod, name)
  n = o.train()
  assert n == MANIFEST[name]["merges"]
  o.save(str(tmp_path / "m"), str(tmp_path / "v"))
  assert (tmp_path / "m").read_bytes() == open(os.path.join(GOLD, name + ".model"), "rb").read()
  assert (tmp_path / "v").read_bytes() == open(os.path.join(GOLD, name + ".vocab"), "rb").read()
  # reference test/bpe_test.cpp:262-270: model size == 12 bytes per merge
  assert os.path.getsize(tmp_path / "m") == 12 * n


def test_oracle_matches_live_reference(oracle_mod, tmp_path):
  if not oracle_mod.ref_available():
    pytest.skip("oracle/_ref not built here (needs /root/reference)")
  from shredword_b200 import synth
  data = bytes(synth.corpus_bytes(synth.small_spec(3_000_000, 60_000, 41, "multi")))
  p = tmp_path / "c.txt"; p.write_bytes(data)
  tm = oracle_mod.run_reference(str(p), 1500, 7, str(tmp_path / "r.model"), str(tmp_path / "r.vocab"))
  o = oracle_mod.Oracle(1500, 0, 0.995, 7); o.load_corpus(str(p))
  assert o.train() == tm["merges"]
  o.save(str(tmp_path / "o.model"), str(tmp_path / "o.vocab"))
  assert (tmp_path / "o.model").read_bytes() == (tmp_path / "r.model").read_bytes()
  assert (tmp_path / "o.vocab").read_bytes() == (tmp_path / "r.vocab").read_bytes()
  # prefix property: a smaller target gives a prefix of the same merge list
  tm2 = oracle_mod.run_reference(str(p), 700, 7, str(tmp_path / "r2.model"), str(tmp_path / "r2.vocab"))
  full = oracle_mod.read_model(str(tmp_path / "r.model")); pre = oracle_mod.read_model(str(tmp_path / "r2.model"))
  assert tm2["merges"] == 444 and np.array_equal(full[:444], pre)


def test_reference_structural_checks(oracle_mod):
  """The 8 checks of reference test/bpe_test.cpp restated against the oracle (unk_id=0)."""
  from shredword_b200 import synth
  o = oracle_mod.Oracle(300, 0, 0.0, 0)          # defaults: coverage 0.0 -> 0.995, min_pair_freq 0 -> 2000
  assert o.load_buffer(synth.reference_test_corpus()) == 0
  assert o.num_words == 32
  o2 = oracle_mod.Oracle(300, 0, 0.995, 2); o2.load_buffer(synth.reference_test_corpus())
  o2.count_bigrams()
  f, s, fr, v = o2.heap()
  assert len(fr) > 0 and fr[0] >= fr[1]            # heap top is a maximum
  o3 = oracle_mod.Oracle(300, 0, 0.995, 2); o3.load_buffer(synth.reference_test_corpus()); o3.init()
  assert o3.merge_batch(1) == 1 and 0 <= o3.merges[0][0] < 1000
  n = o3.train()
  assert 0 < n <= 300 - 256


def test_encode_pinned_by_reference_vocab(oracle_mod):
  """The reference has no encoder; the pin: the token histogram of the encoded training
  corpus, with the trainer's byte map, equals the frequency column of the REFERENCE's .vocab file."""
  for name in ("ascii_ties", "multi_ties", "multi_unk97", "unk_enters_by_delta", "long_words", "ragged", "self_pairs"):
    o = _oracle_for(oracle_mod, name); o.train()
    merges = oracle_mod.read_model(os.path.join(GOLD, name + ".model"))
    assert np.array_equal(merges, o.merges)
    ids = oracle_mod.encode(merges, o.byte_map(cases.kwargs(name).get("unk_id", 0)), cases.corpus(name))
    T = 256 + len(merges)
    hist = np.bincount(ids[(ids >= 0) & (ids < T)], minlength=T)
    vocab = open(os.path.join(GOLD, name + ".vocab"), "rb").read()
    # each line ends with " <freq>\n"; token 10 prints a raw newline, so parse from the right
    freqs = []
    for ln in vocab.split(b"\n"):
      parts = ln.rsplit(b" ", 1)
      if len(parts) == 2 and parts[1].isdigit():
        freqs.append(int(parts[1]))
    assert len(freqs) == T
    assert np.array_equal(hist, np.array(freqs)), name


def test_encode_decode_roundtrip(oracle_mod):
  o = _oracle_for(oracle_mod, "multi_ties"); o.train()
  text = cases.corpus("ragged") + b" " + cases.corpus("multi_ties")[:200_000]
  ids, wn = oracle_mod.encode(o.merges, np.arange(256, dtype=np.int32), text, with_word_counts=True)
  import re
  words = [w for w in re.split(rb"[ \t\r\n]+", text) if w]   # the reference's 4 delimiters only (\v, \f are word bytes)
  assert wn.sum() == len(ids) and len(wn) == len(words)
  assert oracle_mod.decode(o.merges, ids) == text.translate(None, b" \t\r\n")
  # python restatement of base.py:10-36 on a few words (pure-python loop, small case only)
  rank = {(int(a), int(b)): (i, int(n)) for i, (a, b, n) in enumerate(o.merges)}
  out = []
  for w in words[:300]:
    seq = list(w)
    while True:
      cand = [(rank[p][0], p) for p in zip(seq, seq[1:]) if p in rank]
      if not cand:
        break
      _, pair = min(cand)
      new, i, res = rank[pair][1], 0, []
      while i < len(seq):
        if i + 1 < len(seq) and (seq[i], seq[i + 1]) == pair:
          res.append(new); i += 2
        else:
          res.append(seq[i]); i += 1
      seq = res
    out += seq
  assert out == list(ids[: len(out)])


def test_streaming_load_equals_buffer_load(oracle_mod):
  """The chunk-fed loader (used for the 50 GB configuration, which does not fit host memory) builds the same word table."""
  from shredword_b200 import synth
  spec = synth.small_spec(6_000_000, 40_000, 5, "multi")
  kw = dict(target_vocab_size=700, unk_id=0, character_coverage=0.995, min_pair_freq=20)
  a = oracle_mod.Oracle(**kw); a.load_buffer(np.concatenate(list(synth.generate(spec, chunk_words=120_000)))); na = a.train()
  b = oracle_mod.Oracle(**kw); assert b.load_chunks(synth.generate(spec, chunk_words=120_000)) == 0
  assert b.train() == na and np.array_equal(a.merges, b.merges) and np.array_equal(a.token_freq(), b.token_freq())
