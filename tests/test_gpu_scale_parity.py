"""GPU: parity at the sizes the bench numbers are quoted on (BASELINE.json configs 2 and 3).

config 2 (1 GB, vocab 8192): the CUDA path's .model and .vocab must be byte-identical to the files the UNMODIFIED
reference wrote for the same corpus (tests/golden/config2_1GB.*, generated once in the build container by
scripts/make_golden_big.py: a 12-minute single-core run).
config 3 (10 GB, vocab 32768): the reference would need days; the pinned oracle's merge list (tests/golden/
config3_10GB.model, digests.json) is the golden there, plus the size-independent properties (histogram of the
encoded corpus == .vocab frequency column, decode(encode(x)) == x on a slice).
A scripted, host-free launch of the resident kernel (the thing ncu profiles) must leave the same segmentation behind
as the host-driven loop."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIGESTS = json.load(open(os.path.join(GOLD, "digests.json")))


def _corpus(name):
  from shredword_b200 import synth
  return synth.corpus_bytes(synth.CONFIGS[name])


def test_config2_full_size_equals_reference_files(product, tmp_path):
  data = _corpus("config2_1GB")
  man = json.load(open(os.path.join(GOLD, "manifest_big.json")))["config2_1GB"]
  assert hashlib.sha256(data).hexdigest() == man["corpus_sha256"], "corpus generator changed: regenerate the golden files"
  t = product.BPETrainer(**man["kwargs"])
  t.load_buffer(data)
  assert t.train_quiet() == man["merges"] == 7936
  st = t.stats()
  assert st["resident_local_merges"] + st["resident_grid_merges"] == 7936  # the product path: the resident kernel served every merge
  t.save(str(tmp_path / "m"), str(tmp_path / "v"))
  assert (tmp_path / "m").read_bytes() == open(os.path.join(GOLD, "config2_1GB.model"), "rb").read(), ".model differs from the reference's"
  assert (tmp_path / "v").read_bytes() == open(os.path.join(GOLD, "config2_1GB.vocab"), "rb").read(), ".vocab differs from the reference's"
  # encode pinned by the reference's .vocab frequency column, at full size
  enc = t.encoder()
  ids = enc.encode(data)
  T = 256 + 7936
  hist = np.bincount(ids, minlength=T)
  freqs = [int(ln.rsplit(b" ", 1)[1]) for ln in (tmp_path / "v").read_bytes().split(b"\n") if b" " in ln and ln.rsplit(b" ", 1)[1].isdigit()]
  assert len(freqs) == T and np.array_equal(hist, np.array(freqs)), "histogram(encode(corpus)) != .vocab frequency column"
  cut = 2_000_000 + int(np.nonzero(data[2_000_000:2_001_000] == 10)[0][0]) + 1
  bm = t.byte_map()  # bytes the coverage rule dropped come back as unk_id (0)
  table = bytes(int(bm[b]) & 255 for b in range(256))
  assert enc.decode(enc.encode(data[:cut])) == bytes(data[:cut]).translate(None, b" \t\r\n").translate(table)
  t.destroy()


@pytest.mark.skipif("config3_10GB" not in DIGESTS, reason="no golden digest for config 3 committed")
def test_config3_full_size_equals_oracle_digest(product, tmp_path):
  import torch
  if torch.cuda.mem_get_info()[1] < 100e9:
    pytest.skip("needs a 100+ GB GPU")
  g = DIGESTS["config3_10GB"]
  data = _corpus("config3_10GB")
  t = product.BPETrainer(target_vocab_size=32768, unk_id=0, character_coverage=0.995, min_pair_freq=2000)
  t.load_buffer(data)
  n = t.train_quiet()
  m = np.ascontiguousarray(t.merges_array(), dtype="<i4")
  assert n == g["merges"] and hashlib.md5(m.tobytes()).hexdigest() == g["model_md5"], "merge list differs from the oracle's at 10 GB"
  gold_model = os.path.join(GOLD, "config3_10GB.model")
  if os.path.exists(gold_model):
    assert m.tobytes() == open(gold_model, "rb").read()
  t.save(str(tmp_path / "m"), str(tmp_path / "v"))
  assert hashlib.md5((tmp_path / "v").read_bytes()).hexdigest() == g["vocab_md5"], ".vocab (token histogram of the final segmentation) differs from the oracle's"
  t.destroy()


def test_scripted_launch_equals_host_driven_loop(product, oracle_mod):
  """swb_profile_scripted_merges: the same merges, commands from a device-side script, no host in the loop."""
  import cases
  name = "config1_10MB"
  kw = cases.kwargs(name)
  data = cases.corpus(name)
  t = product.BPETrainer(**kw); t.load_buffer(data); n = t.train_quiet()
  merges = t.merges_array(); freq = t.token_freq(); words = t.words()
  t.destroy()
  s = product.BPETrainer(**kw); s.load_buffer(data); s.init()
  ms = s.profile_scripted_merges(merges)
  assert ms > 0
  st = s.stats()
  assert st["resident_local_merges"] + st["resident_grid_merges"] == n and st["hints_taken"] == 0
  assert np.array_equal(s.token_freq()[: 256], freq[: 256])  # (the host never learnt the merges: only the initial ids are in its histogram range)
  w2 = s.words()
  n_sym = int(words[2][-1])
  assert np.array_equal(words[2], w2[2]) and np.array_equal(words[3][:n_sym], w2[3][:n_sym]), "segmentation after the scripted launch differs"
  with pytest.raises(RuntimeError):  # the handle is consumed
    s.merge_batch(1)
  s.destroy()
