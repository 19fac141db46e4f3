"""CPU: the corpus generator names exactly one byte string per (spec, seed), whatever the thread count."""
import hashlib

import numpy as np

from shredword_b200 import synth


def test_generator_is_deterministic_and_exact_size():
  spec = synth.small_spec(5_000_000, 40_000, 3)
  a = synth.corpus_bytes(spec)
  h = hashlib.sha256()
  n = 0
  for c in synth.generate(spec, threads=1):
    h.update(c.tobytes()); n += c.size
  assert n == spec.nbytes == a.size and h.hexdigest() == hashlib.sha256(a.tobytes()).hexdigest()
  assert a[-1] == ord("\n") and set(np.unique(a)) <= set(b"abcdefghijklmnopqrstuvwxyz \n")
  lines = bytes(a[:100_000]).split(b"\n")[:-1]
  assert all(len(l.split()) == synth.WORDS_PER_LINE for l in lines)


def test_prefix_is_stable():
  """bench.py's CPU sample is a prefix of the big corpus generated through a smaller spec."""
  big = synth.corpus_bytes(synth.small_spec(9_000_000, 40_000, 3))
  small = synth.corpus_bytes(synth.small_spec(3_000_000, 40_000, 3))
  assert np.array_equal(big[:2_999_000], small[:2_999_000])


def test_multi_alphabet_is_valid_utf8():
  a = synth.corpus_bytes(synth.small_spec(1_000_000, 20_000, 9, "multi"))
  bytes(a[: a.size - 200]).rsplit(b"\n", 1)[0].decode("utf-8")
