"""Seeded corpora shared by the parity tests (sizes the oracle finishes in seconds)."""
import numpy as np

from shredword_b200 import synth


def _b(x):
  return x if isinstance(x, bytes) else bytes(x)


def long_word_corpus() -> bytes:
  """Words longer than one 128-symbol row (exercise the long-word kernels), lengths around the
  126/127/128 boundary, and repeated so that their pairs pass min_pair_freq."""
  rng = np.random.default_rng(5)
  words = []
  for L in (126, 127, 128, 129, 200, 255, 256, 257, 1000, 5000):
    w = bytes(rng.choice(np.frombuffer(b"abcab", dtype=np.uint8), size=L))
    words.append(w)
  short = [b"ab", b"abc", b"a", b"bca", b"cab", b"abab", b"aaaa", b"aaa", b"bb"]
  parts = []
  for rep in range(40):
    for w in words:
      parts.append(w)
      parts.append(short[rep % len(short)])
    parts.append(b"\n")
  return b" ".join(parts) + b"\n"


def ragged_corpus() -> bytes:
  """Every delimiter kind, runs of delimiters, no trailing newline, bytes >= 0x80, \\v and \\f as word bytes."""
  base = b"\t\tfoo  bar\r\nbaz\tfoo \r \n\n qux\x0bquux  foo\x0cbar caf\xc3\xa9 caf\xc3\xa9 na\xc3\xafve \xe4\xb8\xad\xe6\x96\x87 \xe4\xb8\xad\xe6\x96\x87"
  return (base + b" ") * 30 + b"tail"


def self_pair_corpus() -> bytes:
  """Runs of one letter: greedy left-to-right self-pair merges (a a a -> N a)."""
  ws = [b"a" * k for k in range(1, 12)] + [b"b" + b"a" * 5 + b"b", b"aabaa", b"baaab"]
  return (b" ".join(ws) + b"\n") * 25


CASES = {
  # name: (bytes factory, kwargs for the trainer)
  "ref_fixture": (synth.reference_test_corpus, dict(target_vocab_size=300, min_pair_freq=2)),
  "ref_fixture_f5": (synth.reference_test_corpus, dict(target_vocab_size=400, min_pair_freq=5)),
  "ascii_ties": (lambda: synth.corpus_bytes(synth.small_spec(2_000_000, 50_000, 3)), dict(target_vocab_size=1200, min_pair_freq=5)),
  "multi_ties": (lambda: synth.corpus_bytes(synth.small_spec(2_000_000, 50_000, 5, "multi")), dict(target_vocab_size=1200, min_pair_freq=5)),
  "multi_unk97": (lambda: synth.corpus_bytes(synth.small_spec(2_000_000, 50_000, 5, "multi")),
                  dict(target_vocab_size=900, min_pair_freq=3, unk_id=97, character_coverage=0.9)),
  "unk_enters_by_delta": (lambda: (" ".join(["qab cab dab xab eab fab gab hab"] * 50) + "\n").encode(), dict(target_vocab_size=300, min_pair_freq=2)),
  "long_words": (long_word_corpus, dict(target_vocab_size=330, min_pair_freq=2)),
  "ragged": (ragged_corpus, dict(target_vocab_size=300, min_pair_freq=2)),
  "self_pairs": (self_pair_corpus, dict(target_vocab_size=280, min_pair_freq=2)),
  "heap_exhausted": (synth.reference_test_corpus, dict(target_vocab_size=5000, min_pair_freq=30)),
  "target_below_256": (synth.reference_test_corpus, dict(target_vocab_size=100, min_pair_freq=2)),
  "single_word": (lambda: b"hello", dict(target_vocab_size=300, min_pair_freq=1)),
  "only_delims": (lambda: b" \n\t\r  \n", dict(target_vocab_size=300, min_pair_freq=1)),
  "empty": (lambda: b"", dict(target_vocab_size=300, min_pair_freq=1)),
  "negative_unk": (lambda: (" ".join(["qab cab dab xab eab fab gab hab"] * 50) + "\n").encode(),
                   dict(target_vocab_size=300, min_pair_freq=2, unk_id=-1)),
  "config1_10MB": (lambda: synth.corpus_bytes(synth.CONFIGS["config1_10MB"]), dict(target_vocab_size=500, min_pair_freq=1000)),
}

_cache = {}


def corpus(name: str) -> bytes:
  if name not in _cache:
    _cache[name] = _b(CASES[name][0]())
  return _cache[name]


def kwargs(name: str) -> dict:
  return dict(CASES[name][1])
