"""GPU: differential fuzzing of the whole CUDA path (load, count, merge loop, save histogram, encoder) against the CPU
oracle on random tiny corpora: small alphabets, every delimiter kind, low min_pair_freq (tie-heavy), unk ids inside and
outside 0..255, coverage values that drop several bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("block", range(3))
def test_random_small_corpora_match_oracle(block, product, oracle_mod):
  for seed in range(block * 8, block * 8 + 8):
    rng = np.random.default_rng(5000 + seed)
    alphabet = np.frombuffer(b"abcdefgh\xc3\xa9"[: int(rng.integers(2, 11))], dtype=np.uint8)
    n_words = int(rng.integers(1, 3000))
    words = [bytes(rng.choice(alphabet, size=int(rng.integers(1, 9)))) for _ in range(n_words)]
    if seed % 5 == 0:
      words.append(bytes(rng.choice(alphabet, size=int(rng.integers(120, 400)))))  # a word around / beyond one 128-symbol row
    seps = [b" ", b"\n", b"\t", b"  ", b"\r\n", b" \n "]
    data = b"".join(w + seps[int(rng.integers(0, len(seps)))] for w in words)
    kw = dict(target_vocab_size=int(rng.integers(257, 450)), min_pair_freq=int(rng.integers(1, 6)),
              unk_id=int(rng.choice([0, 0, 0, 97, 98, -1, -7])), character_coverage=float(rng.choice([0.995, 0.9, 0.6])))
    o = oracle_mod.Oracle(kw["target_vocab_size"], kw["unk_id"], kw["character_coverage"], kw["min_pair_freq"])
    assert o.load_buffer(data) == 0
    n_o = o.train()
    t = product.BPETrainer(**kw)
    t.load_buffer(data)
    n_t = t.train_quiet()
    assert n_o == n_t and np.array_equal(o.merges, t.merges_array()), (seed, kw)
    assert np.array_equal(o.token_freq(), t.token_freq()), (seed, kw)
    ids = t.encoder().encode(data)
    assert np.array_equal(ids, oracle_mod.encode(t.merges_array(), t.byte_map(), data)), (seed, kw)
    t.destroy()
