"""Helper of tests/test_gpu_variants.py: trains a few corpora through the C-ABI in THIS process (whose environment
selects a code path of the library: the switches are read once per process) and compares every result with the oracle.
Prints one JSON line with the stats that prove which path ran."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)

import cases  # noqa: E402
import oracle as O  # noqa: E402
from shredword_b200.trainer import BPETrainer  # noqa: E402


def main():
  names = sys.argv[1:]
  out = {}
  for name in names:
    kw = cases.kwargs(name)
    data = cases.corpus(name)
    o = O.Oracle(kw.get("target_vocab_size", 8192), kw.get("unk_id", 0), kw.get("character_coverage", 0.995), kw.get("min_pair_freq", 2000))
    assert o.load_buffer(data) == 0
    n_o = o.train()
    t = BPETrainer(**kw)
    t.load_buffer(data)
    n_t = t.train_quiet()
    assert n_o == n_t, (name, n_o, n_t)
    assert np.array_equal(o.merges, t.merges_array()), f"{name}: merge lists differ"
    assert np.array_equal(o.token_freq(), t.token_freq()), f"{name}: token histograms differ"
    _, _, oso, osy, _ = o.words()
    _, _, tso, tsy, _ = t.words()
    assert np.array_equal(oso, tso) and np.array_equal(osy, tsy), f"{name}: final segmentation differs"
    ids = t.encoder().encode(data)
    assert np.array_equal(ids, O.encode(t.merges_array(), t.byte_map(), data)), f"{name}: encoding differs"
    st = t.stats()
    out[name] = {k: st[k] for k in ("merge_launches", "resident_local_merges", "resident_grid_merges", "hints_taken", "hints_sent", "kernel_launches")}
    out[name]["merges"] = n_t
    t.destroy()
  print(json.dumps(out))


if __name__ == "__main__":
  main()
