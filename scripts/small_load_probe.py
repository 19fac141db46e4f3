"""Where does the load of a SMALL corpus spend its time? (bench.py's like-for-like leg: our arm on the reference's 32 MB sample)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer

spec = synth.CONFIGS["config3_10GB"]
small = synth.CorpusSpec(spec.name, 32_000_000 + 4096, spec.n_types, spec.alphabet, spec.zipf_s, spec.seed)
arr = synth.corpus_bytes(small)
end = 32_000_000
while arr[end - 1] != 10: end -= 1
pinned = torch.empty(arr.size, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = arr
KW = dict(target_vocab_size=32768, unk_id=0, character_coverage=0.995, min_pair_freq=2000)
for label, buf in (("pinned", pinned.numpy()[:end]), ("pageable", arr[:end]), ("pinned", pinned.numpy()[:end])):
  for it in range(3):
    t = BPETrainer(**KW)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t.load_buffer(buf)
    t1 = time.perf_counter()
    n = t.train_quiet()
    t2 = time.perf_counter()
    st = t.stats()
    print(f"{label} #{it}: load wall {1e3*(t1-t0):.1f} ms (stats load {st['load_ms']:.1f}, tokenize {st['tokenize_ms']:.2f}), train wall {1e3*(t2-t1):.1f} ms "
          f"(count {st['count_ms']:.2f}, merge {st['merge_ms']:.1f}), merges {n}, words {st['words']}", flush=True)
    t.destroy()
