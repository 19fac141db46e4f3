"""Development aid: trains config 2 several times and prints a digest of the merge list and the host replica's counters
per repetition -- every line of one corpus must be identical, whatever the hint mode (SWB_NO_HINTS / SWB_NO_EARLY_HINTS)."""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer

name = sys.argv[1] if len(sys.argv) > 1 else "config2_1GB"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec = synth.CONFIGS.get(name) or synth.small_spec(int(name), 2_000_000, 11)
arr = synth.corpus_bytes(spec)
d = torch.from_numpy(arr).cuda()
for rep in range(reps):
  t = BPETrainer(8192, min_pair_freq=2000)
  t.load_device(d.data_ptr(), d.numel())
  t0 = time.perf_counter()
  n = t.train_quiet()
  dt = time.perf_counter() - t0
  st = t.stats()
  print(hashlib.md5(t.merges_array().tobytes()).hexdigest(), n, st["records"], st["heap_pops"], st["heap_pushes"], st["heap_peak"],
        "hints", st["hints_sent"], st["hints_taken"], st["hints_rejected"], f"{dt / max(n, 1) * 1e6:.2f} us/merge", flush=True)
  t.destroy()
