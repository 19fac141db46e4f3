"""Fine-grained host timings of one training step (development aid; not part of the bench contract)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer

name = sys.argv[1] if len(sys.argv) > 1 else "config2_1GB"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
max_merges = int(sys.argv[3]) if len(sys.argv) > 3 else -1
spec = synth.CONFIGS.get(name) or synth.small_spec(int(name), 2_000_000, 11)
arr = synth.corpus_bytes(spec)
d = torch.from_numpy(arr).cuda()
torch.cuda.synchronize()
for rep in range(reps):
  t0 = time.perf_counter()
  t = BPETrainer(8192, min_pair_freq=2000)
  t1 = time.perf_counter()
  t.load_device(d.data_ptr(), d.numel())
  t2 = time.perf_counter()
  t.init()
  t3 = time.perf_counter()
  if max_merges < 0:
    n = t.merge_batch(8192 - 256)
  else:
    n = t.merge_batch(max_merges)
  t4 = time.perf_counter()
  m = t.merges_array()
  st = t.stats()
  t5 = time.perf_counter()
  t.destroy()
  t6 = time.perf_counter()
  print(json.dumps({"create": t1 - t0, "load": t2 - t1, "init": t3 - t2, "merge": t4 - t3, "read": t5 - t4, "destroy": t6 - t5,
                    "merges": n, "us_per_merge": (t4 - t3) / max(n, 1) * 1e6, "stats": st}))
