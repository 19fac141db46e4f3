"""Development aid: one load + short train + one encode on a 256 MB corpus (targets for ncu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer
spec = synth.small_spec(256_000_000, 2_000_000, 11)
arr = synth.corpus_bytes(spec)
d = torch.from_numpy(arr).cuda()
t = BPETrainer(2048, min_pair_freq=2000)
t.load_device(d.data_ptr(), d.numel())
n = t.train_quiet()
enc = t.encoder()
out = torch.empty(d.numel(), dtype=torch.int32, device="cuda")
for _ in range(2):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  ntok = enc.encode_device(d.data_ptr(), d.numel(), out.data_ptr(), out.numel())
  torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"merges {n} tokens {ntok} encode {d.numel()/1e6/dt:.0f} MB/s load_ms {t.stats()['load_ms']:.1f}")
