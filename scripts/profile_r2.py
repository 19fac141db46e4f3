"""Profiling target (ncu): every hot kernel once, on a 256 MB Zipfian corpus, with NO host-interactive launch in the process.

  wt_tokenize   load of the resident corpus
  merge_rows    the training run that produces the merge list goes through the per-launch path (kernel timing on)
  merge_cluster ONE scripted launch (swb_profile_scripted_merges) that replays that merge list with its commands in device
                memory: the same kernel, the same work as the product's resident launch, but replayable by a profiler
  enc_fused     encode of the same corpus, device resident

Prints one JSON line: per-kernel wall numbers measured here without a profiler."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 256_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spec = synth.small_spec(nbytes, 2_000_000, 11)
arr = synth.corpus_bytes(spec)
d = torch.from_numpy(arr).cuda()
kw = dict(target_vocab_size=vocab, min_pair_freq=2000)

t = BPETrainer(**kw)
t.set_kernel_timing(True)  # -> one merge_rows launch per merge, no resident kernel
t.load_device(d.data_ptr(), d.numel())
n = t.train_quiet()
merges = t.merges_array()
st = t.stats()
enc = t.encoder()
out = torch.empty(d.numel() // 2 + 16, dtype=torch.int32, device="cuda")
for _ in range(2):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  ntok = enc.encode_device(d.data_ptr(), d.numel(), out.data_ptr(), out.numel())
  torch.cuda.synchronize(); enc_s = time.perf_counter() - t0
t.destroy()

s = BPETrainer(**kw)
s.load_device(d.data_ptr(), d.numel())
s.init()
ms = s.profile_scripted_merges(merges)
st2 = s.stats()
s.destroy()
print(json.dumps({"corpus_bytes": nbytes, "merges": int(n), "tokenize_ms": st["tokenize_ms"], "tokenize_GB_per_s": nbytes / 1e6 / st["tokenize_ms"],
                  "merge_rows_us_per_launch": st["merge_kernel_ms"] * 1e3 / max(n, 1),
                  "scripted_merge_cluster_ms": ms, "scripted_us_per_merge": ms * 1e3 / max(n, 1),
                  "scripted_local": st2["resident_local_merges"], "scripted_grid": st2["resident_grid_merges"],
                  "scripted_local_us": st2["resident_local_ms"] * 1e3 / max(1, st2["resident_local_merges"]),
                  "scripted_grid_us": st2["resident_grid_ms"] * 1e3 / max(1, st2["resident_grid_merges"]),
                  "alg_bytes_per_merge": st["merge_alg_bytes"] / max(n, 1),
                  "encode_GB_per_s": nbytes / 1e9 / enc_s, "tokens": int(ntok)}))
