"""Scale check (development aid / evidence for DESIGN.md): one of BASELINE.json's large configs end to end on
one GPU -- train, then encode the same corpus with the trained merges -- checked through size-independent
properties: token histogram of encode(corpus) == the trainer's .vocab frequency column, decode(encode(x)) == x
on a slice, and (optionally) the first K merges against the CPU oracle run on the same corpus.

  python scripts/scale_check.py config3_10GB 32768 [oracle_merges]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer

name = sys.argv[1] if len(sys.argv) > 1 else "config3_10GB"
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
oracle_k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
spec = synth.CONFIGS[name]
t0 = time.perf_counter()
host = torch.empty(spec.nbytes, dtype=torch.uint8, pin_memory=True)
arr = host.numpy(); pos = 0
for ch in synth.generate(spec):
  arr[pos:pos + ch.size] = ch; pos += ch.size
print(f"corpus {name}: {spec.nbytes} bytes in {time.perf_counter() - t0:.1f}s", flush=True)
out = {"config": name, "vocab": vocab, "bytes": spec.nbytes}
t = BPETrainer(vocab, min_pair_freq=2000)
t1 = time.perf_counter(); t.load_buffer(arr); t2 = time.perf_counter()
n = t.train_quiet(); t3 = time.perf_counter()
st = t.stats()
out.update(load_s=t2 - t1, train_s=t3 - t2, merges=n, us_per_merge=(t3 - t2) / max(n, 1) * 1e6, words=st["words"], rows=st["rows"],
           train_GB_per_s=spec.nbytes / 1e9 / (t3 - t1), host_split_ms={k: st[k] for k in ("host_pop_ms", "host_wait_ms", "host_apply_ms")})
print(json.dumps(out), flush=True)
freq = t.token_freq()
enc = t.encoder()
# encode in pieces cut on newlines (documents), histogram on the device
piece = 1 << 30
hist = torch.zeros(256 + n, dtype=torch.int64, device="cuda")
d_text = torch.empty(piece + 64, dtype=torch.uint8, device="cuda")
d_out = torch.empty(piece, dtype=torch.int32, device="cuda")
pos = 0; ntok_total = 0; enc_s = 0.0
while pos < spec.nbytes:
  end = min(pos + piece, spec.nbytes)
  while end < spec.nbytes and arr[end - 1] != 10: end -= 1
  d_text[: end - pos].copy_(host[pos:end], non_blocking=True)
  torch.cuda.synchronize(); e0 = time.perf_counter()
  nt = enc.encode_device(d_text.data_ptr(), end - pos, d_out.data_ptr(), d_out.numel())
  torch.cuda.synchronize(); enc_s += time.perf_counter() - e0
  hist += torch.bincount(d_out[:nt], minlength=256 + n)
  ntok_total += nt
  if pos == 0:
    ids = d_out[: 200000].cpu().numpy()
    nw_bytes = enc.decode(ids)
    src = bytes(arr[: len(nw_bytes) * 2])
    # decode drops the delimiters: compare against the words of the source, with the bytes the coverage rule dropped
    # (reference bpe.cpp:257-279: at least one byte value always is) replaced by what unk_id 0 decodes to
    bm = t.byte_map()
    words = np.frombuffer(b"".join(src.split()), dtype=np.uint8)
    want = np.where(bm[words] == words, words, 0).astype(np.uint8).tobytes()[: len(nw_bytes)]
    out["roundtrip_ok"] = bool(nw_bytes == want)
  pos = end
h = hist.cpu().numpy().astype(np.uint64)
out.update(encode_tokens=int(ntok_total), encode_MB_per_s=spec.nbytes / 1e6 / enc_s, hist_equals_vocab_freq=bool(np.array_equal(h, freq)))
print(json.dumps(out), flush=True)
if oracle_k:
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as O
  O.build(ref=False)
  o = O.Oracle(target_vocab_size=256 + oracle_k, unk_id=0, character_coverage=0.995, min_pair_freq=2000)
  c0 = time.perf_counter()
  assert o.load_buffer(arr) == 0
  c1 = time.perf_counter()
  k = o.train()
  c2 = time.perf_counter()
  out.update(oracle_load_s=c1 - c0, oracle_merge_s=c2 - c1, oracle_merges=k,
             first_k_merges_equal=bool(np.array_equal(o.merges[:k], t.merges_array()[:k])))
  print(json.dumps(out), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"scale_{name}.json"), "w"))
