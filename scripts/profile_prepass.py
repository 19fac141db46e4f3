"""Profiling target (ncu): the two optional pre-passes (stream_map<NormEmit>, stream_map<PretokEmit>) on a 256 MB Zipfian ASCII text,
device resident. Prints their wall numbers measured here without a profiler."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shredword_b200 import synth
from shredword_b200.cbase import lib

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 256_000_000
arr = synth.corpus_bytes(synth.small_spec(nbytes, 2_000_000, 11))
d = torch.from_numpy(arr).cuda()
out = torch.empty(2 * d.numel() + 16, dtype=torch.uint8, device="cuda")
res = {"corpus_bytes": nbytes}
for name, fn in (("pretokenize", lib.swb_pretokenize), ("normalize", lib.swb_normalize)):
  for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = fn(ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(out.data_ptr()), out.numel(), 1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
  res[name] = {"out_bytes": int(n), "ms": dt * 1e3, "GB_per_s": nbytes / 1e9 / dt}
print(json.dumps(res))
