"""Development aid: where does the e2e (host buffer) step spend its time outside load_ms?"""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shredword_b200 import synth
from shredword_b200.trainer import BPETrainer
spec = synth.CONFIGS["config2_1GB"]
host = torch.empty(spec.nbytes, dtype=torch.uint8, pin_memory=True)
arr = host.numpy(); pos = 0
for ch in synth.generate(spec):
  arr[pos:pos + ch.size] = ch; pos += ch.size
d = host.cuda()
def one(tag, resident):
  t0 = time.perf_counter()
  t = BPETrainer(8192, min_pair_freq=2000)
  t1 = time.perf_counter()
  if resident: t.load_device(d.data_ptr(), d.numel())
  else: t.load_buffer(arr)
  t2 = time.perf_counter()
  if TRAIN: t.train_quiet()
  st = t.stats()
  t.destroy()
  t3 = time.perf_counter()
  print(f"{tag} resident={resident} create {1e3*(t1-t0):.2f} load {1e3*(t2-t1):.2f} (load_ms {st['load_ms']:.2f}) destroy {1e3*(t3-t2):.2f}", flush=True)
TRAIN = len(sys.argv) > 1
import bench
for i in range(3): one("plain", True)
for i in range(4): one("plain", False)
p = bench.ClockSampler(0); p.start()
time.sleep(1)
for i in range(3): one("smi", True)
for i in range(6): one("smi", False)
print(p.stop())
