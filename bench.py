#!/usr/bin/env python
"""bench.py -- BPE trainer + encoder hot path on B200, one JSON line (contract: see DESIGN.md, "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2_1GB]

A "step" is one full pass of the training hot path over the workload corpus: load (tokenise, unique
word table, row packing) + initial pair count + every merge. Metric: train corpus-GB/s (corpus bytes
per step second); merges/s and the encoder's MB/s ride along in `extra`.

  value : corpus already resident in HBM when the step starts (swb_load_corpus_device + bpe_train)
  e2e   : the same through the host-buffer C-ABI call (swb_load_corpus_buffer from PINNED host memory,
          H2D inside the timed region, merge list read back on the host)
  roofline     : the dominant kernel (merge_rows), algorithmic bytes 4*S_live + 8*W per launch
                  over its CUDA-event time, against MEASURED_PEAKS.json's HBM figure
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, zero-filling malloc), one host core, on a bounded
                 prefix of the same corpus; falls back to the oracle port where oracle/_ref is not built

--impl reference times that same CPU reference on the bounded sample as the step.
Under torchrun (N > 1) the unique words are sharded over the ranks (shredword_b200.distributed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from shredword_b200 import synth  # noqa: E402

TRAIN_KW = dict(target_vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000)
CPU_SAMPLE_BYTES = int(os.environ.get("SWB_BENCH_CPU_SAMPLE_BYTES", 32 * 1000 * 1000))  # (the tests shrink the CPU sample)
SHARDED_MERGE = bool(int(os.environ.get("SWB_BENCH_SHARDED_MERGE", "0")))  # N > 1: shard the merge loop too (slower: one collective per merge)


def log(*a):
  print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
  """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index: int):
    self.idx = gpu_index
    self.proc = None
    self.lines = []

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("SWB_BENCH_CLOCK_MS", "200"),
                                    "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.th = threading.Thread(target=self._read, daemon=True)
      self.th.start()
    except Exception:
      self.proc = None

  def _read(self):
    for ln in self.proc.stdout:
      self.lines.append(ln.strip())

  def stop(self) -> dict:
    if not self.proc:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.25)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for ln in self.lines:
      f = [x.strip() for x in ln.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
        if v.lower().startswith("active"):
          reasons.add(name)
    if not sm:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference / oracle
def cpu_sample_path(corpus: np.ndarray, workdir: str) -> tuple[str, int]:
  """First ~32 MB of the workload corpus, cut after a newline."""
  end = min(CPU_SAMPLE_BYTES, corpus.size)
  while end > 0 and corpus[end - 1] != ord("\n"):
    end -= 1
  p = os.path.join(workdir, "cpu_sample.txt")
  with open(p, "wb") as f:
    f.write(corpus[:end].tobytes())
  return p, end


def run_cpu_once(sample_path: str, nbytes: int, workdir: str) -> dict:
  """One full train of the CPU implementation on the sample: the unmodified reference if oracle/_ref is
  built, else the oracle port. Returns value in corpus-GB/s and details."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as O
  t0 = time.perf_counter()
  if O.ref_available():
    tm = O.run_reference(sample_path, TRAIN_KW["target_vocab_size"], TRAIN_KW["min_pair_freq"], "-", "-",
                         unk_id=TRAIN_KW["unk_id"], coverage=TRAIN_KW["character_coverage"])
    secs = tm["load_s"] + tm["init_s"] + tm["merge_s"]
    kind, merges = "reference", tm["merges"]
    detail = {k: tm[k] for k in ("load_s", "merge_s")}
  else:
    O.build(ref=False)
    o = O.Oracle(**TRAIN_KW)
    o.load_corpus(sample_path)
    merges = o.train()
    secs = time.perf_counter() - t0
    kind, detail = "port", {}
  return {"value": nbytes / 1e9 / secs, "unit": "GB/s", "cores": 1, "kind": kind, "seconds": secs, "merges": merges,
          "merges_per_s": merges / secs if secs > 0 else None,
          "sample": f"first {nbytes} bytes of the workload corpus, full train (load+count+all merges), {TRAIN_KW}", **detail}


# ----------------------------------------------------------------------------- workload
def make_corpus(spec: synth.CorpusSpec, pinned: bool, first_chunk: int = 0):
  """The workload corpus as a uint8 numpy array (backed by pinned host memory when possible)."""
  import torch
  t0 = time.perf_counter()
  if pinned and torch.cuda.is_available():
    host = torch.empty(spec.nbytes, dtype=torch.uint8, pin_memory=True)
  else:
    host = torch.empty(spec.nbytes, dtype=torch.uint8)
  arr = host.numpy()
  pos = 0
  for chunk in synth.generate(spec, first_chunk=first_chunk):
    arr[pos: pos + chunk.size] = chunk
    pos += chunk.size
  assert pos == spec.nbytes
  log(f"[bench] corpus {spec.name}: {spec.nbytes} bytes generated in {time.perf_counter() - t0:.1f}s")
  return host, arr


def main():
  # fd 1 carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr
  real_stdout = os.dup(1)
  os.dup2(2, 1)

  def emit(obj):
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(obj) + "\n").encode())

  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default="config2_1GB", choices=list(synth.CONFIGS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-encode", action="store_true")
  args = ap.parse_args()

  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  spec = synth.CONFIGS[args.workload]
  config = {"workload": f"{args.workload}: train vocab {TRAIN_KW['target_vocab_size']} on a {spec.nbytes / 1e9:g} GB synthetic Zipfian "
                        f"{spec.alphabet} corpus ({spec.n_types} word types, s={spec.zipf_s}, seed {spec.seed}), "
                        f"min_pair_freq {TRAIN_KW['min_pair_freq']}, unk_id 0, coverage 0.995",
            "corpus_bytes": spec.nbytes, "l2": "input (1 GB) and the first-pass word table are larger than L2; no flush between steps",
            "parallelism": "1 GPU" if world == 1 else
                           f"weak scaling: {world} GPUs x one {spec.nbytes / 1e9:g} GB piece each (same word types, disjoint sampling streams) = one "
                           f"{world * spec.nbytes / 1e9:g} GB corpus; range-split tokenising + NCCL word-table exchange (the part that scales with "
                           f"the corpus), then " + ("unique words sharded over the ranks, per-merge NCCL all-gather of delta records, replicated "
                           "frequency table + heap" if SHARDED_MERGE else "every rank runs the latency-bound merge loop on all unique words "
                           "(replicated, no per-merge collective)")}

  # ------------------------------------------------------------------ reference arm
  if args.impl == "reference":
    if rank != 0:
      return
    workdir = tempfile.mkdtemp(prefix="swb_bench_")
    corpus = synth.corpus_bytes(synth.CorpusSpec(spec.name, min(spec.nbytes, CPU_SAMPLE_BYTES + 4096), spec.n_types, spec.alphabet, spec.zipf_s, spec.seed))
    sample, nbytes = cpu_sample_path(corpus, workdir)
    runs = []
    for i in range(args.warmup + args.steps):
      r = run_cpu_once(sample, nbytes, workdir)
      log(f"[bench] reference step {i}: {r['seconds']:.2f}s {r['value']:.5f} GB/s")
      if i >= args.warmup:
        runs.append(r)
    secs = float(np.mean([r["seconds"] for r in runs]))
    value = nbytes / 1e9 / secs
    cb = dict(runs[-1]); cb["value"] = value
    emit(({
      "impl": "reference", "metric": "train_corpus_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "int32 symbols, uint64 counts", "data": "synthetic", "config": config,
      "cpu_baseline": cb, "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "extra": {"merges_per_s": runs[-1]["merges_per_s"], "note": "reference is single-threaded; each step = full train on the bounded sample"},
    }))
    return

  # ------------------------------------------------------------------ our arm
  import torch
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  from shredword_b200 import build as B
  B.build()
  from shredword_b200.cbase import lib
  from shredword_b200.trainer import BPETrainer, bind_host_thread_to_gpu
  lib.swb_set_device(local_rank)
  cpus = None if os.environ.get("SWB_BENCH_NO_BIND") else bind_host_thread_to_gpu(local_rank)  # before any pinned allocation (first touch)
  log(f"[bench] rank {rank}: host thread bound to {len(cpus)} CPUs of the GPU's NUMA node" if cpus else f"[bench] rank {rank}: host thread not bound")
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    from shredword_b200.distributed import DistributedBPETrainer

  # N > 1: rank r holds piece r of an N x 1 GB corpus (bytes [r, r+1) GB of the whole)
  host, arr = make_corpus(spec, pinned=True, first_chunk=rank * synth.PIECE_STRIDE)
  d_corpus = host.to(dev, non_blocking=False)
  nbytes = spec.nbytes
  total_bytes = nbytes * world

  def new_trainer():
    if world > 1:
      return DistributedBPETrainer(**TRAIN_KW, device=dev, sharded_merge=SHARDED_MERGE)
    return BPETrainer(**TRAIN_KW)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  last = {}

  def step(resident: bool, timing: bool):
    w0 = time.perf_counter()
    old = last.pop("trainer", None)
    if old is not None:
      old.destroy()  # the previous step's handle: its device memory goes back before the next step allocates
    w1 = time.perf_counter()
    t = new_trainer()
    t.set_kernel_timing(timing)
    if world > 1:
      t.load_shard(d_corpus if resident else arr, rank * nbytes)
    elif resident:
      t.load_device(d_corpus.data_ptr(), nbytes)
    else:
      t.load_buffer(arr)
    w2 = time.perf_counter()
    merges = t.train_quiet()
    w3 = time.perf_counter()
    m = t.merges_array()  # the step's result, read on the host
    st = t.stats()
    st["merges"] = merges; st["merge_bytes"] = m.nbytes
    st["wall_ms"] = {"destroy_prev": (w1 - w0) * 1e3, "create+load": (w2 - w1) * 1e3, "train": (w3 - w2) * 1e3,
                     "read": (time.perf_counter() - w3) * 1e3}
    last["trainer"] = t
    return st

  def timed(resident: bool, k: int, timing: bool):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    stats = [step(resident, timing) for _ in range(k)]
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ev_ms = e0.elapsed_time(e1)
    ms = max(ev_ms, 0.0)
    if world > 1:
      tt = torch.tensor([ms], device=dev)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      ms = float(tt.item())
    return ms / k, wall_ms / k, stats

  # the sampler is started before the warm-up: spawning nvidia-smi and its first NVML queries stall the
  # driver for tens of ms, which must not land inside the timed region; it keeps sampling through it
  sampler = ClockSampler(local_rank)
  sampler.start()
  for i in range(args.warmup):
    s = step(True, False)
    log(f"[bench] warmup {i}: merges={s['merges']} load={s['load_ms']:.1f}ms count={s['count_ms']:.1f}ms merge={s['merge_ms']:.1f}ms")
  step(False, False)  # e2e warm-up (pinned registration, first H2D)

  ms_res, wall_res, st_res = timed(True, args.steps, timing=False)
  ms_e2e, wall_e2e, st_e2e = timed(False, args.steps, timing=False)
  clocks = sampler.stop()
  # roofline pass: same steps with the dominant kernel bracketed by CUDA events (kept out of `value`,
  # the two event records per launch add latency to a launch-latency-bound loop)
  ms_tim, _, st_tim = timed(True, args.steps, timing=True)

  merges = st_res[-1]["merges"]
  value = total_bytes / 1e9 / (ms_res / 1e3)
  e2e_value = total_bytes / 1e9 / (ms_e2e / 1e3)
  launches = int(sum(s["kernel_launches"] for s in st_res) + sum(s["kernel_launches"] for s in st_e2e))

  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  peak = float(peaks.get("hbm_gbs", 6650.0))
  peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
  # Dominant kernel = the merge kernel. In the product path it is ONE resident launch (merge_cluster) that serves every
  # merge of the step; its per-merge device time (command seen -> result published, %globaltimer, accumulated inside the
  # kernel) is the "launch duration" of the roofline. The CUDA-event pass over the per-launch kernel (merge_rows) rides along.
  def kernel_numbers(stats):
    alg = sum(s["merge_alg_bytes"] for s in stats); kms = sum(s["merge_kernel_ms"] for s in stats)
    n = sum(s["merges"] for s in stats)
    return alg, kms, n
  alg, kms, nl = kernel_numbers(st_res)
  resident = sum(s.get("resident_local_merges", 0) + s.get("resident_grid_merges", 0) for s in st_res) > 0
  if not resident:  # (multi-GPU and fallback paths launch merge_rows per merge: use the CUDA-event pass)
    alg, kms, nl = kernel_numbers(st_tim)
  alg_t, kms_t, nl_t = kernel_numbers(st_tim)
  scan = sum(s["merge_scan_bytes"] for s in st_res)
  achieved = alg / 1e9 / (kms / 1e3) if kms > 0 else None
  traffic = None
  try:  # dram bytes per launch of the merge kernel from the committed ncu capture of this round
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_merge_kernel_traffic.json")))["dram_bytes_per_launch"]
  except Exception:
    pass
  roofline = {"bound": "hbm", "kernel": "merge_cluster (resident, per merge)" if resident else "merge_rows (per launch)",
              "achieved": achieved, "peak": peak, "unit": "GB/s",
              "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
              "alg_bytes_per_launch": alg / nl if nl else None, "launched_over_bytes_per_launch": scan / nl if nl else None,
              "avg_launch_us": kms * 1e3 / nl if nl else None, "launches": nl,
              "kernel_share_of_step": (kms / args.steps) / ms_res if (ms_res and resident) else ((kms / args.steps) / ms_tim if ms_tim else None),
              "resident_split": {"local_merges": sum(s.get("resident_local_merges", 0) for s in st_res),
                                 "grid_merges": sum(s.get("resident_grid_merges", 0) for s in st_res),
                                 "local_us_per_merge": (sum(s.get("resident_local_ms", 0) for s in st_res) * 1e3 /
                                                        max(1, sum(s.get("resident_local_merges", 0) for s in st_res))),
                                 "grid_us_per_merge": (sum(s.get("resident_grid_ms", 0) for s in st_res) * 1e3 /
                                                       max(1, sum(s.get("resident_grid_merges", 0) for s in st_res)))} if resident else None,
              "per_launch_check": {"kernel": "merge_rows (one launch per merge, CUDA events)",
                                   "achieved": alg_t / 1e9 / (kms_t / 1e3) if kms_t > 0 else None,
                                   "avg_launch_us": kms_t * 1e3 / nl_t if nl_t else None, "launches": nl_t},
              "note": "algorithmic bytes = 4*S_live + 8*W per merge (SURVEY.md 8(d), full-scan form); the birth-log index makes a merge touch "
                      "far fewer bytes than that, and the row stream is L2-resident, so achieved may exceed the HBM copy peak; `traffic` = "
                      "dram bytes per launch of the per-launch kernel from the ncu capture in profiles/"}

  out = {
    "metric": "train_corpus_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "int32 symbols, uint64 counts", "data": "synthetic", "config": config, "clocks": clocks,
    "e2e": {"value": e2e_value, "unit": "GB/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(total_bytes),
            "d2h_bytes_per_step": int(st_e2e[-1]["merge_bytes"])},
    "gpu_launches": launches, "roofline": roofline,
    "extra": {"merges": merges, "merges_per_s": merges / (st_res[-1]["merge_ms"] / 1e3) if st_res[-1]["merge_ms"] else None,
              "us_per_merge": st_res[-1]["merge_ms"] * 1e3 / max(merges, 1),
              "phase_ms": {k: st_res[-1][k] for k in ("load_ms", "count_ms", "merge_ms")},
              "e2e_phase_ms": {k: st_e2e[-1][k] for k in ("load_ms", "count_ms", "merge_ms")},
              "wall_ms": [s_["wall_ms"] for s_ in st_res], "e2e_wall_ms": [s_["wall_ms"] for s_ in st_e2e],
              "unique_words": st_res[-1]["words"], "rows": st_res[-1]["rows"],
              "look_ahead": {k: st_res[-1].get(k) for k in ("hints_sent", "hints_taken", "hints_rejected", "host_peek_ms")},
              "host_split_ms": {k: st_res[-1].get(k) for k in ("host_pop_ms", "host_wait_ms", "host_apply_ms")},
              "collectives_per_step": st_res[-1].get("collectives"), "exchange_bytes_per_step": st_res[-1].get("exchange_bytes"), "wall_ms_per_step": wall_res,
              "ms_per_step_with_kernel_timing": ms_tim},
  }

  # ------------------------------------------------------------------ encoder (rides along)
  if not args.no_encode and world == 1:
    enc = last["trainer"].encoder()
    d_out = torch.empty(nbytes, dtype=torch.int32, device=dev)
    h_out = torch.empty(nbytes // 2 + 16, dtype=torch.int32, pin_memory=True)
    for _ in range(2):
      ntok = enc.encode_device(d_corpus.data_ptr(), nbytes, d_out.data_ptr(), d_out.numel())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = enc.kernel_launches
    e0.record()
    for _ in range(args.steps):
      ntok = enc.encode_device(d_corpus.data_ptr(), nbytes, d_out.data_ptr(), d_out.numel())
    e1.record(); torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1) / args.steps
    enc.encode_into(arr[: 64 << 20], h_out.numpy())
    e0.record()
    t0 = time.perf_counter()
    ntok2 = enc.encode_into(arr, h_out.numpy())
    e1.record(); torch.cuda.synchronize()
    enc_e2e_ms = e0.elapsed_time(e1)
    out["gpu_launches"] += int(enc.kernel_launches - l0)
    out["extra"]["encode"] = {"MB_per_s": nbytes / 1e6 / (enc_ms / 1e3), "e2e_MB_per_s": nbytes / 1e6 / (enc_e2e_ms / 1e3),
                              "tokens": int(ntok), "bytes_per_token": nbytes / max(ntok, 1),
                              "alg_bytes_per_input_byte": 1 + 4 * ntok / nbytes,
                              "hbm_frac": (nbytes + 4 * ntok) / 1e9 / (enc_ms / 1e3) / peak,
                              "e2e_h2d_bytes": int(nbytes), "e2e_d2h_bytes": int(4 * ntok2)}
    del d_out

  # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
  if world == 1 and rank == 0 and not args.no_cpu_baseline:
    workdir = tempfile.mkdtemp(prefix="swb_bench_")
    sample, sb = cpu_sample_path(arr, workdir)
    cb = run_cpu_once(sample, sb, workdir)
    log(f"[bench] cpu baseline ({cb['kind']}): {cb['seconds']:.1f}s on {sb} bytes -> {cb['value']:.5f} GB/s")
    out["cpu_baseline"] = cb

  if rank == 0:
    emit(out)
  if world > 1:
    last.pop("trainer").destroy()
    lib.swb_dist_shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
