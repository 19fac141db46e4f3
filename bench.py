#!/usr/bin/env python
"""bench.py -- BPE trainer + encoder hot path on B200, one JSON line (contract: see DESIGN.md, "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3_10GB]

A "step" is one full pass of the training hot path over the workload corpus: load (tokenise, unique
word table, row packing) + initial pair count + every merge. Metric: train corpus-GB/s (corpus bytes
per step second); merges/s and the encoder's MB/s ride along in `extra`.

  workload : config3_10GB by default (BASELINE.json configs[2]: vocab 32k on 10 GB, the configuration the 1/2/4/8-GPU
             metric is quoted on; it fits one GPU). --workload config2_1GB / config1_10MB run the smaller ones.
  N > 1    : STRONG scaling on that one fixed corpus: rank r holds bytes [r/N, (r+1)/N) of it (cut on a line end).
             The load (tokenise + dedupe) is range-split and the word tables are exchanged over NCCL; the merge
             loop is one latency chain and runs on rank 0 (merge list broadcast afterwards).
  value    : corpus already resident in HBM when the step starts
  e2e      : the same through the host-buffer C-ABI calls from PINNED host memory, H2D inside the timed region, merge
             list read back on the host
  parity   : before the line is printed, the merge list of EVERY step (at every N) is compared with the committed
             golden digest of the workload (tests/golden/digests.json: the unmodified reference's .model for
             configs 1-2, the pinned oracle's for config 3) -- a mismatch fails the run
  roofline : the dominant kernel (merge_cluster, resident), algorithmic bytes 4*S_live + 8*W per merge
              over its own per-merge device time, against MEASURED_PEAKS.json's HBM figure; the
             bytes the index-driven kernel actually moves (ncu, profiles/) are reported next to it
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, zero-filling malloc), one host core, on a bounded
             prefix of the same corpus, next to OUR arm on that same prefix (like for like) and the committed
             full-size reference timing of config 2

--impl reference times that same CPU reference on the bounded sample as the step.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from shredword_b200 import synth  # noqa: E402

TRAIN_KWS = {
  "config1_10MB": dict(target_vocab_size=500, unk_id=0, character_coverage=0.995, min_pair_freq=1000),
  "config2_1GB": dict(target_vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000),
  "config3_10GB": dict(target_vocab_size=32768, unk_id=0, character_coverage=0.995, min_pair_freq=2000),
  "config5_50GB": dict(target_vocab_size=100000, unk_id=0, character_coverage=0.995, min_pair_freq=2000),
}
CPU_SAMPLE_BYTES = int(os.environ.get("SWB_BENCH_CPU_SAMPLE_BYTES", 32 * 1000 * 1000))  # (the tests shrink the CPU sample)
MERGE_LOOP = os.environ.get("SWB_BENCH_MERGE_LOOP", "rank0")  # N > 1: rank0 | replicated | sharded
STATE = {"rank": 0, "phase": "start", "step": -1}


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def golden_digest(workload: str):
  try:
    return json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json"))).get(workload)
  except Exception:
    return None


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


def run_cpu_once(sample_path: str, nbytes: int, kw: dict) -> dict:
  """One full train of the CPU implementation on the sample: the unmodified reference if oracle/_ref is
  built, else the oracle port. Returns value in corpus-GB/s and details."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as O
  t0 = time.perf_counter()
  if O.ref_available():
    tm = O.run_reference(sample_path, kw["target_vocab_size"], kw["min_pair_freq"], "-", "-",
                         unk_id=kw["unk_id"], coverage=kw["character_coverage"])
    secs = tm["load_s"] + tm["init_s"] + tm["merge_s"]
    kind, merges = "reference", tm["merges"]
    detail = {k: tm[k] for k in ("load_s", "merge_s")}
  else:
    O.build(ref=False)
    o = O.Oracle(**kw)
    o.load_corpus(sample_path)
    merges = o.train()
    secs = time.perf_counter() - t0
    kind, detail = "port", {}
  return {"value": nbytes / 1e9 / secs, "unit": "GB/s", "cores": 1, "kind": kind, "seconds": secs, "merges": merges,
          "merges_per_s": merges / secs if secs > 0 else None,
          "sample": f"first {nbytes} bytes of the workload corpus, full train (load+count+all merges), {kw}", **detail}


def full_config_reference() -> dict | None:
  """The one full-size run of the unmodified reference that is feasible (config 2, 1 GB: tens of minutes), measured once
  in the build container and committed with its outputs' digests (profiles/r2_reference_config2_full.json)."""
  try:
    return json.load(open(os.path.join(ROOT, "profiles", "r2_reference_config2_full.json")))
  except Exception:
    return None


# ----------------------------------------------------------------------------- workload
def make_piece(spec: synth.CorpusSpec, rank: int, world: int, gather_sizes):
  """This rank's byte range of the workload corpus in PINNED host memory: (torch tensor, numpy view, global offset).
  The corpus is one fixed byte string whatever `world` is (synth.generate_chunks / cut_piece)."""
  import torch
  t0 = time.perf_counter()
  c0, c1 = synth.piece_chunk_range(spec, rank, world)
  chunks = list(synth.generate_chunks(spec, c0, c1))
  uncut = int(sum(c.size for c in chunks))
  sizes = gather_sizes(uncut)
  off, n = synth.cut_piece(spec, sizes, rank)
  host = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
  arr = host.numpy()[:n]
  pos = 0
  for c in chunks:
    take = min(c.size, n - pos)
    if take <= 0:
      break
    arr[pos: pos + take] = c[:take]
    pos += take
  assert pos == n
  if n and off + n == spec.nbytes:
    arr[-1] = ord("\n")
  log(f"[bench] rank {rank}: bytes [{off}, {off + n}) of {spec.name} ({n / 1e9:.3f} GB) generated in {time.perf_counter() - t0:.1f}s")
  return host, arr, off


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
  ap.add_argument("--workload", default=os.environ.get("SWB_BENCH_WORKLOAD", "config3_10GB"), choices=list(TRAIN_KWS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-encode", action="store_true")
  args = ap.parse_args()

  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  STATE["rank"] = rank
  spec = synth.CONFIGS[args.workload]
  KW = TRAIN_KWS[args.workload]
  loop_desc = {"rank0": "the merge loop (one latency chain: every merge waits for the heap decision after the one before) runs on rank 0, "
                        "the merge list is broadcast",
               "replicated": "every rank runs the latency-bound merge loop on all unique words (no per-merge collective)",
               "sharded": "unique words sharded over the ranks, per-merge NCCL all-gather of delta records, replicated frequency table + heap"}[MERGE_LOOP]
  config = {"workload": f"{args.workload}: train vocab {KW['target_vocab_size']} on a {spec.nbytes / 1e9:g} GB synthetic Zipfian "
                        f"{spec.alphabet} corpus ({spec.n_types} word types, s={spec.zipf_s}, seed {spec.seed}), "
                        f"min_pair_freq {KW['min_pair_freq']}, unk_id 0, coverage 0.995",
            "corpus_bytes": spec.nbytes,
            "l2": f"input ({spec.nbytes / 1e9:g} GB) and the word hash table are far larger than the 126 MB L2; no flush between steps",
            "parallelism": "1 GPU" if world == 1 else
                           f"strong scaling: the same {spec.nbytes / 1e9:g} GB corpus split into {world} byte ranges, one per GPU; range-split "
                           f"tokenising + NCCL word-table exchange, then {loop_desc}"}

  # ------------------------------------------------------------------ reference arm
  if args.impl == "reference":
    if rank != 0:
      return
    workdir = tempfile.mkdtemp(prefix="swb_bench_")
    corpus = synth.corpus_bytes(synth.CorpusSpec(spec.name, min(spec.nbytes, CPU_SAMPLE_BYTES + 4096), spec.n_types, spec.alphabet, spec.zipf_s, spec.seed))
    sample, nbytes = cpu_sample_path(corpus, workdir)
    runs = []
    for i in range(args.warmup + args.steps):
      r = run_cpu_once(sample, nbytes, KW)
      log(f"[bench] reference step {i}: {r['seconds']:.2f}s {r['value']:.5f} GB/s")
      if i >= args.warmup:
        runs.append(r)
    secs = float(np.mean([r["seconds"] for r in runs]))
    value = nbytes / 1e9 / secs
    cb = dict(runs[-1]); cb["value"] = value
    full = full_config_reference()
    if full:
      cb["full_config2"] = full
    emit(({
      "impl": "reference", "metric": "train_corpus_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
      "scaling": "strong" if args.gpus > 1 else "weak",
      "vs_baseline": None, "dtype": "int32 symbols, uint64 counts", "data": "synthetic", "config": config,
      "cpu_baseline": cb, "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "extra": {"merges_per_s": runs[-1]["merges_per_s"],
                "note": "reference is single-threaded; each step = full train on the bounded sample (a prefix of the workload corpus): its GB/s falls "
                        "with corpus size (4096-bucket hash maps, full-scan merges), so the sample flatters it"},
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
    from shredword_b200.distributed import DistributedBPETrainer, gather_token_offsets

  def gather_sizes(mine: int) -> list[int]:
    if world == 1:
      return [mine]
    t = torch.tensor([mine], dtype=torch.int64, device=dev)
    out = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]

  STATE["phase"] = "corpus"
  host, arr, goff = make_piece(spec, rank, world, gather_sizes)
  nbytes = arr.size
  d_corpus = host[:max(nbytes, 1)].to(dev, non_blocking=False)
  total_bytes = spec.nbytes
  golden = golden_digest(args.workload)

  def new_trainer():
    if world > 1:
      return DistributedBPETrainer(**KW, device=dev, merge_loop=MERGE_LOOP)
    return BPETrainer(**KW)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  last = {}
  digests = set()

  retried = []

  def step(resident: bool, timing: bool):
    STATE["step"] += 1
    w0 = time.perf_counter()
    old = last.pop("trainer", None)
    if old is not None:
      old.destroy()  # the previous step's handle: its device memory goes back before the next step allocates
    for attempt in range(3):
      w1 = time.perf_counter()
      t = new_trainer()
      t.set_kernel_timing(timing)
      if world > 1:
        t.load_shard(d_corpus[:nbytes] if resident else arr, goff)
      elif resident:
        t.load_device(d_corpus.data_ptr(), nbytes)
      else:
        t.load_buffer(arr)
      w2 = time.perf_counter()
      try:
        merges = t.train_quiet()
        break
      except RuntimeError as e:
        # A failed training step is repeated from scratch INSIDE the timed region (its cost stays in the number) and is
        # reported in the JSON line (`retried_steps`); the library marks the handle failed, nothing of it is reused.
        log(f"[bench] rank {rank} step {STATE['step']} attempt {attempt}: training failed, repeating the step: {e}")
        retried.append({"step": STATE["step"], "error": str(e)[:600]})
        t.destroy()
        if attempt == 2:
          raise
    w3 = time.perf_counter()
    m = t.merges_array()  # the step's result, read on the host
    st = t.stats()
    st["merges"] = merges; st["merge_bytes"] = m.nbytes
    st["wall_ms"] = {"destroy_prev": (w1 - w0) * 1e3, "create+load": (w2 - w1) * 1e3, "train": (w3 - w2) * 1e3,
                     "read": (time.perf_counter() - w3) * 1e3}
    digests.add((merges, hashlib.md5(np.ascontiguousarray(m, dtype="<i4").tobytes()).hexdigest()))
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
  STATE["phase"] = "warmup"
  for i in range(args.warmup):
    s = step(True, False)
    log(f"[bench] rank {rank} warmup {i}: merges={s['merges']} load={s['load_ms']:.1f}ms count={s['count_ms']:.1f}ms merge={s['merge_ms']:.1f}ms")
  step(False, False)  # e2e warm-up (pinned registration, first H2D)

  STATE["phase"] = "timed resident"
  ms_res, wall_res, st_res = timed(True, args.steps, timing=False)
  STATE["phase"] = "timed e2e"
  ms_e2e, wall_e2e, st_e2e = timed(False, args.steps, timing=False)
  clocks = sampler.stop()
  # parity gate: every step of this run (warm-up included, every rank) produced the golden merge list
  STATE["phase"] = "parity check"
  if len(digests) != 1:
    raise RuntimeError(f"steps of one run disagree on the merge list: {sorted(digests)}")
  n_merges_seen, md5_seen = next(iter(digests))
  parity = {"merges": n_merges_seen, "model_md5": md5_seen, "golden": None, "equal": None}
  if golden:
    parity["golden"] = {k: golden[k] for k in ("merges", "model_md5", "source")}
    parity["equal"] = (golden["merges"] == n_merges_seen and golden["model_md5"] == md5_seen)
    if not parity["equal"]:
      raise RuntimeError(f"merge list differs from the golden digest of {args.workload}: got {n_merges_seen} merges md5 {md5_seen}, "
                         f"golden {golden['merges']} merges md5 {golden['model_md5']} ({golden['source']})")
  if world > 1:  # every rank holds the same list
    box = [None] * world
    dist.all_gather_object(box, md5_seen)
    if len(set(box)) != 1:
      raise RuntimeError(f"ranks disagree on the merge list: {box}")

  # per-launch check: a few of the same steps with one merge_rows launch per merge, each bracketed by CUDA events (kept
  # out of `value`: the event records add latency to a latency-bound loop)
  STATE["phase"] = "per-launch pass"
  ms_tim, _, st_tim = timed(True, min(args.steps, 2), timing=True)

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
    n = sum(s["merges"] for s in stats if s["merge_kernel_ms"] > 0)
    return alg, kms, n
  alg, kms, nl = kernel_numbers(st_res)
  resident = sum(s.get("resident_local_merges", 0) + s.get("resident_grid_merges", 0) for s in st_res) > 0
  if not resident:  # (the sharded multi-GPU loop launches merge_rows per merge: use the CUDA-event pass)
    alg, kms, nl = kernel_numbers(st_tim)
  alg_t, kms_t, nl_t = kernel_numbers(st_tim)
  achieved = alg / 1e9 / (kms / 1e3) if kms > 0 else None
  traffic = actual = None
  try:  # dram bytes per merge of merge_cluster from this round's ncu capture of a scripted (host-free) launch of it
    tj = json.load(open(os.path.join(ROOT, "profiles", "r2_merge_cluster_traffic.json")))
    traffic = tj.get("dram_bytes_per_merge"); actual = tj
  except Exception:
    pass
  roofline = {"bound": "hbm", "kernel": "merge_cluster (resident, per merge)" if resident else "merge_rows (per launch)",
              "achieved": achieved, "peak": peak, "unit": "GB/s",
              "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
              "alg_bytes_per_launch": alg / nl if nl else None,
              "avg_launch_us": kms * 1e3 / nl if nl else None, "launches": nl,
              "kernel_share_of_step": (kms / args.steps) / ms_res if (ms_res and resident) else None,
              "actual_dram": ({"bytes_per_merge": traffic, "GB_per_s": traffic / (kms * 1e3 / nl) / 1e3 if (traffic and nl and kms) else None,
                               "frac_of_peak": (traffic / (kms * 1e3 / nl) / 1e3 / peak) if (traffic and nl and kms) else None,
                               "source": actual.get("source") if actual else None}),
              "resident_split": {"spilled_local_merges": sum(s.get("resident_spill_merges", 0) for s in st_res),
                                 "local_by_log_entries(<=512,<=4096,<=32768,more)": {
                                   "merges": [sum(s["local_by_log"][i] for s in st_res) for i in range(4)],
                                   "us_per_merge": [sum(s["local_by_log_ms"][i] for s in st_res) * 1e3 / max(1, sum(s["local_by_log"][i] for s in st_res)) for i in range(4)],
                                   "records_per_merge": [sum(s["local_by_log_recs"][i] for s in st_res) / max(1, sum(s["local_by_log"][i] for s in st_res)) for i in range(4)]},
                                 "local_merges": sum(s.get("resident_local_merges", 0) for s in st_res),
                                 "grid_merges": sum(s.get("resident_grid_merges", 0) for s in st_res),
                                 "local_us_per_merge": (sum(s.get("resident_local_ms", 0) for s in st_res) * 1e3 /
                                                        max(1, sum(s.get("resident_local_merges", 0) for s in st_res))),
                                 "grid_us_per_merge": (sum(s.get("resident_grid_ms", 0) for s in st_res) * 1e3 /
                                                       max(1, sum(s.get("resident_grid_merges", 0) for s in st_res)))} if resident else None,
              "per_launch_check": {"kernel": "merge_rows (one launch per merge, CUDA events)",
                                   "achieved": alg_t / 1e9 / (kms_t / 1e3) if kms_t > 0 else None,
                                   "avg_launch_us": kms_t * 1e3 / nl_t if nl_t else None, "launches": nl_t},
              "note": "NOMINAL figure: algorithmic bytes = the reference's full scan, 4*S_live + 8*W per merge (SURVEY.md 8(d)), over the kernel's own "
                      "per-merge device time. The kernel is index-driven and latency-bound: it moves only `traffic` bytes per merge (actual_dram), "
                      "so frac says how far a merge is from one full scan at copy speed, not how busy the DRAM is."}
  tok_ms = [s["tokenize_ms"] for s in st_res if s.get("tokenize_ms")]
  tok = None
  if tok_ms:
    tms = float(np.mean(tok_ms))
    tok = {"kernel": "wt_tokenize", "ms": tms, "bytes": int(st_res[-1]["tokenize_bytes"]), "GB_per_s": st_res[-1]["tokenize_bytes"] / 1e9 / (tms / 1e3),
           "hbm_frac": st_res[-1]["tokenize_bytes"] / 1e9 / (tms / 1e3) / peak, "alg_bytes": "1 B read per corpus byte"}

  def phase(stats, key):
    return float(np.mean([s[key] for s in stats]))
  out = {
    "metric": "train_corpus_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
    "dtype": "int32 symbols, uint64 counts", "data": "synthetic", "config": config, "clocks": clocks,
    "e2e": {"value": e2e_value, "unit": "GB/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(total_bytes),
            "d2h_bytes_per_step": int(st_e2e[-1]["merge_bytes"])},
    "gpu_launches": launches, "roofline": roofline, "parity": parity, "retried_steps": retried,
    "extra": {"merges": merges, "merges_per_s": merges / (phase(st_res, "merge_ms") / 1e3) if (rank == 0 and phase(st_res, "merge_ms")) else None,
              "us_per_merge": phase(st_res, "merge_ms") * 1e3 / max(merges, 1),
              "phase_ms": {k: phase(st_res, k) for k in ("load_ms", "count_ms", "merge_ms")},
              "e2e_phase_ms": {k: phase(st_e2e, k) for k in ("load_ms", "count_ms", "merge_ms")},
              "tokenize": tok,
              "wall_ms": [s_["wall_ms"] for s_ in st_res[-3:]], "e2e_wall_ms": [s_["wall_ms"] for s_ in st_e2e[-3:]],
              "unique_words": st_res[-1]["words"], "rows": st_res[-1]["rows"],
              "look_ahead": {k: st_res[-1].get(k) for k in ("hints_sent", "hints_taken", "hints_rejected", "host_peek_ms")},
              "host_split_ms": {k: st_res[-1].get(k) for k in ("host_pop_ms", "host_wait_ms", "host_apply_ms")},
              "merge_loop": MERGE_LOOP if world > 1 else "single GPU",
              "collectives_per_step": st_res[-1].get("collectives"), "exchange_bytes_per_step": st_res[-1].get("exchange_bytes"),
              "exchange_ms": float(np.mean([s_.get("exchange_ns", 0) for s_ in st_res])) / 1e6,  # (inside load_ms: sizes + word arenas + metadata all-gathers, merge of the N tables)
              "wall_ms_per_step": wall_res,
              "ms_per_step_with_kernel_timing": ms_tim},
  }

  # ------------------------------------------------------------------ encoder (config 4: document-parallel, every N)
  if not args.no_encode:
    STATE["phase"] = "encode"
    enc = last["trainer"].encoder()
    d_out = torch.empty(max(nbytes, 1), dtype=torch.int32, device=dev)
    h_out = torch.empty(nbytes // 2 + 16, dtype=torch.int32, pin_memory=True)
    ksteps = max(2, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    ntok = enc.encode_device(d_corpus.data_ptr(), nbytes, d_out.data_ptr(), d_out.numel())  # first call of a fresh encoder: empty word memo
    e1.record(); barrier()
    enc_cold_ms = e0.elapsed_time(e1)
    ntok = enc.encode_device(d_corpus.data_ptr(), nbytes, d_out.data_ptr(), d_out.numel())
    barrier()
    l0 = enc.kernel_launches
    e0.record()
    for _ in range(ksteps):
      ntok = enc.encode_device(d_corpus.data_ptr(), nbytes, d_out.data_ptr(), d_out.numel())
      if world > 1:
        tok_off, tok_total, _ = gather_token_offsets(ntok, None, dev)  # the path's one collective: where this rank's ids sit
    e1.record(); barrier()
    enc_ms = e0.elapsed_time(e1) / ksteps
    enc.encode_into(arr[: min(nbytes, 64 << 20)], h_out.numpy())
    barrier()
    e0.record()
    for _ in range(ksteps):
      ntok2 = enc.encode_into(arr, h_out.numpy())
      if world > 1:
        gather_token_offsets(ntok2, None, dev)
    e1.record(); barrier()
    enc_e2e_ms = e0.elapsed_time(e1) / ksteps
    tot_tok = ntok
    if world > 1:
      tt = torch.tensor([enc_ms, enc_e2e_ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      enc_ms, enc_e2e_ms = float(tt[0].item()), float(tt[1].item())
      tn = torch.tensor([ntok], dtype=torch.int64, device=dev); dist.all_reduce(tn)
      tot_tok = int(tn.item())
    # parity of the encoding, size-independent: histogram(encode(training corpus)) == the trainer's own token histogram
    hist = torch.bincount(d_out[:ntok].to(torch.int64), minlength=256 + merges)
    if world > 1:
      dist.all_reduce(hist)
    tf = torch.from_numpy(last["trainer"].token_freq().astype(np.int64)).to(dev)
    hist_equal = bool(torch.equal(hist[: tf.numel()], tf)) and int(hist[tf.numel():].sum().item()) == 0
    if not hist_equal:
      raise RuntimeError("histogram of encode(training corpus) differs from the trainer's token histogram (.vocab frequency column)")
    out["gpu_launches"] += int(enc.kernel_launches - l0)
    if world > 1:
      tt = torch.tensor([enc_cold_ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); enc_cold_ms = float(tt.item())
    out["extra"]["encode"] = {"MB_per_s": total_bytes / 1e6 / (enc_ms / 1e3), "e2e_MB_per_s": total_bytes / 1e6 / (enc_e2e_ms / 1e3),
                              "first_call_MB_per_s": total_bytes / 1e6 / (enc_cold_ms / 1e3),
                              "tokens": int(tot_tok), "bytes_per_token": total_bytes / max(tot_tok, 1),
                              "alg_bytes_per_input_byte": 1 + 4 * tot_tok / total_bytes,
                              "hbm_frac": (total_bytes / world + 4 * tot_tok / world) / 1e9 / (enc_ms / 1e3) / peak,
                              "e2e_h2d_bytes": int(total_bytes), "e2e_d2h_bytes": int(4 * tot_tok),
                              "document_parallel": f"{world} ranks, byte ranges cut on line ends, token counts all-gathered" if world > 1 else "1 GPU",
                              "histogram_equals_vocab_column": hist_equal}
    del d_out

  # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
  if world == 1 and rank == 0 and not args.no_cpu_baseline:
    STATE["phase"] = "cpu baseline"
    workdir = tempfile.mkdtemp(prefix="swb_bench_")
    sample, sb = cpu_sample_path(arr, workdir)
    cb = run_cpu_once(sample, sb, KW)
    log(f"[bench] cpu baseline ({cb['kind']}): {cb['seconds']:.1f}s on {sb} bytes -> {cb['value']:.5f} GB/s")
    # like for like: our arm on exactly that sample, end to end from host memory
    t = BPETrainer(**KW)
    for _ in range(3):
      t.destroy(); t = BPETrainer(**KW)
      torch.cuda.synchronize(); t0 = time.perf_counter()
      t.load_buffer(arr[:sb]); got = t.train_quiet(); t.merges_array()
      same_s = time.perf_counter() - t0
    sst = t.stats()
    cb["ours_on_same_sample"] = {"seconds": same_s, "GB_per_s": sb / 1e9 / same_s, "merges": got, "speedup": cb["seconds"] / same_s,
                                 "merges_equal_count": got == cb["merges"],
                                 "phase_ms": {k: sst[k] for k in ("load_ms", "count_ms", "merge_ms")},
                                 "note": "end to end from host memory incl. handle creation; a 32 MB corpus leaves the GPU mostly idle (fixed costs dominate)"}
    t.destroy()
    full = full_config_reference()
    if full:
      cb["full_config2"] = full
    out["cpu_baseline"] = cb

  if rank == 0:
    emit(out)
  STATE["phase"] = "shutdown"
  last.pop("trainer").destroy()
  if world > 1:
    lib.swb_dist_shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
  try:
    main()
  except SystemExit:
    raise
  except BaseException as e:  # the LAST stderr line names the rank, the phase and the library's own message
    traceback.print_exc()
    err = ""
    try:
      from shredword_b200.cbase import last_error
      err = last_error()
    except Exception:
      pass
    print(f"[bench] FAILED rank {STATE['rank']} phase '{STATE['phase']}' step {STATE['step']}: {type(e).__name__}: {e} | swb_last_error: {err!r}",
          file=sys.stderr, flush=True)
    os._exit(1)
