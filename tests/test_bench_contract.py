"""CPU: bench.py's contract where no GPU is needed -- the reference arm prints exactly one JSON line on stdout with the
keys the driver reads, and the product arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(oracle_mod):
  env = dict(os.environ, SWB_BENCH_CPU_SAMPLE_BYTES="3000000")
  p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                     capture_output=True, text=True, env=env, timeout=600)
  assert p.returncode == 0, p.stderr[-2000:]
  lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1, p.stdout[:500]
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "train_corpus_GB_per_s" and d["unit"] == "GB/s" and d["higher_is_better"] is True
  assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
  cb = d["cpu_baseline"]
  assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
  assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert "workload" in d["config"]


def test_product_arm_needs_a_gpu():
  import torch
  if torch.cuda.is_available():
    return  # (this check is for GPU-less machines)
  p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
  assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
  assert p.stdout.strip() == ""  # nothing that could be mistaken for a result
