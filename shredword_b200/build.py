"""Builds shredword_b200/libtrainer.so in-tree with nvcc for sm_100a (B200).

The library is the product: C++ host core + hand-written CUDA kernels behind the C-ABI of
include/shredword_b200.h. It is named `libtrainer.so` and sits in the package directory because
that is where the reference's ctypes loader looks (reference shredword/cbase.py:5-19).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "capi.cu")
OUT = os.path.join(PKG, "libtrainer.so")
DEPS = [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))] + [
  os.path.join(ROOT, "include", "shredword_b200.h")]


def nvcc_path() -> str:
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise FileNotFoundError("nvcc not found")


def needs_build() -> bool:
  if not os.path.exists(OUT):
    return True
  t = os.path.getmtime(OUT)
  return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
  if not force and not needs_build():
    return OUT
  cmd = [
    nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-shared", "-cudart", "static",
    "-o", OUT + ".tmp", SRC,
  ]
  if os.environ.get("SWB_KERNEL_TRACE"):
    cmd.insert(1, "-DSWB_KERNEL_TRACE")
  if os.environ.get("SWB_DEBUG_BOUNDS"):  # index checks inside the merge kernels (device_util.cuh), reported by the host after a failure
    cmd.insert(1, "-DSWB_DEBUG_BOUNDS")
  if verbose:
    cmd.insert(1, "-Xptxas=-v")
    print(" ".join(cmd), file=sys.stderr)
  subprocess.run(cmd, check=True)
  os.replace(OUT + ".tmp", OUT)
  return OUT


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
