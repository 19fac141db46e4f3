"""Builds shredword_b200/libtrainer.so in-tree with nvcc for sm_100a (B200).

The library is the product: C++ host core + hand-written CUDA kernels behind the C-ABI of
include/shredword_b200.h. It is named `libtrainer.so` and sits in the package directory because
that is where the reference's ctypes loader looks (reference shredword/cbase.py:5-19).
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "capi.cu")
OUT = os.path.join(PKG, "libtrainer.so")
STAMP = OUT + ".srcdigest"  # written next to the library; travels with it
DEPS = [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))] + [
  os.path.join(ROOT, "include", "shredword_b200.h")]


def nvcc_path() -> str:
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise FileNotFoundError("nvcc not found")


def source_digest() -> str:
  """Digest of everything the library is compiled from (file names + contents) and of the build switches."""
  h = hashlib.sha256()
  for d in sorted(DEPS):
    h.update(os.path.basename(d).encode() + b"\0")
    with open(d, "rb") as f:
      h.update(f.read())
  h.update(repr(sorted((k, os.environ.get(k, "")) for k in ("SWB_KERNEL_TRACE", "SWB_DEBUG_BOUNDS"))).encode())
  return h.hexdigest()


def needs_build() -> bool:
  """Stale = the sources' digest differs from the one recorded next to the library when it was built. (Modification times
  do not survive a copy of the tree to another machine.)"""
  if not os.path.exists(OUT) or not os.path.exists(STAMP):
    return True
  try:
    return open(STAMP).read().strip() != source_digest()
  except OSError:
    return True


def build(force: bool = False, verbose: bool = False) -> str:
  if not force and not needs_build():
    return OUT
  # one builder at a time (N ranks of one torchrun call this concurrently): the others wait, then find the library current
  with open(OUT + ".lock", "w") as lock:
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
      if not force and not needs_build():
        return OUT
      digest = source_digest()
      tmp = f"{OUT}.tmp.{os.getpid()}"
      cmd = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
        "-shared", "-cudart", "static",
        "-o", tmp, SRC,
      ]
      if os.environ.get("SWB_KERNEL_TRACE"):
        cmd.insert(1, "-DSWB_KERNEL_TRACE")
      if os.environ.get("SWB_DEBUG_BOUNDS"):  # index checks inside the merge kernels (device_util.cuh), reported by the host after a failure
        cmd.insert(1, "-DSWB_DEBUG_BOUNDS")
      if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
      try:
        subprocess.run(cmd, check=True)
        os.replace(tmp, OUT)
      finally:
        if os.path.exists(tmp):
          os.remove(tmp)
      with open(STAMP + f".{os.getpid()}", "w") as f:
        f.write(digest + "\n")
      os.replace(STAMP + f".{os.getpid()}", STAMP)
    finally:
      fcntl.flock(lock, fcntl.LOCK_UN)
  return OUT


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
