"""CPU: the C-ABI shared library loads, exports every symbol include/*.h declares, keeps the reference's
struct layouts, and fails loudly (never falls back) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
  src = open(os.path.join(ROOT, "include", "shredword_b200.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
  return sorted(set(n for n in names if n.startswith(("swb_", "bpe_", "create_trainer"))))


def test_exports_every_declared_symbol(product):
  from shredword_b200.cbase import lib
  names = _declared_functions()
  assert len(names) >= 40 and "create_trainer" in names and "bpe_merge_batch" in names and "swb_encode" in names
  for n in names:
    assert hasattr(lib, n), f"{n} is declared in include/shredword_b200.h but not exported"


def test_reference_symbols_and_layouts(product):
  from shredword_b200 import cbase
  # the 8 entry points reference cbase.py:44-59 binds
  for n in ("create_trainer", "bpe_trainer_destroy", "bpe_init", "bpe_count_bigrams", "bpe_load_corpus", "bpe_merge_batch",
            "bpe_train", "bpe_save"):
    assert hasattr(cbase.lib, n)
  assert ctypes.sizeof(cbase.BPEConfig) == 24 and cbase.BPEConfig.min_pair_freq.offset == 16 and cbase.BPEConfig.character_coverage.offset == 12
  assert ctypes.sizeof(cbase.HeapEntry) == 24
  T = cbase.Trainer
  assert (T.heap.offset, T.corpus.offset, T.bigram_map.offset, T.next_token.offset, T.num_merges.offset, T.merge_ops.offset) == (24, 48, 72, 88, 96, 104)


def test_create_defaults_and_destroy(product):
  """reference test/bpe_test.cpp:59-94: config copied, defaults applied, num_merges == 0."""
  from shredword_b200 import cbase
  cfg = cbase.BPEConfig(target_vocab_size=1000, unk_id=0, character_coverage=0.0, min_pair_freq=0)
  t = cbase.lib.create_trainer(ctypes.byref(cfg))
  assert t
  c = t.contents
  assert c.config.target_vocab_size == 1000 and abs(c.config.character_coverage - 0.995) < 1e-6 and c.config.min_pair_freq == 2000
  assert c.num_merges == 0 and c.heap.size == 0 and c.heap.cap >= 4096 and bool(c.heap.data)
  cbase.lib.bpe_trainer_destroy(t)
  assert not cbase.lib.create_trainer(None)          # the reference exit()s here; we return NULL
  cbase.lib.bpe_trainer_destroy(None)                # no-op
  assert cbase.lib.bpe_merge_batch(None, 1) == -1 and cbase.lib.bpe_train(None) == -1 and cbase.lib.bpe_load_corpus(None, b"x") == -1


def test_no_gpu_fails_loudly(product, tmp_path):
  from shredword_b200 import cbase
  if cbase.lib.swb_device_count() > 0:
    pytest.skip("a CUDA device is present")
  t = product.BPETrainer(300, min_pair_freq=2)
  p = tmp_path / "c.txt"; p.write_bytes(b"hello world hello\n")
  with pytest.raises(IOError):
    t.load_corpus(str(p))
  assert "no CUDA device" in cbase.last_error() and "no CPU fallback" in cbase.last_error()
  with pytest.raises(IOError):
    t.load_buffer(b"hello world")
  enc = product.BPEEncoder([(104, 101, 256)])
  with pytest.raises(RuntimeError):
    enc.encode(b"hello")
  assert enc.decode([256, 108]) == b"hel"              # decode is a host table lookup


def test_product_never_touches_oracle():
  """The shipped package must not import, link or load anything under oracle/."""
  pkg = os.path.join(ROOT, "shredword_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
        txt = open(os.path.join(dirpath, f), errors="replace").read()
        assert "liboracle" not in txt and "oracle/" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)
  import subprocess
  out = subprocess.run(["ldd", os.path.join(pkg, "libtrainer.so")], capture_output=True, text=True).stdout
  assert "oracle" not in out


def test_reference_python_package_binds_our_library(product, tmp_path):
  """The drop-in claim of INTEGRATION.md: the reference's own cbase.py/trainer.py (symlinked, not copied)
  load our libtrainer.so and drive it. Only where /root/reference exists (this container)."""
  ref = "/root/reference/shredword"
  if not os.path.isdir(ref):
    pytest.skip("/root/reference is not present on this machine")
  import subprocess
  import sys
  pkg = tmp_path / "shredword"
  pkg.mkdir()
  for f in ("__init__.py", "cbase.py", "trainer.py"):
    os.symlink(os.path.join(ref, f), pkg / f)
  os.symlink(os.path.join(ROOT, "shredword_b200", "libtrainer.so"), pkg / "libtrainer.so")
  corpus = tmp_path / "c.txt"; corpus.write_bytes(b"hello world hello\n")
  code = (
    "import sys; sys.path.insert(0, %r)\n"
    "from shredword.trainer import BPETrainer\n"
    "from shredword.cbase import lib\n"
    "t = BPETrainer(target_vocab_size=500, min_pair_freq=1000)\n"
    "assert t.trainer and t.trainer.contents.config.target_vocab_size == 500\n"
    "import ctypes\n"
    "lib.swb_device_count.restype = ctypes.c_int\n"
    "if lib.swb_device_count() == 0:\n"
    "  try:\n"
    "    t.load_corpus(%r); raise SystemExit('expected IOError without a GPU')\n"
    "  except IOError: pass\n"
    "else:\n"
    "  t.load_corpus(%r); t.train(); t.save(%r, %r)\n"
    "t.destroy(); print('BOUND_OK')\n"
  ) % (str(tmp_path), str(corpus), str(corpus), str(tmp_path / "m.model"), str(tmp_path / "m.vocab"))
  r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
  assert "BOUND_OK" in r.stdout, r.stdout + r.stderr


def test_build_staleness_is_decided_by_source_digest_not_mtimes():
  """The tree is copied to the GPU box (modification times change, N ranks import the package at once): the library is
  rebuilt only when the digest of its sources differs from the one recorded next to it, and one builder at a time."""
  import os
  from shredword_b200 import build as B
  B.build()
  assert not B.needs_build()
  stamp = open(B.STAMP).read()
  assert stamp.strip() == B.source_digest()
  now = os.path.getmtime(B.OUT)
  try:
    os.utime(B.OUT, (now - 10_000, now - 10_000))  # an "old" library next to "newer" sources: still current
    assert not B.needs_build()
    open(B.STAMP, "w").write("0" * 64 + "\n")      # another digest: stale
    assert B.needs_build()
  finally:
    open(B.STAMP, "w").write(stamp)
    os.utime(B.OUT, (now, now))
  assert not B.needs_build()


def test_stats_struct_layout_matches_the_header(tmp_path):
  """The ctypes mirror of SwbStats (shredword_b200/cbase.py) against the C declaration in include/shredword_b200.h,
  compiled here with gcc: same size, same offsets of every field."""
  import ctypes
  import subprocess
  from shredword_b200.cbase import SwbStats
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  fields = [k for k, _ in SwbStats._fields_]
  src = tmp_path / "layout.c"
  src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void) { printf("%%zu", sizeof(SwbStats));\n%s\nreturn 0; }\n'
                 % (os.path.join(root, "include", "shredword_b200.h"),
                    "\n".join('printf(" %%zu", offsetof(SwbStats, %s));' % f for f in fields)))
  exe = tmp_path / "layout"
  subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
  got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
  assert got[0] == ctypes.sizeof(SwbStats)
  assert got[1:] == [getattr(SwbStats, f).offset for f in fields]
