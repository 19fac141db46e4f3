"""Generates tests/golden/normalize_cases.json from the REFERENCE's own normalize_line (oracle/_ref/libtrainer_ref.so,
reference csrc/bpe/normalize.cpp:24-59), and checks the oracle's restatement against every case while doing so.

  python tests/golden/make_normalize_golden.py        # needs `make -C oracle ref`
"""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import oracle as O  # noqa: E402


def main():
  O.build(ref=True)
  rng = np.random.default_rng(3)
  cases = [b"", b" ", b"\t\t", b"Hello World", b"  Hello   World  ", b"HELLO\tWORLD\r", b"a", b" a", b"a ", b"A  b \t C",
           b"\xe2\x96\x81x y", b"x \xe2\x96\x81", b"x\xe2\x96\x81\xe2\x96\x81", b"\xe2\x96\x81", b"caf\xc3\x89 NA\xc3\x8fVE", b"MiXeD CaSe 123 !?",
           b"tab\tsep\tline\t", b"\r\rx\r\ry\r\r", b"x" * 300 + b"  " + b"Y" * 10, b"a\nb", b"\nA\n\nB\n"]
  alpha = np.frombuffer(b"abcXYZ  \t\r\n.,\xc3\xa9\xe4\xb8\xad\xe2\x96\x81", dtype=np.uint8)
  for _ in range(120):
    cases.append(bytes(rng.choice(alpha, size=int(rng.integers(0, 80)))))
  out = []
  for c in cases:
    r = O.ref_normalize_line(c)
    assert O.normalize_line(c) == r, (c, r, O.normalize_line(c))
    out.append({"in": base64.b64encode(c).decode(), "out": base64.b64encode(r).decode()})
  json.dump({"generated_by": "tests/golden/make_normalize_golden.py: the reference's own normalize_line (oracle/_ref/libtrainer_ref.so, "
                             "csrc/bpe/normalize.cpp:24-59)", "cases": out}, open(os.path.join(HERE, "normalize_cases.json"), "w"), indent=0)
  print(len(out), "cases; the oracle's restatement agrees on all of them")


if __name__ == "__main__":
  main()
