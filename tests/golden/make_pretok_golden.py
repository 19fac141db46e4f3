"""Generates tests/golden/pretok_cases.json by calling the REFERENCE's own apply_regex (reference shredword/base.py:38-58,
imported from /root/reference in this container) on a fixed set of texts; checks the oracle's restatement on the way.

  python tests/golden/make_pretok_golden.py"""
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import pretok_oracle as P  # noqa: E402
from shredword_b200 import synth  # noqa: E402


def main():
  spec = importlib.util.spec_from_file_location("refbase", "/root/reference/shredword/base.py")
  ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
  texts = ["", " ", "a", "Hello world's  test 12345 it'S  \n\n x!?\n y", "don't I'LL we've they'RE she'd I'm 'tis 'ſ 'x ''s",
           "  leading\tand trailing  ", "\n\n\nline\r\nnext\r\r\n  \n", "tabs\t\tword\t1234567 89.5% (ok)...\n!!!\n\n", "x  \n", "x   ",
           "naïve café ÑANDÚ Straße ΑΒΓ абв 中文字 日本語のテキスト ١٢٣٤٥ ४५६ Ⅻ ½ x²",
           "nbsp here em sp ideographic　sp ls nelvt\x0bff\x0c end", "emoji \U0001f600\U0001f600 ok \U0001f44d\U0001f3fd!", "a1b22c333d4444e55555",
           "price: $1,234.56 -- 50% off!!! (really?) [yes] {no} <maybe> #tag @user", "'s'S'd'D'm'M't'T'll'LL'lL've'VE're'RE'r'l'v",
           " 's", "\t'll", "x's", "1's", "_under_score_ __init__ a_b", "mixed123abc 456def", "CR only\rnext\rlast", "  word", "   \n x",
           " word   word", "x  \ny", "\x1c\x1d kept as they are?"]
  rng = np.random.default_rng(12)
  alpha = list("ab Z9'\n\t\r.!-_ é中٣ s ") + ["'s", "'ll", "  ", "\n\n"]
  for _ in range(150):
    texts.append("".join(rng.choice(alpha, size=int(rng.integers(1, 60)))))
  texts.append(bytes(synth.corpus_bytes(synth.small_spec(20_000, 2_000, 9, "multi"))).decode("utf-8", "ignore"))
  cases = []
  for t in texts:
    want = ref.apply_regex(t)
    assert P.apply_regex(t) == want, t
    assert "".join(want) == t, ("pieces do not cover the text", t)
    if not any(c in t for c in "\x1c\x1d\x1e\x1f"):
      assert P.undo_pretokenize(P.pretokenize_bytes(t.encode("utf-8"))) == t.encode("utf-8")
    cases.append({"text": t, "pieces": want})
  json.dump({"generated_by": "tests/golden/make_pretok_golden.py: the reference's own apply_regex (shredword/base.py:38-58) under regex " +
             __import__("regex").__version__, "cases": cases}, open(os.path.join(HERE, "pretok_cases.json"), "w"), ensure_ascii=True, indent=0)
  print(len(cases), "cases; the oracle's restatement agrees on all of them; pieces always cover the text")


if __name__ == "__main__":
  main()
