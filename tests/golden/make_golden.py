"""Generates tests/golden/* by running the UNMODIFIED reference (oracle/_ref, zero-filling malloc) in this
container, where /root/reference exists. The fixtures are what pins the oracle (and through it the CUDA
path) to the reference: `<case>.model` and `<case>.vocab` are the reference's own output bytes.

  python tests/golden/make_golden.py        # needs `make -C oracle ref`

`manifest.json` records, per case, the sha256 of the corpus the files were generated from, so a change of
the corpus generator is detected instead of silently comparing against fixtures of another input."""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import oracle as O  # noqa: E402

# negative_unk is excluded: the reference's bpe_save writes freq[-1] (heap corruption)
GOLDEN_CASES = [c for c in cases.CASES if c != "negative_unk"]


def main():
  O.build(ref=True)
  assert O.ref_available(), "oracle/_ref is not built"
  manifest = {}
  with tempfile.TemporaryDirectory() as td:
    for name in GOLDEN_CASES:
      data = cases.corpus(name)
      kw = cases.kwargs(name)
      p = os.path.join(td, "c.txt")
      open(p, "wb").write(data)
      tm = O.run_reference(p, kw["target_vocab_size"], kw["min_pair_freq"], os.path.join(HERE, name + ".model"),
                           os.path.join(HERE, name + ".vocab"), unk_id=kw.get("unk_id", 0),
                           coverage=kw.get("character_coverage", 0.995))
      manifest[name] = {"corpus_sha256": hashlib.sha256(data).hexdigest(), "corpus_bytes": len(data), "kwargs": kw,
                        "merges": tm["merges"], "generated_by": "unmodified reference, oracle/_ref/ref_driver"}
      print(name, tm["merges"], "merges")
  json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
  main()
