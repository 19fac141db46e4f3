"""Golden vectors for the configurations the bench numbers are quoted on (run in the build container, CPU only).

  python scripts/make_golden_big.py config2      # ~15 min: the UNMODIFIED reference on the full 1 GB corpus
  python scripts/make_golden_big.py config3      # ~1 h:   the pinned oracle on the full 10 GB corpus (the reference would take days)

config2: oracle/_ref/ref_driver (reference sources, zero-filling malloc) -> tests/golden/config2_1GB.{model,vocab}, its
timing -> profiles/r2_reference_config2_full.json; the oracle is run on the same file and must produce identical bytes.
config3: oracle/bpe_oracle.c (byte-identical to the reference on every golden case incl. config 2) ->
tests/golden/config3_10GB.model (390 KB) + the md5 of its .vocab. Both write tests/golden/digests.json, which bench.py
and tests/test_gpu_scale_parity.py compare the CUDA path's merge list against at every N."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
from shredword_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
KW = {"config2_1GB": dict(target_vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000),
      "config3_10GB": dict(target_vocab_size=32768, unk_id=0, character_coverage=0.995, min_pair_freq=2000)}


def md5(path):
  return hashlib.md5(open(path, "rb").read()).hexdigest()


def update(name, entry):
  p = os.path.join(GOLD, "digests.json")
  d = json.load(open(p)) if os.path.exists(p) else {}
  d[name] = entry
  json.dump(d, open(p, "w"), indent=1, sort_keys=True)


def main():
  which = {"config2": "config2_1GB", "config3": "config3_10GB"}[sys.argv[1]]
  work = os.environ.get("SWB_WORK", "/tmp/work"); os.makedirs(work, exist_ok=True)
  corpus = synth.write_corpus(synth.CONFIGS[which], os.path.join(work, which + ".txt"))
  kw = KW[which]
  O.build(ref=True)
  o = O.Oracle(**kw)
  t0 = time.time(); o.load_corpus(corpus); t1 = time.time(); n = o.train(); t2 = time.time()
  om, ov = os.path.join(work, which + "_oracle.model"), os.path.join(work, which + "_oracle.vocab")
  o.save(om, ov)
  oracle_t = {"load_s": t1 - t0, "train_s": t2 - t1, "merges": n, "unique_words": int(o.num_words)}
  if which == "config2_1GB":
    rm, rv = os.path.join(GOLD, which + ".model"), os.path.join(GOLD, which + ".vocab")
    tm = O.run_reference(corpus, kw["target_vocab_size"], kw["min_pair_freq"], rm, rv, unk_id=0, coverage=0.995)
    same = open(rm, "rb").read() == open(om, "rb").read() and open(rv, "rb").read() == open(ov, "rb").read()
    assert same, "oracle and reference disagree on config 2"
    update(which, {"merges": tm["merges"], "model_md5": md5(rm), "vocab_md5": md5(rv),
                   "source": "unmodified reference (oracle/_ref/ref_driver, zero-filling malloc) on the full 1 GB corpus; the oracle produced byte-identical files"})
    json.dump({"what": "one full run of the UNMODIFIED reference on config 2 (1 core)", "load_s": tm["load_s"], "merge_s": tm["merge_s"],
               "seconds": tm["load_s"] + tm["init_s"] + tm["merge_s"], "merges": tm["merges"],
               "GB_per_s": 1.0 / (tm["load_s"] + tm["init_s"] + tm["merge_s"]), "oracle_same_corpus": oracle_t},
              open(os.path.join(ROOT, "profiles", "r2_reference_config2_full.json"), "w"), indent=1)
  else:
    import shutil
    shutil.copy(om, os.path.join(GOLD, which + ".model"))
    update(which, {"merges": n, "model_md5": md5(om), "vocab_md5": md5(ov), "oracle_timing": oracle_t,
                   "source": "oracle/bpe_oracle.c (pinned: byte-identical to the unmodified reference on 16 golden corpora incl. the full 1 GB config 2) on the "
                             "full 10 GB corpus; a full reference run is infeasible (days)"})
  print(which, "done")


if __name__ == "__main__":
  main()
