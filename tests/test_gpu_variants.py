"""GPU: every environment-selected code path of the library reproduces the oracle bit for bit, and the overflow paths of
the resident merge kernel (which only trigger on their own at 1 GB scale) are FORCED on small corpora by shrinking the
kernel's limits (SWB_TEST_*: ClusterTune in cluster_kernel.cuh, the frequency table's initial size). The switches are
read once per process, so each variant runs tests/variant_check.py in a subprocess."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["ascii_ties", "multi_unk97", "config1_10MB", "self_pairs"]

VARIANTS = {
  "default": {},
  "per_launch_kernel": {"SWB_NO_PERSISTENT": "1"},
  "no_hints": {"SWB_NO_HINTS": "1"},
  "no_early_hints": {"SWB_NO_EARLY_HINTS": "1"},
  "no_initial_pair_index": {"SWB_NO_IP_INDEX": "1"},
  "tiny_initial_pair_lists": {"SWB_IP_LOCAL_MAX": "4"},
  "no_birth_log": {"SWB_NO_BIRTH_LOG": "1", "SWB_NO_PERSISTENT": "1"},
  "host_frequency_table": {"SWB_HOST_TABLE": "1"},
  "no_load_pipeline": {"SWB_NO_LOAD_PIPELINE": "1"},
  "small_load_pieces": {"SWB_LOAD_PIECE": "65536", "SWB_ENCODE_PIECE": "65536"},
  "no_solo_merges": {"SWB_TEST_SOLO_MAX": "0"},          # every LOCAL merge through the whole leader cluster (DSMEM exchange)
  "solo_only_tiny_logs": {"SWB_TEST_SOLO_MAX": "40"},
  # forced overflow paths of merge_cluster
  "grid_from_long_logs": {"SWB_TEST_LOCAL_MAX": "64"},
  "candidate_overflow": {"SWB_TEST_CAND_CAP": "3"},
  "inbox_spill": {"SWB_TEST_INBOX": "2"},
  "table_spill": {"SWB_TEST_MAX_PROBES": "1"},
  "record_stage_overflow": {"SWB_TEST_REC_STAGE": "1"},
  "birth_stage_overflow": {"SWB_TEST_BIRTH_STAGE": "1"},
  "frequency_table_rehash": {"SWB_TEST_SMALL_GT": "1"},
  "word_table_growth": {"SWB_TEST_WT_CAP": "256"},       # the load's hash table starts far too small: overflow, 4x, tokenise again (several times)
  "solo_table_spill": {"SWB_TEST_MAX_PROBES": "1", "SWB_TEST_REC_STAGE": "1", "SWB_TEST_BIRTH_STAGE": "2", "SWB_TEST_CAND_CAP": "4"},
  "everything_small": {"SWB_TEST_LOCAL_MAX": "4096", "SWB_TEST_SOLO_MAX": "300", "SWB_TEST_CAND_CAP": "5", "SWB_TEST_INBOX": "7", "SWB_TEST_MAX_PROBES": "3",
                       "SWB_TEST_REC_STAGE": "2", "SWB_TEST_BIRTH_STAGE": "3", "SWB_TEST_SMALL_GT": "1"},
}


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_variant_matches_oracle(variant, product, oracle_mod):
  env = {k: v for k, v in os.environ.items() if not k.startswith("SWB_")}
  env.update(VARIANTS[variant])
  p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_check.py")] + CASES, capture_output=True, text=True, env=env, timeout=900)
  assert p.returncode == 0, f"{variant}:\n{p.stderr[-3000:]}"
  st = json.loads(p.stdout.strip().splitlines()[-1])
  big = st["config1_10MB"]
  resident = big["resident_local_merges"] + big["resident_grid_merges"]
  if variant in ("per_launch_kernel", "no_birth_log", "host_frequency_table"):
    assert resident == 0 and big["merge_launches"] >= big["merges"]          # one launch per merge
  else:
    assert resident == big["merges"] and big["merge_launches"] < big["merges"]  # the resident kernel served every merge
  if variant == "no_hints":
    assert big["hints_sent"] == 0 and big["hints_taken"] == 0
  if variant == "grid_from_long_logs":
    assert big["resident_grid_merges"] > st["config1_10MB"]["merges"] // 4
