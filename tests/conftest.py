import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
  """The CPU oracle (test infrastructure). Builds oracle/liboracle.so on first use."""
  import oracle as O
  O.build(ref=False)
  return O


@pytest.fixture(scope="session")
def product():
  """The product package; builds shredword_b200/libtrainer.so if it is missing (needs nvcc)."""
  from shredword_b200 import build as B
  B.build()
  import shredword_b200.trainer as T
  return T
