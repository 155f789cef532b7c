import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def connectomics():
  with gzip.open(os.path.join(GOLDEN, "connectomics.npy.gz"), "rb") as f:
    return np.load(f)


@pytest.fixture(scope="session")
def fanc():
  """The reference's C-vs-F regression volume (fanc_bug.npy.gz, automated_test.py:17-20): bool (512, 512, 128, 1)."""
  with gzip.open(os.path.join(GOLDEN, "fanc_bug.npy.gz"), "rb") as f:
    return np.load(f)


@pytest.fixture(scope="session")
def ref_cases():
  return np.load(os.path.join(GOLDEN, "ref_cases.npz"))


@pytest.fixture(scope="session")
def build_all():
  """Compile the CUDA library and the C oracle once per session."""
  import __graft_entry__ as g
  g.build()
  return True
