import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# cora / cora_full ship without their feature blobs (reference snapshot: .MISSING_LARGE_BLOBS); the tests
# opt in to the deterministic label-derived features of SURVEY 8(d).  The product loader raises without it.
os.environ.setdefault("EDIS_SYNTH_FEATURES", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
