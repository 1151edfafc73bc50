import gzip
import json
import sys
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with gzip.open(HERE / "golden" / "ref_cases.json.gz", "rt") as fh:
        cases = json.load(fh)["cases"]
    fixtures = np.load(HERE / "golden" / "fixtures.npz")
    return cases, fixtures
