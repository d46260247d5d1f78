import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    g = np.load(ROOT / "tests" / "golden" / "reference_host.npz")
    meta = json.loads((ROOT / "tests" / "golden" / "reference_host.json").read_text())
    return g, meta
