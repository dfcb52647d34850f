import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kats():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def iris():
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", "iris.npz"))
    return {"S": d["S"], "C": d["C"], "F": d["F"], "names": [str(x) for x in d["names"]],
            "classes": [str(x) for x in d["classes"]]}
