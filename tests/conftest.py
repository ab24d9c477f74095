import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import json
    from oracle import pyref as o
    out = {}
    for sid, S in o.SUITES.items():
        with open(os.path.join(ROOT, "tests", "golden", f"{S.name}_thin.json")) as f:
            out[sid] = json.load(f)
    return out
