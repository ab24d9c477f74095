import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The tests load the in-tree libraries (libavrf_gpu.so, the host emulation, the C oracle); build whatever is
    missing or older than its sources (a no-op when everything is current; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden():
    import json
    from oracle import pyref as o
    out = {}
    for sid, S in o.SUITES.items():
        with open(os.path.join(ROOT, "tests", "golden", f"{S.name}_thin.json")) as f:
            out[sid] = json.load(f)
    return out
