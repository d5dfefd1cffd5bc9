import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum-circuits_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the suite must not read artefacts cached by earlier versions of the planner / generator
import tempfile  # noqa: E402
os.environ.setdefault("QSV_CACHE_DIR", tempfile.mkdtemp(prefix="qsv-cache-test-"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when they are selected on a machine without CUDA they must
    # fail loudly (no silent skip / CPU fallback), so nothing is filtered here.
    return


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gdir, "cases.json")) as f:
        data = json.load(f)
    arrays = np.load(os.path.join(gdir, "arrays.npz"))
    return data["meta"], {c["name"]: c for c in data["cases"]}, arrays
