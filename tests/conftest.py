import importlib
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_c0_state():
    return np.load(os.path.join(GOLDEN_DIR, "c0_ver2_n2000_s10.npz"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/): test infrastructure, never the thing under test on -m gpu."""
    from oracle import oracle as O
    O.build(with_ref=os.path.isdir("/root/reference"))
    return O


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds libnbx.so / nbody.x in-tree if they are missing."""
    p = importlib.import_module("nbody-demo-2023_b200")
    if not (os.path.exists(p.LIB_PATH) and os.path.exists(p.CLI_PATH) and os.path.exists(p.CLI_ALL_PATH)):
        p.build()
    return p


@pytest.fixture(scope="session")
def nbx(pkg):
    return pkg.nbx


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
