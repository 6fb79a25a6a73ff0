import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def valence():
    return load_golden("valence_silicon.npz")


@pytest.fixture(scope="session")
def silicon():
    return load_golden("silicon_12to8.npz")


@pytest.fixture(scope="session")
def silicon_split():
    return load_golden("silicon_split.npz")
