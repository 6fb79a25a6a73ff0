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


@pytest.fixture(scope="session")
def w():
    """the product package (wannier.jl_b200/, loaded by path because of the dot in its name)"""
    import __graft_entry__ as ge
    return ge.load_package()


def make_ctx(w, st, M, nw, **kw):
    """Context from oracle-layout arrays."""
    nk, nn, nb, _ = M.shape
    return w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M), **kw)


def fixture_stencil(d):
    return dict(kpb_k=d["kpb_k"], bvec=d["bvec"], wb=d["wb"])
