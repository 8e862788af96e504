import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def pl_graph(golden):
    """The 60 x 90 power-law graph of the golden file as dense ids + oracle CSR."""
    from oracle import hgr_oracle as O

    u, i = golden["pl_dense_u"], golden["pl_dense_i"]
    n_users, n_items = len(golden["pl_id2user"]), len(golden["pl_id2item"])
    csr = O.build_norm_adj(u, i, n_users, n_items)
    return dict(u=u, i=i, n_users=n_users, n_items=n_items, csr=csr)
