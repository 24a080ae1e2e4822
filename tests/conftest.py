"""pytest configuration: the `gpu` marker, repo-root imports and shared fixtures."""
import json
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
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle as po
    po.build()
    return po.Oracle("port")


@pytest.fixture(scope="session")
def reflib():
    from oracle import pyoracle as po
    if not po.have_reference():
        pytest.skip("oracle/_ref/liblcg_ref.so not present")
    return po.Oracle("reference")


@pytest.fixture(scope="session")
def fixtures():
    """The reference's data/ fixtures (tests/golden/data) + their Jacobi diagonals."""
    from liblcg_b200 import io as lio
    out = {}
    for name in ("10K", "10Kc", "1Kc"):
        A = lio.load_fixture(name)
        A["diag"] = lio.csr_diagonal(A["row_ptr"], A["col"], A["val"])
        out[name] = A
    return out
