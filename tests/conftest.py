"""pytest configuration: markers, shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol checks, gloo multi-process.
`-m gpu`       : parity tests proper -- CUDA path through the C-ABI vs the oracle / golden vectors.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libspmvref.so not built (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def poisson2d():
    text = open(os.path.join(GOLDEN, "poisson2D.mtx")).read()
    b = np.array([float(t) for t in open(os.path.join(GOLDEN, "poisson2D_b.txt")).read().split()])
    z = np.array([float(t) for t in open(os.path.join(GOLDEN, "poisson2D_result.txt")).read().split()])
    return text, b, z


@pytest.fixture(scope="session")
def ref_vectors():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


def l2norm(v):
    v = np.asarray(v, dtype=np.float64)
    return float(np.sqrt(np.dot(v, v)))
