import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def kind_of(g):
    if int(g["is_complex"]):
        return 1
    return 3 if int(g["block"]) == 3 else 0


@pytest.fixture(scope="session")
def golden():
    return load_golden


SYSTEMS = ["poisson_h1p3", "elasticity_h1p4_dim3", "maxwell_hcurlp2", "helmholtz_h1p4_complex",
           "shifted_laplace_complex", "square_h1p4_testsolvers"]


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
