import os
import sys

import numpy as np
import pytest
from scipy.sparse import csr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def unpack(g, prefix):
    return csr_matrix(
        (g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]),
        shape=tuple(int(x) for x in g[prefix + "_shape"]),
    )


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def gpu_available():
    import torch

    return torch.cuda.is_available()
