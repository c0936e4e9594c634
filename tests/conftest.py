import os
import sys

import numpy as np
import pytest
from scipy.sparse import csr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The unmodified reference, installed by baseline/install_ref.sh (travels to the GPU box with the snapshot), and the
# stand-in for its missing `hyperopt` dependency: with them on the path recpack_b200's classes subclass the
# reference's own (recpack_b200/_ref.py) and the reference's Pipeline can drive them.
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
HAVE_REF = os.path.isdir(os.path.join(REF_DIR, "recpack"))
if HAVE_REF:
    for p in (os.path.join(ROOT, "baseline", "stubs"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)


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
