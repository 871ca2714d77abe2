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


@pytest.fixture(scope="session")
def fso():
    """the CPU oracle (test infrastructure; see oracle/fs_oracle.c header)"""
    from oracle import fso as _fso
    _fso.lib()
    return _fso


@pytest.fixture(scope="session")
def ref_meshes():
    return np.load(os.path.join(GOLDEN, "ref_meshes.npz"))


def load_ref_mesh(fso, z, name):
    m = fso.Mesh(z[name + "/xyz"], z[name + "/etype"], z[name + "/eptr"], z[name + "/enodes"], z[name + "/bc"])
    F = z[name + "/forces"] if (name + "/forces") in z.files else None
    return m, F
