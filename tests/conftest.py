import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "hotpath_v1.npz"))


@pytest.fixture(scope="session")
def blosum():
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", "blosum62.npz"))
    return d["mat"], d["pback"]


@pytest.fixture(scope="session")
def oracle(blosum):
    from oracle.pyoracle import Oracle, build
    build(with_ref=False)
    return Oracle(*blosum)


@pytest.fixture(scope="session")
def built_lib():
    from mmseqs2_b200 import build as b
    return b.build()


@pytest.fixture(scope="session")
def submat(blosum, built_lib):
    from mmseqs2_b200 import SubMatrix
    return SubMatrix(*blosum)


@pytest.fixture(scope="session")
def ctx(built_lib):
    from mmseqs2_b200 import Context
    c = Context(0)
    yield c
    c.close()
