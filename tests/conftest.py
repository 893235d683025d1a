import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the oracle restates the reference's *sequential* semantics; the compiled reference must run the same way
os.environ.setdefault("OMP_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def product():
    """The CUDA library.  Fails loudly (no fallback) if it was not built."""
    import spasm_b200
    L = spasm_b200.lib()
    L.spasm_b200_set_verbose(0)
    return L


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle
