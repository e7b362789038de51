import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CODES = os.path.join(ROOT, "codes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the CUDA extension and the oracle are compiled (no GPU needed)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def code225():
    from exp_ldpc_b200.codes import load_code
    return load_code(os.path.join(CODES, "hgp_34_n225.qecc"))


@pytest.fixture(scope="session")
def code1600():
    from exp_ldpc_b200.codes import load_code
    return load_code(os.path.join(CODES, "hgp_34_n1600.qecc"))


@pytest.fixture(scope="session")
def code6400():
    from exp_ldpc_b200.codes import load_code
    return load_code(os.path.join(CODES, "hgp_34_n6400.qecc"))
