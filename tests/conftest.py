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


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a kernel that never returns) must fail on its own instead of taking the whole run with it:
    every gpu-marked test gets a generous timeout (pytest-timeout, when installed)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            # method "thread": the main thread of a hung test sits inside a CUDA call and never returns to the interpreter
            item.add_marker(pytest.mark.timeout(300, method="thread"))


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
