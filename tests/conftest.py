import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _oracle_pow_is_libm_again():
    """the oracle's pow is a process-global: whatever a test swapped in, the next test starts from the libm pow"""
    yield
    try:
        from oracle import binding as ob
        if ob._lib is not None:
            ob.set_pow(None)
            ob.set_id(None)
    except Exception:
        pass
