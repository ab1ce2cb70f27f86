import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """Build the CPU oracle once (port always; _ref only where /root/reference exists)."""
    from oracle import oracle_kernel
    if not os.path.exists(oracle_kernel.PORT_LIB) or (os.path.isdir("/root/reference") and not oracle_kernel.have_ref()):
        oracle_kernel.build()
    return oracle_kernel
