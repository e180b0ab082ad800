import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def core():
    """The product module. Built in-tree; import failure is an error, not a skip."""
    import trimem_b200.core as core
    return core


@pytest.fixture(scope="session")
def gpu_core(core):
    if core.device_count() < 1:
        pytest.fail("test is marked gpu but no CUDA device is visible (there is no CPU fallback)")
    return core


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle restatement (oracle/trimem_oracle.c), compiled on demand with gcc."""
    import oracle_port
    oracle_port.lib()
    return oracle_port


@pytest.fixture(scope="session")
def refmod():
    """The reference build (oracle/_ref) when its .so travelled with the tree, else None."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    try:
        import trimem.core as ref
        return ref
    except ImportError:
        return None
