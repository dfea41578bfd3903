import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand from oracle/tpd_oracle.c."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def built_libs():
    """The product libraries; built in-tree by __graft_entry__.build() (nvcc cross-compiles without a GPU)."""
    from torpedo_b200 import _lib
    if not (os.path.exists(_lib.TPDCU_PATH) and os.path.exists(_lib.TPDHOST_PATH)):
        import __graft_entry__ as g
        g.build()
    return _lib
