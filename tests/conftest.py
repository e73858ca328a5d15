import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native artefacts exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "cramore_b200", "libcramore_cuda.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        g.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    from cramore_b200 import DemuxEngine
    eng = DemuxEngine(0)  # raises without an sm_100 GPU: there is no fallback
    yield eng
    eng.close()
