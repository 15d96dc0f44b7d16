import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One GPU context for the whole session (generator tables are cached in it)."""
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import build
    build.build_lib()
    c = bpg.Context(0)
    yield c
    c.close()
