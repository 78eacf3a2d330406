import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """Session-wide device context; GPU tests only.  Fails loudly without a GPU/extension."""
    from rustqip_b200.state import Context
    c = Context(0)
    yield c
    c.close()
