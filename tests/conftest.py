import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT / "openlifu-python_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lifu_lib():
    """Build (if stale) and load liblifusim.so; GPU tests call through this C ABI only."""
    import __graft_entry__ as g
    g.build()
    from openlifu_b200 import _lib
    return _lib
