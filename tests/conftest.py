import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer-running CPU test")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """Compile the oracle libraries once per session (a no-op when up to date)."""
    from oracle import pyoracle

    pyoracle.build()
