import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def reference():
    """The imported reference hot path; skips where /root/reference is absent (the GPU box)."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present on this machine')
    return ref_import.load()
