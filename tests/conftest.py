import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref_lib():
    """oracle/_ref (the reference's own compiled sources); built on demand when /root/reference exists."""
    from oracle import ref
    if not ref.available():
        try:
            ref.build()
        except Exception:
            pass
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return ref
