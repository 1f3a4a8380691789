import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "360-image-compression_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle", "_ref")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    import __graft_entry__ as g
    g.build_product()
    return g.LIB


@pytest.fixture(scope="session")
def ref_ext():
    """The unmodified reference extension built for sm_100 (oracle/_ref/lic360_ref*.so), or None."""
    try:
        import lic360_ref
        return lic360_ref
    except Exception:
        return None
