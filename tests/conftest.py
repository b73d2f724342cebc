import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """The native artefacts must already be built (python -c 'import __graft_entry__ as g; g.build()');
    when the toolchain is present they are (re)built here so a fresh checkout is testable."""
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "flipengine3d_b200", "libflip_b200.so")
    if not os.path.exists(lib):
        g.build()
    return True
