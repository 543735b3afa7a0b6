import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU oracle runs (set LAGB_SLOW=1 to enable)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("LAGB_SLOW", "0") != "1":
        skip = pytest.mark.skip(reason="slow oracle run: set LAGB_SLOW=1")
        for it in items:
            if "slow" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build (if needed) the product library and the oracle; both are in-tree."""
    import subprocess
    from laghos_b200._lib import LIB_PATH
    from pyoracle import LIB_PATH as ORC
    if not (os.path.exists(LIB_PATH) and os.path.exists(ORC)):
        subprocess.check_call(["make", "-j8", "all"], cwd=ROOT)
    return True
