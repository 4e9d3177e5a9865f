import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

REF_LIB = ROOT / "oracle" / "_ref" / "libhagrid_ref.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from hagrid_b200 import Library
        return Library().device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def lib():
    """The product library; GPU tests fail loudly (not skip) if it cannot run."""
    from hagrid_b200 import Library
    l = Library()
    assert l.impl == "hagrid_b200"
    assert l.device_count() > 0, "no CUDA device visible: -m gpu tests must run on the GPU box"
    return l


@pytest.fixture(scope="session")
def ref_lib():
    """The reference rebuilt for sm_100a behind the same C ABI (oracle/build_ref.sh)."""
    from hagrid_b200 import Library
    if not REF_LIB.exists():
        pytest.skip("oracle/_ref/libhagrid_ref.so not built (run oracle/build_ref.sh where /root/reference exists)")
    l = Library(REF_LIB)
    assert l.impl == "reference"
    return l
