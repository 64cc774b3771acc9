from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are kept out of history): build the product
    library, the microbenchmark and the test oracle once.  The package itself never does this --
    it fails loudly when its library is missing."""
    missing = [p for p in (ROOT / "taxi2_b200" / "lib" / "libtaxi2_b200.so", ROOT / "oracle" / "libtaxi_oracle.so") if not p.exists()]
    if missing:
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return GOLDEN
