import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "labrador-snark_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """CUDA context of the product library; GPU tests are skipped (not passed) when no device exists."""
    import labrador_b200 as lb
    try:
        c = lb.Context(0)
    except lb.LabError as e:
        if e.status == 5:
            pytest.skip(f"no CUDA device: {e}")
        raise
    yield c
    c.close()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle
