import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


import pytest


@pytest.fixture(params=["fused", "rows"])
def modtable_variant(request, monkeypatch):
    """Runs a GPU test once per modification-table variant (JTK_MODTABLE): 'fused' keeps the DP matrices on chip, 'rows' parks
    the forward rows in HBM.  The library reads the variable at every call."""
    monkeypatch.setenv("JTK_MODTABLE", request.param)
    return request.param
