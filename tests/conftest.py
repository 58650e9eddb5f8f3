"""Shared fixtures.  `-m gpu` tests need a B200 and call the product only through the
C ABI (libifl_b200.so); everything else runs on CPU."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ifl():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("incremental-fluids_b200")


@pytest.fixture(scope="session")
def port():
    from oracle import portapi
    portapi.build()
    return portapi


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
