import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name starts with a digit, so importlib)."""
    return importlib.import_module("4dflownet_b200")


@pytest.fixture(scope="session")
def oracle():
    return importlib.import_module("oracle.sr4d_oracle")
