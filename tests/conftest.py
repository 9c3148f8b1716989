import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def mp3():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("swift-mp3_b200")


@pytest.fixture(scope="session")
def orc():
    import oracle_binding
    oracle_binding.lib()
    return oracle_binding
