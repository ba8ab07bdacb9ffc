import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_binding():
    """The package directory is called `ldpc-lib_b200` (not an identifier), so load its binding by path."""
    name = "pyldpcb200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "ldpc-lib_b200", "pyldpcb200.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ldpc():
    return load_binding()


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.oracle()
    return pyoracle
