"""pytest plumbing: the `gpu` marker, and import paths for the hyphenated package directory."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rtl-sdr-scanner-cpp_b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def load_b2s():
    """The package directory name contains '-', so it is loaded by path."""
    if "b2s" in sys.modules:
        return sys.modules["b2s"]
    spec = importlib.util.spec_from_file_location("b2s", os.path.join(PKG, "b2s.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2s"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def b2s():
    return load_b2s()


@pytest.fixture(scope="session")
def engine(b2s):
    """One engine per session. Fails loudly (no CPU fallback) when the CUDA library or the GPU is missing."""
    e = b2s.Engine(0)
    yield e
    e.close()
