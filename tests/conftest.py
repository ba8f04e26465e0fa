import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def pf():
    return importlib.import_module("permafrost-engine_b200")


@pytest.fixture(scope="session")
def pforacle():
    import pforacle as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def pfref():
    """The compiled reference (oracle/_ref/libpfref.so); built here when /root/reference is present."""
    import pfref as m
    if not m.available():
        if os.path.isdir("/root/reference/src"):
            import subprocess
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    if not m.available():
        pytest.skip("oracle/_ref/libpfref.so not built and /root/reference absent")
    m.lib()
    return m


@pytest.fixture(scope="session")
def nav(pf):
    n = pf.capi.Nav(0)
    yield n
    n.close()
