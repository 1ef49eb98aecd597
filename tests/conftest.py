"""pytest configuration: markers, import paths, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "vision-processor_b200", "python"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import oracle as O
    return O.Oracle("port")


@pytest.fixture(scope="session")
def clref():
    import oracle as O
    if not O.have_reference():
        pytest.skip("oracle/_ref/libvp_clref.so not built (needs /root/reference at build time)")
    return O.Oracle("reference")


@pytest.fixture(scope="session")
def ctx():
    from vpb200 import lib
    c = lib.Context(0)  # raises without a GPU: there is no fallback
    yield c
    c.close()
