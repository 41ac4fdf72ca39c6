import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_ref():
    """The unmodified reference, compiled by oracle/Makefile (prebuilt file on the GPU box)."""
    from oracle import ref
    if not ref.have_ref():
        if os.path.isdir("/root/reference/src"):
            ref.build()
        else:
            pytest.skip("oracle/_ref/libqrdm_ref.so not present and /root/reference absent")
    ref.load_ref()
    return ref


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import ref
    ref.load_port()
    return ref
