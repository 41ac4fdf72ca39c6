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


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Table of every reference comparison: exact, or how many blocks the margin rule excluded."""
    try:
        import parity
    except Exception:
        return
    if not parity.REPORT:
        return
    tr = terminalreporter
    tr.section("parity vs the unmodified reference (tests/parity.py::graded_check)")
    for e in parity.REPORT:
        tr.write_line(f"{e['case']:42s} {str(tuple(e['shape'])):18s} {e['family']:9s} {e['mode']:22s} "
                      f"blocks {e['blocks_trusted']}/{e['blocks_total']} excluded {e['blocks_excluded']}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        import json
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_report.json"), "w") as f:
            json.dump(parity.REPORT, f, indent=1)
    except OSError:
        pass
