"""CPU suite, part 2: the C-ABI library loads, exports every symbol the header declares, checks
its arguments like the reference does, and has NO CPU fallback (no compute calls succeed here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from qrdm_b200 import _lib
    return _lib.lib


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "qrdm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", hdr)
    names = [n for n in names if not n.startswith("__")]
    assert {"dgeqrdm", "dgeqrdm_work", "dgeqrdm_dev"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qrdm_b200.h but not exported"
    from qrdm_b200 import _lib
    assert set(_lib.EXPORTS) <= set(names)


def _call(lib, m, n, lda=None, thres=(0.9, 0.15), nb=64, layout=102, jpvt0=0, entry="dgeqrdm"):
    A = np.zeros((max(m, 1), max(n, 1)), order="F")
    jpvt = np.full(max(n, 1), jpvt0, dtype=np.int32)
    tau = np.zeros(max(min(m, n), 1))
    ncols = np.zeros(max(n, 1), dtype=np.int32)
    th = np.array(list(thres) + [0.0] * (3 - len(thres)))
    return getattr(lib, entry)(layout, m, n, A.ctypes.data, m if lda is None else lda, jpvt.ctypes.data,
                               tau.ctypes.data, ncols.ctypes.data, th.ctypes.data, nb)


@pytest.mark.parametrize("kw", [dict(m=0, n=4), dict(m=4, n=0), dict(m=4, n=4, lda=3),
                                dict(m=4, n=4, thres=(1.5, 0.1)), dict(m=4, n=4, thres=(0.5, -0.1)),
                                dict(m=4, n=4, nb=0), dict(m=4, n=4, layout=7), dict(m=4, n=4, layout=101)])
@pytest.mark.parametrize("entry", ["dgeqrdm", "dgeqrdm_work"])
def test_argument_errors_return_minus_one(lib, kw, entry, capfd):
    """reference src/dgeqrdm_work.c:559-581: every argument failure prints and returns -1
    (layout 101 passes the reference's check but never worked there: rejected here)."""
    assert _call(lib, entry=entry, **kw) == -1
    assert "DGEQRDM" in capfd.readouterr().err


def test_unsupported_inputs(lib, capfd):
    assert _call(lib, 4, 4, nb=257) == -102   # QRDM_NB_MAX = 256 (64 < nb <= 256 runs in micro-panels of 64 columns)
    capfd.readouterr()


def test_fixed_columns_take_the_gpu_path(lib, capfd):
    """jpvt != 0 on entry (fixed columns, reference src/dgeqrdm_work.c:592-635) is a supported input since round 2: on a
    box without a GPU it must therefore fail like every other valid call (no CPU fallback), not with -102."""
    import torch
    rc = _call(lib, 4, 4, jpvt0=1)
    assert rc == (0 if torch.cuda.is_available() else -100)
    capfd.readouterr()


def test_no_cpu_fallback(lib, capfd):
    """Without a CUDA device a valid call must fail loudly with QRDM_ERR_CUDA, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _call(lib, 8, 8) == -100
    assert "CUDA error" in capfd.readouterr().err


def _ormqr(lib, trans, m, n, k, lda=None, ldc=None):
    a = np.zeros((max(m, 1), max(k, 1)), order="F")
    c = np.zeros((max(m, 1), max(n, 1)), order="F")
    tau = np.zeros(max(k, 1))
    return lib.qrdm_b200_dormqr(trans, m, n, k, a.ctypes.data, m if lda is None else lda, tau.ctypes.data,
                                c.ctypes.data, m if ldc is None else ldc)


@pytest.mark.parametrize("kw", [dict(trans=b"N", m=0, n=4, k=0), dict(trans=b"N", m=4, n=0, k=2),
                                dict(trans=b"N", m=4, n=4, k=5), dict(trans=b"T", m=4, n=4, k=2, lda=3),
                                dict(trans=b"T", m=4, n=4, k=2, ldc=3)])
def test_dormqr_argument_errors(lib, kw, capfd):
    """The Q application (include/qrdm_b200.h) checks its arguments before touching the device."""
    assert _ormqr(lib, **kw) == -1
    capfd.readouterr()


def test_dormqr_has_no_cpu_fallback(lib, capfd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _ormqr(lib, b"N", 8, 8, 4) == -100
    capfd.readouterr()


def test_product_never_imports_oracle():
    """The shipped package must not reference oracle/ (parity claims are void otherwise)."""
    pkg = os.path.join(ROOT, "qrdm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle loaders", ""), f"{f} mentions oracle/"
