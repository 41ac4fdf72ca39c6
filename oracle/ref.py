"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the two CPU checkers.

* ``load_ref()``  -> ``oracle/_ref/libqrdm_ref.so``: the unmodified reference
  (`/root/reference/src/{dlarfb,dlarf,dlarfg,dgeqr2,dgeqp3,dgeqrdm_work,dgeqrdm}.c`,
  the list in the reference's ``setup_QRDM.py:20-30``) compiled by ``oracle/Makefile``.
* ``load_port()`` -> ``oracle/_build/libqrdm_port.so``: our plain-C restatement
  (``oracle/qrdm_port.c``), no BLAS, with decision-margin logging.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference/cpu_baseline
legs may import this module.  The product package ``qrdm_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libqrdm_ref.so")
REF_TIMED_SO = os.path.join(HERE, "_ref", "libqrdm_ref_timed.so")
PORT_SO = os.path.join(HERE, "_build", "libqrdm_port.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(quiet: bool = True) -> None:
    """Run ``make -C oracle`` (port always; reference only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_ref = None
_port = None


def load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle` where /root/reference exists")
        lib = C.CDLL(REF_SO)
        for name in ("dgeqrdm", "dgeqrdm_work"):
            f = getattr(lib, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _dp, _ip, _dp, C.c_int]
        lib.dgeqp3.restype = C.c_int
        lib.dgeqp3.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _dp]
        _ref = lib
    return _ref


def load_port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build()
        lib = C.CDLL(PORT_SO)
        lib.qrdm_port_dgeqrdm.restype = C.c_int
        lib.qrdm_port_dgeqrdm.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _dp, _ip, _dp,
                                          C.c_int, _dp]
        _port = lib
    return _port


def set_ref_threads(n: int) -> None:
    """Set the thread count of the OpenBLAS the reference .so is linked against."""
    lib = load_ref()
    import glob
    import scipy
    path = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs",
                                  "libscipy_openblas*.so"))[0]
    ob = C.CDLL(path)
    ob.scipy_openblas_set_num_threads(C.c_int(int(n)))
    del lib


def _prep(A, thres, stop_mode):
    A = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    m, n = A.shape
    jpvt = np.zeros(n, dtype=np.int32)
    tau = np.zeros(min(m, n), dtype=np.float64)
    ncols = np.zeros(max(n, 1), dtype=np.int32)
    ncols[0] = stop_mode
    th = np.zeros(3, dtype=np.float64)
    th[: len(thres)] = thres
    return A, m, n, jpvt, tau, ncols, th


def ref_dgeqrdm(A, thres=(0.9, 0.15), nb=64, stop_mode=0, layout=102, lda=None):
    """Reference dgeqrdm (include/QRDM.h:19-22) on a copy of A.  Returns dict of outputs."""
    lib = load_ref()
    A, m, n, jpvt, tau, ncols, th = _prep(A, thres, stop_mode)
    info = lib.dgeqrdm(layout, m, n, A.ctypes.data_as(_dp), m if lda is None else lda,
                       jpvt.ctypes.data_as(_ip), tau.ctypes.data_as(_dp), ncols.ctypes.data_as(_ip),
                       th.ctypes.data_as(_dp), nb)
    return dict(info=info, A=A, jpvt=jpvt, tau=tau, ncols=ncols)


def port_dgeqrdm(A, thres=(0.9, 0.15), nb=64, stop_mode=0, layout=102, lda=None):
    """Plain-C restatement; additionally returns per-iteration decision margins
    (``margins[it]`` = smallest relative margin of any data-dependent decision in block it)."""
    lib = load_port()
    A, m, n, jpvt, tau, ncols, th = _prep(A, thres, stop_mode)
    margins = np.full((max(n, 1), 5), np.inf, dtype=np.float64)
    info = lib.qrdm_port_dgeqrdm(layout, m, n, A.ctypes.data_as(_dp), m if lda is None else lda,
                                 jpvt.ctypes.data_as(_ip), tau.ctypes.data_as(_dp),
                                 ncols.ctypes.data_as(_ip), th.ctypes.data_as(_dp), nb,
                                 margins.ctypes.data_as(_dp))
    return dict(info=info, A=A, jpvt=jpvt, tau=tau, ncols=ncols, margins=margins)


_ref_timed = None


def have_ref_timed() -> bool:
    return os.path.exists(REF_TIMED_SO)


def ref_dgeqrdm_timed(A, thres=(0.9, 0.15), nb=64, stop_mode=0):
    """The reference again, built with call-site timers (oracle/ref_timing_shim.c, sources untouched): returns the
    usual outputs plus ``split`` = seconds in dgeqr2_mia / LAPACKE_dlarft / LAPACKE_dlarfb_mia and the total."""
    global _ref_timed
    import time
    if _ref_timed is None:
        lib = C.CDLL(REF_TIMED_SO)
        lib.dgeqrdm.restype = C.c_int
        lib.dgeqrdm.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _dp, _ip, _dp, C.c_int]
        lib.qt_reset.restype = None
        lib.qt_get.restype = None
        lib.qt_get.argtypes = [_dp]
        _ref_timed = lib
    lib = _ref_timed
    A, m, n, jpvt, tau, ncols, th = _prep(A, thres, stop_mode)
    lib.qt_reset()
    t0 = time.perf_counter()
    info = lib.dgeqrdm(102, m, n, A.ctypes.data_as(_dp), m, jpvt.ctypes.data_as(_ip), tau.ctypes.data_as(_dp),
                       ncols.ctypes.data_as(_ip), th.ctypes.data_as(_dp), nb)
    total = time.perf_counter() - t0
    t = np.zeros(3)
    lib.qt_get(t.ctypes.data_as(_dp))
    split = {"dgeqr2_mia (panel)": float(t[0]), "LAPACKE_dlarft (T factor)": float(t[1]),
             "LAPACKE_dlarfb_mia (trailing update)": float(t[2]),
             "DM_perm + norm_update + driver (remainder)": float(total - t.sum()), "total": float(total)}
    return dict(info=info, A=A, jpvt=jpvt, tau=tau, ncols=ncols, split=split)


def ref_dgeqp3(A):
    """Reference's dgeqp3 copy (src/dgeqp3.c:39-93) = LAPACKE_dgeqp3 without the NaN check."""
    lib = load_ref()
    A = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    m, n = A.shape
    jpvt = np.zeros(n, dtype=np.int32)
    tau = np.zeros(min(m, n), dtype=np.float64)
    info = lib.dgeqp3(102, m, n, A.ctypes.data_as(_dp), m, jpvt.ctypes.data_as(_ip), tau.ctypes.data_as(_dp))
    return dict(info=info, A=A, jpvt=jpvt, tau=tau)
