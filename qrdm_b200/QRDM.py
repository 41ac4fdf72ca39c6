"""Python module ``QRDM`` — the same five functions, argument order and in-place semantics as the
reference's CPython extension (reference QRDM_wrapper.c:131-139), so that the reference's
``test.ipynb`` / ``auxil.py`` run unchanged with ``from qrdm_b200 import QRDM``.

* ``QRDM(matrix_layout, m, n, A, lda, jpvt, tau, ncols, thres, nb)``  (QRDM_wrapper.c:71-101)
  calls the C-ABI entry point ``dgeqrdm`` of libqrdm_b200.so on the NumPy buffers: the B200 path.
* ``QP3``, ``QRF``, ``DORMQR`` (QRDM_wrapper.c:15-69, 104-126) stay host LAPACK pass-throughs, as
  in the reference (they are the CPU comparator / verification helpers).
* ``init`` (QRDM_wrapper.c:9-13) is a no-op kept for compatibility.

The reference casts ``->data`` unchecked (QRDM_wrapper.c:155-159); here dtype and contiguity are
verified and a TypeError is raised instead of reading garbage.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def init():
    return None


def _buf(x, dtype, name):
    if not isinstance(x, np.ndarray) or x.dtype != dtype:
        raise TypeError(f"{name} must be a numpy array of dtype {np.dtype(dtype).name}")
    if not (x.flags.c_contiguous or x.flags.f_contiguous):
        raise TypeError(f"{name} must be contiguous")
    if not x.flags.writeable:
        raise TypeError(f"{name} must be writeable")
    return x.ctypes.data


def _colmajor_view(A, m, n, lda):
    """m x n column-major view (leading dimension lda) of A's raw buffer, as LAPACK sees it."""
    flat = A.reshape(-1, order="A")
    return np.lib.stride_tricks.as_strided(flat, shape=(m, n), strides=(8, 8 * lda), writeable=True)


def QRDM(matrix_layout, m, n, A, lda, jpvt, tau, ncols, thres, nb):
    """QR with Deviation Maximization pivoting on the GPU; returns info (QRDM_wrapper.c:100)."""
    pa = _buf(A, np.float64, "A")
    pj = _buf(jpvt, np.int32, "jpvt")
    pt = _buf(tau, np.float64, "tau")
    pn = _buf(ncols, np.int32, "ncols")
    ph = _buf(thres, np.float64, "thres")
    if m > 0 and n > 0 and (A.size < lda * (n - 1) + m or jpvt.size < n or tau.size < min(m, n)
                            or thres.size < 2 or ncols.size < min(m, n)):
        # ncols: the driver writes one entry per iteration, up to min(m, n) of them (reference doc: dimension N)
        raise ValueError("array too small for the given m, n, lda (ncols needs min(m, n) entries)")
    if thres.size < 3 and ncols.size > 0 and int(ncols.flat[0]) == 3:
        raise ValueError("stop mode 3 reads thres[2]")
    return int(_lib.lib.dgeqrdm(int(matrix_layout), int(m), int(n), pa, int(lda), pj, pt, pn, ph, int(nb)))


def QP3(matrix_layout, m, n, A, lda, jpvt, tau):
    """LAPACK dgeqp3 on the host (QRDM_wrapper.c:15-41)."""
    from scipy.linalg import lapack
    if matrix_layout != 102:
        raise NotImplementedError("only matrix_layout=102 is supported")
    _buf(A, np.float64, "A")
    _buf(jpvt, np.int32, "jpvt")
    _buf(tau, np.float64, "tau")
    V = _colmajor_view(A, m, n, lda)
    qr, p, t, work, info = lapack.dgeqp3(np.asfortranarray(V))
    V[...] = qr
    jpvt[:n] = p
    tau[: min(m, n)] = t
    return int(info)


def QRF(matrix_layout, m, n, A, lda, tau):
    """LAPACK dgeqrf on the host (QRDM_wrapper.c:44-69)."""
    from scipy.linalg import lapack
    if matrix_layout != 102:
        raise NotImplementedError("only matrix_layout=102 is supported")
    _buf(A, np.float64, "A")
    _buf(tau, np.float64, "tau")
    V = _colmajor_view(A, m, n, lda)
    qr, t, work, info = lapack.dgeqrf(np.asfortranarray(V))
    V[...] = qr
    tau[: min(m, n)] = t
    return int(info)


def DORMQR(matrix_layout, m, n, k, A, lda, tau, Cmat, ldc):
    """LAPACK dormqr('L','N') on the host: C <- Q C with the first k reflectors of A
    (QRDM_wrapper.c:104-126)."""
    from scipy.linalg import lapack
    if matrix_layout != 102:
        raise NotImplementedError("only matrix_layout=102 is supported")
    _buf(A, np.float64, "A")
    _buf(tau, np.float64, "tau")
    _buf(Cmat, np.float64, "C")
    V = _colmajor_view(A, m, k, lda)
    Cv = _colmajor_view(Cmat, m, n, ldc)
    a = np.asfortranarray(V)
    c = np.asfortranarray(Cv)
    lwork = int(lapack.dormqr("L", "N", a, tau[:k], c, -1)[1][0])
    cq, work, info = lapack.dormqr("L", "N", a, tau[:k], c, lwork)
    Cv[...] = cq
    return int(info)
