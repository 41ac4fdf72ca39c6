"""Convenience layer over the C ABI (no numerics here)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _params(n, thres, stop_mode):
    ncols = np.zeros(max(n, 1), dtype=np.int32)
    ncols[0] = stop_mode
    th = np.zeros(3, dtype=np.float64)
    th[: len(thres)] = thres
    return ncols, th


def dgeqrdm(A, thres=(0.9, 0.15), nb=64, stop_mode=0, layout=102, lda=None, inplace=False):
    """Factor a host matrix through the reference-facing entry point ``dgeqrdm`` (host pointers;
    H2D/D2H inside the call).  Returns dict(info, A, jpvt, tau, ncols) like the oracle loaders."""
    A = np.asarray(A, dtype=np.float64)
    if not (inplace and A.flags.f_contiguous):
        A = np.array(A, dtype=np.float64, order="F", copy=True)
    m, n = A.shape
    jpvt = np.zeros(n, dtype=np.int32)
    tau = np.zeros(min(m, n), dtype=np.float64)
    ncols, th = _params(n, thres, stop_mode)
    info = _lib.lib.dgeqrdm(int(layout), m, n, A.ctypes.data, int(m if lda is None else lda),
                            jpvt.ctypes.data, tau.ctypes.data, ncols.ctypes.data, th.ctypes.data, int(nb))
    return dict(info=int(info), A=A, jpvt=jpvt, tau=tau, ncols=ncols)


def dgeqrdm_device(dA, m, n, lda, d_jpvt, d_tau, thres=(0.9, 0.15), nb=64, stop_mode=0, stream=None):
    """Factor a device-resident column-major matrix in place (``dgeqrdm_dev``).  ``dA``, ``d_jpvt``
    (int32, n) and ``d_tau`` (float64, min(m,n)) are torch CUDA tensors (or raw device pointers);
    returns (info, ncols) with ncols a host int32 array."""
    def ptr(x):
        return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)
    ncols, th = _params(n, thres, stop_mode)
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    info = _lib.lib.dgeqrdm_dev(int(m), int(n), ptr(dA), int(lda), ptr(d_jpvt), ptr(d_tau),
                                ncols.ctypes.data, th.ctypes.data, int(nb), C.c_void_p(int(stream)))
    return int(info), ncols


def dgeqrdm_batched(As, thres=(0.9, 0.15), nb=64, stop_mode=0):
    """Factor a batch of equally-shaped host matrices (``As``: array-like (batch, m, n)) through
    ``dgeqrdm_batched``.  Returns dict(info, infos, A (batch, m, n), jpvt, tau, ncols)."""
    As = np.asarray(As, dtype=np.float64)
    batch, m, n = As.shape
    buf = np.empty((batch, n, m), dtype=np.float64)          # each [b] is column-major m x n
    buf[...] = As.transpose(0, 2, 1)
    jpvt = np.zeros((batch, n), dtype=np.int32)
    tau = np.zeros((batch, min(m, n)), dtype=np.float64)
    ncols = np.zeros((batch, n), dtype=np.int32)
    ncols[:, 0] = stop_mode
    infos = np.zeros(batch, dtype=np.int32)
    th = np.zeros(3, dtype=np.float64)
    th[: len(thres)] = thres
    info = _lib.lib.dgeqrdm_batched(batch, m, n, buf.ctypes.data, m, m * n, jpvt.ctypes.data, tau.ctypes.data,
                                    ncols.ctypes.data, th.ctypes.data, int(nb), infos.ctypes.data)
    return dict(info=int(info), infos=infos, A=buf.transpose(0, 2, 1), jpvt=jpvt, tau=tau, ncols=ncols)


def dgeqrdm_batched_device(batch, m, n, d_a, lda, stride_a, d_jpvt, d_tau, d_ncols, d_infos=0, thres=(0.9, 0.15), nb=64,
                           stream=0):
    """Device-resident batch (raw device pointers as ints, e.g. ``tensor.data_ptr()``): one launch of the
    one-CTA-per-matrix kernel.  ``d_ncols`` is [batch][n] int32 with the stop mode in [b][0] on entry."""
    th = np.zeros(3, dtype=np.float64)
    th[: len(thres)] = thres
    return int(_lib.lib.dgeqrdm_batched_dev(int(batch), int(m), int(n), C.c_void_p(int(d_a)), int(lda), int(stride_a),
                                            C.c_void_p(int(d_jpvt)), C.c_void_p(int(d_tau)), C.c_void_p(int(d_ncols)),
                                            C.c_void_p(int(d_infos)) if d_infos else None, th.ctypes.data, int(nb),
                                            C.c_void_p(int(stream)) if stream else None))


def dormqr(F, tau, Cm, k=None, trans="N"):
    """C <- Q C (trans "N") or Q' C ("T") on the GPU; F = factored matrix (host), reflectors in its first k
    columns (default len(tau)); returns (info, new C).  ``qrdm_b200_dormqr`` (host pointers)."""
    F = np.asfortranarray(F, dtype=np.float64)
    Cm = np.array(Cm, dtype=np.float64, order="F", copy=True)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    m = F.shape[0]
    k = len(tau) if k is None else int(k)
    assert Cm.shape[0] == m and k <= F.shape[1]
    info = _lib.lib.qrdm_b200_dormqr(trans.encode(), m, Cm.shape[1], k, F.ctypes.data, m, tau.ctypes.data,
                                     Cm.ctypes.data, m)
    return int(info), Cm


def low_rank(F, tau, k=0):
    """Rank-k approximation from a rank-revealing factorisation, the reference's ``auxil.low_rank``
    (auxil.py:156-200) with the Q application on the GPU:  A P ~ A_k = Q [R11 R12; 0 0], R11 k x k.
    F, tau: outputs of dgeqrdm (column-major convention); k = 0 or k > min(m, n) means min(m, n).
    Returns A_k (m x n, columns in pivoted order)."""
    F = np.asfortranarray(F, dtype=np.float64)
    m, n = F.shape
    if k == 0 or k > min(m, n):
        k = min(m, n)
    R = np.zeros((m, n), order="F")
    R[:k, :] = np.triu(F)[:k, :]
    info, Ak = dormqr(F, tau, R, k=k, trans="N")
    if info != 0:
        raise RuntimeError(f"qrdm_b200_dormqr failed: info={info}")
    return Ak


def dormqr_device(trans, m, n, k, dA, lda, d_tau, dC, ldc, stream=None):
    """Device-resident Q application (``qrdm_b200_dormqr_dev``): torch CUDA tensors or raw device pointers."""
    def ptr(x):
        return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    return int(_lib.lib.qrdm_b200_dormqr_dev(trans.encode(), int(m), int(n), int(k), ptr(dA), int(lda), ptr(d_tau),
                                             ptr(dC), int(ldc), C.c_void_p(int(stream))))


def dgeqp3(A):
    """QR with classical column pivoting on the GPU (LAPACK dgeqp3's blocked algorithm): the comparator the
    reference's wrapper exposes as QP3 (QRDM_wrapper.c:15-41).  Returns dict(info, A, jpvt (1-based), tau)."""
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    F = np.asfortranarray(A.copy())
    jpvt = np.zeros(n, dtype=np.int32)
    tau = np.zeros(min(m, n), dtype=np.float64)
    info = _lib.lib.qrdm_b200_dgeqp3(m, n, F.ctypes.data, m, jpvt.ctypes.data, tau.ctypes.data)
    return dict(info=int(info), A=F, jpvt=jpvt, tau=tau)


def dgeqp3_device(dA, m, n, lda, d_jpvt, d_tau, stream=None):
    """Device-resident variant: dA is an (n, lda) row-major float64 CUDA tensor (= column-major m x n)."""
    def ptr(x):
        return x.data_ptr() if hasattr(x, "data_ptr") else int(x)
    return int(_lib.lib.qrdm_b200_dgeqp3_dev(m, n, ptr(dA), lda, ptr(d_jpvt), ptr(d_tau), stream or 0))


def stats():
    return _lib.stats()


def set_profile(mode: int):
    """0 off; 1 every stage timed (syncs after each stage); 2 light: panel + trailing only, no syncs."""
    _lib.lib.qrdm_b200_set_profile(int(mode))


def fp64_peak(use_dmma=True, stream=0):
    """Measured FP64 peak of this GPU in TFLOP/s (DMMA.8x8x4 or DFMA chains on every SM)."""
    return float(_lib.lib.qrdm_b200_measure_fp64_peak(1 if use_dmma else 0, C.c_void_p(int(stream))))


def copy_gbs(nbytes=1 << 30, stream=0):
    """Measured device-to-device copy bandwidth in GB/s (bytes read + written per second)."""
    return float(_lib.lib.qrdm_b200_measure_copy_gbs(int(nbytes), C.c_void_p(int(stream))))
