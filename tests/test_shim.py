"""CPU suite, part 3: the Python `QRDM` module mirrors the reference's extension
(reference QRDM_wrapper.c:131-139): host LAPACK pass-throughs and argument validation."""
import numpy as np
import pytest
import scipy.linalg as sla

from qrdm_b200 import generators as g


@pytest.fixture(scope="module")
def Q():
    from qrdm_b200 import QRDM
    return QRDM


def test_module_surface(Q):
    for name in ("init", "QRF", "QP3", "QRDM", "DORMQR"):
        assert callable(getattr(Q, name))
    assert Q.init() is None


def test_qp3_qrf_dormqr_roundtrip(Q):
    """The notebook's usage (test.ipynb cells 2-6): a C-ordered array passed with layout 102,
    i.e. the routine factors X^T; auxil.checkQR's residual definitions (auxil.py:92-101)."""
    n = 48
    X = np.ascontiguousarray(g.graded(n, seed=5))
    A = X.copy()
    jpvt = np.zeros(n, dtype=np.int32)
    tau = np.zeros(n)
    assert Q.QP3(102, n, n, A, n, jpvt, tau) == 0
    Qm = np.eye(n)
    assert Q.DORMQR(102, n, n, n, A, n, tau, Qm, n) == 0
    Qm = Qm.T                      # column-major buffer seen as C-order
    R = np.triu(A.T)
    assert np.linalg.norm(np.eye(n) - Qm.T @ Qm) < 1e-13
    assert np.linalg.norm(X.T[:, jpvt - 1] - Qm @ R) < 1e-13
    # QRF agrees with scipy
    B = X.copy()
    tau2 = np.zeros(n)
    assert Q.QRF(102, n, n, B, n, tau2) == 0
    qr_ref, tau_ref = sla.qr(X.T, mode="raw")[0]
    assert np.allclose(B.T, qr_ref) and np.allclose(tau2, tau_ref)


def test_qrdm_rejects_wrong_dtypes(Q):
    A = np.zeros((4, 4), order="F")
    with pytest.raises(TypeError):
        Q.QRDM(102, 4, 4, A, 4, np.zeros(4, dtype=np.int64), np.zeros(4), np.zeros(4, dtype=np.int32),
               np.array([0.9, 0.15]), 64)
    with pytest.raises(TypeError):
        Q.QRDM(102, 4, 4, A.astype(np.float32), 4, np.zeros(4, dtype=np.int32), np.zeros(4),
               np.zeros(4, dtype=np.int32), np.array([0.9, 0.15]), 64)
    with pytest.raises(ValueError):
        Q.QRDM(102, 8, 8, A, 8, np.zeros(8, dtype=np.int32), np.zeros(8), np.zeros(8, dtype=np.int32),
               np.array([0.9, 0.15]), 64)


def test_generators_shapes_and_flops():
    assert g.gaussian(5, 3, 0).flags.f_contiguous
    K = g.kahan(6)
    assert np.allclose(np.tril(K, -1), 0) and K[0, 0] == 1.0
    X = g.graded(32, seed=0)
    s = np.linalg.svd(X, compute_uv=False)
    assert (s > 5e-3).sum() == 15  # r-1 of the shifted singular values, r = 16
    assert g.flops(10, 10, 10) == pytest.approx(2 * 10 * 100 - 2 / 3 * 1000)
