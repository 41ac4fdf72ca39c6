"""GPU suite, QR with classical column pivoting on the device: `qrdm_b200_dgeqp3` against LAPACK dgeqp3 —
the routine behind the reference wrapper's QP3 (QRDM_wrapper.c:15-41, src/dgeqp3.c:39-93).  Same blocked algorithm
(dlaqps: partial-norm downdating with the cancellation test, a panel ends at the first column that fails it), so on
inputs without near-ties among the column norms the pivots are LAPACK's; |diag R| to 1e-10; residual and orthogonality of
the factors by the invariants of tests/parity.py."""
import numpy as np
import pytest
import scipy.linalg as sla

import parity
from qrdm_b200 import generators as g

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU suite selected but no CUDA device: the product has no CPU fallback")
    import qrdm_b200
    return qrdm_b200


def _invariants(A, out):
    m, n = A.shape
    o = dict(A=out["A"], tau=out["tau"], jpvt=out["jpvt"], ncols=np.array([min(m, n)]))
    return parity.qr_invariants(A, o)


CASES = [
    ("gauss300x200", lambda: g.gaussian(300, 200, 8)),
    ("gauss1000", lambda: g.gaussian(1000, 1000, 0)),
    ("gauss700x1900_wide", lambda: g.gaussian(700, 1900, 3)),
    ("gauss2501x601_odd", lambda: g.gaussian(2501, 601, 4)),
    ("gauss64x64", lambda: g.gaussian(64, 64, 5)),
    ("gauss1x17", lambda: g.gaussian(1, 17, 6)),
    ("gauss33x1", lambda: g.gaussian(33, 1, 7)),
    ("gauss5000x130_tall", lambda: g.gaussian(5000, 130, 9)),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_dgeqp3_matches_lapack(name, make, q):
    A = make()
    m, n = A.shape
    out = q.dgeqp3(A)
    assert out["info"] == 0
    qr, jp, tau, _, info = sla.lapack.dgeqp3(np.asfortranarray(A))
    assert info == 0
    assert sorted(out["jpvt"].tolist()) == list(range(1, n + 1))
    assert np.array_equal(out["jpvt"], jp), (name, np.flatnonzero(out["jpvt"] != jp)[:5])
    k = min(m, n)
    assert np.allclose(np.abs(np.diag(out["A"]))[:k], np.abs(np.diag(qr))[:k], rtol=1e-10, atol=0)
    res, orth = _invariants(A, out)
    tol = parity.invariant_tol(A.shape)
    assert res <= tol and orth <= tol, (res, orth, tol)


@pytest.mark.parametrize("name,make", [("graded512", lambda: g.graded(512, seed=3)),
                                       ("kahan200", lambda: g.kahan(200, theta=1.2, perturb=1e3, seed=1)),
                                       ("graded777x1200", lambda: g.graded(1200, seed=4, m=777))])
def test_dgeqp3_rank_deficient(name, make, q):
    """Past the numerical rank the pivots are rounding noise (two summation orders never agree there): the pivots must be
    LAPACK's while |R_jj| stays above the noise floor, |diag R| must be non-increasing up to dlaqps's downdating slack, and
    the factorisation must be a factorisation."""
    A = make()
    m, n = A.shape
    out = q.dgeqp3(A)
    assert out["info"] == 0
    qr, jp, tau, _, info = sla.lapack.dgeqp3(np.asfortranarray(A))
    d = np.abs(np.diag(qr))[: min(m, n)]
    floor = 100.0 * max(m, n) * parity.EPS * d.max()
    trusted = int(np.argmax(d <= floor)) if np.any(d <= floor) else d.size
    assert trusted >= 1
    gd = np.abs(np.diag(out["A"]))[:trusted]
    if name.startswith("kahan"):
        # the Kahan matrix is THE tie case of column pivoting: all trailing partial norms agree to ~1e-13, the winner depends
        # on the last bits of the norm arithmetic (dnrm2 vs sums of squares); whoever wins, |R_kk| is the same to ~1e-6
        assert sorted(out["jpvt"].tolist()) == list(range(1, n + 1))
        assert np.allclose(gd, d[:trusted], rtol=1e-6, atol=0)
    else:
        assert np.array_equal(out["jpvt"][:trusted], jp[:trusted])
        assert np.allclose(gd, d[:trusted], rtol=1e-8, atol=0)
    res, orth = _invariants(A, out)
    tol = parity.invariant_tol(A.shape)
    assert res <= tol and orth <= tol, (res, orth, tol)


def test_dgeqp3_device_entry_point_and_fixed_columns_rejected(q):
    import torch
    m, n, lda = 900, 500, 902
    A = g.gaussian(m, n, 12)
    buf = torch.full((n, lda), 7.5, dtype=torch.float64, device="cuda")
    buf[:, :m] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
    d_jpvt = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_tau = torch.zeros(min(m, n), dtype=torch.float64, device="cuda")
    assert q.dgeqp3_device(buf, m, n, lda, d_jpvt, d_tau) == 0
    out = buf.cpu().numpy()
    assert np.all(out[:, m:] == 7.5)
    qr, jp, tau, _, info = sla.lapack.dgeqp3(np.asfortranarray(A))
    assert np.array_equal(d_jpvt.cpu().numpy(), jp)
    assert np.allclose(np.abs(np.diag(out[:, :m].T)), np.abs(np.diag(qr)), rtol=1e-10, atol=0)
    # the host entry point takes free columns only
    from qrdm_b200 import _lib
    F = np.asfortranarray(A.copy())
    jpvt = np.zeros(n, dtype=np.int32)
    jpvt[3] = 1
    tau2 = np.zeros(min(m, n))
    assert _lib.lib.qrdm_b200_dgeqp3(m, n, F.ctypes.data, m, jpvt.ctypes.data, tau2.ctypes.data) == -102
