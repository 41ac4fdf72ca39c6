"""GPU suite, SURVEY.md 8(f)-1: Q application on the device (qrdm_b200_dormqr / _dev) and checkQR at full size.

Oracle for this row = LAPACK dormqr, which is what the reference wrapper's DORMQR calls
(reference QRDM_wrapper.c:104-126, LAPACKE_dormqr('L','N')) and what auxil.checkQR (reference
auxil.py:20-105) builds Q and Q R with.  The factorisations fed to it come from LAPACK dgeqrf (so the
check does not depend on our own dgeqrdm) and from dgeqrdm itself.
"""
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


def _lapack_ormqr(F, tau, Cm, k, trans):
    a = np.asfortranarray(F[:, :k])
    c = np.asfortranarray(Cm)
    lw = int(sla.lapack.dormqr("L", trans, a, tau[:k], c, -1)[1][0])
    cq, _, info = sla.lapack.dormqr("L", trans, a, tau[:k], c, lw)
    assert info == 0
    return cq


@pytest.mark.parametrize("m,n,k,p", [(300, 200, 200, 50), (1000, 1000, 1000, 1000), (1501, 777, 777, 33),
                                     (640, 640, 70, 129), (257, 131, 3, 7), (20000, 128, 128, 64)],
                         ids=["300x200", "1000sq", "odd1501x777", "k70", "k3", "tall20000x128"])
@pytest.mark.parametrize("trans", ["N", "T"])
def test_dormqr_matches_lapack(m, n, k, p, trans, q):
    rng = np.random.default_rng(m + n + k)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F, tau, _, info = sla.lapack.dgeqrf(A)          # LAPACK's own reflectors
    assert info == 0
    Cm = np.asfortranarray(rng.standard_normal((m, p)))
    info, got = q.dormqr(F, tau, Cm, k=k, trans=trans)
    assert info == 0 and q.stats()["launches"] > 0
    exp = _lapack_ormqr(F, tau, Cm, k, trans)
    scale = np.linalg.norm(Cm)
    assert np.linalg.norm(got - exp) <= 50 * max(m, p) * parity.EPS * scale, np.linalg.norm(got - exp) / scale


def test_dormqr_on_dgeqrdm_output_gives_checkqr_metrics(q):
    """auxil.checkQR on the device path: ||A P - Q R|| / ||A|| and ||I - Q'Q|| from the GPU Q application
    agree with the host LAPACK evaluation used everywhere else in this suite."""
    A = g.gaussian(700, 500, seed=3)
    out = q.dgeqrdm(A)
    r = int(out["ncols"].sum())
    res_h, orth_h = parity.qr_invariants(A, out)
    R = np.triu(out["A"])[:, :]
    R[r:, :r] = 0.0
    info, QR = q.dormqr(out["A"], out["tau"], R, k=r, trans="N")
    assert info == 0
    res = np.linalg.norm(A[:, out["jpvt"] - 1] - QR) / np.linalg.norm(A)
    info, Q = q.dormqr(out["A"], out["tau"], np.eye(700), k=r, trans="N")
    orth = np.linalg.norm(np.eye(700) - Q.T @ Q)
    tol = parity.invariant_tol(A.shape)
    assert res <= tol and orth <= tol and abs(res - res_h) <= tol and abs(orth - orth_h) <= tol, (res, res_h, orth, orth_h)
    # Q' (Q C) = C
    Cm = np.random.default_rng(0).standard_normal((700, 40))
    _, QC = q.dormqr(out["A"], out["tau"], Cm, k=r, trans="N")
    _, back = q.dormqr(out["A"], out["tau"], QC, k=r, trans="T")
    assert np.linalg.norm(back - Cm) <= tol * np.linalg.norm(Cm)


@pytest.mark.parametrize("m,n", [(16384, 16384), (1000000, 512)], ids=["C3_16384sq", "C4_1000000x512"])
def test_checkqr_at_full_size_on_device(m, n, q):
    """BASELINE.json sizes: factor on the device, then form Q R and the thin Q with qrdm_b200_dormqr_dev and
    grade ||A P - Q R|| / ||A|| and ||I - Q'Q|| <= 10 max(m,n) eps — no host copy of these matrices, no m x m Q."""
    import torch
    import bench
    dev = torch.device("cuda", 0)
    A0 = bench.make_matrix_torch(torch, m, n, "gaussian", seed=0, device=dev)   # (n, m) row-major = m x n column-major
    F = A0.clone()
    d_jpvt = torch.zeros(n, dtype=torch.int32, device=dev)
    d_tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
    info, ncols = q.dgeqrdm_device(F, m, n, m, d_jpvt, d_tau)
    assert info == 0
    r = int(ncols.sum())
    assert r == min(m, n)
    nrmA = float(torch.linalg.norm(A0))
    # R = upper triangle of the factored matrix (stored transposed: F[c, r] = A[r, c])
    R = torch.tril(F)                       # keeps entries with row index <= column index of the m x n matrix
    assert q.dormqr_device("N", m, n, r, F, m, d_tau, R, m) == 0
    st = q.stats()
    P = d_jpvt.long() - 1
    res = float(torch.linalg.norm(A0[P, :] - R)) / nrmA
    del R
    Qt = torch.zeros((r, m), dtype=torch.float64, device=dev)                   # thin Q, column-major m x r
    Qt.diagonal().fill_(1.0)
    assert q.dormqr_device("N", m, r, r, F, m, d_tau, Qt, m) == 0
    G = Qt @ Qt.T
    G.diagonal().sub_(1.0)
    orth = float(torch.linalg.norm(G))
    tol = parity.invariant_tol((m, n))
    print(f"checkQR {m}x{n}: residual {res:.2e} orthogonality {orth:.2e} (tol {tol:.2e}); Q R formed in "
          f"{st['ms_total']:.1f} ms = {st['trailing_flops'] / st['ms_total'] / 1e9:.1f} TFLOP/s")
    assert res <= tol and orth <= tol, (res, orth, tol)


def test_low_rank_matches_host_evaluation(q):
    """auxil.low_rank (reference auxil.py:156-200) with the GPU Q application: A_k = Q [R11 R12; 0 0]; its error
    against A P is ||R22||, and it agrees with the LAPACK evaluation of the same product."""
    A = g.graded(256, seed=1)                       # numerical rank 127
    out = q.dgeqrdm(A)
    k = 100
    Ak = q.low_rank(out["A"], out["tau"], k)
    R = np.zeros_like(A, order="F")
    R[:k, :] = np.triu(out["A"])[:k, :]
    exp = _lapack_ormqr(out["A"], out["tau"], R, k, "N")
    assert np.linalg.norm(Ak - exp) <= 50 * 256 * parity.EPS * np.linalg.norm(A)
    AP = A[:, out["jpvt"] - 1]
    R22 = np.triu(out["A"])[k:, k:]
    assert abs(np.linalg.norm(AP - Ak) - np.linalg.norm(R22)) <= 1e-10 * np.linalg.norm(A)


def test_nb1_is_column_pivoted_qr(q):
    """With one candidate per iteration (nb = 1) Deviation Maximisation degenerates to classical column pivoting:
    the pivots and |diag R| of LAPACK dgeqp3 (the reference wrapper's QP3, QRDM_wrapper.c:15-41)."""
    A = g.gaussian(300, 200, seed=8)
    out = q.dgeqrdm(A, nb=1)
    assert out["info"] == 0 and int(out["ncols"].sum()) == 200 and out["ncols"][:200].max() == 1
    qr, jp, tau, _, info = sla.lapack.dgeqp3(np.asfortranarray(A))
    assert info == 0
    assert np.array_equal(out["jpvt"], jp)
    assert np.allclose(np.abs(np.diag(out["A"])), np.abs(np.diag(qr)), rtol=1e-10, atol=0)
