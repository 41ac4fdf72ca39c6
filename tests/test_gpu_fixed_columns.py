"""Fixed columns: jpvt[j] != 0 on entry pins column j to the front, LAPACK dgeqp3 style.  The
reference sets out to do this at src/dgeqrdm_work.c:592-635 (swap the fixed columns up front, LAPACKE_dgeqrf on them,
LAPACKE_dormqr on the rest, DM on the free block) but its continuation mis-indexes its auxiliary arrays for nfxd > 0
(SURVEY.md 2a), so the oracle is assembled from the pieces that DO work: the same swap sequence replayed in NumPy,
LAPACK dgeqrf / dormqr for the fixed block, and the unmodified reference's dgeqrdm on the free block."""
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


def _replay_swaps(jpvt_in):
    """src/dgeqrdm_work.c:592-607, 0-based: (source column of every position, jpvt after the loop, nfxd)."""
    n = len(jpvt_in)
    src, jp, nfxd = list(range(n)), list(jpvt_in), 0
    for c in range(n):
        if jp[c] != 0:
            if c != nfxd:
                src[c], src[nfxd] = src[nfxd], src[c]
                jp[c] = jp[nfxd]
                jp[nfxd] = c + 1
            else:
                jp[c] = c + 1
            nfxd += 1
        else:
            jp[c] = c + 1
    return np.array(src), np.array(jp, dtype=np.int32), nfxd


def _expected(A, jpvt_in, oracle_ref, **kw):
    m, n = A.shape
    src, jp, nfxd = _replay_swaps(jpvt_in)
    Ap = np.asfortranarray(A[:, src])
    nf = min(m, nfxd)
    qr, tau_f, _, info = sla.lapack.dgeqrf(np.asfortranarray(Ap[:, :nf]))
    assert info == 0
    out = dict(nf=nf, jpvt=jp.copy(), diag=np.diag(qr)[:nf].copy(), tau=tau_f.copy(), ncols=np.zeros(n, dtype=np.int32))
    if nf < min(m, n):
        rest = np.asfortranarray(Ap[:, nf:])
        lw = int(sla.lapack.dormqr("L", "T", qr, tau_f, rest, -1)[1][0])
        C = sla.lapack.dormqr("L", "T", qr, tau_f, rest, lw)[0]
        tail = oracle_ref.ref_dgeqrdm(C[nf:, :], **kw)
        assert tail["info"] == 0
        out["jpvt"][nf:] = jp[nf:][tail["jpvt"] - 1]
        out["diag"] = np.concatenate([out["diag"], np.diag(tail["A"])[: min(m - nf, n - nf)]])
        out["tau"] = np.concatenate([tau_f, tail["tau"]])
        out["ncols"] = tail["ncols"]
    return out


def _run(q, A, jpvt_in, **kw):
    from qrdm_b200 import _lib
    m, n = A.shape
    F = np.array(A, order="F", copy=True)
    jpvt = np.array(jpvt_in, dtype=np.int32)
    tau = np.zeros(min(m, n))
    ncols = np.zeros(n, dtype=np.int32)
    ncols[0] = kw.get("stop_mode", 0)
    th = np.zeros(3)
    th[:2] = kw.get("thres", (0.9, 0.15))
    info = _lib.lib.dgeqrdm(102, m, n, F.ctypes.data, m, jpvt.ctypes.data, tau.ctypes.data, ncols.ctypes.data, th.ctypes.data,
                            kw.get("nb", 64))
    return dict(info=info, A=F, jpvt=jpvt, tau=tau, ncols=ncols)


CASES = [
    ("gauss300x200_four_fixed", lambda: g.gaussian(300, 200, 1), [3, 50, 51, 120], {}),
    ("gauss300x200_first_fixed", lambda: g.gaussian(300, 200, 2), [0], {}),
    ("gauss400x300_100_fixed", lambda: g.gaussian(400, 300, 3), list(range(5, 205, 2)), {}),          # two forced blocks: 64 + 36
    ("gauss200x150_all_fixed", lambda: g.gaussian(200, 150, 4), list(range(150)), {}),                # plain unpivoted QR
    ("gauss50x200_wide_80_fixed", lambda: g.gaussian(50, 200, 5), list(range(60, 140)), {}),          # na = min(m, nfxd) = 50
    ("gauss1500x900_nb32", lambda: g.gaussian(1500, 900, 6), [899, 0, 450], dict(nb=32, thres=(0.7, 0.3))),
    ("graded256_fixed_stop1", lambda: g.graded(256, seed=5), [10, 200], dict(stop_mode=1)),
]


@pytest.mark.parametrize("name,make,fixed,kw", CASES, ids=[c[0] for c in CASES])
def test_fixed_columns(name, make, fixed, kw, q, oracle_ref):
    A = make()
    m, n = A.shape
    jin = np.zeros(n, dtype=np.int32)
    jin[fixed] = 1
    got = _run(q, A, jin, **kw)
    exp = _expected(A, jin, oracle_ref, **kw)
    assert got["info"] == 0
    nf = exp["nf"]
    # the fixed columns sit up front in the order the reference's swaps leave them, untouched by pivoting
    assert np.array_equal(got["jpvt"][:nf], exp["jpvt"][:nf])
    assert sorted(got["jpvt"].tolist()) == list(range(1, n + 1))
    nblk, ncol = parity.trusted_prefix(exp["ncols"], exp["diag"][nf:], (m - nf, n - nf)) if nf < min(m, n) else (0, 0)
    assert np.array_equal(got["ncols"][:nblk], exp["ncols"][:nblk])
    assert np.array_equal(got["jpvt"][: nf + ncol], exp["jpvt"][: nf + ncol])
    gd, ed = np.abs(np.diag(got["A"]))[: nf + ncol], np.abs(exp["diag"])[: nf + ncol]
    assert np.max(np.abs(gd - ed) / np.maximum(ed, np.finfo(float).tiny)) <= 1e-10
    assert np.allclose(got["tau"][: nf + ncol], exp["tau"][: nf + ncol], rtol=1e-9, atol=1e-13)
    if name.startswith("gauss"):
        assert nblk == np.count_nonzero(exp["ncols"])       # Gaussian: every block, every pivot
        # A P = Q R and Q'Q = I with r = nf + sum(ncols) reflectors
        r = nf + int(got["ncols"].sum())
        assert r == min(m, n)
        tol = parity.invariant_tol((m, n))
        Qr = sla.lapack.dorgqr(np.asfortranarray(got["A"][:, :r]), got["tau"][:r])[0] if m >= n else None
        if Qr is not None:
            R = np.triu(got["A"][:r, :])
            assert np.linalg.norm(np.eye(r) - Qr.T @ Qr) <= tol
            assert np.linalg.norm(A[:, got["jpvt"] - 1] - Qr @ R) / np.linalg.norm(A) <= tol


def test_no_fixed_columns_is_unchanged(q):
    """jpvt all zero on entry takes the ordinary path (bit-identical to the convenience wrapper)."""
    A = g.gaussian(500, 320, 9)
    a = q.dgeqrdm(A)
    b = _run(q, A, np.zeros(320, dtype=np.int32))
    assert np.array_equal(a["A"], b["A"]) and np.array_equal(a["jpvt"], b["jpvt"]) and np.array_equal(a["ncols"], b["ncols"])
