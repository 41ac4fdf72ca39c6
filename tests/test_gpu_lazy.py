"""GPU suite, deferred trailing update (k_fused / k_wapply rows mode / k_rankk<list> / k_colupd / flush).

By default the deferred path only engages while the trailing matrix has >= 7168^2 elements, so the
small fixtures of test_gpu_parity.py never reach it.  Here QRDM_B200_LAZY_MIN=1 forces it from the
first iteration on (every block deferred until fewer than 65 columns remain), which covers:
fused pass-2/pass-1, eager completion of the leading 64 positions + candidates, flagged-norm columns
brought up to date before their exact recompute (graded / Kahan), the flush on an early stop, and
leftover panel columns after a DM early stop (k < fjb).  Checked against the same oracles as the
eager path: golden fixtures (unmodified reference outputs), the reference build, and invariants.
"""
import os

import numpy as np
import pytest

import parity
from golden.cases import CASES
from qrdm_b200 import generators as g

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def q():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU suite selected but no CUDA device: the product has no CPU fallback")
    import qrdm_b200
    return qrdm_b200


@pytest.fixture()
def forced_lazy(monkeypatch):
    monkeypatch.setenv("QRDM_B200_LAZY_MIN", "1")
    monkeypatch.setenv("QRDM_B200_LAZY", "1")


def _golden(name):
    c = CASES[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    A = z["A"] if c["store_input"] else c["make"]()
    exp = dict(info=int(z["info"]), jpvt=z["jpvt"], ncols=z["ncols"], tau=z["tau"], diagR=z["diagR"])
    return c, A, exp


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_fixtures_deferred(name, q, oracle_port, forced_lazy, capfd):
    c, A, exp = _golden(name)
    got = q.dgeqrdm(A, thres=c["thres"], nb=c["nb"], stop_mode=c["stop_mode"])
    margins = None
    if exp["info"] == 0 and not c["exact"]:
        margins = oracle_port.port_dgeqrdm(A, thres=c["thres"], nb=c["nb"], stop_mode=c["stop_mode"])["margins"]
    if name == "inf_in_panel":
        assert got["info"] in (-8, -13)
        return
    parity.check_against(got, exp, A.shape, margins=margins, exact=c["exact"])
    if exp["info"] == 0:
        res, orth = parity.qr_invariants(A, got)
        tol = parity.invariant_tol(A.shape)
        assert res <= tol and orth <= tol, (res, orth, tol)
    capfd.readouterr()


LAZY_CASES = [
    ("gauss1000", lambda: g.gaussian(1000, 1000, 0), {}, "1"),
    ("gauss700x1900_wide", lambda: g.gaussian(700, 1900, 3), {}, "1"),
    ("gauss1501x777_odd", lambda: g.gaussian(1501, 777, 4), {}, "1"),            # odd m: non-VEC16 path
    ("gauss1024_nb24_d05", lambda: g.gaussian(1024, 1024, 6), dict(thres=(0.5, 0.6), nb=24), "1"),
    ("kahan300_perturbed", lambda: g.kahan(300, theta=1.2, perturb=1e3, seed=1), {}, "1"),
    ("graded1024_stop1", lambda: g.graded(1024, seed=3), dict(stop_mode=1), "1"),  # flush on the early stop
    ("graded777x1200", lambda: g.graded(1200, seed=4, m=777), {}, "1"),
    ("gauss3000x2600_min1500", lambda: g.gaussian(3000, 2600, 7), {}, "1500"),  # deferred first, then eager
]


@pytest.mark.parametrize("name,make,kw,lazy_min", LAZY_CASES, ids=[c[0] for c in LAZY_CASES])
def test_deferred_against_reference(name, make, kw, lazy_min, q, oracle_ref, oracle_port, monkeypatch):
    if lazy_min is None:
        monkeypatch.delenv("QRDM_B200_LAZY_MIN", raising=False)
    else:
        monkeypatch.setenv("QRDM_B200_LAZY_MIN", lazy_min)
    monkeypatch.setenv("QRDM_B200_LAZY", "1")
    A = make()
    got = q.dgeqrdm(A, **kw)
    exp = oracle_ref.ref_dgeqrdm(A, **kw)
    fam = "graded" if name.startswith("graded") else ("kahan" if name.startswith("kahan") else "gaussian")
    st = parity.graded_check("deferred/" + name, got, exp, A.shape, family=fam, require_full=(fam != "graded"),
                             margins_fn=lambda: oracle_port.port_dgeqrdm(A, **kw)["margins"])
    assert st["cols_trusted"] >= 1
    if max(A.shape) <= 3000:
        res, orth = parity.qr_invariants(A, got)
        tol = parity.invariant_tol(A.shape)
        assert res <= tol and orth <= tol, (res, orth, tol)


def test_deferred_matches_eager_path(q, monkeypatch):
    """Same input through both schedules: identical pivots and block sizes, R equal to rounding."""
    A = g.gaussian(2500, 2300, 11)
    monkeypatch.setenv("QRDM_B200_LAZY", "0")
    a = q.dgeqrdm(A)
    monkeypatch.setenv("QRDM_B200_LAZY", "1")
    monkeypatch.setenv("QRDM_B200_LAZY_MIN", "1")
    b = q.dgeqrdm(A)
    assert a["info"] == b["info"] == 0
    assert np.array_equal(a["jpvt"], b["jpvt"]) and np.array_equal(a["ncols"], b["ncols"])
    da, db = np.abs(np.diag(a["A"])), np.abs(np.diag(b["A"]))
    assert np.allclose(da, db, rtol=1e-10, atol=0)
    assert np.allclose(np.triu(a["A"]), np.triu(b["A"]), rtol=0, atol=1e-9 * np.abs(da).max())


def test_deferred_bitwise_determinism(q, forced_lazy):
    A = g.gaussian(900, 800, 5)
    a = q.dgeqrdm(A)
    b = q.dgeqrdm(A)
    assert np.array_equal(a["A"], b["A"]) and np.array_equal(a["jpvt"], b["jpvt"]) and np.array_equal(a["tau"], b["tau"])


def test_deferred_device_entry_point_odd_lda(q, oracle_ref, forced_lazy):
    """Odd leading dimension on the device with the deferred schedule forced on: k_fused<false> and
    k_rankk<false, LIST> (8-byte copies, scalar C loads/stores); the padding rows must stay untouched."""
    import torch
    for (m, n, lda) in [(300, 210, 301), (1100, 700, 1103), (640, 900, 641)]:
        A = g.gaussian(m, n, 40 + m)
        buf = torch.full((n, lda), 3.25, dtype=torch.float64, device="cuda")
        buf[:, :m] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
        d_jpvt = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_tau = torch.zeros(min(m, n), dtype=torch.float64, device="cuda")
        info, ncols = q.dgeqrdm_device(buf, m, n, lda, d_jpvt, d_tau)
        assert info == 0
        out = buf.cpu().numpy()
        assert np.all(out[:, m:] == 3.25)
        got = dict(info=info, A=out[:, :m].T, jpvt=d_jpvt.cpu().numpy(), tau=d_tau.cpu().numpy(), ncols=ncols)
        parity.check_against(got, oracle_ref.ref_dgeqrdm(A), (m, n))
        res, orth = parity.qr_invariants(A, got)
        tol = parity.invariant_tol(A.shape)
        assert res <= tol and orth <= tol, (res, orth, tol)


def test_deferred_dormqr_device_odd_ldc(q):
    """Q application with an odd leading dimension of C on the device (non-vectorised K6 paths)."""
    import scipy.linalg as sla
    import torch
    m, n, p, ldc = 500, 300, 77, 501
    rng = np.random.default_rng(5)
    F, tau, _, info = sla.lapack.dgeqrf(np.asfortranarray(rng.standard_normal((m, n))))
    Cm = rng.standard_normal((m, p))
    dF = torch.from_numpy(np.ascontiguousarray(F.T)).cuda()
    dtau = torch.from_numpy(tau).cuda()
    dC = torch.full((p, ldc), -1.5, dtype=torch.float64, device="cuda")
    dC[:, :m] = torch.from_numpy(np.ascontiguousarray(Cm.T)).cuda()
    assert q.dormqr_device("N", m, p, n, dF, m, dtau, dC, ldc) == 0
    out = dC.cpu().numpy()
    assert np.all(out[:, m:] == -1.5)
    lw = int(sla.lapack.dormqr("L", "N", F, tau, np.asfortranarray(Cm), -1)[1][0])
    exp = sla.lapack.dormqr("L", "N", F, tau, np.asfortranarray(Cm), lw)[0]
    assert np.linalg.norm(out[:, :m].T - exp) <= 50 * m * parity.EPS * np.linalg.norm(Cm)
