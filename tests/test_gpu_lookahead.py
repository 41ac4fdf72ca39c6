"""GPU suite, look-ahead of the deferred trailing update.

Pass 2 of the pending block is applied to the last columns of the trailing matrix by a second, least-priority stream
(k_rankk in side mode, short CTAs) while the main stream runs the next selection chain and the next panel; the next
k_fused then makes pass 1 only on those columns (qrdm_prob::pre_col0) and its CTAs take unit ranges of equal COST.
The share is sized from the idle SM-time of that window (QRDM_B200_SIDE_US, QRDM_B200_SIDE_COL_US), so on the small matrices of a test suite the default hands
the whole matrix beyond the eager set to the side stream; tiny windows leave most of it to k_fused and make the
boundary fall inside a column tile.  Whatever the split, pivots and block sizes must be those of the run without
look-ahead (and of the reference), and the factors equal to rounding — every column gets the same update exactly once.
"""
import numpy as np
import pytest

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


SETTINGS = [
    ("window0.5us", {"QRDM_B200_SIDE_US": "0.5", "QRDM_B200_SIDE_COL_US": "0"}),
    ("window3us", {"QRDM_B200_SIDE_US": "3", "QRDM_B200_SIDE_COL_US": "0"}),
    ("window12us", {"QRDM_B200_SIDE_US": "12", "QRDM_B200_SIDE_COL_US": "0.1"}),
    ("default", {}),
    ("ends_before_panel", {"QRDM_B200_SIDE_PANEL": "0", "QRDM_B200_SIDE_US": "6"}),
    ("one_unit_per_cta", {"QRDM_B200_SIDE_UPC": "1", "QRDM_B200_SIDE_US": "9"}),
    ("seven_units_per_cta", {"QRDM_B200_SIDE_UPC": "7"}),
]

CASES = [
    ("gauss2500x2300", lambda: g.gaussian(2500, 2300, 11), {}),
    ("gauss1501x777_odd_m", lambda: g.gaussian(1501, 777, 4), {}),                 # non-VEC16 kernels
    ("gauss700x1900_wide", lambda: g.gaussian(700, 1900, 3), {}),
    ("gauss1400_nb24_d05", lambda: g.gaussian(1400, 1400, 6), dict(thres=(0.5, 0.6), nb=24)),
    ("kahan300_perturbed", lambda: g.kahan(300, theta=1.2, perturb=1e3, seed=1), {}),  # one-column blocks, flagged norms
    ("graded1024_stop1", lambda: g.graded(1024, seed=3), dict(stop_mode=1)),       # flush with a side update in flight
    ("graded777x1200", lambda: g.graded(1200, seed=4, m=777), {}),
]


def _run(q, monkeypatch, A, kw, env):
    for k in ("QRDM_B200_SIDE", "QRDM_B200_SIDE_US", "QRDM_B200_SIDE_PANEL", "QRDM_B200_SIDE_COL_US", "QRDM_B200_SIDE_UPC", "QRDM_B200_SIDE_EFF"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return q.dgeqrdm(A, **kw)


@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_lookahead_equals_plain_deferred_schedule(name, make, kw, q, monkeypatch):
    monkeypatch.setenv("QRDM_B200_LAZY", "1")
    monkeypatch.setenv("QRDM_B200_LAZY_MIN", "1")
    A = make()
    base = _run(q, monkeypatch, A, kw, {"QRDM_B200_SIDE": "0"})
    assert base["info"] == 0
    graded = name.startswith("graded")  # past the numerical rank the pivots are rounding noise (parity.py): prefix rule
    for label, env in SETTINGS:
        got = _run(q, monkeypatch, A, kw, env)
        assert got["info"] == 0, label
        st = parity.check_against(got, base, A.shape, exact=not graded)
        assert st["cols"] >= 1, label
        r = st["cols"]
        assert np.allclose(got["tau"][:r], base["tau"][:r], rtol=0, atol=1e-10), label
        scale = np.abs(np.diag(base["A"])).max()
        assert np.abs(np.triu(got["A"][:, :r]) - np.triu(base["A"][:, :r])).max() <= 1e-10 * scale, label
        res, orth = parity.qr_invariants(A, got)
        tol = parity.invariant_tol(A.shape)
        assert res <= tol and orth <= tol, (label, res, orth, tol)
        assert q.stats()["launches"] > 0


def test_lookahead_against_reference_partial_window(q, oracle_ref, monkeypatch):
    """The boundary between the side stream's columns and k_fused's falls inside a column tile and moves every iteration."""
    monkeypatch.setenv("QRDM_B200_LAZY", "1")
    monkeypatch.setenv("QRDM_B200_LAZY_MIN", "1")
    monkeypatch.setenv("QRDM_B200_SIDE_US", "5")
    monkeypatch.setenv("QRDM_B200_SIDE_COL_US", "0.05")
    A = g.gaussian(3000, 2600, 7)
    got = q.dgeqrdm(A)
    exp = oracle_ref.ref_dgeqrdm(A)
    st = parity.graded_check("lookahead/gauss3000x2600_window5us", got, exp, A.shape, family="gaussian", require_full=True)
    assert st["cols_trusted"] >= 1
    res, orth = parity.qr_invariants(A, got)
    tol = parity.invariant_tol(A.shape)
    assert res <= tol and orth <= tol, (res, orth, tol)


def test_lookahead_bitwise_determinism(q, monkeypatch):
    monkeypatch.setenv("QRDM_B200_LAZY", "1")
    monkeypatch.setenv("QRDM_B200_LAZY_MIN", "1")
    monkeypatch.setenv("QRDM_B200_SIDE_US", "4")
    monkeypatch.setenv("QRDM_B200_SIDE_COL_US", "0.05")
    A = g.gaussian(1800, 1700, 21)
    a = q.dgeqrdm(A)
    b = q.dgeqrdm(A)
    assert np.array_equal(a["A"], b["A"]) and np.array_equal(a["jpvt"], b["jpvt"]) and np.array_equal(a["tau"], b["tau"])
