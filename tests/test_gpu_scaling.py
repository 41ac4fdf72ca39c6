"""Badly scaled inputs (VERDICT r1 missing #5): entries around 1e+-200, where plain sums of squares
leave the double range.  The reference factors them (scaled cblas_dnrm2 at src/dgeqrdm_work.c:69,96,673; dlarfg's
safmin loop, src/dlarfg.c:144-182); round 1 silently returned wrong pivots with info = 0.  The driver now multiplies
such a matrix by one power of two, factors it with the unchanged kernels and divides the R-like entries back
(dgeqrdm_host.c: prescale_input; k_small.cu: small_prescale).  Power-of-two scaling is exact, so for HUGE inputs the
results must be bit-identical to those of the well-scaled matrix up to that factor.  For TINY inputs the reference is
itself not scale invariant: its norm downdate sums unscaled squares (src/dgeqrdm_work.c:81-86), which underflow to 0 at
~1e-200, so its partial norms go stale and it picks other pivots than for the same matrix at unit scale (and for huge
inputs the sum overflows and forces an exact recompute every iteration).  The downdate is therefore evaluated in the
caller's scale (qrdm_prob::inv_scale) — and every case here must equal the REFERENCE on the same input."""
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


def _split(F, r):
    """(R-like part, V part) of a factored matrix of rank r."""
    m, n = F.shape
    R = np.triu(F).copy()
    R[:, r:] = F[:, r:]
    V = np.tril(F, -1).copy()
    V[:, r:] = 0.0
    return R, V


@pytest.mark.parametrize("expo", [600, 700])
@pytest.mark.parametrize("shape,kw", [((300, 200), {}), ((257, 300), {}), ((1200, 700), dict(nb=32, thres=(0.7, 0.3)))],
                         ids=["300x200", "257x300", "1200x700_nb32"])
def test_power_of_two_scaling_is_exact(expo, shape, kw, q):
    A = g.gaussian(*shape, seed=7)
    base = q.dgeqrdm(A, **kw)
    sc = q.dgeqrdm(np.ldexp(A, expo), **kw)
    assert base["info"] == 0 and sc["info"] == 0
    assert np.array_equal(sc["jpvt"], base["jpvt"]) and np.array_equal(sc["ncols"], base["ncols"])
    assert np.array_equal(sc["tau"], base["tau"])
    r = int(base["ncols"].sum())
    Rb, Vb = _split(base["A"], r)
    Rs, Vs = _split(sc["A"], r)
    assert np.array_equal(Vs, Vb)
    assert np.array_equal(Rs, np.ldexp(Rb, expo))


@pytest.mark.parametrize("factor", [1e200, 1e-200, 1e250, 3e-290, 2.0 ** -600, 2.0 ** 650, 1e-140, 1e120])
def test_scaled_input_against_reference(factor, q, oracle_ref):
    A = g.gaussian(400, 260, seed=11) * factor
    got = q.dgeqrdm(A)
    exp = oracle_ref.ref_dgeqrdm(A)
    assert exp["info"] == 0
    e = parity.graded_check(f"gauss400x260 * {factor:g}", got, exp, A.shape, family="gaussian", require_full=True)
    assert e["mode"] == "exact"
    r = int(exp["ncols"].sum())
    assert np.allclose(got["tau"][:r], exp["tau"][:r], rtol=1e-9, atol=1e-13)


def test_scaled_graded_stop_rule(q, oracle_ref, oracle_port):
    """Graded rank-deficient input at 1e-210 with the stop rule: the trusted prefix (through the numerical rank) must
    equal the reference's.  (Past it the reference's partial norms are stale at this scale — its downdate underflows —
    so its stop rule fires late or never; that tail is noise and is not compared.)"""
    A = g.graded(256, seed=2) * 1e-210
    got = q.dgeqrdm(A, stop_mode=1)
    exp = oracle_ref.ref_dgeqrdm(A, stop_mode=1)
    parity.graded_check("graded256 * 1e-210 stop1", got, exp, A.shape, family="graded",
                        margins_fn=lambda: oracle_port.port_dgeqrdm(A, stop_mode=1)["margins"])
    assert int(got["ncols"].sum()) >= 127 and int(exp["ncols"].sum()) >= 127


def test_batched_kernel_scaling(q, oracle_ref):
    """One-CTA-per-matrix kernel: two of the five matrices are badly scaled (one huge: bit-identical to the well-scaled
    result up to the factor; one tiny: equal to the reference, whose downdate underflows there)."""
    expo = 620
    As = np.stack([g.gaussian(96, 96, seed=s) for s in range(5)])
    base = q.dgeqrdm_batched(As)
    Sc = As.copy()
    Sc[1] = np.ldexp(Sc[1], expo)
    Sc[3] = np.ldexp(Sc[3], -expo)
    sc = q.dgeqrdm_batched(Sc)
    assert base["info"] == 0 and sc["info"] == 0 and not sc["infos"].any()
    for b in (0, 1, 2, 4):
        e = expo if b == 1 else 0
        assert np.array_equal(sc["jpvt"][b], base["jpvt"][b]) and np.array_equal(sc["ncols"][b], base["ncols"][b])
        assert np.array_equal(sc["tau"][b], base["tau"][b])
        r = int(base["ncols"][b].sum())
        Rb, Vb = _split(base["A"][b], r)
        Rs, Vs = _split(sc["A"][b], r)
        assert np.array_equal(Vs, Vb) and np.array_equal(Rs, np.ldexp(Rb, e))
    for b in (1, 3):
        exp = oracle_ref.ref_dgeqrdm(Sc[b])
        got = dict(info=0, A=sc["A"][b], jpvt=sc["jpvt"][b], tau=sc["tau"][b], ncols=sc["ncols"][b])
        parity.check_against(got, exp, (96, 96), exact=True)


def test_zero_matrix_and_inf_are_not_rescaled(q):
    """maxnrm = 0 and maxnrm = Inf trigger the max |a_ij| pass but must leave the established behaviour alone."""
    Z = np.zeros((40, 30), order="F")
    out = q.dgeqrdm(Z)
    assert out["info"] == 0 and np.array_equal(out["jpvt"], np.arange(1, 31)) and not out["A"].any()
    B = g.gaussian(40, 30, 0)
    B[3, 4] = np.inf
    assert q.dgeqrdm(B)["info"] in (-8, -13)
