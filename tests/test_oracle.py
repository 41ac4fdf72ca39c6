"""CPU suite, part 1: pin the oracle.

The reference ships no reproducible golden vectors, so the pins are
(1) the committed fixtures tests/golden/*.npz, produced by the unmodified reference compiled from
/root/reference (tests/golden/make_golden.py), (2) the reference .so itself where present, and
(3) the stored notebook accuracy levels as order-of-magnitude sanity.
"""
import os

import numpy as np
import pytest

import parity
from golden.cases import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    c = CASES[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    A = z["A"] if c["store_input"] else c["make"]()
    assert tuple(z["shape"]) == A.shape
    fin = A[np.isfinite(A)]
    assert np.isclose(np.sum(np.abs(fin)), float(z["checksum"]), rtol=1e-12), "generator drifted"
    exp = dict(info=int(z["info"]), jpvt=z["jpvt"], ncols=z["ncols"], tau=z["tau"], diagR=z["diagR"])
    return c, A, exp


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_golden(name, oracle_port):
    c, A, exp = load_case(name)
    got = oracle_port.port_dgeqrdm(A, thres=c["thres"], nb=c["nb"], stop_mode=c["stop_mode"])
    st = parity.check_against(got, exp, A.shape, margins=got["margins"], exact=c["exact"])
    if exp["info"] == 0:
        res, orth = parity.qr_invariants(A, got)
        tol = parity.invariant_tol(A.shape)
        assert res <= tol and orth <= tol, (res, orth, tol)
        assert st["cols"] >= 1


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_matches_golden(name, oracle_ref):
    """The compiled reference reproduces its own fixtures bit-for-decision (same box or not)."""
    c, A, exp = load_case(name)
    got = oracle_ref.ref_dgeqrdm(A, thres=c["thres"], nb=c["nb"], stop_mode=c["stop_mode"])
    parity.check_against(got, exp, A.shape, exact=c["exact"])


def test_trusted_prefix_stops_at_noise():
    """graded128: reference blocks [48,14,1,23,...]; true numerical rank 63 -> prefix = 3 blocks."""
    c, A, exp = load_case("graded128")
    nblk, ncol = parity.trusted_prefix(exp["ncols"], exp["diagR"], A.shape)
    assert (nblk, ncol) == (3, 63)


@pytest.mark.parametrize("kw", [dict(thres=(1.5, 0.15)), dict(thres=(0.5, -0.1)), dict(nb=0),
                                dict(layout=7), dict(lda=3)])
def test_port_argument_errors(kw, oracle_port, capfd):
    """All argument failures return -1 (src/dgeqrdm_work.c:559-581)."""
    from qrdm_b200 import generators as g
    got = oracle_port.port_dgeqrdm(g.gaussian(20, 10, 0), **kw)
    assert got["info"] == -1
    capfd.readouterr()


def test_notebook_accuracy_levels(oracle_port):
    """test.ipynb cells 9/11 print ||Q'Q-I|| ~1.5e-14 and ||A[:,p]-QR|| ~3e-15 at n=128: same order."""
    from qrdm_b200 import generators as g
    A = g.graded(128, seed=11)
    got = oracle_port.port_dgeqrdm(A)
    res, orth = parity.qr_invariants(A, got)
    assert orth < 1e-13 and res * np.linalg.norm(A) < 1e-13
