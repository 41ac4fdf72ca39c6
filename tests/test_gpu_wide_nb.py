"""Block sizes above 64 (VERDICT r1 missing #4): the reference accepts any nb > 0 (src/dgeqrdm_work.c:573-575); round 1
rejected nb > 64.  Up to nb = 256 the selection now runs at full width (k_wide.cu: wide Gram, greedy pick on the full
cosine matrix, permute_marked replayed literally) and the selected block is factored in micro-panels of 64 columns.
Every case against the unmodified reference with the same nb."""
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


CASES = [
    ("gauss600x500_nb65", lambda: g.gaussian(600, 500, 1), dict(nb=65)),
    ("gauss600x500_nb100", lambda: g.gaussian(600, 500, 1), dict(nb=100)),
    ("gauss900x700_nb128", lambda: g.gaussian(900, 700, 2), dict(nb=128)),
    ("gauss2000x1500_nb200", lambda: g.gaussian(2000, 1500, 3), dict(nb=200)),
    ("gauss1100x1100_nb256", lambda: g.gaussian(1100, 1100, 4), dict(nb=256)),
    ("gauss300x900_wide_nb150", lambda: g.gaussian(300, 900, 5), dict(nb=150)),
    ("gauss800x600_nb128_d05_t06", lambda: g.gaussian(800, 600, 6), dict(nb=128, thres=(0.5, 0.6))),   # many rejections
    ("gauss45000x300_tall_nb128", lambda: g.gaussian(45000, 300, 7), dict(nb=128)),                   # blocked tall panel
    ("kahan300_perturbed_nb128", lambda: g.kahan(300, theta=1.2, perturb=1e3, seed=1), dict(nb=128)),
    ("graded512_nb128_stop1", lambda: g.graded(512, seed=3), dict(nb=128, stop_mode=1)),              # early stops
    ("graded640x512_nb256", lambda: g.graded(512, seed=4, m=640), dict(nb=256)),
]


@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_wide_blocks_against_reference(name, make, kw, q, oracle_ref, oracle_port):
    A = make()
    got = q.dgeqrdm(A, **kw)
    exp = oracle_ref.ref_dgeqrdm(A, **kw)
    fam = "graded" if name.startswith("graded") else ("kahan" if name.startswith("kahan") else "gaussian")
    e = parity.graded_check("nb>64/" + name, got, exp, A.shape, family=fam, require_full=(fam != "graded"),
                            margins_fn=lambda: oracle_port.port_dgeqrdm(A, **kw)["margins"])
    assert e["cols_trusted"] >= 1
    assert int(exp["ncols"].max()) > 64 or fam != "gaussian", "the case must actually produce a block wider than 64"
    r = e["cols_trusted"]
    assert np.allclose(got["tau"][:r], exp["tau"][:r], rtol=1e-9, atol=1e-13)
    if max(A.shape) <= 2100:
        res, orth = parity.qr_invariants(A, got)
        tol = parity.invariant_tol(A.shape)
        assert res <= tol and orth <= tol, (res, orth, tol)


def test_nb64_path_untouched_by_the_wide_mode(q):
    """nb <= 64 never enters k_wide.cu: bit-identical to itself across calls interleaved with wide ones."""
    A = g.gaussian(700, 500, 9)
    a = q.dgeqrdm(A)
    q.dgeqrdm(g.gaussian(400, 300, 1), nb=128)
    b = q.dgeqrdm(A)
    assert np.array_equal(a["A"], b["A"]) and np.array_equal(a["jpvt"], b["jpvt"])
