"""Development helper: run the CUDA path against the oracle on a list of cases and print
diagnostics (not a test; see tests/ for the gated versions)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import qrdm_b200  # noqa: E402
from oracle import ref  # noqa: E402
from qrdm_b200 import generators as g  # noqa: E402


def run(name, A, use_port=False, **kw):
    t = time.time()
    got = qrdm_b200.dgeqrdm(A, **kw)
    tg = time.time() - t
    st = qrdm_b200.stats()
    t = time.time()
    exp = ref.port_dgeqrdm(A, **kw) if use_port or not ref.have_ref() else ref.ref_dgeqrdm(A, **kw)
    te = time.time() - t
    msg = "OK"
    try:
        pre = parity.check_against(got, exp, A.shape, margins=exp.get("margins"))
        msg = f"OK prefix {pre['blocks']} blk/{pre['cols']} cols"
    except AssertionError as e:
        msg = "MISMATCH: " + str(e)[:150]
    inv = ""
    if got["info"] == 0 and max(A.shape) <= 2500:
        res, orth = parity.qr_invariants(A, got)
        inv = f" res {res:.1e} orth {orth:.1e} (tol {parity.invariant_tol(A.shape):.1e})"
    nb = int(np.count_nonzero(got["ncols"]))
    print(f"{name:26s} info {got['info']:4d}/{exp['info']:4d} rank {int(got['ncols'].sum()):5d}/{int(exp['ncols'].sum()):5d} "
          f"it {nb:4d} gpu {tg*1e3:8.1f} ms (dev {st['ms_total']:8.2f}) cpu {te*1e3:8.1f} ms  {msg}{inv}", flush=True)
    if os.environ.get("QRDM_B200_PROFILE"):
        ms = st["ms_stage"]; ln = st["stage_launches"]
        print("      stages ms: " + " ".join(f"{k}={v:.2f}" for k, v in ms.items() if v > 0) + f" | launches {st['launches']}", flush=True)
    return got, exp


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    run("gauss8", g.gaussian(8, 8, 0))
    run("gauss200", g.gaussian(200, 200, 0))
    run("gauss500x150", g.gaussian(500, 150, 1))
    run("gauss120x300", g.gaussian(120, 300, 2))
    run("gauss200 nb16", g.gaussian(200, 200, 0), thres=(0.6, 0.5), nb=16)
    run("gauss257x131 nb32", g.gaussian(257, 131, 4), nb=32)
    run("kahan96", g.kahan(96))
    run("zeros16", np.zeros((16, 16)))
    run("eye16", np.eye(16))
    run("graded128", g.graded(128, seed=0), use_port=True)
    run("graded128 stop1", g.graded(128, seed=0), use_port=True, stop_mode=1)
    run("gauss1000", g.gaussian(1000, 1000, 0))
    A = g.gaussian(200, 200, 0); A[100, 150] = np.nan
    run("nan", A)
    if which != "small":
        run("gauss2048", g.gaussian(2048, 2048, 1))
        run("gauss4096", g.gaussian(4096, 4096, 2))
        run("graded1024 stop1", g.graded(1024, seed=3), stop_mode=1)
        run("gauss20000x512", g.gaussian(20000, 512, 3))
        run("kahan512", g.kahan(512))
