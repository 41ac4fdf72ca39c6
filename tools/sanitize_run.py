"""Small runs of every new code path for compute-sanitizer (memcheck / racecheck):
deferred trailing update forced on (QRDM_B200_LAZY_MIN=1), early-stop flush, odd sizes, Q application."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QRDM_B200_LAZY_MIN"] = "1"
import numpy as np

import qrdm_b200
from qrdm_b200 import generators as g

for name, A, kw in (("gauss 700x900", g.gaussian(700, 900, 1), {}),
                    ("gauss 333x257 (odd m: 8-byte paths)", g.gaussian(333, 257, 2), {}),
                    ("graded 512 stop rule 1 (flush, flagged norms)", g.graded(512, seed=3), dict(stop_mode=1)),
                    ("kahan 200 perturbed (k = 1 blocks)", g.kahan(200, theta=1.2, perturb=1e3, seed=1), {})):
    out = qrdm_b200.dgeqrdm(A, **kw)
    r = int(out["ncols"].sum())
    print(name, "info", out["info"], "rank", r, "launches", qrdm_b200.stats()["launches"], flush=True)
    info, QR = qrdm_b200.dormqr(out["A"], out["tau"], np.triu(out["A"]), k=min(r, min(A.shape)), trans="N")
    if kw.get("stop_mode", 0) == 0:
        res = np.linalg.norm(A[:, out["jpvt"] - 1] - QR) / np.linalg.norm(A)
        print("   dormqr info", info, "residual", f"{res:.2e}", flush=True)
