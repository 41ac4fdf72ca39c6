"""compute-sanitizer targets of the second half of round 2: grouped sub-panel kernel of tall panels (slab-resident and
streaming), V'V pass of the tall-skinny trailing update, GPU QRCP (k_qp3.cu), grouped register panel with a DM early stop
and a guard break (planted dependencies).  usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_r2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg as sla

import qrdm_b200
from qrdm_b200 import generators as g

for name, A in (("gauss 70000x130 (tall grouped, slab-resident, V'V pass)", g.gaussian(70000, 130, 1)),
                ("gauss 600000x24 (tall grouped, streaming)", g.gaussian(600000, 24, 2)),
                ("planted 1500x400 (guard break + DM stop in the grouped panel)", g.planted(1500, 400, seed=3, eps=1e-7)),
                ("graded_tall 45000x100 (rounds in the grouped sub-panel)", g.graded_tall(45000, 100, seed=5))):
    out = qrdm_b200.dgeqrdm(A)
    print(name, "info", out["info"], "rank", int(out["ncols"].sum()), "launches", qrdm_b200.stats()["launches"], flush=True)
A = g.gaussian(700, 500, 4)
out = qrdm_b200.dgeqp3(A)
jp = sla.lapack.dgeqp3(np.asfortranarray(A))[1]
print("dgeqp3 700x500 info", out["info"], "pivots equal to LAPACK:", bool(np.array_equal(out["jpvt"], jp)), flush=True)
