"""Cluster-exchange panel check: factor a matrix whose panel grid is > 32 CTAs and verify residual / orthogonality
on the host.  usage: python tools/cl_check.py m n"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import parity
import qrdm_b200
from qrdm_b200 import generators as g

m, n = int(sys.argv[1]), int(sys.argv[2])
A = g.gaussian(m, n, 1)
out = qrdm_b200.dgeqrdm(A)
r = int(out["ncols"].sum())
R = np.triu(out["A"])
info, QR = qrdm_b200.dormqr(out["A"], out["tau"], R, k=r, trans="N")
res = np.linalg.norm(A[:, out["jpvt"] - 1] - QR) / np.linalg.norm(A)
print(f"{m}x{n}: info {out['info']} rank {r} blocks {out['ncols'][:6].tolist()} residual {res:.2e} (tol {parity.invariant_tol(A.shape):.1e})", flush=True)
