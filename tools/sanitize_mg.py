"""compute-sanitizer target: the row-sharded code path (1-rank peer context, QRDM_B200_FORCE_MG) on small inputs.
    compute-sanitizer --tool memcheck python tools/sanitize_mg.py"""
import ctypes as C
import os
import sys

os.environ["QRDM_B200_FORCE_MG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from qrdm_b200 import _lib, generators as g, sharded  # noqa: E402

_lib.lib.qrdm_b200_init(0)
h = C.create_string_buffer(64)
assert _lib.lib.qrdm_b200_peer_handle(h) == 0 and _lib.lib.qrdm_b200_peer_open(0, 1, h.raw) == 0
for name, A in [("gauss600x200", g.gaussian(600, 200, 0)), ("kahan200", g.kahan(200)), ("gauss100x80", g.gaussian(100, 80, 5))]:
    m, n = A.shape
    lda = m + (m & 1)
    loc = torch.zeros((n, lda), dtype=torch.float64, device="cuda")
    loc[:, :m] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
    jp = torch.zeros(n, dtype=torch.int32, device="cuda"); tau = torch.zeros(min(m, n), dtype=torch.float64, device="cuda")
    info, nc = sharded.dgeqrdm_sharded(loc, m, m, 0, 1, n, lda, jp, tau)
    torch.cuda.synchronize()
    print(name, "info", info, "rank", int(nc.sum()), "iterations", int(np.count_nonzero(nc)), flush=True)
