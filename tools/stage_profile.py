"""Per-stage device time of one factorisation (profile mode 1: event pair + sync per stage).
usage: python tools/stage_profile.py m n [gaussian|graded|kahan] [stop_mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import qrdm_b200
from qrdm_b200 import generators as g

m, n = int(sys.argv[1]), int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "gaussian"
stop = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
if kind == "gaussian":
    A0 = torch.randn((n, m), dtype=torch.float64, device=dev)
elif kind == "kahan":
    A0 = torch.from_numpy(np.ascontiguousarray(g.kahan(n).T)).to(dev)
else:
    import bench
    A0 = bench.make_matrix_torch(torch, m, n, "graded", 0, dev)
A = A0.clone()
jp = torch.zeros(n, dtype=torch.int32, device=dev)
tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
for mode in (0, 0, 1):
    qrdm_b200.set_profile(mode)
    A.copy_(A0)
    torch.cuda.synchronize()
    info, nc = qrdm_b200.dgeqrdm_device(A, m, n, m, jp, tau, stop_mode=stop)
    st = qrdm_b200.stats()
    print(f"{m}x{n} {kind} stop={stop} mode {mode}: info {info} rank {int(nc.sum())} iters {st['iterations']} "
          f"ms_total {st['ms_total']:.2f} launches {st['launches']}",
          {k: round(v, 2) for k, v in st["ms_stage"].items() if v > 0})
