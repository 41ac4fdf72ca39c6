"""One device-resident factorisation (for ncu): python tools/one_run.py m n [lazy]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qrdm_b200

m, n = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3:
    os.environ["QRDM_B200_LAZY"] = sys.argv[3]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
A = torch.randn((n, m), dtype=torch.float64, device=dev, generator=gen)
jp = torch.zeros(n, dtype=torch.int32, device=dev)
tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
info, nc = qrdm_b200.dgeqrdm_device(A, m, n, m, jp, tau)
st = qrdm_b200.stats()
print(f"{m}x{n}: info {info} rank {int(nc.sum())} ms_total {st['ms_total']:.2f} launches {st['launches']}")
