"""GPU QRCP (qrdm_b200_dgeqp3_dev) beside dgeqrdm on the same device-resident matrix.  usage: python tools/qp3_bench.py n [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qrdm_b200

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
A0 = torch.randn((n, n), dtype=torch.float64, device=dev, generator=gen)
A = A0.clone()
jp = torch.zeros(n, dtype=torch.int32, device=dev)
tau = torch.zeros(n, dtype=torch.float64, device=dev)
for r in range(reps):
    A.copy_(A0)
    torch.cuda.synchronize()
    info = qrdm_b200.dgeqp3_device(A, n, n, n, jp, tau)
    st = qrdm_b200.stats()
    print(f"dgeqp3 (GPU, blocked QRCP) {n}x{n}: info {info} ms_total {st['ms_total']:.1f} launches {st['launches']} "
          f"= {4 / 3 * n ** 3 / (st['ms_total'] * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)
for r in range(reps):
    A.copy_(A0)
    torch.cuda.synchronize()
    info, nc = qrdm_b200.dgeqrdm_device(A, n, n, n, jp, tau)
    st = qrdm_b200.stats()
    print(f"dgeqrdm (GPU, DM pivoting)  {n}x{n}: info {info} rank {int(nc.sum())} ms_total {st['ms_total']:.1f} "
          f"= {4 / 3 * n ** 3 / (st['ms_total'] * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)
