"""Deferred vs eager trailing update on one device-resident matrix (CUDA events inside dgeqrdm_dev).
usage: python tools/lazy_bench.py m n [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import qrdm_b200

m, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
A0 = torch.randn((n, m), dtype=torch.float64, device=dev, generator=gen)
A = A0.clone()
jp = torch.zeros(n, dtype=torch.int32, device=dev)
tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
qrdm_b200.set_profile(2)
ref = None
for lazy in ("0", "1", "0", "1"):
    os.environ["QRDM_B200_LAZY"] = lazy
    for r in range(reps):
        A.copy_(A0)
        torch.cuda.synchronize()
        info, nc = qrdm_b200.dgeqrdm_device(A, m, n, m, jp, tau)
        st = qrdm_b200.stats()
    tf = st["trailing_flops"] / (st["ms_stage"]["trailing"] * 1e-3) / 1e12
    d = torch.abs(torch.diagonal(A.T)[: min(m, n)]).clone()
    same = ""
    if ref is None:
        ref = (jp.clone(), d)
    else:
        same = f" jpvt_equal={bool(torch.equal(jp, ref[0]))} max_rel_diag_diff={float(torch.max(torch.abs(d - ref[1]) / ref[1])):.2e}"
    print(f"{m}x{n} lazy={lazy}: info {info} rank {int(nc.sum())} iters {st['iterations']} ms_total {st['ms_total']:.2f} "
          f"panel {st["ms_stage"]["panel"]:.2f} trailing {st['ms_stage']['trailing']:.2f} ms = {tf:.2f} TFLOP/s, launches {st['launches']}{same}",
          flush=True)
