"""Device-resident timing of the batched mode (one CTA per matrix).  usage: batched_bench.py [batch] [n] [kind]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qrdm_b200
from qrdm_b200 import generators as g

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
kind = sys.argv[3] if len(sys.argv) > 3 else "kahan"
distinct = min(batch, 37)
if kind == "kahan":
    base = np.stack([g.kahan(n, theta=1.1 + 0.2 * b / distinct, perturb=1e3, seed=b).T.copy() for b in range(distinct)])
else:
    base = np.stack([g.gaussian(n, n, b).T.copy() for b in range(distinct)])
d_base = torch.from_numpy(base).cuda()
idx = torch.arange(batch, device="cuda") % distinct
d_a = d_base[idx].contiguous()
d_jpvt = torch.zeros((batch, n), dtype=torch.int32, device="cuda")
d_tau = torch.zeros((batch, n), dtype=torch.float64, device="cuda")
d_ncols = torch.zeros((batch, n), dtype=torch.int32, device="cuda")
d_infos = torch.zeros((batch,), dtype=torch.int32, device="cuda")
for rep in range(3):
    d_a.copy_(d_base[idx]); d_ncols.zero_(); torch.cuda.synchronize()
    rc = qrdm_b200.api.dgeqrdm_batched_device(batch, n, n, d_a.data_ptr(), n, n * n, d_jpvt.data_ptr(), d_tau.data_ptr(),
                                             d_ncols.data_ptr(), d_infos.data_ptr())
    ms = qrdm_b200.stats()["ms_total"]
    rank = d_ncols.sum(dim=1)
    its = (d_ncols > 0).sum(dim=1)
    fl = sum(g.flops(n, n, int(r)) for r in rank[:distinct].tolist()) / distinct * batch
    print(f"{kind} n={n} batch={batch}: rc={rc} {ms:.1f} ms  {batch / ms * 1e3:.0f} matrices/s  {fl / ms / 1e6:.1f} GFLOP/s "
          f"mean iterations {its.float().mean().item():.1f} infos!=0: {int((d_infos != 0).sum())}")
