"""Multi-GPU development check (run under torchrun, one rank per GPU): row-sharded dgeqrdm vs the
single-GPU result on the same seeded matrix."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qrdm_b200  # noqa: E402
from qrdm_b200 import generators as g, sharded  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from qrdm_b200 import _lib  # noqa: E402
_lib.lib.qrdm_b200_init(lr)
sharded.init_comm(rank, world, device=dev)

def check(name, A, **kw):
    m, n = A.shape
    parts = sharded.row_partition(m, world)
    row0, ml = parts[rank]
    # single-GPU result (every rank computes it for comparison)
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).to(dev)
    jp1 = torch.zeros(n, dtype=torch.int32, device=dev); tau1 = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
    info1, nc1 = qrdm_b200.dgeqrdm_device(dA, m, n, m, jp1, tau1, **kw)
    # sharded
    lda = max(ml, 2) + (ml & 1)
    loc = torch.zeros((n, lda), dtype=torch.float64, device=dev)
    loc[:, :ml] = torch.from_numpy(np.ascontiguousarray(A[row0:row0 + ml, :].T)).to(dev)
    jp = torch.zeros(n, dtype=torch.int32, device=dev); tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    info, nc = sharded.dgeqrdm_sharded(loc, ml, m, row0, world, n, lda, jp, tau, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # graded inputs: only the trusted prefix (blocks above the rounding-noise floor) is reproducible
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    d1 = torch.diagonal(dA.T)[: min(m, n)].cpu().numpy()
    nblk, ncol = parity.trusted_prefix(nc1, d1, (m, n))
    ok_nc = np.array_equal(nc[:nblk], nc1[:nblk]); ok_jp = bool(torch.equal(jp[:ncol], jp1[:ncol]))
    full = nblk == int(np.count_nonzero(nc1))
    ref_rows = dA.T[row0:row0 + ml, :]
    got_rows = loc.T[:ml, :]
    scale = float(torch.linalg.norm(dA)) or 1.0
    err = float(torch.linalg.norm(got_rows - ref_rows)) / scale if ml > 0 else 0.0
    terr = float(torch.max(torch.abs(tau - tau1))) if min(m, n) > 0 else 0.0
    print(f"[rank {rank}] {name:22s} info {info}/{info1} rank {int(nc.sum())}/{int(nc1.sum())} ncols_eq {ok_nc} jpvt_eq {ok_jp} "
          f"rows[{row0}:{row0+ml}] rel.diff {err:.2e} tau diff {terr:.2e} time {dt*1e3:.1f} ms", flush=True)
    return ok_nc and ok_jp and (err < 1e-10 or not full)

ok = True
ok &= check("gauss600x200", g.gaussian(600, 200, 0))
ok &= check("gauss500x500", g.gaussian(500, 500, 1))
ok &= check("gauss3000x384", g.gaussian(3000, 384, 2))
ok &= check("gauss257x300_wide", g.gaussian(257, 300, 3))
ok &= check("kahan200", g.kahan(200))
ok &= check("graded256_stop1", g.graded(256, seed=2), stop_mode=1)
if len(sys.argv) > 1 and sys.argv[1] == "big":
    ok &= check("gauss200000x512", g.gaussian(200000, 512, 4))
    ok &= check("gauss4096", g.gaussian(4096, 4096, 5))
print(f"[rank {rank}] ALL OK" if ok else f"[rank {rank}] MISMATCH", flush=True)
dist.destroy_process_group()
