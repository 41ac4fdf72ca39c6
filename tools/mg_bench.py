"""Row-sharded configs[3] under torchrun (one rank per GPU): timing of the two transports (peer-memory LL exchange
inside the panel kernel vs the round-1 ncclAllReduce per panel column) and the parity object of bench.py.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mg_bench.py [rows]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qrdm_b200  # noqa: E402
from qrdm_b200 import _lib, sharded  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
assert _lib.lib.qrdm_b200_init(lr) == 0
sharded.init_comm(rank, world, device=dev)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
n = 512
row0, ml = sharded.row_partition(m, world)[rank]
gen = torch.Generator(device=dev); gen.manual_seed(4321 + rank)
A0 = torch.randn((n, ml), dtype=torch.float64, device=dev, generator=gen)
A = torch.empty_like(A0)
jp = torch.zeros(n, dtype=torch.int32, device=dev); tau = torch.zeros(n, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream()
out = {"world": world, "rows": m, "rows_per_gpu": ml}
for mode in ("peer", "legacy_nccl_per_column", "peer"):
    os.environ.pop("QRDM_B200_MG_LEGACY", None)
    if mode.startswith("legacy"):
        os.environ["QRDM_B200_MG_LEGACY"] = "1"
    times = []
    for it in range(4):
        A.copy_(A0)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        info, nc = sharded.dgeqrdm_sharded(A, ml, m, row0, world, n, ml, jp, tau, stream=stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
        assert info == 0, info
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(float(t))
    out.setdefault(mode, []).append({"ms": times, "launches": qrdm_b200.stats()["launches"], "rank": int(nc.sum())})
os.environ.pop("QRDM_B200_MG_LEGACY", None)
# per-stage split of the sharded path (profile mode 1: event pair + sync around every stage, every rank alike)
qrdm_b200.set_profile(1)
A.copy_(A0)
torch.cuda.synchronize(); dist.barrier()
info, nc = sharded.dgeqrdm_sharded(A, ml, m, row0, world, n, ml, jp, tau, stream=stream.cuda_stream)
st = qrdm_b200.stats()
qrdm_b200.set_profile(0)
out["stage_profile_rank0"] = {"ms_total": st["ms_total"], "ms_stage": st["ms_stage"], "stage_launches": st["stage_launches"]}
del A0, A
torch.cuda.empty_cache()
out["parity_vs_single_gpu"] = bench.sharded_parity(torch, dist, qrdm_b200, rank, world, dev)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
