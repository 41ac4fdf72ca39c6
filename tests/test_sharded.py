"""Row-sharded (multi-GPU) path, SURVEY.md §8e.

CPU part: the host-side plumbing under torch.distributed with the gloo backend, world_size 2 —
row partition, the broadcast of the NCCL unique id, and max-over-ranks timing reduction.
GPU part (one GPU): the sharded CODE PATH (per-column panel kernels, W-slot fold, split norm
update, ncclAllReduce on a 1-rank communicator) against the unmodified reference.
The 2/4/8-GPU runs are exercised by tools/mg_check.py and bench.py under torchrun.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batch_partition_covers_all_matrices():
    from qrdm_b200.sharded import batch_partition
    for batch, w in [(8192, 8), (8192, 1), (10, 3), (2, 4), (0, 2), (1184, 5)]:
        parts = batch_partition(batch, w)
        assert len(parts) == w and sum(c for _, c in parts) == batch
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        batch_partition(4, 0)


def test_row_partition_covers_all_rows():
    from qrdm_b200.sharded import row_partition
    for m, w in [(2_000_000, 8), (2_000_000, 4), (100, 8), (31, 2), (0, 3), (4096, 1), (257, 2)]:
        parts = row_partition(m, w)
        assert len(parts) == w
        pos = 0
        for lo, rows in parts:
            assert lo == min(pos, m) and rows >= 0
            assert lo % 32 == 0 or rows == 0
            pos = lo + rows
        assert pos == m
    with pytest.raises(ValueError):
        row_partition(10, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from qrdm_b200.sharded import broadcast_unique_id, row_partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fake = bytes(range(128))                      # stands in for ncclGetUniqueId on a CPU box
    uid = broadcast_unique_id(lambda: fake, rank)
    # the partial-sum -> all-reduce -> replicated-decision pattern of the sharded driver, in numpy
    rng = np.random.default_rng(0)
    A = rng.standard_normal((100, 7))
    lo, rows = row_partition(100, world)[rank]
    part = torch.from_numpy((A[lo:lo + rows] ** 2).sum(axis=0))
    dist.all_reduce(part)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # bench.py: max over ranks
    q.put((rank, uid == fake, np.allclose(part.numpy(), (A ** 2).sum(axis=0)), float(t)))
    dist.destroy_process_group()


def test_gloo_world2_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] and r[2] for r in res)
    assert all(r[3] == 2.0 for r in res)


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl_per_column", "peer_fused_panel"])
def test_sharded_code_path_on_one_gpu(transport, oracle_ref, monkeypatch):
    """QRDM_B200_FORCE_MG routes a 1-rank job through every sharded kernel: once with the NCCL-only transport (one
    kernel + ncclAllReduce per panel column, the fallback when no peer memory is open) and once with peer memory open
    (k_panel_tall<true>: exchange inside the persistent sub-panel kernel).  world_size > 1 is covered by
    tests/test_gpu_sharded_mp.py."""
    import ctypes as C
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    from qrdm_b200 import _lib, generators as g, sharded
    raw = C.create_string_buffer(128)
    assert _lib.lib.qrdm_b200_comm_unique_id(raw) == 0
    assert _lib.lib.qrdm_b200_comm_init(0, 1, raw.raw) == 0
    if transport == "peer_fused_panel":
        h = C.create_string_buffer(64)
        assert _lib.lib.qrdm_b200_peer_handle(h) == 0
        assert _lib.lib.qrdm_b200_peer_open(0, 1, h.raw) == 0
    monkeypatch.setenv("QRDM_B200_FORCE_MG", "1")
    try:
        for A, kw in [(g.gaussian(700, 300, 21), {}), (g.gaussian(300, 450, 22), {}),
                      (g.kahan(130), {}), (g.gaussian(5000, 160, 23), dict(nb=32, thres=(0.7, 0.3))),
                      (g.gaussian(40000, 100, 24), {})]:   # > 16384 rows per rank: blocked sharded panel
            m, n = A.shape
            lda = m + (m & 1)
            loc = torch.zeros((n, lda), dtype=torch.float64, device="cuda")
            loc[:, :m] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
            jp = torch.zeros(n, dtype=torch.int32, device="cuda")
            tau = torch.zeros(min(m, n), dtype=torch.float64, device="cuda")
            info, ncols = sharded.dgeqrdm_sharded(loc, m, m, 0, 1, n, lda, jp, tau, **kw)
            got = dict(info=info, A=loc.cpu().numpy().T[:m, :], jpvt=jp.cpu().numpy(), tau=tau.cpu().numpy(), ncols=ncols)
            exp = oracle_ref.ref_dgeqrdm(A, **kw)
            parity.check_against(got, exp, (m, n), exact=True)
            if max(m, n) <= 5000:
                res, orth = parity.qr_invariants(A, got)
                tol = parity.invariant_tol(A.shape)
                assert res <= tol and orth <= tol, (res, orth, tol)
    finally:
        _lib.lib.qrdm_b200_peer_close()
        _lib.lib.qrdm_b200_comm_destroy()
