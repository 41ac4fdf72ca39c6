"""Row-sharded dgeqrdm with world_size > 1 in the GPU suite (the sharded path was only ever run by
pytest on a 1-rank communicator, so row0 > 0, ranks without active rows and cross-rank reductions were untested).

The driver's GPU suite has ONE GPU.  NCCL refuses two ranks on one device, but the library's own transport does
not: the ranks are separate processes that share cuda:0 (time-sliced), map each other's receive buffers with CUDA
IPC and exchange LL packets exactly as they would over NVLink (`QRDM_B200_COLL=peer` routes every all-reduce of
the sharded driver through that transport; rendezvous and handle exchange over gloo).  Every case is compared,
rank by rank, with the single-GPU factorisation of the same matrix AND with the unmodified reference:
jpvt / ncols / tau identical on every rank, each rank's rows of the factored matrix equal to rounding.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _cases():
    from qrdm_b200 import generators as g
    return [
        ("gauss600x200", g.gaussian(600, 200, 0), {}),
        ("gauss257x300_wide_ragged", g.gaussian(257, 300, 3), {}),           # m not a multiple of 32 * world
        ("gauss100x80_rank_runs_dry", g.gaussian(100, 80, 5), {}),          # rank 0 has no active rows once j >= 64
        ("gauss3000x384", g.gaussian(3000, 384, 2), {}),
        ("gauss40000x96_tall", g.gaussian(40000, 96, 7), {}),               # > 16384 rows per rank
        ("gauss1500x260_nb24", g.gaussian(1500, 260, 8), dict(nb=24, thres=(0.5, 0.6))),
        ("kahan200", g.kahan(200), {}),                                     # 199 one-column iterations
        ("graded256_stop1", g.graded(256, seed=2), dict(stop_mode=1)),      # early stop inside the panel + stop rule
        ("gauss600x200_times_1e200", g.gaussian(600, 200, 0) * 1e200, {}),  # pre-scaling: every rank must pick the same factor
    ]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        os.environ["QRDM_B200_COLL"] = "peer"      # all collectives through the library's own peer transport
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch
        import torch.distributed as dist
        import parity
        import qrdm_b200
        from qrdm_b200 import _lib, sharded
        from oracle import ref
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        assert _lib.lib.qrdm_b200_init(0) == 0
        sharded.init_comm(rank, world, nccl=False, peer=True)
        out = []
        for name, A, kw in _cases():
            m, n = A.shape
            row0, ml = sharded.row_partition(m, world)[rank]
            # single-GPU result of the same matrix (every rank computes it: it is also the time-slicing partner's load)
            dA = torch.from_numpy(np.ascontiguousarray(A.T)).to(dev)
            jp1 = torch.zeros(n, dtype=torch.int32, device=dev)
            tau1 = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
            info1, nc1 = qrdm_b200.dgeqrdm_device(dA, m, n, m, jp1, tau1, **kw)
            lda = max(ml, 2) + (ml & 1)
            loc = torch.full((n, lda), 9.5, dtype=torch.float64, device=dev)
            if ml > 0:
                loc[:, :ml] = torch.from_numpy(np.ascontiguousarray(A[row0:row0 + ml, :].T)).to(dev)
            jp = torch.zeros(n, dtype=torch.int32, device=dev)
            tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            dist.barrier()
            info, nc = sharded.dgeqrdm_sharded(loc, ml, m, row0, world, n, lda, jp, tau, **kw)
            if info != 0:   # the other ranks are waiting for this one: say which case before they trap
                raise RuntimeError(f"rank {rank}: dgeqrdm_dev_sharded returned {info} on {name} (rows [{row0}, {row0 + ml}))")
            torch.cuda.synchronize()
            F1 = dA.cpu().numpy().T                       # m x n single-GPU factor
            d1 = np.diag(F1)[: min(m, n)]
            # graded input: pivots are only reproducible on the trusted prefix (tests/parity.py: |R_jj| above the noise
            # floor and every decision of the block separated by > 1e-12, margins logged by the C port)
            margins = ref.port_dgeqrdm(A, **kw)["margins"] if name.startswith("graded") else None
            nblk, ncol = parity.trusted_prefix(nc1, d1, (m, n), margins)
            full = nblk == int(np.count_nonzero(nc1))
            rows = loc.cpu().numpy()[:, :ml].T            # this rank's rows of the sharded factor
            mx = float(np.abs(F1).max()) or 1.0           # norms of 1e200-sized entries overflow: normalise first
            scale = np.linalg.norm(F1 / mx) or 1.0
            rel = float(np.linalg.norm((rows - F1[row0:row0 + ml, :]) / mx) / scale) if ml > 0 else 0.0
            pad_ok = bool(np.all(loc.cpu().numpy()[:, ml:] == 9.5))
            # against the unmodified reference: assemble nothing — jpvt / ncols / tau / diag(R) live replicated or in
            # the rank that owns the diagonal rows; compare the replicated outputs here
            exp = ref.ref_dgeqrdm(A, **kw)
            r = ncol
            rec = dict(case=name, rank=rank, info=int(info), info1=int(info1),
                       ncols_eq_single=bool(np.array_equal(nc[:nblk], nc1[:nblk])),
                       jpvt_eq_single=bool(np.array_equal(jp.cpu().numpy()[:ncol], jp1.cpu().numpy()[:ncol])),
                       ncols_eq_ref=bool(np.array_equal(nc[:nblk], exp["ncols"][:nblk])),
                       jpvt_eq_ref=bool(np.array_equal(jp.cpu().numpy()[:ncol], exp["jpvt"][:ncol])),
                       tau_close=bool(np.allclose(tau.cpu().numpy()[:r], exp["tau"][:r], rtol=1e-9, atol=1e-13)),
                       rows_rel_diff=rel if full else 0.0, pad_ok=pad_ok, full=bool(full), blocks=int(nblk),
                       perm_ok=sorted(jp.cpu().numpy().tolist()) == list(range(1, n + 1)),
                       rank_sum=int(nc.sum()), rank_sum_single=int(nc1.sum()),
                       ncols=nc[:np.count_nonzero(nc)].tolist(), ncols_single=nc1[:np.count_nonzero(nc1)].tolist(),
                       ncols_ref=exp["ncols"][:np.count_nonzero(exp["ncols"])].tolist())
            out.append(rec)
        dist.barrier()
        _lib.lib.qrdm_b200_peer_close()
        dist.destroy_process_group()
        q.put((rank, out, None))
    except BaseException as exc:  # noqa: BLE001 - the parent must always get an answer
        import traceback
        q.put((rank, None, traceback.format_exc() + repr(exc)))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranks_share_one_gpu(world):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU suite selected but no CUDA device")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, errs = {}, []
    try:
        for _ in procs:
            rank, out, err = q.get(timeout=900)
            if err is not None:
                errs.append(f"rank {rank} failed:\n{err}")
            res[rank] = out
        assert not errs, "\n".join(errs)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert sorted(res) == list(range(world))
    ncase = len(res[0])
    for ci in range(ncase):
        for r in range(world):
            e = res[r][ci]
            msg = "\n".join(f"{k}: {v}" for k, v in e.items())
            assert e["info"] == 0 and e["info1"] == 0, msg
            assert e["ncols_eq_single"] and e["jpvt_eq_single"], msg
            assert e["ncols_eq_ref"] and e["jpvt_eq_ref"] and e["tau_close"], msg
            assert e["perm_ok"] and e["pad_ok"], msg
            assert e["rows_rel_diff"] < 1e-12, msg
            assert e["blocks"] >= 1
        # replicated decisions: every rank reports the same revealed rank
        assert len({res[r][ci]["rank_sum"] for r in range(world)}) == 1
