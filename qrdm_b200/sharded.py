"""1-D block-row sharded dgeqrdm across the GPUs of one node, one process per GPU.

`torch.distributed` is only the launcher-side plumbing (rendezvous, broadcasting the NCCL unique
id, all-gathering the CUDA-IPC handles of the peer receive buffers); the data-path exchanges are
issued by the C host driver on the compute stream (`dgeqrdm_dev_sharded` in include/qrdm_b200.h):
LL packets over NVLink peer memory from inside the panel kernel and a one-kernel LL all-reduce for
the small vectors, `ncclAllReduce` for the bandwidth-bound ones.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def row_partition(m: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row blocks [row0, row0 + rows) per rank, multiples of 32 rows except the last
    (row tiles of the trailing kernels start at multiples of 32)."""
    if world < 1 or m < 0:
        raise ValueError("bad partition request")
    per = -(-m // world)
    per = -(-per // 32) * 32
    out = []
    for r in range(world):
        lo = min(m, r * per)
        hi = min(m, lo + per)
        out.append((lo, hi - lo))
    return out


def batch_partition(batch: int, world: int) -> list[tuple[int, int]]:
    """Batched mode (independent units, no collective): matrices [first, first + count) per rank; the
    first ``batch % world`` ranks take one extra matrix."""
    if world < 1 or batch < 0:
        raise ValueError("bad partition request")
    base, extra = divmod(batch, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def broadcast_unique_id(make_id, rank: int, group=None, device=None) -> bytes:
    """Rank 0 produces the 128-byte NCCL unique id (`make_id()`), everybody receives it through
    torch.distributed (works with the gloo and the nccl backend)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def allgather_handles(my_handle: bytes, rank: int, world: int, group=None, device=None) -> bytes:
    """All-gather of the 64-byte CUDA-IPC handles, rank order (gloo or nccl backend)."""
    import torch
    import torch.distributed as dist
    if len(my_handle) != 64:
        raise ValueError("a CUDA IPC handle is 64 bytes")
    mine = torch.frombuffer(bytearray(my_handle), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def init_peer(rank: int, world: int, group=None, device=None) -> None:
    """Open NVLink peer memory between the ranks (include/qrdm_b200.h: qrdm_b200_peer_handle / _open)."""
    import torch.distributed as dist
    from . import _lib
    raw = C.create_string_buffer(64)
    if _lib.lib.qrdm_b200_peer_handle(raw) != 0:
        raise RuntimeError("qrdm_b200_peer_handle failed (cudaIpcGetMemHandle)")
    allh = allgather_handles(raw.raw, rank, world, group=group, device=device)
    if _lib.lib.qrdm_b200_peer_open(int(rank), int(world), allh) != 0:
        raise RuntimeError("qrdm_b200_peer_open failed (cudaIpcOpenMemHandle: no peer access between the GPUs?)")
    dist.barrier(group=group)   # nobody sends before everybody has mapped and zeroed


def init_comm(rank: int, world: int, group=None, device=None, nccl: bool = True, peer: bool = True) -> None:
    """Set up the library's transports for this process: the NCCL communicator and NVLink peer memory."""
    from . import _lib

    def make_id():
        raw = C.create_string_buffer(128)
        if _lib.lib.qrdm_b200_comm_unique_id(raw) != 0:
            raise RuntimeError("qrdm_b200_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return raw.raw

    if nccl:
        uid = broadcast_unique_id(make_id, rank, group=group, device=device)
        if _lib.lib.qrdm_b200_comm_init(int(rank), int(world), uid) != 0:
            raise RuntimeError("qrdm_b200_comm_init failed")
    if peer:
        init_peer(rank, world, group=group, device=device)


def dgeqrdm_sharded(dA_local, m_local, m_global, row0, world, n, lda, d_jpvt, d_tau,
                    thres=(0.9, 0.15), nb=64, stop_mode=0, stream=None):
    """Factor the row-sharded matrix in place; returns (info, ncols).  `dA_local` is this rank's
    (n, lda) row-major float64 CUDA tensor = its m_local x n column-major row block."""
    from . import _lib
    import torch
    ncols = np.zeros(max(n, 1), dtype=np.int32)
    ncols[0] = stop_mode
    th = np.zeros(3, dtype=np.float64)
    th[: len(thres)] = thres
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    info = _lib.lib.dgeqrdm_dev_sharded(int(m_local), int(m_global), int(row0), int(world), int(n),
                                        C.c_void_p(int(dA_local.data_ptr())), int(lda),
                                        C.c_void_p(int(d_jpvt.data_ptr())), C.c_void_p(int(d_tau.data_ptr())),
                                        ncols.ctypes.data, th.ctypes.data, int(nb), C.c_void_p(int(stream)))
    return int(info), ncols
