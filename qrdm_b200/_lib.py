"""ctypes binding of libqrdm_b200.so (the C ABI declared in include/qrdm_b200.h).

There is no fallback of any kind: if the library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C qrdm_b200/csrc``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libqrdm_b200.so")


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("rank", C.c_int), ("launches", C.c_longlong),
                ("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("ms_stage", C.c_double * 12), ("stage_launches", C.c_longlong * 12),
                ("trailing_flops", C.c_double), ("panel_cols", C.c_double),
                ("stage_bytes", C.c_double * 12), ("side_flops", C.c_double), ("side_launches", C.c_longlong),
                ("fused_flops", C.c_double), ("fused_launches", C.c_longlong)]


STAGES = ["norm_init", "select", "gram", "pick", "permute", "panel", "vtv", "trailing", "wsolve",
          "rankk", "norm_update", "sync"]

# every symbol include/qrdm_b200.h declares (checked by tests/test_abi.py)
EXPORTS = ["dgeqrdm", "dgeqrdm_work", "dgeqrdm_dev", "dgeqrdm_dev_sharded", "dgeqrdm_batched", "dgeqrdm_batched_dev", "qrdm_b200_get_stats",
           "qrdm_b200_set_profile", "qrdm_b200_init", "qrdm_b200_shutdown", "qrdm_b200_measure_fp64_peak",
           "qrdm_b200_version", "qrdm_b200_comm_unique_id", "qrdm_b200_comm_init", "qrdm_b200_comm_destroy",
           "qrdm_b200_dormqr", "qrdm_b200_dormqr_dev", "qrdm_b200_peer_handle", "qrdm_b200_peer_open",
           "qrdm_b200_peer_close", "qrdm_b200_measure_copy_gbs"]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C qrdm_b200/csrc` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    sig = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
           C.c_int]
    for name in ("dgeqrdm", "dgeqrdm_work"):
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = sig
    lib.dgeqrdm_dev.restype = C.c_int
    lib.dgeqrdm_dev.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_void_p]
    lib.qrdm_b200_get_stats.restype = None
    lib.qrdm_b200_get_stats.argtypes = [C.POINTER(Stats)]
    lib.qrdm_b200_set_profile.restype = None
    lib.qrdm_b200_set_profile.argtypes = [C.c_int]
    lib.qrdm_b200_init.restype = C.c_int
    lib.qrdm_b200_init.argtypes = [C.c_int]
    lib.qrdm_b200_shutdown.restype = None
    lib.qrdm_b200_measure_fp64_peak.restype = C.c_double
    lib.qrdm_b200_measure_fp64_peak.argtypes = [C.c_int, C.c_void_p]
    lib.qrdm_b200_version.restype = C.c_char_p
    lib.qrdm_b200_comm_unique_id.restype = C.c_int
    lib.qrdm_b200_comm_unique_id.argtypes = [C.c_char_p]
    lib.qrdm_b200_comm_init.restype = C.c_int
    lib.qrdm_b200_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    lib.qrdm_b200_comm_destroy.restype = None
    lib.qrdm_b200_peer_handle.restype = C.c_int
    lib.qrdm_b200_peer_handle.argtypes = [C.c_char_p]
    lib.qrdm_b200_peer_open.restype = C.c_int
    lib.qrdm_b200_peer_open.argtypes = [C.c_int, C.c_int, C.c_char_p]
    lib.qrdm_b200_peer_close.restype = None
    lib.qrdm_b200_measure_copy_gbs.restype = C.c_double
    lib.qrdm_b200_measure_copy_gbs.argtypes = [C.c_size_t, C.c_void_p]
    lib.dgeqrdm_batched.restype = C.c_int
    lib.dgeqrdm_batched.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.dgeqrdm_batched_dev.restype = C.c_int
    lib.dgeqrdm_batched_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.dgeqrdm_dev_sharded.restype = C.c_int
    lib.dgeqrdm_dev_sharded.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.qrdm_b200_dormqr_dev.restype = C.c_int
    lib.qrdm_b200_dormqr_dev.argtypes = [C.c_char, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p]
    lib.qrdm_b200_dgeqp3.restype = C.c_int
    lib.qrdm_b200_dgeqp3.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.qrdm_b200_dgeqp3_dev.restype = C.c_int
    lib.qrdm_b200_dgeqp3_dev.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qrdm_b200_dormqr.restype = C.c_int
    lib.qrdm_b200_dormqr.argtypes = [C.c_char, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_int]
    return lib


lib = _load()


def stats() -> dict:
    s = Stats()
    lib.qrdm_b200_get_stats(C.byref(s))
    return dict(iterations=s.iterations, rank=s.rank, launches=s.launches, ms_total=s.ms_total,
                ms_h2d=s.ms_h2d, ms_d2h=s.ms_d2h, trailing_flops=s.trailing_flops, panel_cols=s.panel_cols,
                ms_stage={STAGES[i]: s.ms_stage[i] for i in range(12)},
                stage_launches={STAGES[i]: s.stage_launches[i] for i in range(12)},
                stage_bytes={STAGES[i]: s.stage_bytes[i] for i in range(12)},
                side_flops=s.side_flops, side_launches=s.side_launches,
                fused_flops=s.fused_flops, fused_launches=s.fused_launches)
