"""Wall time of the reference-facing entry point `dgeqrdm` for the three host-buffer modes (pinned, pageable through the
bounce pipeline of hostio.c, pageable with plain cudaMemcpy2D) — development probe for bench.py's e2e figures."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qrdm_b200  # noqa: E402
from qrdm_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = n
rng = np.random.default_rng(0)
A0 = np.asfortranarray(rng.standard_normal((m, n)))
th = np.array([0.9, 0.15, 0.0])


def run(buf_ptr, label, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    jp = np.zeros(n, dtype=np.int32); tau = np.zeros(n); nc = np.zeros(n, dtype=np.int32)
    t0 = time.perf_counter()
    info = _lib.lib.dgeqrdm(102, m, n, buf_ptr, m, jp.ctypes.data, tau.ctypes.data, nc.ctypes.data, th.ctypes.data, 64)
    dt = time.perf_counter() - t0
    st = qrdm_b200.stats()
    for k in (env or {}):
        os.environ.pop(k, None)
    print(f"{label:34s} info {info} wall {dt*1e3:8.1f} ms  device {st['ms_total']:7.1f} ms  h2d {st['ms_h2d']:6.1f}  d2h-tail {st['ms_d2h']:6.1f}", flush=True)
    return jp, tau, nc


hA = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
for rep in range(2):
    hA.copy_(torch.from_numpy(A0.T))
    ref = run(hA.data_ptr(), "pinned")
    ref_A = hA.numpy().copy()
    A = A0.copy(order="F")
    out = run(A.ctypes.data, "pageable, bounce pipeline")
    assert np.array_equal(A.T, ref_A) and all(np.array_equal(a, b) for a, b in zip(out, ref)), "bounce result differs"
    A = A0.copy(order="F")
    out = run(A.ctypes.data, "pageable, bounce, no overlap", {"QRDM_B200_NO_OVERLAP": "1"})
    assert np.array_equal(A.T, ref_A)
    A = A0.copy(order="F")
    out = run(A.ctypes.data, "pageable, plain cudaMemcpy2D", {"QRDM_B200_NO_BOUNCE": "1"})
    assert np.array_equal(A.T, ref_A)
print("results identical across the three transfer modes")
