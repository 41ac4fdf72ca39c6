"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libqrdm_ref.so,
compiled from /root/reference by oracle/Makefile) on the seeded inputs of cases.py.

Dev-container only (needs /root/reference to build the .so).  Usage:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden.cases import CASES  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    ref.build()
    out = os.path.dirname(os.path.abspath(__file__))
    for name, c in CASES.items():
        A = c["make"]()
        r = ref.ref_dgeqrdm(A, thres=c["thres"], nb=c["nb"], stop_mode=c["stop_mode"])
        k = min(A.shape)
        payload = dict(info=np.int32(r["info"]), jpvt=r["jpvt"], ncols=r["ncols"], tau=r["tau"],
                       diagR=np.diag(r["A"])[:k].copy(), shape=np.array(A.shape, dtype=np.int64),
                       checksum=np.float64(np.nansum(np.abs(A[np.isfinite(A)]))))
        if c["store_input"]:
            payload["A"] = A
        np.savez_compressed(os.path.join(out, name + ".npz"), **payload)
        print(f"{name:28s} info {r['info']:4d} rank {int(r['ncols'].sum()):4d} blocks "
              f"{r['ncols'][:np.count_nonzero(r['ncols'])].tolist()[:10]}")


if __name__ == "__main__":
    main()
