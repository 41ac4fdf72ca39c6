"""One factorisation through the ROW-SHARDED code path on a 1-rank peer context (QRDM_B200_FORCE_MG), for an ncu launch
list of the sharded kernels at one rank's share of configs[3]:  python tools/one_run_mg.py 250016 512"""
import ctypes as C
import os
import sys

os.environ["QRDM_B200_FORCE_MG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qrdm_b200  # noqa: E402
from qrdm_b200 import _lib, sharded  # noqa: E402

m = int(sys.argv[1]); n = int(sys.argv[2])
_lib.lib.qrdm_b200_init(0)
h = C.create_string_buffer(64)
assert _lib.lib.qrdm_b200_peer_handle(h) == 0 and _lib.lib.qrdm_b200_peer_open(0, 1, h.raw) == 0
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
A0 = torch.randn((n, m), dtype=torch.float64, device="cuda", generator=gen)
jp = torch.zeros(n, dtype=torch.int32, device="cuda"); tau = torch.zeros(n, dtype=torch.float64, device="cuda")
for rep in range(2):
    A = A0.clone()
    info, nc = sharded.dgeqrdm_sharded(A, m, m, 0, 1, n, m, jp, tau)
    torch.cuda.synchronize()
    st = qrdm_b200.stats()
    print(f"{m}x{n} sharded path (1 rank): info {info} rank {int(nc.sum())} ms_total {st['ms_total']:.2f} launches {st['launches']}")
