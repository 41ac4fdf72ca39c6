"""hostio.c (the pageable-host bounce pipeline of the reference-facing entry point) on a CPU-only box: the same
source linked against a mock runtime whose asynchronous copies are deferred until somebody waits for them
(tests/support/mock_rt.c).  Covers: column-group tiles, row-segment tiles of very long columns, lda > m on both
sides, ragged sizes, the streamed write-back ring (full ring, partial takes, tail download)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "qrdm_b200", "csrc")


@pytest.fixture(scope="module")
def hio(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostio") / "libhostio_mock.so"
    subprocess.run(["gcc", "-O1", "-g", "-fPIC", "-std=gnu11", "-shared", "-I", CSRC, "-o", str(out),
                    os.path.join(CSRC, "hostio.c"), os.path.join(ROOT, "tests", "support", "mock_rt.c"), "-lpthread"],
                   check=True)
    lib = C.CDLL(str(out))
    vp, ip = C.c_void_p, C.c_int
    lib.qrdm_hostio_create.argtypes = [C.POINTER(vp), ip]
    lib.qrdm_hostio_destroy.argtypes = [vp]
    lib.qrdm_hostio_destroy.restype = None
    lib.qrdm_hostio_upload.argtypes = [vp, vp, ip, vp, ip, ip, ip]
    lib.qrdm_hostio_download.argtypes = [vp, vp, ip, vp, ip, ip, ip, ip]
    lib.qrdm_hostio_wb_begin.argtypes = [vp, vp, ip, vp, ip, ip, vp]
    lib.qrdm_hostio_wb_push.argtypes = [vp, ip, ip]
    lib.qrdm_hostio_wb_end.argtypes = [vp]
    lib.qrdm_rt_stream_create.argtypes = [C.POINTER(vp)]
    h = vp()
    assert lib.qrdm_hostio_create(C.byref(h), 0) == 0
    yield lib, h
    lib.qrdm_hostio_destroy(h)


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("m,n,ldh,ldd", [(7, 5, 7, 8), (1000, 333, 1003, 1000), (1, 1, 1, 2), (4097, 70, 4097, 4098),
                                         (1_300_000, 3, 1_300_001, 1_300_000),   # a column longer than a bounce buffer
                                         (1_048_576, 2, 1_048_576, 1_048_576)])  # exactly one buffer per column
def test_upload_download_roundtrip(hio, m, n, ldh, ldd):
    lib, h = hio
    rng = np.random.default_rng(m + n)
    H = np.full((n, ldh), -7.0)
    H[:, :m] = rng.standard_normal((n, m))          # row c of H = column c of the matrix (column-major, lda = ldh)
    D = np.full((n, ldd), 3.0)
    assert lib.qrdm_hostio_upload(h, _p(D), ldd, _p(H), ldh, m, n) == 0
    assert np.array_equal(D[:, :m], H[:, :m]) and np.all(D[:, m:] == 3.0)
    D[:, :m] *= 2.0
    H2 = np.full((n, ldh), -7.0)
    c0 = n // 3
    assert lib.qrdm_hostio_download(h, _p(H2), ldh, _p(D), ldd, m, c0, n) == 0
    assert np.array_equal(H2[c0:, :m], D[c0:, :m])
    assert np.all(H2[:c0] == -7.0) and np.all(H2[:, m:] == -7.0)      # nothing outside columns [c0, n) x rows [0, m)


@pytest.mark.parametrize("m,n,step", [(500, 700, 64), (16384, 300, 64), (100, 5000, 640), (3, 130, 1)])
def test_streamed_writeback(hio, m, n, step):
    lib, h = hio
    rng = np.random.default_rng(m * 3 + n)
    D = rng.standard_normal((n, m + 2))
    H = np.zeros((n, m + 5))
    s = C.c_void_p()
    assert lib.qrdm_rt_stream_create(C.byref(s)) == 0
    assert lib.qrdm_hostio_wb_begin(h, _p(H), m + 5, _p(D), m + 2, m, s) == 0
    done = 0
    for j in range(step, n + 1, step):              # "iterations": columns [done, j) become final
        taken = lib.qrdm_hostio_wb_push(h, done, j)
        assert 0 <= taken <= j - done
        done += taken
    assert lib.qrdm_hostio_wb_end(h) == 0
    assert np.array_equal(H[:done, :m], D[:done, :m]) and np.all(H[done:] == 0) and np.all(H[:, m:] == 0)
    assert lib.qrdm_hostio_download(h, _p(H), m + 5, _p(D), m + 2, m, done, n) == 0   # the tail
    assert np.array_equal(H[:, :m], D[:, :m])


def test_writeback_declines_huge_columns(hio):
    lib, h = hio
    m = 2_000_000                                   # 64 columns = 1 GB: more than the ring is worth
    x = np.zeros(8)
    assert lib.qrdm_hostio_wb_begin(h, _p(x), m, _p(x), m, m, None) == 1
