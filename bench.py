#!/usr/bin/env python
"""bench.py — the dgeqrdm FP64 hot path on B200, one JSON line (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|C1|C4] [--impl reference]

A "step" is one complete dgeqrdm factorisation of one synthetic matrix.  N = 1 runs the
workload on one GPU; N > 1 (under torchrun) runs N independent replicas, one matrix per rank
("batched mode spreads independent matrices across GPUs": no data-path collective, weak scaling).

value   = whole-job GFLOP/s with A resident in HBM (dgeqrdm_dev, CUDA events, max over ranks);
          algorithmic FLOPs F(m,n,r) = 4mnr - 2(m+n)r^2 + (4/3)r^3, r = sum(ncols).
e2e     = the same metric through the reference-facing C ABI `dgeqrdm` with PINNED HOST buffers:
          H2D of A, the factorisation, D2H of A/jpvt/tau all inside the timed region.
e2e_pageable = the same call the reference's user makes: `QRDM.QRDM(...)` on PAGEABLE NumPy arrays (the library's
          multi-threaded pinned bounce pipeline + streamed write-back, hostio.c), wall clock.
roofline_hbm = the bandwidth-bound stages (K1 norms, K3b candidate Gram, K3d column exchange, K4 tall panel, K2 norm
          downdate) of configs[3] on one GPU: algorithmic bytes / CUDA-event stage time against
          MEASURED_PEAKS.json's hbm_gbs.
roofline= the trailing update (K6: k_fused / k_vtc + k_tinv + k_wapply + k_rankk, FP64 DMMA) timed with CUDA events
          on the launching stream inside the same timed steps, against the FP64 DMMA peak measured
          live by the library's micro-benchmark (MEASURED_PEAKS.json carries no FP64 figure).
cpu_baseline / --impl reference = the UNMODIFIED reference (oracle/_ref, compiled from
          /root/reference, OpenBLAS on all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner
# to stdout under torchrun), so keep a private copy of the real stdout for the JSON line and point fd 1
# at stderr for everybody else.
_JSON_OUT = None


def protect_stdout():
    """Called by main() only (importing this module must not touch the caller's file descriptors)."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

WORKLOADS = {
    # name: (m, n, generator, stop_mode, description = the BASELINE.json config it is)
    "C1": (1000, 1000, "gaussian", 0, "1000x1000 random Gaussian double matrix (configs[0])"),
    "C2": (4096, 4096, "graded", 1, "4096x4096 graded-spectrum rank-deficient (rank 2048) (configs[1])"),
    "C3": (16384, 16384, "gaussian", 0, "16384x16384 dense Gaussian FP64 (configs[2], trailing-GEMM roofline case)"),
    "C4s": (250000, 512, "gaussian", 0, "tall-skinny 250000x512 slice of configs[3] (one GPU's share at 8 GPUs)"),
}
THRES = (0.9, 0.15)
NB = 64


def flops(m, n, r):
    m, n, r = float(m), float(n), float(r)
    return 4 * m * n * r - 2 * (m + n) * r * r + (4.0 / 3.0) * r ** 3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_matrix_torch(torch, m, n, kind, seed, device):
    """Synthetic input generated on the device (float64, column-major = a (n, m) row-major tensor)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + seed)
    if kind == "gaussian":
        return torch.randn((n, m), dtype=torch.float64, device=device, generator=gen)  # At[c, r] = A[r, c]
    if kind == "graded":
        # test.ipynb cell 3 recipe: X = U diag(sv) V', sv_i = 2^(1-i)+1e-17, sv[:r] += .01 (r-i)
        k = min(m, n)
        r = k // 2
        U, _ = torch.linalg.qr(torch.randn((m, m), dtype=torch.float64, device=device, generator=gen))
        V, _ = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device=device, generator=gen))
        i = torch.arange(1, k + 1, dtype=torch.float64, device=device)
        sv = torch.pow(2.0, 1.0 - i) + 1e-17
        sv[:r] += 0.01 * (r - i[:r])
        X = (U[:, :k] * sv) @ V[:, :k].T          # m x n
        return X.T.contiguous()                    # stored as (n, m) row-major = column-major m x n
    raise ValueError(kind)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref) on the
    box's host cores, on a bounded sample of the workload.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import ref
    m, n, kind, stop_mode, desc = WORKLOADS[args.workload]
    cores = len(os.sched_getaffinity(0))
    line = {"impl": "reference", "metric": "dgeqrdm_fp64_gflops", "unit": "GFLOP/s", "higher_is_better": True,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "dtype": "f64", "data": "synthetic",
            "scaling": "weak", "vs_baseline": None}
    if not ref.have_ref():
        line["unavailable"] = "oracle/_ref/libqrdm_ref.so not present (build it where /root/reference exists)"
        emit(line)
        return
    ref.set_ref_threads(cores)
    # bounded sample: the leading sm x sn block of the same matrix family, sized for a few s per step
    sm_, sn_ = (min(m, args.ref_sample), min(n, args.ref_sample)) if n > 1024 else (m, n)
    if args.workload == "C4s":
        sm_, sn_ = min(m, 100000), n
    from qrdm_b200 import generators as g
    A0 = g.gaussian(sm_, sn_, 0) if kind == "gaussian" else g.graded(sn_, seed=0, m=sm_)
    times, rk = [], 0
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = ref.ref_dgeqrdm(A0, thres=THRES, nb=NB, stop_mode=stop_mode)
        dt = time.perf_counter() - t0
        rk = int(out["ncols"].sum())
        if s >= args.warmup:
            times.append(dt)
        if s == 0 and dt * (args.warmup + args.steps) > 240:  # keep the arm within a few minutes
            times = [dt]
            break
    tot = sum(times)
    val = flops(sm_, sn_, rk) * len(times) / tot / 1e9
    whole = (sm_, sn_) == (m, n)
    sample = (f"reference dgeqrdm (oracle/_ref, unmodified sources, OpenBLAS {cores} threads) on "
              + (f"the FULL {sm_}x{sn_} {kind} matrix of the workload" if whole else
                 f"a {sm_}x{sn_} {kind} matrix (leading-block sample of {desc})")
              + f", rank {rk}, {len(times)} timed run(s)"
              + ("" if len(times) > 1 else " (a single run: warm-up + steps would exceed the few-minute budget)"))
    line.update({"value": val, "ms_per_step": tot / len(times) * 1e3,
                 "config": {"workload": desc, "shape": [m, n], "sample_shape": [sm_, sn_], "same_config": whole,
                            "thres": list(THRES), "nb": NB, "stop_mode": stop_mode},
                 "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "reference",
                                  "sample": sample},
                 "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


def cpu_baseline(args, m, n, kind, stop_mode, desc):
    """Bounded CPU sample on rank 0 at N=1: the unmodified reference where its .so travelled with
    the snapshot, else the plain-C port."""
    from oracle import ref
    from qrdm_b200 import generators as g
    cores = len(os.sched_getaffinity(0))
    if ref.have_ref():
        ref.set_ref_threads(cores)
        sm_, sn_ = (min(m, args.cpu_sample), min(n, args.cpu_sample)) if n > 1024 else (m, n)
        if args.workload == "C4s":
            sm_, sn_ = min(m, 100000), n
        A0 = g.gaussian(sm_, sn_, 0) if kind == "gaussian" else g.graded(sn_, seed=0, m=sm_)
        ref.ref_dgeqrdm(g.gaussian(512, 512, 1))  # warm the BLAS threads
        t0 = time.perf_counter()
        out = ref.ref_dgeqrdm(A0, thres=THRES, nb=NB, stop_mode=stop_mode)
        dt = time.perf_counter() - t0
        rk = int(out["ncols"].sum())
        t1 = time.perf_counter()
        ref.ref_dgeqp3(A0)
        dt3 = time.perf_counter() - t1
        split = None
        if ref.have_ref_timed():   # the same sources with call-site timers (oracle/ref_timing_shim.c)
            try:
                split = ref.ref_dgeqrdm_timed(A0, thres=THRES, nb=NB, stop_mode=stop_mode)["split"]
            except Exception as exc:
                split = {"error": str(exc)[:120]}
        return {"value": flops(sm_, sn_, rk) / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "reference",
                "seconds": dt, "dgeqp3_seconds": dt3, "stage_split_seconds": split,
                "dgeqp3_gflops": flops(sm_, sn_, min(sm_, sn_)) / dt3 / 1e9,
                "sample": f"one run of the unmodified reference dgeqrdm (oracle/_ref, OpenBLAS {cores} threads) on a "
                          f"{sm_}x{sn_} {kind} matrix = leading-block sample of {desc}; rank {rk}; "
                          f"LAPACK dgeqp3 on the same sample timed beside it"}
    sm_ = min(m, 1024)
    sn_ = min(n, 1024)
    A0 = g.gaussian(sm_, sn_, 0)
    t0 = time.perf_counter()
    out = ref.port_dgeqrdm(A0, thres=THRES, nb=NB, stop_mode=stop_mode)
    dt = time.perf_counter() - t0
    rk = int(out["ncols"].sum())
    return {"value": flops(sm_, sn_, rk) / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "seconds": dt, "sample": f"scalar C port (oracle/qrdm_port.c) on a {sm_}x{sn_} Gaussian sample"}


def sharded_parity(torch, dist, qrdm_b200, rank, world, dev, m=200_000, n=512):
    """parity_vs_single_gpu: one seeded m x n Gaussian matrix factored row-sharded over `world`
    GPUs and on ONE GPU (every rank does the single-GPU run itself, it fits); jpvt / ncols must be equal and every
    rank's rows of the factor must agree with the single-GPU rows."""
    from qrdm_b200 import sharded
    gen = torch.Generator(device=dev)
    gen.manual_seed(777)                               # the same matrix on every rank
    A0 = torch.randn((n, m), dtype=torch.float64, device=dev, generator=gen)
    row0, ml = sharded.row_partition(m, world)[rank]
    loc = A0[:, row0:row0 + ml].clone()                # .contiguous() would alias A0 when world == 1
    jp, tau = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    info, nc = sharded.dgeqrdm_sharded(loc, ml, m, row0, world, n, ml, jp, tau, thres=THRES, nb=NB)
    one = A0.clone()
    jp1, tau1 = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    info1, nc1 = qrdm_b200.dgeqrdm_device(one, m, n, m, jp1, tau1, thres=THRES, nb=NB)
    scale = float(torch.linalg.norm(one))
    diff = float(torch.linalg.norm(loc - one[:, row0:row0 + ml])) / scale
    stat = torch.tensor([float(info != 0 or info1 != 0), float(not torch.equal(jp, jp1)),
                         float(not np.array_equal(nc, nc1)), diff,
                         float(torch.max(torch.abs(tau - tau1)))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stat, op=dist.ReduceOp.MAX)
    bad, jd, nd, diff, td = (float(x) for x in stat.tolist())
    del A0, one, loc
    return {"matrix": f"{m}x{n} Gaussian, seed 777, the same on every rank", "info_ok": bad == 0.0,
            "jpvt_equal": jd == 0.0, "ncols_equal": nd == 0.0, "rows_rel_diff": diff, "tau_max_abs_diff": td,
            "revealed_rank": int(nc.sum()), "reduced_over_ranks": "max"}


def run_row_sharded(torch, dist, qrdm_b200, rank, world, dev, m, n, steps, warmup, with_cpu=False):
    """BASELINE config C4: tall-skinny m x n Gaussian, 1-D block-row sharded over `world` GPUs (strong scaling: the
    matrix is fixed, each rank holds m/world rows).  Exchanges in the data path: LL packets over NVLink peer memory
    from inside the panel kernel + a one-kernel LL all-reduce for the small vectors, ncclAllReduce for the rest.
    world == 1 runs the ordinary single-GPU path on the whole matrix."""
    from qrdm_b200 import sharded
    row0, ml = sharded.row_partition(m, world)[rank]
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    A0 = torch.randn((n, ml), dtype=torch.float64, device=dev, generator=gen)
    A = torch.empty_like(A0)
    d_jpvt = torch.zeros(n, dtype=torch.int32, device=dev)
    d_tau = torch.zeros(min(m, n), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    transport = "none"
    if world > 1:
        sharded.init_comm(rank, world, device=dev)
        transport = ("NVLink peer memory (CUDA IPC): LL packets inside k_panel_tall<true> per panel column + one-kernel "
                     "LL all-reduce for vectors <= 96 Ki doubles; ncclAllReduce above that")
        if os.environ.get("QRDM_B200_MG_LEGACY"):
            transport = "ncclAllReduce per panel column (round-1 path, QRDM_B200_MG_LEGACY)"

    def step():
        A.copy_(A0)
        if world > 1:
            info, ncols = sharded.dgeqrdm_sharded(A, ml, m, row0, world, n, ml, d_jpvt, d_tau, thres=THRES, nb=NB,
                                                  stream=stream.cuda_stream)
        else:
            info, ncols = qrdm_b200.dgeqrdm_device(A, m, n, m, d_jpvt, d_tau, thres=THRES, nb=NB,
                                                   stream=stream.cuda_stream)
        if info != 0:
            raise SystemExit(f"row-sharded dgeqrdm failed: info={info}")
        return ncols

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(steps):
        ncols = step()
        launches += qrdm_b200.stats()["launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rk = int(ncols.sum())
    out = {"workload": f"tall-skinny {m}x{n} Gaussian, row-sharded over {world} GPU(s) (configs[3])",
           "n_gpus": world, "scaling": "strong", "ms_per_step": ms / steps, "seconds": ms / steps / 1e3,
           "value": steps * flops(m, n, rk) / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "revealed_rank": rk,
           "rows_per_gpu": ml, "steps": steps, "warmup": warmup, "gpu_launches_per_step": launches / steps,
           "collectives": transport}
    # ---- per-stage split + HBM rooflines of the bandwidth-bound stages (one GPU, every stage event-timed with a sync) ----
    if world == 1:
        try:
            qrdm_b200.set_profile(1)
            step()
            st = qrdm_b200.stats()
            qrdm_b200.set_profile(0)
            out["stage_profile"] = {"ms_total": st["ms_total"], "ms_stage": st["ms_stage"], "stage_bytes": st["stage_bytes"],
                                    "trailing_flops": st["trailing_flops"],
                                    "note": "profile mode 1: event pair + sync around every stage (perturbs the total)"}
        except Exception as exc:
            out["stage_profile"] = {"error": str(exc)[:200]}
    del A0, A
    torch.cuda.empty_cache()
    try:
        out["parity_vs_single_gpu"] = sharded_parity(torch, dist, qrdm_b200, rank, world, dev)
    except SystemExit:
        raise
    except Exception as exc:
        out["parity_vs_single_gpu"] = {"error": str(exc)[:300]}
    torch.cuda.empty_cache()
    if with_cpu and rank == 0:
        try:
            from oracle import ref as oracle
            from qrdm_b200 import generators as g
            cores = len(os.sched_getaffinity(0))
            oracle.set_ref_threads(cores)
            cm = min(m, 500_000)
            Ac = g.gaussian(cm, n, 0)
            t0 = time.perf_counter()
            o = oracle.ref_dgeqrdm(Ac, thres=THRES, nb=NB)
            dt = time.perf_counter() - t0
            rkc = int(o["ncols"].sum())
            out["cpu_reference"] = {"seconds": dt, "gflops": flops(cm, n, rkc) / dt / 1e9, "cores": cores, "shape": [cm, n],
                                    "seconds_extrapolated_to_full_rows": dt * m / cm,
                                    "sample": f"unmodified reference dgeqrdm (oracle/_ref, OpenBLAS {cores} threads) on a {cm}x{n} "
                                              f"Gaussian matrix (a quarter of the rows of configs[3]; cost is linear in m)"}
            del Ac, o
        except Exception as exc:
            out["cpu_reference"] = {"error": str(exc)[:200]}
    return out


def _cpu_one_matrix(args):
    """Pool worker of the C5 CPU baseline: one matrix, one BLAS thread."""
    theta, seed, n, pert = args
    from oracle import ref as oracle
    from qrdm_b200 import generators as g
    oracle.set_ref_threads(1)
    A = g.kahan(n, theta=theta, perturb=pert, seed=seed)
    t0 = time.perf_counter()
    o = oracle.ref_dgeqrdm(A, thres=THRES, nb=NB)
    return time.perf_counter() - t0, int(np.count_nonzero(o["ncols"]))


def run_batched(torch, dist, qrdm_b200, rank, world, dev, total, n, steps, warmup, with_cpu, pure_theta=None, nb=None):
    """BASELINE config C5: `total` Kahan-type n x n matrices with the seeded diagonal perturbation 1e3*eps*(n..1), split
    evenly over `world` GPUs as independent units (no collective); every rank factors its share with ONE launch of the
    one-CTA-per-matrix kernel (dgeqrdm_batched_dev).  pure_theta=None: per-matrix theta in [1.1, 1.3] with the seeded
    perturbation (mixed block sizes, ~55 iterations per matrix); pure_theta=1.25: the survey's unperturbed Kahan matrix,
    511 one-column iterations per matrix."""
    from qrdm_b200 import generators as g
    from qrdm_b200 import sharded
    nb = NB if nb is None else nb
    per = sharded.batch_partition(total, world)[rank][1]
    distinct = 37
    thetas = [pure_theta if pure_theta is not None else 1.1 + 0.2 * b / distinct for b in range(distinct)]
    # the pure-theta batch is the survey's UNPERTURBED Kahan matrix (jpvt = identity, 511 one-column iterations at n = 512)
    pert = 0.0 if pure_theta is not None else 1e3
    base = np.stack([np.ascontiguousarray(g.kahan(n, theta=thetas[b], perturb=pert, seed=1000 * rank + b).T)
                     for b in range(distinct)])
    d_base = torch.from_numpy(base).to(dev)
    idx = torch.arange(per, device=dev) % distinct
    d_a = torch.empty((per, n, n), dtype=torch.float64, device=dev)
    d_jpvt = torch.zeros((per, n), dtype=torch.int32, device=dev)
    d_tau = torch.zeros((per, n), dtype=torch.float64, device=dev)
    d_ncols = torch.zeros((per, n), dtype=torch.int32, device=dev)
    d_infos = torch.zeros((per,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        torch.index_select(d_base, 0, idx, out=d_a)
        d_ncols.zero_()
        rc = qrdm_b200.api.dgeqrdm_batched_device(per, n, n, d_a.data_ptr(), n, n * n, d_jpvt.data_ptr(), d_tau.data_ptr(),
                                                  d_ncols.data_ptr(), d_infos.data_ptr(), thres=THRES, nb=nb,
                                                  stream=stream.cuda_stream)
        if rc != 0:
            raise SystemExit(f"dgeqrdm_batched_dev failed: {rc}")

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    bad = int((d_infos != 0).sum())
    ranks_ = d_ncols[:distinct].sum(dim=1).tolist()
    fl = sum(flops(n, n, int(r)) for r in ranks_) / distinct * per
    iters = float((d_ncols[:distinct] > 0).sum(dim=1).float().mean())
    agg = torch.tensor([ms, fl, bad], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        ms, fl, bad = float(mx[0]), float(agg[1]), int(agg[2])
    out = {"workload": f"{total} Kahan-type {n}x{n} matrices ("
                       + (f"theta = {pure_theta}, unperturbed" if pure_theta is not None else "theta in [1.1, 1.3], perturbed diagonal")
                       + (f", nb = {nb}" if nb != NB else "")
                       + f") split over {world} GPU(s), independent units (configs[4])",
           "n_gpus": world, "scaling": "strong", "matrices_per_gpu": per, "ms_per_step": ms / steps,
           "matrices_per_s": total * steps / (ms * 1e-3), "value": steps * fl / (ms * 1e-3) / 1e9, "unit": "GFLOP/s",
           "mean_iterations_per_matrix": iters, "nonzero_infos": bad, "gpu_launches_per_step": 1,
           "timed_region": "K x (device-side restore of the batch + one k_small launch), CUDA events, max over ranks"}
    # L2/HBM view of the one-CTA-per-matrix kernel: every iteration streams the trailing matrix of every matrix once in
    # and once out (16 * m_r * n_c bytes, SURVEY.md 8d K6 minimum); with 148 x 2 MB in flight the set exceeds L2
    if rank == 0:
        itv = d_ncols[:distinct].cpu().numpy()
        tb = 0.0
        for b in range(distinct):
            jj = 0
            for k_ in itv[b][itv[b] > 0]:
                tb += 16.0 * (n - jj) * (n - jj - int(k_))
                jj += int(k_)
        tb = tb / distinct * total
        out["roofline"] = {"bound": "hbm", "kernel": "k_small (whole factorisation of one matrix per CTA)",
                           "algorithmic_bytes_per_step": tb, "achieved": tb / (ms / steps * 1e-3) / 1e9, "unit": "GB/s",
                           "note": "16*m_r*n_c bytes per iteration per matrix summed over the batch; peak = MEASURED_PEAKS hbm_gbs"}
    if with_cpu and rank == 0 and nb == NB:
        try:
            import multiprocessing as mp
            cores = len(os.sched_getaffinity(0))
            cnt = 256 if pure_theta is None else 64          # ~0.1 s (mixed) / ~0.1-0.2 s (pure) per matrix and thread
            jobs = [(thetas[b % distinct], 1000 * rank + (b % distinct), n, pert) for b in range(cnt)]
            with mp.get_context("spawn").Pool(cores) as pool:
                pool.map(_cpu_one_matrix, jobs[:cores])      # start the workers, load the libraries
                t0 = time.perf_counter()
                res = pool.map(_cpu_one_matrix, jobs, chunksize=1)
                wall = time.perf_counter() - t0
            out["cpu_reference"] = {"matrices": cnt, "pool_processes": cores, "blas_threads_each": 1, "wall_seconds": wall,
                                    "matrices_per_s": cnt / wall, "seconds_for_the_whole_batch_extrapolated": total * wall / cnt,
                                    "mean_ms_per_matrix_1_thread": 1e3 * sum(r[0] for r in res) / cnt,
                                    "mean_iterations": sum(r[1] for r in res) / cnt,
                                    "sample": f"{cnt}-matrix subsample of the batch, unmodified reference (oracle/_ref), process "
                                              f"pool of {cores} x 1 BLAS thread, extrapolated to {total} matrices (SURVEY.md 8d)"}
        except Exception as exc:
            out["cpu_reference"] = {"error": str(exc)[:200]}
    del d_a, d_base
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=6144, help="edge of the CPU-baseline sample block")
    ap.add_argument("--ref-sample", type=int, default=16384,
                    help="edge of the --impl reference sample block (default = the full configs[2] matrix: same config as the GPU arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-row-sharded", action="store_true", help="skip the configs[3] row-sharded leg")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C1/C2 context timings")
    ap.add_argument("--sharded-rows", type=int, default=2_000_000)
    ap.add_argument("--no-batched", action="store_true", help="skip the configs[4] batched leg")
    ap.add_argument("--batch-total", type=int, default=8192)
    args = ap.parse_args()
    protect_stdout()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)  # contract says W >= 3; honour smaller only for ncu captures

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import qrdm_b200  # fails loudly if libqrdm_b200.so is missing: there is no fallback
    from qrdm_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if _lib.lib.qrdm_b200_init(local_rank) != 0:
        raise SystemExit("qrdm_b200_init failed")

    m, n, kind, stop_mode, desc = WORKLOADS[args.workload]
    if os.environ.get("QRDM_BENCH_SHAPE"):  # experiments only: "m,n" Gaussian
        m, n = (int(x) for x in os.environ["QRDM_BENCH_SHAPE"].split(","))
        kind, stop_mode, desc = "gaussian", 0, f"experiment {m}x{n} Gaussian"
    minmn = min(m, n)
    stream = torch.cuda.current_stream()

    # ---- synthetic input, resident in HBM; (n, m) row-major tensor == m x n column-major ----
    A0 = make_matrix_torch(torch, m, n, kind, seed=rank, device=dev)
    lda = m + int(os.environ.get("QRDM_BENCH_LDA_PAD", "0"))  # experiment: non-power-of-two column stride
    if lda != m:
        A0p = torch.zeros((n, lda), dtype=torch.float64, device=dev)
        A0p[:, :m] = A0
        A0 = A0p
    A = torch.empty_like(A0)
    d_jpvt = torch.zeros(n, dtype=torch.int32, device=dev)
    d_tau = torch.zeros(minmn, dtype=torch.float64, device=dev)
    peak_dmma = qrdm_b200.fp64_peak(True, stream.cuda_stream)
    peak_dfma = qrdm_b200.fp64_peak(False, stream.cuda_stream)

    def one_step():
        A.copy_(A0)  # restore the input (D2D, 8mn bytes read + written; inside the timed region)
        info, ncols = qrdm_b200.dgeqrdm_device(A, m, n, lda, d_jpvt, d_tau, thres=THRES, nb=NB,
                                               stop_mode=stop_mode, stream=stream.cuda_stream)
        if info != 0:
            raise SystemExit(f"dgeqrdm_dev failed: info={info}")
        return ncols

    qrdm_b200.set_profile(2)  # light: event pairs around panel + trailing stages, no syncs
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    trailing_ms = panel_ms = trailing_flops = side_flops = side_ms = fused_ms = fused_flops = 0.0
    trailing_launches = side_launches = fused_launches = 0
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        ncols = one_step()
        st = qrdm_b200.stats()
        launches += st["launches"]
        trailing_ms += st["ms_stage"]["trailing"] + st["ms_stage"]["vtv"]   # vtv = the k_fused launches, timed on their own
        fused_ms += st["ms_stage"]["vtv"]
        fused_flops += st["fused_flops"]
        fused_launches += st["fused_launches"]
        panel_ms += st["ms_stage"]["panel"]
        trailing_flops += st["trailing_flops"]
        trailing_launches += st["stage_launches"]["trailing"] + st["stage_launches"]["vtv"]
        side_flops += st["side_flops"]          # look-ahead: pass-2 FLOPs applied by the side stream, outside the stage
        side_ms += st["ms_stage"]["rankk"]
        side_launches += st["side_launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    rk = int(ncols.sum())
    iters = int(np.count_nonzero(ncols))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * flops(m, n, rk) / (ms * 1e-3) / 1e9
    qrdm_b200.set_profile(0)

    # ---- QRCP on the same device: LAPACK dgeqp3's blocked algorithm on the GPU, one run, same matrix ----
    qrcp_gpu = None
    if rank == 0 and not args.no_other_configs:
        try:
            A.copy_(A0)
            torch.cuda.synchronize()
            qinfo = qrdm_b200.dgeqp3_device(A, m, n, lda, d_jpvt, d_tau, stream=stream.cuda_stream)
            qst = qrdm_b200.stats()
            qrcp_gpu = {"what": "qrdm_b200_dgeqp3_dev: blocked Householder QR with classical column pivoting (dlaqps) on the same "
                                "device-resident matrix, CUDA events inside the call, one run",
                        "info": qinfo, "seconds": qst["ms_total"] * 1e-3, "launches": qst["launches"],
                        "gflops": flops(m, n, minmn) / (qst["ms_total"] * 1e-3) / 1e9,
                        "dgeqrdm_speedup_same_device": (qst["ms_total"] / (ms / args.steps)) if ms > 0 else None}
        except Exception as exc:  # noqa: BLE001
            qrcp_gpu = {"error": str(exc)[:200]}

    # ---- end to end through the reference-facing C ABI with pinned host buffers ----
    hA0 = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
    hA0.copy_(A0[:, :m])
    hA = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
    h_jpvt = np.zeros(n, dtype=np.int32)
    h_tau = np.zeros(minmn, dtype=np.float64)
    h_ncols = np.zeros(n, dtype=np.int32)
    th = np.array([THRES[0], THRES[1], 0.0])
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        hA.copy_(hA0)          # host-side restore, NOT timed
        h_jpvt[:] = 0
        h_ncols[:] = 0
        h_ncols[0] = stop_mode
        t0 = time.perf_counter()
        info = _lib.lib.dgeqrdm(102, m, n, hA.data_ptr(), m, h_jpvt.ctypes.data, h_tau.ctypes.data,
                                h_ncols.ctypes.data, th.ctypes.data, NB)
        dt = time.perf_counter() - t0
        if info != 0:
            raise SystemExit(f"dgeqrdm failed: info={info}")
        return dt

    e2e_step()
    if world > 1:
        dist.barrier()
    e2e_t = sum(e2e_step() for _ in range(e2e_steps))
    if world > 1:
        t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    e2e_rank = int(h_ncols.sum())
    e2e_val = world * e2e_steps * flops(m, n, e2e_rank) / e2e_t / 1e9

    # ---- the call a user of the reference makes: QRDM.QRDM on pageable NumPy arrays (QRDM_wrapper.c:89-96) ----
    e2e_pg = None
    try:
        from qrdm_b200 import QRDM as shim
        A_np = np.empty((n, m), dtype=np.float64)          # C-ordered (n, m) buffer = column-major m x n, pageable
        src_np = hA0.numpy()
        th2 = np.array([THRES[0], THRES[1]])                # the notebook passes a 2-element thres

        def pg_step():
            np.copyto(A_np, src_np)                        # host-side restore, NOT timed
            h_jpvt[:] = 0
            h_ncols[:] = 0
            h_ncols[0] = stop_mode
            t0 = time.perf_counter()
            info = shim.QRDM(102, m, n, A_np, m, h_jpvt, h_tau, h_ncols, th2, NB)
            dt = time.perf_counter() - t0
            if info != 0:
                raise SystemExit(f"QRDM.QRDM failed: info={info}")
            return dt

        pg_step()
        if world > 1:
            dist.barrier()
        pg_t = sum(pg_step() for _ in range(e2e_steps))
        if world > 1:
            t = torch.tensor([pg_t], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pg_t = float(t.item())
        same = bool(np.array_equal(A_np, hA.numpy()))      # bit-identical to the pinned-buffer result
        e2e_pg = {"value": world * e2e_steps * flops(m, n, int(h_ncols.sum())) / pg_t / 1e9, "unit": "GFLOP/s",
                  "ms_per_step": pg_t / e2e_steps * 1e3, "steps": e2e_steps, "ratio_to_pinned_time": pg_t / e2e_t,
                  "result_bit_identical_to_pinned": same,
                  "api": "qrdm_b200.QRDM.QRDM (the reference's Python call, QRDM_wrapper.c:71-101) on pageable NumPy arrays: "
                         "multi-threaded pinned bounce H2D, streamed D2H through a pinned ring (hostio.c)"}
        del A_np
    except SystemExit:
        raise
    except Exception as exc:
        e2e_pg = {"error": str(exc)[:300]}
    clocks = sampler.stop() if rank == 0 else None

    # ---- configs[3]: the row-sharded tall-skinny case (every rank takes part) ----
    row_sharded = None
    if not args.no_row_sharded and not os.environ.get("QRDM_BENCH_SHAPE"):
        del A0, A, hA0, hA
        torch.cuda.empty_cache()
        try:
            row_sharded = run_row_sharded(torch, dist, qrdm_b200, rank, world, dev, args.sharded_rows, 512,
                                          steps=3, warmup=2, with_cpu=(world == 1 and not args.no_cpu_baseline))
        except SystemExit:
            raise
        except Exception as exc:  # never sink the headline measurement
            row_sharded = {"error": str(exc)[:300]}

    batched = None
    if not args.no_batched and not os.environ.get("QRDM_BENCH_SHAPE"):
        try:
            batched = run_batched(torch, dist, qrdm_b200, rank, world, dev, args.batch_total, 512, steps=2, warmup=1,
                                  with_cpu=(world == 1 and not args.no_cpu_baseline))
        except SystemExit:
            raise
        except Exception as exc:
            batched = {"error": str(exc)[:300]}
        try:
            pure = run_batched(torch, dist, qrdm_b200, rank, world, dev, args.batch_total, 512, steps=1, warmup=1,
                               with_cpu=(world == 1 and not args.no_cpu_baseline), pure_theta=1.25)
            if isinstance(batched, dict):
                batched["pure_theta_1.25"] = pure
        except SystemExit:
            raise
        except Exception as exc:
            if isinstance(batched, dict):
                batched["pure_theta_1.25"] = {"error": str(exc)[:300]}
        if isinstance(batched, dict) and isinstance(batched.get("pure_theta_1.25"), dict):
            batched["pure_theta_1.25"]["note"] = (
                "unperturbed Kahan: all trailing partial norms are equal to the last bit (ORDER margin 0 in every iteration, "
                "the 1e-12 rule exempts it).  The reference (and the one-matrix GPU path, tests kahan512: 511/511 blocks exact) "
                "break the ties towards 511 one-column iterations; the one-CTA kernel sums the norms in another order and "
                "breaks them differently (see mean_iterations_per_matrix).  The forced one-column case is the nb = 1 batch below.")
        try:
            one = run_batched(torch, dist, qrdm_b200, rank, world, dev, args.batch_total // 8, 512, steps=1, warmup=1,
                              with_cpu=False, pure_theta=1.25, nb=1)
            if isinstance(batched, dict):
                one["note"] = ("nb = 1: every iteration triangularises exactly one column (512 iterations per matrix, the "
                               "latency-bound worst case of the kernel); an eighth of the batch to bound the run time")
                batched["one_column_iterations_nb1"] = one
        except SystemExit:
            raise
        except Exception as exc:
            if isinstance(batched, dict):
                batched["one_column_iterations_nb1"] = {"error": str(exc)[:300]}

    # ---- the other single-GPU BASELINE configs (parity-test cases, reported for context) ----
    other = {}
    if world == 1 and not args.no_other_configs and not os.environ.get("QRDM_BENCH_SHAPE"):
        extra = dict(WORKLOADS)
        extra["G8192"] = (8192, 8192, "gaussian", 0, "8192x8192 dense Gaussian (north_star: K6 >= 60 % of FP64 peak for n >= 8192)")
        for name in ("C1", "C2", "G8192"):
            try:
                om, on, okind, ostop, odesc = extra[name]
                B0 = make_matrix_torch(torch, om, on, okind, seed=0, device=dev)
                B = torch.empty_like(B0)
                bj = torch.zeros(on, dtype=torch.int32, device=dev)
                bt = torch.zeros(min(om, on), dtype=torch.float64, device=dev)
                best = None
                k6 = None
                qrdm_b200.set_profile(2)   # event pairs around the panel and trailing stages only, no syncs
                for _ in range(4):
                    B.copy_(B0)
                    torch.cuda.synchronize()
                    oinfo, oncols = qrdm_b200.dgeqrdm_device(B, om, on, om, bj, bt, thres=THRES, nb=NB, stop_mode=ostop,
                                                             stream=stream.cuda_stream)
                    st = qrdm_b200.stats()
                    if best is None or st["ms_total"] < best:
                        best = st["ms_total"]
                        k6 = (st["trailing_flops"] - st["side_flops"], st["ms_stage"]["trailing"] + st["ms_stage"]["vtv"],
                              st["ms_stage"]["panel"], st["side_flops"], st["fused_flops"], st["ms_stage"]["vtv"])
                qrdm_b200.set_profile(0)
                ork = int(oncols.sum())
                k6_tf = k6[0] / (k6[1] * 1e-3) / 1e12 if k6 and k6[1] > 0 else None
                other[name] = {"workload": odesc, "ms": best, "revealed_rank": ork, "stop_mode": ostop,
                               "gflops": flops(om, on, ork) / (best * 1e-3) / 1e9, "iterations": int(np.count_nonzero(oncols)),
                               "roofline": {"kernel": "K6 trailing update", "bound": "tensor", "achieved": k6_tf, "peak": peak_dmma,
                                            "unit": "TFLOP/s", "frac": (k6_tf / peak_dmma) if k6_tf else None,
                                            "ms": k6[1] if k6 else None, "panel_ms": k6[2] if k6 else None,
                                            "flops_in_stage": k6[0] if k6 else None,
                                            "k_fused_alone": ({"tflops": k6[4] / (k6[5] * 1e-3) / 1e12, "frac": k6[4] / (k6[5] * 1e-3) / 1e12 / peak_dmma,
                                                               "ms": k6[5]} if k6 and k6[5] > 0 else None),
                                            "flops_on_side_stream": k6[3] if k6 else None,
                                            "whole_step_frac_of_peak": flops(om, on, ork) / (best * 1e-3) / 1e12 / peak_dmma},
                               "timing": "best of 3 after 1 warm-up, device-resident (CUDA events inside dgeqrdm_dev)"}
                del B0, B
            except Exception as exc:
                other[name] = {"error": str(exc)[:200]}
        qrdm_b200.set_profile(0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # Dominant kernel: k_fused (k_vtc for the first deferred block), its own CUDA-event pairs inside the timed steps and
    # the FLOPs executed inside those launches.  The whole K6 stage (+ k_tinv, k_wapply, the eager k_rankk<list>, and
    # k_vtc + k_rankk below the break-even) is reported beside it; FLOPs the look-ahead moved to the side stream are
    # executed outside the stage and do not count for either.
    achieved = fused_flops / (fused_ms * 1e-3) / 1e12 if fused_ms > 0 else None
    stage_achieved = (trailing_flops - side_flops) / (trailing_ms * 1e-3) / 1e12 if trailing_ms > 0 else None
    line = {
        "metric": "dgeqrdm_fp64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "shape": [m, n], "thres": list(THRES), "nb": NB, "stop_mode": stop_mode,
                   "revealed_rank": rk, "iterations": iters, "parallelism": f"replicas x{world}" if world > 1 else "1 GPU",
                   "l2": "input 8mn bytes >> 126 MB L2 and restored from HBM every step (no flush needed)"
                         if 8 * m * n > 4 * 126e6 else "input restored by a D2D copy every step; fits L2 partially",
                   "timed_region": "K x (D2D restore of A + dgeqrdm_dev), CUDA events on the launching stream"},
        "e2e": {"value": e2e_val, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * m * n + 8 * minmn,
                "d2h_bytes_per_step": 8 * m * n + 8 * minmn + 4 * n, "ms_per_step": e2e_t / e2e_steps * 1e3,
                "steps": e2e_steps, "api": "dgeqrdm (C ABI, pinned host buffers, wall clock around the blocking call)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "k_fused (K6, DMMA.8x8x4): pass 2 of the pending block fused into pass 1 of the current one; pass 1 only "
                               "on the columns the look-ahead's side stream completed", "bound": "tensor",
                     "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s",
                     "frac": (achieved / peak_dmma) if achieved else None,
                     "peak_source": "FP64 DMMA peak measured live by qrdm_b200_measure_fp64_peak "
                                    f"(DFMA pipe: {peak_dfma:.2f}); MEASURED_PEAKS.json has no FP64 entry "
                                    f"(hbm_gbs={peaks_file.get('hbm_gbs')})",
                     "algorithmic_flops_per_step": fused_flops / args.steps,
                     "launches_per_step": fused_launches / args.steps,
                     "avg_launch_ms": fused_ms / fused_launches if fused_launches else None,
                     "k6_stage": {"what": "the whole trailing-update stage on the main stream: k_fused + k_tinv + k_wapply + eager "
                                          "k_rankk<list>, and k_vtc + k_rankk below the 2048^2 break-even",
                                  "achieved": stage_achieved, "frac": (stage_achieved / peak_dmma) if stage_achieved else None,
                                  "algorithmic_flops_per_step": (trailing_flops - side_flops) / args.steps,
                                  "launches_per_step": trailing_launches / args.steps, "ms_per_step": trailing_ms / args.steps},
                     "lookahead": {"what": "pass 2 of the pending block on the last columns, applied by k_rankk (side mode) on a "
                                           "least-priority stream beside the next selection / panel (SURVEY 8f-2); its FLOPs are "
                                           "excluded from `achieved` above, its event time (waits for SMs included) is reported here",
                                   "flops_per_step": side_flops / args.steps, "launches_per_step": side_launches / args.steps,
                                   "side_stream_ms_per_step": side_ms / args.steps,
                                   "share_of_k6_flops": side_flops / trailing_flops if trailing_flops > 0 else None},
                     "ms_per_step": fused_ms / args.steps, "share_of_step": fused_ms / ms,
                     "traffic": 2.075419e9 + 1.768631e9,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, ncu --set full at iteration 10 (m_r = 15744, "
                                     "15680 trailing columns): 2.075 GB read + 1.769 GB written; algorithmic for that launch: 1.975 GB "
                                     "read (every trailing element once) + ~1.77 GB written (the columns the side stream completed are "
                                     "read for pass 1 but not written): ratio 1.03, no wasted re-reads; traffic shrinks with the trailing "
                                     "matrix, `achieved` is the average over all launches of a step; profiles/r02_ncu_lookahead_and_grouped_panel.txt"},
        "stages": {"panel_ms_per_step": panel_ms / args.steps, "trailing_ms_per_step": trailing_ms / args.steps},
    }
    if e2e_pg is not None:
        line["e2e_pageable"] = e2e_pg
    if qrcp_gpu is not None:
        line["qrcp_same_device"] = qrcp_gpu
    # ---- HBM rooflines of the bandwidth-bound stages, measured on configs[3] (one GPU) ----
    if row_sharded is not None and isinstance(row_sharded.get("stage_profile"), dict) and "ms_stage" in row_sharded["stage_profile"]:
        sp = row_sharded["stage_profile"]
        hbm_peak = float(peaks_file.get("hbm_gbs") or 6550.0)
        copy_live = qrdm_b200.copy_gbs(1 << 30, stream.cuda_stream)
        names = {"norm_init": "K1 initial column norms (k_colnorm_partial/finalize), 8mn bytes",
                 "gram": "K3b candidate Gram, 8*m_r*nc bytes per iteration",
                 "permute": "K3d column exchange (k_permute), 32*m bytes per exchange",
                 "panel": "K4 tall panel (k_panel_tall + k_skinny), >= 16*m_r*fjb bytes per iteration",
                 "norm_update": "K2 norm downdate, 8*k*n_r bytes per iteration",
                 "trailing": "K6 trailing update at n = 512 (24*m_r*n_c bytes per iteration; compute-bound, for reference)"}
        rl = []
        for key, label in names.items():
            ms_ = sp["ms_stage"].get(key, 0.0)
            by_ = sp["stage_bytes"].get(key, 0.0)
            if ms_ > 0 and by_ > 0:
                gbs = by_ / (ms_ * 1e-3) / 1e9
                rl.append({"stage": key, "kernel": label, "bound": "hbm", "algorithmic_bytes": by_, "ms": ms_,
                           "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak})
        line["roofline_hbm"] = {"workload": row_sharded["workload"], "peak_source": "MEASURED_PEAKS.json hbm_gbs"
                                if peaks_file.get("hbm_gbs") else "fallback 6550 GB/s (B200_PROFILING.md)",
                                "copy_gbs_measured_live": copy_live, "stages": rl,
                                "timing": "CUDA event pair + sync around every stage of one factorisation (profile mode 1)"}
    if row_sharded is not None:
        line["row_sharded"] = row_sharded
    if batched is not None:
        line["batched"] = batched
    if other:
        line["other_configs"] = other
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args, m, n, kind, stop_mode, desc)
        except Exception as exc:  # the baseline must never sink the measurement
            line["cpu_baseline"] = {"value": None, "error": str(exc)[:200]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
