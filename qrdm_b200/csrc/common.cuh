// common.cuh — shared device helpers for the qrdm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "qrdm_dev.h"

extern long long g_qrdm_launches;  // kernels launched by this process (stats: gpu_launches)
extern "C" int qrdm_rt_device_generation(void);  // bumped when the library (re)initialises on a device

#define QRDM_LAUNCH_CHECK()                   \
  do {                                        \
    ++g_qrdm_launches;                        \
    cudaError_t e__ = cudaGetLastError();     \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

// first local row that is still active when j columns are done (row-sharded: global row j)
__device__ __forceinline__ int qrdm_jr(const qrdm_prob& P, int j) {
  const int x = j - P.row0;
  return x < 0 ? 0 : (x > P.m ? P.m : x);
}

// Geometry the kernels work on: normally the iteration's panel (j, fjb, fjb_cmp) with the trailing
// matrix to its right; in the blocked tall-panel mode (P.sub = s + 1) the 8-column sub-panel that
// starts at panel column s, whose "trailing matrix" is the rest of the panel.
struct QrdmGeom {
  int j;      // first column (= first global row) of the block being applied
  int fjb;    // columns of that block (the update starts right after them)
  int k;      // reflectors in the block
  int n_end;  // one past the last column the update touches
  int voff;   // column of Vc that holds reflector 0 of the block
};
__device__ __forceinline__ QrdmGeom qrdm_geom(const qrdm_prob& P) {
  const qrdm_ctrl* c = P.ctrl;
  QrdmGeom g;
  if (P.pend) {  // flush of a deferred block: rows >= pend_r0 of the columns >= pend_c0
    g.j = c->pend_r0; g.fjb = c->pend_c0 - c->pend_r0; g.k = c->pend_k; g.n_end = P.n; g.voff = 0;
  } else if (P.sub == 0) {
    g.j = c->j; g.fjb = c->fjb; g.k = c->fjb_cmp; g.n_end = P.n; g.voff = 0;
  } else {
    const int s = P.sub - 1;
    // After a DM early stop only the sub-panels BEHIND the one that stopped are dead.  The stopping sub-panel itself
    // still owes its k = sub_k reflectors to the rest of the panel (the reference applies every reflector to all
    // remaining panel columns as it goes, src/dgeqr2.c:179-186) — round 1 marked it dead as well, so a stop in a
    // sub-panel s > 0 left the panel columns behind it one block reflector short.
    const bool dead = s >= c->fjb || (s > 0 && c->tall_done && s > c->tall_stop_s);
    g.j = c->j + s;
    g.fjb = dead ? 0 : min(QRDM_TALL_B, c->fjb - s);
    g.k = dead ? 0 : c->sub_k;
    g.n_end = c->j + c->fjb;
    g.voff = s;
  }
  return g;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// Deterministic block-wide sum; every thread gets the result. scratch: >= 32 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += scratch[i];
  return t;
}

// cp.async helpers (LDGSTS).  16-byte copies with a source size that may be 0..16: the remainder
// is zero-filled, which is how row/column tails and masked tiles are handled.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  Lane l: g = l>>2, t = l&3 holds
// A[g][t], B[t][g], D[g][2t], D[g][2t+1].  SASS: DMMA.8x8x4 — measured at the full 37.0 TFLOP/s
// FP64 peak on B200 (profiles/r01_fp64_peak_microbench.txt) vs 33.5-34 for DFMA loops.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- mbarrier + bulk async copy (cp.async.bulk: SASS UBLKCP, the TMA engine's linear mode) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // a copy that never lands must not hang the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
