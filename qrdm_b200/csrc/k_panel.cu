// k_panel.cu — K4: Householder panel factorisation with the Deviation-Maximisation early stop.
//
// Replaces: dgeqr2_mia (reference src/dgeqr2.c:29-195), dlarfg_mia + d_sign
// (src/dlarfg.c:21-188) and dlarf_ (src/dlarf.c:28-211) as called at src/dgeqrdm_work.c:735-738.
//
// One cooperative grid, rows of the m_r x fjb panel split over CTAs; each CTA keeps its row slab
// in shared memory for the whole panel (global-memory mode when the slab does not fit: tall
// matrices).  Per column ONE grid-wide reduction, fused: the sweep that applies H_i also
// accumulates, for the next column x' = P[i+2:, i+1], the dot products x'.P[:, j] for all
// remaining j (j = i+1 gives ||x'||^2), so  w_j = P[i+1, j] + (x'.P_j)/(alpha - beta)  needs no
// second pass.  Partials are combined in fixed CTA order by every CTA, so all
// CTAs take bit-identical decisions (stop test, tau) and the result is run-to-run deterministic.
// Also emits the clean copy Vc of the reflectors (unit diagonal, zeros above, zero-padded to a
// multiple of 8 columns) that the trailing-update kernels consume.
#include <cstdlib>

#include "common.cuh"

#define PANEL_THREADS 512
#define PANEL_WARPS (PANEL_THREADS / 32)
#define PANEL_CPW ((63 + PANEL_WARPS - 1) / PANEL_WARPS)  // columns per warp in the sweep
#define PANEL_RG (PANEL_THREADS / 64)                      // reduction groups

// Grid-wide barrier on a monotonically increasing counter (reset by k_pick).  Release/acquire at
// gpu scope instead of __threadfence(): no L1 invalidation (CCTL.IVALL) on the critical path; data
// written by other CTAs is read with ld.global.cg (L2), see the reduction below.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target, unsigned nctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nctas;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(bar) : "memory");
    while (ld_acquire_u32(bar) < target) { }
  }
  __syncthreads();
}

template <bool SMEM>
__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel(qrdm_prob P, int rpc) {
  extern __shared__ __align__(16) double slab[];  // SMEM mode: [64][rpc]
  __shared__ double sred[PANEL_RG][64];
  __shared__ double S_[64], rowv[64], wv[64];
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, fjb = ctrl->fjb;
  if (fjb <= 0) return;
  const int rows = P.m - j, lda = P.lda;
  const int G = gridDim.x, b = blockIdx.x;
  const int r0 = min(rows, b * rpc), r1 = min(rows, r0 + rpc), nr = r1 - r0;
  double* Ap = P.a + (size_t)j * lda + j;
  const int lds = rpc;
  double* part = P.panel_part;  // [2][PANEL_MAXCTA][64]
  double* rowb = P.panel_row;   // [2][64]
  unsigned bar_target = 0;

#define PX(r, c) (SMEM ? slab[(c) * lds + (r)] : Ap[(size_t)(c) * lda + (size_t)(r0 + (r))])

  if (SMEM) {
    for (int c = wid; c < fjb; c += PANEL_WARPS)
      for (int r = lane; r < nr; r += 32) slab[c * lds + r] = Ap[(size_t)c * lda + r0 + r];
    __syncthreads();
  }

  // partial sums for column 0: S_j = sum_{R>0} P[R,0] P[R,j]; row 0 itself goes to rowb[0]
  for (int jj = wid; jj < fjb; jj += PANEL_WARPS) {
    double acc = 0.0;
    for (int r = lane; r < nr; r += 32) {
      const int R = r0 + r;
      const double p = PX(r, jj);
      if (R > 0) acc = fma(PX(r, 0), p, acc);
      else rowb[jj] = p;
    }
    acc = warp_sum(acc);
    if (lane == 0) part[(size_t)b * 64 + jj] = acc;
  }
  grid_barrier(&ctrl->panel_bar, bar_target, G);

  double thres = 5e-14;  // reference src/dgeqr2.c:40
  int k = fjb;
  for (int i = 0; i < fjb; ++i) {
    const int cur = i & 1, nxt = cur ^ 1;
    // ---- fixed-order reduction of the partials (identical on every CTA => identical decisions) ----
    {
      const int jj = tid & 63, q = tid >> 6;
      double s = 0.0;
      if (jj >= i && jj < fjb) {
        const double* src = part + ((size_t)cur * QRDM_PANEL_MAXCTA) * 64 + jj;
        for (int bb0 = q; bb0 < G; bb0 += 8 * PANEL_RG) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int bb = bb0 + u * PANEL_RG;
            v[u] = bb < G ? __ldcg(src + (size_t)bb * 64) : 0.0;  // L2: written by other SMs
          }
          s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
      }
      sred[q][jj] = s;
    }
    __syncthreads();
    if (tid < 64) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < PANEL_RG; ++q) s += sred[q][tid];
      S_[tid] = s;
      rowv[tid] = __ldcg(&rowb[cur * 64 + tid]);
    }
    __syncthreads();
    // ---- reflector scalars: dlarfg_mia (src/dlarfg.c:120-185), redundantly on every thread ----
    const double alpha = rowv[i];
    const int len = rows - i;
    double tau = 0.0, beta = alpha, scale = 1.0;
    if (len > 1) {
      const double xnorm = sqrt(S_[i]);
      if (i > 0 && xnorm < thres) { k = i; break; }  // DM early stop: column i left untouched
      if (xnorm != 0.0) {
        const double h = hypot(alpha, xnorm);
        beta = (alpha >= 0.0) ? -h : h;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
    }
    if (i == 0 && fjb > 1 && P.tau_ > 0.0) thres = P.tau_ * fabs(beta);  // src/dgeqr2.c:176-177
    if (b == 0 && tid == 0) {
      P.tau[j + i] = tau;
      if (tau != tau && ctrl->err == 0) ctrl->err = -8;  // LAPACKE_dlarft's NaN screen of tau
    }
    const bool last = i + 1 >= fjb;
    if (tid < 64 && tid > i && tid < fjb) wv[tid] = tau * (rowv[tid] + S_[tid] * scale);
    const double w1 = last ? 0.0 : tau * (rowv[i + 1] + S_[i + 1] * scale);
    // ---- phase 1: v = x * scale, diagonal = beta; column i+1 gets H_i right away ----
    for (int r = tid; r < nr; r += PANEL_THREADS) {
      const int R = r0 + r;
      if (R < i) continue;
      double v = 1.0;
      if (R > i) {
        v = PX(r, i);
        if (tau != 0.0) { v *= scale; PX(r, i) = v; }
      } else {
        PX(r, i) = beta;
      }
      if (!last) {
        const double p = fma(-v, w1, PX(r, i + 1));
        PX(r, i + 1) = p;
        if (R == i + 1) rowb[nxt * 64 + i + 1] = p;
      }
    }
    __syncthreads();
    if (last) break;
    // ---- phase 2: remaining columns + the fused dot products for the next reflector ----
    {
      double wreg[PANEL_CPW], acc[PANEL_CPW];
#pragma unroll
      for (int c = 0; c < PANEL_CPW; ++c) {
        const int jj = i + 1 + wid + c * PANEL_WARPS;
        wreg[c] = jj < fjb ? wv[jj] : 0.0;
        acc[c] = 0.0;
      }
      for (int r = lane; r < nr; r += 32) {
        const int R = r0 + r;
        if (R < i) continue;
        const double v = (R == i) ? 1.0 : PX(r, i);
        const double x1 = PX(r, i + 1);
        const bool below = R > i + 1;
#pragma unroll
        for (int c = 0; c < PANEL_CPW; ++c) {
          const int jj = i + 1 + wid + c * PANEL_WARPS;
          if (jj < fjb) {
            double p;
            if (jj == i + 1) {
              p = x1;  // already updated in phase 1: only its squared norm is needed
            } else {
              p = fma(-v, wreg[c], PX(r, jj));
              PX(r, jj) = p;
              if (R == i + 1) rowb[nxt * 64 + jj] = p;
            }
            if (below) acc[c] = fma(x1, p, acc[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < PANEL_CPW; ++c) {
        const int jj = i + 1 + wid + c * PANEL_WARPS;
        const double a = warp_sum(acc[c]);
        if (lane == 0 && jj < fjb) part[((size_t)nxt * QRDM_PANEL_MAXCTA + b) * 64 + jj] = a;
      }
    }
    grid_barrier(&ctrl->panel_bar, bar_target, G);
  }
  __syncthreads();
  if (b == 0 && tid == 0) ctrl->fjb_cmp = k;

  // ---- write the slab back and emit Vc ----
  if (SMEM) {
    for (int c = wid; c < fjb; c += PANEL_WARPS)
      for (int r = lane; r < nr; r += 32) Ap[(size_t)c * lda + r0 + r] = slab[c * lds + r];
  }
  const int kpad = (k + 7) & ~7;
  for (int q = wid; q < kpad; q += PANEL_WARPS) {
    double* vcol = P.vc + (size_t)q * P.ldv + j;
    for (int r = lane; r < nr; r += 32) {
      const int R = r0 + r;
      double v = 0.0;
      if (q < k) v = (R > q) ? PX(r, q) : (R == q ? 1.0 : 0.0);
      vcol[R] = v;
    }
    if (b == 0) {  // rows between the aligned tile start and j must read as zero
      const int jal = j & ~(QRDM_ROWALIGN - 1);
      for (int g = jal + lane; g < j; g += 32) P.vc[(size_t)q * P.ldv + g] = 0.0;
    }
  }
#undef PX
}

extern "C" int qrdm_k_panel(const qrdm_prob* p, int j_host, void* stream) {
  static bool attr_set = false;
  static int rows_per_cta = 128;
  const int smem_cap = 200 * 1024;
  if (!attr_set) {
    cudaFuncSetAttribute(k_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
    const char* e = getenv("QRDM_PANEL_ROWS");
    if (e && atoi(e) >= 32) rows_per_cta = atoi(e);
    attr_set = true;
  }
  const int rows = p->m - j_host;
  if (rows <= 0) return 0;
  // few, fat CTAs: the per-column cost is the grid barrier + the all-to-all read of the partials,
  // both proportional to the CTA count; grow the grid only when the slab would not fit in smem
  const int gmax = p->sm_count < QRDM_PANEL_MAXCTA ? p->sm_count : QRDM_PANEL_MAXCTA;
  int G = (rows + rows_per_cta - 1) / rows_per_cta;
  const int gfit = (rows + (smem_cap / 512) - 1) / (smem_cap / 512);  // CTAs needed for smem residency
  if (G < gfit) G = gfit;
  if (G > gmax) G = gmax;
  if (G < 1) G = 1;
  int rpc = (rows + G - 1) / G;
  qrdm_prob prob = *p;
  void* args[] = {(void*)&prob, (void*)&rpc};
  const size_t smem = (size_t)rpc * 64 * sizeof(double);
  cudaError_t e;
  if (smem <= (size_t)smem_cap)
    e = cudaLaunchCooperativeKernel((void*)k_panel<true>, dim3(G), dim3(PANEL_THREADS), args, smem, (cudaStream_t)stream);
  else
    e = cudaLaunchCooperativeKernel((void*)k_panel<false>, dim3(G), dim3(PANEL_THREADS), args, 0, (cudaStream_t)stream);
  ++g_qrdm_launches;
  if (e != cudaSuccess) return (int)e;
  return 0;
}
