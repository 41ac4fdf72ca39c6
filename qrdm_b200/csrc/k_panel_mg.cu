// k_panel_mg.cu — K4 for the 1-D block-row sharded build: the Householder panel
// when the rows of the panel live on several GPUs.
//
// Same algorithm and same fused reduction as k_panel.cu (one vector [||x||^2, x'C_panel, pivot row]
// per column), but the per-column exchange crosses GPUs, so every column is one kernel followed by
// one ncclAllReduce of 128 doubles issued by the host driver on the same stream:
//     init  -> allreduce -> step 0 -> allreduce -> step 1 -> ... -> finish
// All ranks see identical reduced vectors, hence take identical decisions (stop test, tau); the
// host launches the steps blindly, steps after the early stop (or beyond fjb) return immediately.
// Inside a rank the CTAs' partial vectors are combined by the last CTA to arrive, in CTA order
// (deterministic).  The slab stays in global memory (it does not fit on chip for tall matrices;
// 250000 x 64 doubles = 128 MB is L2-sized).
#include "common.cuh"

#define MG_THREADS 256
#define MG_WARPS (MG_THREADS / 32)

struct MgGeom {
  int j, jr, fjb, rows, pr0, rpc, r0, nr;  // pr0: panel-relative global index of local row jr
  int sub_s, voff, jmain, fjb_main;        // blocked tall-panel mode: sub-panel at panel column sub_s
};
__device__ __forceinline__ MgGeom mg_geom(const qrdm_prob& P) {
  MgGeom g;
  const QrdmGeom q = qrdm_geom(P);
  g.j = q.j; g.fjb = q.fjb; g.voff = q.voff;
  g.sub_s = P.sub ? P.sub - 1 : 0;
  g.jmain = P.ctrl->j; g.fjb_main = P.ctrl->fjb;
  g.jr = qrdm_jr(P, g.j);
  g.rows = P.m - g.jr;                 // local active rows
  g.pr0 = P.row0 + g.jr - g.j;         // >= 0
  g.rpc = (g.rows + gridDim.x - 1) / gridDim.x;
  g.r0 = min(g.rows, (int)blockIdx.x * g.rpc);
  g.nr = min(g.rows, g.r0 + g.rpc) - g.r0;
  return g;
}

// combine the per-CTA vectors part[b][128] into dst[128]: done by the last CTA to arrive, b ascending
__device__ __forceinline__ void mg_combine(const qrdm_prob& P, double* dst, int lo, int hi) {
  __shared__ bool is_last;
  double* part = P.mg_buf + 512;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(P.mg_cnt, 1u);
    is_last = (prev == gridDim.x - 1);
    if (is_last) *P.mg_cnt = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < 128; e += MG_THREADS) {
    double s = 0.0;
    const int jj = e & 63;
    if (jj >= lo && jj < hi)
      for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(&part[(size_t)b * 128 + e]);
    dst[e] = s;
  }
}

#define PA(r, c) Ap[(size_t)(c) * lda + (size_t)(g.r0 + (r))]

__global__ void __launch_bounds__(MG_THREADS) k_panel_mg_init(qrdm_prob P) {
  const MgGeom g = mg_geom(P);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lda = P.lda;
  if (g.fjb <= 0) return;
  if (blockIdx.x == 0 && tid == 0) { P.ctrl->mg_k = -1; if (g.sub_s == 0) P.ctrl->mg_thres2 = P.thres0 * P.thres0; }
  double* Ap = P.a + (size_t)g.j * lda + g.jr;
  double* mine = P.mg_buf + 512 + (size_t)blockIdx.x * 128;
  for (int e = tid; e < 128; e += MG_THREADS) mine[e] = 0.0;
  __syncthreads();
  for (int jj = wid; jj < g.fjb; jj += MG_WARPS) {
    double acc = 0.0;
    for (int r = lane; r < g.nr; r += 32) {
      const int R = g.pr0 + g.r0 + r;
      const double p = PA(r, jj);
      if (R > 0) acc = fma(PA(r, 0), p, acc);
      else mine[64 + jj] = p;
    }
    acc = warp_sum(acc);
    if (lane == 0) mine[jj] = acc;
  }
  mg_combine(P, P.mg_buf, 0, g.fjb);
}

__global__ void __launch_bounds__(MG_THREADS) k_panel_mg_step(qrdm_prob P, int i) {
  __shared__ double wv[64];
  qrdm_ctrl* ctrl = P.ctrl;
  const MgGeom g = mg_geom(P);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lda = P.lda;
  if (i >= g.fjb || ctrl->mg_k >= 0) return;
  const double* red = P.mg_buf + (size_t)(i & 1) * 128;   // all-reduced: [0,64) sums, [64,128) pivot row
  double* Ap = P.a + (size_t)g.j * lda + g.jr;
  double* mine = P.mg_buf + 512 + (size_t)blockIdx.x * 128;
  // ---- reflector scalars (dlarfg_mia, src/dlarfg.c:120-185) ----
  const double alpha = red[64 + i], xn2 = red[i];
  const int len = P.m_glob - g.j - i;
  double thres2 = ctrl->mg_thres2;
  double tau = 0.0, beta = alpha, scale = 1.0;
  if (len > 1) {
    if (g.sub_s + i > 0 && xn2 < thres2) {  // DM early stop: column i left untouched (every rank, every CTA alike)
      if (blockIdx.x == 0 && tid == 0) ctrl->mg_k = i;
      return;
    }
    if (xn2 != 0.0) {
      const double h = sqrt(fma(alpha, alpha, xn2));
      beta = (alpha >= 0.0) ? -h : h;
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    if (g.sub_s + i == 0 && g.fjb_main > 1 && P.tau_ > 0.0) { const double th = P.tau_ * fabs(beta); ctrl->mg_thres2 = th * th; }
    P.tau[g.j + i] = tau;
    if (tau != tau && ctrl->err == 0) ctrl->err = -8;
  }
  const bool last = i + 1 >= g.fjb;
  if (tid < 64 && tid > i && tid < g.fjb) wv[tid] = tau * (red[64 + tid] + red[tid] * scale);
  for (int e = tid; e < 128; e += MG_THREADS) mine[e] = 0.0;
  __syncthreads();
  // ---- phase 1: v = x * scale, diagonal = beta; the next pivot column gets H_i right away ----
  const double w1 = last ? 0.0 : wv[i + 1];
  for (int r = tid; r < g.nr; r += MG_THREADS) {
    const int R = g.pr0 + g.r0 + r;
    if (R < i) continue;
    double v = 1.0;
    if (R > i) {
      v = PA(r, i);
      if (tau != 0.0) { v *= scale; PA(r, i) = v; }
    } else {
      PA(r, i) = beta;
    }
    if (!last) {
      const double p = fma(-v, w1, PA(r, i + 1));
      PA(r, i + 1) = p;
      if (R == i + 1) mine[64 + i + 1] = p;
    }
  }
  __syncthreads();
  if (last) return;
  // ---- phase 2: remaining columns + fused dot products for the next reflector ----
  for (int jj = i + 1 + wid; jj < g.fjb; jj += MG_WARPS) {
    double acc = 0.0;
    if (jj == i + 1) {
      for (int r = lane; r < g.nr; r += 32)
        if (g.pr0 + g.r0 + r > i + 1) { const double x = PA(r, i + 1); acc = fma(x, x, acc); }
    } else {
      const double wj = wv[jj];
      for (int r = lane; r < g.nr; r += 32) {
        const int R = g.pr0 + g.r0 + r;
        if (R < i) continue;
        const double v = (R == i) ? 1.0 : PA(r, i);
        const double p = fma(-v, wj, PA(r, jj));
        PA(r, jj) = p;
        if (R > i + 1) acc = fma(PA(r, i + 1), p, acc);
        else if (R == i + 1) mine[64 + jj] = p;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) mine[jj] = acc;
  }
  mg_combine(P, P.mg_buf + (size_t)((i + 1) & 1) * 128, i + 1, g.fjb);
}

__global__ void __launch_bounds__(MG_THREADS) k_panel_mg_finish(qrdm_prob P) {
  qrdm_ctrl* ctrl = P.ctrl;
  const MgGeom g = mg_geom(P);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lda = P.lda;
  if (g.fjb <= 0) return;
  const int k = ctrl->mg_k >= 0 ? ctrl->mg_k : g.fjb;
  if (blockIdx.x == 0 && tid == 0) {  // read by later kernels only
    if (P.sub) {
      const int tk = (g.sub_s == 0 ? 0 : ctrl->tall_k) + k;
      ctrl->sub_k = k;
      ctrl->tall_k = tk;
      ctrl->tall_done = (k < g.fjb) ? 1 : 0;
      if (k < g.fjb) ctrl->tall_stop_s = g.sub_s;
      ctrl->fjb_cmp = tk;
    } else {
      ctrl->fjb_cmp = k;
    }
  }
  const double* Ap = P.a + (size_t)g.j * lda + g.jr;
  const int kpad = P.sub ? min(QRDM_TALL_B, 64 - g.voff) : ((k + 7) & ~7);
  const int jal = qrdm_jr(P, g.jmain) & ~(QRDM_ROWALIGN - 1);
  for (int q = wid; q < kpad; q += MG_WARPS) {
    double* vcol = P.vc + (size_t)(g.voff + q) * P.ldv + g.jr;
    for (int r = lane; r < g.nr; r += 32) {
      const int R = g.pr0 + g.r0 + r;
      double v = 0.0;
      if (q < k) v = (R > q) ? PA(r, q) : (R == q ? 1.0 : 0.0);
      vcol[g.r0 + r] = v;
    }
    if (blockIdx.x == 0)
      for (int gg = jal + lane; gg < g.jr; gg += 32) P.vc[(size_t)(g.voff + q) * P.ldv + gg] = 0.0;
  }
}
#undef PA

static int mg_grid(const qrdm_prob* p, int j_host) {
  if (p->sub) j_host += p->sub - 1;
  int jr = j_host - p->row0;
  jr = jr < 0 ? 0 : (jr > p->m ? p->m : jr);
  const int rows = p->m - jr;
  int G = (rows + 511) / 512;
  if (G > 2 * p->sm_count) G = 2 * p->sm_count;
  if (G < 1) G = 1;  // a rank without rows still runs one CTA: it must contribute zeros
  return G;
}
extern "C" int qrdm_k_panel_mg_init(const qrdm_prob* p, int j_host, void* stream) {
  k_panel_mg_init<<<mg_grid(p, j_host), MG_THREADS, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
extern "C" int qrdm_k_panel_mg_step(const qrdm_prob* p, int j_host, int step, void* stream) {
  k_panel_mg_step<<<mg_grid(p, j_host), MG_THREADS, 0, (cudaStream_t)stream>>>(*p, step);
  QRDM_LAUNCH_CHECK();
  return 0;
}
extern "C" int qrdm_k_panel_mg_finish(const qrdm_prob* p, int j_host, void* stream) {
  k_panel_mg_finish<<<mg_grid(p, j_host), MG_THREADS, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
