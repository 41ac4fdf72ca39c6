// k_wide.cu — block sizes above 64 (64 < nb <= 256).
//
// The reference accepts any nb > 0 (src/dgeqrdm_work.c:573-575; workspaces sized from nb at :650-665): up to nb
// candidates by norm, their nc x nc cosine matrix, a greedy pick of up to nb columns, ONE panel of fjb <= nb columns and
// ONE compact-WY update.  The kernels of this library are built around 64-column tiles, so a wide block runs as
//   selection at full width   k_select (unchanged: the candidate arrays of qrdm_ctrl hold 256 entries)
//                             k_gram_wide   nc x nc Gram in 64 x 64 blocks of column-group pairs (FMA, split over rows)
//                             k_pick_wide   greedy delta test on the full cosine matrix + permute_marked replayed
//                                           literally (src/dgeqrdm_work.c:149-262) on a marks array -> a swap list
//                             k_swap_cols   the swaps applied to the columns of A (each row by one thread, in order)
//   factorisation in MICRO-PANELS of <= 64 of the selected columns: k_micro_begin(t) points qrdm_ctrl at micro-panel t,
//                             the ordinary panel + trailing-update kernels run on it (the trailing update covers the
//                             selected columns still to come as well — the reference applies each reflector to the
//                             rest of its panel at once, src/dgeqr2.c:179-186, the product of reflectors is the same),
//                             the DM stop threshold is carried from micro-panel to micro-panel (ctrl->micro_thres2) and
//                             a stop ends the block; k_micro_end restores (j, fjb, fjb_cmp) of the whole block for the
//                             norm downdate and the next selection.
// Correctness path, not a tuned one: nb = 64 is the reference's only documented setting (test.ipynb cell 9).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

#define WD_LD QRDM_CANDMAX  // leading dimension of the wide Gram matrix

// ---------------------------------------------------------------- Gram: 64 x 64 block (bx, by) of X'X, bx <= by
#define WG_ROWS 32
#define WG_LDS 66
__global__ void __launch_bounds__(256) k_gram_wide_partial(qrdm_prob P, double* part, int chunk) {
  __shared__ __align__(16) double tx[WG_ROWS * WG_LDS], ty[WG_ROWS * WG_LDS];
  __shared__ int sx[64], sy[64];
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nc = ctrl->nc, j = ctrl->j;
  if (nc <= 1) return;
  // block pair index -> (bx, by), bx <= by, over nb64 = ceil(nc / 64) groups
  const int nb64 = (nc + 63) / 64;
  int pr = blockIdx.y, bx = 0, by = 0;
  for (bx = 0; bx < nb64; ++bx) {
    const int cnt = nb64 - bx;
    if (pr < cnt) { by = bx + pr; break; }
    pr -= cnt;
  }
  if (bx >= nb64) return;
  if (tid < 64) {
    const int cx = 64 * bx + tid, cy = 64 * by + tid;
    sx[tid] = cx < nc ? ctrl->cand[cx] : -1;
    sy[tid] = cy < nc ? ctrl->cand[cy] : -1;
  }
  __syncthreads();
  const int r_lo = qrdm_jr(P, j), r_hi = P.m;
  const int my_lo = r_lo + blockIdx.x * chunk, my_hi = min(r_hi, my_lo + chunk);
  const double* base = P.a + (size_t)j * P.lda;
  const int ty4 = (tid >> 4) * 4, tx4 = (tid & 15) * 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int r0 = my_lo; r0 < my_hi; r0 += WG_ROWS) {
    for (int q = 0; q < 8; ++q) {
      const int c = wid + 8 * q, rr = lane, r = r0 + rr;
      const bool ok = r < my_hi;
      tx[rr * WG_LDS + c] = (ok && sx[c] >= 0) ? base[(size_t)sx[c] * P.lda + r] : 0.0;
      ty[rr * WG_LDS + c] = (ok && sy[c] >= 0) ? base[(size_t)sy[c] * P.lda + r] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WG_ROWS; ++r) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { av[a] = tx[r * WG_LDS + ty4 + a]; bv[a] = ty[r * WG_LDS + tx4 + a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  double* out = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4096;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) out[(ty4 + a) * 64 + tx4 + b] = acc[a][b];
}

// fixed-order sum over the row chunks; G (nc x nc, ld = WD_LD) gets both triangles
__global__ void __launch_bounds__(256) k_gram_wide_reduce(qrdm_prob P, const double* part, double* G, int nparts) {
  const qrdm_ctrl* ctrl = P.ctrl;
  const int nc = ctrl->nc;
  if (nc <= 1) return;
  const int nb64 = (nc + 63) / 64;
  int pr = blockIdx.y, bx = 0, by = 0;
  for (bx = 0; bx < nb64; ++bx) {
    const int cnt = nb64 - bx;
    if (pr < cnt) { by = bx + pr; break; }
    pr -= cnt;
  }
  if (bx >= nb64) return;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 4096; e += gridDim.x * blockDim.x) {
    const double* src = part + (size_t)blockIdx.y * nparts * 4096 + e;
    double s = 0.0;
    for (int q = 0; q < nparts; ++q) s += src[(size_t)q * 4096];
    const int r = 64 * bx + (e >> 6), c = 64 * by + (e & 63);
    if (r < nc && c < nc) { G[(size_t)r * WD_LD + c] = s; G[(size_t)c * WD_LD + r] = s; }
  }
}

// ---------------------------------------------------------------- greedy pick + permute_marked, one CTA
// swaps: [0] = count, then pairs (p, q) of column offsets relative to j, in the order the reference performs them
__global__ void __launch_bounds__(256) k_pick_wide(qrdm_prob P, const double* G, int* marks, int* swaps) {
  __shared__ double inv[QRDM_CANDMAX], mxc[QRDM_CANDMAX];
  __shared__ int selpos[QRDM_CANDMAX], sel[QRDM_CANDMAX];
  __shared__ int s_fjb;
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x;
  const int j = ctrl->j, kmax = ctrl->kmax, nc = ctrl->nc, cols = P.n - j;
  if (tid == 0) swaps[0] = 0;
  if (kmax == 0) return;
  if (ctrl->forced) {  // fixed columns: the next nc columns as they stand
    for (int s = tid; s < nc; s += blockDim.x) ctrl->sel[s] = s;
    if (tid == 0) { ctrl->fjb = nc; ctrl->ncyc = 0; ctrl->cyc_start[0] = 0; }
    return;
  }
  for (int t = tid; t < nc; t += blockDim.x) {
    inv[t] = 1.0 / ctrl->candnrm[t];  // cc = 1 / norm, src/dgeqrdm_work.c:366
  }
  __syncthreads();
  // running maximum of |cos| against the accepted columns; a NaN cosine is ignored like the reference's `maxval < fabs()`
  for (int t = tid; t < nc; t += blockDim.x) mxc[t] = nc > 1 ? fmax(0.0, fabs(G[t] * inv[0] * inv[t])) : 0.0;
  if (tid == 0) { selpos[0] = 0; s_fjb = 1; }
  __syncthreads();
  for (int t = 1; t < nc; ++t) {  // src/dgeqrdm_work.c:382-403
    const bool accept = mxc[t] < P.delta && s_fjb < kmax;  // uniform: shared values, read before the barrier below
    __syncthreads();
    if (accept) {
      if (tid == 0) { selpos[s_fjb] = t; s_fjb = s_fjb + 1; }
      for (int u = tid; u < nc; u += blockDim.x) mxc[u] = fmax(mxc[u], fabs(G[(size_t)t * WD_LD + u] * inv[t] * inv[u]));
    }
    __syncthreads();
  }
  const int fjb = s_fjb;
  for (int s = tid; s < fjb; s += blockDim.x) {
    const int c = ctrl->cand[selpos[s]];
    sel[s] = c;
    ctrl->sel[s] = c;
    marks[c] = 1;
  }
  __syncthreads();
  __threadfence_block();
  if (tid == 0) {
    // permute_marked with nz = 0 (src/dgeqrdm_work.c:149-262), literally: swap(p, q) exchanges marks, jpvt and vn1
    // (NOT vn2: reference quirk) and is recorded for k_swap_cols
    int nsw = 0, jb = 0, jt = cols - 1;
    auto swp = [&](int p, int q) {
      const int mt = marks[p]; marks[p] = marks[q]; marks[q] = mt;
      const int jp = P.jpvt[j + p]; P.jpvt[j + p] = P.jpvt[j + q]; P.jpvt[j + q] = jp;
      const double v = P.vn1[j + p]; P.vn1[j + p] = P.vn1[j + q]; P.vn1[j + q] = v;
      swaps[1 + 2 * nsw] = p; swaps[2 + 2 * nsw] = q;
      ++nsw;
    };
    const int cap = 2 * QRDM_CANDMAX + 4;
    bool overflow = false;
    for (int s = 0; s < fjb && !overflow; ++s) {
      const int jc = sel[s];
      while (jb < jt && marks[jt] == 1) {
        if (nsw >= cap) { overflow = true; break; }
        swp(jt, jb);
        while (jb < cols && marks[jb] != 0) ++jb;
      }
      if (overflow) break;
      if (marks[jc] == 1) {
        while (jb < cols && marks[jb] != 0) ++jb;
        if (jc <= jb || jc < fjb) continue;
        if (jb < cols && marks[jb] == 0) {
          if (nsw >= cap) { overflow = true; break; }
          swp(jc, jb);
          ++jb;
        }
      }
    }
    // every selected column now sits in one of the leading fjb slots: clear the marks for the next iteration
    for (int p = 0; p < cols && p < fjb + 2; ++p) marks[p] = 0;
    for (int s = 0; s < fjb; ++s) marks[sel[s]] = 0;
    swaps[0] = nsw;
    ctrl->fjb = fjb;
    ctrl->ncyc = 0;
    ctrl->cyc_start[0] = 0;
    ctrl->stat_perm_cols += 2 * nsw;
    if (overflow) ctrl->err = QRDM_ERR_INTERNAL;
  }
}

// the recorded exchanges applied to the full-height columns of A: thread <-> row, swaps in order
__global__ void __launch_bounds__(256) k_swap_cols(qrdm_prob P, const int* swaps) {
  const int nsw = swaps[0];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (nsw <= 0 || r >= P.m) return;
  double* a = P.a + (size_t)P.ctrl->j * P.lda + r;
  for (int s = 0; s < nsw; ++s) {
    const size_t p = (size_t)swaps[1 + 2 * s] * P.lda, q = (size_t)swaps[2 + 2 * s] * P.lda;
    const double t = a[p];
    a[p] = a[q];
    a[q] = t;
  }
}

// ---------------------------------------------------------------- micro-panels
__device__ __forceinline__ void micro_collect(qrdm_ctrl* c) {
  if (c->w_pending) {  // account for the micro-panel that just ran
    c->w_k += c->fjb_cmp;
    if (c->fjb_cmp < c->fjb) c->w_stop = 1;  // DM early stop inside it: the block ends here
    c->w_pending = 0;
  }
}
__global__ void k_micro_begin(qrdm_prob P, int t) {
  qrdm_ctrl* c = P.ctrl;
  if (t == 0) { c->w_j0 = c->j; c->w_fjb = c->fjb; c->w_k = 0; c->w_stop = 0; c->w_pending = 0; }
  micro_collect(c);
  const int rem = c->w_fjb - QRDM_KMAX * t;
  c->micro_t = t;
  c->j = c->w_j0 + QRDM_KMAX * t;
  c->fjb_cmp = 0;
  if (c->w_stop || rem <= 0 || c->err != 0) {
    c->fjb = 0;  // the panel and trailing kernels return at once
  } else {
    c->fjb = rem < QRDM_KMAX ? rem : QRDM_KMAX;
    c->w_pending = 1;
  }
}
__global__ void k_micro_end(qrdm_prob P) {
  qrdm_ctrl* c = P.ctrl;
  micro_collect(c);
  c->j = c->w_j0;
  c->fjb = c->w_fjb;
  c->fjb_cmp = c->w_k;
  c->micro_t = 0;
}

extern "C" {

int qrdm_k_gram_wide(const qrdm_prob* p, double* part, double* G, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int nbmax = p->nb < QRDM_CANDMAX ? p->nb : QRDM_CANDMAX;
  const int nb64 = (nbmax + 63) / 64, npairs = nb64 * (nb64 + 1) / 2;
  int g = (rows_hint + 255) / 256;
  if (g < 1) g = 1;
  if (g > QRDM_WIDE_ROWCTAS) g = QRDM_WIDE_ROWCTAS;
  int chunk = (rows_hint + g - 1) / g;
  chunk = (chunk + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  k_gram_wide_partial<<<dim3(g, npairs), 256, 0, s>>>(*p, part, chunk);
  QRDM_LAUNCH_CHECK();
  k_gram_wide_reduce<<<dim3(4, npairs), 256, 0, s>>>(*p, part, G, g);
  QRDM_LAUNCH_CHECK();
  return 0;
}
int qrdm_k_pick_wide(const qrdm_prob* p, const double* G, int* marks, int* swaps, void* stream) {
  k_pick_wide<<<1, 256, 0, (cudaStream_t)stream>>>(*p, G, marks, swaps);
  QRDM_LAUNCH_CHECK();
  k_swap_cols<<<(p->m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, swaps);
  QRDM_LAUNCH_CHECK();
  return 0;
}
int qrdm_k_micro_begin(const qrdm_prob* p, int t, void* stream) {
  k_micro_begin<<<1, 1, 0, (cudaStream_t)stream>>>(*p, t);
  QRDM_LAUNCH_CHECK();
  return 0;
}
int qrdm_k_micro_end(const qrdm_prob* p, void* stream) {
  k_micro_end<<<1, 1, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
