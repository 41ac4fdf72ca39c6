// k_norms.cu — K1 (column-norm initialisation) and K2 (partial-norm downdate with the
// dgeqp3-style recompute guard), the fused HBM-bound pass of BASELINE.json's north_star.
//
// Replaces: reference src/dgeqrdm_work.c:672-682 (initial cblas_dnrm2 loop) and
//           norm_update, src/dgeqrdm_work.c:36-122.
// Algorithmic bytes: K1 8*m*n; K2 8*k*n_r (+ 8*m_r per recomputed column).
#include "common.cuh"

// Sum of squares of A[r0:m, c] for the columns of a list (or all columns c0..n-1), rows split in
// `nsplit` contiguous segments (gridDim.y) so tall matrices fill the machine; partial results go
// to part[split * n + idx] and are combined in fixed order by k_colnorm_finalize (deterministic;
// in the row-sharded build the all-reduce sits between the two kernels).
__global__ void __launch_bounds__(256) k_colnorm_partial(qrdm_prob P, int use_list) {
  __shared__ double scratch[32];
  const qrdm_ctrl* ctrl = P.ctrl;
  int r0, c0, count;
  if (use_list == 2) {  // row-sharded recompute: flag array over the active columns
    const int cfirst = ctrl->j + ctrl->fjb_cmp;
    r0 = qrdm_jr(P, cfirst);
    const int len2 = P.m - r0, nsplit2 = gridDim.y, split2 = blockIdx.y;
    int seg2 = (len2 + nsplit2 - 1) / nsplit2;
    seg2 = (seg2 + 1) & ~1;
    const int lo2 = min(len2, split2 * seg2), hi2 = min(len2, lo2 + seg2);
    for (int c = cfirst + blockIdx.x; c < P.n; c += gridDim.x) {
      double t = 0.0;
      if (ctrl->nflag > 0 && P.flag_list[c]) {
        const double* col = P.a + (size_t)c * P.lda + r0;
        double s0 = 0.0;
        for (int q = lo2 + threadIdx.x; q < hi2; q += blockDim.x) s0 = fma(col[q], col[q], s0);
        t = block_sum(s0, scratch);
      }
      if (threadIdx.x == 0) P.nrm_part[(size_t)split2 * P.n + c] = t;
    }
    return;
  }
  if (use_list) {
    count = ctrl->nflag;
    r0 = qrdm_jr(P, ctrl->j + ctrl->fjb_cmp);  // (local) rows below the block just factored
    c0 = 0;
  } else {
    count = P.n;
    r0 = 0;
    c0 = 0;
  }
  const int len = P.m - r0;
  const int nsplit = gridDim.y, split = blockIdx.y;
  int seg = (len + nsplit - 1) / nsplit;
  seg = (seg + 1) & ~1;
  const int lo = min(len, split * seg), hi = min(len, lo + seg);
  for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const int c = use_list ? P.flag_list[idx] : c0 + idx;
    const double* col = P.a + (size_t)c * P.lda + r0;
    double s0 = 0.0, s1 = 0.0;
    int i = lo + threadIdx.x * 2;
    // 16-byte vector loads when the segment start is 16B aligned, scalar otherwise
    if (((uintptr_t)(col + lo) & 15) == 0) {
      for (; i + 1 < hi; i += 2 * blockDim.x) {
        const double2 v = *reinterpret_cast<const double2*>(col + i);
        s0 = fma(v.x, v.x, s0);
        s1 = fma(v.y, v.y, s1);
      }
      if (i < hi) s0 = fma(col[i], col[i], s0);
    } else {
      for (int q = lo + threadIdx.x; q < hi; q += blockDim.x) s0 = fma(col[q], col[q], s0);
    }
    const double t = block_sum(s0 + s1, scratch);
    if (threadIdx.x == 0) P.nrm_part[(size_t)split * P.n + idx] = t;
  }
}

__global__ void __launch_bounds__(256) k_colnorm_finalize(qrdm_prob P, int use_list, int nsplit) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (use_list == 2) {  // row-sharded recompute: nrm_part[q*n + c] holds the all-reduced partials of flagged columns
    const int c = P.ctrl->j + P.ctrl->fjb_cmp + idx;
    if (c >= P.n || P.ctrl->nflag == 0 || !P.flag_list[c]) return;
    double s2 = 0.0;
    for (int q = 0; q < nsplit; ++q) s2 += P.nrm_part[(size_t)q * P.n + c];
    const double v2 = sqrt(s2);
    P.vn1[c] = v2;
    P.vn2[c] = v2;
    return;
  }
  const int count = use_list ? P.ctrl->nflag : P.n;
  if (idx >= count) return;
  double s = 0.0;
  for (int q = 0; q < nsplit; ++q) s += P.nrm_part[(size_t)q * P.n + idx];
  const double v = sqrt(s);
  const int c = use_list ? P.flag_list[idx] : idx;
  P.vn1[c] = v;
  P.vn2[c] = v;
  if (!use_list && !P.keep_jpvt) P.jpvt[c] = c + 1;  // src/dgeqrdm_work.c:596-609 with every column free
}

static int colnorm_nsplit(const qrdm_prob* p, int* gx_out) {
  const int len = p->m;
  int nsplit = 1;
  const long target = 4L * p->sm_count;
  if (p->n < target) nsplit = (int)min((long)p->nrm_splits, max(1L, min(target / max(1, p->n), (long)(len / 4096))));
  if (nsplit < 1 || p->nranks > 1) nsplit = 1;  // row-sharded: the all-reduce length must agree on all ranks
  *gx_out = (int)min((long)p->n, max(1L, target / nsplit));
  return nsplit;
}
extern "C" int qrdm_k_colnorm_part(const qrdm_prob* p, int use_list, int* nsplit_out, void* stream) {
  int gx;
  const int nsplit = colnorm_nsplit(p, &gx);
  k_colnorm_partial<<<dim3(gx, nsplit), 256, 0, (cudaStream_t)stream>>>(*p, use_list);
  QRDM_LAUNCH_CHECK();
  *nsplit_out = nsplit;
  return 0;
}
extern "C" int qrdm_k_colnorm_fin(const qrdm_prob* p, int use_list, int nsplit, void* stream) {
  k_colnorm_finalize<<<(p->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, use_list, nsplit);
  QRDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int qrdm_k_colnorm(const qrdm_prob* p, int use_list, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  // rows per column segment >= 4096 so each CTA streams >= 32 KB; enough CTAs for ~4 per SM
  const int len = p->m;
  int nsplit = 1;
  const long target = 4L * p->sm_count;
  if (p->n < target) nsplit = (int)min((long)p->nrm_splits, max(1L, min(target / max(1, p->n), (long)(len / 4096))));
  if (nsplit < 1) nsplit = 1;
  const int gx = (int)min((long)p->n, max(1L, target / nsplit));
  k_colnorm_partial<<<dim3(gx, nsplit), 256, 0, s>>>(*p, use_list);
  QRDM_LAUNCH_CHECK();
  k_colnorm_finalize<<<(p->n + 255) / 256, 256, 0, s>>>(*p, use_list, nsplit);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// K2: one warp per active column c >= j + k.  d = sum of squares of the k new R rows of that
// column; t = max(0,(1+d/vn1)(1-d/vn1)); if t*(vn1/vn2)^2 <= sqrt(eps) the column goes on the
// exact-recompute list, else vn1 *= sqrt(t)  (reference src/dgeqrdm_work.c:81-108).
// MODE 0: fused (single GPU); 1: partial sums only; 2: apply the guard to all-reduced sums
template <int MODE>
__global__ void __launch_bounds__(256) k_norm_update(qrdm_prob P, double tol3z) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int j = ctrl->j, k = ctrl->fjb_cmp;
  const int lane = threadIdx.x & 31;
  const int c = j + k + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= P.n) return;
  const double v1 = P.vn1[c];
  if (v1 == 0.0) { if (MODE == 1 && lane == 0) P.nrm_part[c] = 0.0; return; }
  double d = 0.0;
  if (MODE != 2) {  // sum of squares of the (local part of the) k new R rows of this column
    const int rlo = qrdm_jr(P, j), rhi = qrdm_jr(P, j + k);
    const double* col = P.a + (size_t)c * P.lda;
    for (int r = rlo + lane; r < rhi; r += 32) d = fma(col[r], col[r], d);
    d = warp_sum(d);
    if (MODE == 1) {  // row-sharded, part 1: publish the partial; the all-reduce follows
      if (lane == 0) P.nrm_part[c] = d;
      return;
    }
  } else {
    d = P.nrm_part[c];  // row-sharded, part 2: all-reduced sum
  }
  if (lane == 0) {
    // in the caller's scale (inv_scale = 1 unless the input was pre-scaled: then x * 1.0 is exact and nothing changes)
    const double dt = (d * P.inv_scale) * P.inv_scale;
    double t = sqrt(fabs(dt)) / (v1 * P.inv_scale);
    t = (t + 1.0) * (1.0 - t);
    t = (0.0 >= t) ? 0.0 : t;
    const double q = v1 / P.vn2[c];
    const double t2 = t * (q * q);
    if (MODE == 2) P.flag_list[c] = 0;  // row-sharded: flag_list is a per-column flag ARRAY so that every
                                        // rank recomputes the same columns in the same slots
    if (t2 <= tol3z) {
      if (P.m_glob - (j + k) > 0) {
        if (MODE == 2) { P.flag_list[c] = 1; atomicAdd(&ctrl->nflag, 1); }
        else { const int slot = atomicAdd(&ctrl->nflag, 1); P.flag_list[slot] = c; }
      } else {
        P.vn1[c] = 0.0;
        P.vn2[c] = 0.0;
      }
    } else {
      P.vn1[c] = v1 * sqrt(t);
    }
  }
}

extern "C" int qrdm_k_norm_update(const qrdm_prob* p, int j_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int maxcols = p->n - j_host - 1;  // k >= 1
  if (maxcols <= 0) return 0;
  k_norm_update<0><<<(maxcols + 7) / 8, 256, 0, s>>>(*p, 1.0536712127723509e-08 /* tol3z = sqrt(dlamch('e')) = sqrt(2^-53), src/dgeqrdm_work.c:528-529 */);
  QRDM_LAUNCH_CHECK();
  return qrdm_k_colnorm(p, 1, stream);
}

// End of the fixed-column phase: the reference computes the norms of the free columns from scratch on the rows below
// the fixed block (src/dgeqrdm_work.c:672-682) — flag every remaining column and run the exact recompute of K2.
__global__ void __launch_bounds__(256) k_flag_all(qrdm_prob P) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int c0 = ctrl->j + ctrl->fjb_cmp, cnt = P.n - c0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) P.flag_list[i] = c0 + i;
  if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->nflag = cnt > 0 ? cnt : 0;
}
extern "C" int qrdm_k_norm_recompute_all(const qrdm_prob* p, int j_host, void* stream) {
  const int maxcols = p->n - j_host - 1;
  if (maxcols <= 0) return 0;
  k_flag_all<<<(maxcols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return qrdm_k_colnorm(p, 1, stream);
}

// deferred trailing update: the flagged columns are brought up to date before their exact recompute
extern "C" int qrdm_k_norm_update_lazy(const qrdm_prob* p, int j_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int maxcols = p->n - j_host - 1;
  if (maxcols <= 0) return 0;
  k_norm_update<0><<<(maxcols + 7) / 8, 256, 0, s>>>(*p, 1.0536712127723509e-08);
  QRDM_LAUNCH_CHECK();
  const int rc = qrdm_k_colupd(p, 1, j_host, stream);
  if (rc) return rc;
  return qrdm_k_colnorm(p, 1, stream);
}

extern "C" int qrdm_k_norm_dpart(const qrdm_prob* p, int j_host, void* stream) {
  const int maxcols = p->n - j_host - 1;
  if (maxcols <= 0) return 0;
  k_norm_update<1><<<(maxcols + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*p, 0.0);
  QRDM_LAUNCH_CHECK();
  return 0;
}
extern "C" int qrdm_k_norm_apply(const qrdm_prob* p, int j_host, void* stream) {
  const int maxcols = p->n - j_host - 1;
  if (maxcols <= 0) return 0;
  k_norm_update<2><<<(maxcols + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*p, 1.0536712127723509e-08);
  QRDM_LAUNCH_CHECK();
  return 0;
}


// ---- badly scaled inputs -------------------------------------------------------------------------
// The reference computes norms with the scaled cblas_dnrm2 (src/dgeqrdm_work.c:69,96,673) and dlarfg rescales by
// safmin (src/dlarfg.c:144-182), so it factors matrices with entries around 1e+-200; plain sums of squares do not.
// Rather than carrying scale factors through every reduction, the driver multiplies such an input by ONE power of
// two (exact: no rounding, the pivots, tau and V are invariant, R comes out scaled by the same factor), factors it
// with the unchanged kernels and divides the R-like entries back at the end.  Only taken when the largest column
// norm leaves [2^-300, 2^300]; every other input never sees these kernels.
__global__ void __launch_bounds__(256) k_amax(qrdm_prob P, double* out) {
  __shared__ double scratch[8];
  double mx = 0.0;
  const size_t total = (size_t)P.m * (size_t)P.n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t c = e / (size_t)P.m, r = e - c * (size_t)P.m;
    mx = fmax(mx, fabs(P.a[c * (size_t)P.lda + r]));  // fmax drops NaNs: they are the NaN screen's business
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t = fmax(t, scratch[w]);
    out[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_scale(qrdm_prob P, double s, int mode, int r) {
  const size_t total = (size_t)P.m * (size_t)P.n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t c = e / (size_t)P.m, row = e - c * (size_t)P.m;
    // mode 1: global row index of a local row (row-sharded: row0 + row); R-like <=> row <= column, or column >= r
    if (mode == 1 && (int)c < r && (long long)P.row0 + (long long)row > (long long)c) continue;
    P.a[c * (size_t)P.lda + row] *= s;
  }
}
extern "C" int qrdm_k_amax(const qrdm_prob* p, double* out, int* nparts, void* stream) {
  const int g = 4 * p->sm_count < 1024 ? 4 * p->sm_count : 1024;
  k_amax<<<g, 256, 0, (cudaStream_t)stream>>>(*p, out);
  QRDM_LAUNCH_CHECK();
  *nparts = g;
  return 0;
}
extern "C" int qrdm_k_scale(const qrdm_prob* p, double s, int mode, int r, void* stream) {
  const int g = 8 * p->sm_count;
  k_scale<<<g, 256, 0, (cudaStream_t)stream>>>(*p, s, mode, r);
  QRDM_LAUNCH_CHECK();
  return 0;
}
