// k_gram.cu — K3b / K5: tall-skinny Gram matrix of up to 64 gathered columns, G = X' X.
//
// Used twice per iteration:
//   of_v = 0: X = A[j:m, j + cand[0..nc)]  — the candidates' Gram matrix; the cosine matrix of
//             DM_perm is G[s][t] / (vn1_s vn1_t).  Replaces the rescale + cblas_dsyrk of reference
//             src/dgeqrdm_work.c:365-379 without materialising the scaled copy `as`.
//   of_v = 1: X = Vc[:, 0..kpad)            — V'V, from which the trailing update applies T'
//             by forward substitution (replaces LAPACKE_dlarft, src/dgeqrdm_work.c:751-754).
// Rows are split over CTAs (split-K); partial 64x64 blocks go to gram_part[cta] and are summed in
// fixed order by k_gram_reduce (deterministic; row-sharded build: all-reduce in between).
// HBM/L2-bound: algorithmic bytes 8 * rows * nc.
#include "common.cuh"

#define GR_ROWS 64
#define GR_LD 66  // doubles per smem row: 64 columns + 2 pad (keeps 16B alignment of float4 reads)

__global__ void __launch_bounds__(256) k_gram_partial(qrdm_prob P, int of_v) {
  __shared__ __align__(16) double tile[GR_ROWS * GR_LD];
  __shared__ int scol[64];
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j;
  int nc, r_lo, r_hi, ld;
  const double* base;
  if (of_v) {
    nc = (ctrl->fjb_cmp + 7) & ~7;
    base = P.vc; ld = P.ldv;
    r_lo = qrdm_jr(P, j); r_hi = P.m;
    if (tid < 64) scol[tid] = tid;
  } else {
    nc = ctrl->nc;
    if (nc <= 1) return;  // a single candidate is always taken, no cosines needed
    base = P.a + (size_t)j * P.lda; ld = P.lda;
    r_lo = qrdm_jr(P, j); r_hi = P.m;
    if (tid < 64) scol[tid] = tid < nc ? ctrl->cand[tid] : 0;
  }
  __syncthreads();
  const int rows = r_hi - r_lo;
  int chunk = (rows + gridDim.x - 1) / gridDim.x;
  chunk = (chunk + GR_ROWS - 1) / GR_ROWS * GR_ROWS;
  const int my_lo = r_lo + blockIdx.x * chunk, my_hi = min(r_hi, my_lo + chunk);

  const int ty = tid >> 4, tx = tid & 15;  // 4x4 block (rows 4ty.., cols 4tx..) of G per thread
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  // software pipeline: the global loads of sub-chunk i+1 are in flight while sub-chunk i is
  // multiplied out of shared memory (warps over columns, lanes along rows: coalesced)
  double pre[2][8];  // 16 independent loads in flight per thread
  auto fetch = [&](int r0) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = wid + 8 * q, r = r0 + lane + 32 * h;
        pre[h][q] = (c < nc && r < my_hi) ? base[(size_t)scol[c] * ld + r] : 0.0;
      }
  };
  if (my_lo < my_hi) fetch(my_lo);
  for (int r0 = my_lo; r0 < my_hi; r0 += GR_ROWS) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 8; ++q) tile[(lane + 32 * h) * GR_LD + wid + 8 * q] = pre[h][q];
    __syncthreads();
    if (r0 + GR_ROWS < my_hi) fetch(r0 + GR_ROWS);
#pragma unroll 4
    for (int r = 0; r < GR_ROWS; ++r) {
      const double2 a01 = *reinterpret_cast<const double2*>(&tile[r * GR_LD + 4 * ty]);
      const double2 a23 = *reinterpret_cast<const double2*>(&tile[r * GR_LD + 4 * ty + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&tile[r * GR_LD + 4 * tx]);
      const double2 b23 = *reinterpret_cast<const double2*>(&tile[r * GR_LD + 4 * tx + 2]);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y};
      const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  double* out = P.gram_part + (size_t)blockIdx.x * 4096;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) out[(4 * ty + a) * 64 + 4 * tx + b] = acc[a][b];
}

__global__ void __launch_bounds__(64) k_gram_reduce(qrdm_prob P, int of_v, int nparts) {
  if (!of_v && P.ctrl->nc <= 1) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // 4096 entries
  const double* src = P.gram_part + e;
  // 16 loads in flight per thread (the kernel is pure L2 latency); fixed association order: deterministic
  double s[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) s[u] = 0.0;
  int q = 0;
  for (; q + 15 < nparts; q += 16) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = src[(size_t)(q + u) * 4096];
#pragma unroll
    for (int u = 0; u < 16; ++u) s[u] += v[u];
  }
  for (; q < nparts; ++q) s[0] += src[(size_t)q * 4096];
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int u = 0; u < w; ++u) s[u] += s[u + w];
  P.gram[e] = s[0];
}

// rows_hint: host-side upper bound of the number of rows (m - j) used to size the grid.
extern "C" int qrdm_k_gram(const qrdm_prob* p, int of_v, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  // >= 128 rows (two pipelined sub-chunks) per CTA: the reduce kernel then sums 128 partial blocks at
  // m = 16384 instead of 256 (it was 30 us of pure L2 latency per iteration)
  int g = (rows_hint + 127) / 128;
  if (g < 1) g = 1;
  if (g > QRDM_GRAM_MAXCTA) g = QRDM_GRAM_MAXCTA;
  if (g > 2 * p->sm_count) g = 2 * p->sm_count;
  k_gram_partial<<<g, 256, 0, s>>>(*p, of_v);
  QRDM_LAUNCH_CHECK();
  k_gram_reduce<<<64, 64, 0, s>>>(*p, of_v, g);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// row-sharded: partial + local reduce into p->gram; the caller all-reduces p->gram (4096 doubles)
extern "C" int qrdm_k_gram_part(const qrdm_prob* p, int rows_hint, void* stream) { return qrdm_k_gram(p, 0, rows_hint, stream); }
