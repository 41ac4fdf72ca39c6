// k_skinny.cu — block-reflector update of the REST OF A TALL PANEL by one 8-column sub-panel
// (blocked tall-panel mode, see qrdm_k_panel): C_p <- (I - V T V')' C_p with V = m_r x k (k <= 8)
// and C_p = the <= 56 remaining panel columns.  The K6 kernels are built for k = 64 and wide C; on
// this shape they were launch/overhead bound (1.9 ms per call against 0.14 ms of HBM time), so the
// skinny case gets three bandwidth-bound FMA kernels:
//   k_sub_w      partial  W_ext = V' [V | C_p]  (8 x 64) per 2048-row chunk, V chunk staged in smem
//   k_sub_w2     fixed-order reduction over chunks, T' = (I + D N)^-1 D from V'V, W2 = -T' W
//   k_sub_apply  C_p += V W2
// Algorithmic bytes: 8*m_r*(2*ncp + 16) read + 8*m_r*ncp written.
#include "common.cuh"

#define SK_RC 2048           // rows per CTA chunk
#define SK_THREADS 512       // thread = (column c = tid >> 6, row lane rl = tid & 63)
#define SK_SMEM (SK_RC * 8 * 8)

struct SkGeom { int j, jr, k, ncp, rows, c0, voff; };
__device__ __forceinline__ SkGeom sk_geom(const qrdm_prob& P) {
  const QrdmGeom q = qrdm_geom(P);
  SkGeom g;
  g.j = q.j; g.k = q.k; g.voff = q.voff;
  g.jr = qrdm_jr(P, q.j);
  g.rows = P.m - g.jr;
  g.c0 = q.j + q.fjb;              // first column of the rest of the panel
  g.ncp = q.n_end - g.c0;
  if (q.fjb <= 0) g.k = 0;
  return g;
}
__device__ __forceinline__ void sk_load_v(const qrdm_prob& P, const SkGeom& g, double* Vs, int row0, int nrc) {
  for (int e = threadIdx.x; e < SK_RC * 8; e += SK_THREADS) {
    const int q = e / SK_RC, r = e - q * SK_RC;  // consecutive threads -> consecutive rows (coalesced)
    Vs[r * 8 + q] = (r < nrc && q < g.k) ? P.vc[(size_t)(g.voff + q) * P.ldv + g.jr + row0 + r] : 0.0;
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1) k_sub_w(qrdm_prob P) {
  extern __shared__ __align__(16) double Vs[];  // [row][8]
  __shared__ double red[SK_THREADS / 32][8];
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, c = tid >> 6, rl = tid & 63;
  const int row0 = blockIdx.x * SK_RC;
  if (row0 >= g.rows) return;
  const int nrc = min(SK_RC, g.rows - row0);
  sk_load_v(P, g, Vs, row0, nrc);
  __syncthreads();
  const int ngroups = 1 + (g.ncp + 7) / 8;
  double* out = P.gram_part + (size_t)blockIdx.x * 512;  // [group][q][c]: 8 groups x 64
  for (int gi = 0; gi < ngroups; ++gi) {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    const int col = (gi - 1) * 8 + c;
    const bool ok = gi == 0 || col < g.ncp;
    const double* src = P.a + (size_t)(g.c0 + (gi == 0 ? 0 : col)) * P.lda + g.jr + row0;
#pragma unroll 4
    for (int r = rl; r < nrc; r += 64) {
      const double x = gi == 0 ? Vs[r * 8 + c] : (ok ? src[r] : 0.0);
      const double2 v01 = *reinterpret_cast<const double2*>(Vs + r * 8);
      const double2 v23 = *reinterpret_cast<const double2*>(Vs + r * 8 + 2);
      const double2 v45 = *reinterpret_cast<const double2*>(Vs + r * 8 + 4);
      const double2 v67 = *reinterpret_cast<const double2*>(Vs + r * 8 + 6);
      acc[0] = fma(v01.x, x, acc[0]); acc[1] = fma(v01.y, x, acc[1]);
      acc[2] = fma(v23.x, x, acc[2]); acc[3] = fma(v23.y, x, acc[3]);
      acc[4] = fma(v45.x, x, acc[4]); acc[5] = fma(v45.y, x, acc[5]);
      acc[6] = fma(v67.x, x, acc[6]); acc[7] = fma(v67.y, x, acc[7]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = warp_sum(acc[q]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) red[wid][q] = acc[q];
    }
    __syncthreads();
    if (tid < 64) {  // (q, cc): the two warps of column cc
      const int q = tid >> 3, cc = tid & 7;
      out[gi * 64 + q * 8 + cc] = red[2 * cc][q] + red[2 * cc + 1][q];
    }
  }
}

__global__ void __launch_bounds__(512) k_sub_w2(qrdm_prob P, int nchunks_max) {
  __shared__ double W[8 * 64];   // [group][q][c]
  __shared__ double T[64];       // T'[q][p]
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int tid = threadIdx.x;
  const int nchunks = P.w_reduced ? 1 : min(nchunks_max, (g.rows + SK_RC - 1) / SK_RC);  // w_reduced: chunk 0 = all-reduced sum
  const int ngroups = 1 + (g.ncp + 7) / 8;
  if (tid < ngroups * 64) {
    double s = 0.0;
    for (int b = 0; b < nchunks; ++b) s += P.gram_part[(size_t)b * 512 + tid];  // fixed order
    W[tid] = s;
  }
  __syncthreads();
  if (tid < 8) {  // column p = tid of X = (I + D N)^-1, N = strictly-lower V'V, D = diag(tau); T' = X D
    const int p = tid;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double a = (i == p) ? 1.0 : 0.0;
      const double ti = i < g.k ? P.tau[g.j + i] : 0.0;
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < i) a = fma(-ti * W[s * 8 + i], x[s], a);  // (V'V)[s][i] = W[group 0][q = s][c = i]
      x[i] = a;
    }
    const double tp = p < g.k ? P.tau[g.j + p] : 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) T[i * 8 + p] = (i < g.k) ? x[i] * tp : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < 8 * g.ncp; e += blockDim.x) {  // W2[q][col] = -sum_p T'[q][p] W[p][col]
    const int q = e / g.ncp, col = e - q * g.ncp;
    const int gi = 1 + col / 8, cc = col & 7;
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < 8; ++p) s = fma(T[q * 8 + p], W[gi * 64 + p * 8 + cc], s);
    P.w2[(size_t)q * P.ldw + col] = -s;
    if (s != s) atomicCAS(&P.ctrl->err, 0, -13);
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1) k_sub_apply(qrdm_prob P) {
  extern __shared__ __align__(16) double Vs[];  // [row][8]
  __shared__ double W2s[8 * 64];
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int tid = threadIdx.x, c = tid >> 6, rl = tid & 63;
  const int row0 = blockIdx.x * SK_RC;
  if (row0 >= g.rows) return;
  const int nrc = min(SK_RC, g.rows - row0);
  sk_load_v(P, g, Vs, row0, nrc);
  for (int e = tid; e < 8 * 64; e += SK_THREADS) {
    const int q = e >> 6, col = e & 63;
    W2s[e] = col < g.ncp ? P.w2[(size_t)q * P.ldw + col] : 0.0;
  }
  __syncthreads();
  const int ngroups = (g.ncp + 7) / 8;
  for (int gi = 0; gi < ngroups; ++gi) {
    const int col = gi * 8 + c;
    if (col >= g.ncp) continue;
    double w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) w[q] = W2s[q * 64 + col];
    double* dst = P.a + (size_t)(g.c0 + col) * P.lda + g.jr + row0;
#pragma unroll 4
    for (int r = rl; r < nrc; r += 64) {
      const double2 v01 = *reinterpret_cast<const double2*>(Vs + r * 8);
      const double2 v23 = *reinterpret_cast<const double2*>(Vs + r * 8 + 2);
      const double2 v45 = *reinterpret_cast<const double2*>(Vs + r * 8 + 4);
      const double2 v67 = *reinterpret_cast<const double2*>(Vs + r * 8 + 6);
      double x = dst[r];
      x = fma(v01.x, w[0], x); x = fma(v01.y, w[1], x); x = fma(v23.x, w[2], x); x = fma(v23.y, w[3], x);
      x = fma(v45.x, w[4], x); x = fma(v45.y, w[5], x); x = fma(v67.x, w[6], x); x = fma(v67.y, w[7], x);
      dst[r] = x;
    }
  }
}

// row-sharded build: fold the chunk partials into chunk 0 (fixed order) so that 512 doubles can be
// all-reduced; a rank without rows contributes zeros
__global__ void __launch_bounds__(512) k_sub_wred(qrdm_prob P, int nchunks_max) {
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int nchunks = min(nchunks_max, g.rows > 0 ? (g.rows + SK_RC - 1) / SK_RC : 0);
  double s = 0.0;
  for (int b = 0; b < nchunks; ++b) s += P.gram_part[(size_t)b * 512 + threadIdx.x];
  P.gram_part[threadIdx.x] = s;
}

// rows_hint: host-side upper bound of the rows of the sub-panel
extern "C" int qrdm_k_skinny_update(const qrdm_prob* p, int rows_hint, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sub_w, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
    cudaFuncSetAttribute(k_sub_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  k_sub_w<<<nch, SK_THREADS, SK_SMEM, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  k_sub_w2<<<1, 512, 0, s>>>(*p, nch);
  QRDM_LAUNCH_CHECK();
  k_sub_apply<<<nch, SK_THREADS, SK_SMEM, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// row-sharded variant, split at the all-reduce: part 1 = partials folded into gram_part[0..512)
extern "C" int qrdm_k_skinny_part(const qrdm_prob* p, int rows_hint, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sub_w, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
    cudaFuncSetAttribute(k_sub_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  k_sub_w<<<nch, SK_THREADS, SK_SMEM, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  k_sub_wred<<<1, 512, 0, s>>>(*p, nch);
  QRDM_LAUNCH_CHECK();
  return 0;
}
extern "C" int qrdm_k_skinny_finish(const qrdm_prob* p, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  k_sub_w2<<<1, 512, 0, s>>>(*p, 1);
  QRDM_LAUNCH_CHECK();
  k_sub_apply<<<nch, SK_THREADS, SK_SMEM, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
