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
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------
// K3b on the TMA + DMMA path (candidates' Gram, 16-byte aligned even-lda matrices: every production shape).
//
// The m_r x nc candidate panel is gathered column by column — a column chunk of a column-major matrix is one
// contiguous run of bytes — by `cp.async.bulk` (SASS UBLKCP) into a 3-stage shared-memory ring, each stage
// 128 rows x 64 columns, completion signalled on an mbarrier (expect_tx / complete_tx); a dedicated producer warp
// issues the copies and waits on the stage's "empty" barrier, the 8 consumer warps wait on "full".  The product
// runs on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA) with M = N = candidate index, K = row index.  A lane's A
// fragment for column block bi and its B fragment for block bj are the SAME register when bi = bj's block, so a
// k-step costs NB8 shared-memory loads for NB8 (NB8 + 1) / 2 DMMAs (upper block triangle only: G is symmetric).
// Consumer warp w owns the k-steps w, w + 8, ... of every stage (split-K inside the CTA, all warps balanced) and
// keeps the whole upper triangle in registers (72 doubles at nc = 64); the 8 partial triangles are added in warp
// order through shared memory (deterministic) and the CTA writes one 64 x 64 partial block; k_gram_reduce sums the
// CTAs' blocks in fixed order as before.  Column stride GT_TRP = 132 doubles: 16-byte aligned bulk-copy
// destinations and conflict-free fragment loads (lane (g, t) reads word g * 132 + t: banks 8g + 2t mod 32).
// Bound: FP64 tensor pipe at nc = 64 (9 DMMA per row: 36 clk/row/SM against 23 B/clk/SM of HBM = 22 clk per
// 512-byte row) — 0.25 ms per 2,000,000-row Gram against 0.16 ms of pure HBM time; HBM-bound for nc <= 32.
#define GT_TR 128
#define GT_TRP (GT_TR + 4)
#define GT_STAGES 3
#define GT_CONSUMERS 8
#define GT_THREADS ((GT_CONSUMERS + 1) * 32)
#define GT_STAGE_DOUBLES (64 * GT_TRP)
#define GT_SMEM (GT_STAGES * GT_STAGE_DOUBLES * 8 + 128)

// SUBW = true: the same machinery computes the skinny product of the blocked tall panel (k_skinny.cu),
//   W_ext = V' [V | C_p]   (8 x (8 + ncp)) = the first 8 ROWS of the Gram matrix of X = [V_sub | C_p],
// where V_sub = the 8 clean reflector columns of the sub-panel (Vc) and C_p = the <= 56 panel columns behind it: only
// the block row bi = 0 is accumulated (NB8 DMMAs per k-step) and every column is read exactly once — the round-1 FMA
// kernel re-read V once per 8 columns and spent more shuffles on its per-chunk reductions than FMAs on the product
// (402 us per call at 2,000,000 rows against 98 us of HBM time, profiles/r02_launches_c4_summary.txt).
// Output: the CTA's partial in the layout k_sub_w2 folds, gram_part[cta * 512 + group * 64 + q * 8 + c].
// SRC: 0 = the candidates of the selection, 1 = SUBW (above), 2 = the block's clean reflectors Vc (V'V for k_tinv when
// the trailing update of a tall-skinny matrix skips its V'V tile, see k_trailing.cu: P.no_vtv)
template <int NB8, int SRC>
__global__ void __launch_bounds__(GT_THREADS, 1) k_gram_tma(qrdm_prob P, int chunk) {
  extern __shared__ __align__(128) unsigned char gt_smem[];
  double* stage = reinterpret_cast<double*>(gt_smem);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(gt_smem + (size_t)GT_STAGES * GT_STAGE_DOUBLES * 8);
  unsigned long long* empty = full + GT_STAGES;
  __shared__ const double* scolp[64];  // column c of X, local row 0
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr bool SUBW = SRC == 1;
  int nc, r_lo;
  if (SRC == 2) {
    const QrdmGeom q = qrdm_geom(P);
    if (q.k <= 0) return;
    nc = (q.k + 7) & ~7;
    r_lo = qrdm_jr(P, q.j);
    if (tid < 64) scolp[tid] = P.vc + (size_t)(q.voff + (tid < nc ? tid : 0)) * P.ldv;
  } else if (SUBW) {
    const QrdmGeom q = qrdm_geom(P);
    const int c0 = q.j + q.fjb, ncp = q.n_end - c0;
    if (q.fjb <= 0 || q.k <= 0 || ncp <= 0) return;  // dead sub-panel / nothing behind it (k_sub_w2 returns alike)
    nc = 8 + ncp;
    r_lo = qrdm_jr(P, q.j);
    if (tid < 64) scolp[tid] = tid < 8 ? P.vc + (size_t)(q.voff + tid) * P.ldv : P.a + (size_t)(c0 + (tid < nc ? tid - 8 : 0)) * P.lda;
  } else {
    nc = ctrl->nc;
    if (nc <= 1) return;  // a single candidate is always taken, no cosines needed
    const int j = ctrl->j;
    r_lo = qrdm_jr(P, j);
    if (tid < 64) scolp[tid] = P.a + (size_t)(j + (tid < nc ? ctrl->cand[tid] : 0)) * P.lda;
  }
  const int r_hi = P.m;
  const int r_al = r_lo & ~1;  // bulk copies need 16-byte aligned sources: start on an even row, mask the extra one
  const int my_lo = r_al + blockIdx.x * chunk, my_hi = min(r_hi, my_lo + chunk);
  const int nchunks = my_hi > my_lo ? (my_hi - my_lo + GT_TR - 1) / GT_TR : 0;
  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GT_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  constexpr int NBLK = SUBW ? NB8 : NB8 * (NB8 + 1) / 2;
  if (wid == GT_CONSUMERS) {
    // ---- producer warp: one bulk copy per candidate column and stage ----
    for (int it = 0; it < nchunks; ++it) {
      const int s = it % GT_STAGES;
      const unsigned ph = (unsigned)(it / GT_STAGES) & 1u;
      if (it >= GT_STAGES) mbar_wait(&empty[s], ph ^ 1u);
      const int r0 = my_lo + it * GT_TR;
      const int vr = min(GT_TR, my_hi - r0);
      const unsigned bytes = (unsigned)((vr + 1) & ~1) * 8u;  // an odd tail reads one padding row (lda > m there), masked below
      if (lane == 0) mbar_expect_tx(&full[s], bytes * (unsigned)nc);
      __syncwarp();
      for (int c = lane; c < nc; c += 32)
        bulk_g2s(stage + (size_t)s * GT_STAGE_DOUBLES + (size_t)c * GT_TRP, scolp[c] + r0, bytes, &full[s]);
    }
    return;
  }

  // ---- consumer warps ----
  const int g = lane >> 2, t = lane & 3;
  double acc[NBLK][2];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) { acc[b][0] = 0.0; acc[b][1] = 0.0; }
  for (int it = 0; it < nchunks; ++it) {
    const int s = it % GT_STAGES;
    const unsigned ph = (unsigned)(it / GT_STAGES) & 1u;
    const int r0 = my_lo + it * GT_TR;
    const int vr = min(GT_TR, my_hi - r0);
    const int lo = r0 < r_lo ? r_lo - r0 : 0;
    const double* st = stage + (size_t)s * GT_STAGE_DOUBLES;
    mbar_wait(&full[s], ph);
    const int nks = (vr + 3) >> 2;
    for (int ks = wid; ks < nks; ks += GT_CONSUMERS) {
      const int row = ks * 4 + t;
      const bool rv = row >= lo && row < vr;
      double f[NB8];
#pragma unroll
      for (int b = 0; b < NB8; ++b) {
        const int c = 8 * b + g;
        const double x = st[(size_t)c * GT_TRP + row];
        f[b] = (rv && c < nc) ? x : 0.0;
      }
      int q = 0;
#pragma unroll
      for (int bi = 0; bi < (SUBW ? 1 : NB8); ++bi)
#pragma unroll
        for (int bj = bi; bj < NB8; ++bj, ++q) dmma884(acc[q][0], acc[q][1], f[bi], f[bj]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  // ---- add the 8 warps' partial triangles in warp order (deterministic), then write the CTA's 64 x 64 block ----
  asm volatile("bar.sync 1, %0;\n" ::"n"(GT_CONSUMERS * 32) : "memory");  // every stage consumed, no copy in flight: reuse stage 0
  double* Gs = stage;  // [64][66]
  for (int w = 0; w < GT_CONSUMERS; ++w) {
    if (wid == w) {
      int q = 0;
#pragma unroll
      for (int bi = 0; bi < (SUBW ? 1 : NB8); ++bi)
#pragma unroll
        for (int bj = bi; bj < NB8; ++bj, ++q) {
          double* d = &Gs[(8 * bi + g) * 66 + 8 * bj + 2 * t];
          if (w == 0) { d[0] = acc[q][0]; d[1] = acc[q][1]; }
          else { d[0] += acc[q][0]; d[1] += acc[q][1]; }
        }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(GT_CONSUMERS * 32) : "memory");
  }
  if (SUBW) {
    double* out = P.gram_part + (size_t)blockIdx.x * 512;
    for (int e = tid; e < 512; e += GT_CONSUMERS * 32) {  // e = group * 64 + q * 8 + c  <->  G[q][8 * group + c]
      const int gi = e >> 6, q = (e >> 3) & 7, c = e & 7;
      out[e] = gi < NB8 ? Gs[q * 66 + 8 * gi + c] : 0.0;
    }
    return;
  }
  double* out = P.gram_part + (size_t)blockIdx.x * 4096;
  for (int e = tid; e < 4096; e += GT_CONSUMERS * 32) {
    const int r = e >> 6, c = e & 63;
    double v = 0.0;
    if (r < 8 * NB8 && c < 8 * NB8) v = ((r >> 3) <= (c >> 3)) ? Gs[r * 66 + c] : Gs[c * 66 + r];
    out[e] = v;
  }
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

static void gram_tma_attrs() {
  cudaFuncSetAttribute(k_gram_tma<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
  cudaFuncSetAttribute(k_gram_tma<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
  cudaFuncSetAttribute(k_gram_tma<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
  cudaFuncSetAttribute(k_gram_tma<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
  cudaFuncSetAttribute(k_gram_tma<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
}

// V'V of the current block's clean reflectors on the TMA + DMMA path -> p->gram (64 x 64, zero beyond kpad).  Used by
// the trailing update of tall-skinny matrices instead of its V'V tile (P.no_vtv): one pass over V at HBM speed against
// a full 128-wide DMMA tile per 32 rows.  Returns -1 when Vc cannot be bulk-copied (never: Vc is library-allocated).
extern "C" int qrdm_k_vtv(const qrdm_prob* p, int rows_hint, void* stream) {
  static int attr_gen = -1;
  if (attr_gen != qrdm_rt_device_generation()) {
    gram_tma_attrs();
    attr_gen = qrdm_rt_device_generation();
  }
  int gt = (rows_hint + 1 + GT_TR - 1) / GT_TR;
  if (gt > p->sm_count) gt = p->sm_count;
  if (gt < 1) gt = 1;
  int chunk = (rows_hint + 1 + gt - 1) / gt;
  chunk = (chunk + GT_TR - 1) / GT_TR * GT_TR;
  k_gram_tma<8, 2><<<gt, GT_THREADS, GT_SMEM, (cudaStream_t)stream>>>(*p, chunk);
  QRDM_LAUNCH_CHECK();
  k_gram_reduce<<<64, 64, 0, (cudaStream_t)stream>>>(*p, 1, gt);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// Skinny product of the blocked tall panel on the TMA + DMMA path; *nparts = number of 512-double partials written to
// gram_part (one per CTA, CTAs without rows write zeros).  Returns -1 when the matrix is not 16-byte aligned with an
// even lda (the caller then uses the FMA kernel k_sub_w).
extern "C" int qrdm_k_subw_tma(const qrdm_prob* p, int rows_hint, int* nparts, void* stream) {
  static const char* e_old = getenv("QRDM_SUBW_OLD");  // experiment switch: the round-1 FMA kernel
  if (!p->vec16 || (e_old && atoi(e_old))) return -1;
  static int attr_gen = -1;
  if (attr_gen != qrdm_rt_device_generation()) {
    gram_tma_attrs();
    attr_gen = qrdm_rt_device_generation();
  }
  int gt = (rows_hint + 1 + GT_TR - 1) / GT_TR;
  if (gt > p->sm_count) gt = p->sm_count;
  if (gt < 1) gt = 1;
  int chunk = (rows_hint + 1 + gt - 1) / gt;
  chunk = (chunk + GT_TR - 1) / GT_TR * GT_TR;
  k_gram_tma<8, 1><<<gt, GT_THREADS, GT_SMEM, (cudaStream_t)stream>>>(*p, chunk);
  QRDM_LAUNCH_CHECK();
  *nparts = gt;
  return 0;
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
  static const char* e_old = getenv("QRDM_GRAM_OLD");  // experiment switch: the round-1 FMA kernel
  if (!of_v && p->vec16 && !(e_old && atoi(e_old))) {
    // TMA + DMMA path: one CTA per SM at most, contiguous row ranges that are multiples of the 128-row stage
    static int attr_gen = -1;
    if (attr_gen != qrdm_rt_device_generation()) {
      gram_tma_attrs();
      attr_gen = qrdm_rt_device_generation();
    }
    int gt = (rows_hint + 1 + GT_TR - 1) / GT_TR;  // + 1: the range may start one row early (even alignment)
    if (gt > p->sm_count) gt = p->sm_count;
    if (gt < 1) gt = 1;
    int chunk = (rows_hint + 1 + gt - 1) / gt;
    chunk = (chunk + GT_TR - 1) / GT_TR * GT_TR;
    // nc is only known on the device: the widest variant is always correct (narrower ones are an optimisation the
    // host can take when nb bounds the candidate count)
    const int nbmax = p->nb < QRDM_KMAX ? p->nb : QRDM_KMAX;
    if (nbmax <= 16) k_gram_tma<2, 0><<<gt, GT_THREADS, GT_SMEM, s>>>(*p, chunk);
    else if (nbmax <= 32) k_gram_tma<4, 0><<<gt, GT_THREADS, GT_SMEM, s>>>(*p, chunk);
    else k_gram_tma<8, 0><<<gt, GT_THREADS, GT_SMEM, s>>>(*p, chunk);
    QRDM_LAUNCH_CHECK();
    k_gram_reduce<<<64, 64, 0, s>>>(*p, of_v, gt);
    QRDM_LAUNCH_CHECK();
    return 0;
  }
  k_gram_partial<<<g, 256, 0, s>>>(*p, of_v);
  QRDM_LAUNCH_CHECK();
  k_gram_reduce<<<64, 64, 0, s>>>(*p, of_v, g);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// row-sharded: partial + local reduce into p->gram; the caller all-reduces p->gram (4096 doubles)
extern "C" int qrdm_k_gram_part(const qrdm_prob* p, int rows_hint, void* stream) { return qrdm_k_gram(p, 0, rows_hint, stream); }
