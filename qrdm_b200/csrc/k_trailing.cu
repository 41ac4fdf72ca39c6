// k_trailing.cu — K6: compact-WY trailing update  C <- (I - V T V')' C = C - V (T' (V' C))
// on C = A[j:m, j+fjb:n], the one dense contraction of the algorithm (the roofline kernel).
//
// Replaces: LAPACKE_dlarfb_mia -> LAPACKE_dlarfb_work -> dlarfb_ (reference src/dlarfb.c:40-151,
// call at src/dgeqrdm_work.c:762-767) and the T factor of LAPACKE_dlarft (:751-754).
//
// Three kernels:
//   k_vtc     W_s = V_s' C_s for row split s       DMMA, K = rows, 3-stage cp.async pipeline
//   k_wsolve  y = T' (sum_s W_s) per column by forward substitution with V'V and tau:
//             T^-1 = striu(V'V) + diag(1/tau)  =>  y_i = tau_i (w_i - sum_{s<i} (V'V)[s,i] y_s)
//             (tau_i = 0 gives y_i = 0 = dlarft's zero column).  Also the NaN screen of C (-13).
//   k_rankk   C -= V y                              DMMA, K = k (<= 64) resident in smem
// FP64 math is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): measured 37.0 TFLOP/s = the B200 FP64 peak,
// vs 33.5-34 for a DFMA loop (profiles/r01_fp64_peak_microbench.txt); tcgen05/wgmma have no FP64
// kind.  Algorithmic FLOPs 4*m_r*n_c*k; minimum HBM traffic 24*m_r*n_c bytes (C read twice,
// written once) -> 10.7 flop/B at k = 64, compute-bound; HBM-bound for small k.
//
// Row tiles start at multiples of QRDM_ROWALIGN in GLOBAL row numbering (so cp.async sources are
// 16-byte aligned whenever lda is even); Vc is zero in rows [j_aligned, j) and >= m, which makes
// those rows contribute nothing and leaves the R rows above the block untouched.
#include "common.cuh"

// ------------------------------------------------------------------ k_vtc
#define VT_BN 128
#define VT_BK 32
#define VT_LD 36  // == 4 (mod 16): conflict-free DMMA fragment loads
#define VT_STAGES 3
#define VT_STAGE_DOUBLES ((64 + VT_BN) * VT_LD)
#define VT_SMEM (VT_STAGES * VT_STAGE_DOUBLES * 8)

template <bool VEC16>
__device__ __forceinline__ void load_rowpair(double* dst, const double* src, int r, int m, bool col_ok) {
  // copies rows r, r+1 (r even) of one column into smem, zero-filling rows >= m / masked columns
  if (VEC16) {
    int bytes = col_ok ? (m - r) * 8 : 0;
    bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
    cp_async16(dst, src, bytes);
  } else {
    cp_async8(dst, src, (col_ok && r < m) ? 8 : 0);
    cp_async8(dst + 1, src + 1, (col_ok && r + 1 < m) ? 8 : 0);
  }
}

template <bool VEC16>
__global__ void __launch_bounds__(256, 1) k_vtc(qrdm_prob P, int splits) {
  extern __shared__ __align__(16) double sm[];
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int j = ctrl->j, fjb = ctrl->fjb, k = ctrl->fjb_cmp;
  const int nc = P.n - j - fjb;
  const int c0 = blockIdx.x * VT_BN;
  if (c0 >= nc || k <= 0) return;
  const int kpad = (k + 7) & ~7, MT = kpad >> 3;
  const int jal = j & ~(QRDM_ROWALIGN - 1);
  const int mpad = (P.m + VT_BK - 1) / VT_BK * VT_BK;
  const int nchunks = (mpad - jal) / VT_BK;
  const int cps = (nchunks + splits - 1) / splits;
  const int ch_lo = blockIdx.y * cps, ch_hi = min(nchunks, ch_lo + cps);
  const double* Cg = P.a + (size_t)(j + fjb) * P.lda;

  double acc[8][2][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  auto issue = [&](int chunk, int stage) {
    double* Vs = sm + (size_t)stage * VT_STAGE_DOUBLES;
    double* Cs = Vs + 64 * VT_LD;
    const int r0 = jal + chunk * VT_BK;
    // V: kpad columns x 16 row pairs
    for (int id = tid; id < kpad * 16; id += 256) {
      const int q = id >> 4, rp = (id & 15) * 2;
      cp_async16(Vs + q * VT_LD + rp, P.vc + (size_t)q * P.ldv + r0 + rp, 16);
    }
    for (int id = tid; id < VT_BN * 16; id += 256) {
      const int c = id >> 4, rp = (id & 15) * 2;
      const bool ok = c0 + c < nc;
      const double* src = ok ? Cg + (size_t)(c0 + c) * P.lda + r0 + rp : Cg;
      load_rowpair<VEC16>(Cs + c * VT_LD + rp, src, r0 + rp, P.m, ok);
    }
  };

  const int nmy = ch_hi - ch_lo;
#pragma unroll
  for (int s = 0; s < VT_STAGES - 1; ++s) {
    if (s < nmy) issue(ch_lo + s, s);
    cp_async_commit();
  }
  for (int it = 0; it < nmy; ++it) {
    cp_async_wait<VT_STAGES - 2>();
    __syncthreads();
    const int nxt = it + VT_STAGES - 1;
    if (nxt < nmy) issue(ch_lo + nxt, nxt % VT_STAGES);
    cp_async_commit();
    const double* Vs = sm + (size_t)(it % VT_STAGES) * VT_STAGE_DOUBLES;
    const double* Cs = Vs + 64 * VT_LD;
    const double* bp0 = Cs + (wid * 16 + g) * VT_LD + t;
    const double* bp1 = bp0 + 8 * VT_LD;
    const double* ap = Vs + g * VT_LD + t;
#pragma unroll
    for (int ks = 0; ks < VT_BK / 4; ++ks) {
      const double b0 = bp0[ks * 4], b1 = bp1[ks * 4];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        if (mt < MT) {
          const double a = ap[mt * 8 * VT_LD + ks * 4];
          dmma884(acc[mt][0][0], acc[mt][0][1], a, b0);
          dmma884(acc[mt][1][0], acc[mt][1][1], a, b1);
        }
      }
    }
  }
  cp_async_wait<0>();
  // store the partial W (row-major [split][q][ldw]); columns >= nc are never read
  double* W = P.wp + (size_t)blockIdx.y * 64 * P.ldw;
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    if (mt < MT) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int q = mt * 8 + g, c = c0 + wid * 16 + nt * 8 + 2 * t;
        if (c < nc) *reinterpret_cast<double2*>(W + (size_t)q * P.ldw + c) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      }
    }
  }
}

// ------------------------------------------------------------------ k_wsolve
#define WS_THREADS 128
#define WS_SMEM ((64 * 64 + 64 + 64 * WS_THREADS) * 8)

__global__ void __launch_bounds__(WS_THREADS) k_wsolve(qrdm_prob P, int splits) {
  extern __shared__ __align__(16) double sm[];
  double* G = sm;                 // V'V, [s*64 + i]
  double* taus = sm + 4096;       // 64
  double* ys = sm + 4096 + 64;    // [i][thread]
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x;
  const int j = ctrl->j, fjb = ctrl->fjb, k = ctrl->fjb_cmp;
  const int nc = P.n - j - fjb;
  const int c = blockIdx.x * WS_THREADS + tid;
  if (blockIdx.x * WS_THREADS >= nc || k <= 0) return;
  const int kpad = (k + 7) & ~7;
  for (int e = tid; e < k * 64; e += WS_THREADS) G[e] = P.gram[e];
  if (tid < 64) taus[tid] = tid < k ? P.tau[j + tid] : 0.0;
  if (c < nc) {
    for (int i = 0; i < k; ++i) {
      double w = 0.0;
      for (int s = 0; s < splits; ++s) w += P.wp[((size_t)s * 64 + i) * P.ldw + c];
      ys[i * WS_THREADS + tid] = w;
    }
  }
  __syncthreads();
  if (c >= nc) return;
  bool bad = false;
  for (int i = 0; i < k; ++i) {
    double a0 = ys[i * WS_THREADS + tid], a1 = 0.0;
    int s = 0;
    for (; s + 1 < i; s += 2) {
      a0 = fma(-G[s * 64 + i], ys[s * WS_THREADS + tid], a0);
      a1 = fma(-G[(s + 1) * 64 + i], ys[(s + 1) * WS_THREADS + tid], a1);
    }
    if (s < i) a0 = fma(-G[s * 64 + i], ys[s * WS_THREADS + tid], a0);
    const double y = taus[i] * (a0 + a1);
    bad |= (y != y);
    ys[i * WS_THREADS + tid] = y;
    P.w2[(size_t)i * P.ldw + c] = y;
  }
  for (int i = k; i < kpad; ++i) P.w2[(size_t)i * P.ldw + c] = 0.0;
  if (bad) atomicCAS(&ctrl->err, 0, -13);  // LAPACKE_dlarfb_mia: NaN in C (src/dlarfb.c:73-75)
}

// ------------------------------------------------------------------ k_rankk
#define RK_BM 128  // rows
#define RK_BN 64   // columns
#define RK_LDV (RK_BM + 4)
#define RK_LDW (RK_BN + 4)
#define RK_SMEM ((64 * RK_LDV + 64 * RK_LDW) * 8)

template <bool VEC16>
__global__ void __launch_bounds__(256, 2) k_rankk(qrdm_prob P) {
  extern __shared__ __align__(16) double sm[];
  double* Vs = sm;                 // [q][RK_LDV]
  double* Ws = sm + 64 * RK_LDV;   // [q][RK_LDW]
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int j = ctrl->j, fjb = ctrl->fjb, k = ctrl->fjb_cmp;
  const int nc = P.n - j - fjb;
  const int c0 = blockIdx.x * RK_BN;
  const int jal = j & ~(QRDM_ROWALIGN - 1);
  const int R0 = jal + blockIdx.y * RK_BM;
  if (c0 >= nc || R0 >= P.m || k <= 0) return;
  const int kpad = (k + 7) & ~7;
  double* Cg = P.a + (size_t)(j + fjb) * P.lda;

  for (int id = tid; id < kpad * (RK_BM / 2); id += 256) {
    const int q = id / (RK_BM / 2), rp = (id % (RK_BM / 2)) * 2;
    cp_async16(Vs + q * RK_LDV + rp, P.vc + (size_t)q * P.ldv + R0 + rp, 16);  // ldv covers the tile
  }
  for (int id = tid; id < kpad * (RK_BN / 2); id += 256) {
    const int q = id / (RK_BN / 2), cp = (id % (RK_BN / 2)) * 2;
    cp_async16(Ws + q * RK_LDW + cp, P.w2 + (size_t)q * P.ldw + c0 + cp, 16);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int wr = wid & 3, wc = wid >> 2;  // warp tile: rows wr*32.., cols wc*32..
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const double* ap = Ws + t * RK_LDW + wc * 32 + g;   // A[m=c][k=q]
  const double* bp = Vs + t * RK_LDV + wr * 32 + g;   // B[k=q][n=r]
  for (int ks = 0; ks < kpad / 4; ++ks) {
    double a[4], b[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      a[x] = ap[ks * 4 * RK_LDW + x * 8];
      b[x] = bp[ks * 4 * RK_LDV + x * 8];
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
  // C[r, c] -= acc ; lane holds (c = ..+g, r = ..+2t, 2t+1)
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int c = c0 + wc * 32 + mt * 8 + g;
    if (c >= nc) continue;
    double* col = Cg + (size_t)c * P.lda;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r = R0 + wr * 32 + nt * 8 + 2 * t;
      if (VEC16) {
        if (r >= j && r + 1 < P.m) {
          double2 v = *reinterpret_cast<double2*>(col + r);
          v.x -= acc[mt][nt][0];
          v.y -= acc[mt][nt][1];
          *reinterpret_cast<double2*>(col + r) = v;
          continue;
        }
      }
      if (r >= j && r < P.m) col[r] -= acc[mt][nt][0];
      if (r + 1 >= j && r + 1 < P.m) col[r + 1] -= acc[mt][nt][1];
    }
  }
}

extern "C" int qrdm_k_trailing(const qrdm_prob* p, int j_host, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_vtc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM);
    cudaFuncSetAttribute(k_vtc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM);
    cudaFuncSetAttribute(k_wsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
    cudaFuncSetAttribute(k_rankk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM);
    cudaFuncSetAttribute(k_rankk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM);
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int ncmax = p->n - j_host - 1;  // fjb >= 1
  if (ncmax <= 0) return 0;
  const int jal = j_host & ~(QRDM_ROWALIGN - 1);
  const int mpad = (p->m + VT_BK - 1) / VT_BK * VT_BK;
  const int nchunks = (mpad - jal) / VT_BK;
  const int ntiles = (ncmax + VT_BN - 1) / VT_BN;
  int splits = (2 * p->sm_count + ntiles - 1) / ntiles;
  if (splits > nchunks) splits = nchunks;
  const size_t cap = p->wp_elems / ((size_t)64 * p->ldw);
  if ((size_t)splits > cap) splits = (int)cap;
  if (splits < 1) splits = 1;

  if (p->vec16) k_vtc<true><<<dim3(ntiles, splits), 256, VT_SMEM, s>>>(*p, splits);
  else k_vtc<false><<<dim3(ntiles, splits), 256, VT_SMEM, s>>>(*p, splits);
  QRDM_LAUNCH_CHECK();
  k_wsolve<<<(ncmax + WS_THREADS - 1) / WS_THREADS, WS_THREADS, WS_SMEM, s>>>(*p, splits);
  QRDM_LAUNCH_CHECK();
  dim3 grid((ncmax + RK_BN - 1) / RK_BN, (p->m - jal + RK_BM - 1) / RK_BM);
  if (p->vec16) k_rankk<true><<<grid, 256, RK_SMEM, s>>>(*p);
  else k_rankk<false><<<grid, 256, RK_SMEM, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
