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
  for (int e = tid; e < 4096; e += WS_THREADS) G[e] = (e >> 6) < k ? P.gram[e] : 0.0;
  if (tid < 64) taus[tid] = tid < k ? P.tau[j + tid] : 0.0;
  if (c < nc) {
    for (int i0 = 0; i0 < k; i0 += 8) {  // 8 independent load chains in flight
      double w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = 0.0;
      for (int s = 0; s < splits; ++s) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + u < k) w[u] += P.wp[((size_t)s * 64 + i0 + u) * P.ldw + c];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u < k) ys[(i0 + u) * WS_THREADS + tid] = w[u];
    }
  }
  __syncthreads();
  if (c >= nc) return;
  bool bad = false;
  // blocked forward substitution, 16 reflectors at a time: the contributions of the earlier blocks
  // are 16 independent FMA chains (ILP), only the 16x16 triangle is a dependent recurrence
  for (int b0 = 0; b0 < k; b0 += 16) {
    double a[16], y[16];
#pragma unroll
    for (int ii = 0; ii < 16; ++ii) a[ii] = (b0 + ii < k) ? ys[(b0 + ii) * WS_THREADS + tid] : 0.0;
    for (int s = 0; s < b0; ++s) {
      const double ysv = ys[s * WS_THREADS + tid];
      const double2* grow = reinterpret_cast<const double2*>(G + s * 64 + b0);
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const double2 gv = grow[ii];
        a[2 * ii] = fma(-gv.x, ysv, a[2 * ii]);
        a[2 * ii + 1] = fma(-gv.y, ysv, a[2 * ii + 1]);
      }
    }
#pragma unroll
    for (int ii = 0; ii < 16; ++ii) {
      double acc = a[ii];
#pragma unroll
      for (int s2 = 0; s2 < ii; ++s2) acc = fma(-G[(b0 + s2) * 64 + b0 + ii], y[s2], acc);
      y[ii] = taus[(b0 + ii) & 63] * acc;
    }
#pragma unroll
    for (int ii = 0; ii < 16; ++ii) {
      if (b0 + ii < k) {
        bad |= (y[ii] != y[ii]);
        ys[(b0 + ii) * WS_THREADS + tid] = y[ii];
        P.w2[(size_t)(b0 + ii) * P.ldw + c] = -y[ii];  // negated: k_rankk computes C + V (-T'W)
      }
    }
  }
  for (int i = k; i < kpad; ++i) P.w2[(size_t)i * P.ldw + c] = 0.0;
  if (bad) atomicCAS(&ctrl->err, 0, -13);  // LAPACKE_dlarfb_mia: NaN in C (src/dlarfb.c:73-75)
}

// ------------------------------------------------------------------ k_rankk
// Persistent: one CTA per SM walks a contiguous range of (row block, column tile) units; the
// 128 x k tile of V stays resident in smem while the column tiles of one row block stream by, the
// next W tile arrives by cp.async during the current tile's MMAs, and the next C tile is
// prefetched into registers, so HBM latency is never exposed.  k_wsolve stores -T'W, which lets
// the accumulators be initialised with C itself: C_new = C + V (-T'W) comes straight out of DMMA.
#define RK_BM 128  // rows
#define RK_BN 64   // columns
#define RK_LDV (RK_BM + 4)
#define RK_LDW (RK_BN + 4)
#define RK_SMEM ((64 * RK_LDV + 2 * 64 * RK_LDW) * 8)

template <bool VEC16>
__global__ void __launch_bounds__(256, 1) k_rankk(qrdm_prob P) {
  extern __shared__ __align__(16) double sm[];
  double* Vs = sm;                 // [q][RK_LDV]
  double* Wsb = sm + 64 * RK_LDV;  // 2 x [q][RK_LDW]
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int j = ctrl->j, fjb = ctrl->fjb, k = ctrl->fjb_cmp;
  const int nc = P.n - j - fjb;
  if (nc <= 0 || k <= 0) return;
  const int kpad = (k + 7) & ~7;
  const int jal = j & ~(QRDM_ROWALIGN - 1);
  const int RB = (P.m - jal + RK_BM - 1) / RK_BM, CT = (nc + RK_BN - 1) / RK_BN;
  const long long U = (long long)RB * CT;
  const long long lo = U * blockIdx.x / gridDim.x, hi = U * (blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;
  double* Cg = P.a + (size_t)(j + fjb) * P.lda;
  const int wr = wid & 3, wc = wid >> 2;  // warp tile: rows wr*32.., cols wc*32..
  const size_t lda = (size_t)P.lda;
  // this lane's element (mt, nt, e): column c0 + wc*32 + mt*8 + g, rows R0 + wr*32 + nt*8 + 2t + e
  const size_t lane_off = (size_t)(wc * 32 + g) * lda + (size_t)(wr * 32 + 2 * t);

  auto issue_w = [&](int ct, int buf) {
    const int c0 = ct * RK_BN;
    double* Ws = Wsb + buf * 64 * RK_LDW;
    for (int id = tid; id < kpad * (RK_BN / 2); id += 256) {
      const int q = id / (RK_BN / 2), cp = (id % (RK_BN / 2)) * 2;
      cp_async16(Ws + q * RK_LDW + cp, P.w2 + (size_t)q * P.ldw + c0 + cp, 16);
    }
  };
  auto issue_v = [&](int rb) {
    const int R0 = jal + rb * RK_BM;
    for (int id = tid; id < kpad * (RK_BM / 2); id += 256) {
      const int q = id / (RK_BM / 2), rp = (id % (RK_BM / 2)) * 2;
      cp_async16(Vs + q * RK_LDV + rp, P.vc + (size_t)q * P.ldv + R0 + rp, 16);  // ldv covers the tile
    }
  };
  auto interior = [&](int rb, int ct) {
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    return VEC16 && R0 >= j && R0 + RK_BM <= P.m && c0 + RK_BN <= nc;
  };
  auto load_c = [&](int rb, int ct, double (&dst)[4][4][2]) {
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    const double* base = Cg + (size_t)c0 * lda + R0 + lane_off;
    if (interior(rb, ct)) {  // the common case: 16 unconditional 16-byte loads in flight
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const double2 v = *reinterpret_cast<const double2*>(base + (size_t)(mt * 8) * lda + nt * 8);
          dst[mt][nt][0] = v.x; dst[mt][nt][1] = v.y;
        }
    } else {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int c = c0 + wc * 32 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r = R0 + wr * 32 + nt * 8 + 2 * t;
          const double* ptr = base + (size_t)(mt * 8) * lda + nt * 8;
          dst[mt][nt][0] = (c < nc && r >= j && r < P.m) ? ptr[0] : 0.0;
          dst[mt][nt][1] = (c < nc && r + 1 >= j && r + 1 < P.m) ? ptr[1] : 0.0;
        }
      }
    }
  };
  auto store_c = [&](int rb, int ct, const double (&src)[4][4][2]) {
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    double* base = Cg + (size_t)c0 * lda + R0 + lane_off;
    if (interior(rb, ct)) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<double2*>(base + (size_t)(mt * 8) * lda + nt * 8) = make_double2(src[mt][nt][0], src[mt][nt][1]);
    } else {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int c = c0 + wc * 32 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r = R0 + wr * 32 + nt * 8 + 2 * t;
          double* ptr = base + (size_t)(mt * 8) * lda + nt * 8;
          if (c < nc && r >= j && r < P.m) ptr[0] = src[mt][nt][0];
          if (c < nc && r + 1 >= j && r + 1 < P.m) ptr[1] = src[mt][nt][1];
        }
      }
    }
  };

  int rb = (int)(lo / CT), ct = (int)(lo % CT), buf = 0;
  long long left = hi - lo;
  // one unit: X holds C(u) (prefetched), Y receives C(u+1) while the MMAs of u run
  auto step = [&](double (&X)[4][4][2], double (&Y)[4][4][2]) {
    cp_async_wait<0>();
    __syncthreads();  // W(u) (and V) landed for everyone; everyone is done with W(u-1)
    const bool more = left > 1;
    int nrb = rb, nct = ct + 1;
    if (nct == CT) { nct = 0; ++nrb; }
    if (more && nrb == rb) { issue_w(nct, buf ^ 1); cp_async_commit(); }
    if (more) load_c(nrb, nct, Y);
    const double* Ws = Wsb + buf * 64 * RK_LDW;
    const double* ap = Ws + t * RK_LDW + wc * 32 + g;  // A[m=c][k=q] = W[q][c]
    const double* bp = Vs + t * RK_LDV + wr * 32 + g;  // B[k=q][n=r] = V[r][q]
#pragma unroll 2
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
        for (int nt = 0; nt < 4; ++nt) dmma884(X[mt][nt][0], X[mt][nt][1], a[mt], b[nt]);
    }
    store_c(rb, ct, X);
    if (more && nrb != rb) {
      __syncthreads();  // all warps finished reading the old V tile
      issue_v(nrb);
      issue_w(nct, buf ^ 1);
      cp_async_commit();
    }
    rb = nrb; ct = nct; buf ^= 1; --left;
  };

  double accA[4][4][2], accB[4][4][2];
  issue_v(rb);
  issue_w(ct, 0);
  cp_async_commit();
  load_c(rb, ct, accA);
  while (left > 0) {
    step(accA, accB);
    if (left > 0) step(accB, accA);
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
  // row splits: k_vtc runs one CTA per SM, so pick the split count whose CTA total fills whole
  // waves (and whose chunk ranges are balanced) — wave quantisation cost 14% at 381 CTAs / 148 SMs
  const size_t cap = p->wp_elems / ((size_t)64 * p->ldw);
  int smax = (2 * p->sm_count + ntiles - 1) / ntiles;
  if (smax < 16) smax = 16;
  if (smax > nchunks) smax = nchunks;
  if ((size_t)smax > cap) smax = (int)cap;
  if (smax < 1) smax = 1;
  int splits = 1;
  double best = 0.0;
  for (int sp = 1; sp <= smax; ++sp) {
    const long ctas = (long)ntiles * sp;
    const long waves = (ctas + p->sm_count - 1) / p->sm_count;
    const int cps = (nchunks + sp - 1) / sp;
    const double eff = (double)ctas / (double)(waves * p->sm_count) * (double)nchunks / ((double)cps * sp);
    if (eff > best + 0.015) { best = eff; splits = sp; }
  }

  if (p->vec16) k_vtc<true><<<dim3(ntiles, splits), 256, VT_SMEM, s>>>(*p, splits);
  else k_vtc<false><<<dim3(ntiles, splits), 256, VT_SMEM, s>>>(*p, splits);
  QRDM_LAUNCH_CHECK();
  k_wsolve<<<(ncmax + WS_THREADS - 1) / WS_THREADS, WS_THREADS, WS_SMEM, s>>>(*p, splits);
  QRDM_LAUNCH_CHECK();
  {
    const long long units = (long long)((ncmax + RK_BN - 1) / RK_BN) * ((p->m - jal + RK_BM - 1) / RK_BM);
    const int grid = (int)(units < p->sm_count ? units : p->sm_count);
    if (p->vec16) k_rankk<true><<<grid, 256, RK_SMEM, s>>>(*p);
    else k_rankk<false><<<grid, 256, RK_SMEM, s>>>(*p);
  }
  QRDM_LAUNCH_CHECK();
  return 0;
}
