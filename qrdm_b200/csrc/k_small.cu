// k_small.cu — batched mode (BASELINE.json configs[4]: 8192 Kahan-type 512 x 512 matrices): one CTA
// factors one small matrix from the first norm to the last block, so a batch runs as `batch`
// independent CTAs with no host round trips and no inter-CTA traffic ("independent
// units").  The matrix stays in global memory (2 MB at 512^2: L2-resident while it is being worked
// on); the control block, the partial norms, the Gram/T' matrices and all staging tiles live in
// shared memory, reached through generic pointers so that the selection logic is literally the
// same code as the one-kernel-per-stage path (select_body.cuh, pick_body.cuh).
//
// Per matrix this is the whole loop of reference src/dgeqrdm_work.c:640-787:
//   norms (:672-682) -> [ DM_perm (:280-418) -> permute_marked (:149-262) -> dgeqr2_mia panel
//   (src/dgeqr2.c:148-191) -> dlarft + dlarfb (:751-767) -> norm_update (:36-122) -> stop rule
//   (:782-785) ]*
// Differences in HOW (never in what is decided):
//   * the candidates' cosines are evaluated lazily: first only against candidate 0 (always
//     accepted); candidates that already fail the delta test there can never be accepted and are
//     dropped before the full Gram matrix of the survivors is formed — for Kahan-type matrices that
//     removes the 64 x 64 x rows Gram product from every one of the 511 iterations;
//   * the panel is blocked in 8-column sub-panels held in registers;
//   * the trailing update applies the block's reflectors in order (H_k ... H_1 C, the product the
//     compact-WY form of dlarfb evaluates) with each trailing column held in the registers of one warp:
//     C is read once and written once, no T factor is formed; for blocks of more than 8 reflectors the
//     reflectors stream through shared memory in double-buffered groups of 4.
#include <cstdio>
#include "pick_body.cuh"
#include "select_body.cuh"

#define SM_NT 512
#define SM_NW (SM_NT / 32)
#define SM_MAXDIM 1024  // m, n <= 1024 (the column-in-registers path holds 32 rows per lane)
#define SM_KSMALL 8
#define SM_VPAD 96     // zero rows after V[p][0..1024): r_lo + 32 * NB may overshoot by < 64 + 32
#define SM_LD 66        // doubles per row of the [row][column] staging tiles

#ifdef SM_DEBUG  // per-phase cycle counts of CTA 0 (development aid)
__device__ long long g_pt[12];
#define SM_PS(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_pt[i] += t_ - S.ptq2; S.ptq2 = t_; } } while (0)
#define SM_PS0 do { if (threadIdx.x == 0 && blockIdx.x == 0) S.ptq2 = clock64(); } while (0)
#define SM_PT(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); if ((i) > 0) g_pt[i] += t_ - S.ptq; S.ptq = t_; } } while (0)
#define SM_TDECL long long tph_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tq_ = 0
#define SM_T(i) do { const long long t_ = clock64(); if ((i) > 0) tph_[i] += t_ - tq_; tq_ = t_; } while (0)
#define SM_TPRINT do { if (tid == 0 && blockIdx.x == 0) { printf("cycles: cosines %lld pick %lld permute %lld panel %lld tfactor %lld trailing %lld normupd %lld select %lld\n", tph_[1], tph_[2], tph_[3], tph_[4], tph_[5], tph_[6], tph_[7], tph_[8]); printf("panel: load+dots %lld steps %lld writeback %lld blockupd %lld\n", g_pt[1], g_pt[2], g_pt[3], g_pt[4]); printf("step: scalars %lld update+dots %lld sync1 %lld reduce %lld sync2 %lld\n", g_pt[5], g_pt[6], g_pt[7], g_pt[8], g_pt[9]); } } while (0)
#else
#define SM_PT(i)
#define SM_PS(i)
#define SM_PS0
#define SM_TDECL
#define SM_T(i)
#define SM_TPRINT
#endif

struct SmallArgs {
  int batch, m, n, lda, nb;
  long long stride_a;
  double delta, tau_, eps, eta3;
  double* a;
  int* jpvt;
  double* tau;
  int* ncols;  // [batch][n]; [b][0] holds the stop-rule mode on entry
  int* infos;  // [batch] or NULL
};

struct SmallShared {
  qrdm_ctrl ctrl;
  alignas(16) double vn1[SM_MAXDIM];
  double vn2[SM_MAXDIM];
  alignas(16) double gram[4096];  // candidates' Gram matrix
  alignas(16) double red[SM_NW * 8];   // panel: warp partials of a sub-panel's first reduction
  alignas(16) double part[8 * SM_NT];  // panel: per-thread partial dots, [column][thread]
  double S_[64], rowv[64], wv[64], taus[64];
  double sc[2][4];  // panel: {tau, beta, scale, stop} of the current / next column
  int xcol[64], xdiag[64], ycol[64], ydiag[64];
  int flag[SM_MAXDIM];
  double eta;
  double thres0, inv_scale;  // 5e-14 x the input scale / its reciprocal (1 unless the matrix was pre-scaled)
  long long ptq, ptq2;
  int stop_mode, it, bad, keep[64];
  union alignas(16) {
    SelShared sel;
    PickShared pick;
    struct { double tx[64 * SM_LD], ty[64 * SM_LD]; } xty;
    struct { double v[SM_KSMALL][SM_MAXDIM + SM_VPAD]; } sk;  // reflectors of the current sub-panel, zero outside their rows
    struct { double vbuf[SM_MAXDIM], xold[SM_MAXDIM]; } pan;
  } u;
};

// out[s][t] = sum over rows r_lo <= r < r_hi of X[r][s] * Y[r][t]  (64 x 64, row-major, smem).
// Column q of X is column xcol[q] of A; if xdiag[q] >= 0 it is a Householder vector stored in place:
// implicit 1 at row xdiag[q], zeros above.  `same`: Y == X, only the upper block triangle is
// computed and mirrored.  Two row groups of 256 threads, 4 x 4 accumulators per thread, the next
// 64-row chunk is in flight (registers) while the current one is multiplied out of shared memory.
__device__ __noinline__ void small_xty(const double* __restrict__ a, int lda, int r_lo, int r_hi, const int* xcol, const int* xdiag,
                                       int nx, const int* ycol, const int* ydiag, int ny, bool same, double* out, double* tx,
                                       double* ty) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int grp = tid >> 8, t256 = tid & 255, ty4 = (t256 >> 4) * 4, tx4 = (t256 & 15) * 4;
  const bool active = ty4 < nx && tx4 < ny && !(same && tx4 < ty4);
  double acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
  double preX[8], preY[8];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = wid + SM_NW * q, r = r0 + lane + 32 * h;
        double vx = 0.0, vy = 0.0;
        if (r < r_hi) {
          if (c < nx) { const int d = xdiag[c]; vx = r < d ? 0.0 : (r == d ? 1.0 : a[(size_t)xcol[c] * lda + r]); }
          if (!same && c < ny) { const int d = ydiag[c]; vy = r < d ? 0.0 : (r == d ? 1.0 : a[(size_t)ycol[c] * lda + r]); }
        }
        preX[q * 2 + h] = vx;
        preY[q * 2 + h] = vy;
      }
  };
  const double* tyy = same ? tx : ty;
  if (r_lo < r_hi) fetch(r_lo);
  for (int r0 = r_lo; r0 < r_hi; r0 += 64) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        tx[(lane + 32 * h) * SM_LD + wid + SM_NW * q] = preX[q * 2 + h];
        if (!same) ty[(lane + 32 * h) * SM_LD + wid + SM_NW * q] = preY[q * 2 + h];
      }
    __syncthreads();
    if (r0 + 64 < r_hi) fetch(r0 + 64);
    if (active) {
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int r = grp * 32 + rr;
        const double2 a01 = *reinterpret_cast<const double2*>(&tx[r * SM_LD + ty4]);
        const double2 a23 = *reinterpret_cast<const double2*>(&tx[r * SM_LD + ty4 + 2]);
        const double2 b01 = *reinterpret_cast<const double2*>(&tyy[r * SM_LD + tx4]);
        const double2 b23 = *reinterpret_cast<const double2*>(&tyy[r * SM_LD + tx4 + 2]);
        const double av[4] = {a01.x, a01.y, a23.x, a23.y};
        const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fma(av[p], bv[q], acc[p][q]);
      }
    }
    __syncthreads();
  }
  // combine the two row groups in fixed order (tx doubles as scratch: 64*66 >= 4096)
  if (grp == 1) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) tx[(ty4 + p) * 64 + tx4 + q] = acc[p][q];
  }
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) out[(ty4 + p) * 64 + tx4 + q] = acc[p][q] + tx[(ty4 + p) * 64 + tx4 + q];
  }
  __syncthreads();
  if (same) {
    for (int e = tid; e < 4096; e += SM_NT) {
      const int s = e >> 6, t = e & 63;
      if ((t >> 2) < (s >> 2)) out[e] = out[t * 64 + s];
    }
    __syncthreads();
  }
}

// ---- cosines: lazy evaluation of DM_perm's Gram matrix (src/dgeqrdm_work.c:365-403) ----
__device__ __forceinline__ void small_cosines(const qrdm_prob& P, SmallShared& S) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, nc = ctrl->nc;
  if (nc <= 1) return;
  // (1) dots of candidate 0 with every candidate: warp per candidate
  const double* c0 = P.a + (size_t)(j + ctrl->cand[0]) * P.lda;
  for (int t = wid; t < nc; t += SM_NW) {
    const double* ct = P.a + (size_t)(j + ctrl->cand[t]) * P.lda;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = j + lane;
    for (; r + 96 < P.m; r += 128) {
      s0 = fma(c0[r], ct[r], s0);
      s1 = fma(c0[r + 32], ct[r + 32], s1);
      s2 = fma(c0[r + 64], ct[r + 64], s2);
      s3 = fma(c0[r + 96], ct[r + 96], s3);
    }
    for (; r < P.m; r += 32) s0 = fma(c0[r], ct[r], s0);
    const double d = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) S.S_[t] = d;
  }
  __syncthreads();
  // (2) keep candidate 0 and those whose |cos| against it is < delta (same expression as the pick:
  // G * (1/n_s) * (1/n_t); a NaN cosine is ignored there, so it is kept here)
  if (tid == 0) {
    const double i0 = 1.0 / ctrl->candnrm[0];
    int cnt = 1;
    S.keep[0] = 0;
    for (int t = 1; t < nc; ++t) {
      const double cs = fabs(S.S_[t] * i0 * (1.0 / ctrl->candnrm[t]));
      if (!(cs >= P.delta)) S.keep[cnt++] = t;
    }
    for (int q = 1; q < cnt; ++q) {  // compact in place (keep[q] >= q)
      ctrl->cand[q] = ctrl->cand[S.keep[q]];
      ctrl->candnrm[q] = ctrl->candnrm[S.keep[q]];
    }
    ctrl->nc = cnt;
  }
  __syncthreads();
  const int nc2 = ctrl->nc;
  if (nc2 <= 1) return;
  if (tid < 64) { S.xcol[tid] = tid < nc2 ? j + ctrl->cand[tid] : 0; S.xdiag[tid] = -1; }
  __syncthreads();
  small_xty(P.a, P.lda, j, P.m, S.xcol, S.xdiag, nc2, S.xcol, S.xdiag, nc2, true, S.gram, S.u.xty.tx, S.u.xty.ty);
}

// ---- K3d: rotate the columns along the planned cycles (all rows, all cycles in parallel) ----
__device__ __forceinline__ void small_permute(const qrdm_prob& P) {
  const qrdm_ctrl* ctrl = P.ctrl;
  const int ncyc = ctrl->ncyc, j = ctrl->j;
  const int total = ncyc * P.m;
  for (int e = threadIdx.x; e < total; e += SM_NT) {
    const int cyc = e / P.m, r = e - cyc * P.m;
    const int b = ctrl->cyc_start[cyc], en = ctrl->cyc_start[cyc + 1];
    double* a = P.a + (size_t)j * P.lda + r;
    const double first = a[(size_t)ctrl->cyc_pos[b] * P.lda];
    for (int q = b; q < en - 1; ++q) a[(size_t)ctrl->cyc_pos[q] * P.lda] = a[(size_t)ctrl->cyc_pos[q + 1] * P.lda];
    a[(size_t)ctrl->cyc_pos[en - 1] * P.lda] = first;
  }
}

// Apply `nref` Householder reflectors, in order, to columns [c_lo, c_hi) of the block whose origin
// (row j, column j) is A0: y <- (I - tau_p v_p v_p') y for p = 0..nref-1, which is what the reference's
// unblocked loop (src/dlarf.c) and, mathematically, its dlarfb call do.  One warp holds NC whole columns in
// registers (rows r_lo + lane + 32 q, q < NB), so every column is read once and written once; V[p][r]
// lives in shared memory and is zero outside the reflector's rows (no row predicates in the inner loops).
// Returns true if a dot product came out NaN (the -13 screen of LAPACKE_dlarfb_mia, src/dlarfb.c:73-75).
template <int NB, int NC>
__device__ __noinline__ bool small_apply_seq(double* A0, int lda, int c_lo, int c_hi, int r_lo, int rows,
                                             const double (*V)[SM_MAXDIM + SM_VPAD], const double* taus, int nref) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool bad = false;
  for (int c = c_lo + wid * NC; c < c_hi; c += SM_NW * NC) {
    double y[NC][NB];
#pragma unroll
    for (int u = 0; u < NC; ++u)
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int r = r_lo + lane + 32 * q;
        y[u][q] = (c + u < c_hi && r < rows) ? A0[(size_t)(c + u) * lda + r] : 0.0;
      }
    for (int p = 0; p < nref; ++p) {
      const double* vp = V[p] + r_lo + lane;
      double d[NC][2];
#pragma unroll
      for (int u = 0; u < NC; ++u) d[u][0] = d[u][1] = 0.0;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const double vv = vp[32 * q];
#pragma unroll
        for (int u = 0; u < NC; ++u) d[u][q & 1] = fma(vv, y[u][q], d[u][q & 1]);
      }
      const double tp = taus[p];
      double f[NC];
#pragma unroll
      for (int u = 0; u < NC; ++u) {
        const double t = warp_sum(d[u][0] + d[u][1]);
        bad |= t != t;
        f[u] = tp * t;
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const double vv = vp[32 * q];
#pragma unroll
        for (int u = 0; u < NC; ++u) y[u][q] = fma(-vv, f[u], y[u][q]);
      }
    }
#pragma unroll
    for (int u = 0; u < NC; ++u)
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int r = r_lo + lane + 32 * q;
        if (c + u < c_hi && r < rows) A0[(size_t)(c + u) * lda + r] = y[u][q];
      }
  }
  return bad;
}
__device__ __forceinline__ bool small_apply(double* A0, int lda, int c_lo, int c_hi, int r_lo, int rows,
                                            const double (*V)[SM_MAXDIM + SM_VPAD], const double* taus, int nref) {
  const int nblk = (rows - r_lo + 31) >> 5;
  if (nblk <= 4) return small_apply_seq<4, 2>(A0, lda, c_lo, c_hi, r_lo, rows, V, taus, nref);
  if (nblk <= 8) return small_apply_seq<8, 2>(A0, lda, c_lo, c_hi, r_lo, rows, V, taus, nref);
  if (nblk <= 16) return small_apply_seq<16, 2>(A0, lda, c_lo, c_hi, r_lo, rows, V, taus, nref);
  return small_apply_seq<32, 1>(A0, lda, c_lo, c_hi, r_lo, rows, V, taus, nref);
}

// ---- K4: Householder panel with the DM early stop (src/dgeqr2.c:148-191, src/dlarfg.c:120-185,
// src/dlarf.c:133-185).  Blocked so that the column-by-column part never waits on global memory:
//   * the panel is processed in sub-panels of 8 columns held in REGISTERS (thread t owns rows t and
//     t + 512 of all 8 columns); per column one block-wide reduction delivers, fused, the squared
//     norm of the next pivot column and its dot products with the columns to its right
//     (S_[c] = x_{i+1}' x_c below the next pivot row; rowv[c] = the next pivot row);
//   * the 8 reflectors are then applied to the rest of the panel one column per warp, the column in
//     registers (read once, written once), so later sub-panels start from fully updated columns —
//     exactly the values the reference's unblocked loop would see, hence the same early stop.
// The same for a block of k > 8 reflectors (the trailing update of a wide block): the columns stay in
// registers while the reflectors stream through shared memory in groups of 4, double-buffered — group
// g+1 is fetched from the panel (global/L2) into registers while group g is applied, then stored to the
// other buffer.  FLOPs are the 4*rows*cols*k of the compact-WY form, no T factor, no second pass over C.
template <int NB, int NC>
__device__ __noinline__ bool small_apply_blocked(double* A0, int lda, int c_lo, int c_hi, int rows, int k,
                                                 double (*Vb)[SM_MAXDIM + SM_VPAD], const double* taus) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int SPAN = NB * 32;               // rows covered by the register tile (>= rows)
  constexpr int EPT = (4 * SPAN) / SM_NT;     // staged elements per thread and group
  const int ng = (k + 3) >> 2;
  bool bad = false;
  double pre[EPT];
  auto fetch = [&](int g) {
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
      const int e = tid + SM_NT * t, p = e / SPAN, r = e % SPAN, gp = 4 * g + p;
      pre[t] = (gp < k && r < rows) ? (r < gp ? 0.0 : (r == gp ? 1.0 : A0[(size_t)gp * lda + r])) : 0.0;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
      const int e = tid + SM_NT * t, p = e / SPAN, r = e % SPAN;
      Vb[buf * 4 + p][r] = pre[t];
    }
  };
  for (int cb = c_lo; cb < c_hi; cb += SM_NW * NC) {
    const int c = cb + wid * NC;
    double y[NC][NB];
#pragma unroll
    for (int u = 0; u < NC; ++u)
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int r = lane + 32 * q;
        y[u][q] = (c + u < c_hi && r < rows) ? A0[(size_t)(c + u) * lda + r] : 0.0;
      }
    fetch(0);
    __syncthreads();
    stash(0);
    __syncthreads();
    for (int g = 0; g < ng; ++g) {
      if (g + 1 < ng) fetch(g + 1);
      const int np = min(4, k - 4 * g);
      for (int p = 0; p < np; ++p) {
        const double* vp = Vb[(g & 1) * 4 + p] + lane;
        double d[NC][2];
#pragma unroll
        for (int u = 0; u < NC; ++u) d[u][0] = d[u][1] = 0.0;
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const double vv = vp[32 * q];
#pragma unroll
          for (int u = 0; u < NC; ++u) d[u][q & 1] = fma(vv, y[u][q], d[u][q & 1]);
        }
        const double tp = taus[4 * g + p];
        double f[NC];
#pragma unroll
        for (int u = 0; u < NC; ++u) {
          const double t = warp_sum(d[u][0] + d[u][1]);
          bad |= t != t;
          f[u] = tp * t;
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          const double vv = vp[32 * q];
#pragma unroll
          for (int u = 0; u < NC; ++u) y[u][q] = fma(-vv, f[u], y[u][q]);
        }
      }
      if (g + 1 < ng) stash((g + 1) & 1);
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < NC; ++u)
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int r = lane + 32 * q;
        if (c + u < c_hi && r < rows) A0[(size_t)(c + u) * lda + r] = y[u][q];
      }
  }
  return bad;
}
__device__ __forceinline__ bool small_apply_wide(double* A0, int lda, int c_lo, int c_hi, int rows, int k,
                                                 double (*Vb)[SM_MAXDIM + SM_VPAD], const double* taus) {
  const int nblk = (rows + 31) >> 5;
  if (nblk <= 4) return small_apply_blocked<4, 2>(A0, lda, c_lo, c_hi, rows, k, Vb, taus);
  if (nblk <= 8) return small_apply_blocked<8, 2>(A0, lda, c_lo, c_hi, rows, k, Vb, taus);
  if (nblk <= 16) return small_apply_blocked<16, 2>(A0, lda, c_lo, c_hi, rows, k, Vb, taus);
  return small_apply_blocked<32, 1>(A0, lda, c_lo, c_hi, rows, k, Vb, taus);
}

// dlarfg_mia's scalars for one column (src/dlarfg.c:120-185) + the DM stop test (:129-133), computed by ONE
// thread and broadcast through shared memory: FP64 sqrt/div on every thread would saturate the FP64 pipe.
__device__ __forceinline__ void small_hh_scalars(double alpha, double xn2, int len, bool can_stop, double thres2, double* out) {
  double tau = 0.0, beta = alpha, scale = 1.0, stop = 0.0;
  if (len > 1 && can_stop && xn2 < thres2) {
    stop = 1.0;
  } else if (len > 1 && xn2 != 0.0) {
    const double hy = sqrt(fma(alpha, alpha, xn2));
    beta = (alpha >= 0.0) ? -hy : hy;
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  out[0] = tau; out[1] = beta; out[2] = scale; out[3] = stop;
}
#define SM_PB 8
__device__ __forceinline__ void small_panel(const qrdm_prob& P, SmallShared& S) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, fjb = ctrl->fjb, lda = P.lda;
  const int rows = P.m - j;
  double* Ap = P.a + (size_t)j * lda + j;
  double(*red)[SM_PB] = reinterpret_cast<double(*)[SM_PB]>(S.red);
  double* part = S.part;
  double* Sd = S.S_;    // [2][8] reduced dots, double-buffered by column parity
  double* Rv = S.rowv;  // [2][8] pivot-row entries
  double thres2 = S.thres0 * S.thres0;  // (src/dgeqr2.c:40)^2, 5e-14 x the input scale
  int k = fjb;
  bool stopped = false;
  SM_PT(0);
  for (int s0 = 0; s0 < fjb && !stopped; s0 += SM_PB) {
    const int w = min(SM_PB, fjb - s0);
    SM_PT(0);
    for (int e = tid; e < SM_PB * SM_VPAD; e += SM_NT) S.u.sk.v[e / SM_VPAD][SM_MAXDIM + e % SM_VPAD] = 0.0;
    double x[2][SM_PB];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < SM_PB; ++c) {
        const int r = tid + SM_NT * h;
        x[h][c] = (c < w && r >= s0 && r < rows) ? Ap[(size_t)(s0 + c) * lda + r] : 0.0;
      }
    // dots of the sub-panel's first column with all of its columns (rows below the pivot row s0)
    {
      double acc[SM_PB];
#pragma unroll
      double xm[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = tid + SM_NT * h;
        xm[h] = r > s0 ? x[h][0] : 0.0;
        if (r == s0) {
#pragma unroll
          for (int c = 0; c < SM_PB; ++c) Rv[c] = x[h][c];
        }
      }
#pragma unroll
      for (int c = 0; c < SM_PB; ++c) acc[c] = warp_sum(fma(xm[0], x[0][c], xm[1] * x[1][c]));
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < SM_PB; ++c) red[wid][c] = acc[c];
      }
      __syncthreads();
      if (tid < SM_PB) {
        double t = 0.0;
        for (int q = 0; q < SM_NW; ++q) t += red[q][tid];
        Sd[tid] = t;
        if (tid == 0) small_hh_scalars(Rv[0], t, rows - s0, s0 > 0, thres2, S.sc[0]);
      }
      __syncthreads();
    }
    SM_PT(1);
    int nref = w;  // reflectors this sub-panel ends up producing
#pragma unroll
    for (int i = 0; i < SM_PB; ++i) {
      if (i < w && !stopped) {
        const int gi = s0 + i, cur = (i & 1) * SM_PB, nxt = ((i + 1) & 1) * SM_PB;
        SM_PS0;
        const double tau = S.sc[i & 1][0], beta = S.sc[i & 1][1], scale = S.sc[i & 1][2];
        if (S.sc[i & 1][3] != 0.0) {  // DM early stop: column gi left untouched
          k = gi; nref = i; stopped = true;
        } else {
          if (gi == 0 && fjb > 1 && P.tau_ > 0.0) { const double th = P.tau_ * fabs(beta); thres2 = th * th; }
          if (tid == 0) {
            P.tau[j + gi] = tau;
            S.taus[i] = tau;
            if (tau != tau && ctrl->err == 0) ctrl->err = -8;  // LAPACKE_dlarft's NaN screen of tau
          }
          SM_PS(5);
          double v[2], wc[SM_PB];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = tid + SM_NT * h;
            v[h] = 0.0;
            if (r < rows) {
              if (r > gi) { v[h] = x[h][i]; if (tau != 0.0) { v[h] *= scale; x[h][i] = v[h]; } }
              else if (r == gi) { v[h] = 1.0; x[h][i] = beta; }
            }
            S.u.sk.v[i][r] = v[h];  // zero above the pivot row and below the matrix
          }
#pragma unroll
          for (int c = 0; c < SM_PB; ++c) wc[c] = (c > i && c < w) ? tau * (Rv[cur + c] + Sd[cur + c] * scale) : 0.0;
          if (i + 1 < w) {
            double acc[SM_PB];
#pragma unroll
            for (int c = 0; c < SM_PB; ++c) {
              acc[c] = 0.0;
              if (c > i) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int r = tid + SM_NT * h;
                  x[h][c] = fma(-v[h], wc[c], x[h][c]);  // H_gi applied (rows < gi have v = 0)
                }
              }
            }
            // branch-free: the multiplicand is zeroed for rows at or above the next pivot row
            double xm[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = tid + SM_NT * h;
              xm[h] = r > gi + 1 ? x[h][i + 1 < SM_PB ? i + 1 : i] : 0.0;
              if (r == gi + 1) {  // one thread: the next pivot row
#pragma unroll
                for (int c = 0; c < SM_PB; ++c)
                  if (c > i) Rv[nxt + c] = x[h][c];
              }
            }
#pragma unroll
            for (int c = 0; c < SM_PB; ++c) {
              if (c > i) {
                acc[c] = fma(xm[0], x[0][c], xm[1] * x[1][c]);
                part[c * SM_NT + tid] = acc[c];
              }
            }
            SM_PS(6);
            __syncthreads();
            SM_PS(7);
            if (wid > i && wid < w) {  // warp c totals the 512 partials of column c (fixed order)
              const double* pc = part + wid * SM_NT + lane;
              double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
              for (int q = 0; q < SM_NW; q += 4) {
                t0 += pc[32 * q]; t1 += pc[32 * (q + 1)]; t2 += pc[32 * (q + 2)]; t3 += pc[32 * (q + 3)];
              }
              const double t = warp_sum((t0 + t1) + (t2 + t3));
              if (lane == 0) {
                Sd[nxt + wid] = t;
                if (wid == i + 1) small_hh_scalars(Rv[nxt + wid], t, rows - (gi + 1), true, thres2, S.sc[(i + 1) & 1]);
              }
            }
            SM_PS(8);
            __syncthreads();
            SM_PS(9);
          }
        }
      }
    }
    SM_PT(2);
    // sub-panel back to global memory (rows >= s0; rows above hold R entries of earlier blocks)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < SM_PB; ++c) {
        const int r = tid + SM_NT * h;
        if (c < w && r >= s0 && r < rows) Ap[(size_t)(s0 + c) * lda + r] = x[h][c];
      }
    __syncthreads();  // v[][] and taus[] of this sub-panel complete
    SM_PT(3);
    // apply its nref reflectors to the rest of the panel (columns in registers, read once, written once)
    if (nref > 0 && s0 + w < fjb) small_apply(Ap, lda, s0 + w, fjb, s0, rows, S.u.sk.v, S.taus, nref);
    __syncthreads();
    SM_PT(4);
  }
  if (tid == 0) ctrl->fjb_cmp = k;
}

// ---- K2: partial-norm downdate with the recompute guard (src/dgeqrdm_work.c:36-122) ----
__device__ __forceinline__ void small_norm_update(const qrdm_prob& P, SmallShared& S, double tol3z) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, k = ctrl->fjb_cmp;
  for (int c = j + k + tid; c < P.n; c += SM_NT) {
    const double v1 = S.vn1[c];
    if (v1 == 0.0) continue;
    const double* col = P.a + (size_t)c * P.lda;
    double d0 = 0.0, d1 = 0.0;
    int r = j;
    for (; r + 1 < j + k; r += 2) { d0 = fma(col[r], col[r], d0); d1 = fma(col[r + 1], col[r + 1], d1); }
    if (r < j + k) d0 = fma(col[r], col[r], d0);
    const double d = d0 + d1;
    const double dt = (d * S.inv_scale) * S.inv_scale;  // caller's scale, see qrdm_prob::inv_scale
    double t = sqrt(fabs(dt)) / (v1 * S.inv_scale);
    t = (t + 1.0) * (1.0 - t);
    t = (0.0 >= t) ? 0.0 : t;
    const double q = v1 / S.vn2[c];
    const double t2 = t * (q * q);
    if (t2 <= tol3z) {
      if (P.m - (j + k) > 0) S.flag[atomicAdd(&ctrl->nflag, 1)] = c;
      else { S.vn1[c] = 0.0; S.vn2[c] = 0.0; }
    } else {
      S.vn1[c] = v1 * sqrt(t);
    }
  }
  __syncthreads();
  const int nflag = ctrl->nflag;
  for (int f = wid; f < nflag; f += SM_NW) {  // exact recompute, warp per flagged column
    const int c = S.flag[f];
    const double* col = P.a + (size_t)c * P.lda;
    double s0 = 0.0, s1 = 0.0;
    int r = j + k + lane;
    for (; r + 32 < P.m; r += 64) { s0 = fma(col[r], col[r], s0); s1 = fma(col[r + 32], col[r + 32], s1); }
    if (r < P.m) s0 = fma(col[r], col[r], s0);
    const double v = sqrt(warp_sum(s0 + s1));
    if (lane == 0) { S.vn1[c] = v; S.vn2[c] = v; }
  }
}

// initial norms (src/dgeqrdm_work.c:672-682), jpvt = identity (:596-609 with every column free)
// (the noinline helpers take scalars, not the qrdm_prob: taking its address would move the whole descriptor to local
// memory for the hot loops of the kernel as well)
__device__ __noinline__ void small_init_norms(const double* a, int lda, int m, int n, int* jpvt, SmallShared& S) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  struct { const double* a; int lda, m, n; int* jpvt; } P = {a, lda, m, n, jpvt};
  for (int c = wid; c < P.n; c += SM_NW) {
    const double* col = P.a + (size_t)c * P.lda;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = lane;
    for (; r + 96 < P.m; r += 128) {
      s0 = fma(col[r], col[r], s0);
      s1 = fma(col[r + 32], col[r + 32], s1);
      s2 = fma(col[r + 64], col[r + 64], s2);
      s3 = fma(col[r + 96], col[r + 96], s3);
    }
    for (; r < P.m; r += 32) s0 = fma(col[r], col[r], s0);
    const double v = sqrt(warp_sum((s0 + s1) + (s2 + s3)));
    if (lane == 0) { S.vn1[c] = v; S.vn2[c] = v; P.jpvt[c] = c + 1; }
  }
}

// max |a_ij| of the matrix; if it is finite and its exponent is beyond +-200, multiply the matrix by the power of two
// that brings it into [1, 2) and return that factor (1.0 otherwise).  Every thread returns the same value.
__device__ __noinline__ double small_prescale(double* a, int lda, int m, int n, SmallShared& S) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  struct { double* a; int lda, m, n; } P = {a, lda, m, n};
  double mx = 0.0;
  for (int c = wid; c < P.n; c += SM_NW) {
    const double* col = P.a + (size_t)c * P.lda;
    for (int r = lane; r < P.m; r += 32) mx = fmax(mx, fabs(col[r]));  // fmax drops NaNs (the NaN screen's business)
  }
  mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) S.red[wid] = mx;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < SM_NW; ++w) t = fmax(t, S.red[w]);
    double sc = 1.0;
    if (t > 0.0 && !isinf(t)) {
      const int e = ilogb(t);
      if (e <= -200 || e >= 200) sc = ldexp(1.0, -e);
    }
    S.red[0] = sc;
  }
  __syncthreads();
  const double sc = S.red[0];
  __syncthreads();
  if (sc != 1.0)
    for (int c = wid; c < P.n; c += SM_NW) {
      double* col = P.a + (size_t)c * P.lda;
      for (int r = lane; r < P.m; r += 32) col[r] *= sc;
    }
  return sc;
}

__global__ void __launch_bounds__(SM_NT, 1) k_small(SmallArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallShared& S = *reinterpret_cast<SmallShared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.x;
  if (b >= A.batch) return;

  qrdm_prob P = {};
  P.m = A.m; P.n = A.n; P.lda = A.lda; P.nb = A.nb;
  P.delta = A.delta; P.tau_ = A.tau_;
  P.a = A.a + (size_t)A.stride_a * b;
  P.jpvt = A.jpvt + (size_t)A.n * b;
  P.tau = A.tau + (size_t)min(A.m, A.n) * b;
  P.vn1 = S.vn1; P.vn2 = S.vn2; P.ctrl = &S.ctrl; P.gram = S.gram;
  P.row0 = 0; P.m_glob = A.m; P.nranks = 1; P.sub = 0; P.debug = 0;
  P.thres0 = 5e-14; P.inv_scale = 1.0;
  int* ncols = A.ncols + (size_t)A.n * b;
  const int minmn = min(A.m, A.n);

  for (int i = tid; i < (int)(sizeof(qrdm_ctrl) / sizeof(int)); i += SM_NT) reinterpret_cast<int*>(&S.ctrl)[i] = 0;
  if (tid == 0) {
    const int mode = ncols[0];  // src/dgeqrdm_work.c:531-543
    S.stop_mode = (mode >= 1 && mode <= 3) ? mode : 0;
    S.eta = mode == 1 ? A.eps * A.n : (mode == 2 ? A.eps * sqrt((double)A.n) : (mode == 3 ? A.eta3 : 0.0));
    S.it = 0;
    S.bad = 0;
  }
  small_init_norms(P.a, A.lda, A.m, A.n, P.jpvt, S);
  __syncthreads();
  qrdm_select_body<SM_NT>(P, S.u.sel);
  __syncthreads();
  // badly scaled matrix (largest column norm Inf / NaN, or so small that squares underflow): one exact power-of-two
  // scaling of the whole matrix, undone on the R-like entries at the end (same scheme as the one-matrix driver,
  // dgeqrdm_host.c prescale_input; the reference gets there with cblas_dnrm2 and dlarfg's safmin loop)
  double in_scale = 1.0;
  if (!(S.ctrl.maxnrm <= 0x1p300) || S.ctrl.maxnrm < 0x1p-300) {
    in_scale = small_prescale(P.a, A.lda, A.m, A.n, S);
    if (in_scale != 1.0) {
      __syncthreads();
      for (int i = tid; i < (int)(sizeof(qrdm_ctrl) / sizeof(int)); i += SM_NT) reinterpret_cast<int*>(&S.ctrl)[i] = 0;
      __syncthreads();
      small_init_norms(P.a, A.lda, A.m, A.n, P.jpvt, S);
      __syncthreads();
      qrdm_select_body<SM_NT>(P, S.u.sel);
      __syncthreads();
    }
  }
  if (tid == 0) {
    S.thres0 = 5e-14 * in_scale;  // src/dgeqr2.c:40
    S.inv_scale = 1.0 / in_scale;
    S.eta *= S.ctrl.maxnrm;  // :684
  }
  __syncthreads();

  SM_TDECL;
  while (true) {
    const int j = S.ctrl.j;
    if (j >= minmn) break;
    const int cols = A.n - j;
    SM_T(0); small_cosines(P, S);
    __syncthreads();
    SM_T(1);
    qrdm_pick_body(P, S.u.pick);
    __syncthreads();
    SM_T(2);
    small_permute(P);
    __syncthreads();
    SM_T(3);
    small_panel(P, S);
    __syncthreads();
    SM_T(4);
    const int fjb = S.ctrl.fjb, k = S.ctrl.fjb_cmp;
    if (k > 0 && A.n - j - fjb > 0) {
      if (k <= SM_KSMALL) {
        // narrow block: its reflectors are still in shared memory (sub-panel 0 of the panel) -> apply them
        // in order, one trailing column per warp in registers; no T factor needed
        if (small_apply(P.a + (size_t)j * A.lda + j, A.lda, fjb, A.n - j, 0, A.m - j, S.u.sk.v, S.taus, k)) S.bad = 1;
        SM_T(5);
      } else {
        // wide block: all k taus, then the reflectors stream through shared memory in groups of 4
        if (tid < 64) S.taus[tid] = tid < k ? P.tau[j + tid] : 0.0;
        __syncthreads();
        SM_T(5);
        if (small_apply_wide(P.a + (size_t)j * A.lda + j, A.lda, fjb, A.n - j, A.m - j, k, S.u.sk.v, S.taus)) S.bad = 1;
      }
      __syncthreads();
      if (tid == 0 && S.bad && S.ctrl.err == 0) S.ctrl.err = -13;
    }
    __syncthreads();
    SM_T(6);
    small_norm_update(P, S, 1.0536712127723509e-08 /* sqrt(dlamch('e')), src/dgeqrdm_work.c:528-529 */);
    __syncthreads();
    SM_T(7);
    qrdm_select_body<SM_NT>(P, S.u.sel);  // next iteration's prologue: j += k, max norm, candidates
    __syncthreads();
    SM_T(8);
    const int kk = S.ctrl.last_k;
    if (tid == 0) ncols[S.it++] = kk;  // :740
    if (S.ctrl.err != 0) break;
    if (kk <= 0) { if (tid == 0) S.ctrl.err = QRDM_ERR_INTERNAL; break; }
    if (S.stop_mode && S.ctrl.maxnrm * sqrt((double)(cols - kk)) <= S.eta) break;  // :782-785
  }
  __syncthreads();
  if (in_scale != 1.0) {  // R back to the caller's scale: rows <= c of the columns c < rank, whole columns c >= rank
    const int rk = S.ctrl.j;
    const double inv = 1.0 / in_scale;
    for (int c = wid; c < A.n; c += SM_NW) {
      double* col = P.a + (size_t)c * A.lda;
      const int hi = c < rk ? min(c + 1, A.m) : A.m;
      for (int r = lane; r < hi; r += 32) col[r] *= inv;
    }
  }
  SM_TPRINT;
  if (tid == 0 && A.infos) A.infos[b] = S.ctrl.err;
}

extern "C" int qrdm_k_small_supported(int m, int n) { return m <= SM_MAXDIM && n <= SM_MAXDIM; }

extern "C" int qrdm_k_small(int batch, int m, int n, double* d_a, int lda, long long stride_a, int* d_jpvt, double* d_tau,
                            int* d_ncols, int* d_infos, double delta, double tau_, double eta3, int nb, void* stream) {
  static int attr_gen = -1;  // per-device attribute, see qrdm_rt_device_generation
  if (attr_gen != qrdm_rt_device_generation()) {
    cudaError_t e = cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
    if (e != cudaSuccess) return (int)e;
    attr_gen = qrdm_rt_device_generation();
  }
  SmallArgs A;
  A.batch = batch; A.m = m; A.n = n; A.lda = lda; A.nb = nb;
  A.stride_a = stride_a;
  A.delta = delta; A.tau_ = tau_; A.eps = 1.1102230246251565e-16 /* dlamch('e') */; A.eta3 = eta3;
  A.a = d_a; A.jpvt = d_jpvt; A.tau = d_tau; A.ncols = d_ncols; A.infos = d_infos;
  k_small<<<batch, SM_NT, sizeof(SmallShared), (cudaStream_t)stream>>>(A);
  QRDM_LAUNCH_CHECK();
  return 0;
}
