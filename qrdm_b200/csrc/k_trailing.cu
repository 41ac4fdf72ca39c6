// k_trailing.cu — K6: compact-WY trailing update  C <- (I - V T V')' C = C - V (T' (V' C))
// on C = A[j:m, j+fjb:n], the one dense contraction of the algorithm (the roofline kernel).
//
// Replaces: LAPACKE_dlarfb_mia -> LAPACKE_dlarfb_work -> dlarfb_ (reference src/dlarfb.c:40-151,
// call at src/dgeqrdm_work.c:762-767) and the T factor of LAPACKE_dlarft (:751-754).
//
// Kernels:
//   k_vtc      W_s = V_s' C_s per partial-W slot        DMMA, K = rows, 2-stage cp.async ring
//   k_tinv     T' = (I + D N)^-1 D from V'V (k_vtc's / k_fused's tile 0) and tau, one CTA
//   k_wapply   W2 = -T' (sum_s W_s) per column tile, DMMA; NaN screen of C (-13); optionally the k new R rows
//              of the tile (deferred schedule) and T instead of T' (Q application)
//   k_rankk    C += V W2                                  DMMA, K = k (<= 64) resident in smem; LIST mode = the same
//              update on a gathered column list (eager set of the deferred schedule)
//   k_fused    deferred schedule: pass 2 of the pending block fused into pass 1 of the current one (C read once,
//              written once per iteration); k_colupd / k_w2mask: its small helpers
//   k_vc_build rebuilds the clean V of one block of a factored matrix (Q application, qrdm_b200_dormqr)
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
// Persistent: one CTA per SM walks a contiguous, perfectly balanced range of (column tile,
// 32-row chunk) units; the accumulators are flushed to a partial-W "slot" whenever the walk
// leaves a column tile, so a tile has only 1-3 partials (vs one per row split) and there is no
// wave quantisation.  Tile 0 is V itself: its product is V'V, which k_wsolve needs — the separate
// Gram pass over V is gone.  k_wsolve recomputes the same unit->slot map (vt_* helpers).
#define VT_BN 128
#define VT_BK 32
// smem tile layout: dense [column][32 rows], rows in 16-byte pairs whose index is XOR-swizzled with
// the column parity (bit 2), so the 16-byte fragment loads of a quarter-warp hit 8 distinct banks
#define VT_LD 32
__device__ __forceinline__ int vt_sw(int col, int rowpair_idx) { return col * VT_LD + ((rowpair_idx ^ ((col & 1) << 2)) << 1); }
#define VT_STAGES 2  // two CTAs per SM hide each other's barrier / refill bubbles
#define VT_STAGE_DOUBLES ((64 + VT_BN) * VT_LD)
#define VT_SMEM (VT_STAGES * VT_STAGE_DOUBLES * 8)

struct VtGeom {
  int j, jr, fjb, k, nc, kpad, jal, NCH, CT, voff;  // j: columns done; jr: first active LOCAL row; CT counts the V'V tile
  int T0;           // first tile the unit walk covers: 0 = the V'V tile, 1 when P.no_vtv (V'V comes from qrdm_k_vtv)
  int Tpre;         // k_fused: tiles T >= Tpre already hold the pending update (look-ahead, P.pre_col0): pass 1 only
  long long U;      // units = CT * NCH
  long long Usplit; // = Tpre * NCH (== U without look-ahead): units beyond it cost VT_WB instead of VT_WA
  long long Ctot;   // total cost
  int wb;           // cost of a unit beyond Usplit (VT_WB unless overridden)
};
// Relative cost of a k_fused unit with / without phase A (measured: k_fused 2.08 ms against k_vtc 1.20 ms for the same
// trailing matrix).  The CTAs take contiguous unit ranges of EQUAL COST, so a mix of both kinds stays balanced.
#define VT_WA 7
#define VT_WB 4
__device__ __forceinline__ VtGeom vt_geom(const qrdm_prob& P, int bn = VT_BN) {
  VtGeom g;
  const QrdmGeom q = qrdm_geom(P);
  g.j = q.j; g.fjb = q.fjb; g.k = q.k; g.voff = q.voff;
  g.nc = q.n_end - g.j - g.fjb;
  g.kpad = (g.k + 7) & ~7;
  g.jr = qrdm_jr(P, g.j);
  g.jal = g.jr & ~(QRDM_ROWALIGN - 1);
  const int mpad = (P.m + VT_BK - 1) / VT_BK * VT_BK;
  g.NCH = (mpad - g.jal) / VT_BK;
  g.CT = 1 + (g.nc > 0 ? (g.nc + bn - 1) / bn : 0);
  g.T0 = (P.no_vtv && !P.pend && P.sub == 0 && g.CT > 1) ? 1 : 0;
  g.U = (long long)(g.CT - g.T0) * g.NCH;
  g.Tpre = g.CT;
  if (P.pre_col0 > 0 && !P.pend && P.sub == 0) {  // first tile whose first column is >= pre_col0
    const int d = P.pre_col0 - (g.j + g.fjb);
    const int tp = d <= 0 ? 1 : (d + bn - 1) / bn + 1;
    if (tp < g.CT) g.Tpre = tp;
  }
  g.Usplit = (long long)(g.Tpre - g.T0) * g.NCH;
  g.wb = P.vt_wb > 0 ? P.vt_wb : VT_WB;
  g.Ctot = VT_WA * g.Usplit + g.wb * (g.U - g.Usplit);
  return g;
}
__device__ __forceinline__ long long vt_cost(const VtGeom& ge, long long u) {
  return u <= ge.Usplit ? VT_WA * u : VT_WA * ge.Usplit + ge.wb * (u - ge.Usplit);
}
// first unit of CTA b: the units are cut where the accumulated cost passes b/G of the total (lo(0) = 0, lo(G) = U)
__device__ __forceinline__ long long vt_lo(const VtGeom& ge, int G, int b) {
  const long long x = ge.Ctot * b / G;
  return x <= VT_WA * ge.Usplit ? x / VT_WA : ge.Usplit + (x - VT_WA * ge.Usplit) / ge.wb;
}
// first CTA whose unit range reaches into tile T
__device__ __forceinline__ int vt_bfirst(const VtGeom& ge, int G, int T) {
  const long long X = (long long)(T - ge.T0) * ge.NCH;
  int b = (int)(vt_cost(ge, X) * G / ge.Ctot);
  if (b >= G) b = G - 1;
  while (b > 0 && vt_lo(ge, G, b) > X) --b;
  while (b + 1 < G && vt_lo(ge, G, b + 1) <= X) ++b;
  return b;
}

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

template <bool VEC16, bool FULLK>
__device__ __forceinline__ void vtc_body(const qrdm_prob& P, const VtGeom& ge, int wslot_stride_cols, double* sm) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int G = gridDim.x, b = blockIdx.x;
  const long long lo = vt_lo(ge, G, b), hi = vt_lo(ge, G, b + 1);
  if (lo >= hi) return;
  const int MT = FULLK ? 8 : (ge.kpad >> 3);
  const double* Cg = P.a + (size_t)(ge.j + ge.fjb) * P.lda;

  double acc[8][2][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

  auto issue = [&](long long u, int stage) {
    const int T = ge.T0 + (int)(u / ge.NCH), chunk = (int)(u - (long long)(T - ge.T0) * ge.NCH);
    double* Vs = sm + (size_t)stage * VT_STAGE_DOUBLES;
    double* Cs = Vs + 64 * VT_LD;
    const int r0 = ge.jal + chunk * VT_BK;
    for (int id = tid; id < ge.kpad * 16; id += 256) {  // V: kpad columns x 16 row pairs
      const int q = id >> 4, rp = (id & 15) * 2;
      cp_async16(Vs + vt_sw(q, rp >> 1), P.vc + (size_t)(ge.voff + q) * P.ldv + r0 + rp, 16);
    }
    if (T == 0) {  // the "C" tile is V itself (columns >= kpad read as zero)
      for (int id = tid; id < VT_BN * 16; id += 256) {
        const int c = id >> 4, rp = (id & 15) * 2;
        const bool ok = c < ge.kpad;
        cp_async16(Cs + vt_sw(c, rp >> 1), ok ? P.vc + (size_t)(ge.voff + c) * P.ldv + r0 + rp : P.vc, ok ? 16 : 0);
      }
    } else {
      const int c0 = (T - 1) * VT_BN;
      for (int id = tid; id < VT_BN * 16; id += 256) {
        const int c = id >> 4, rp = (id & 15) * 2;
        const bool ok = c0 + c < ge.nc;
        const double* src = ok ? Cg + (size_t)(c0 + c) * P.lda + r0 + rp : Cg;
        load_rowpair<VEC16>(Cs + vt_sw(c, rp >> 1), src, r0 + rp, P.m, ok);
      }
    }
  };
  auto flush = [&](int T) {
    const int slot = b - vt_bfirst(ge, G, T);
    double* W = P.wp + (size_t)slot * 64 * wslot_stride_cols + (size_t)T * VT_BN;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      if (mt < MT) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int q = mt * 8 + g, c = wid * 16 + nt * 8 + 2 * t;
          *reinterpret_cast<double2*>(W + (size_t)q * wslot_stride_cols + c) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
          acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        }
      }
    }
  };

  const int nmy = (int)(hi - lo);
  issue(lo, 0);
  cp_async_commit();
  int curT = ge.T0 + (int)(lo / ge.NCH);
  for (int it = 0; it < nmy; ++it) {
    const int T = ge.T0 + (int)((lo + it) / ge.NCH);
    if (T != curT) { flush(curT); curT = T; }
    cp_async_wait<0>();
    __syncthreads();
    if (it + 1 < nmy) issue(lo + it + 1, (it + 1) & 1);
    cp_async_commit();
    const double* Vs = sm + (size_t)(it & 1) * VT_STAGE_DOUBLES;
    const double* Cs = Vs + 64 * VT_LD;
    // K is consumed 8 rows at a time: lane t owns rows 2t, 2t+1 of the group, so one 16-byte LDS per
    // column feeds the fragments of two DMMA k-steps (k = {0,2,4,6} then {1,3,5,7}).  One group per
    // trip, NOT unrolled: ptxas otherwise strings all k-steps of an accumulator into a dependent chain.
    const int swz = (g & 1) << 2;
    const double* bp0 = Cs + (wid * 16 + g) * VT_LD;
    const double* bp1 = bp0 + 8 * VT_LD;
    const double* ap = Vs + g * VT_LD;
#pragma unroll 1
    for (int kb = 0; kb < VT_BK / 8; ++kb) {
      const int off = ((kb * 4 + t) ^ swz) << 1;
      const double2 b0 = *reinterpret_cast<const double2*>(bp0 + off);
      const double2 b1 = *reinterpret_cast<const double2*>(bp1 + off);
      double2 a[8];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt)
        if (mt < MT) a[mt] = *reinterpret_cast<const double2*>(ap + mt * 8 * VT_LD + off);
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        if (mt < MT) {
          dmma884(acc[mt][0][0], acc[mt][0][1], a[mt].x, b0.x);
          dmma884(acc[mt][1][0], acc[mt][1][1], a[mt].x, b1.x);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        if (mt < MT) {
          dmma884(acc[mt][0][0], acc[mt][0][1], a[mt].y, b0.y);
          dmma884(acc[mt][1][0], acc[mt][1][1], a[mt].y, b1.y);
        }
      }
    }
  }
  cp_async_wait<0>();
  flush(curT);
}

template <bool VEC16>
__global__ void __launch_bounds__(256, 2) k_vtc(qrdm_prob P, int wslot_stride_cols) {
  extern __shared__ __align__(16) double sm[];
  const VtGeom ge = vt_geom(P);
  if (ge.k <= 0 || ge.nc <= 0 || ge.jr >= P.m) return;
  if (ge.kpad == 64) vtc_body<VEC16, true>(P, ge, wslot_stride_cols, sm);  // the common case: no predicates
  else vtc_body<VEC16, false>(P, ge, wslot_stride_cols, sm);
}

// ------------------------------------------------------------------ k_tinv + k_wapply
// T' W without ever forming T by dlarft's recurrence:  y = T'w satisfies
//   y_i = tau_i (w_i - sum_{s<i} (V'V)[s,i] y_s)   <=>   (I + D N) y = D w,
// N = strictly-lower part of V'V, D = diag(tau).  k_tinv (one CTA) sums the V'V slots and inverts
// the unit lower-triangular I + D N by blocked substitution: M = (I + D N)^-1 D = T'.  A zero tau_i
// simply gives a zero row/column, like dlarft.  k_wapply then computes  W2 = -M (sum_s W_s)  for
// each 128-column tile with DMMA, fusing the slot reduction, the sign (k_rankk adds) and the NaN
// screen of C (any NaN in C poisons its W column): LAPACKE_dlarfb_mia's -13, src/dlarfb.c:73-75.
#define TI_THREADS 1024
#define TI_SMEM (2 * 64 * 65 * 8)

__device__ __forceinline__ int vt_slot_list(const VtGeom& ge, int vt_grid, int T, int* list, int cap) {
  // indices b - bfirst of the CTAs that wrote a partial for tile T, in fixed (ascending) order
  if (T < ge.T0) return 0;
  const int bf = vt_bfirst(ge, vt_grid, T);
  const long long Tend = (long long)(T - ge.T0 + 1) * ge.NCH;
  int n = 0;
  (void)cap;
  for (int b = bf; b < vt_grid && vt_lo(ge, vt_grid, b) < Tend && n < cap; ++b)
    if (vt_lo(ge, vt_grid, b + 1) > vt_lo(ge, vt_grid, b)) list[n++] = b - bf;
  return n;
}

__global__ void __launch_bounds__(TI_THREADS) k_tinv(qrdm_prob P, int vt_grid, int wslot_stride_cols, int bn) {
  extern __shared__ __align__(16) double sm[];
  double* B = sm;            // B[i][s] = tau_i * (V'V)[s][i] for s < i
  double* X = sm + 64 * 65;  // X[i][p]: column p of (I + B)^-1, one thread per column
  __shared__ double taus[64];
  __shared__ int slots[QRDM_PANEL_MAXCTA * 2 + 8];
  __shared__ int nslots;
  const int tid = threadIdx.x;
  const VtGeom ge = vt_geom(P, bn);
  const int k = ge.k;
  if (k <= 0 || ge.nc <= 0) return;
  if (tid == 0) { if (P.w_reduced) { slots[0] = 0; nslots = 1; } else nslots = vt_slot_list(ge, vt_grid, 0, slots, QRDM_PANEL_MAXCTA * 2 + 8); }
  if (tid < 64) taus[tid] = tid < k ? P.tau[ge.j + tid] : 0.0;
  __syncthreads();
  const size_t sstride = (size_t)64 * wslot_stride_cols;
  for (int e = tid; e < 4096; e += TI_THREADS) {
    const int s = e >> 6, i = e & 63;  // (V'V)[s][i], needed for s < i < k
    double g = 0.0;
    if (P.no_vtv && ge.T0 == 1) {  // V'V was formed by qrdm_k_vtv (reduced 64 x 64 block in P.gram; overwritten with M below)
      if (s < i && i < k) g = P.gram[s * 64 + i];
    } else if (s < i && i < k) {
      const double* src = P.wp + (size_t)s * wslot_stride_cols + i;
      for (int q = 0; q < nslots; q += 4) {  // up to 4 slot loads in flight; fixed summation order
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (q + u < nslots) ? src[(size_t)slots[q + u] * sstride] : 0.0;
        g += v[0]; g += v[1]; g += v[2]; g += v[3];
      }
    }
    B[i * 65 + s] = taus[i] * g;
  }
  __syncthreads();
  if (tid < 64) {
    const int p = tid;  // solve (I + B) x = e_p, 16 rows at a time
    for (int b0 = 0; b0 < 64; b0 += 16) {
      double a[16], y[16];
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) a[ii] = (b0 + ii == p) ? 1.0 : 0.0;
      for (int s = 0; s < b0; ++s) {
        const double xs = X[s * 65 + p];
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) a[ii] = fma(-B[(b0 + ii) * 65 + s], xs, a[ii]);
      }
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) {
        double acc = a[ii];
#pragma unroll
        for (int s2 = 0; s2 < ii; ++s2) acc = fma(-B[(b0 + ii) * 65 + b0 + s2], y[s2], acc);
        y[ii] = acc;
      }
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) X[(b0 + ii) * 65 + p] = y[ii];
    }
  }
  __syncthreads();
  // M = (I + B)^-1 D  -> P.gram (row-major 64 x 64; rows/columns >= k are zero)
  for (int e = tid; e < 4096; e += TI_THREADS) {
    const int q = e >> 6, pp = e & 63;
    P.gram[e] = (q < k && pp < k) ? X[q * 65 + pp] * taus[pp] : 0.0;
  }
}

#define WA_LDM 68
#define WA_SMEM_BN(BN) ((2 * 64 * WA_LDM + 64 * ((BN) + 4)) * 8)
#define WA_SMEM WA_SMEM_BN(VT_BN)

// BN = width of the column tiles the partial-W slots were produced with (128: k_vtc, 64: k_fused)
// rows_mode (deferred update): the tile's part of the k new R rows is finished here as well,
//   C[j:j+k, tile] += V[j:j+k, :] W2[:, tile]   (the norm downdate needs them before the next selection;
// everything below row j+k waits for k_fused), and the pending block is recorded in ctrl.
template <int BN>
__global__ void __launch_bounds__(2 * BN) k_wapply(qrdm_prob P, int vt_grid, int wslot_stride_cols, int rows_mode) {
  constexpr int NT = 2 * BN, LDW = BN + 4;
  extern __shared__ __align__(16) double sm[];
  double* Ms = sm;                  // [q][WA_LDM]
  double* Vt = sm + 64 * WA_LDM;    // [r][WA_LDM]: rows j..j+63 of V (rows_mode)
  double* Ws = sm + 2 * 64 * WA_LDM;  // [p][LDW]
  __shared__ int slots[QRDM_PANEL_MAXCTA * 2 + 8];
  __shared__ int nslots;
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const VtGeom ge = vt_geom(P, BN);
  const int c0 = blockIdx.x * BN;
  if ((rows_mode & 1) && blockIdx.x == 0 && tid == 0) { ctrl->pend_k = ge.k; ctrl->pend_c0 = ge.j + ge.fjb; ctrl->pend_r0 = ge.j + ge.k; }
  if (ge.k <= 0 || c0 >= ge.nc) return;
  const int T = blockIdx.x + 1;
  if (tid == 0) { if (P.w_reduced) { slots[0] = 0; nslots = 1; } else nslots = vt_slot_list(ge, vt_grid, T, slots, QRDM_PANEL_MAXCTA * 2 + 8); }
  if (rows_mode & 2) {  // apply Q instead of Q': T = (T')'
    for (int e = tid; e < 4096; e += NT) Ms[(e & 63) * WA_LDM + (e >> 6)] = P.gram[e];
  } else {
    for (int e = tid; e < 4096; e += NT) Ms[(e >> 6) * WA_LDM + (e & 63)] = P.gram[e];
  }
  if (rows_mode & 1)
    for (int e = tid; e < 4096; e += NT) {
      const int q = e >> 6, r = e & 63;  // consecutive threads -> consecutive rows of one column of V
      Vt[r * WA_LDM + q] = (q < ge.k && r < ge.k && ge.j + r < P.m) ? P.vc[(size_t)(ge.voff + q) * P.ldv + ge.j + r] : 0.0;
    }
  __syncthreads();
  const size_t sstride = (size_t)64 * wslot_stride_cols;
  const int ns = nslots;
  for (int e = tid; e < 64 * (BN / 2); e += NT) {  // fixed-order slot sum, 16-byte accesses
    const int pq = e / (BN / 2), cp = (e % (BN / 2)) * 2;
    double2 sacc = make_double2(0.0, 0.0);
    if (pq < ge.kpad) {
      const double* src = P.wp + (size_t)pq * wslot_stride_cols + (size_t)T * BN + cp;
      for (int q = 0; q < ns; q += 4) {  // up to 4 slot loads in flight; fixed summation order
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          v[u] = (q + u < ns) ? *reinterpret_cast<const double2*>(src + (size_t)slots[q + u] * sstride) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 4; ++u) { sacc.x += v[u].x; sacc.y += v[u].y; }
      }
    }
    *reinterpret_cast<double2*>(Ws + pq * LDW + cp) = sacc;
  }
  __syncthreads();
  double acc[8][2][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
  const double* ap = Ms + g * WA_LDM + t;             // A[m=q][k=p] = M[q][p]
  const double* bp = Ws + t * LDW + wid * 16 + g;      // B[k=p][n=c] = W[p][c]
  const int ksteps = ge.kpad >> 2;
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    const double b0 = bp[ks * 4 * LDW], b1 = bp[ks * 4 * LDW + 8];
    double a[8];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) a[mt] = ap[mt * 8 * WA_LDM + ks * 4];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      dmma884(acc[mt][0][0], acc[mt][0][1], a[mt], b0);
      dmma884(acc[mt][1][0], acc[mt][1][1], a[mt], b1);
    }
  }
  bool bad = false;
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    const int q = mt * 8 + g;
    if (q < ge.kpad) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int c = c0 + wid * 16 + nt * 8 + 2 * t;
        const double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
        if (c < ge.nc) bad |= (v0 != v0);
        if (c + 1 < ge.nc) bad |= (v1 != v1);
        *reinterpret_cast<double2*>(P.w2 + (size_t)q * P.ldw + c) = make_double2(-v0, -v1);
      }
    }
  }
  if (bad) atomicCAS(&ctrl->err, 0, -13);
  if (!(rows_mode & 1)) return;
  // ---- the k new R rows of this tile.  Each warp keeps to its own 16 columns of Ws (the ones it read
  // above), so overwriting them with W2 needs no block barrier ----
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int q = mt * 8 + g, c = wid * 16 + nt * 8 + 2 * t;
      const bool ok = q < ge.kpad;
      *reinterpret_cast<double2*>(Ws + q * LDW + c) = make_double2(ok ? -acc[mt][nt][0] : 0.0, ok ? -acc[mt][nt][1] : 0.0);
      acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    }
  __syncwarp();
  const double* ap2 = Vt + g * WA_LDM + t;  // A[m=r][k=q] = V[j+r][q]
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    const double b0 = bp[ks * 4 * LDW], b1 = bp[ks * 4 * LDW + 8];
    double a[8];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) a[mt] = ap2[mt * 8 * WA_LDM + ks * 4];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      dmma884(acc[mt][0][0], acc[mt][0][1], a[mt], b0);
      dmma884(acc[mt][1][0], acc[mt][1][1], a[mt], b1);
    }
  }
  double* Cg = P.a + (size_t)(ge.j + ge.fjb) * P.lda + ge.j;
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    const int r = mt * 8 + g;
    if (r < ge.k && ge.j + r < P.m) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int c = c0 + wid * 16 + nt * 8 + 2 * t;
        if (c < ge.nc) Cg[(size_t)c * P.lda + r] += acc[mt][nt][0];
        if (c + 1 < ge.nc) Cg[(size_t)(c + 1) * P.lda + r] += acc[mt][nt][1];
      }
    }
  }
}

// ------------------------------------------------------------------ k_rankk
// Persistent rank-k update, TWO independent 4-warp CTAs per SM.  What the profiles showed:
//  * a warp can issue one DMMA.8x8x4 every 16 cycles (stall 15 + NOP), exactly the pipe rate, so the
//    two warps a scheduler holds must not do their per-unit bookkeeping at the same time: with one
//    8-warp CTA per SM all warps hit the unit barrier together and the pipe idled 35-40% of the time;
//    two small CTAs drift apart and one's loads/stores/barrier hide under the other's MMAs;
//  * per-unit instruction overhead must be small against the unit's DMMAs: 32 x 32 warp tiles give
//    256 DMMAs per warp per unit and 8 LDS per 16 DMMAs; all slot addressing is hoisted.
// Each CTA walks a contiguous range of (row block, column tile) units with the 128 x k tile of V
// resident in smem, the next W tile arriving by cp.async and the next C tile prefetched into a
// second register set (ping-pong) during the current unit's MMAs.  k_wapply stores -T'W, so the
// accumulators start as C and C_new = C + V (-T'W) goes back to HBM straight from registers.
// Measured (ncu, 16384^2, iteration 10): 93% DMMA-active with the C traffic ablated, 72% with the
// loads, 65% with loads + stores — the same 65% as a 3-stage cp.async ring for C, register prefetch
// with one 8-warp CTA, or LDS.128 interleaved tiles: the strided 1-KB column segments of C, not the
// SM side, bound this kernel.
#define RK_BM 128  // rows
#define RK_BN 32   // columns
#define RK_LDV (RK_BM + 4)
#define RK_LDW (RK_BN + 4)
#define RK_THREADS 128
#define RK_SMEM ((64 * RK_LDV + 2 * 64 * RK_LDW) * 8)

// LIST = true (deferred update, with P.pend): the columns are not a contiguous range but the EAGER SET of the
// pending block — the leading min(64, cols) positions of the new trailing matrix plus the candidates
// k_select just chose, i.e. everything the next Gram / pick / permutation / panel touches.  They are
// completed here (rows >= pend_r0) and stamped in upd_eager so that k_fused / the flush skip them.
template <bool VEC16, bool LIST>
__global__ void __launch_bounds__(RK_THREADS, 2) k_rankk(qrdm_prob P) {
  extern __shared__ __align__(16) double sm[];
  double* Vs = sm;                 // [q][RK_LDV]
  double* Wsb = sm + 64 * RK_LDV;  // 2 x [q][RK_LDW]
  __shared__ int slist[LIST ? 128 : 1];  // absolute column of list entry e, or -1
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5, g = lane >> 2, t = lane & 3;
  const QrdmGeom qg = qrdm_geom(P);
  const int jc = qg.j, fjb = qg.fjb, k = qg.k;
  int nc = qg.n_end - jc - fjb;
  if (LIST) {
    const int jn = ctrl->j, cols = P.n - jn, ncand = ctrl->nc, c0p = jc + fjb;
    int col = -1;
    if (tid < 64) col = tid < cols ? jn + tid : -1;
    else if (tid - 64 < ncand) { const int off = ctrl->cand[tid - 64]; col = off >= 64 ? jn + off : -1; }
    if (col >= 0) {
      if (blockIdx.x == 0) P.upd_eager[col] = P.stamp;
      if (P.upd_flag[col] == P.stamp || col < c0p) col = -1;  // done by k_colupd / leftover panel column
    }
    slist[tid] = col;
    __syncthreads();
    nc = 64 + ncand;
  }
  // side mode (look-ahead, with P.pend): only the columns >= side_col0, and among them only those nobody else owns —
  // a column stamped in upd_eager / upd_flag was completed eagerly and may be permuted or factored by the main stream
  // while this kernel runs, so it is never stored here (not even with its old value)
  const bool side = !LIST && P.side_col0 > 0;
  int coff = 0;
  if (side) { coff = P.side_col0 - (jc + fjb); if (coff < 0) coff = 0; nc -= coff; }
  if (nc <= 0 || k <= 0) return;
  (void)ctrl;
  const int j = qrdm_jr(P, jc);  // first active local row (== jc on a single GPU)
  if (j >= P.m) return;          // row-sharded: this rank has no rows left
  const int kpad = (k + 7) & ~7;
  const int jal = j & ~(QRDM_ROWALIGN - 1);
  const int RB = (P.m - jal + RK_BM - 1) / RK_BM, CT = (nc + RK_BN - 1) / RK_BN;
  const long long U = (long long)RB * CT;
  const long long lo = U * blockIdx.x / gridDim.x, hi = U * (blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;
  double* Cg = P.a + (size_t)(jc + fjb + coff) * P.lda;
  const size_t lda = (size_t)P.lda;
  // warp wr owns rows wr*32..+31 of the tile, all 32 columns: DMMA M = columns (4 tiles), N = rows (4)
  // lane element (mt, nt, e): column c0 + mt*8 + g, rows R0 + wr*32 + nt*8 + 2t + e
  const size_t lane_off = (size_t)g * lda + (size_t)(wr * 32 + 2 * t);
  const int w_q = tid >> 4, w_cp = (tid & 15) * 2;  // W tile slots of this thread: q = w_q + 8i
  const int w_iters = kpad >> 3;

  auto issue_w = [&](int ct, int buf) {
    double* dst = Wsb + buf * 64 * RK_LDW + w_q * RK_LDW + w_cp;
    if (LIST) {  // gathered columns: two 8-byte copies, zero-filled for masked entries
      const int col0 = slist[ct * RK_BN + w_cp], col1 = slist[ct * RK_BN + w_cp + 1];
      const double* s0 = P.w2 + (size_t)w_q * P.ldw + (col0 >= 0 ? col0 - (jc + fjb) : 0);
      const double* s1 = P.w2 + (size_t)w_q * P.ldw + (col1 >= 0 ? col1 - (jc + fjb) : 0);
      for (int i = 0; i < w_iters; ++i) {
        cp_async8(dst + i * 8 * RK_LDW, s0 + (size_t)i * 8 * P.ldw, col0 >= 0 ? 8 : 0);
        cp_async8(dst + i * 8 * RK_LDW + 1, s1 + (size_t)i * 8 * P.ldw, col1 >= 0 ? 8 : 0);
      }
      return;
    }
    const double* src = P.w2 + (size_t)w_q * P.ldw + coff + ct * RK_BN + w_cp;
    if (coff & 1) {  // side mode with an odd column offset: the W2 rows are only 8-byte aligned
      for (int i = 0; i < w_iters; ++i) {
        cp_async8(dst + i * 8 * RK_LDW, src + (size_t)i * 8 * P.ldw, 8);
        cp_async8(dst + i * 8 * RK_LDW + 1, src + (size_t)i * 8 * P.ldw + 1, 8);
      }
      return;
    }
    for (int i = 0; i < w_iters; ++i) cp_async16(dst + i * 8 * RK_LDW, src + (size_t)i * 8 * P.ldw, 16);
  };
  // side mode: bit mt = the lane's column c0 + mt*8 + g may be stored; bit 4 = every column of the tile may (warp-uniform)
  auto keep_mask = [&](int ct) -> int {
    if (!side) return 31;
    int mk = 0;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int c = ct * RK_BN + mt * 8 + g;
      const int col = jc + fjb + coff + c;
      if (c >= nc || (P.upd_eager[col] != P.stamp && P.upd_flag[col] != P.stamp)) mk |= 1 << mt;
    }
    if (__all_sync(0xffffffffu, mk == 15)) mk |= 16;
    return mk;
  };
  auto issue_v = [&](int rb) {
    const int R0 = jal + rb * RK_BM;
    for (int id = tid; id < kpad * (RK_BM / 2); id += RK_THREADS) {
      const int q = id >> 6, rp = (id & 63) * 2;
      cp_async16(Vs + q * RK_LDV + rp, P.vc + (size_t)(qg.voff + q) * P.ldv + R0 + rp, 16);  // ldv covers the tile
    }
  };
  auto interior = [&](int rb, int ct) {
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    return VEC16 && R0 >= j && R0 + RK_BM <= P.m && c0 + RK_BN <= nc;
  };
  // LIST: lane (g, .) works on list entries c0 + mt*8 + g; per-column pointers, rows as usual
  auto list_rw = [&](int rb, int ct, double (&x)[4][4][2], bool store) {
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    const bool rows_in = VEC16 && R0 >= j && R0 + RK_BM <= P.m;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int col = slist[c0 + mt * 8 + g];
      double* base = P.a + (size_t)(col >= 0 ? col : 0) * lda + R0 + wr * 32 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = R0 + wr * 32 + nt * 8 + 2 * t;
        double* ptr = base + nt * 8;
        if (!store) {
          double v0 = 0.0, v1 = 0.0;
          if (col >= 0) {
            if (rows_in) { const double2 v = *reinterpret_cast<const double2*>(ptr); v0 = v.x; v1 = v.y; }
            else {
              if (r >= j && r < P.m) v0 = ptr[0];
              if (r + 1 >= j && r + 1 < P.m) v1 = ptr[1];
            }
          }
          x[mt][nt][0] = v0; x[mt][nt][1] = v1;
        } else if (col >= 0) {
          if (rows_in) *reinterpret_cast<double2*>(ptr) = make_double2(x[mt][nt][0], x[mt][nt][1]);
          else {
            if (r >= j && r < P.m) ptr[0] = x[mt][nt][0];
            if (r + 1 >= j && r + 1 < P.m) ptr[1] = x[mt][nt][1];
          }
        }
      }
    }
  };
  auto load_c = [&](int rb, int ct, double (&dst)[4][4][2]) {
    if (LIST) { list_rw(rb, ct, dst, false); return; }
    if (P.debug & 2) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dst[mt][nt][0] = dst[mt][nt][1] = 0.0;
      return;
    }
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    const double* base = Cg + (size_t)c0 * lda + R0 + lane_off;
    if (interior(rb, ct)) {  // the common case: 16 unconditional 16-byte loads in flight
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const double2* ptr2 = reinterpret_cast<const double2*>(base + (size_t)(mt * 8) * lda + nt * 8);
          const double2 v = *ptr2;  // default caching: __ldcs/__stcs hints measured 3-5% slower here
          dst[mt][nt][0] = v.x; dst[mt][nt][1] = v.y;
        }
    } else {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int c = c0 + mt * 8 + g;
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
  auto store_c = [&](int rb, int ct, double (&src)[4][4][2], int mk) {
    if (LIST) { list_rw(rb, ct, src, true); return; }
    if (P.debug & 1) return;
    const int R0 = jal + rb * RK_BM, c0 = ct * RK_BN;
    double* base = Cg + (size_t)c0 * lda + R0 + lane_off;
    if (interior(rb, ct) && (mk & 16)) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<double2*>(base + (size_t)(mt * 8) * lda + nt * 8) = make_double2(src[mt][nt][0], src[mt][nt][1]);
    } else {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int c = c0 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r = R0 + wr * 32 + nt * 8 + 2 * t;
          double* ptr = base + (size_t)(mt * 8) * lda + nt * 8;
          if (c < nc && ((mk >> mt) & 1) && r >= j && r < P.m) ptr[0] = src[mt][nt][0];
          if (c < nc && ((mk >> mt) & 1) && r + 1 >= j && r + 1 < P.m) ptr[1] = src[mt][nt][1];
        }
      }
    }
  };

  int rb = (int)(lo / CT), ct = (int)(lo - (long long)rb * CT), buf = 0;
  int left = (int)(hi - lo);
  const int a_off = t * RK_LDW + g;             // A[m=c][k=q] = W[q][c]
  const int b_off = t * RK_LDV + wr * 32 + g;   // B[k=q][n=r] = V[r][q]
  // one unit: X holds C(u) (prefetched), Y receives C(u+1) while the MMAs of u run
  auto step = [&](double (&X)[4][4][2], double (&Y)[4][4][2], int& mkX, int& mkY) {
    cp_async_wait<0>();
    __syncthreads();  // W(u) (and V) landed for everyone; everyone is done with W(u-1)
    const bool more = left > 1;
    int nrb = rb, nct = ct + 1;
    if (nct == CT) { nct = 0; ++nrb; }
    if (more && nrb == rb) { issue_w(nct, buf ^ 1); cp_async_commit(); }
    const double* ap = Wsb + buf * 64 * RK_LDW + a_off;
    const double* bp = Vs + b_off;
    auto ksteps = [&](int ks0, int ks1) {
#pragma unroll 2
      for (int ks = ks0; ks < ks1; ++ks) {  // 8 LDS.64 feed 16 independent DMMAs
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
    };
    // The prefetch of C(u+1) is issued only after the first use of X (see k_fused: a wait for X that sits
    // behind freshly issued loads on the same scoreboard costs a full memory latency per unit).
    ksteps(0, 2);
    if (more) { load_c(nrb, nct, Y); mkY = keep_mask(nct); }
    ksteps(2, kpad / 4);
    store_c(rb, ct, X, mkX);
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
  int mkA = keep_mask(ct), mkB = 31;
  while (left > 0) {
    step(accA, accB, mkA, mkB);
    if (left > 0) step(accB, accA, mkB, mkA);
  }
}

// Row-sharded build: fold the partial-W slots of every tile into slot 0 (fixed order) so that ONE
// contiguous [64][stride] block can be all-reduced over NVLink; ranks that have no rows left
// contribute zeros.  k_tinv / k_wapply then read slot 0 only (P.w_reduced).
// Tall-skinny matrices have few column tiles and many CTAs per tile (296 CTAs over <= 4 tiles at n = 512: ~74 slots per
// tile), so the fold is spread over gridDim.y = 16 row groups of the 64 x 128 tile and every thread keeps four slot
// loads in flight (fixed association order ((s0 + s1) + (s2 + s3)), deterministic).  Also used on one GPU when a tile
// has more than 8 slots: k_tinv / k_wapply then read one slot instead of summing 74 serially on 1 + CT CTAs.
#define WR_GY 16
__global__ void __launch_bounds__(256) k_wreduce(qrdm_prob P, int vt_grid, int wslot_stride_cols) {
  __shared__ int slots[QRDM_PANEL_MAXCTA * 2 + 8];
  __shared__ int nslots;
  const VtGeom ge = vt_geom(P);
  const int T = blockIdx.x;
  if (ge.k <= 0 || T >= ge.CT || T < ge.T0) return;
  const bool have_rows = ge.jr < P.m && ge.nc > 0;
  if (threadIdx.x == 0) nslots = have_rows ? vt_slot_list(ge, vt_grid, T, slots, QRDM_PANEL_MAXCTA * 2 + 8) : 0;
  __syncthreads();
  const size_t sstride = (size_t)64 * wslot_stride_cols;
  const int ns = nslots;
  constexpr int QPG = 64 / WR_GY;  // rows of the tile per row group
  for (int e = threadIdx.x; e < QPG * VT_BN; e += 256) {
    const int q = blockIdx.y * QPG + e / VT_BN, c = e % VT_BN;
    double* dst = P.wp + (size_t)q * wslot_stride_cols + (size_t)T * VT_BN + c;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (q < ge.kpad) {
      int i = 0;
      for (; i + 3 < ns; i += 4) {
        const double v0 = dst[(size_t)slots[i] * sstride], v1 = dst[(size_t)slots[i + 1] * sstride];
        const double v2 = dst[(size_t)slots[i + 2] * sstride], v3 = dst[(size_t)slots[i + 3] * sstride];
        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
      }
      for (; i < ns; ++i) s0 += dst[(size_t)slots[i] * sstride];
    }
    dst[0] = (s0 + s1) + (s2 + s3);
  }
}

// ------------------------------------------------------------------ k_fused (deferred update)
// Pass 2 of the PENDING block (iteration i-1) fused into pass 1 of the CURRENT block (iteration i):
//     C_new = C + V_prev W2_prev        (what k_rankk would have done an iteration earlier)
//     W_cur += V_cur' C_new             (what k_vtc does)
// so every element of the trailing matrix is read once and written once per iteration (16 B per
// 4k FLOPs = 16 flop/B at k = 64, against 8 flop/B for k_rankk alone, whose mixed read+write
// stream bound it at 65% DMMA-active).  The columns the next selection touches (leading 64
// positions, the candidates, columns whose norm is recomputed exactly) were brought up to date
// eagerly (k_rankk<., LIST>, k_colupd) and carry the block's stamp in upd_eager / upd_flag: their W2 column reads
// as zero here.  The k new R rows were finished by k_wapply (rows mode); rows < j are never stored.
//
// Two 4-warp CTAs per SM; a unit is a 32-row chunk of a 64-column tile; warp w owns columns
// 16w..16w+15 of the tile.  C goes global -> registers in DMMA fragment layout (prefetched one unit
// ahead, like k_rankk); with M = columns, N = rows in phase A the accumulator fragment of lane
// (g, t) holds rows 2t, 2t+1 of column g, which is exactly the B-operand fragment phase B needs
// when its k-steps take the rows {2t+e}: C_new never leaves the registers between the two MMAs.
// V_prev / V_cur chunks arrive by cp.async (2-stage ring), the W2 tile stays in smem for the whole
// walk down a column tile; partial W goes to the same slot structure as k_vtc (tile width 64).
#define FU_BN 64
#define FU_BK 32
#define FU_LDP 36  // V_prev chunk [q][36]: phase-A B operand, lanes g -> consecutive rows, t -> q stride
#define FU_LDN 40  // V_cur  chunk [q][40]: phase-B A operand, 16-byte row pairs per lane
#define FU_LDW 68  // W2 tile [q][68]
#define FU_THREADS 128
#define FU_STAGE_DOUBLES (64 * FU_LDP + 64 * FU_LDN)
#define FU_SMEM ((64 * FU_LDW + 2 * FU_STAGE_DOUBLES) * 8)

template <bool VEC16, bool FULLK>
__device__ __forceinline__ void fused_body(const qrdm_prob& P, const VtGeom& ge, int wslot_stride_cols, double* sm) {
  double* W2s = sm;                  // [q][FU_LDW]
  double* stage0 = sm + 64 * FU_LDW;
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int G = gridDim.x, b = blockIdx.x;
  const long long lo = vt_lo(ge, G, b), hi = vt_lo(ge, G, b + 1);
  if (lo >= hi) return;
  const int kp_prev = (ctrl->pend_k + 7) & ~7, pend_c0 = ctrl->pend_c0;
  const int pre_c0 = (P.pre_col0 > 0 && P.pre_col0 < P.n) ? P.pre_col0 : P.n;  // look-ahead: columns >= pre_c0 are done
  const int ccol0 = ge.j + ge.fjb;  // first trailing column
  const int jrow = ge.jr;           // first active row
  const size_t lda = (size_t)P.lda;
  double* Cg = P.a + (size_t)ccol0 * lda;
  const double* Vn_g = P.vc + (size_t)ge.voff * P.ldv;
  const int MT = FULLK ? 8 : (ge.kpad >> 3);

  double acc[8][2][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;

  auto issue = [&](int T, int chunk, int stage) {
    double* Vp = stage0 + (size_t)stage * FU_STAGE_DOUBLES;
    double* Vn = Vp + 64 * FU_LDP;
    const int r0 = ge.jal + chunk * FU_BK;
    for (int id = tid; id < ge.kpad * 16; id += FU_THREADS) {
      const int q = id >> 4, rp = (id & 15) * 2;
      cp_async16(Vn + q * FU_LDN + rp, Vn_g + (size_t)q * P.ldv + r0 + rp, 16);
    }
    if (T > 0 && T < ge.Tpre)
      for (int id = tid; id < kp_prev * 16; id += FU_THREADS) {
        const int q = id >> 4, rp = (id & 15) * 2;
        cp_async16(Vp + q * FU_LDP + rp, P.vc_prev + (size_t)q * P.ldv + r0 + rp, 16);
      }
  };
  auto load_w2 = [&](int T) {  // W2 tile of the pending block; zero where the column is already up to date
    const int c0 = (T - 1) * FU_BN;
    for (int e = tid; e < kp_prev * FU_BN; e += FU_THREADS) {
      const int q = e >> 6, c = e & 63;
      const int col = ccol0 + c0 + c, idx = col - pend_c0;
      double v = 0.0;
      if (idx >= 0 && col < pre_c0 && P.upd_eager[col] != P.stamp && P.upd_flag[col] != P.stamp) v = P.w2[(size_t)q * P.ldw + idx];
      W2s[q * FU_LDW + c] = v;
    }
  };
  auto interior = [&](int T, int chunk) {
    const int R0 = ge.jal + chunk * FU_BK, c0 = (T - 1) * FU_BN;
    return VEC16 && T > 0 && R0 >= jrow && R0 + FU_BK <= P.m && c0 + FU_BN <= ge.nc;
  };
  auto load_c = [&](int T, int chunk, double (&dst)[2][4][2]) {
    const int R0 = ge.jal + chunk * FU_BK;
    if (T == 0) {  // the "C" tile is V_cur itself: its product is V'V (k_tinv)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        const int c = wid * 16 + ct * 8 + g;
        const double* base = Vn_g + (size_t)c * P.ldv + R0 + 2 * t;
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) {
          double2 v = make_double2(0.0, 0.0);
          if (c < ge.kpad) v = *reinterpret_cast<const double2*>(base + rt * 8);
          dst[ct][rt][0] = v.x; dst[ct][rt][1] = v.y;
        }
      }
      return;
    }
    const int c0 = (T - 1) * FU_BN;
    const double* base = Cg + (size_t)(c0 + wid * 16 + g) * lda + R0 + 2 * t;
    if (interior(T, chunk)) {
#pragma unroll
      for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) {
          const double2 v = *reinterpret_cast<const double2*>(base + (size_t)(ct * 8) * lda + rt * 8);
          dst[ct][rt][0] = v.x; dst[ct][rt][1] = v.y;
        }
    } else {
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        const int c = c0 + wid * 16 + ct * 8 + g;
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) {
          const int r = R0 + rt * 8 + 2 * t;
          const double* ptr = base + (size_t)(ct * 8) * lda + rt * 8;
          dst[ct][rt][0] = (c < ge.nc && r >= jrow && r < P.m) ? ptr[0] : 0.0;
          dst[ct][rt][1] = (c < ge.nc && r + 1 >= jrow && r + 1 < P.m) ? ptr[1] : 0.0;
        }
      }
    }
  };
  auto store_c = [&](int T, int chunk, const double (&src)[2][4][2]) {
    const int R0 = ge.jal + chunk * FU_BK, c0 = (T - 1) * FU_BN;
    double* base = Cg + (size_t)(c0 + wid * 16 + g) * lda + R0 + 2 * t;
    if (interior(T, chunk)) {
#pragma unroll
      for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int rt = 0; rt < 4; ++rt)
          *reinterpret_cast<double2*>(base + (size_t)(ct * 8) * lda + rt * 8) = make_double2(src[ct][rt][0], src[ct][rt][1]);
    } else {
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        const int c = c0 + wid * 16 + ct * 8 + g;
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) {
          const int r = R0 + rt * 8 + 2 * t;
          double* ptr = base + (size_t)(ct * 8) * lda + rt * 8;
          if (c < ge.nc && r >= jrow && r < P.m) ptr[0] = src[ct][rt][0];
          if (c < ge.nc && r + 1 >= jrow && r + 1 < P.m) ptr[1] = src[ct][rt][1];
        }
      }
    }
  };
  auto flush = [&](int T) {
    const int slot = b - vt_bfirst(ge, G, T);
    double* W = P.wp + (size_t)slot * 64 * wslot_stride_cols + (size_t)T * FU_BN;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      if (mt < MT) {
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          const int q = mt * 8 + g, c = wid * 16 + ct * 8 + 2 * t;
          *reinterpret_cast<double2*>(W + (size_t)q * wslot_stride_cols + c) = make_double2(acc[mt][ct][0], acc[mt][ct][1]);
          acc[mt][ct][0] = acc[mt][ct][1] = 0.0;
        }
      }
    }
  };

  int left = (int)(hi - lo), st = 0;
  int T = (int)(lo / ge.NCH), chunk = (int)(lo - (long long)T * ge.NCH);
  int curT = T;
  if (curT > 0 && curT < ge.Tpre) load_w2(curT);
  issue(T, chunk, 0);
  cp_async_commit();
  const int KSA = FULLK ? 16 : (kp_prev >> 2);  // k-steps of phase A

  // one unit: X holds C(u) (prefetched), Y receives C(u+1) while the MMAs of u run
  auto step = [&](double (&X)[2][4][2], double (&Y)[2][4][2]) {
    cp_async_wait<0>();
    __syncthreads();  // V chunks of u (and the W2 tile) landed; everyone is done with unit u-1
    if (T != curT) {
      flush(curT);
      curT = T;
      if (T < ge.Tpre) {
        load_w2(T);
        __syncthreads();
      }
    }
    int nT = T, nchunk = chunk + 1;
    if (nchunk == ge.NCH) { nchunk = 0; ++nT; }
    if (left > 1) {
      issue(nT, nchunk, st ^ 1);
      cp_async_commit();
    }
    const double* Vp = stage0 + (size_t)st * FU_STAGE_DOUBLES;
    const double* Vn = Vp + 64 * FU_LDP;
    // phase B: acc[q][c] += sum_r V_cur[r][q] X[c][r]     (M = q, N = columns, K = rows {2t+e})
    auto phase_b = [&](bool prefetch_mid) {  // prefetch_mid (a literal): fetch C(u+1) after the first row group
      const double* ap = Vn + g * FU_LDN + 2 * t;
#pragma unroll
      for (int rt = 0; rt < 4; ++rt) {
        if (prefetch_mid && rt == 1 && left > 1) load_c(nT, nchunk, Y);
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // 4 q-tiles at a time
          double2 a[4];
#pragma unroll
          for (int x = 0; x < 4; ++x)
            if (h * 4 + x < MT) a[x] = *reinterpret_cast<const double2*>(ap + (h * 4 + x) * 8 * FU_LDN + rt * 8);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            if (h * 4 + x < MT) {
              dmma884(acc[h * 4 + x][0][0], acc[h * 4 + x][0][1], a[x].x, X[0][rt][0]);
              dmma884(acc[h * 4 + x][1][0], acc[h * 4 + x][1][1], a[x].x, X[1][rt][0]);
            }
          }
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            if (h * 4 + x < MT) {
              dmma884(acc[h * 4 + x][0][0], acc[h * 4 + x][0][1], a[x].y, X[0][rt][1]);
              dmma884(acc[h * 4 + x][1][0], acc[h * 4 + x][1][1], a[x].y, X[1][rt][1]);
            }
          }
        }
      }
    };
    // The prefetch of C(u+1) into Y is issued only AFTER the first use of X, and the V'V tile (T == 0, no
    // phase A) gets its own copy of phase B: ptxas gives the loads of X and of Y the same scoreboard, so a
    // first use of X that sits behind freshly issued loads of Y — which is what the merge point of the two
    // paths looked like to the compiler — waits a full memory latency per unit (ncu: 8-10% of all stall
    // samples on one DMMA).
    if (T > 0 && T < ge.Tpre) {
      // phase A: X[c][r] += sum_q W2[q][c] V_prev[r][q]   (M = columns, N = rows, K = q)
      const double* ap = W2s + t * FU_LDW + wid * 16 + g;
      const double* bp = Vp + t * FU_LDP + g;
      auto kstep = [&](int ks) {
        double a[2], bb[4];
#pragma unroll
        for (int x = 0; x < 2; ++x) a[x] = ap[ks * 4 * FU_LDW + x * 8];
#pragma unroll
        for (int x = 0; x < 4; ++x) bb[x] = bp[ks * 4 * FU_LDP + x * 8];
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
          for (int rt = 0; rt < 4; ++rt) dmma884(X[ct][rt][0], X[ct][rt][1], a[ct], bb[rt]);
      };
      // (fully unrolling the 16 k-steps of a full block makes the kernel 8% SLOWER: 150 KB of code; issuing the
      // prefetch right after the first k-step instead of after the store: 3% slower)
#pragma unroll 4
      for (int ks = 0; ks < KSA; ++ks) kstep(ks);
      store_c(T, chunk, X);
      if (left > 1) load_c(nT, nchunk, Y);
      phase_b(false);
    } else if (T > 0) {
      // look-ahead tile: the pending update was applied by the side stream (k_rankk, side mode) — pass 1 only.
      // Same rule as above: the prefetch goes out after the first use of X, the rest of the MMAs cover its latency.
      phase_b(true);
    } else {
      phase_b(false);
      if (left > 1) load_c(nT, nchunk, Y);
    }
    T = nT; chunk = nchunk; st ^= 1; --left;
  };

  double XA[2][4][2], XB[2][4][2];
  load_c(T, chunk, XA);
  while (left > 0) {
    step(XA, XB);
    if (left > 0) step(XB, XA);
  }
  flush(curT);
}

template <bool VEC16>
__global__ void __launch_bounds__(FU_THREADS, 2) k_fused(qrdm_prob P, int wslot_stride_cols) {
  extern __shared__ __align__(16) double sm[];
  // the geometry lives in shared memory: the body needs most of it only at tile changes, and 255 registers are all taken
  __shared__ VtGeom ge;
  if (threadIdx.x == 0) ge = vt_geom(P, FU_BN);
  __syncthreads();
  if (ge.k <= 0 || ge.nc <= 0 || ge.jr >= P.m) return;
  if (ge.kpad == 64 && P.ctrl->pend_k > 56) fused_body<VEC16, true>(P, ge, wslot_stride_cols, sm);  // the common case: both blocks full, no predicates
  else fused_body<VEC16, false>(P, ge, wslot_stride_cols, sm);
}

// Completion of the pending update (rows >= pend_r0) on the columns whose partial norm is about to be recomputed
// exactly (flag_list of k_norm_update: any length, usually empty), plain FMA.  A finished column gets the block's
// stamp in upd_flag so that nobody applies the update twice.  (The other eager set — leading positions +
// candidates — goes through the DMMA kernel, k_rankk<., LIST>.)
// Thread <-> row (its row of the pending V in registers), 8 list entries per CTA column group.
#define CU_ROWS 128
#define CU_GROUP 8
__global__ void __launch_bounds__(CU_ROWS) k_colupd(qrdm_prob P) {
  __shared__ __align__(16) double Wc[CU_GROUP][64];
  __shared__ int colof[CU_GROUP];
  const qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x;
  const int kprev = ctrl->pend_k, c0p = ctrl->pend_c0, r0p = ctrl->pend_r0;
  const int nlist = ctrl->nflag;
  if (kprev <= 0 || nlist <= 0) return;
  const int ngroups = (nlist + CU_GROUP - 1) / CU_GROUP;
  const int r = r0p + blockIdx.x * CU_ROWS + tid;
  const bool rok = r < P.m;
  double v[64];
  bool vloaded = false;
  for (int gi = blockIdx.y; gi < ngroups; gi += gridDim.y) {
    __syncthreads();
    if (tid < CU_GROUP) {
      const int e = gi * CU_GROUP + tid;
      int col = e < nlist ? P.flag_list[e] : -1;
      if (col >= 0 && blockIdx.x == 0) P.upd_flag[col] = P.stamp;
      if (col >= 0 && col - c0p < 0) col = -1;  // leftover panel column: nothing pending
      colof[tid] = col;
    }
    __syncthreads();
    for (int e = tid; e < CU_GROUP * 64; e += CU_ROWS) {
      const int x = e >> 6, q = e & 63;
      const int col = colof[x];
      Wc[x][q] = (col >= 0 && q < kprev) ? P.w2[(size_t)q * P.ldw + (col - c0p)] : 0.0;
    }
    __syncthreads();
    if (!rok) continue;
    if (!vloaded) {
#pragma unroll
      for (int q = 0; q < 64; ++q) v[q] = q < kprev ? P.vc_prev[(size_t)q * P.ldv + r] : 0.0;
      vloaded = true;
    }
    double s[CU_GROUP];
#pragma unroll
    for (int x = 0; x < CU_GROUP; ++x) s[x] = 0.0;
#pragma unroll
    for (int q = 0; q < 64; q += 2) {  // 8 independent FMA chains, W2 broadcast from smem two q at a time
#pragma unroll
      for (int x = 0; x < CU_GROUP; ++x) {
        const double2 w = *reinterpret_cast<const double2*>(&Wc[x][q]);
        s[x] = fma(v[q], w.x, s[x]);
        s[x] = fma(v[q + 1], w.y, s[x]);
      }
    }
#pragma unroll
    for (int x = 0; x < CU_GROUP; ++x) {
      const int col = colof[x];
      if (col >= 0) P.a[(size_t)col * P.lda + r] += s[x];
    }
  }
}

// flush helper: zero the W2 columns that carry the stamp, so that the plain k_rankk can finish the block
__global__ void __launch_bounds__(256) k_w2mask(qrdm_prob P) {
  const qrdm_ctrl* ctrl = P.ctrl;
  const int c0p = ctrl->pend_c0;
  const int col = c0p + blockIdx.x * 256 + threadIdx.x;
  if (col >= P.n) return;
  if (P.upd_eager[col] == P.stamp || P.upd_flag[col] == P.stamp || (P.pre_col0 > 0 && col >= P.pre_col0))
    for (int q = 0; q < 64; ++q) P.w2[(size_t)q * P.ldw + (col - c0p)] = 0.0;
}

static void trailing_attrs() {
  static int attr_gen = -1;  // per-device attributes, see qrdm_rt_device_generation
  if (attr_gen == qrdm_rt_device_generation()) return;
  cudaFuncSetAttribute(k_vtc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM);
  cudaFuncSetAttribute(k_vtc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, VT_SMEM);
  cudaFuncSetAttribute(k_wapply<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM_BN(128));
  cudaFuncSetAttribute(k_wapply<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM_BN(64));
  cudaFuncSetAttribute(k_tinv, cudaFuncAttributeMaxDynamicSharedMemorySize, TI_SMEM);
  cudaFuncSetAttribute(k_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
  cudaFuncSetAttribute(k_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM);
  cudaFuncSetAttribute(k_rankk<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM);
  cudaFuncSetAttribute(k_rankk<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM);
  cudaFuncSetAttribute(k_rankk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM);
  cudaFuncSetAttribute(k_rankk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM);
  attr_gen = qrdm_rt_device_generation();
}
static int host_jr(const qrdm_prob* p, int j) {
  const int x = j - p->row0;
  return x < 0 ? 0 : (x > p->m ? p->m : x);
}

// pass 1: W slots (and V'V in tile 0)
extern "C" int qrdm_k_vtc_only(const qrdm_prob* p, int j_host, int* stride_out, int* grid_out, void* stream) {
  trailing_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const int ncmax = p->n - j_host - 1;  // fjb >= 1
  *stride_out = 0; *grid_out = 0;
  if (ncmax <= 0) return 0;
  const int jal = host_jr(p, j_host) & ~(QRDM_ROWALIGN - 1);
  const int mpad = (p->m + VT_BK - 1) / VT_BK * VT_BK;
  const int nchunks = (mpad - jal) / VT_BK;
  const int ct_ub = 1 + (ncmax + VT_BN - 1) / VT_BN;  // + the V'V tile
  // persistent grid: two CTAs per SM, never more CTAs than units (device-side nc <= ncmax keeps
  // U >= nchunks * 2 >= grid whenever anything is left to update)
  long long units_lb = (long long)nchunks * 2;
  int vt_grid = 2 * p->sm_count;
  if (vt_grid > units_lb) vt_grid = (int)units_lb;
  if (vt_grid < 1) vt_grid = 1;
  // partial-W slots are laid out [slot][64][stride]; (grid/CT + 2) slots always fit (see host alloc)
  const int stride = ct_ub * VT_BN;
  *stride_out = stride; *grid_out = vt_grid;
  if (nchunks <= 0) return 0;  // row-sharded: no local rows left
  if (p->vec16) k_vtc<true><<<vt_grid, 256, VT_SMEM, s>>>(*p, stride);
  else k_vtc<false><<<vt_grid, 256, VT_SMEM, s>>>(*p, stride);
  QRDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int qrdm_k_wreduce(const qrdm_prob* p, int j_host, int vt_grid, int stride, void* stream) {
  if (stride <= 0) return 0;
  k_wreduce<<<dim3(stride / VT_BN, WR_GY), 256, 0, (cudaStream_t)stream>>>(*p, vt_grid, stride);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// T' and W2 = -T' W from the partial-W slots of k_vtc (bn = 128) or k_fused (bn = 64)
extern "C" int qrdm_k_w2(const qrdm_prob* p, int j_host, int vt_grid, int stride, int bn_and_rows, void* stream) {
  // bit 0: also finish the k new R rows (deferred update); bit 1: use T instead of T' (apply Q, qrdm_b200_dormqr)
  const int bn = bn_and_rows & 0xff0, rows_mode = bn_and_rows & 3;
  trailing_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const int ncmax = p->n - j_host - 1;
  if (ncmax <= 0 || stride <= 0) return 0;
  k_tinv<<<1, TI_THREADS, TI_SMEM, s>>>(*p, vt_grid, stride, bn);
  QRDM_LAUNCH_CHECK();
  if (bn == 64) k_wapply<64><<<(ncmax + 63) / 64, 128, WA_SMEM_BN(64), s>>>(*p, vt_grid, stride, rows_mode);
  else k_wapply<128><<<(ncmax + 127) / 128, 256, WA_SMEM_BN(128), s>>>(*p, vt_grid, stride, rows_mode);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// pass 2 alone: C += V W2 on rows >= j (or, with p->pend, the deferred block on rows >= pend_r0)
extern "C" int qrdm_k_rankk(const qrdm_prob* p, int j_host, void* stream) {
  trailing_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const int ncmax = p->n - j_host - 1;
  if (ncmax <= 0) return 0;
  const int jal = host_jr(p, j_host) & ~(QRDM_ROWALIGN - 1);
  if (p->m - jal > 0) {
    const long long units = (long long)((ncmax + RK_BN - 1) / RK_BN) * ((p->m - jal + RK_BM - 1) / RK_BM);
    const int grid = (int)(units < 2 * p->sm_count ? units : 2 * p->sm_count);
    if (p->vec16) k_rankk<true, false><<<grid, RK_THREADS, RK_SMEM, s>>>(*p);
    else k_rankk<false, false><<<grid, RK_THREADS, RK_SMEM, s>>>(*p);
    QRDM_LAUNCH_CHECK();
  }
  return 0;
}

// T', W2 = -T' W, and pass 2
extern "C" int qrdm_k_trailing_finish(const qrdm_prob* p, int j_host, int vt_grid, int stride, void* stream) {
  const int ncmax = p->n - j_host - 1;
  if (ncmax <= 0 || stride <= 0) return 0;
  const int rc = qrdm_k_w2(p, j_host, vt_grid, stride, VT_BN, stream);
  if (rc) return rc;
  return qrdm_k_rankk(p, j_host, stream);
}

extern "C" int qrdm_k_trailing(const qrdm_prob* p_in, int j_host, void* stream) {
  int stride = 0, grid = 0;
  // Tall-skinny trailing matrices: k_vtc's V'V tile is a full 128-wide DMMA tile per 32 rows of which half is padding —
  // 1.2 of the 5.4 ms of the first iteration of configs[3] (ncu: DMMA pipe 82 % active on 30 % padding + V'V work).
  // There V'V comes from one pass over V at HBM speed (qrdm_k_vtv: TMA + DMMA upper triangle, 0.25 ms) instead.
  qrdm_prob pv = *p_in;
  {
    const int rows = p_in->m - host_jr(p_in, j_host), ncmax = p_in->n - j_host - 1;
    const char* e = getenv("QRDM_B200_NO_VTV");  // experiment switch: 0 keeps the V'V tile everywhere
    pv.no_vtv = (!p_in->sub && !p_in->pend && p_in->nranks == 1 && p_in->vec16 && rows >= 65536 && ncmax <= 1024 && !(e && atoi(e) == 0)) ? 1 : 0;
  }
  const qrdm_prob* p = &pv;
  if (p->no_vtv) {
    const int rcv = qrdm_k_vtv(p, p->m - host_jr(p, j_host), stream);
    if (rcv) return rcv;
  }
  int rc = qrdm_k_vtc_only(p, j_host, &stride, &grid, stream);
  if (rc) return rc;
  // tall-skinny: many CTAs per column tile => fold the partial-W slots in parallel first (see k_wreduce)
  if (stride > 0 && !p->w_reduced && grid > 8 * (stride / VT_BN)) {
    rc = qrdm_k_wreduce(p, j_host, grid, stride, stream);
    if (rc) return rc;
    qrdm_prob p2 = *p;
    p2.w_reduced = 1;
    return qrdm_k_trailing_finish(&p2, j_host, grid, stride, stream);
  }
  return qrdm_k_trailing_finish(p, j_host, grid, stride, stream);
}

// ---- deferred-update launchers ----
extern "C" int qrdm_k_fused(const qrdm_prob* p, int j_host, int* stride_out, int* grid_out, void* stream) {
  trailing_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const int ncmax = p->n - j_host - 1;
  *stride_out = 0; *grid_out = 0;
  if (ncmax <= 0) return 0;
  const int jal = host_jr(p, j_host) & ~(QRDM_ROWALIGN - 1);
  const int mpad = (p->m + FU_BK - 1) / FU_BK * FU_BK;
  const int nchunks = (mpad - jal) / FU_BK;
  const int ct_ub = 1 + (ncmax + FU_BN - 1) / FU_BN;
  long long units_lb = (long long)nchunks * 2;
  int grid = 2 * p->sm_count;
  if (grid > units_lb) grid = (int)units_lb;
  if (grid < 1) grid = 1;
  const int stride = ct_ub * FU_BN;
  *stride_out = stride; *grid_out = grid;
  if (nchunks <= 0) return 0;
  if (p->vec16) k_fused<true><<<grid, FU_THREADS, FU_SMEM, s>>>(*p, stride);
  else k_fused<false><<<grid, FU_THREADS, FU_SMEM, s>>>(*p, stride);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// j_host: first column of the iteration that created the pending block (rows >= j_host + 1 may be active).
// Called in that same iteration, so the block's V is still the "current" buffer p->vc.
extern "C" int qrdm_k_colupd(const qrdm_prob* p, int mode, int j_host, void* stream) {
  trailing_attrs();
  const int rows = p->m - j_host - 1;
  if (rows <= 0) return 0;
  qrdm_prob q = *p;
  if (mode == 0) {  // eager set: DMMA rank-k update on the gathered columns (<= 128 = 4 column tiles)
    q.pend = 1;
    const int jal = host_jr(p, j_host) & ~(QRDM_ROWALIGN - 1);
    const long long units = 4LL * ((p->m - jal + RK_BM - 1) / RK_BM);
    const int grid = (int)(units < 2 * p->sm_count ? units : 2 * p->sm_count);
    if (p->vec16) k_rankk<true, true><<<grid, RK_THREADS, RK_SMEM, (cudaStream_t)stream>>>(q);
    else k_rankk<false, true><<<grid, RK_THREADS, RK_SMEM, (cudaStream_t)stream>>>(q);
  } else {          // flagged-norm list (arbitrary length, usually empty): FMA kernel
    const int gx = (rows + CU_ROWS - 1) / CU_ROWS;
    q.vc_prev = p->vc;
    k_colupd<<<dim3(gx, gx >= 64 ? 4 : 16), CU_ROWS, 0, (cudaStream_t)stream>>>(q);
  }
  QRDM_LAUNCH_CHECK();
  return 0;
}

// Look-ahead: pass 2 of the pending block on the columns >= p->side_col0, launched on the LEAST-priority
// side stream while the main stream runs the next block's Gram / pick / permutation / panel.  Called in the iteration
// that created the pending block (its V is still p->vc), after k_select and the eager completion (all stamps are set).
// Unlike every other launch of k_rankk this one is NOT persistent: ~units_per_cta units (128 rows x 32 columns, ~6.5 us
// each) per CTA, so that CTA slots come free every few microseconds and the block scheduler can hand them to the main
// stream's kernels the moment those are launched.
extern "C" int qrdm_k_side(const qrdm_prob* p, int j_host, int units_per_cta, void* stream) {
  trailing_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  const int ncols = p->n - p->side_col0;
  if (p->side_col0 <= 0 || ncols <= 0) return 0;
  const int jal = host_jr(p, j_host) & ~(QRDM_ROWALIGN - 1);
  if (p->m - jal <= 0) return 0;
  qrdm_prob q = *p;
  q.pend = 1;
  q.pre_col0 = 0;
  const long long units = (long long)((ncols + RK_BN - 1) / RK_BN) * ((p->m - jal + RK_BM - 1) / RK_BM);
  if (units_per_cta < 1) units_per_cta = 1;
  long long grid = (units + units_per_cta - 1) / units_per_cta;
  if (grid < 1) grid = 1;
  if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;
  if (p->vec16) k_rankk<true, false><<<(unsigned)grid, RK_THREADS, RK_SMEM, s>>>(q);
  else k_rankk<false, false><<<(unsigned)grid, RK_THREADS, RK_SMEM, s>>>(q);
  QRDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int qrdm_k_flush(const qrdm_prob* p, int j_host, void* stream) {
  const int ncmax = p->n - j_host - 1;
  if (ncmax <= 0) return 0;
  k_w2mask<<<(ncmax + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  qrdm_prob q = *p;
  q.pend = 1;
  q.vc = p->vc_prev;
  return qrdm_k_rankk(&q, j_host, stream);
}

// ---- Q application (qrdm_b200_dormqr*): the block reflector of columns j0 .. j0+k-1 of an already factored
// matrix is rebuilt as the clean copy Vc (unit diagonal, zeros above, zero-padded to 8 columns) and ctrl is
// pointed at that block, after which the trailing kernels above apply it to any m x n matrix C.
__global__ void __launch_bounds__(256) k_vc_build(qrdm_prob P, const double* af, int ldf, int j0, int k) {
  const int q = blockIdx.y;  // < kpad
  if (blockIdx.x == 0 && q == 0 && threadIdx.x == 0) {
    qrdm_ctrl* c = P.ctrl;
    c->j = j0; c->fjb = k; c->fjb_cmp = k;
  }
  const int jal = j0 & ~(QRDM_ROWALIGN - 1);
  const int r = jal + blockIdx.x * 256 + threadIdx.x;
  if (r >= P.m) return;
  double v = 0.0;
  if (q < k) {
    if (r == j0 + q) v = 1.0;
    else if (r > j0 + q) v = af[(size_t)(j0 + q) * ldf + r];
  }
  P.vc[(size_t)q * P.ldv + r] = v;
}

extern "C" int qrdm_k_vc_build(const qrdm_prob* p, const double* d_af, int ldf, int j0, int k, void* stream) {
  const int kpad = (k + 7) & ~7;
  const int jal = j0 & ~(QRDM_ROWALIGN - 1);
  const int rows = p->m - jal;
  if (rows <= 0 || k <= 0) return 0;
  k_vc_build<<<dim3((rows + 255) / 256, kpad), 256, 0, (cudaStream_t)stream>>>(*p, d_af, ldf, j0, k);
  QRDM_LAUNCH_CHECK();
  return 0;
}
