// k_panel.cu — K4: Householder panel factorisation with the Deviation-Maximisation early stop.
//
// Replaces: dgeqr2_mia (reference src/dgeqr2.c:29-195), dlarfg_mia + d_sign
// (src/dlarfg.c:21-188) and dlarf_ (src/dlarf.c:28-211) as called at src/dgeqrdm_work.c:735-738.
//
// One cooperative grid, rows of the m_r x fjb panel split over CTAs; each CTA keeps its row slab
// in shared memory for the whole panel (global-memory mode when the slab does not fit: tall
// matrices).  Per column ONE grid-wide reduction, fused: the sweep that applies H_i also
// accumulates, for the next column x' = P[i+2:, i+1], the dot products x'.P[:, j] for all
// remaining j (j = i+1 gives ||x'||^2), so  w_j = P[i+1, j] + (x'.P_j)/(alpha - beta)  needs no
// second pass.  Partials are combined in fixed CTA order by every CTA, so all
// CTAs take bit-identical decisions (stop test, tau) and the result is run-to-run deterministic.
// Also emits the clean copy Vc of the reflectors (unit diagonal, zeros above, zero-padded to a
// multiple of 8 columns) that the trailing-update kernels consume.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ll.cuh"

#define PANEL_THREADS 512
#define PANEL_WARPS (PANEL_THREADS / 32)
#define PANEL_CPW ((63 + PANEL_WARPS - 1) / PANEL_WARPS)  // columns per warp in the sweep
#define PANEL_RG (PANEL_THREADS / 64)                      // reduction groups
#define QRDM_PANEL_AG_GAIN 0.45                            // us per column saved by the one-hop exchange (G <= 32), measured

template <bool SMEM>
__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel(qrdm_prob P, int rpc, unsigned epoch) {
  extern __shared__ __align__(16) double slab[];  // SMEM mode: [64][rpc]
  __shared__ double sred[PANEL_WARPS];
  __shared__ double S_[64], rowv[64], wv[64];
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // blocked tall-panel mode (P.sub = s + 1): this launch factors the 8-column sub-panel that starts
  // at panel column s; stop threshold / counters are carried in ctrl (see qrdm_geom)
  const QrdmGeom qg = qrdm_geom(P);
  const int j = qg.j, fjb = qg.fjb, sub_s = P.sub ? P.sub - 1 : 0;
  const int jmain = ctrl->j, fjb_main = ctrl->fjb;
  if (fjb <= 0) return;
  const int rows = P.m - j, lda = P.lda;
  const int G = gridDim.x, b = blockIdx.x;
  const int r0 = min(rows, b * rpc), r1 = min(rows, r0 + rpc), nr = r1 - r0;
  double* Ap = P.a + (size_t)j * lda + j;
  const int lds = rpc;
  LLPacket* part = reinterpret_cast<LLPacket*>(P.panel_part);  // [2][PANEL_MAXCTA][64]
  LLPacket* bcast = reinterpret_cast<LLPacket*>(P.panel_row);  // [2][128]: 64 totals + 64 pivot-row entries
  const unsigned tag_base = epoch << 8;

#define PX(r, c) (SMEM ? slab[(c) * lds + (r)] : Ap[(size_t)(c) * lda + (size_t)(r0 + (r))])

  if (SMEM) {
    for (int c = wid; c < fjb; c += PANEL_WARPS)
      for (int r = lane; r < nr; r += 32) slab[c * lds + r] = Ap[(size_t)c * lda + r0 + r];
    __syncthreads();
  }

  // partial sums for column 0: S_j = sum_{R>0} P[R,0] P[R,j]; row 0 itself is published directly
  for (int jj = wid; jj < fjb; jj += PANEL_WARPS) {
    double acc = 0.0;
    for (int r = lane; r < nr; r += 32) {
      const int R = r0 + r;
      const double p = PX(r, jj);
      if (R > 0) acc = fma(PX(r, 0), p, acc);
      else ll_store(&bcast[64 + jj], p, tag_base + 1);
    }
    acc = warp_sum(acc);
    if (lane == 0) ll_store(&part[(size_t)b * 64 + jj], acc, tag_base + 1);
  }

  double thres = (sub_s == 0) ? P.thres0 : ctrl->tall_thres;  // reference src/dgeqr2.c:40 (5e-14 x input scale)
  int k = fjb;
  long long tph[5] = {0, 0, 0, 0, 0};  // QRDM_B200_DEBUG & 8: per-phase cycle counts of CTA 0
  const bool timing = (P.debug & 8) && b == 0 && tid == 0;
  for (int i = 0; i < fjb; ++i) {
    const int cur = i & 1, nxt = cur ^ 1;
    const unsigned tag = tag_base + i + 1;
    long long tq0 = timing ? clock64() : 0;
    // ---- reduce-scatter: this CTA totals the columns jj == b (mod G) ----
    for (int jj = i + ((b - i % G + G) % G); jj < fjb; jj += G) {
      double v = 0.0;
      if (tid < G) v = ll_load(&part[((size_t)cur * QRDM_PANEL_MAXCTA + tid) * 64 + jj], tag);
      v = warp_sum(v);                       // lanes <-> CTAs: fixed association order
      if (lane == 0) sred[wid] = v;
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < PANEL_WARPS; ++w) tot += sred[w];
        ll_store(&bcast[cur * 128 + jj], tot, tag);
      }
      __syncthreads();
    }
    if (timing) { const long long tq = clock64(); tph[0] += tq - tq0; tq0 = tq; }
    // ---- broadcast: everybody picks up the totals and the pivot row ----
    if (tid < 128) {
      const int jj = tid & 63;
      if (jj >= i && jj < fjb) {
        const double v = ll_load(&bcast[cur * 128 + tid], tag);
        if (tid < 64) S_[jj] = v; else rowv[jj] = v;
      }
    }
    __syncthreads();
    if (timing) { const long long tq = clock64(); tph[1] += tq - tq0; tq0 = tq; }
    // ---- reflector scalars: dlarfg_mia (src/dlarfg.c:120-185), redundantly on every thread ----
    const double alpha = rowv[i];
    const int len = rows - i;
    double tau = 0.0, beta = alpha, scale = 1.0;
    if (len > 1) {
      const double xnorm = sqrt(S_[i]);
      if (sub_s + i > 0 && xnorm < thres && !ctrl->forced) { k = i; break; }  // DM early stop: column i left untouched
      if (xnorm != 0.0) {
        const double h = hypot(alpha, xnorm);
        beta = (alpha >= 0.0) ? -h : h;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
    }
    if (sub_s + i == 0 && fjb_main > 1 && P.tau_ > 0.0) thres = P.tau_ * fabs(beta);  // src/dgeqr2.c:176-177
    if (b == 0 && tid == 0) {
      P.tau[j + i] = tau;
      if (tau != tau && ctrl->err == 0) ctrl->err = -8;  // LAPACKE_dlarft's NaN screen of tau
    }
    const bool last = i + 1 >= fjb;
    if (tid < 64 && tid > i && tid < fjb) wv[tid] = tau * (rowv[tid] + S_[tid] * scale);
    const double w1 = last ? 0.0 : tau * (rowv[i + 1] + S_[i + 1] * scale);
    // ---- phase 1: v = x * scale, diagonal = beta; column i+1 gets H_i right away ----
    for (int r = tid; r < nr; r += PANEL_THREADS) {
      const int R = r0 + r;
      if (R < i) continue;
      double v = 1.0;
      if (R > i) {
        v = PX(r, i);
        if (tau != 0.0) { v *= scale; PX(r, i) = v; }
      } else {
        PX(r, i) = beta;
      }
      if (!last) {
        const double p = fma(-v, w1, PX(r, i + 1));
        PX(r, i + 1) = p;
        if (R == i + 1) ll_store(&bcast[nxt * 128 + 64 + i + 1], p, tag + 1);
      }
    }
    __syncthreads();
    if (timing) { const long long tq = clock64(); tph[2] += tq - tq0; tq0 = tq; }
    if (last) break;
    // ---- phase 2: remaining columns + the fused dot products for the next reflector ----
    {
      double wreg[PANEL_CPW], acc[PANEL_CPW];
#pragma unroll
      for (int c = 0; c < PANEL_CPW; ++c) {
        const int jj = i + 1 + wid + c * PANEL_WARPS;
        wreg[c] = jj < fjb ? wv[jj] : 0.0;
        acc[c] = 0.0;
      }
      for (int r = lane; r < nr; r += 32) {
        const int R = r0 + r;
        if (R < i) continue;
        const double v = (R == i) ? 1.0 : PX(r, i);
        const double x1 = PX(r, i + 1);
        const bool below = R > i + 1;
#pragma unroll
        for (int c = 0; c < PANEL_CPW; ++c) {
          const int jj = i + 1 + wid + c * PANEL_WARPS;
          if (jj < fjb) {
            double p;
            if (jj == i + 1) {
              p = x1;  // already updated in phase 1: only its squared norm is needed
            } else {
              p = fma(-v, wreg[c], PX(r, jj));
              PX(r, jj) = p;
              if (R == i + 1) ll_store(&bcast[nxt * 128 + 64 + jj], p, tag + 1);
            }
            if (below) acc[c] = fma(x1, p, acc[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < PANEL_CPW; ++c) {
        const int jj = i + 1 + wid + c * PANEL_WARPS;
        const double a = warp_sum(acc[c]);
        if (lane == 0 && jj < fjb) ll_store(&part[((size_t)nxt * QRDM_PANEL_MAXCTA + b) * 64 + jj], a, tag + 1);
      }
    }
    __syncthreads();  // sred / S_ / rowv / wv are rewritten by the next step
    if (timing) { const long long tq = clock64(); tph[3] += tq - tq0; tq0 = tq; }
  }
  if (timing && j == 640)
    printf("panel j=%d G=%d rpc=%d fjb=%d cycles: reduce %lld bcast %lld scalars+phase1 %lld phase2 %lld\n", j, G, rpc, fjb,
           tph[0], tph[1], tph[2], tph[3]);
  __syncthreads();
  if (b == 0 && tid == 0) {
    if (P.sub) {
      const int tk = (sub_s == 0 ? 0 : ctrl->tall_k) + k;
      ctrl->sub_k = k;
      ctrl->tall_k = tk;
      ctrl->tall_done = (k < fjb) ? 1 : 0;  // dead sub-panels never get here, so 0 is right after a full one
      if (k < fjb) ctrl->tall_stop_s = sub_s;
      ctrl->tall_thres = thres;
      ctrl->fjb_cmp = tk;
    } else {
      ctrl->fjb_cmp = k;
    }
  }

  // ---- write the slab back and emit Vc ----
  if (SMEM) {
    for (int c = wid; c < fjb; c += PANEL_WARPS)
      for (int r = lane; r < nr; r += 32) Ap[(size_t)c * lda + r0 + r] = slab[c * lds + r];
  }
  const int kpad = P.sub ? min(QRDM_TALL_B, 64 - qg.voff) : ((k + 7) & ~7);
  for (int q = wid; q < kpad; q += PANEL_WARPS) {
    double* vcol = P.vc + (size_t)(qg.voff + q) * P.ldv + j;
    for (int r = lane; r < nr; r += 32) {
      const int R = r0 + r;
      double v = 0.0;
      if (q < k) v = (R > q) ? PX(r, q) : (R == q ? 1.0 : 0.0);
      vcol[R] = v;
    }
    if (b == 0) {  // rows between the aligned tile start and j must read as zero
      const int jal = jmain & ~(QRDM_ROWALIGN - 1);
      for (int g = jal + lane; g < j; g += 32) P.vc[(size_t)(qg.voff + q) * P.ldv + g] = 0.0;
    }
  }
#undef PX
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant for panels of <= 128 rows per CTA (m_r <= 128 * #SMs = 18944 on B200:
// every square BASELINE config).  The cycle counters of the smem version (QRDM_B200_DEBUG=8) showed
// the per-column time was NOT the cross-CTA exchange but the SM-local work: ~3100 cycles of sweep
// (shared-memory bandwidth: every element read and written through smem each column) and ~3700
// cycles of reflector scalars computed redundantly by 512 threads on the 64-lane FP64 pipe.  Here
// each thread keeps its 16 slab entries (rows lane+32*ri, columns wid+16*c) in registers for the
// whole panel; only the two active columns (v and the next pivot column) pass through smem, the
// scalars are computed once per warp with a single sqrt (beta^2 = alpha^2 + ||x||^2, stop test on
// ||x||^2 < thres^2), and the exchange stays the LL reduce-scatter + broadcast of the kernel above.
// RI = rows per lane: 1, 2, 4 or 8 (32 ... 256 rows per CTA); the launcher picks the cheapest that fits
// Exchange per column (MODE):
//   1  G <= 32: every CTA totals all columns itself from the G partial packets — ONE cross-CTA hop;
//   2  larger grids, thread-block clusters of PANEL_CL: the CTAs of a cluster first add their partials in the
//      leader's shared memory (DSMEM stores + one cluster barrier), the leader publishes ONE packet per column,
//      and every CTA totals the G/PANEL_CL (<= 35) cluster packets itself — one cluster barrier + one global hop;
//   0  fallback: reduce-scatter to CTA (jj mod G) + broadcast — two global hops.
#define PANEL_CL 4
template <int RI, int MODE>
__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel_reg(qrdm_prob P, int rpc, unsigned epoch) {
  constexpr bool allgather = MODE != 0;
  __shared__ double sred[PANEL_WARPS];
  __shared__ double red[MODE != 0 ? 7 : 1][64];
  __shared__ double cpart[MODE == 2 ? 2 : 1][MODE == 2 ? PANEL_CL : 1][64];  // leader's copy: [buf][rank][column]
  __shared__ double S_[64], rowv[64], wv[64];
  __shared__ double vbuf[32 * RI], xbuf[32 * RI];
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, fjb = ctrl->fjb;
  if (fjb <= 0) return;
  const bool forced = ctrl->forced != 0;  // fixed columns: plain Householder QR, no early stop
  const int rows = P.m - j, lda = P.lda;
  const int G = gridDim.x, b = blockIdx.x;
  const int r0 = min(rows, b * rpc), r1 = min(rows, r0 + rpc), nr = r1 - r0;  // nr <= 128
  double* Ap = P.a + (size_t)j * lda + j;
  LLPacket* part = reinterpret_cast<LLPacket*>(P.panel_part);
  LLPacket* bcast = reinterpret_cast<LLPacket*>(P.panel_row);
  const unsigned tag_base = epoch << 8;
  // MODE 2: publishers are clusters; pub = index of this CTA's packet row, GP = number of packet rows to gather
  const int crank = MODE == 2 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
  const int pub = MODE == 2 ? b / PANEL_CL : b, GP = MODE == 2 ? G / PANEL_CL : G;
  double* cpart_leader = MODE == 2 ? cooperative_groups::this_cluster().map_shared_rank(&cpart[0][0][0], 0) : nullptr;
  // DSMEM rule: nobody may write into the leader's shared memory before the leader CTA has started running
  // (found by compute-sanitizer racecheck and by a wrong factorisation under ncu, whose scheduling differs)
  if (MODE == 2) cooperative_groups::this_cluster().sync();
  // publish the block sums acc[c] of columns wid + 16c (c < 4, columns > lo only) for exchange buffer `buf`
  auto publish = [&](double (&acc)[4], int buf, int lo, unsigned ptag) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int jj = wid + 16 * c;
      const double a = warp_sum(acc[c]);
      if (lane == 0 && jj > lo && jj < fjb) {
        if (MODE == 2) cpart_leader[((size_t)buf * PANEL_CL + crank) * 64 + jj] = a;
        else ll_store(&part[((size_t)buf * QRDM_PANEL_MAXCTA + b) * 64 + jj], a, ptag);
      }
    }
  };
  // MODE 2, after the cluster barrier: the leader adds the 8 partials of every column and publishes them
  auto publish_cluster = [&](int buf, int lo, unsigned ptag) {
    if (MODE == 2 && crank == 0 && tid < 64 && tid > lo && tid < fjb) {
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < PANEL_CL; ++r) t += cpart[buf][r][tid];
      ll_store(&part[((size_t)buf * QRDM_PANEL_MAXCTA + pub) * 64 + tid], t, ptag);
    }
  };

  double reg[RI][4];  // [ri][c]: row r0 + lane + 32*ri, column wid + 16*c
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = lane + 32 * ri, jj = wid + 16 * c;
      reg[ri][c] = (r < nr && jj < fjb) ? Ap[(size_t)jj * lda + r0 + r] : 0.0;
    }
  // column 0 through smem so that everybody can form the first dot products
  if (wid == 0) {
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) xbuf[lane + 32 * ri] = reg[ri][0];
  }
  __syncthreads();
  {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = lane + 32 * ri, R = r0 + r;
      const double x0 = xbuf[r];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = wid + 16 * c;
        if (r < nr && jj < fjb) {
          if (R > 0) acc[c] = fma(x0, reg[ri][c], acc[c]);
          else ll_store(&bcast[64 + jj], reg[ri][c], tag_base + 1);
        }
      }
    }
    publish(acc, 0, -1, tag_base + 1);
  }
  if (MODE == 2) { cooperative_groups::this_cluster().sync(); publish_cluster(0, -1, tag_base + 1); }
  __syncthreads();

  // nb > 64: this panel may CONTINUE a wider block (micro-panel t > 0, k_wide.cu): the stop test then applies from its
  // first column on and the threshold set by the block's first column is carried in ctrl
  const bool cont = ctrl->micro_t > 0;
  double thres2 = cont ? ctrl->micro_thres2 : P.thres0 * P.thres0;  // (reference src/dgeqr2.c:40)^2, 5e-14 x input scale
  int k = fjb;
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // QRDM_B200_DEBUG & 8: per-phase cycle counts of one CTA
  const bool timing = (P.debug & 8) && b == (G > 40 ? 40 : 0) && tid == 0;
  for (int i = 0; i < fjb; ++i) {
    const int cur = i & 1, nxt = cur ^ 1;
    const unsigned tag = tag_base + i + 1;
    long long tq0 = timing ? clock64() : 0;
    if (allgather) {
      // ---- one hop: thread (jj, grp) sums the partials of CTAs grp, grp+7, ... (<= 5) of column jj; threads
      // 448..511 pick up the pivot row meanwhile; 64 threads then add the 7 group sums in fixed order.  Every
      // CTA adds the same packets in the same order => bit-identical totals (and decisions) everywhere ----
      const int jj = tid & 63, grp = tid >> 6;
      if (grp < 7) {
        double v = 0.0;
        if (jj >= i && jj < fjb)
          v = ll_gather_sum_strided(&part[(size_t)cur * QRDM_PANEL_MAXCTA * 64 + jj], 64, grp, 7, GP, tag);
        red[grp][jj] = v;
      } else if (jj >= i && jj < fjb) {
        rowv[jj] = ll_load(&bcast[cur * 128 + 64 + jj], tag);
      }
      __syncthreads();
      if (tid < 64 && tid >= i && tid < fjb)
        S_[tid] = (((((red[0][tid] + red[1][tid]) + red[2][tid]) + red[3][tid]) + red[4][tid]) + red[5][tid]) + red[6][tid];
      if (timing) { const long long tq = clock64(); tph[0] += tq - tq0; tq0 = tq; }
    } else {
    // ---- reduce-scatter: column jj is totalled by warp (jj - i - off) / G of CTA jj mod G, so a small
    // grid (few rows) spreads its columns over the warps instead of looping over them; lanes <-> CTAs,
    // fixed association order, no block barrier ----
    for (int jj = i + ((b - i % G + G) % G) + wid * G; jj < fjb; jj += PANEL_WARPS * G) {
      double v = ll_gather_sum(&part[(size_t)cur * QRDM_PANEL_MAXCTA * 64 + jj], 64, lane, G, tag);
      v = warp_sum(v);
      if (lane == 0) ll_store(&bcast[cur * 128 + jj], v, tag);
    }
    if (timing) { const long long tq = clock64(); tph[0] += tq - tq0; tq0 = tq; }
    // ---- broadcast: everybody picks up the totals and the pivot row ----
    if (tid < 128) {
      const int jj = tid & 63;
      if (jj >= i && jj < fjb) {
        const double v = ll_load(&bcast[cur * 128 + tid], tag);
        if (tid < 64) S_[jj] = v; else rowv[jj] = v;
      }
    }
    }
    __syncthreads();
    if (timing) { const long long tq = clock64(); tph[1] += tq - tq0; tq0 = tq; }
    // ---- reflector scalars (dlarfg_mia, src/dlarfg.c:120-185), one sqrt, two divisions ----
    const double alpha = rowv[i], xn2 = S_[i];
    const int len = rows - i;
    double tau = 0.0, beta = alpha, scale = 1.0;
    if (len > 1) {
      if ((i > 0 || cont) && xn2 < thres2 && !forced) { k = i; break; }  // DM early stop: column i left untouched
      // Only the warps that consume the scalars compute them: warps 0-1 (wv, tau), the owner of column i
      // (scale, beta) and, in column 0, everybody (thres2 needs beta).  The chain itself costs ~350 cycles
      // (measured by running it twice); the 2.5-3.5 k cycles QRDM_B200_DEBUG=8 books under "scalars" are the
      // broadcast wait: a clock read right after __syncthreads captures the barrier's ISSUE, not its release.
      if (xn2 != 0.0 && (i == 0 || wid < 2 || wid == (i & 15))) {
        const double h = sqrt(fma(alpha, alpha, xn2));
        beta = (alpha >= 0.0) ? -h : h;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
    }
    if (i == 0 && !cont && fjb > 1 && P.tau_ > 0.0) { const double th = P.tau_ * fabs(beta); thres2 = th * th; }
    if (b == 0 && tid == 0) {
      P.tau[j + i] = tau;
      if (tau != tau && ctrl->err == 0) ctrl->err = -8;  // LAPACKE_dlarft's NaN screen of tau
    }
    if (timing) { const long long tq = clock64(); tph[5] += tq - tq0; }
    const bool last = i + 1 >= fjb;
    const int wi = i & 15, ci = i >> 4;
    // ---- publish v (column i, scaled) and the old next pivot column through smem ----
    if (wid == wi) {
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) {
        const int r = lane + 32 * ri, R = r0 + r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c == ci) {
            double v = 0.0;
            if (r < nr) {
              if (R > i) { v = reg[ri][c]; if (tau != 0.0) { v *= scale; reg[ri][c] = v; } }
              else if (R == i) { v = 1.0; reg[ri][c] = beta; }
            }
            vbuf[r] = v;
          }
        }
      }
    }
    if (!last && wid == ((i + 1) & 15)) {
      const int c1 = (i + 1) >> 4;
#pragma unroll
      for (int ri = 0; ri < RI; ++ri)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c == c1) xbuf[lane + 32 * ri] = reg[ri][c];
    }
    if (tid < 64 && tid > i && tid < fjb) wv[tid] = tau * (rowv[tid] + S_[tid] * scale);
    if (timing) { const long long tq = clock64(); tph[6] += tq - tq0; }
    __syncthreads();
    if (timing) { const long long tq = clock64(); tph[2] += tq - tq0; tq0 = tq; }
    if (last) break;
    // ---- sweep: apply H_i to the remaining columns, fused dot products for the next reflector ----
    {
      const double w1 = wv[i + 1];
      double wreg[4], acc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = wid + 16 * c;
        wreg[c] = (jj > i && jj < fjb) ? wv[jj] : 0.0;
        acc[c] = 0.0;
      }
#pragma unroll
      for (int ri = 0; ri < RI; ++ri) {
        const int r = lane + 32 * ri, R = r0 + r;
        const double v = vbuf[r];
        const double x1 = fma(-v, w1, xbuf[r]);  // next pivot column after H_i
        const bool below = R > i + 1 && r < nr;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jj = wid + 16 * c;
          if (jj > i && jj < fjb) {
            const double p = fma(-v, wreg[c], reg[ri][c]);
            reg[ri][c] = p;
            if (below) acc[c] = fma(x1, p, acc[c]);
            else if (R == i + 1 && r < nr) ll_store(&bcast[nxt * 128 + 64 + jj], p, tag + 1);
          }
        }
      }
      publish(acc, nxt, i, tag + 1);
    }
    if (timing) { const long long tq = clock64(); tph[3] += tq - tq0; tq0 = tq; }
    if (MODE == 2) { cooperative_groups::this_cluster().sync(); publish_cluster(nxt, i, tag + 1); }
    else __syncthreads();  // vbuf / xbuf / S_ / rowv / wv are rewritten by the next step
    if (timing) { const long long tq = clock64(); tph[4] += tq - tq0; tq0 = tq; }
  }
  if (timing && j == 640)
    printf("panel_reg j=%d G=%d rpc=%d fjb=%d cycles: reduce %lld bcast %lld scalars+publish %lld (scalars %lld, to-sync %lld) sweep %lld endsync %lld\n", j, G, rpc, fjb,
           tph[0], tph[1], tph[2], tph[5], tph[6], tph[3], tph[4]);
  __syncthreads();
  if (b == 0 && tid == 0) { ctrl->fjb_cmp = k; ctrl->micro_thres2 = thres2; }

  // ---- write the slab back and emit Vc ----
  const int kpad = (k + 7) & ~7;
  const int jal = j & ~(QRDM_ROWALIGN - 1);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int jj = wid + 16 * c;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = lane + 32 * ri, R = r0 + r;
      if (r < nr && jj < fjb) Ap[(size_t)jj * lda + R] = reg[ri][c];
      if (r < nr && jj < kpad) {
        double v = 0.0;
        if (jj < k) v = (R > jj) ? reg[ri][c] : (R == jj ? 1.0 : 0.0);
        P.vc[(size_t)jj * P.ldv + j + R] = v;
      }
    }
    if (b == 0 && jj < kpad)
      for (int g = jal + lane; g < j; g += 32) P.vc[(size_t)jj * P.ldv + g] = 0.0;
  }
  if (MODE == 2) cooperative_groups::this_cluster().sync();  // no CTA of a cluster leaves while its shared memory may still be addressed
}

// ---------------------------------------------------------------------------------------------
// Grouped register panel: ONE cross-CTA exchange per group of up to S columns instead of one per column.
//
// The per-column kernel above spends its 3.9 us per column (16384 rows, 128 CTAs) almost entirely on the exchange that
// turns the CTAs' partial sums into [||x||^2, x'C] — 64 dependent exchanges per panel.  But everything a reflector
// needs besides the slab itself is a handful of inner products, and inner products can be carried through a
// Householder update algebraically.  For a group that starts at panel column i0 the exchange delivers, for its columns
// u = 0..S-1 and every remaining column jj >= i0,
//     M[u][jj]  = sum over rows R > i0 of  c_{i0+u}[R] c_jj[R]        (S rows of the panel's Gram matrix)
//     Rw[u][jj] = c_jj[i0+u]                                          (the group's S pivot rows)
// and every CTA then runs the same scalar recurrence on these replicated numbers.  Step t (column col = i0 + t) takes
// alpha = Rw[t][col], ||x||^2 = M[t][col], forms beta, tau, scale and w_jj = tau (Rw[t][jj] + scale M[t][jj]) exactly
// as dlarfg/dlarf do (src/dlarfg.c:120-185, src/dlarf.c:173-184), sweeps the slab registers with H = I - tau v v', and
// carries the state to step t + 1 (v = scale x below the diagonal, sums over R > col):
//     Rw'[u][jj] = Rw[u][jj] - (scale Rw[u][col]) w_jj                                              (the sweep, row i0+u)
//     M~[u][jj]  = M[u][jj] - w_jj (scale M[t][cu]) - w_cu (scale M[t][jj]) + w_cu w_jj scale^2 M[t][col]   (cu = i0+u)
//     M'[u][jj]  = M~[u][jj] - Rw'[t+1][cu] Rw'[t+1][jj]                            (row col + 1 leaves the sums)
// The pivot-row recurrence repeats the sweep's own operations, so alpha is bit-identical to the slab's entry; the
// downdated sums are exact up to eps times the ORIGINAL column norms, i.e. they lose accuracy only when a column
// loses most of its norm inside the group.  That is tested: if ||x||^2 comes out below 2^-7 of the column's sum at the
// start of the group, the group ends there and a fresh exchange delivers exact sums (same decision on every CTA:
// all of them hold the same numbers).  Error amplification is therefore bounded by ~2^7 in ||x||^2 (1.4e-14
// relative) and ~11 in w; Gaussian panels never trip the test, DM-selected panels rarely (their columns are far from
// parallel by construction).  Exchanges are numbered (tag, buffer parity) by a counter, not by the column index.
// MODE 1 (<= 32 CTAs) / 2 (clusters) as in k_panel_reg; larger grids without clusters keep the per-column kernel.
#define GRP_GUARD 0x1p-7
__device__ __forceinline__ void bar_sync_64() { asm volatile("bar.sync 1, 64;\n" ::: "memory"); }  // warps 0-1 only
// w_c = tau v'c = tau (c[i] + scale x'c) for a column with pivot-row entry r and sum m; and the carry of a Gram entry
// through H_i (see the header of k_panel_grp).  Fixed rounding: both phases of the recurrence call exactly these.
__device__ __forceinline__ double grp_w(double tau, double scale, double m, double r) { return __dmul_rn(tau, __fma_rn(m, scale, r)); }
__device__ __forceinline__ double grp_m_update(double m_uj, double w_j, double sv_j, double r1_j, double w_u, double sv_u, double r1_u, double vv) {
  double t = __fma_rn(-w_j, sv_u, m_uj);
  t = __fma_rn(-w_u, sv_j, t);
  t = __fma_rn(__dmul_rn(w_u, w_j), vv, t);
  return __fma_rn(-r1_u, r1_j, t);
}
template <int RI, int MODE, int S>
__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel_grp(qrdm_prob P, int rpc, unsigned epoch) {
  static_assert(MODE == 1 || MODE == 2, "all-gather exchange only");
  static_assert(S == 2 || S == 4, "group width");
  constexpr int NV = S * 64;                 // values per exchange
  constexpr int NGRP = PANEL_THREADS / NV;   // gather groups (4 / 2)
  constexpr int GB = 6;                      // packet loads a thread keeps in flight
  __shared__ double red[NGRP][NV];
  __shared__ double cpart[MODE == 2 ? 2 : 1][MODE == 2 ? PANEL_CL : 1][NV];  // leader's copy: [buf][rank][value]
  // the group's Gram rows / pivot rows, double-buffered by step parity (a step reads [t & 1], writes the rows u > t of [~t & 1])
  __shared__ double Mx[S][64], Rw[S][64];  // the group's Gram rows / pivot rows as delivered by the exchange
  __shared__ double Wt[S][64 + S], SC[S][4], SCX[S][2];  // per step: w (0 for columns <= i and beyond the panel), {tau, beta, scale, vv}, {vr1, sc}
  __shared__ double GW[S][S], GSV[S][S], GR1[S][S], GVR[S][S];  // per step, per group column u: w_u, sv_u, r1_u, vr_u
  __shared__ double SEFF[64], BETA[64];   // per finished column: the scale of its reflector (1 when tau = 0) and beta
  __shared__ int ginfo[2];                // reflectors produced by the group, DM stop flag
  __shared__ double xbuf[S][32 * RI];
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, fjb = ctrl->fjb;
  if (fjb <= 0) return;
  const bool forced = ctrl->forced != 0;
  const int rows = P.m - j, lda = P.lda;
  const int G = gridDim.x, b = blockIdx.x;
  const int r0 = min(rows, b * rpc), r1 = min(rows, r0 + rpc), nr = r1 - r0;
  double* Ap = P.a + (size_t)j * lda + j;
  LLPacket* part = reinterpret_cast<LLPacket*>(P.panel_part) + QRDM_PANEL_PART_PKTS;  // [2][MAXCTA][NV]
  LLPacket* brow = reinterpret_cast<LLPacket*>(P.panel_row) + QRDM_PANEL_ROW_PKTS;    // [2][NV]
  const unsigned tag_base = epoch << 8;
  const int crank = MODE == 2 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
  const int pub = MODE == 2 ? b / PANEL_CL : b, GP = MODE == 2 ? G / PANEL_CL : G;
  double* cpart_leader = MODE == 2 ? cooperative_groups::this_cluster().map_shared_rank(&cpart[0][0][0], 0) : nullptr;
  if (MODE == 2) cooperative_groups::this_cluster().sync();  // DSMEM rule, see k_panel_reg

  double reg[RI][4];  // [ri][c]: row r0 + lane + 32*ri, column wid + 16*c
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = lane + 32 * ri, jj = wid + 16 * c;
      reg[ri][c] = (r < nr && jj < fjb) ? Ap[(size_t)jj * lda + r0 + r] : 0.0;
    }
  double xg[S][RI];  // this thread's rows of the group's columns (private copy, swept along with the slab)

  // ---- local part of exchange e for the group that starts at column i0: partial Gram rows + the pivot rows ----
  auto contribute = [&](int i0, int sg, int e) {
    const int buf = e & 1;
    const unsigned tag = tag_base + e + 1;
#pragma unroll
    for (int u = 0; u < S; ++u) {  // the group's columns through smem (their owner warps hold them)
      const int cu = i0 + u;
      if (u < sg && wid == (cu & 15)) {
        const int cs = cu >> 4;  // (selects, not a dynamic index: the slab must stay in registers)
#pragma unroll
        for (int ri = 0; ri < RI; ++ri)
          xbuf[u][lane + 32 * ri] = cs == 0 ? reg[ri][0] : (cs == 1 ? reg[ri][1] : (cs == 2 ? reg[ri][2] : reg[ri][3]));
      }
    }
    __syncthreads();
    double acc[S][4];
#pragma unroll
    for (int u = 0; u < S; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[u][c] = 0.0;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = lane + 32 * ri, R = r0 + r;
#pragma unroll
      for (int u = 0; u < S; ++u) xg[u][ri] = (u < sg && r < nr) ? xbuf[u][r] : 0.0;
      if (r < nr && R >= i0) {
        if (R > i0) {
#pragma unroll
          for (int u = 0; u < S; ++u)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[u][c] = fma(xg[u][ri], reg[ri][c], acc[u][c]);
        }
        if (R < i0 + sg) {  // pivot row R - i0 of the group
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int jj = wid + 16 * c;
            if (jj >= i0 && jj < fjb) ll_store(&brow[(size_t)buf * NV + (R - i0) * 64 + jj], reg[ri][c], tag);
          }
        }
      }
    }
    // Transposing warp reduction of the 4S accumulators: each butterfly stage halves the number of values a lane
    // carries (it keeps the half selected by its lane bit and adds the partner's partial of that half), so the 4S
    // sums cost 4S shuffles instead of 5 x 4S, and they end up on 4S different lanes, which publish in parallel.
    {
      constexpr int NA = 4 * S;
      double a[NA];
#pragma unroll
      for (int u = 0; u < S; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[u * 4 + c] = acc[u][c];
      int idx = 0;  // index of the value this lane ends up with
#pragma unroll
      for (int half = NA / 2, o = 16; half >= 1; half >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int q = 0; q < half; ++q) {
          const double send = up ? a[q] : a[q + half];
          const double keep = up ? a[q + half] : a[q];
          a[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        idx = idx * 2 + (up ? 1 : 0);
      }
      // remaining lane bits (2 for S = 2, 1 for S = 4): plain butterfly on the single value left
#pragma unroll
      for (int o = (S == 4 ? 1 : 2); o >= 1; o >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
      const int u = idx >> 2, jj = wid + 16 * (idx & 3);
      const bool first = (lane & (S == 4 ? 1 : 3)) == 0;  // one of the lanes that hold the same total
      if (first && u < sg && jj >= i0 && jj < fjb) {
        if (MODE == 2) cpart_leader[((size_t)buf * PANEL_CL + crank) * NV + u * 64 + jj] = a[0];
        else ll_store(&part[((size_t)buf * QRDM_PANEL_MAXCTA + b) * NV + u * 64 + jj], a[0], tag);
      }
    }
    if (MODE == 2) {
      cooperative_groups::this_cluster().sync();
      if (crank == 0 && tid < NV) {
        const int u = tid >> 6, jj = tid & 63;
        if (u < sg && jj >= i0 && jj < fjb) {
          double t = 0.0;
#pragma unroll
          for (int r = 0; r < PANEL_CL; ++r) t += cpart[buf][r][tid];
          ll_store(&part[((size_t)buf * QRDM_PANEL_MAXCTA + pub) * NV + tid], t, tag);
        }
      }
    }
  };
  // ---- gather exchange e: the same packets added in the same order everywhere => identical M / Rw in every CTA.
  // MODE 1: every CTA reads all G x NV packets.  MODE 2: that all-gather is L2-bandwidth bound (124 CTAs x 31
  // publishers x 256 packets x 16 B = 15.7 MB per exchange), so the CTAs of a cluster share it: CTA r totals the r-th
  // quarter of the values and stores its totals into the Mx of all four CTAs through DSMEM (+ one cluster barrier) ----
  constexpr int QV = MODE == 2 ? NV / PANEL_CL : NV;   // values this CTA totals
  constexpr int QGRP = PANEL_THREADS / QV;             // publisher groups (threads per value)
  double* mx_peer[PANEL_CL];
#pragma unroll
  for (int rr = 0; rr < PANEL_CL; ++rr)
    mx_peer[rr] = MODE == 2 ? cooperative_groups::this_cluster().map_shared_rank(&Mx[0][0], rr) : &Mx[0][0];
  double* redq = &red[0][0];  // [QGRP][QV], same storage
  auto gather = [&](int i0, int sg, int e) {
    const int buf = e & 1;
    const unsigned tag = tag_base + e + 1;
    const int xq = tid % QV, grp = tid / QV;
    const int x = (MODE == 2 ? crank * QV : 0) + xq, u = x >> 6, jj = x & 63;
    const bool live = u < sg && jj >= i0 && jj < fjb;
    double v = 0.0;
    if (live) {
      const LLPacket* base = &part[(size_t)buf * QRDM_PANEL_MAXCTA * NV + x];
      for (int c0 = grp; c0 < GP; c0 += QGRP * GB) {  // publishers grp, grp + QGRP, ...: GB loads in flight
        unsigned lo[GB], t0[GB], hi[GB], t1[GB];
#pragma unroll
        for (int q = 0; q < GB; ++q) {
          const int c = c0 + q * QGRP;
          if (c < GP)
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                         : "=r"(lo[q]), "=r"(t0[q]), "=r"(hi[q]), "=r"(t1[q]) : "l"(base + (size_t)c * NV) : "memory");
        }
#pragma unroll
        for (int q = 0; q < GB; ++q) {
          const int c = c0 + q * QGRP;
          if (c < GP) {
            unsigned spins = 0;
            while (t0[q] != tag || t1[q] != tag) {
              asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                           : "=r"(lo[q]), "=r"(t0[q]), "=r"(hi[q]), "=r"(t1[q]) : "l"(base + (size_t)c * NV) : "memory");
              if (++spins > LL_SPIN_LIMIT) __trap();
            }
            v += __longlong_as_double((long long)(((unsigned long long)hi[q] << 32) | lo[q]));
          }
        }
      }
    }
    redq[grp * QV + xq] = v;
    if (tid < NV) {  // pivot row i0 + u' (rows beyond the panel do not exist: zero); every CTA reads all of them
      const int u2 = tid >> 6, j2 = tid & 63;
      double rv = 0.0;
      if (u2 < sg && j2 >= i0 && j2 < fjb && i0 + u2 < rows) rv = ll_load(&brow[(size_t)buf * NV + tid], tag);
      Rw[u2][j2] = rv;
    }
    __syncthreads();
    if (tid < QV) {
      double t = redq[tid];
#pragma unroll
      for (int q = 1; q < QGRP; ++q) t += redq[q * QV + tid];
      const int xo = (MODE == 2 ? crank * QV : 0) + tid;  // index into Mx viewed as [NV]
#pragma unroll
      for (int rr = 0; rr < (MODE == 2 ? PANEL_CL : 1); ++rr) mx_peer[rr][xo] = t;
    }
    if (MODE == 2) cooperative_groups::this_cluster().sync();
    else __syncthreads();
  };

  const bool cont = ctrl->micro_t > 0;
  double thres2 = cont ? ctrl->micro_thres2 : P.thres0 * P.thres0;  // (reference src/dgeqr2.c:40)^2; kept by warps 0-1
  int k = fjb, e = 0, i0 = 0;
  bool stopped = false;
  long long tph[6] = {0, 0, 0, 0, 0, 0};  // QRDM_B200_DEBUG & 8: per-phase cycle counts of one CTA
  const bool timing = (P.debug & 8) && b == (G > 40 ? 40 : 0) && tid == 0;
  long long tq0 = timing ? clock64() : 0;
#define GRP_TICK(slot) do { if (timing) { const long long tq = clock64(); tph[slot] += tq - tq0; tq0 = tq; } } while (0)
  while (i0 < fjb && !stopped) {
    const int sg = min(S, fjb - i0);
    contribute(i0, sg, e);  // (single call site: the lambda must be inlined, it works on the register slab)
    GRP_TICK(0);
    gather(i0, sg, e);
    GRP_TICK(1);
    // ---- the scalar recurrence of the whole group, on replicated numbers only (no slab access), in two phases.
    // Phase 1 (warp 0, critical path): the reflector scalars of all steps depend only on the S x S block of the
    // group's own columns; lane (u, u') carries M[u][cu'] and Rw[u][cu'] in registers and fetches what it needs of
    // row t by shuffles — no shared-memory round trip, no barrier between steps.  Per step it publishes the scalars and
    // the group-column terms (w_u, sv_u, r1_u, vr_u) that every column's update needs.
    // Phase 2 (warps 0-1, one thread per column): each column runs its own S-step recurrence in registers against
    // those published terms and produces w for the sweep.  Both phases evaluate grp_w / grp_m_update with the same
    // rounding, so a group column's phase-2 values are bit-identical to the block's. ----
    if (tid < 64) {
      const int jj = tid;
      double m[S], rw[S];
#pragma unroll
      for (int u = 0; u < S; ++u) { m[u] = Mx[u][jj]; rw[u] = Rw[u][jj]; }
      if (wid == 0) {
        const int u = lane / S, up = lane % S;
        const bool inblk = lane < S * S && u < sg && up < sg;
        double gm = inblk ? Mx[u][i0 + up] : 0.0, grw = inblk ? Rw[u][i0 + up] : 0.0;
        const double gm0 = gm;
        int t = 0;
        bool stop = false, go = true;
#pragma unroll
        for (int tt = 0; tt < S; ++tt) {
          if (tt < sg && go) {
            const int i = i0 + tt;
            const double alpha = __shfl_sync(0xffffffffu, grw, tt * S + tt), xn2 = __shfl_sync(0xffffffffu, gm, tt * S + tt);
            const double d0t = __shfl_sync(0xffffffffu, gm0, tt * S + tt);
            // a column that lost most of its norm inside the group: its downdated sums are no longer trustworthy
            if (tt > 0 && !(xn2 >= GRP_GUARD * d0t)) go = false;
            const int len = rows - i;
            if (go && len > 1 && (i > 0 || cont) && xn2 < thres2 && !forced) { stop = true; go = false; }  // DM early stop
            if (go) {
              double tau = 0.0, beta = alpha, scale = 1.0;
              if (len > 1 && xn2 != 0.0) {
                // dlarfg_mia, src/dlarfg.c:120-185.  The serial chain of the panel runs through these scalars once per
                // column: one reciprocal square root + one reciprocal (206 cycles) instead of sqrt + two divisions (345,
                // tools/fp64_lat.cu): h = s2 * rsqrt(s2), 1/beta = -+rsqrt(s2); beta and tau differ from the divided
                // forms by at most an ulp or two, far inside the 1e-10 parity tolerance
                const double s2 = fma(alpha, alpha, xn2);
                if (s2 > 0x1p-1000 && s2 < 0x1p1000) {
                  const double rh = rsqrt(s2), h = s2 * rh;
                  beta = (alpha >= 0.0) ? -h : h;
                  tau = (beta - alpha) * ((alpha >= 0.0) ? -rh : rh);
                } else {  // (never with the driver's pre-scaling; keeps Inf / NaN behaviour of the plain formulas)
                  const double h = sqrt(s2);
                  beta = (alpha >= 0.0) ? -h : h;
                  tau = (beta - alpha) / beta;
                }
                scale = 1.0 / (alpha - beta);
              }
              if (i == 0 && !cont && fjb > 1 && P.tau_ > 0.0) { const double th = P.tau_ * fabs(beta); thres2 = th * th; }
              const double sc = tau != 0.0 ? scale : 0.0;  // tau == 0: H = I, only the row leaves the sums
              const double vs = tau != 0.0 ? scale : 1.0;
              const double vv = sc * sc * xn2;
              // row tt of the block for this lane's two columns, row tt + 1, and column i of rows u and tt + 1
              const double Mt_u = __shfl_sync(0xffffffffu, gm, (tt * S + u) & 31), Mt_up = __shfl_sync(0xffffffffu, gm, tt * S + up);
              const double Rt_u = __shfl_sync(0xffffffffu, grw, (tt * S + u) & 31), Rt_up = __shfl_sync(0xffffffffu, grw, tt * S + up);
              const double R1_u = __shfl_sync(0xffffffffu, grw, ((tt + 1) * S + u) & 31), R1_up = __shfl_sync(0xffffffffu, grw, ((tt + 1) * S + up) & 31);
              const double Ru_i = __shfl_sync(0xffffffffu, grw, (u * S + tt) & 31), R1_i = __shfl_sync(0xffffffffu, grw, ((tt + 1) * S + tt) & 31);
              const double vr1 = R1_i * vs, vr_u = Ru_i * vs;
              const double w_u = grp_w(tau, scale, Mt_u, Rt_u), w_up = grp_w(tau, scale, Mt_up, Rt_up);
              const double sv_u = __dmul_rn(sc, Mt_u), sv_up = __dmul_rn(sc, Mt_up);
              const double r1_u = __fma_rn(-vr1, w_u, R1_u), r1_up = __fma_rn(-vr1, w_up, R1_up);
              if (lane == 0) {
                SC[tt][0] = tau; SC[tt][1] = beta; SC[tt][2] = scale; SC[tt][3] = vv;
                SCX[tt][0] = vr1; SCX[tt][1] = sc;
                SEFF[i] = vs; BETA[i] = beta;
                if (b == 0) {
                  P.tau[j + i] = tau;
                  if (tau != tau && ctrl->err == 0) ctrl->err = -8;  // LAPACKE_dlarft's NaN screen of tau
                }
              }
              if (lane < S * S && up == 0) { GW[tt][u] = w_u; GSV[tt][u] = sv_u; GR1[tt][u] = r1_u; GVR[tt][u] = vr_u; }
              if (u > tt && up > tt) {  // the block after H_i, row i0 + tt + 1 out of the sums
                grw = __fma_rn(-vr_u, w_up, grw);
                gm = grp_m_update(gm, w_up, sv_up, r1_up, w_u, sv_u, r1_u, vv);
              }
              t = tt + 1;
            }
          }
        }
        if (lane == 0) { ginfo[0] = t; ginfo[1] = stop ? 1 : 0; }
      }
      if (timing) { tph[5] += clock64() - tq0; }
      bar_sync_64();
      const int nst = ginfo[0];
#pragma unroll
      for (int t = 0; t < S; ++t) {
        if (t < nst) {
          const int i = i0 + t;
          const double tau = SC[t][0], scale = SC[t][2], vv = SC[t][3], vr1 = SCX[t][0], sc = SCX[t][1];
          const double w_j = (jj > i && jj < fjb) ? grp_w(tau, scale, m[t], rw[t]) : 0.0;  // 0: the sweep's FMA leaves the column alone
          Wt[t][jj] = w_j;
          if (jj < S) Wt[t][64 + jj] = 0.0;
          if (t + 1 < S) {
            const double sv_j = __dmul_rn(sc, m[t]);
            const double r1_j = __fma_rn(-vr1, w_j, rw[t + 1]);
#pragma unroll
            for (int u = 1; u < S; ++u)
              if (u > t) {
                rw[u] = __fma_rn(-GVR[t][u], w_j, rw[u]);  // the sweep's own operation on the replicated row
                m[u] = grp_m_update(m[u], w_j, sv_j, r1_j, GW[t][u], GSV[t][u], GR1[t][u], vv);
              }
          }
        }
      }
    }
    GRP_TICK(2);
    __syncthreads();
    GRP_TICK(3);
    const int nsteps = ginfo[0];
    stopped = ginfo[1] != 0;
    // ---- apply the group's reflectors to the slab registers, one after the other, no barrier in between.  The code
    // is kept lean on purpose (16 warps issue every instruction of it): unconditional FMAs against w = 0 for the
    // columns that must not change; a reflector's own column keeps the raw x in the slab (its private copy xg feeds
    // the later steps) and is scaled / gets beta only when the slab is written back (SEFF, BETA) ----
#pragma unroll
    for (int t = 0; t < S; ++t) {
      if (t < nsteps) {
        const int i = i0 + t;
        const double seff = SC[t][0] != 0.0 ? SC[t][2] : 1.0;
        double wreg[4], wg[S];
#pragma unroll
        for (int c = 0; c < 4; ++c) wreg[c] = Wt[t][wid + 16 * c];
#pragma unroll
        for (int u = 0; u < S; ++u) wg[u] = u > t ? Wt[t][i0 + u] : 0.0;
#pragma unroll
        for (int ri = 0; ri < RI; ++ri) {
          const int r = lane + 32 * ri, R = r0 + r;
          // rows beyond this CTA's slab hold zeros in xg, so R > i needs no r < nr test
          const double v = R > i ? xg[t][ri] * seff : ((R == i && r < nr) ? 1.0 : 0.0);
#pragma unroll
          for (int c = 0; c < 4; ++c) reg[ri][c] = fma(-v, wreg[c], reg[ri][c]);
#pragma unroll
          for (int u = 0; u < S; ++u)
            if (u > t) xg[u][ri] = fma(-v, wg[u], xg[u][ri]);
        }
      }
    }
    GRP_TICK(4);
    i0 += nsteps;
    if (stopped) k = i0;
    ++e;
  }
#undef GRP_TICK
  if (timing && j == 640)
    printf("panel_grp<S=%d> j=%d G=%d rpc=%d fjb=%d exchanges=%d cycles: contribute %lld gather %lld recurrence %lld (phase 1: %lld) sync %lld sweeps %lld\n",
           S, j, G, rpc, fjb, e, tph[0], tph[1], tph[2], tph[5], tph[3], tph[4]);
  __syncthreads();
  if (b == 0 && tid == 0) { ctrl->fjb_cmp = k; ctrl->micro_thres2 = thres2; }

  // ---- write the slab back and emit Vc ----
  const int kpad = (k + 7) & ~7;
  const int jal = j & ~(QRDM_ROWALIGN - 1);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int jj = wid + 16 * c;
    const double sf = jj < k ? SEFF[jj] : 1.0, bt = jj < k ? BETA[jj] : 0.0;
#pragma unroll
    for (int ri = 0; ri < RI; ++ri) {
      const int r = lane + 32 * ri, R = r0 + r;
      double val = reg[ri][c];
      if (jj < k) val = R > jj ? val * sf : (R == jj ? bt : val);  // the reflector itself: v below the diagonal, beta on it
      if (r < nr && jj < fjb) Ap[(size_t)jj * lda + R] = val;
      if (r < nr && jj < kpad) {
        double v = 0.0;
        if (jj < k) v = (R > jj) ? val : (R == jj ? 1.0 : 0.0);
        P.vc[(size_t)jj * P.ldv + j + R] = v;
      }
    }
    if (b == 0 && jj < kpad)
      for (int g = jal + lane; g < j; g += 32) P.vc[(size_t)jj * P.ldv + g] = 0.0;
  }
  if (MODE == 2) cooperative_groups::this_cluster().sync();  // no CTA of a cluster leaves while its shared memory may still be addressed
}

// ---------------------------------------------------------------------------------------------
// Sub-panel kernel of the blocked tall panel: QRDM_TALL_B (= 8) columns, slab in global memory.
// With so few columns the row dimension is what has to be parallel: threads <-> rows (coalesced,
// several independent rows in flight per thread), the <= 8 column values of a row live in registers
// for the whole sweep, and the per-column sums are reduced block-wide once per step.  The cross-CTA
// exchange is a single hop: every CTA reads all G partials of the <= 8 columns (G*8 LL packets).
#define TALL_B QRDM_TALL_B
// MG = true: the rows of the panel live on several GPUs (1-D block-row sharding).  Same kernel, but
// the per-column exchange has a second level that crosses NVLink INSIDE the kernel: CTA 0 (the rank leader) totals
// the rank's G partial packets, stores the <= 8 rank sums and the pivot-row entries it owns (zeros otherwise) as LL
// packets into slot [rank] of EVERY rank's receive buffer (k_peer.cu), and every CTA of every rank then polls the
// nranks slots of its own buffer and adds them in rank order.  Same packets, same order everywhere => all CTAs of
// all ranks compute bit-identical tau / beta / stop decisions; no kernel boundary, no ncclAllReduce per column
// (the round-1 sharded panel paid one launch + one 1-KB ncclAllReduce per column: ~70 us x 512 columns on C4).
// Row bookkeeping: this rank holds global rows [row0, row0 + m); lr0 = first local row of the sub-panel, goff = the
// panel-relative index of that row (0 on the rank that owns the diagonal block).  Single GPU: lr0 = j, goff = 0.
// SM = true: the CTA's rows of the 8-column sub-panel fit in shared memory (rpc * 64 bytes <= 200 KB: up to ~3200 rows
// per CTA, 470,000 rows per GPU — every rank of configs[3] from 4 GPUs on): the slab is loaded once, the 8 sweeps run
// out of shared memory and it is written back once, instead of streaming the remaining columns through L2 / HBM for
// every column (16 column passes instead of 64).
template <bool MG, bool SM>
__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel_tall(qrdm_prob P, int rpc, unsigned epoch, PeerCtx pc) {
  extern __shared__ __align__(16) double tall_slab[];  // SM: [TALL_B][lds]
  __shared__ double S_[TALL_B], rowv[TALL_B];
  __shared__ double sacc[PANEL_WARPS][TALL_B];
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const QrdmGeom qg = qrdm_geom(P);
  const int j = qg.j, fjb = qg.fjb, sub_s = P.sub - 1;
  const int jmain = ctrl->j, fjb_main = ctrl->fjb;
  if (fjb <= 0) return;
  const bool forced = ctrl->forced != 0;  // fixed columns: plain Householder QR, no early stop
  const int lr0 = qrdm_jr(P, j);          // first local active row
  const int goff = P.row0 + lr0 - j;      // its index relative to the sub-panel's first row (>= 0)
  const int rows_l = P.m - lr0;           // local active rows
  const int rows = P.m_glob - j;          // active rows of the whole (global) panel
  const int lda = P.lda;
  const int G = gridDim.x, b = blockIdx.x;
  const int r0 = min(rows_l, b * rpc), r1 = min(rows_l, r0 + rpc), nr = r1 - r0;
  double* Ag = P.a + (size_t)j * lda + lr0 + r0;  // local row r of this CTA, sub-panel column c: Ag[c*lda + r]
  const int ldp = SM ? ((rpc + 1) & ~1) : lda;
  double* Ap = SM ? tall_slab : Ag;                // the working copy: Ap[c*ldp + r]
  if (SM) {
    for (int c = 0; c < fjb; ++c)
      for (int r = threadIdx.x; r < nr; r += PANEL_THREADS) Ap[(size_t)c * ldp + r] = Ag[(size_t)c * lda + r];
    __syncthreads();
  }
  LLPacket* part = reinterpret_cast<LLPacket*>(P.panel_part);  // [2][PANEL_MAXCTA][64]
  LLPacket* bcast = reinterpret_cast<LLPacket*>(P.panel_row);  // [2][128]
  const unsigned tag_base = epoch << 8;
  const int me = MG ? pc.rank : 0, NR = MG ? pc.nranks : 1;
  // cross-GPU exchanges are numbered by a counter that lives on the device and advances identically on every rank
  // (it counts the exchanges actually performed — the DM early stop makes that data dependent): exchange x uses slot
  // parity x & 1 and tag x + 1, so consecutive exchanges never share a slot, within or across launches
  const unsigned px0 = MG ? *pc.xseq : 0u;

  // block-wide sums of acc[0..TALL_B) -> LL packets part[buf][b][jj] for jj in [lo, fjb)
  auto publish = [&](double (&acc)[TALL_B], int buf, int lo, unsigned tag) {
#pragma unroll
    for (int c = 0; c < TALL_B; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < TALL_B; ++c) sacc[wid][c] = acc[c];
    }
    __syncthreads();
    if (tid < TALL_B && tid >= lo && tid < fjb) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < PANEL_WARPS; ++w) t += sacc[w][tid];
      ll_store(&part[((size_t)buf * QRDM_PANEL_MAXCTA + b) * 64 + tid], t, tag);
    }
    __syncthreads();
  };

  {  // dot products of column 0 with every column (rows below the diagonal); row 0 is the pivot row
    double acc[TALL_B];
#pragma unroll
    for (int c = 0; c < TALL_B; ++c) acc[c] = 0.0;
    for (int r = tid; r < nr; r += PANEL_THREADS) {
      const int R = goff + r0 + r;
      double v[TALL_B];
#pragma unroll
      for (int c = 0; c < TALL_B; ++c) v[c] = c < fjb ? Ap[(size_t)c * ldp + r] : 0.0;
      if (R > 0) {
#pragma unroll
        for (int c = 0; c < TALL_B; ++c) acc[c] = fma(v[0], v[c], acc[c]);
      } else {
#pragma unroll
        for (int c = 0; c < TALL_B; ++c)
          if (c < fjb) ll_store(&bcast[64 + c], v[c], tag_base + 1);
      }
    }
    publish(acc, 0, 0, tag_base + 1);
  }

  // squared stop threshold (carried across the sub-panels in ctrl->tall_thres): the test runs on ||x||^2 and the
  // reflector needs ONE square root, beta^2 = alpha^2 + ||x||^2 — the same scalar chain as k_panel_reg (the
  // sqrt + hypot pair of the first version cost more than the sweep of a slab-resident sub-panel)
  const bool cont = ctrl->micro_t > 0;  // nb > 64: the panel continues a wider block (k_wide.cu)
  double thres2 = (sub_s == 0) ? (cont ? ctrl->micro_thres2 : P.thres0 * P.thres0) : ctrl->tall_thres;
  int k = fjb;
  for (int i = 0; i < fjb; ++i) {
    const int cur = i & 1, nxt = cur ^ 1;
    const unsigned tag = tag_base + i + 1;
    // ---- gather: warp jj totals column jj over all CTAs (lanes <-> CTAs, fixed order); pivot row ----
    if (!MG) {
      if (wid < TALL_B && wid >= i && wid < fjb) {
        double v = 0.0;
        v += ll_gather_sum(&part[(size_t)cur * QRDM_PANEL_MAXCTA * 64 + wid], 64, lane, G, tag);
        v = warp_sum(v);
        if (lane == 0) { S_[wid] = v; rowv[wid] = ll_load(&bcast[cur * 128 + 64 + wid], tag); }
      }
    } else if (wid < TALL_B && wid >= i && wid < fjb) {
      const unsigned ptag = px0 + (unsigned)i + 1u;
      const int pcur = (int)((px0 + (unsigned)i) & 1u);
      if (b == 0) {
        // rank leader: total of this rank's CTAs, then one packet per (column, peer) over NVLink.  Only the rank that
        // owns global row j + i has a pivot-row entry; the others contribute an exact 0.
        double v = ll_gather_sum(&part[(size_t)cur * QRDM_PANEL_MAXCTA * 64 + wid], 64, lane, G, tag);
        v = warp_sum(v);
        const int grow = j + i;  // global row of the pivot
        double rv = 0.0;
        if (grow >= P.row0 && grow < P.row0 + P.m) {
          if (lane == 0) rv = ll_load(&bcast[cur * 128 + 64 + wid], tag);
          rv = __shfl_sync(0xffffffffu, rv, 0);
        }
        if (lane < NR) ll_store(peer_panel_slot(pc.recv[lane], pcur, me, wid), v, ptag);
        else if (lane >= 8 && lane < 8 + NR) ll_store(peer_panel_slot(pc.recv[lane - 8], pcur, me, 64 + wid), rv, ptag);
      }
      // everybody: the nranks rank sums from the local receive buffer, added in rank order
      double x = 0.0;
      if (lane < NR) x = ll_load(peer_panel_slot(pc.recv[me], pcur, lane, wid), ptag);
      else if (lane >= 8 && lane < 8 + NR) x = ll_load(peer_panel_slot(pc.recv[me], pcur, lane - 8, 64 + wid), ptag);
      double sv = 0.0, rvs = 0.0;
      for (int r = 0; r < NR; ++r) {
        sv += __shfl_sync(0xffffffffu, x, r);
        rvs += __shfl_sync(0xffffffffu, x, 8 + r);
      }
      if (lane == 0) { S_[wid] = sv; rowv[wid] = rvs; }
    }
    __syncthreads();
    // ---- reflector scalars (dlarfg_mia) ----
    const double alpha = rowv[i], xn2 = S_[i];
    const int len = rows - i;
    double tau = 0.0, beta = alpha, scale = 1.0;
    if (len > 1) {
      if ((sub_s + i > 0 || cont) && xn2 < thres2 && !forced) { k = i; break; }  // DM early stop (never for fixed columns)
      if (xn2 != 0.0) {
        const double h = sqrt(fma(alpha, alpha, xn2));
        beta = (alpha >= 0.0) ? -h : h;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
    }
    if (sub_s + i == 0 && !cont && fjb_main > 1 && P.tau_ > 0.0) { const double th = P.tau_ * fabs(beta); thres2 = th * th; }
    if (b == 0 && tid == 0) {
      P.tau[j + i] = tau;
      if (tau != tau && ctrl->err == 0) ctrl->err = -8;
    }
    double w[TALL_B];
#pragma unroll
    for (int c = 0; c < TALL_B; ++c) w[c] = (c > i && c < fjb) ? tau * (rowv[c] + S_[c] * scale) : 0.0;
    const bool last = i + 1 >= fjb;
    // ---- sweep: one pass over the rows, all remaining columns of a row in registers ----
    double acc[TALL_B];
#pragma unroll
    for (int c = 0; c < TALL_B; ++c) acc[c] = 0.0;
    auto do_row = [&](int r) {  // general row (may be the pivot row, the next pivot row, or above the diagonal)
      const int R = goff + r0 + r;
      if (R < i) return;
      double v = 1.0;
      if (R > i) {
        v = Ap[(size_t)i * ldp + r];
        if (tau != 0.0) { v *= scale; Ap[(size_t)i * ldp + r] = v; }
      } else {
        Ap[(size_t)i * ldp + r] = beta;
      }
      if (last) return;
      double pv[TALL_B];
#pragma unroll
      for (int c = 0; c < TALL_B; ++c)
        if (c > i && c < fjb) pv[c] = Ap[(size_t)c * ldp + r];
      double x1 = 0.0;
#pragma unroll
      for (int c = 0; c < TALL_B; ++c) {
        if (c > i && c < fjb) {
          pv[c] = fma(-v, w[c], pv[c]);
          Ap[(size_t)c * ldp + r] = pv[c];
          if (c == i + 1) x1 = pv[c];
        }
      }
      if (R > i + 1) {
#pragma unroll
        for (int c = 0; c < TALL_B; ++c)
          if (c > i && c < fjb) acc[c] = fma(x1, pv[c], acc[c]);
      } else if (R == i + 1) {
#pragma unroll
        for (int c = 0; c < TALL_B; ++c)
          if (c > i && c < fjb) ll_store(&bcast[nxt * 128 + 64 + c], pv[c], tag + 1);
      }
    };
    // Two rows per trip whenever both lie strictly below the next pivot row (all but the first few rows of the rank
    // that owns the diagonal block): 2 x (8 - i) independent loads in flight per thread before the first FMA — the
    // one-row loop left the kernel latency-bound at ~2 TB/s (profiles/r02_launches_c4_summary.txt).  Same per-row
    // arithmetic and the same per-thread accumulation order (rows ascending), so results are bit-identical.
    int r = tid;
    for (; !SM && r + PANEL_THREADS < nr; r += 2 * PANEL_THREADS) {
      const int rb = r + PANEL_THREADS;
      if (goff + r0 + r <= i + 1 || last) { do_row(r); do_row(rb); continue; }
      double va = Ap[(size_t)i * ldp + r], vb = Ap[(size_t)i * ldp + rb];
      double pa[TALL_B], pb[TALL_B];
#pragma unroll
      for (int c = 0; c < TALL_B; ++c)
        if (c > i && c < fjb) { pa[c] = Ap[(size_t)c * ldp + r]; pb[c] = Ap[(size_t)c * ldp + rb]; }
      if (tau != 0.0) { va *= scale; vb *= scale; Ap[(size_t)i * ldp + r] = va; Ap[(size_t)i * ldp + rb] = vb; }
      double xa = 0.0, xb = 0.0;
#pragma unroll
      for (int c = 0; c < TALL_B; ++c) {
        if (c > i && c < fjb) {
          pa[c] = fma(-va, w[c], pa[c]);
          pb[c] = fma(-vb, w[c], pb[c]);
          Ap[(size_t)c * ldp + r] = pa[c];
          Ap[(size_t)c * ldp + rb] = pb[c];
          if (c == i + 1) { xa = pa[c]; xb = pb[c]; }
        }
      }
#pragma unroll
      for (int c = 0; c < TALL_B; ++c)
        if (c > i && c < fjb) { acc[c] = fma(xa, pa[c], acc[c]); acc[c] = fma(xb, pb[c], acc[c]); }
    }
    for (; r < nr; r += PANEL_THREADS) do_row(r);
    if (last) break;
    publish(acc, nxt, i + 1, tag + 1);
  }
  __syncthreads();
  if (b == 0 && tid == 0) {
    const int tk = (sub_s == 0 ? 0 : ctrl->tall_k) + k;
    ctrl->sub_k = k;
    ctrl->tall_k = tk;
    ctrl->tall_done = (k < fjb) ? 1 : 0;
    if (k < fjb) ctrl->tall_stop_s = sub_s;
    ctrl->tall_thres = thres2;
    ctrl->micro_thres2 = thres2;
    ctrl->fjb_cmp = tk;
    if (MG) *pc.xseq = px0 + (unsigned)(k < fjb ? k + 1 : fjb);  // exchanges performed by this launch
  }
  if (SM) {  // the factored slab goes back to global memory once (the block-wide barrier above ordered the last sweep)
    for (int c = 0; c < fjb; ++c)
      for (int r = tid; r < nr; r += PANEL_THREADS) Ag[(size_t)c * lda + r] = Ap[(size_t)c * ldp + r];
  }
  // ---- clean copy of the sub-panel's reflectors into Vc columns voff .. voff + 7 (Vc is indexed by LOCAL row) ----
  const int kpad = min(TALL_B, 64 - qg.voff);
  const int jal = qrdm_jr(P, jmain) & ~(QRDM_ROWALIGN - 1);
  for (int q = 0; q < kpad; ++q) {
    double* vcol = P.vc + (size_t)(qg.voff + q) * P.ldv + lr0 + r0;
    for (int r = tid; r < nr; r += PANEL_THREADS) {
      const int R = goff + r0 + r;
      double v = 0.0;
      if (q < k) v = (R > q) ? Ap[(size_t)q * ldp + r] : (R == q ? 1.0 : 0.0);
      vcol[r] = v;
    }
    if (b == 0)
      for (int g = jal + tid; g < lr0; g += PANEL_THREADS) P.vc[(size_t)(qg.voff + q) * P.ldv + g] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Grouped sub-panel kernel of the blocked tall panel: the whole 8-column sub-panel with ONE exchange (normally).
// Same algebra as k_panel_grp with S = 8 = all columns of the sub-panel: pass 1 streams the CTA's rows once and
// accumulates the upper triangle of the 8 x 8 Gram matrix of the rows below the first pivot row (36 sums) and
// publishes the 8 pivot rows; one exchange (within the GPU: all-gather of the G partials; MG: + one LL exchange with
// the other GPUs over NVLink peer memory — ONE per sub-panel instead of one per column); the scalar recurrence on
// the replicated 8 x 8 numbers gives all tau / beta / scale / w; pass 2 streams the rows again and applies the
// reflectors to the row's 8 values in registers (read once, written once).  The per-column kernel above made 8
// read+write sweeps over the remaining columns and 8 exchanges.  The accuracy guard of k_panel_grp applies: a column
// whose ||x||^2 falls below 2^-7 of its sum at the start of the round ends the round; the kernel then makes another
// round (pass 1 / exchange / recurrence / pass 2) from that column on — every CTA (and every rank) takes the same
// decision from the same numbers, so the number of exchanges is the same everywhere.
#define TG_NSUM 36  // upper triangle of the 8 x 8 Gram block: index of (u, c), u <= c
__device__ __forceinline__ constexpr int tg_idx(int u, int c) { return u * TALL_B - u * (u - 1) / 2 + (c - u); }
template <bool MG, bool SM>
__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel_tall_grp(qrdm_prob P, int rpc, unsigned epoch, PeerCtx pc) {
  extern __shared__ __align__(16) double tall_slab[];  // SM: [TALL_B][lds]
  __shared__ double sacc[PANEL_WARPS][TG_NSUM];
  __shared__ double Gs[TG_NSUM], Rs[TALL_B][TALL_B];            // totals delivered by the exchange
  __shared__ double Mb[2][TALL_B][TALL_B], Rb[2][TALL_B][TALL_B];  // recurrence state, double-buffered by step parity
  __shared__ double Wt[TALL_B][TALL_B], SC[TALL_B][4];            // per step: w_t[c], {tau, beta, seff}
  __shared__ int ginfo[2];
  __shared__ double gth2;
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const QrdmGeom qg = qrdm_geom(P);
  const int j = qg.j, fjb = qg.fjb, sub_s = P.sub - 1;
  const int jmain = ctrl->j, fjb_main = ctrl->fjb;
  if (fjb <= 0) return;
  const bool forced = ctrl->forced != 0;
  const int lr0 = qrdm_jr(P, j);
  const int goff = P.row0 + lr0 - j;
  const int rows_l = P.m - lr0;
  const int rows = P.m_glob - j;
  const int lda = P.lda;
  const int G = gridDim.x, b = blockIdx.x;
  const int r0 = min(rows_l, b * rpc), r1 = min(rows_l, r0 + rpc), nr = r1 - r0;
  double* Ag = P.a + (size_t)j * lda + lr0 + r0;
  const int ldp = SM ? ((rpc + 1) & ~1) : lda;
  double* Ap = SM ? tall_slab : Ag;
  if (SM) {
    for (int c = 0; c < fjb; ++c)
      for (int r = threadIdx.x; r < nr; r += PANEL_THREADS) Ap[(size_t)c * ldp + r] = Ag[(size_t)c * lda + r];
    __syncthreads();
  }
  LLPacket* part = reinterpret_cast<LLPacket*>(P.panel_part);  // [2][PANEL_MAXCTA][64]
  LLPacket* bcast = reinterpret_cast<LLPacket*>(P.panel_row);  // [2][128]: pivot rows at 64 + u * 8 + c
  const unsigned tag_base = epoch << 8;
  const int me = MG ? pc.rank : 0, NR = MG ? pc.nranks : 1;
  const unsigned px0 = MG ? *pc.xseq : 0u;
  const bool cont = ctrl->micro_t > 0;
  double thres2 = (sub_s == 0) ? (cont ? ctrl->micro_thres2 : P.thres0 * P.thres0) : ctrl->tall_thres;
  int k = fjb, i0 = 0, round = 0;
  bool stopped = false, vc_done = false;
  while (i0 < fjb && !stopped) {
    const int cur = round & 1;
    const unsigned tag = tag_base + round + 1;
    // ---- pass 1: Gram block of the columns >= i0 over the rows below pivot row i0; the pivot rows i0 .. fjb-1 ----
    {
      double acc[TG_NSUM];
#pragma unroll
      for (int x = 0; x < TG_NSUM; ++x) acc[x] = 0.0;
      auto row1 = [&](int r) {
        const int R = goff + r0 + r;
        if (R < i0) return;
        double v[TALL_B];
#pragma unroll
        for (int c = 0; c < TALL_B; ++c) v[c] = (c >= i0 && c < fjb) ? Ap[(size_t)c * ldp + r] : 0.0;
        if (R > i0) {
#pragma unroll
          for (int u = 0; u < TALL_B; ++u)
#pragma unroll
            for (int c = u; c < TALL_B; ++c) acc[tg_idx(u, c)] = fma(v[u], v[c], acc[tg_idx(u, c)]);
        }
        if (R < fjb) {
#pragma unroll
          for (int c = 0; c < TALL_B; ++c)
            if (c >= i0 && c < fjb) ll_store(&bcast[cur * 128 + 64 + R * TALL_B + c], v[c], tag);
        }
      };
      // two rows per trip below the pivot rows (16 independent loads in flight per thread; same accumulation order)
      int r = tid;
      for (; !SM && r + PANEL_THREADS < nr; r += 2 * PANEL_THREADS) {
        const int rb = r + PANEL_THREADS;
        if (goff + r0 + r < TALL_B) { row1(r); row1(rb); continue; }
        double va[TALL_B], vb[TALL_B];
#pragma unroll
        for (int c = 0; c < TALL_B; ++c) {
          const bool on = c >= i0 && c < fjb;
          va[c] = on ? Ap[(size_t)c * ldp + r] : 0.0;
          vb[c] = on ? Ap[(size_t)c * ldp + rb] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < TALL_B; ++u)
#pragma unroll
          for (int c = u; c < TALL_B; ++c) {
            acc[tg_idx(u, c)] = fma(va[u], va[c], acc[tg_idx(u, c)]);
            acc[tg_idx(u, c)] = fma(vb[u], vb[c], acc[tg_idx(u, c)]);
          }
      }
      for (; r < nr; r += PANEL_THREADS) row1(r);
#pragma unroll
      for (int x = 0; x < TG_NSUM; ++x) acc[x] = warp_sum(acc[x]);
      if (lane == 0) {
#pragma unroll
        for (int x = 0; x < TG_NSUM; ++x) sacc[wid][x] = acc[x];
      }
      __syncthreads();
      if (tid < TG_NSUM) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PANEL_WARPS; ++w) t += sacc[w][tid];
        ll_store(&part[((size_t)cur * QRDM_PANEL_MAXCTA + b) * 64 + tid], t, tag);
      }
    }
    // ---- exchange: totals of the 36 sums and the 8 x 8 pivot rows, the same numbers in every CTA (and rank) ----
    if (!MG) {
      for (int x = wid; x < TG_NSUM; x += PANEL_WARPS) {
        double v = ll_gather_sum(&part[(size_t)cur * QRDM_PANEL_MAXCTA * 64 + x], 64, lane, G, tag);
        v = warp_sum(v);
        if (lane == 0) Gs[x] = v;
      }
      if (tid >= 256 && tid < 256 + 64) {
        const int u = (tid - 256) >> 3, c = (tid - 256) & 7;
        double rv = 0.0;
        if (u >= i0 && u < fjb && c >= i0 && c < fjb && u < rows) rv = ll_load(&bcast[cur * 128 + 64 + u * TALL_B + c], tag);
        Rs[u][c] = rv;
      }
    } else {
      const unsigned ptag = px0 + (unsigned)round + 1u;
      const int pcur = (int)((px0 + (unsigned)round) & 1u);
      if (b == 0) {
        // rank leader: totals of this rank's CTAs and the pivot-row entries this rank owns (exact zeros otherwise),
        // one packet per (value, peer) over NVLink
        for (int x = wid; x < TG_NSUM; x += PANEL_WARPS) {
          double v = ll_gather_sum(&part[(size_t)cur * QRDM_PANEL_MAXCTA * 64 + x], 64, lane, G, tag);
          v = warp_sum(v);
          if (lane < NR) ll_store(peer_panel_slot(pc.recv[lane], pcur, me, x), v, ptag);
        }
        for (int x = wid; x < 64; x += PANEL_WARPS) {
          const int u = x >> 3, c = x & 7;
          const int grow = j + u;  // global row of pivot row u
          double rv = 0.0;
          if (u >= i0 && u < fjb && c >= i0 && c < fjb && grow >= P.row0 && grow < P.row0 + P.m) {
            if (lane == 0) rv = ll_load(&bcast[cur * 128 + 64 + u * TALL_B + c], tag);
            rv = __shfl_sync(0xffffffffu, rv, 0);
          }
          if (lane < NR) ll_store(peer_panel_slot(pc.recv[lane], pcur, me, 64 + x), rv, ptag);
        }
      }
      // everybody: the nranks contributions from the local receive buffer, added in rank order
      for (int x = wid; x < TG_NSUM + 64; x += PANEL_WARPS) {
        const int e = x < TG_NSUM ? x : 64 + (x - TG_NSUM);
        double v = 0.0;
        if (lane < NR) v = ll_load(peer_panel_slot(pc.recv[me], pcur, lane, e), ptag);
        double t = 0.0;
        for (int r = 0; r < NR; ++r) t += __shfl_sync(0xffffffffu, v, r);
        if (lane == 0) {
          if (x < TG_NSUM) Gs[x] = t;
          else Rs[(x - TG_NSUM) >> 3][(x - TG_NSUM) & 7] = t;
        }
      }
    }
    __syncthreads();
    // ---- scalar recurrence on the replicated 8 x 8 block: threads (u, c) of warps 0-1, one 64-thread barrier per step ----
    if (tid < 64) {
      const int u = tid >> 3, c = tid & 7;
      Mb[0][u][c] = (u >= i0 && c >= i0 && u < fjb && c < fjb) ? Gs[u <= c ? tg_idx(u, c) : tg_idx(c, u)] : 0.0;
      Rb[0][u][c] = Rs[u][c];
      bar_sync_64();
      const double d0 = Mb[0][c][c];  // (read before anything is overwritten; used as the guard reference of column c)
      int t = i0;
      bool stop = false;
      for (; t < fjb; ++t) {
        const int pb = (t - i0) & 1;
        const double alpha = Rb[pb][t][t], xn2 = Mb[pb][t][t];
        const double d0t = __shfl_sync(0xffffffffu, d0, t & 7, 8);  // lanes c = t of every group of 8 hold it
        if (t > i0 && !(xn2 >= GRP_GUARD * d0t)) break;
        const int len = rows - t;
        double tau = 0.0, beta = alpha, scale = 1.0;
        if (len > 1) {
          if ((sub_s + t > 0 || cont) && xn2 < thres2 && !forced) { stop = true; break; }  // DM early stop
          if (xn2 != 0.0) {
            const double h = sqrt(fma(alpha, alpha, xn2));
            beta = (alpha >= 0.0) ? -h : h;
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
          }
        }
        if (sub_s + t == 0 && !cont && fjb_main > 1 && P.tau_ > 0.0) { const double th = P.tau_ * fabs(beta); thres2 = th * th; }
        const double sc = tau != 0.0 ? scale : 0.0, vs = tau != 0.0 ? scale : 1.0, vv = sc * sc * xn2;
        if (tid == 0) {
          SC[t][0] = tau; SC[t][1] = beta; SC[t][2] = vs;
          if (b == 0) {
            P.tau[j + t] = tau;
            if (tau != tau && ctrl->err == 0) ctrl->err = -8;
          }
        }
        const double w_c = c > t ? grp_w(tau, scale, Mb[pb][t][c], Rb[pb][t][c]) : 0.0;
        if (u == 0) Wt[t][c] = w_c;
        if (u > t && c > t && t + 1 < fjb) {
          const double w_u = grp_w(tau, scale, Mb[pb][t][u], Rb[pb][t][u]);
          const double sv_c = __dmul_rn(sc, Mb[pb][t][c]), sv_u = __dmul_rn(sc, Mb[pb][t][u]);
          const double vr1 = Rb[pb][t + 1][t] * vs, vr_u = Rb[pb][u][t] * vs;
          const double r1_c = __fma_rn(-vr1, w_c, Rb[pb][t + 1][c]), r1_u = __fma_rn(-vr1, w_u, Rb[pb][t + 1][u]);
          Rb[pb ^ 1][u][c] = __fma_rn(-vr_u, w_c, Rb[pb][u][c]);
          Mb[pb ^ 1][u][c] = grp_m_update(Mb[pb][u][c], w_c, sv_c, r1_c, w_u, sv_u, r1_u, vv);
        }
        bar_sync_64();
      }
      if (tid == 0) { ginfo[0] = t - i0; ginfo[1] = stop ? 1 : 0; gth2 = thres2; }
    }
    __syncthreads();
    const int nsteps = ginfo[0];
    stopped = ginfo[1] != 0;
    thres2 = gth2;
    // ---- pass 2: apply the round's reflectors to every row of the slab (the row's values in registers).  In the
    // common case — one round took the whole sub-panel — the clean copy Vc is written from the same registers ----
    const bool emit_vc = round == 0 && nsteps == fjb && !SM;
    if (emit_vc) vc_done = true;
    if (nsteps > 0) {
      const int kpad_e = min(TALL_B, 64 - qg.voff);
      // (the small tables stay in shared memory: 28 w's + 16 scalars do not fit next to the row in 128 registers)
      auto row2 = [&](int r, double (&v)[TALL_B]) {
        const int R = goff + r0 + r;
#pragma unroll
        for (int t = 0; t < TALL_B; ++t) {
          if (t >= i0 && t < i0 + nsteps && R >= t) {
            const double vt = R > t ? v[t] * SC[t][2] : 1.0;
#pragma unroll
            for (int c = 0; c < TALL_B; ++c)
              if (c > t) v[c] = fma(-vt, Wt[t][c], v[c]);
            v[t] = R > t ? vt : SC[t][1];
          }
        }
#pragma unroll
        for (int c = 0; c < TALL_B; ++c)
          if (c >= i0 && c < fjb) Ap[(size_t)c * ldp + r] = v[c];
        if (emit_vc) {
#pragma unroll
          for (int q = 0; q < TALL_B; ++q)
            if (q < kpad_e) P.vc[(size_t)(qg.voff + q) * P.ldv + lr0 + r0 + r] = q < fjb ? (R > q ? v[q] : (R == q ? 1.0 : 0.0)) : 0.0;
        }
      };
      int r = tid;
      for (; !SM && r + PANEL_THREADS < nr; r += 2 * PANEL_THREADS) {
        const int rb = r + PANEL_THREADS;
        if (goff + r0 + r < i0) {  // (rows above the round's first pivot row: only at the top of the diagonal block)
          for (int q = 0; q < 2; ++q) {
            const int rr = q ? rb : r;
            if (goff + r0 + rr < i0) continue;
            double v[TALL_B];
#pragma unroll
            for (int c = 0; c < TALL_B; ++c) v[c] = (c >= i0 && c < fjb) ? Ap[(size_t)c * ldp + rr] : 0.0;
            row2(rr, v);
          }
          continue;
        }
        double va[TALL_B], vb[TALL_B];
#pragma unroll
        for (int c = 0; c < TALL_B; ++c) {
          const bool on = c >= i0 && c < fjb;
          va[c] = on ? Ap[(size_t)c * ldp + r] : 0.0;
          vb[c] = on ? Ap[(size_t)c * ldp + rb] : 0.0;
        }
        row2(r, va);
        row2(rb, vb);
      }
      for (; r < nr; r += PANEL_THREADS) {
        if (goff + r0 + r < i0) continue;
        double v[TALL_B];
#pragma unroll
        for (int c = 0; c < TALL_B; ++c) v[c] = (c >= i0 && c < fjb) ? Ap[(size_t)c * ldp + r] : 0.0;
        row2(r, v);
      }
    }
    i0 += nsteps;
    if (stopped) k = i0;
    ++round;
    __syncthreads();  // the small tables are rewritten by the next round
  }
  if (b == 0 && tid == 0) {
    const int tk = (sub_s == 0 ? 0 : ctrl->tall_k) + k;
    ctrl->sub_k = k;
    ctrl->tall_k = tk;
    ctrl->tall_done = (k < fjb) ? 1 : 0;
    if (k < fjb) ctrl->tall_stop_s = sub_s;
    ctrl->tall_thres = thres2;
    ctrl->micro_thres2 = thres2;
    ctrl->fjb_cmp = tk;
    if (MG) *pc.xseq = px0 + (unsigned)round;  // exchanges performed by this launch
  }
  if (SM) {
    for (int c = 0; c < fjb; ++c)
      for (int r = tid; r < nr; r += PANEL_THREADS) Ag[(size_t)c * lda + r] = Ap[(size_t)c * ldp + r];
  }
  // ---- clean copy of the sub-panel's reflectors into Vc columns voff .. voff + 7 (Vc is indexed by LOCAL row) ----
  const int kpad = min(TALL_B, 64 - qg.voff);
  const int jal = qrdm_jr(P, jmain) & ~(QRDM_ROWALIGN - 1);
  for (int q = 0; q < kpad; ++q) {
    double* vcol = P.vc + (size_t)(qg.voff + q) * P.ldv + lr0 + r0;
    for (int r = tid; r < nr && !vc_done; r += PANEL_THREADS) {
      const int R = goff + r0 + r;
      double v = 0.0;
      if (q < k) v = (R > q) ? Ap[(size_t)q * ldp + r] : (R == q ? 1.0 : 0.0);
      vcol[r] = v;
    }
    if (b == 0)
      for (int g = jal + tid; g < lr0; g += PANEL_THREADS) P.vc[(size_t)(qg.voff + q) * P.ldv + g] = 0.0;
  }
}

// LL tags = (epoch << 8) + step.  Three disjoint epoch ranges share the exchange buffers: [1, 2^21) blocked tall
// panel, [2^21, 2^22) smem/global panel, [2^22, 2^23) register panel.  When a range wraps, packets of its previous
// cycle could carry tags that match again, so the buffers are cleared (stream-ordered, between two panel launches no
// packet is live) before the range restarts.
static unsigned panel_next_epoch(unsigned& e, unsigned lo, unsigned hi, const qrdm_prob* p, void* stream) {
  if (e + 1 >= hi) {
    cudaMemsetAsync(p->panel_part, 0, (size_t)16 * 2 * QRDM_PANEL_MAXCTA * 64, (cudaStream_t)stream);
    cudaMemsetAsync(p->panel_row, 0, (size_t)16 * 2 * 128, (cudaStream_t)stream);
    e = lo;
  } else {
    ++e;
  }
  return e;
}
static unsigned g_epoch_tall = 0, g_epoch_plain = 0x200000, g_epoch_reg = 0x400000;
// the grouped register panel has its own packet regions (behind the per-column ones) and its own epoch range [2^23, 2^24)
static unsigned g_epoch_grp = 0x800000;
static unsigned panel_grp_epoch(const qrdm_prob* p, void* stream) {
  if (g_epoch_grp + 1 >= 0xffffffu) {
    cudaMemsetAsync(reinterpret_cast<LLPacket*>(p->panel_part) + QRDM_PANEL_PART_PKTS, 0, (size_t)16 * QRDM_PANEL_GPART_PKTS, (cudaStream_t)stream);
    cudaMemsetAsync(reinterpret_cast<LLPacket*>(p->panel_row) + QRDM_PANEL_ROW_PKTS, 0, (size_t)16 * QRDM_PANEL_GROW_PKTS, (cudaStream_t)stream);
    g_epoch_grp = 0x800000;
  } else {
    ++g_epoch_grp;
  }
  return g_epoch_grp;
}
// dynamic shared memory of the slab-resident sub-panel kernel for `rpc` rows per CTA, or 0 when the slab does not fit
#define TALL_SLAB_CAP (200 * 1024)
// QRDM_TALL_S=1: the per-column sub-panel kernel (k_panel_tall); default: the grouped one (k_panel_tall_grp)
static bool tall_grouped() {
  const char* e = getenv("QRDM_TALL_S");
  return !(e && atoi(e) == 1);
}
static size_t tall_slab_bytes(int rpc) {
  static int attr_gen = -1;
  static const char* e_off = getenv("QRDM_PANEL_TALL_SM");  // experiment switch: 0 = always stream from global memory
  if (e_off && atoi(e_off) == 0) return 0;
  const size_t bytes = (size_t)((rpc + 1) & ~1) * QRDM_TALL_B * sizeof(double);
  if (bytes == 0 || bytes > TALL_SLAB_CAP) return 0;
  if (attr_gen != qrdm_rt_device_generation()) {
    cudaFuncSetAttribute(k_panel_tall<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TALL_SLAB_CAP);
    cudaFuncSetAttribute(k_panel_tall<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TALL_SLAB_CAP);
    cudaFuncSetAttribute(k_panel_tall_grp<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TALL_SLAB_CAP);
    cudaFuncSetAttribute(k_panel_tall_grp<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TALL_SLAB_CAP);
    attr_gen = qrdm_rt_device_generation();
  }
  return bytes;
}
static unsigned qrdm_panel_tall_epoch(const qrdm_prob* p, void* stream) { return panel_next_epoch(g_epoch_tall, 1, 0x200000, p, stream); }

// Row-sharded panel: every sharded panel is run BLOCKED — 8-column sub-panels by k_panel_tall<true>
// with the cross-GPU exchange inside the kernel; the caller (dgeqrdm_host.c) applies each sub-panel's block
// reflector to the rest of the panel with the skinny kernels + one small all-reduce.  One launch per sub-panel.
extern "C" int qrdm_k_panel_tall_mg(const qrdm_prob* p, int j_host, void* stream) {
  const PeerCtx* pc = qrdm_peer_ctx();
  if (!pc || p->sub <= 0) return (int)cudaErrorInvalidValue;
  const int jsub = j_host + p->sub - 1;
  int lr0 = jsub - p->row0;
  lr0 = lr0 < 0 ? 0 : (lr0 > p->m ? p->m : lr0);
  const int rows_l = p->m - lr0;
  const int gmax = p->sm_count < QRDM_PANEL_MAXCTA ? p->sm_count : QRDM_PANEL_MAXCTA;
  int Gs = (rows_l + 1023) / 1024;
  if (Gs > gmax) Gs = gmax;
  if (Gs < 1) Gs = 1;  // a rank without active rows still runs its leader CTA: it contributes zeros
  int rpcs = (rows_l + Gs - 1) / Gs;
  unsigned epoch = qrdm_panel_tall_epoch(p, stream);
  qrdm_prob prob_s = *p;
  PeerCtx pcv = *pc;
  void* args_s[] = {(void*)&prob_s, (void*)&rpcs, (void*)&epoch, (void*)&pcv};
  const size_t slab = tall_slab_bytes(rpcs);
  const bool grp = tall_grouped();
  void* fn_sm = grp ? (void*)k_panel_tall_grp<true, true> : (void*)k_panel_tall<true, true>;
  void* fn_gl = grp ? (void*)k_panel_tall_grp<true, false> : (void*)k_panel_tall<true, false>;
  cudaError_t es = slab ? cudaLaunchCooperativeKernel(fn_sm, dim3(Gs), dim3(PANEL_THREADS), args_s, slab, (cudaStream_t)stream)
                        : cudaLaunchCooperativeKernel(fn_gl, dim3(Gs), dim3(PANEL_THREADS), args_s, 0, (cudaStream_t)stream);
  ++g_qrdm_launches;
  return es == cudaSuccess ? 0 : (int)es;
}

static int cl_max_ctas = -1;  // CTAs that fit as whole clusters of PANEL_CL (per device; queried on first use)
// Rows per CTA (32 x RI) and exchange mode by a measured cost model (us per column, B200, square Gaussian
// inputs): ~0.3 per RI (the in-CTA sweep) + 0.0045 per CTA (skew / fan-in of the LL exchange); the one-hop
// exchange (<= 32 CTAs) saves 0.45 (1000 rows: 2.90 -> 2.44), the cluster exchange 0.3 (8192 rows: 3.58 -> 3.27).
static void panel_plan(int rows, int gmax, int* per_out, int* mode_out) {
  static const char* e_ag = getenv("QRDM_PANEL_AG");  // experiment switch: 0 disables the one-hop exchange
  const bool ag_ok = !(e_ag && atoi(e_ag) == 0);
  const int clmax = cl_max_ctas < 0 ? 0 : cl_max_ctas;
  int per = 256, mode = 0;
  double best = 1e30;
  for (int ri = 1; ri <= 8; ri *= 2) {
    const int g = (rows + 32 * ri - 1) / (32 * ri);
    if (g > gmax) continue;
    const int gpad = (g + PANEL_CL - 1) / PANEL_CL * PANEL_CL;
    const int md = (ag_ok && g <= 32) ? 1 : (gpad <= clmax ? 2 : 0);
    const double cost = 0.30 * ri + 0.0045 * g - (md == 1 ? 0.45 : md == 2 ? 0.30 : 0.0);
    if (cost < best) { best = cost; per = 32 * ri; mode = md; }
  }
  const char* e = getenv("QRDM_PANEL_PER");  // experiment switch
  if (e) {
    const int v = atoi(e);
    if ((v == 32 || v == 64 || v == 128 || v == 256) && rows <= v * gmax) {
      per = v;
      const int g = (rows + v - 1) / v, gpad = (g + PANEL_CL - 1) / PANEL_CL * PANEL_CL;
      mode = (ag_ok && g <= 32) ? 1 : (gpad <= clmax ? 2 : 0);
    }
  }
  *per_out = per; *mode_out = mode;
}
// SMs the panel of `rows` rows will occupy (one CTA per SM): what the look-ahead's side stream cannot count on
extern "C" int qrdm_k_panel_ctas(const qrdm_prob* p, int rows) {
  const int gmax = p->sm_count < QRDM_PANEL_MAXCTA ? p->sm_count : QRDM_PANEL_MAXCTA;
  if (rows <= 0) return 0;
  if (rows > 256 * gmax) return gmax;
  int per = 256, mode = 0;
  panel_plan(rows, gmax, &per, &mode);
  const int g = (rows + per - 1) / per;
  return mode == 2 ? (g + PANEL_CL - 1) / PANEL_CL * PANEL_CL : g;
}

extern "C" int qrdm_k_panel(const qrdm_prob* p, int j_host, void* stream) {
  static int attr_gen = -1;  // per-device function attributes: re-applied when the library moves to another device
  static int rows_per_cta = 128;
  const int smem_cap = 200 * 1024;
  if (attr_gen != qrdm_rt_device_generation()) {
    cudaFuncSetAttribute(k_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
    const char* e = getenv("QRDM_PANEL_ROWS");
    if (e && atoi(e) >= 32) rows_per_cta = atoi(e);
    cl_max_ctas = -1;  // cluster occupancy is a per-device figure too
    attr_gen = qrdm_rt_device_generation();
  }
  const int rows = p->m - j_host;
  if (rows <= 0) return 0;
  const int gmax = p->sm_count < QRDM_PANEL_MAXCTA ? p->sm_count : QRDM_PANEL_MAXCTA;
  if (rows <= 256 * gmax && !getenv("QRDM_PANEL_NOREG")) {  // register-resident slabs, 32 ... 256 rows per CTA
    static const char* e_cl = getenv("QRDM_PANEL_CL");  // experiment switch: 0 disables the cluster exchange
    // clusters of PANEL_CL CTAs (MODE 2): how many CTAs can be co-resident as whole clusters (GPC granularity:
    // 33 clusters of 4 = 132 CTAs on a 148-SM B200, only 15 clusters of 8); queried once
    static cudaLaunchAttribute cl_attrs[2];
    if (cl_max_ctas < 0) {
      cl_max_ctas = 0;
      // The library never looks at its environment to guess whether a tool is attached (round 1 did, and profiled a
      // kernel it does not ship).  Nsight Compute's KERNEL replay re-runs a kernel that communicates through global
      // memory flags out of context; profile the cluster panel with a single-pass metric set or with
      // --replay-mode application (profiles/README), or force the two-hop kernel with QRDM_PANEL_CL=0.
      const int want = e_cl ? atoi(e_cl) : 1;
      if (want >= 1) {
        cudaLaunchConfig_t q;
        memset(&q, 0, sizeof(q));
        q.gridDim = dim3(PANEL_CL); q.blockDim = dim3(PANEL_THREADS);
        cl_attrs[0].id = cudaLaunchAttributeClusterDimension;
        cl_attrs[0].val.clusterDim.x = PANEL_CL; cl_attrs[0].val.clusterDim.y = 1; cl_attrs[0].val.clusterDim.z = 1;
        cl_attrs[1].id = cudaLaunchAttributeCooperative;
        cl_attrs[1].val.cooperative = 1;
        q.attrs = cl_attrs; q.numAttrs = 2;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, (void*)k_panel_reg<4, 2>, &q) == cudaSuccess) cl_max_ctas = ncl * PANEL_CL;
        (void)cudaGetLastError();
        if (cl_max_ctas > 35 * PANEL_CL) cl_max_ctas = 35 * PANEL_CL;  // gather width of the MODE 2 kernel
      }
    }
    int per = 256, mode = 0;
    panel_plan(rows, gmax, &per, &mode);
    int Gr = (rows + per - 1) / per, rpcr = (rows + Gr - 1) / Gr;
    // columns per exchange: 4 (default) or 2 = the grouped kernel k_panel_grp, 1 = one exchange per column (k_panel_reg)
    const char* e_s = getenv("QRDM_PANEL_S");
    int grp_s = e_s ? atoi(e_s) : 4;
    if (per == 256) grp_s = 1;  // 256-row slabs + the group's private columns and accumulators do not fit in 128 registers
    // No cluster launch available for this grid (more than 132 CTAs, QRDM_PANEL_CL=0, or a failed cluster launch earlier):
    // the grouped kernel runs its plain all-gather exchange (MODE 1) on any grid — more L2 traffic per exchange, but still
    // one exchange per 4 columns.  This is also the variant Nsight Compute can profile: under ncu the cooperative +
    // cluster launch of 31 clusters does not become co-resident and the run ends in the spin-limit trap.
    const int gmode = (mode == 0 && (grp_s == 2 || grp_s == 4)) ? 1 : mode;
    if (grp_s == 2 || grp_s == 4) {
      unsigned epoch_g = panel_grp_epoch(p, stream);
      qrdm_prob prob_g = *p;
      void* args_g[] = {(void*)&prob_g, (void*)&rpcr, (void*)&epoch_g};
#define PANEL_GFN(MODE, S) (per == 32 ? (void*)k_panel_grp<1, MODE, S> : per == 64 ? (void*)k_panel_grp<2, MODE, S> \
                            : per == 128 ? (void*)k_panel_grp<4, MODE, S> : (void*)k_panel_grp<8, MODE, S>)
      bool plain = gmode != 2;
      if (gmode == 2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((Gr + PANEL_CL - 1) / PANEL_CL * PANEL_CL);  // padded with CTAs that own no rows
        cfg.blockDim = dim3(PANEL_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
        cfg.attrs = cl_attrs; cfg.numAttrs = 2;
        const cudaError_t e2 = cudaLaunchKernelExC(&cfg, grp_s == 2 ? PANEL_GFN(2, 2) : PANEL_GFN(2, 4), args_g);
        if (e2 == cudaSuccess) { ++g_qrdm_launches; return 0; }
        (void)cudaGetLastError();
        cl_max_ctas = 0;  // no cluster launches from now on
        plain = true;
      }
      if (plain) {
        cudaError_t eg = cudaLaunchCooperativeKernel(grp_s == 2 ? PANEL_GFN(1, 2) : PANEL_GFN(1, 4), dim3(Gr), dim3(PANEL_THREADS),
                                                     args_g, 0, (cudaStream_t)stream);
        ++g_qrdm_launches;
        return eg == cudaSuccess ? 0 : (int)eg;
      }
#undef PANEL_GFN
    }
    unsigned epoch_r = panel_next_epoch(g_epoch_reg, 0x400000, 0x7fffff, p, stream);
    qrdm_prob prob_r = *p;
    void* args_r[] = {(void*)&prob_r, (void*)&rpcr, (void*)&epoch_r};
#define PANEL_FN(MODE) (per == 32 ? (void*)k_panel_reg<1, MODE> : per == 64 ? (void*)k_panel_reg<2, MODE> \
                        : per == 128 ? (void*)k_panel_reg<4, MODE> : (void*)k_panel_reg<8, MODE>)
    if (mode == 2) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((Gr + PANEL_CL - 1) / PANEL_CL * PANEL_CL);  // padded with CTAs that own no rows
      cfg.blockDim = dim3(PANEL_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
      cfg.attrs = cl_attrs; cfg.numAttrs = 2;
      const cudaError_t e2 = cudaLaunchKernelExC(&cfg, PANEL_FN(2), args_r);
      if (e2 == cudaSuccess) { ++g_qrdm_launches; return 0; }
      if (getenv("QRDM_PANEL_VERBOSE"))
        fprintf(stderr, "qrdm_b200: cluster panel launch failed (%s, grid %d); two-hop kernel from now on\n",
                cudaGetErrorString(e2), (int)cfg.gridDim.x);
      (void)cudaGetLastError();
      cl_max_ctas = 0;
      mode = 0;
    }
    void* fn = mode == 1 ? PANEL_FN(1) : PANEL_FN(0);
#undef PANEL_FN
    cudaError_t er = cudaLaunchCooperativeKernel(fn, dim3(Gr), dim3(PANEL_THREADS), args_r, 0, (cudaStream_t)stream);
    ++g_qrdm_launches;
    return er == cudaSuccess ? 0 : (int)er;
  }
  if (!getenv("QRDM_PANEL_NOBLOCK")) {
    // Tall panel (the slab does not fit in the register file of the whole chip): blocked Householder.
    // 8-column sub-panels are factored by the LL kernel above on 8 columns only, and each sub-panel's
    // block reflector is applied to the rest of the panel with the trailing-update kernels (K6) — the
    // unblocked sweep streamed all remaining panel columns through HBM for every single column
    // (371 of 439 ms on 2,000,000 x 512).  The DM early stop is unchanged: the stop test runs inside
    // the sub-panel kernel with the threshold carried in ctrl, and the reflectors produced before a
    // stop are still applied to the rest of the panel, as the unblocked reference does.
    int kmax_h = p->nb;
    if (kmax_h > p->n - j_host) kmax_h = p->n - j_host;
    if (kmax_h > p->m_glob - j_host) kmax_h = p->m_glob - j_host;
    const PeerCtx nopeer{};
    for (int sb = 0; sb < kmax_h; sb += QRDM_TALL_B) {
      const int rows_s = rows - sb;
      if (rows_s <= 0) break;
      int Gs = (rows_s + 1023) / 1024;  // >= 2 rows per thread; every SM streams its share of the rows
      if (Gs > gmax) Gs = gmax;
      if (Gs < 1) Gs = 1;
      int rpcs = (rows_s + Gs - 1) / Gs;
      unsigned epoch_t = qrdm_panel_tall_epoch(p, stream);
      qrdm_prob prob_s = *p;
      prob_s.sub = sb + 1;
      void* args_s[] = {(void*)&prob_s, (void*)&rpcs, (void*)&epoch_t, (void*)&nopeer};
      const size_t slab = tall_slab_bytes(rpcs);
      const bool grp = tall_grouped();
      void* fn_sm = grp ? (void*)k_panel_tall_grp<false, true> : (void*)k_panel_tall<false, true>;
      void* fn_gl = grp ? (void*)k_panel_tall_grp<false, false> : (void*)k_panel_tall<false, false>;
      cudaError_t es = slab ? cudaLaunchCooperativeKernel(fn_sm, dim3(Gs), dim3(PANEL_THREADS), args_s, slab, (cudaStream_t)stream)
                            : cudaLaunchCooperativeKernel(fn_gl, dim3(Gs), dim3(PANEL_THREADS), args_s, 0, (cudaStream_t)stream);
      ++g_qrdm_launches;
      if (es != cudaSuccess) return (int)es;
      if (sb + QRDM_TALL_B < kmax_h) {  // apply the sub-panel's reflectors to the rest of the panel
        const int rc = qrdm_k_skinny_update(&prob_s, rows_s, stream);
        if (rc) return rc;
      }
    }
    return 0;
  }
  int G = (rows + rows_per_cta - 1) / rows_per_cta;
  const int gfit = (rows + (smem_cap / 512) - 1) / (smem_cap / 512);  // CTAs needed for smem residency
  if (G < gfit) G = gfit;
  if (G > gmax) G = gmax;
  if (G < 1) G = 1;
  int rpc = (rows + G - 1) / G;
  unsigned epoch = panel_next_epoch(g_epoch_plain, 0x200000, 0x400000, p, stream);
  qrdm_prob prob = *p;
  void* args[] = {(void*)&prob, (void*)&rpc, (void*)&epoch};
  const size_t smem = (size_t)rpc * 64 * sizeof(double);
  cudaError_t e;
  if (smem <= (size_t)smem_cap)
    e = cudaLaunchCooperativeKernel((void*)k_panel<true>, dim3(G), dim3(PANEL_THREADS), args, smem, (cudaStream_t)stream);
  else
    e = cudaLaunchCooperativeKernel((void*)k_panel<false>, dim3(G), dim3(PANEL_THREADS), args, 0, (cudaStream_t)stream);
  ++g_qrdm_launches;
  if (e != cudaSuccess) return (int)e;
  return 0;
}
