// pick_body.cuh — body of K3c (greedy pick + permutation plan), shared by k_pick.cu and the
// one-CTA-per-matrix batched kernel (k_small.cu).  See k_pick.cu for what it replaces.
#pragma once
#include "common.cuh"

struct PickShared {
  double cosm[64 * 65];
  double inv[64];
  int cand[64];
  int selpos[64];
  int sel[64];
  int hi_pos[64];  // current position of selected column s when it sits at a position >= 64, else -1
  int cyc_start[QRDM_MAXEX + 2];
  int cyc_pos[2 * QRDM_MAXEX + 4];
  int fjb, ncyc;
};

// The exchange sequence of permute_marked has a simple structure (derivation in DESIGN.md):
//  * "tail phase": if the LAST active column is selected it is exchanged with slot jb (0, then
//    the first unselected slot) — at most two exchanges, one 2- or 3-cycle; this is the only
//    place where two selected columns can be swapped "pointlessly" (e.g. identity -> jpvt 16,1,2..);
//  * every selected column sitting at a position >= fjb is exchanged, in acceptance order, with
//    the successive unselected slots among the leading positions: disjoint 2-cycles; selected
//    columns already inside the leading fjb slots stay where they are.
// One thread replays the reference's loop verbatim on a 64-bit mask of the leading slots (every
// exchange has jb < 64), so the plan costs a few hundred instructions instead of a warp-wide
// search per flag lookup.
__device__ __forceinline__ void qrdm_pick_body(const qrdm_prob& P, PickShared& S) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, kmax = ctrl->kmax, nc = ctrl->nc, cols = P.n - j;
  if (kmax == 0) return;
  if (ctrl->forced) {  // fixed columns: all nc leading columns, in place (no greedy test, no exchange)
    for (int s = tid; s < nc; s += blockDim.x) ctrl->sel[s] = s;
    if (tid == 0) {
      S.fjb = nc; S.ncyc = 0; S.cyc_start[0] = 0;
      ctrl->fjb = nc; ctrl->ncyc = 0; ctrl->cyc_start[0] = 0; ctrl->panel_bar = 0u;
    }
    __syncthreads();
    return;
  }

  if (tid < 64) {
    S.cand[tid] = tid < kmax ? ctrl->cand[tid] : -1;
    if (tid < nc) S.inv[tid] = 1.0 / ctrl->candnrm[tid];  // cc = 1/norm, src/dgeqrdm_work.c:366
  }
  __syncthreads();
  if (nc > 1) {
    for (int e = tid; e < nc * nc; e += blockDim.x) {
      const int s = e / nc, t = e - s * nc;
      S.cosm[s * 65 + t] = P.gram[s * 64 + t] * S.inv[s] * S.inv[t];
    }
  }
  __syncthreads();

  if (wid == 0) {
    // ---- greedy pick (lock-step warp, scalars warp-uniform), src/dgeqrdm_work.c:382-403 ----
    // lane l tracks, for candidates l and l+32, the largest |cos| against the columns accepted so
    // far; candidate t is accepted iff that running maximum is < delta (one shuffle per candidate,
    // one pair of smem reads per acceptance)
    int fjb = 1;
    if (lane == 0) S.selpos[0] = 0;
    double m0 = (nc > 1 && lane < nc) ? fmax(0.0, fabs(S.cosm[lane])) : 0.0;  // vs candidate 0 (a NaN cosine is ignored,
    double m1 = (nc > 1 && lane + 32 < nc) ? fmax(0.0, fabs(S.cosm[lane + 32])) : 0.0;  // like the reference's `maxval < fabs()`)
    for (int t = 1; t < nc; ++t) {
      const double mine = (t & 32) ? m1 : m0;
      const double mx = __shfl_sync(0xffffffffu, mine, t & 31);
      if (mx < P.delta && fjb < kmax) {
        if (lane == 0) S.selpos[fjb] = t;
        ++fjb;
        if (lane < nc) m0 = fmax(m0, fabs(S.cosm[t * 65 + lane]));
        if (lane + 32 < nc) m1 = fmax(m1, fabs(S.cosm[t * 65 + lane + 32]));
      }
    }
    __syncwarp();
    for (int s = lane; s < fjb; s += 32) {
      const int c = S.cand[S.selpos[s]];
      S.sel[s] = c;
      S.hi_pos[s] = c >= 64 ? c : -1;
      ctrl->sel[s] = c;
    }
    __syncwarp();

    if (lane == 0) {
      // ---- exchange plan = permute_marked (src/dgeqrdm_work.c:149-262), replayed serially ----
      unsigned long long low = 0ull;
      for (int s = 0; s < fjb; ++s)
        if (S.sel[s] < 64) low |= 1ull << S.sel[s];
      auto marked = [&](int x) -> bool {
        if (x < 64) return (low >> x) & 1ull;
        for (int s = 0; s < fjb; ++s)
          if (S.hi_pos[s] == x) return true;
        return false;
      };
      auto skip = [&](int jb) -> int {  // while (jb < cols && marked[jb]) ++jb
        while (jb < cols && marked(jb)) {
          if (jb < 64) {
            const unsigned long long freebits = ~low >> jb;
            jb = freebits ? jb + __ffsll((long long)freebits) - 1 : 64;
          } else {
            ++jb;
          }
        }
        return jb;
      };
      auto unmark_hi = [&](int x) {
        for (int s = 0; s < fjb; ++s)
          if (S.hi_pos[s] == x) S.hi_pos[s] = -1;
      };
      int jb = 0, ncyc = 0, total = 0;
      const int jt = cols - 1;
      bool overflow = false;
      // tail phase (can only fire on the first pass of the reference's outer loop)
      int tail[3], ntail = 0;
      while (jb < jt && marked(jt) && ntail < 3) {
        if (jb >= 64) { overflow = true; break; }  // cannot happen: see header comment
        const bool mb = (low >> jb) & 1ull;
        if (ntail == 0) { tail[0] = jb; tail[1] = jt; ntail = 2; } else { tail[ntail++] = jb; }
        if (!mb) {  // flags differ: the selected column moves from jt to jb
          low |= 1ull << jb;
          if (jt < 64) low &= ~(1ull << jt); else unmark_hi(jt);
        }
        jb = skip(jb);
      }
      if (ntail >= 2) {  // exchanges (jt,a)[,(jt,b)] compose to the cycle a <- jt [<- b]
        S.cyc_start[ncyc++] = total;
        for (int q = 0; q < ntail; ++q) S.cyc_pos[total++] = tail[q];
      }
      for (int s = 0; s < fjb; ++s) {
        const int jc = S.sel[s];
        const bool m = jc < 64 ? ((low >> jc) & 1ull) : (S.hi_pos[s] == jc);
        if (!m) continue;
        jb = skip(jb);
        if (jc <= jb || jc < fjb) continue;
        if (jb < cols && !marked(jb)) {
          if (jb >= 64 || ncyc >= QRDM_MAXEX) { overflow = true; break; }
          low |= 1ull << jb;
          if (jc < 64) low &= ~(1ull << jc); else S.hi_pos[s] = -1;
          S.cyc_start[ncyc++] = total;
          S.cyc_pos[total++] = jc;
          S.cyc_pos[total++] = jb;
          ++jb;
        }
      }
      S.cyc_start[ncyc] = total;
      S.ncyc = ncyc;
      S.fjb = fjb;
      ctrl->ncyc = ncyc;
      ctrl->fjb = fjb;
      ctrl->panel_bar = 0u;
      if (overflow) ctrl->err = QRDM_ERR_INTERNAL;
    }
  }
  __syncthreads();
  // ---- publish the cycles; rotate jpvt and vn1 along them (NOT vn2: reference quirk) ----
  // cycle (p_0 .. p_{L-1}): new[p_k] = old[p_{k+1}], new[p_{L-1}] = old[p_0]
  const int ncyc = S.ncyc;
  for (int c = tid; c <= ncyc; c += blockDim.x) ctrl->cyc_start[c] = S.cyc_start[c];
  for (int e = tid; e < S.cyc_start[ncyc]; e += blockDim.x) ctrl->cyc_pos[e] = S.cyc_pos[e];
  for (int c = tid; c < ncyc; c += blockDim.x) {
    const int b = S.cyc_start[c], e = S.cyc_start[c + 1];
    const int p0 = j + S.cyc_pos[b];
    const int j0 = P.jpvt[p0];
    const double n0 = P.vn1[p0];
    for (int q = b; q < e - 1; ++q) {
      const int dst = j + S.cyc_pos[q], src = j + S.cyc_pos[q + 1];
      P.jpvt[dst] = P.jpvt[src];
      P.vn1[dst] = P.vn1[src];
    }
    const int last = j + S.cyc_pos[e - 1];
    P.jpvt[last] = j0;
    P.vn1[last] = n0;
  }
}

