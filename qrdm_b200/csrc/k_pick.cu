// k_pick.cu — K3c (greedy delta-orthogonality pick) + the column-permutation plan, and K3d (the
// batched full-height column moves).
//
// Replaces: the greedy loop of DM_perm (reference src/dgeqrdm_work.c:382-403) and
// permute_marked (src/dgeqrdm_work.c:149-262, nz = 0).  The plan reproduces the reference's
// exchange sequence exactly — including its quirks: selected columns already inside the leading
// fjb slots stay put, the result is NOT in norm order, "pointless" exchanges of two selected
// columns still happen, and vn2 is never exchanged — because jpvt parity depends on all of them.
// The exchanges are then composed into disjoint cycles so that K3d moves every affected column
// exactly once (bytes = 16*m per moved column) with (row-chunks x cycles) parallelism.
#include "common.cuh"

struct PickShared {
  double cosm[64 * 65];
  double inv[64];
  double tmpn[QRDM_MAXPOS];
  int tmpj[QRDM_MAXPOS];
  int selpos[64];
  int mk[64];
  int ex_p[QRDM_MAXEX], ex_q[QRDM_MAXEX];
  int pos[QRDM_MAXPOS], cur[QRDM_MAXPOS];
  int visited[QRDM_MAXPOS];
  int fjb, nex, npos;
};

__device__ __forceinline__ bool wl_marked(const int* mk, int fjb, int x, int lane) {
  bool f = false;
  for (int s = lane; s < fjb; s += 32) f |= (mk[s] == x);
  return __any_sync(0xffffffffu, f);
}
__device__ __forceinline__ int wl_find(const int* pos, int npos, int x, int lane) {
  for (int base = 0; base < npos; base += 32) {
    const int i = base + lane;
    const unsigned b = __ballot_sync(0xffffffffu, i < npos && pos[i] == x);
    if (b) return base + __ffs(b) - 1;
  }
  return -1;
}

__global__ void __launch_bounds__(256) k_pick(qrdm_prob P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PickShared& S = *reinterpret_cast<PickShared*>(smem_raw);
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = ctrl->j, kmax = ctrl->kmax, nc = ctrl->nc, cols = P.n - j;
  if (kmax == 0) return;

  if (nc > 1) {
    if (tid < nc) S.inv[tid] = 1.0 / ctrl->candnrm[tid];  // cc = 1/norm, src/dgeqrdm_work.c:366
    __syncthreads();
    for (int e = tid; e < nc * nc; e += blockDim.x) {
      const int s = e / nc, t = e - s * nc;
      S.cosm[s * 65 + t] = P.gram[s * 64 + t] * S.inv[s] * S.inv[t];
    }
  }
  __syncthreads();

  if (wid == 0) {
    // ---- greedy pick (all lanes in lock-step, scalars are warp-uniform) ----
    int fjb = 1;
    if (lane == 0) S.selpos[0] = 0;
    __syncwarp();
    for (int t = 1; t < nc; ++t) {
      double mx = 0.0;
      for (int s = lane; s < fjb; s += 32) mx = fmax(mx, fabs(S.cosm[S.selpos[s] * 65 + t]));
      mx = warp_max(mx);
      if (mx < P.delta && fjb < kmax) {
        if (lane == 0) S.selpos[fjb] = t;
        ++fjb;
      }
      __syncwarp();
    }
    for (int s = lane; s < fjb; s += 32) {
      const int c = ctrl->cand[S.selpos[s]];
      S.mk[s] = c;
      ctrl->sel[s] = c;
    }
    __syncwarp();

    // ---- exchange plan = permute_marked ----
    int jb = 0, nex = 0;
    const int jt = cols - 1;
    bool overflow = false;
    auto exchange = [&](int p, int q) {
      const bool mp = wl_marked(S.mk, fjb, p, lane), mq = wl_marked(S.mk, fjb, q, lane);
      if (mp != mq) {
        const int from = mp ? p : q, to = mp ? q : p;
        for (int s = lane; s < fjb; s += 32)
          if (S.mk[s] == from) S.mk[s] = to;
      }
      if (nex < QRDM_MAXEX) {
        if (lane == 0) { S.ex_p[nex] = p; S.ex_q[nex] = q; }
        ++nex;
      } else {
        overflow = true;
      }
      __syncwarp();
    };
    for (int s = 0; s < fjb; ++s) {
      const int jc = ctrl->sel[s];
      while (jb < jt && wl_marked(S.mk, fjb, jt, lane)) {
        exchange(jt, jb);
        while (jb < cols && wl_marked(S.mk, fjb, jb, lane)) ++jb;
      }
      if (wl_marked(S.mk, fjb, jc, lane)) {
        while (jb < cols && wl_marked(S.mk, fjb, jb, lane)) ++jb;
        if (jc <= jb || jc < fjb) continue;
        if (jb < cols && !wl_marked(S.mk, fjb, jb, lane)) {
          exchange(jc, jb);
          ++jb;
        }
      }
    }

    // ---- compose exchanges: cur[i] = original position whose column ends up at pos[i] ----
    int npos = 0;
    auto slot_of = [&](int x) {
      int i = wl_find(S.pos, npos, x, lane);
      if (i < 0) {
        i = npos;
        if (lane == 0) { S.pos[i] = x; S.cur[i] = x; S.visited[i] = 0; }
        ++npos;
        __syncwarp();
      }
      return i;
    };
    for (int e = 0; e < nex; ++e) {
      const int ip = slot_of(S.ex_p[e]), iq = slot_of(S.ex_q[e]);
      if (lane == 0) { const int t = S.cur[ip]; S.cur[ip] = S.cur[iq]; S.cur[iq] = t; }
      __syncwarp();
    }
    // ---- cycles: new[p_k] = old[p_{k+1}], new[p_last] = old[p_0] ----
    int ncyc = 0, total = 0;
    for (int i = 0; i < npos; ++i) {
      if (S.visited[i] || S.cur[i] == S.pos[i]) continue;
      if (lane == 0) ctrl->cyc_start[ncyc] = total;
      const int start = S.pos[i];
      int idx = i;
      while (true) {
        if (lane == 0) { S.visited[idx] = 1; ctrl->cyc_pos[total] = S.pos[idx]; }
        ++total;
        __syncwarp();
        const int nxt = S.cur[idx];
        if (nxt == start) break;
        idx = wl_find(S.pos, npos, nxt, lane);
      }
      ++ncyc;
    }
    if (lane == 0) {
      ctrl->cyc_start[ncyc] = total;
      ctrl->ncyc = ncyc;
      ctrl->fjb = fjb;
      ctrl->panel_bar = 0u;
      if (overflow) ctrl->err = QRDM_ERR_INTERNAL;
      S.fjb = fjb; S.nex = nex; S.npos = npos;
    }
  }
  __syncthreads();
  // ---- apply the composed permutation to jpvt and vn1 (NOT vn2: reference quirk) ----
  const int npos = S.npos;
  for (int i = tid; i < npos; i += blockDim.x) {
    S.tmpj[i] = P.jpvt[j + S.cur[i]];
    S.tmpn[i] = P.vn1[j + S.cur[i]];
  }
  __syncthreads();
  for (int i = tid; i < npos; i += blockDim.x) {
    P.jpvt[j + S.pos[i]] = S.tmpj[i];
    P.vn1[j + S.pos[i]] = S.tmpn[i];
  }
}

// K3d: rotate the full-height columns along each cycle.  grid = (row chunks, cycles).
__global__ void __launch_bounds__(256) k_permute(qrdm_prob P) {
  const qrdm_ctrl* ctrl = P.ctrl;
  const int cyc = blockIdx.y;
  if (cyc >= ctrl->ncyc) return;
  const int j = ctrl->j;
  const int b = ctrl->cyc_start[cyc], e = ctrl->cyc_start[cyc + 1];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.m) return;
  double* a = P.a + (size_t)j * P.lda + r;
  const double first = a[(size_t)ctrl->cyc_pos[b] * P.lda];
  double nxt = a[(size_t)ctrl->cyc_pos[b + 1] * P.lda];
  for (int k = b; k < e - 1; ++k) {
    const double v = nxt;
    if (k + 2 < e) nxt = a[(size_t)ctrl->cyc_pos[k + 2] * P.lda];
    a[(size_t)ctrl->cyc_pos[k] * P.lda] = v;
  }
  a[(size_t)ctrl->cyc_pos[e - 1] * P.lda] = first;
}

extern "C" int qrdm_k_pick(const qrdm_prob* p, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_pick, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PickShared));
    attr_set = true;
  }
  k_pick<<<1, 256, sizeof(PickShared), (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int qrdm_k_permute(const qrdm_prob* p, void* stream) {
  // at most 2*nb exchanges -> at most 2*nb cycles of length >= 2
  const int maxcyc = 2 * (p->nb < QRDM_KMAX ? p->nb : QRDM_KMAX);
  dim3 grid((p->m + 255) / 256, maxcyc);
  k_permute<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
