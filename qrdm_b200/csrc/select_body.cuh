// select_body.cuh — body of K3a (k_select), shared by the one-kernel-per-stage path (k_select.cu)
// and the one-CTA-per-matrix batched kernel (k_small.cu).  See k_select.cu for what it replaces.
#pragma once
#include "common.cuh"

struct SelShared {
  unsigned long long red[32];
  unsigned long long key[QRDM_SELCAP];
  int idx[QRDM_SELCAP];
  unsigned hist[256];
  unsigned wsum[32];
  int count;
  int j, cols, kmax;
  // radix state
  unsigned long long prefix;  // value of the key bits above `shift` that survivors must match
  int shift, need, greater, bucket, done;
};

// Partial norms as radix keys.  A NaN norm (NaN somewhere in the column) never wins a comparison
// in the reference (cmpStruct, `work[j] > maxnrm`), so it must not look like the largest key:
// map it to the smallest one; the NaN is then caught by the block-reflector screen (-13).
__device__ __forceinline__ unsigned long long ldkey(const unsigned long long* keys, int c) {
  const unsigned long long v = keys[c];
  return v > 0x7ff0000000000000ull ? 0ull : v;
}

__device__ __forceinline__ bool before(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return ka > kb || (ka == kb && ia < ib);
}

template <int NT>
__device__ __forceinline__ void qrdm_select_body(const qrdm_prob& P, SelShared& S) {
  qrdm_ctrl* ctrl = P.ctrl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (tid == 0) {
    const int k = ctrl->fjb_cmp;
    const int j = ctrl->j + k;
    const int rows = P.m_glob - j, cols = P.n - j;
    int kmax = min(P.nb, min(rows, cols));
    if (kmax < 0) kmax = 0;
    S.j = j; S.cols = cols; S.kmax = kmax;
    ctrl->j = j;
    ctrl->last_k = k;
    ctrl->fjb_cmp = 0;
    ctrl->fjb = 0;
    ctrl->nc = 0;
    ctrl->ncyc = 0;
    ctrl->nflag = 0;
    ctrl->forced = 0;
    ctrl->kmax = kmax;
    if (k > 0) ctrl->it += 1;
    S.count = 0;
  }
  __syncthreads();
  const int j = S.j, cols = S.cols, kmax = S.kmax;
  if (cols <= 0) {
    if (tid == 0) ctrl->maxnrm = 0.0;
    return;
  }
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(P.vn1 + j);

  // pass 0: min / max key
  unsigned long long kmx = 0ull, kmn = ~0ull;
  for (int c = tid; c < cols; c += NT) {
    const unsigned long long v = ldkey(keys, c);
    kmx = v > kmx ? v : kmx;
    kmn = v < kmn ? v : kmn;
  }
  kmx = warp_max_u64(kmx);
  kmn = warp_min_u64(kmn);
  if (lane == 0) S.red[wid] = kmx;
  __syncthreads();
  if (wid == 0) {
    unsigned long long v = lane < NT / 32 ? S.red[lane] : 0ull;
    v = warp_max_u64(v);
    if (lane == 0) S.red[0] = v;
  }
  __syncthreads();
  kmx = S.red[0];
  __syncthreads();
  if (lane == 0) S.red[wid] = kmn;
  __syncthreads();
  if (wid == 0) {
    unsigned long long v = lane < NT / 32 ? S.red[lane] : ~0ull;
    v = warp_min_u64(v);
    if (lane == 0) S.red[0] = v;
  }
  __syncthreads();
  kmn = S.red[0];
  if (tid == 0) ctrl->maxnrm = __longlong_as_double((long long)kmx);
  if (kmax == 0) return;
  if (j < P.nfxd) {
    // Fixed columns (jpvt[j] != 0 on entry, moved up front by the driver as src/dgeqrdm_work.c:592-607 does): the
    // reference factors them with LAPACKE_dgeqrf and applies Q' with LAPACKE_dormqr (:612-635).  Here they go through
    // the same panel + trailing-update kernels in blocks of <= nb columns, taken as they stand.
    const int kf = min(kmax, P.nfxd - j);
    for (int c = tid; c < kf; c += NT) { ctrl->cand[c] = c; ctrl->candnrm[c] = P.vn1[j + c]; }
    __syncthreads();
    if (tid == 0) { ctrl->nc = kf; ctrl->forced = 1; }
    return;
  }

  // ---- narrow down to <= SELCAP survivors containing the top kmax ----
  bool collect_all = cols <= QRDM_SELCAP;
  if (!collect_all) {
    if (tid == 0) {
      const unsigned long long x = kmn ^ kmx;
      S.need = kmax; S.greater = 0; S.done = 0; S.bucket = cols;
      if (x == 0ull) { S.shift = 0; S.prefix = kmx; S.done = 1; }  // all norms identical
      else { const int hb = 63 - __clzll((long long)x); S.shift = (hb / 8) * 8; S.prefix = (S.shift + 8 >= 64) ? 0ull : (kmx >> (S.shift + 8)); }
    }
    __syncthreads();
    while (!S.done) {
      const int shift = S.shift;
      const unsigned long long pre = S.prefix;
      if (tid < 256) S.hist[tid] = 0;
      __syncthreads();
      for (int c = tid; c < cols; c += NT) {
        const unsigned long long v = ldkey(keys, c);
        const bool match = (shift + 8 >= 64) ? true : ((v >> (shift + 8)) == pre);
        if (match) atomicAdd(&S.hist[(unsigned)(v >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (wid == 0) {
        // lane l owns digits 255-8l .. 248-8l (descending); find where the running count from the
        // top reaches `need`
        unsigned loc[8], tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { loc[q] = S.hist[255 - 8 * lane - q]; tot += loc[q]; }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        const unsigned excl = incl - tot;
        const unsigned need = (unsigned)S.need;
        const bool mine = excl < need && incl >= need;
        if (mine) {
          unsigned cum = excl;
          int q = 0;
          for (; q < 8; ++q) { if (cum + loc[q] >= need) break; cum += loc[q]; }
          const int digit = 255 - 8 * lane - q;
          S.greater += (int)cum;
          S.need = (int)(need - cum);
          S.bucket = (int)loc[q];
          S.prefix = (shift + 8 >= 64) ? (unsigned long long)digit : ((pre << 8) | (unsigned long long)digit);
          if (S.greater + (int)loc[q] <= QRDM_SELCAP || shift == 0) S.done = 1; else S.shift = shift - 8;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- collect survivors ----
  if (collect_all) {
    for (int c = tid; c < cols; c += NT) { S.key[c] = ldkey(keys, c); S.idx[c] = c; }
    if (tid == 0) S.count = cols;
  } else {
    const int shift = S.shift;
    const unsigned long long pre = S.prefix;  // survivors: (key >> shift) >= pre
    const bool ties_overflow = S.greater + S.bucket > QRDM_SELCAP;  // only possible when shift == 0
    for (int c = tid; c < cols; c += NT) {
      const unsigned long long v = ldkey(keys, c);
      const unsigned long long hi = v >> shift;
      if (hi > pre || (hi == pre && !ties_overflow)) {
        const int slot = atomicAdd(&S.count, 1);
        S.key[slot] = v; S.idx[slot] = c;
      }
    }
    __syncthreads();
    if (ties_overflow) {
      // more exact ties than fit: take the `need` lowest-index columns whose norm == pre
      // (ordered stream compaction, chunks of NT columns in index order)
      int taken = 0;
      const int need = S.need;
      for (int base = 0; base < cols && taken < need; base += NT) {
        const int c = base + tid;
        const bool f = c < cols && ldkey(keys, c) == pre;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) S.wsum[wid] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
        for (int w = 0; w < NT / 32; ++w) { const int v = (int)S.wsum[w]; if (w < wid) off += v; total += v; }
        const int pos = taken + off + __popc(bal & ((1u << lane) - 1u));
        if (f && pos < need) { const int slot = S.greater + pos; S.key[slot] = pre; S.idx[slot] = c; }
        taken += total;
        __syncthreads();
      }
      if (tid == 0) S.count = S.greater + need;
    }
  }
  __syncthreads();

  // ---- exact rank among survivors (value descending, index ascending) ----
  const int L = S.count;
  for (int e = tid; e < L; e += NT) {
    const unsigned long long ke = S.key[e];
    const int ie = S.idx[e];
    int rank = 0;
    for (int f = 0; f < L; ++f) rank += before(S.key[f], S.idx[f], ke, ie) ? 1 : 0;
    if (rank < kmax) {
      ctrl->cand[rank] = ie;
      ctrl->candnrm[rank] = __longlong_as_double((long long)ke);
    }
  }
  __syncthreads();
  __threadfence_block();
  if (tid == 0) {
    // candidates = leading run with norm > tau_ * max norm, at most kmax (src/dgeqrdm_work.c:348-351)
    const double thr = P.tau_ * ctrl->candnrm[0];
    int nc = 0;
    while (nc < kmax && ctrl->candnrm[nc] > thr) ++nc;
    ctrl->nc = nc;
  }
}

