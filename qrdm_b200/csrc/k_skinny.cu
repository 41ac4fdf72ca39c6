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
#include "ll.cuh"

#define SK_RC 2048           // rows per CTA chunk
#define SK_THREADS 256       // thread <-> rows tid, tid+256, ... of the chunk (coalesced per column)
#define SK_RPT (SK_RC / SK_THREADS)

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

// W_ext partial of one 2048-row chunk.  Each thread keeps the 8 reflector entries of its row in
// registers and, per group of 8 columns, issues 8 independent loads and 64 FMAs into an 8x8
// register tile; the tiles are reduced block-wide once per group (fixed order).
__global__ void __launch_bounds__(SK_THREADS, 1) k_sub_w(qrdm_prob P) {
  __shared__ double red[SK_THREADS / 32][64];
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int row0 = blockIdx.x * SK_RC;
  if (row0 >= g.rows) return;
  const int nrc = min(SK_RC, g.rows - row0);
  const int ngroups = 1 + (g.ncp + 7) / 8;
  double* out = P.gram_part + (size_t)blockIdx.x * 512;  // [group][q][c]: 8 groups x 64
  const double* vbase = P.vc + (size_t)g.voff * P.ldv + g.jr + row0;
  const double* abase = P.a + (size_t)g.c0 * P.lda + g.jr + row0;
  for (int gi = 0; gi < ngroups; ++gi) {
    double acc[8][8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = 0.0;
    for (int it = 0; it < SK_RPT; ++it) {
      const int r = tid + it * SK_THREADS;
      if (r >= nrc) break;
      double v[8], x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = q < g.k ? vbase[(size_t)q * P.ldv + r] : 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = (gi - 1) * 8 + c;
        x[c] = gi == 0 ? v[c] : (col < g.ncp ? abase[(size_t)col * P.lda + r] : 0.0);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[q][c] = fma(v[q], x[c], acc[q][c]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double t = warp_sum(acc[q][c]);
        if (lane == 0) red[wid][q * 8 + c] = t;
      }
    __syncthreads();
    if (tid < 64) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < SK_THREADS / 32; ++w) t += red[w][tid];
      out[gi * 64 + tid] = t;
    }
    __syncthreads();
  }
}

// MG = true (row-sharded, peer memory open): the fold of this rank's chunk partials, the sum over the ranks and the
// T' / W2 step in ONE single-CTA kernel — thread e stores its folded value as an LL packet into every peer's receive
// buffer (k_peer.cu, generic region) and adds the nranks packets of its own buffer in rank order, so all ranks hold
// bit-identical W; replaces k_sub_wred + all-reduce kernel + k_sub_w2 (three launches per sub-panel).
template <bool MG>
__global__ void __launch_bounds__(1024) k_sub_w2(qrdm_prob P, int nchunks_max, PeerCtx pc, unsigned ptag, int pparity) {
  __shared__ double W[8 * 64];   // [group][q][c]
  __shared__ double T[64];       // T'[q][p]
  const SkGeom g = sk_geom(P);
  const int tid = threadIdx.x;
  if (g.k <= 0 || g.ncp <= 0) {
    // Nothing to update (replicated decision: every rank gets here alike).  The host has already drawn a sequence
    // number for this exchange, and the two-parity slot scheme of k_peer.cu is only safe if consecutive PERFORMED
    // exchanges alternate parity — so the exchange is still performed, on one dummy element.  (Skipping it let a fast
    // rank overwrite, two sequence numbers later, a slot a slow rank had not read yet: the waiting rank then spun into
    // its trap — seen with two ranks time-slicing one GPU on Kahan inputs, where every panel has one column.)
    if (MG && tid == 0) {
      const int me = pc.rank, N = pc.nranks;
      for (int r = 0; r < N; ++r)
        if (r != me) ll_store(peer_gen_slot(pc.recv[r], pparity, me, 0), 0.0, ptag);
      for (int r = 0; r < N; ++r)
        if (r != me) (void)ll_load(peer_gen_slot(pc.recv[me], pparity, r, 0), ptag);
    }
    return;
  }
  // nchunks_max < 0: exactly -nchunks_max partials were written (TMA + DMMA producer, CTAs without rows write zeros);
  // > 0: the FMA producer k_sub_w, whose CTAs beyond the last row chunk do not write
  const int nchunks = nchunks_max < 0 ? -nchunks_max
                      : MG ? min(nchunks_max, g.rows > 0 ? (g.rows + SK_RC - 1) / SK_RC : 0)
                           : (P.w_reduced ? 1 : min(nchunks_max, (g.rows + SK_RC - 1) / SK_RC));  // w_reduced: chunk 0 = all-reduced sum
  const int ngroups = 1 + (g.ncp + 7) / 8;
  // fold of the partials: the CTA has 1024 threads, half h sums the partials b = h, h + 2, ... of entry e with four
  // loads in flight, then the two halves are added (fixed order: deterministic)
  __shared__ double Wh[2][512];
  {
    const int h = tid >> 9, e = tid & 511;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (e < ngroups * 64) {
      int b = h;
      for (; b + 6 < nchunks; b += 8) {
        s0 += P.gram_part[(size_t)b * 512 + e];
        s1 += P.gram_part[(size_t)(b + 2) * 512 + e];
        s2 += P.gram_part[(size_t)(b + 4) * 512 + e];
        s3 += P.gram_part[(size_t)(b + 6) * 512 + e];
      }
      for (; b < nchunks; b += 2) s0 += P.gram_part[(size_t)b * 512 + e];
    }
    Wh[h][e] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  if (tid < ngroups * 64) {
    const double mine = Wh[0][tid] + Wh[1][tid];
    if (MG) {
      const int me = pc.rank, N = pc.nranks;
      for (int r = 0; r < N; ++r)
        if (r != me) ll_store(peer_gen_slot(pc.recv[r], pparity, me, (size_t)tid), mine, ptag);
      double tot = 0.0;
      for (int r = 0; r < N; ++r) tot += (r == me) ? mine : ll_load(peer_gen_slot(pc.recv[me], pparity, r, (size_t)tid), ptag);
      W[tid] = tot;
    } else {
      W[tid] = mine;
    }
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

// C_p += V W2: per group of 8 columns the 8x8 block of W2 sits in registers; every row needs 8
// independent loads, 64 FMAs and 8 stores.
__global__ void __launch_bounds__(SK_THREADS, 2) k_sub_apply(qrdm_prob P) {
  __shared__ double W2s[8 * 64];
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * SK_RC;
  if (row0 >= g.rows) return;
  const int nrc = min(SK_RC, g.rows - row0);
  for (int e = tid; e < 8 * 64; e += SK_THREADS) {
    const int q = e >> 6, col = e & 63;
    W2s[e] = col < g.ncp ? P.w2[(size_t)q * P.ldw + col] : 0.0;
  }
  __syncthreads();
  const double* vbase = P.vc + (size_t)g.voff * P.ldv + g.jr + row0;
  double* abase = P.a + (size_t)g.c0 * P.lda + g.jr + row0;
  const int ngroups = (g.ncp + 7) / 8;
  for (int it = 0; it < SK_RPT; ++it) {
    const int r = tid + it * SK_THREADS;
    if (r >= nrc) break;
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = q < g.k ? vbase[(size_t)q * P.ldv + r] : 0.0;
    for (int gi = 0; gi < ngroups; ++gi) {
      double x[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = gi * 8 + c;
        x[c] = col < g.ncp ? abase[(size_t)col * P.lda + r] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int q = 0; q < 8; ++q) x[c] = fma(v[q], W2s[q * 64 + gi * 8 + c], x[c]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = gi * 8 + c;
        if (col < g.ncp) abase[(size_t)col * P.lda + r] = x[c];
      }
    }
  }
}

// row-sharded build: fold the chunk partials into chunk 0 (fixed order) so that 512 doubles can be
// all-reduced; a rank without rows contributes zeros
__global__ void __launch_bounds__(512) k_sub_wred(qrdm_prob P, int nchunks_max) {
  const SkGeom g = sk_geom(P);
  if (g.k <= 0 || g.ncp <= 0) return;
  const int nchunks = min(nchunks_max, g.rows > 0 ? (g.rows + SK_RC - 1) / SK_RC : 0);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int b = 0;
  for (; b + 3 < nchunks; b += 4) {
    s0 += P.gram_part[(size_t)b * 512 + threadIdx.x];
    s1 += P.gram_part[(size_t)(b + 1) * 512 + threadIdx.x];
    s2 += P.gram_part[(size_t)(b + 2) * 512 + threadIdx.x];
    s3 += P.gram_part[(size_t)(b + 3) * 512 + threadIdx.x];
  }
  for (; b < nchunks; ++b) s0 += P.gram_part[(size_t)b * 512 + threadIdx.x];
  P.gram_part[threadIdx.x] = (s0 + s1) + (s2 + s3);
}

// rows_hint: host-side upper bound of the rows of the sub-panel
extern "C" int qrdm_k_subw_tma(const qrdm_prob* p, int rows_hint, int* nparts, void* stream);  // k_gram.cu

extern "C" int qrdm_k_skinny_update(const qrdm_prob* p, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  int nparts = 0, fold = nch;
  const int rc = qrdm_k_subw_tma(p, rows_hint, &nparts, stream);
  if (rc > 0) return rc;
  if (rc == 0) fold = -nparts;
  else { k_sub_w<<<nch, SK_THREADS, 0, s>>>(*p); QRDM_LAUNCH_CHECK(); }
  k_sub_w2<false><<<1, 1024, 0, s>>>(*p, fold, PeerCtx{}, 0u, 0);
  QRDM_LAUNCH_CHECK();
  k_sub_apply<<<nch, SK_THREADS, 0, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}

// row-sharded variant, split at the all-reduce: part 1 = partials folded into gram_part[0..512)
extern "C" int qrdm_k_skinny_part(const qrdm_prob* p, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  k_sub_w<<<nch, SK_THREADS, 0, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  k_sub_wred<<<1, 512, 0, s>>>(*p, nch);
  QRDM_LAUNCH_CHECK();
  return 0;
}
extern "C" int qrdm_k_skinny_finish(const qrdm_prob* p, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  k_sub_w2<false><<<1, 1024, 0, s>>>(*p, 1, PeerCtx{}, 0u, 0);
  QRDM_LAUNCH_CHECK();
  k_sub_apply<<<nch, SK_THREADS, 0, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
// row-sharded with peer memory: partial products, then fold + cross-GPU sum + W2 in one kernel, then the update
extern "C" int qrdm_k_skinny_update_mg(const qrdm_prob* p, int rows_hint, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const PeerCtx* pc = qrdm_peer_ctx();
  if (!pc) return (int)cudaErrorInvalidValue;
  int nch = (rows_hint + SK_RC - 1) / SK_RC;
  if (nch < 1) nch = 1;
  int nparts = 0, fold = nch;
  const int rc = qrdm_k_subw_tma(p, rows_hint, &nparts, stream);
  if (rc > 0) return rc;
  if (rc == 0) fold = -nparts;
  else { k_sub_w<<<nch, SK_THREADS, 0, s>>>(*p); QRDM_LAUNCH_CHECK(); }
  unsigned tag = 0;
  int parity = 0;
  qrdm_peer_next_gen(&tag, &parity);
  k_sub_w2<true><<<1, 1024, 0, s>>>(*p, fold, *pc, tag, parity);
  QRDM_LAUNCH_CHECK();
  k_sub_apply<<<nch, SK_THREADS, 0, s>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
