// k_qp3.cu — blocked Householder QR with classical column pivoting on the GPU (LAPACK dgeqp3 / dlaqps), SURVEY 8f-4:
// the QRCP comparator the reference exports (src/dgeqp3.c:39-93, QRDM_wrapper.c:15-41 -> LAPACKE_dgeqp3) and the
// notebook times against dgeqrdm, here on the same device so that the comparison is GPU against GPU.
//
// The algorithm is dlaqps's, panel by panel (nb = 64): per column ONE pass over the not yet updated trailing matrix
// (the BLAS-2 half of QRCP that Deviation Maximisation exists to avoid), then one rank-kb update per panel on the DMMA
// kernel of the trailing update (k_rankk, pending-block geometry: the pivot rows are already up to date).
//   step k of the panel that starts at column j (rk = j + k):
//     k_qp3_pivot    pvt = first argmax of vn1[rk:]; jpvt / vn1 / vn2 / row k of F exchanged            (1 CTA)
//     k_qp3_swapupd  columns rk <-> pvt of A exchanged; the pivot column brought up to date:
//                    A[rk:, rk] -= A[rk:, j:rk] F[k, 0:k]'                                               (rows)
//     k_qp3_gemv     g_c = A[rk+1:, c]' x for every column c >= j (x = the pivot column below the diagonal): gives
//                    ||x||^2 (c = rk), the V'x terms of the panel's earlier reflectors and A' x of the trailing columns
//                    — dlarfg's norm and dlaqps's two dgemv in one pass                                  (columns x row splits)
//     k_qp3_finish   tau / beta / scale; F[c, k] = tau (A[rk, c] + scale g_c) + sum_q F[c, q] auxv_q,
//                    auxv_q = -tau (A[rk, j+q] + scale g_{j+q}); pivot row A[rk, c] -= A[rk, j:rk+1] F[c, 0:k+1]';
//                    norm downdate with dlaqps's cancellation test (a column that fails it is flagged: the panel ends
//                    after this step and its norm is recomputed — every later step kernel of the panel sees
//                    state->stopped and returns, the host reads the panel's length once per panel)     (columns)
//     k_qp3_scale    v = x * scale, beta on the diagonal                                                  (rows)
//   per panel: Vc (k_vc_build), W2 = -F' (k_qp3_w2), C += V W2 on rows / columns >= j + kb (k_rankk), flagged norms.
// F lives in P.wp as Ft[q][c - j] (row-major, ld = P.ldw), the g partials behind it.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

#define QP3_NB 64
#define QP3_RC 4096   // rows of x a GEMV CTA keeps in shared memory at a time
#define QP3_MAXSPLIT 12
struct Qp3State {
  int pvt, stopped, kb, stop_req;  // stop_req: set by k_qp3_finish, turned into `stopped` by k_qp3_scale (a CTA of the
                                   // same launch must not see the flag another CTA has just raised)
  double tau, beta, scale, pad;
};
__device__ __forceinline__ Qp3State* qp3_state(const qrdm_prob& P) { return reinterpret_cast<Qp3State*>(P.gram); }
__device__ __forceinline__ double* qp3_ft(const qrdm_prob& P) { return P.wp; }                             // [64][ldw]
__device__ __forceinline__ double* qp3_gpart(const qrdm_prob& P) { return P.wp + (size_t)64 * P.ldw; }     // [nsplit][ldw]

__global__ void __launch_bounds__(256) k_qp3_norms(qrdm_prob P) {  // vn1 = vn2 = column norms, jpvt = identity (1-based)
  __shared__ double sc[32];
  const int c = blockIdx.x;
  const double* col = P.a + (size_t)c * P.lda;
  double s = 0.0;
  for (int r = threadIdx.x; r < P.m; r += 256) s = fma(col[r], col[r], s);
  s = block_sum(s, sc);
  if (threadIdx.x == 0) {
    const double nrm = sqrt(s);
    P.vn1[c] = nrm; P.vn2[c] = nrm; P.jpvt[c] = c + 1; P.flag_list[c] = 0;
  }
}

__global__ void __launch_bounds__(1024) k_qp3_pivot(qrdm_prob P, int j, int k) {
  __shared__ double sv[32];
  __shared__ int si[32];
  Qp3State* st = qp3_state(P);
  if (k == 0 && threadIdx.x == 0) { st->stopped = 0; st->stop_req = 0; st->kb = 0; }
  __syncthreads();
  if (k > 0 && st->stopped) return;
  const int rk = j + k, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double best = -1.0;
  int bi = rk;
  for (int c = rk + tid; c < P.n; c += 1024) {  // ascending within a thread: the first maximum wins, as idamax
    const double v = P.vn1[c];
    if (v > best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { sv[wid] = best; si[wid] = bi; }
  __syncthreads();
  if (wid == 0) {
    best = sv[lane]; bi = si[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      st->pvt = bi;
      if (bi != rk) {
        const int t = P.jpvt[bi]; P.jpvt[bi] = P.jpvt[rk]; P.jpvt[rk] = t;
        P.vn1[bi] = P.vn1[rk]; P.vn2[bi] = P.vn2[rk];  // (entry rk is not looked at again)
        const int f = P.flag_list[bi]; P.flag_list[bi] = P.flag_list[rk]; P.flag_list[rk] = f;
      }
    }
    const int pv = __shfl_sync(0xffffffffu, bi, 0);
    if (pv != rk) {  // rows pv - j and k of F change places (columns 0 .. k-1)
      double* Ft = qp3_ft(P);
      for (int q = lane; q < k; q += 32) {
        double* a = Ft + (size_t)q * P.ldw + (pv - j);
        double* b = Ft + (size_t)q * P.ldw + k;
        const double t = *a; *a = *b; *b = t;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_qp3_swapupd(qrdm_prob P, int j, int k) {
  __shared__ double fk[QP3_NB];
  const Qp3State* st = qp3_state(P);
  if (st->stopped) return;
  const int rk = j + k, pvt = st->pvt;
  if (threadIdx.x < k) fk[threadIdx.x] = qp3_ft(P)[(size_t)threadIdx.x * P.ldw + k];
  __syncthreads();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= P.m) return;
  double* ak = P.a + (size_t)rk * P.lda + r;
  double v = *ak;
  if (pvt != rk) {
    double* ap = P.a + (size_t)pvt * P.lda + r;
    const double t = *ap; *ap = v; v = t;
  }
  if (r >= rk) {
    const double* pc = P.a + (size_t)j * P.lda + r;
    for (int q = 0; q < k; ++q) v = fma(-pc[(size_t)q * P.lda], fk[q], v);
  }
  *ak = v;
}

__global__ void __launch_bounds__(256) k_qp3_gemv(qrdm_prob P, int j, int k, int nsplit) {
  __shared__ double xs[QP3_RC];
  const Qp3State* st = qp3_state(P);
  if (st->stopped) return;
  const int rk = j + k, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r_lo = rk + 1, rows = P.m - r_lo;
  const int per = ((rows + nsplit - 1) / nsplit + 31) & ~31;
  const int my_lo = r_lo + blockIdx.y * per, my_hi = min(P.m, my_lo + per);
  const int c0 = j + blockIdx.x * 64;
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0;
  const double* xcol = P.a + (size_t)rk * P.lda;
  for (int base = my_lo; base < my_hi; base += QP3_RC) {
    const int cnt = min(QP3_RC, my_hi - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += 256) xs[i] = xcol[base + i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // warp wid: columns c0 + wid + 8 i
      const int c = c0 + wid + 8 * i;
      if (c < P.n) {
        const double* col = P.a + (size_t)c * P.lda + base;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int t = lane;
        for (; t + 96 < cnt; t += 128) {
          s0 = fma(col[t], xs[t], s0); s1 = fma(col[t + 32], xs[t + 32], s1);
          s2 = fma(col[t + 64], xs[t + 64], s2); s3 = fma(col[t + 96], xs[t + 96], s3);
        }
        for (; t < cnt; t += 32) s0 = fma(col[t], xs[t], s0);
        acc[i] += (s0 + s1) + (s2 + s3);
      }
    }
  }
  double* gp = qp3_gpart(P) + (size_t)blockIdx.y * P.ldw;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + wid + 8 * i;
    const double s = warp_sum(acc[i]);
    if (lane == 0 && c < P.n) gp[c - j] = s;
  }
}

__global__ void __launch_bounds__(256) k_qp3_finish(qrdm_prob P, int j, int k, int nsplit, int lastrk) {
  __shared__ double auxv[QP3_NB], vrow[QP3_NB + 1];
  Qp3State* st = qp3_state(P);
  if (st->stopped) return;
  const int rk = j + k;
  const double* gp = qp3_gpart(P);
  double* Ft = qp3_ft(P);
  auto gsum = [&](int crel) {
    double s = 0.0;
    for (int q = 0; q < nsplit; ++q) s += gp[(size_t)q * P.ldw + crel];
    return s;
  };
  // reflector scalars, recomputed by every thread from the same numbers (dlarfg: beta = -sign(alpha) ||(alpha, x)||)
  const bool rows_below = rk + 1 < P.m;
  const double alpha = P.a[(size_t)rk * P.lda + rk];
  const double xn2 = rows_below ? gsum(k) : 0.0;
  double tau = 0.0, beta = alpha, scale = 1.0;
  if (rows_below && xn2 != 0.0) {
    const double h = sqrt(fma(alpha, alpha, xn2));
    beta = (alpha >= 0.0) ? -h : h;
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  if (threadIdx.x < k) {  // earlier reflectors of the panel: their entry in row rk and auxv = -tau V'v
    const double vr = P.a[(size_t)(j + threadIdx.x) * P.lda + rk];
    vrow[threadIdx.x] = vr;
    auxv[threadIdx.x] = -tau * (vr + (rows_below ? scale * gsum(threadIdx.x) : 0.0));
  }
  __syncthreads();
  const int crel = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.tau[rk] = tau;
    st->tau = tau; st->beta = beta; st->scale = (tau != 0.0) ? scale : 1.0;
    st->kb = k + 1;
  }
  const int c = j + crel;
  if (c >= P.n || crel <= k) return;
  double* arow = P.a + (size_t)c * P.lda + rk;
  const double a_rc = *arow;
  double f = tau * (a_rc + (rows_below ? scale * gsum(crel) : 0.0));
  double upd = 0.0;
  for (int q = 0; q < k; ++q) {
    const double fq = Ft[(size_t)q * P.ldw + crel];
    f = fma(fq, auxv[q], f);
    upd = fma(vrow[q], fq, upd);
  }
  Ft[(size_t)k * P.ldw + crel] = f;
  const double a_new = a_rc - (upd + f);  // v_k[rk] = 1
  *arow = a_new;
  if (rk < lastrk) {  // dlaqps's partial-norm downdate
    const double v1 = P.vn1[c];
    if (v1 != 0.0) {
      double temp = fabs(a_new) / v1;
      temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
      const double ratio = v1 / P.vn2[c];
      const double temp2 = temp * (ratio * ratio);
      if (temp2 <= 1.0536712127723509e-08 /* tol3z = sqrt(dlamch('Epsilon')) = 2^-26.5 */) {
        P.flag_list[c] = 1;
        st->stop_req = 1;  // (everybody stores 1; k_qp3_scale turns it into `stopped` for the next step's kernels)
      } else {
        P.vn1[c] = v1 * sqrt(temp);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_qp3_scale(qrdm_prob P, int j, int k) {
  const Qp3State* st = qp3_state(P);
  Qp3State* stw = qp3_state(P);
  if (st->kb != k + 1) return;  // the step did not run (the panel stopped before it)
  if (blockIdx.x == 0 && threadIdx.x == 0 && st->stop_req) stw->stopped = 1;
  const int rk = j + k;
  const int r = rk + blockIdx.x * 256 + threadIdx.x;
  if (r >= P.m) return;
  double* x = P.a + (size_t)rk * P.lda + r;
  *x = (r == rk) ? st->beta : *x * st->scale;
}

// W2 = -F' in the layout k_rankk reads (row q, column offset relative to j + kb), pending-block geometry in ctrl
__global__ void __launch_bounds__(256) k_qp3_w2(qrdm_prob P, int j, int kb) {
  qrdm_ctrl* ctrl = P.ctrl;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { ctrl->pend_k = kb; ctrl->pend_c0 = j + kb; ctrl->pend_r0 = j + kb; }
  const int q = blockIdx.y;
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= P.ldw) return;
  const int crel = x + kb;
  double v = 0.0;
  if (q < kb && j + crel < P.n) v = -qp3_ft(P)[(size_t)q * P.ldw + crel];
  P.w2[(size_t)q * P.ldw + x] = v;
}

__global__ void __launch_bounds__(128) k_qp3_renorm(qrdm_prob P, int j_next) {  // exact norms of the flagged columns, rows >= j_next
  __shared__ double sc[32];
  const int c = j_next + blockIdx.x;
  if (c >= P.n || !P.flag_list[c]) return;
  const double* col = P.a + (size_t)c * P.lda;
  double s = 0.0;
  for (int r = j_next + threadIdx.x; r < P.m; r += 128) s = fma(col[r], col[r], s);
  s = block_sum(s, sc);
  if (threadIdx.x == 0) { const double nrm = sqrt(s); P.vn1[c] = nrm; P.vn2[c] = nrm; P.flag_list[c] = 0; }
}

// The whole factorisation on device-resident data.  p: a, lda, m, n, jpvt (out, 1-based), tau (out), vn1, vn2, wp, w2,
// ldw, vc, ldv, ctrl, gram (state), flag_list, sm_count, vec16 as filled by the host driver.  mailbox: pinned, >= 64 bytes.
extern "C" int qrdm_qp3_dev(const qrdm_prob* p, void* mailbox, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int m = p->m, n = p->n, minmn = m < n ? m : n;
  qrdm_prob P = *p;
  P.pend = 0; P.sub = 0; P.side_col0 = 0; P.pre_col0 = 0; P.no_vtv = 0; P.stamp = 0x7fffffff;
  cudaMemsetAsync(P.ctrl, 0, sizeof(qrdm_ctrl), s);
  k_qp3_norms<<<n, 256, 0, s>>>(P);
  QRDM_LAUNCH_CHECK();
  const int lastrk = minmn - 1;  // dlaqps: LASTRK = min(M, N + OFFSET), 0-based last pivot row that still downdates
  for (int j = 0; j < minmn;) {
    const int nbp = QP3_NB < minmn - j ? QP3_NB : minmn - j;
    const int ncols = n - j;
    for (int k = 0; k < nbp; ++k) {
      const int rk = j + k, rows = m - rk - 1;
      int nsplit = rows > 0 ? (rows + QP3_RC - 1) / QP3_RC : 1;
      // enough CTAs to fill the GPU when few columns are left, never more splits than the partial buffer holds
      const int ctiles = (ncols + 63) / 64;
      while (nsplit < QP3_MAXSPLIT && ctiles * nsplit < 2 * p->sm_count && rows / (nsplit + 1) >= 256) ++nsplit;
      if (nsplit > QP3_MAXSPLIT) nsplit = QP3_MAXSPLIT;
      k_qp3_pivot<<<1, 1024, 0, s>>>(P, j, k);
      QRDM_LAUNCH_CHECK();
      k_qp3_swapupd<<<(m + 255) / 256, 256, 0, s>>>(P, j, k);
      QRDM_LAUNCH_CHECK();
      if (rows > 0) {
        k_qp3_gemv<<<dim3(ctiles, nsplit), 256, 0, s>>>(P, j, k, nsplit);
        QRDM_LAUNCH_CHECK();
      }
      k_qp3_finish<<<(ncols + 255) / 256, 256, 0, s>>>(P, j, k, nsplit, lastrk);
      QRDM_LAUNCH_CHECK();
      k_qp3_scale<<<(m - rk + 255) / 256, 256, 0, s>>>(P, j, k);
      QRDM_LAUNCH_CHECK();
    }
    cudaError_t e = cudaMemcpyAsync(mailbox, P.gram, sizeof(Qp3State), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
    const int kb = reinterpret_cast<const Qp3State*>(mailbox)->kb;
    if (kb <= 0 || kb > nbp) return (int)cudaErrorUnknown;
    if (j + kb < n && j + kb < m) {  // block update of everything behind the panel
      int rc = qrdm_k_vc_build(&P, P.a, P.lda, j, kb, s);
      if (rc) return rc;
      k_qp3_w2<<<dim3((P.ldw + 255) / 256, 64), 256, 0, s>>>(P, j, kb);
      QRDM_LAUNCH_CHECK();
      qrdm_prob q = P;
      q.pend = 1;
      rc = qrdm_k_rankk(&q, j + kb - 1, s);  // (j_host only bounds the grid: rows >= j + kb are updated)
      if (rc) return rc;
    }
    j += kb;
    if (j < minmn) {
      k_qp3_renorm<<<n - j, 128, 0, s>>>(P, j);
      QRDM_LAUNCH_CHECK();
    }
  }
  return 0;
}
