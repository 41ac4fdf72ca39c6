/* qrdm_b200.h — C ABI of libqrdm_b200.so: the B200-native (sm_100a) drop-in for the
 * factorisation hot path of mdessole/qrdm.
 *
 * The first two entry points are exactly what the reference exports and what its only caller
 * (the CPython wrapper, reference QRDM_wrapper.c:96) binds; the rest are additions for
 * device-resident timing, multi-GPU and statistics.  Plain C linkage, plain pointers and sizes,
 * `int` = the reference's `lapack_int` (reference include/lapacke_config.h:46-50).
 *
 * There is no CPU fallback: every numeric stage runs as a CUDA kernel; without a usable device the
 * calls return QRDM_ERR_CUDA.
 */
#ifndef QRDM_B200_H_
#define QRDM_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QRDM_COL_MAJOR 102 /* reference include/cblas.h:10, src/dgeqrdm_work.c:17-18 */
#define QRDM_ROW_MAJOR 101 /* accepted by the reference's check but never worked there: rejected */
#define QRDM_NB_MAX 256    /* largest supported block size / candidate count (above 64: micro-panels of 64 columns) */

/* error codes beyond the reference's (0, -1 bad argument, -8/-6/-13 NaN screens of
 * LAPACKE_dlarft / LAPACKE_dlarfb_mia, reference src/dlarfb.c:73-86) */
#define QRDM_ERR_CUDA (-100)       /* CUDA runtime failure (no device, launch error, OOM) */
#define QRDM_ERR_COMM (-101)       /* NCCL failure in the row-sharded path */
#define QRDM_ERR_UNSUPPORTED (-102) /* nb > QRDM_NB_MAX (nb > 64 in the row-sharded and batched entry points) */

/* Threading: the reference keeps no global state (src/dgeqrdm_work.c:650-665, 816-826) and may be called from
 * several host threads at once.  Here the device workspace is per process; every entry point below takes one
 * process-wide recursive lock, so concurrent callers are safe and are served one after the other.  The workspace
 * follows the caller's current CUDA device: a call made after cudaSetDevice(other) re-creates it there. */

/* Replaces: reference include/QRDM.h:19-22, src/dgeqrdm.c:5-16.
 * QR factorisation with Deviation-Maximisation block pivoting, A P = Q R, FP64.
 *   matrix_layout  QRDM_COL_MAJOR (102)
 *   a              m x n, column-major, lda >= m, HOST memory; overwritten with R (upper
 *                  triangle of the first r = sum(ncols) columns), the Householder vectors below
 *                  it (dgeqrf convention) and the Q'-updated R12/R22 in columns >= r
 *   jpvt[n]        in: zero = free column, non-zero = FIXED column (LAPACK dgeqp3 convention, reference
 *                  src/dgeqrdm_work.c:592-635): fixed columns are moved to the front with the reference's swap sequence,
 *                  factored without pivoting, and DM pivoting runs on the rest.  out: 1-based permutation
 *   tau[min(m,n)]  out: reflector scalars, first r entries
 *   ncols[n]       in: ncols[0] = stop rule (0 none, 1 eps*n, 2 eps*sqrt(n), 3 thres[2]);
 *                  out: ncols[it] = columns triangularised in DM iteration it; revealed rank = (number of fixed
 *                  columns) + sum (the reference counts DM iterations only); at least min(m, n) entries
 *   thres          thres[0] = delta (cosine bound), thres[1] = tau_ (norm fraction), [2] = eta
 *   nb             maximum block size / number of candidates, 1..QRDM_NB_MAX (= 256; above 64 the selected block is
 *                  factored in micro-panels of 64 columns)
 * Returns info as the reference does (0; -1 after an xerbla-style message for bad arguments;
 * -8/-6/-13 when NaNs reach the block reflector). */
int dgeqrdm(int matrix_layout, int m, int n, double *a, int lda, int *jpvt, double *tau, int *ncols,
            double *thres, int nb);

/* Replaces: reference src/dgeqrdm_work.c:420-425 (same contract; dgeqrdm forwards to it). */
int dgeqrdm_work(int matrix_layout, int m, int n, double *a, int lda, int *jpvt, double *tau,
                 int *ncols, double *thres, int nb);

/* Addition: device-resident variant used for compute-only timing and by callers that already
 * hold A in HBM.  d_a, d_jpvt (n ints) and d_tau (min(m,n) doubles) are DEVICE pointers on the
 * current device; ncols and thres are HOST arrays.  d_jpvt need not be initialised (treated as
 * all-free).  `stream` is a cudaStream_t passed as void* (NULL = default stream).  The call
 * returns after the factorisation has completed on the stream. */
int dgeqrdm_dev(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, int *ncols,
                const double *thres, int nb, void *stream);

/* Addition: batched mode — `batch` independent m x n HOST matrices, matrix b
 * at a + b*stride_a; jpvt (b*n), tau (b*min(m,n)), ncols (b*n, [b*n] = stop-rule mode on entry) laid out
 * per matrix; infos[b] per matrix (may be NULL).  Returns 0 or the first non-zero info.  Matrices with
 * m, n <= 1024 are factored by ONE kernel launch per chunk of 592 matrices, one CTA per matrix, with the
 * upload of the next chunk and the download of the previous one overlapped; larger ones run one after the
 * other through the one-matrix path.  Across GPUs each rank passes its share (independent units). */
int dgeqrdm_batched(int batch, int m, int n, double *a, int lda, long long stride_a, int *jpvt, double *tau,
                    int *ncols, double *thres, int nb, int *infos);

/* Device-resident batched variant (m, n <= 1024, else QRDM_ERR_UNSUPPORTED): every array is a DEVICE
 * pointer except thres, which MUST hold 3 doubles here (the stop modes live on the device, so thres[2] is always
 * read; the host-pointer variant above reads it only if some matrix asks for stop mode 3); d_ncols is [batch][n] with the stop-rule mode in [b][0] on entry and the block
 * sizes on exit; d_infos [batch] (may be NULL) receives the per-matrix info.  One launch; returns after
 * it has completed on the stream. */
int dgeqrdm_batched_dev(int batch, int m, int n, double *d_a, int lda, long long stride_a, int *d_jpvt,
                        double *d_tau, int *d_ncols, int *d_infos, const double *thres, int nb, void *stream);

/* Addition: 1-D block-row sharded factorisation across the GPUs of one node, one
 * process per GPU.  Rank p passes its rows [row0, row0 + m_local) of all n columns (device memory,
 * column-major, lda >= m_local); d_jpvt / d_tau / ncols come back replicated on every rank.  The
 * all-reduces (column-norm partials, candidate Gram, one 128-double vector per panel column, the
 * V'C slots) run on `stream`.  Two transports (both may be set up; QRDM_B200_COLL=peer|nccl forces one):
 *   NCCL   rank 0 calls qrdm_b200_comm_unique_id, the 128 bytes are broadcast by the launcher (e.g.
 *          torch.distributed), every rank calls qrdm_b200_comm_init.  Carries the bandwidth-bound vectors.
 *   peer   NVLink peer memory (CUDA IPC, one node, <= 8 ranks): every rank calls qrdm_b200_peer_handle, the launcher
 *          all-gathers the 64-byte handles (rank order), every rank calls qrdm_b200_peer_open and the launcher
 *          runs one barrier.  With peer memory open the panel's per-column reduction happens INSIDE one persistent
 *          kernel per 8-column sub-panel (LL packets stored straight into the peers' buffers) and the small vectors
 *          go through a one-kernel LL all-reduce; without it the round-1 path (one kernel + one 1-KB ncclAllReduce
 *          per panel column) is used. */
int qrdm_b200_comm_unique_id(char *out128);
int qrdm_b200_comm_init(int rank, int nranks, const char *id128);
void qrdm_b200_comm_destroy(void);
int qrdm_b200_peer_handle(char *out64);
int qrdm_b200_peer_open(int rank, int nranks, const char *handles64);
void qrdm_b200_peer_close(void);
int dgeqrdm_dev_sharded(int m_local, int m_global, int row0, int nranks, int n, double *d_a, int lda,
                        int *d_jpvt, double *d_tau, int *ncols, const double *thres, int nb, void *stream);

/* Addition: apply Q = H_0 H_1 ... H_{k-1} (trans = 'N') or Q' (trans = 'T') from the left to an
 * m x n matrix C, the reflectors being the first k columns of a matrix factored by dgeqrdm / dgeqrf (vectors
 * below the diagonal, unit diagonal implicit) and tau.  Replaces the LAPACKE_dormqr('L', trans, ...) behind the
 * reference wrapper's DORMQR (reference QRDM_wrapper.c:104-126), which auxil.checkQR (reference auxil.py:20-105)
 * uses to form Q and Q R; runs the K6 trailing-update kernels block by block (64 reflectors per block).
 * _dev: every pointer is a DEVICE pointer, column-major, lda/ldc >= m.  Returns 0 (or -13 if NaNs went through). */
int qrdm_b200_dormqr_dev(char trans, int m, int n, int k, const double *d_a, int lda, const double *d_tau, double *d_c,
                         int ldc, void *stream);
int qrdm_b200_dormqr(char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc);

/* QR with classical column pivoting on the GPU: LAPACK dgeqp3's blocked algorithm (dlaqps) — what the
 * reference exports as dgeqp3 (src/dgeqp3.c:39-93, a renamed LAPACKE_dgeqp3) and its wrapper calls as QP3
 * (QRDM_wrapper.c:15-41) — so that dgeqrdm is compared with QRCP on the same device.  Column-major, all columns free
 * (the host variant rejects jpvt[c] != 0 on entry with QRDM_ERR_UNSUPPORTED); on exit A holds R and the Householder
 * vectors, tau the min(m, n) scalars, jpvt the 1-based permutation.  _dev: device pointers. */
int qrdm_b200_dgeqp3(int m, int n, double *a, int lda, int *jpvt, double *tau);
int qrdm_b200_dgeqp3_dev(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, void *stream);

/* Addition: per-call statistics of the last dgeqrdm*() on this thread's device. */
typedef struct qrdm_b200_stats {
  int iterations;        /* DM iterations (= number of ncols entries written) */
  int rank;              /* sum of ncols */
  long long launches;    /* CUDA kernels launched by the call */
  double ms_total;       /* device time of the factorisation proper (CUDA events) */
  double ms_h2d, ms_d2h; /* host<->device copies of the host-pointer entry points */
  /* per-stage device time in ms, only filled when profiling is enabled (QRDM_B200_PROFILE=1
   * or qrdm_b200_set_profile(1)): stage order = QRDM_STAGE_* */
  double ms_stage[12];
  long long stage_launches[12];
  double trailing_flops; /* FLOPs executed by the trailing-update kernels (4*rows*cols*k summed) */
  double panel_cols;     /* total panel columns processed */
  /* algorithmic bytes per stage, summed by the host from the mailbox: K1 8mn; K3b 8*m_r*nc;
   * K3d 32*m per exchange; K4 16*m_r*fjb; K2 8*k*n_r; K6 16*m_r*n_c (deferred schedule) or 24*m_r*n_c */
  double stage_bytes[12];
  /* look-ahead: FLOPs of the pending blocks' pass 2 that the side stream applied beside the selection / panel of the
   * next block (part of trailing_flops; NOT executed inside the trailing stage, whose time is ms_stage[TRAILING]); the
   * side kernels' own event time is ms_stage[RANKK] (includes their waits for SMs: they run at the least priority) */
  double side_flops;
  long long side_launches;
  /* the dominant kernel alone: FLOPs executed inside the k_fused launches (pass 1 of the current block on all trailing
   * columns + pass 2 of the pending block on the columns neither the eager set nor the side stream completed; k_vtc for
   * the first deferred block), whose own event time is ms_stage[VTV] in profile mode 2 */
  double fused_flops;
  long long fused_launches;
} qrdm_b200_stats;

enum {
  QRDM_STAGE_NORM_INIT = 0,
  QRDM_STAGE_SELECT,
  QRDM_STAGE_GRAM,
  QRDM_STAGE_PICK,
  QRDM_STAGE_PERMUTE,
  QRDM_STAGE_PANEL,
  QRDM_STAGE_VTV,
  QRDM_STAGE_VTC,
  QRDM_STAGE_WSOLVE,
  QRDM_STAGE_RANKK,
  QRDM_STAGE_NORM_UPDATE,
  QRDM_STAGE_SYNC,
  QRDM_STAGE_COUNT
};

void qrdm_b200_get_stats(qrdm_b200_stats *out);
void qrdm_b200_set_profile(int mode); /* 0 off, 1 per-stage with syncs, 2 light (panel + trailing, no syncs) */

/* Optional: create / release the per-process device workspace ahead of the first call
 * (otherwise created lazily and grown on demand).  init returns 0 or QRDM_ERR_CUDA. */
int qrdm_b200_init(int device);
void qrdm_b200_shutdown(void);

/* Micro-benchmarks used by bench.py for the roofline denominators (measured live, on the stream
 * given): FP64 peak in TFLOP/s (use_dmma = 1: DMMA.8x8x4 chains, 0: DFMA chains), and a device-to-device copy of
 * `bytes` bytes in GB/s (bytes read + bytes written per second; best of 5). */
double qrdm_b200_measure_fp64_peak(int use_dmma, void *stream);
double qrdm_b200_measure_copy_gbs(size_t bytes, void *stream);
const char *qrdm_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QRDM_B200_H_ */
