/* dgeqrdm_host.c — host orchestration of the B200-native dgeqrdm, plain C over the thin CUDA
 * C-ABI layer of qrdm_dev.h (BASELINE.json north_star: "Host orchestration stays in C").
 *
 * Replaces the driver of the reference: dgeqrdm (src/dgeqrdm.c:5-16) and dgeqrdm_work
 * (src/dgeqrdm_work.c:420-838): argument checks (:559-589), stop-rule set-up (:531-543, 684),
 * the main block loop (:694-787) and its exit test (:782-785).  Every numeric stage is a CUDA
 * kernel; there is no CPU fallback.  The loop's only host<->device synchronisation is one 64-byte
 * mailbox copy per iteration (block size, error flag, max partial norm), needed because the block
 * size — hence the trip count and the stop rule — is data dependent.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/qrdm_b200.h"
#include "hostio.h"
#include "qrdm_dev.h"

#define QRDM_VERSION "qrdm_b200 0.2 (round 2)"

typedef struct {
  int ready, sm_count;
  int device; /* the device the workspace lives on (a call on another current device re-creates it) */
  /* capacities */
  int cap_m, cap_n;
  size_t cap_a_bytes;
  /* device buffers */
  qrdm_ctrl *ctrl;
  double *vn1, *vn2, *gram_part, *gram, *panel_part, *panel_row, *vc, *wp, *w2, *nrm_part;
  int *flag_list, *upd_marks; /* upd_marks: [2][cap_n] stamps of the deferred trailing update */
  double *mg_buf;
  unsigned *mg_cnt;
  /* nb > 64 (k_wide.cu), allocated on first use */
  double *wide_part, *wide_gram;
  int *wide_marks, *wide_swaps;
  int wide_marks_cap;
  size_t wp_elems;
  int ldv, ldw, nrm_splits;
  /* staging for the host-pointer entry points */
  double *d_a, *d_tau;
  int *d_jpvt;
  size_t cap_tau, cap_jpvt;
  /* pinned mailbox + events */
  qrdm_ctrl *mailbox;
  void *ev[4];
  void *ev_stage[2];
  void *compute_stream, *copy_stream; /* non-blocking streams of the host-pointer entry points */
  void *h2d_stream;                   /* third stream of the batched pipeline (created on first use) */
  void *side_stream;                  /* look-ahead: pass 2 of the pending block beside the next selection / panel (LEAST priority) */
  void *hp_stream;                    /* look-ahead: the factorisation itself runs here (GREATEST priority), see factor_device */
  void *ev_hop[2];                    /* caller's stream -> hp_stream at entry, hp_stream -> caller's stream at exit */
  void *ev_side[2];                   /* [0] main -> side (stamps are set), [1] side -> main (side update finished) */
  int side_busy;                      /* a side launch is in flight that the main stream has not waited for yet */
  qrdm_hostio *io;                    /* bounce pipeline for pageable host buffers (created on first use) */
} qrdm_workspace;

/* 1-D block-row sharding: this rank holds global rows [row0, row0 + m_local) */
typedef struct {
  int row0, m_glob, nranks;
} qrdm_shard;

/* optional streaming write-back of finished columns (host-pointer entry point, pinned buffers) */
typedef struct {
  double *h_a;
  int h_lda;
  void *copy_stream;
  int done_cols; /* columns [0, done_cols) already enqueued for D2H */
  qrdm_hostio *io; /* non-NULL: pageable host buffer, finished columns go through the pinned ring of hostio.c */
} qrdm_writeback;

static qrdm_workspace g_ws;
static qrdm_b200_stats g_stats;
static int g_profile = -1;

/* The reference allocates and frees its workspace per call and keeps no global state (src/dgeqrdm_work.c:650-665,
 * 816-826), so two host threads may call it concurrently.  Here the device workspace, the statistics and the LL
 * epochs of the kernel launchers are per process: every public entry point runs under one recursive lock, which
 * makes concurrent callers safe (they are serialised — one GPU is one resource anyway). */
static pthread_mutex_t g_api_lock;
static pthread_once_t g_api_once = PTHREAD_ONCE_INIT;
static void api_lock_init(void) {
  pthread_mutexattr_t at;
  pthread_mutexattr_init(&at);
  pthread_mutexattr_settype(&at, PTHREAD_MUTEX_RECURSIVE);
  pthread_mutex_init(&g_api_lock, &at);
  pthread_mutexattr_destroy(&at);
}
static void api_lock(void) { pthread_once(&g_api_once, api_lock_init); pthread_mutex_lock(&g_api_lock); }
static void api_unlock(void) { pthread_mutex_unlock(&g_api_lock); }
#define API_BODY(call) do { api_lock(); int r__ = (call); api_unlock(); return r__; } while (0)

static int roundup(int x, int a) { return (x + a - 1) / a * a; }

#define CU(call)                                                                          \
  do {                                                                                    \
    int e__ = (call);                                                                     \
    if (e__ != 0) {                                                                       \
      fprintf(stderr, "qrdm_b200: CUDA error %d (%s) at %s:%d\n", e__, qrdm_rt_errstr(e__), \
              __FILE__, __LINE__);                                                        \
      return QRDM_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

const char *qrdm_b200_version(void) { return QRDM_VERSION; }
void qrdm_b200_get_stats(qrdm_b200_stats *out) { api_lock(); *out = g_stats; api_unlock(); }
void qrdm_b200_set_profile(int mode) { api_lock(); g_profile = mode < 0 ? 0 : (mode > 2 ? 2 : mode); api_unlock(); }
double qrdm_b200_measure_fp64_peak(int use_dmma, void *stream) {
  api_lock();
  const double r = qrdm_rt_fp64_peak(use_dmma, stream);
  api_unlock();
  return r;
}
double qrdm_b200_measure_copy_gbs(size_t bytes, void *stream) {
  api_lock();
  const double r = qrdm_rt_copy_gbs(bytes, stream);
  api_unlock();
  return r;
}

static void ws_free_sized(qrdm_workspace *w) {
  void *bufs[] = {w->vn1, w->vn2, w->vc, w->wp, w->w2, w->nrm_part, w->flag_list, w->upd_marks};
  for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i)
    if (bufs[i]) qrdm_rt_free(bufs[i]);
  w->vn1 = w->vn2 = w->vc = w->wp = w->w2 = w->nrm_part = NULL;
  w->flag_list = NULL;
  w->upd_marks = NULL;
  w->cap_m = w->cap_n = 0;
}

static void shutdown_impl(void) {
  qrdm_workspace *w = &g_ws;
  if (!w->ready) return;
  int cur = -1;
  qrdm_rt_get_device(&cur);
  if (cur != w->device) qrdm_rt_set_device(w->device); /* streams / events are destroyed on their own device */
  qrdm_rt_peer_destroy();
  ws_free_sized(w);
  void *fixed[] = {w->ctrl, w->gram_part, w->gram, w->panel_part, w->panel_row, w->d_a, w->d_tau, w->d_jpvt, w->mg_buf, w->mg_cnt,
                   w->wide_part, w->wide_gram, w->wide_marks, w->wide_swaps};
  for (size_t i = 0; i < sizeof(fixed) / sizeof(fixed[0]); ++i)
    if (fixed[i]) qrdm_rt_free(fixed[i]);
  if (w->mailbox) qrdm_rt_host_free(w->mailbox);
  for (int i = 0; i < 4; ++i) qrdm_rt_event_destroy(w->ev[i]);
  for (int i = 0; i < 2; ++i) qrdm_rt_event_destroy(w->ev_stage[i]);
  if (w->compute_stream) qrdm_rt_stream_destroy(w->compute_stream);
  if (w->copy_stream) qrdm_rt_stream_destroy(w->copy_stream);
  if (w->h2d_stream) qrdm_rt_stream_destroy(w->h2d_stream);
  if (w->side_stream) qrdm_rt_stream_destroy(w->side_stream);
  if (w->hp_stream) qrdm_rt_stream_destroy(w->hp_stream);
  for (int i = 0; i < 2; ++i) if (w->ev_side[i]) qrdm_rt_event_destroy(w->ev_side[i]);
  for (int i = 0; i < 2; ++i) if (w->ev_hop[i]) qrdm_rt_event_destroy(w->ev_hop[i]);
  if (w->io) qrdm_hostio_destroy(w->io);
  memset(w, 0, sizeof(*w));
  if (cur >= 0 && cur != w->device) qrdm_rt_set_device(cur);
}
void qrdm_b200_shutdown(void) { api_lock(); shutdown_impl(); api_unlock(); }

static int init_impl(int device) {
  qrdm_workspace *w = &g_ws;
  if (device >= 0) CU(qrdm_rt_set_device(device));
  int cur = 0;
  CU(qrdm_rt_get_device(&cur));
  if (w->ready && w->device == cur) return 0;
  if (w->ready) { /* the caller switched devices: buffers, streams and kernel attributes belong to the old one */
    shutdown_impl();
    CU(qrdm_rt_set_device(cur));
  }
  qrdm_rt_new_device_generation(); /* launchers re-apply their per-device function attributes */
  w->device = cur;
  size_t freeb = 0;
  CU(qrdm_rt_device_info(&w->sm_count, &freeb));
  CU(qrdm_rt_malloc((void **)&w->ctrl, sizeof(qrdm_ctrl)));
  CU(qrdm_rt_malloc((void **)&w->gram_part, sizeof(double) * 4096 * QRDM_GRAM_MAXCTA));
  CU(qrdm_rt_malloc((void **)&w->gram, sizeof(double) * 4096));
  CU(qrdm_rt_malloc((void **)&w->panel_part, 16 * (QRDM_PANEL_PART_PKTS + QRDM_PANEL_GPART_PKTS))); /* LL packets */
  CU(qrdm_rt_memset(w->panel_part, 0, 16 * (QRDM_PANEL_PART_PKTS + QRDM_PANEL_GPART_PKTS), NULL));
  CU(qrdm_rt_malloc((void **)&w->panel_row, 16 * (QRDM_PANEL_ROW_PKTS + QRDM_PANEL_GROW_PKTS)));
  CU(qrdm_rt_memset(w->panel_row, 0, 16 * (QRDM_PANEL_ROW_PKTS + QRDM_PANEL_GROW_PKTS), NULL));
  CU(qrdm_rt_malloc((void **)&w->mg_buf, sizeof(double) * (512 + 128 * 2 * 160)));
  CU(qrdm_rt_malloc((void **)&w->mg_cnt, 64));
  CU(qrdm_rt_memset(w->mg_cnt, 0, 64, NULL));
  CU(qrdm_rt_host_alloc((void **)&w->mailbox, sizeof(qrdm_ctrl)));
  for (int i = 0; i < 4; ++i) CU(qrdm_rt_event_create(&w->ev[i]));
  for (int i = 0; i < 2; ++i) CU(qrdm_rt_event_create(&w->ev_stage[i]));
  CU(qrdm_rt_stream_create(&w->compute_stream));
  CU(qrdm_rt_stream_create(&w->copy_stream));
  CU(qrdm_rt_stream_create_prio(&w->side_stream, 0));
  CU(qrdm_rt_stream_create_prio(&w->hp_stream, 1));
  for (int i = 0; i < 2; ++i) CU(qrdm_rt_event_create(&w->ev_side[i]));
  for (int i = 0; i < 2; ++i) CU(qrdm_rt_event_create(&w->ev_hop[i]));
  w->side_busy = 0;
  w->ready = 1;
  if (g_profile < 0) {
    const char *e = getenv("QRDM_B200_PROFILE");
    g_profile = e ? atoi(e) : 0;
    if (g_profile < 0 || g_profile > 2) g_profile = 0;
  }
  return 0;
}
int qrdm_b200_init(int device) { API_BODY(init_impl(device)); }

static int ws_ensure(int m, int n) {
  qrdm_workspace *w = &g_ws;
  int rc = init_impl(-1);
  if (rc) return rc;
  if (m <= w->cap_m && n <= w->cap_n) return 0;
  int cm = m > w->cap_m ? m : w->cap_m, cn = n > w->cap_n ? n : w->cap_n;
  ws_free_sized(w);
  w->ldv = roundup(cm, QRDM_ROWALIGN) + 128; /* k_rankk row tiles may overhang by < 128 rows */
  w->ldw = roundup(cn, 128) + 128;           /* k_rankk / k_vtc column tiles overhang likewise */
  w->nrm_splits = 64;
  /* partial-W capacity: enough splits for ~2 CTAs per SM at any trailing width (see k_trailing) */
  /* (second term: with the V'V tile skipped (qrdm_prob::no_vtv) a single 128-column tile can carry all 2 x SMs slots:
   * slots x 64 x stride <= 64 x 128 x (grid / t + 2)(t + 2) for t real tiles, largest at t = 1) */
  w->wp_elems = (size_t)64 * w->ldw * 16 + (size_t)64 * 128 * (6 * (size_t)w->sm_count + 64);
  CU(qrdm_rt_malloc((void **)&w->vn1, sizeof(double) * cn));
  CU(qrdm_rt_malloc((void **)&w->vn2, sizeof(double) * cn));
  CU(qrdm_rt_malloc((void **)&w->vc, sizeof(double) * (size_t)w->ldv * 64 * 2)); /* two buffers: current / pending block */
  CU(qrdm_rt_malloc((void **)&w->wp, sizeof(double) * w->wp_elems));
  CU(qrdm_rt_malloc((void **)&w->w2, sizeof(double) * (size_t)w->ldw * 64));
  CU(qrdm_rt_malloc((void **)&w->nrm_part, sizeof(double) * (size_t)cn * w->nrm_splits));
  CU(qrdm_rt_malloc((void **)&w->flag_list, sizeof(int) * (size_t)cn));
  CU(qrdm_rt_malloc((void **)&w->upd_marks, sizeof(int) * (size_t)cn * 2));
  if ((size_t)(cm / 2048 + 2) * 512 > (size_t)4096 * QRDM_GRAM_MAXCTA) { /* skinny-update partials of very tall matrices */
    if (w->gram_part) qrdm_rt_free(w->gram_part);
    w->gram_part = NULL;
    CU(qrdm_rt_malloc((void **)&w->gram_part, sizeof(double) * (size_t)(cm / 2048 + 2) * 512));
  }
  w->cap_m = cm;
  w->cap_n = cn;
  return 0;
}

/* xerbla-style message of the reference (src/dgeqrdm_work.c:577-581) */
static int bad_argument(int pos) {
  fprintf(stderr, " ** On entry to DGEQRDM parameter number %2d had an illegal value\n", pos);
  return -1;
}

static int check_args(int matrix_layout, int m, int n, int lda, const double *thres, int nb) {
  if (matrix_layout != QRDM_COL_MAJOR) return bad_argument(1); /* 101 never worked upstream */
  if (m <= 0) return bad_argument(2);
  if (n <= 0) return bad_argument(3);
  if (lda < (m > 1 ? m : 1)) return bad_argument(5);
  if (thres[0] < 0.0 || thres[0] > 1.0) return bad_argument(9);
  if (thres[1] < 0.0 || thres[1] > 1.0) return bad_argument(9);
  if (nb <= 0) return bad_argument(10);
  if (nb > QRDM_NB_MAX) {
    fprintf(stderr, "qrdm_b200: nb = %d > %d is not supported\n", nb, QRDM_NB_MAX);
    return QRDM_ERR_UNSUPPORTED;
  }
  return 0;
}

/* Profiling modes (QRDM_B200_PROFILE / qrdm_b200_set_profile):
 *   0 off;  1 every stage timed with an event pair and a sync (perturbs the run: for stage splits);
 *   2 light: only the panel and trailing-update stages get event pairs from a pool, no syncs —
 *     the mode bench.py uses to time the roofline kernel inside an otherwise undisturbed run. */
#define EV_POOL 8192
static void *g_ev_pool[EV_POOL];
static int g_ev_stage_of[EV_POOL / 2];
static int g_ev_used = 0, g_ev_created = 0;

static int stage_begin(int stage, void *stream) {
  if (g_profile == 1) return qrdm_rt_event_record(g_ws.ev_stage[0], stream);
  if (g_profile == 2 && (stage == QRDM_STAGE_PANEL || stage == QRDM_STAGE_VTC || stage == QRDM_STAGE_RANKK || stage == QRDM_STAGE_VTV) && g_ev_used + 2 <= EV_POOL) {
    while (g_ev_created < g_ev_used + 2) {
      int e = qrdm_rt_event_create(&g_ev_pool[g_ev_created]);
      if (e) return e;
      ++g_ev_created;
    }
    g_ev_stage_of[g_ev_used / 2] = stage;
    return qrdm_rt_event_record(g_ev_pool[g_ev_used], stream);
  }
  return 0;
}
static int stage_end(int stage, long long launches_before, void *stream) {
  if (g_profile == 2 && (stage == QRDM_STAGE_PANEL || stage == QRDM_STAGE_VTC || stage == QRDM_STAGE_RANKK || stage == QRDM_STAGE_VTV) && g_ev_used + 2 <= EV_POOL) {
    int e = qrdm_rt_event_record(g_ev_pool[g_ev_used + 1], stream);
    g_ev_used += 2;
    g_stats.stage_launches[stage] += qrdm_rt_launch_count() - launches_before;
    return e;
  }
  if (g_profile == 1) {
    int e = qrdm_rt_event_record(g_ws.ev_stage[1], stream);
    if (e) return e;
    e = qrdm_rt_event_sync(g_ws.ev_stage[1]);
    if (e) return e;
    g_stats.ms_stage[stage] += qrdm_rt_event_ms(g_ws.ev_stage[0], g_ws.ev_stage[1]);
    g_stats.stage_launches[stage] += qrdm_rt_launch_count() - launches_before;
  }
  return 0;
}
#define STAGE(id, call)                               \
  do {                                                \
    long long lb__ = qrdm_rt_launch_count();          \
    CU(stage_begin(id, stream));                      \
    CU(call);                                         \
    CU(stage_end(id, lb__, stream));                  \
  } while (0)

/* Sum all-reduce of the row-sharded path.  Two transports:
 *   peer  one-shot LL all-reduce over NVLink peer memory (k_peer.cu): one kernel, ~one NVLink store + one L2 poll of
 *         latency — used for everything up to QRDM_COLL_PEER_MAX doubles (the per-iteration vectors of a tall-skinny
 *         factorisation: Gram 4096, skinny products 512, folded W <= 64 x ~600, norm sums n);
 *   nccl  ncclAllReduce (bandwidth-bound messages: W of wide matrices, norm partials of many columns; and whenever the
 *         launcher did not open peer memory).
 * QRDM_B200_COLL=peer|nccl forces one of them (tests run 2 ranks on ONE GPU with "peer": NCCL refuses that). */
#define QRDM_COLL_PEER_MAX ((size_t)96 * 1024)
static int mg_allreduce(double *buf, size_t count, void *stream) {
  static int mode = -1; /* 0 auto, 1 peer, 2 nccl */
  if (mode < 0) {
    const char *e = getenv("QRDM_B200_COLL");
    mode = !e ? 0 : (strcmp(e, "peer") == 0 ? 1 : (strcmp(e, "nccl") == 0 ? 2 : 0));
  }
  const int have_peer = qrdm_rt_peer_available() > 0;
  if (mode == 1 && !have_peer) {
    fprintf(stderr, "qrdm_b200: QRDM_B200_COLL=peer but no peer memory is open (qrdm_b200_peer_open)\n");
    return -1;
  }
  if (have_peer && (mode == 1 || (mode == 0 && count <= QRDM_COLL_PEER_MAX))) return qrdm_k_peer_allreduce(buf, count, stream);
  return qrdm_rt_allreduce(buf, count, stream);
}

static int read_mailbox(const qrdm_prob *p, void *stream);

/* Badly scaled input (largest column norm Inf, or below 2^-300 where squares underflow): find max |a_ij|, and if it is
 * finite and far from 1 multiply A by the power of two that brings it to [1, 2).  Exact, so pivots / tau / V are those
 * of the unscaled matrix and R is off by the same factor, which factor_device divides out at the end.  The reference
 * gets there with scaled norms (cblas_dnrm2) and dlarfg's safmin loop (src/dlarfg.c:144-182). */
static int mg_allreduce(double *buf, size_t count, void *stream);
static int prescale_input(qrdm_prob *P, int mg, double *scale_out, void *stream) {
  qrdm_workspace *w = &g_ws;
  int nparts = 0;
  *scale_out = 1.0;
  CU(qrdm_k_amax(P, P->nrm_part, &nparts, stream));
  double *h = (double *)malloc(sizeof(double) * (size_t)(nparts > 16 ? nparts : 16));
  if (!h) return QRDM_ERR_CUDA;
  int rc = qrdm_rt_d2h(h, P->nrm_part, sizeof(double) * (size_t)nparts, stream);
  if (!rc) rc = qrdm_rt_sync(stream);
  double amax = 0.0;
  for (int i = 0; i < nparts && !rc; ++i) amax = h[i] > amax ? h[i] : amax;
  if (!rc && mg && P->nranks > 1) {
    /* the all-reduce sums: sum_r amax_r lies in [max, nranks * max] — as good as the max for picking a power of two,
     * and bit-identical on every rank, so all ranks scale by the same factor */
    double v[2] = {amax, 0.0};
    rc = qrdm_rt_h2d(P->mg_buf, v, sizeof(v), stream);
    if (!rc) rc = mg_allreduce(P->mg_buf, 2, stream) ? QRDM_ERR_COMM : 0;
    if (!rc) rc = qrdm_rt_d2h(v, P->mg_buf, sizeof(v), stream);
    if (!rc) rc = qrdm_rt_sync(stream);
    if (!rc) amax = v[0];
  }
  free(h);
  if (rc) return rc == QRDM_ERR_COMM ? rc : QRDM_ERR_CUDA;
  if (!(amax > 0.0) || isinf(amax) || isnan(amax)) return 0; /* zero matrix, Inf or NaN input: nothing to gain */
  const int e = ilogb(amax);
  if (e > -200 && e < 200) return 0;                         /* a lone tiny or huge column: the matrix itself is fine */
  const double sc = ldexp(1.0, -e);
  CU(qrdm_k_scale(P, sc, 0, 0, stream));
  *scale_out = sc;
  (void)w;
  return 0;
}

static int read_mailbox(const qrdm_prob *p, void *stream) {
  CU(qrdm_rt_d2h(g_ws.mailbox, p->ctrl, QRDM_MAILBOX_BYTES, stream));
  CU(qrdm_rt_sync(stream));
  return 0;
}

/* nb > 64: workspace of the wide selection (k_wide.cu) */
static int ws_ensure_wide(int n) {
  qrdm_workspace *w = &g_ws;
  if (!w->wide_part) {
    const size_t pairs = (QRDM_CANDMAX / 64) * (QRDM_CANDMAX / 64 + 1) / 2;
    CU(qrdm_rt_malloc((void **)&w->wide_part, sizeof(double) * pairs * QRDM_WIDE_ROWCTAS * 4096));
    CU(qrdm_rt_malloc((void **)&w->wide_gram, sizeof(double) * QRDM_CANDMAX * QRDM_CANDMAX));
    CU(qrdm_rt_malloc((void **)&w->wide_swaps, sizeof(int) * (2 * (2 * QRDM_CANDMAX + 4) + 1)));
  }
  if (n > w->wide_marks_cap) {
    if (w->wide_marks) qrdm_rt_free(w->wide_marks);
    w->wide_marks = NULL;
    w->wide_marks_cap = 0;
    CU(qrdm_rt_malloc((void **)&w->wide_marks, sizeof(int) * (size_t)n));
    w->wide_marks_cap = n;
  }
  return 0;
}

/* The factorisation proper on device-resident data. */
static int factor_device_impl(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, int *ncols,
                              const double *thres, int nb, void *stream, qrdm_writeback *wb,
                              const qrdm_shard *sh, int nfxd_in);

/* The factorisation runs on the library's own GREATEST-priority stream, ordered after everything the caller has put on
 * `stream` and before everything the caller puts there afterwards (two event hops).  That is what lets the look-ahead's
 * side stream (LEAST priority, many short CTAs) share the GPU with it: the block scheduler serves the pending CTAs of
 * the selection / panel kernels first — it even keeps whole SMs free for a panel CTA while side CTAs would still fit
 * (measured, tools/prio_probe.cu: 120 full-SM CTAs of a high-priority kernel all start within 11 us of their launch
 * into a saturated GPU, and the low-priority grid carries on on the 28 SMs they leave; with equal priorities they wait
 * for the whole low grid).  On every exit, error paths included, the caller's stream waits for both streams. */
static int factor_device(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, int *ncols,
                         const double *thres, int nb, void *stream, qrdm_writeback *wb,
                         const qrdm_shard *sh, int nfxd_in) {
  int rc = init_impl(-1);
  if (rc) return rc;
  qrdm_workspace *w = &g_ws;
  CU(qrdm_rt_event_record(w->ev_hop[0], stream));
  CU(qrdm_rt_stream_wait_event(w->hp_stream, w->ev_hop[0]));
  rc = factor_device_impl(m, n, d_a, lda, d_jpvt, d_tau, ncols, thres, nb, w->hp_stream, wb, sh, nfxd_in);
  if (w->side_busy) { /* an error return left a side launch in flight */
    if (qrdm_rt_stream_wait_event(w->hp_stream, w->ev_side[1]) == 0) w->side_busy = 0;
  }
  CU(qrdm_rt_event_record(w->ev_hop[1], w->hp_stream));
  CU(qrdm_rt_stream_wait_event(stream, w->ev_hop[1]));
  return rc;
}

static int factor_device_impl(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, int *ncols,
                              const double *thres, int nb, void *stream, qrdm_writeback *wb,
                              const qrdm_shard *sh, int nfxd_in) {
  qrdm_workspace *w = &g_ws;
  const double eps = DBL_EPSILON * 0.5; /* dlamch('e'), src/dgeqrdm_work.c:528 */
  /* row-sharded over several GPUs (QRDM_B200_FORCE_MG=1 exercises that path on a 1-rank communicator) */
  const int mg = sh && (sh->nranks > 1 || getenv("QRDM_B200_FORCE_MG") != NULL);
  const int m_glob = sh ? sh->m_glob : m;
  const int minmn = m_glob < n ? m_glob : n;
  int stop_mode = 0;
  double eta = 0.0;
  if (ncols[0] == 1) { stop_mode = 1; eta = eps * n; }                    /* :531-534 */
  else if (ncols[0] == 2) { stop_mode = 2; eta = eps * sqrt((double)n); } /* :535-538 */
  else if (ncols[0] == 3) { stop_mode = 3; eta = thres[2]; }              /* :539-543 */

  int rc = ws_ensure(m, n);
  if (rc) return rc;

  qrdm_prob P;
  memset(&P, 0, sizeof(P));
  P.m = m; P.n = n; P.lda = lda; P.nb = nb;
  P.delta = thres[0]; P.tau_ = thres[1];
  P.a = d_a; P.jpvt = d_jpvt; P.tau = d_tau;
  P.vn1 = w->vn1; P.vn2 = w->vn2; P.ctrl = w->ctrl;
  P.gram_part = w->gram_part; P.gram = w->gram;
  P.panel_part = w->panel_part; P.panel_row = w->panel_row;
  P.vc = w->vc; P.ldv = w->ldv;
  P.wp = w->wp; P.wp_elems = w->wp_elems; P.w2 = w->w2; P.ldw = w->ldw;
  P.nrm_part = w->nrm_part; P.nrm_splits = w->nrm_splits; P.flag_list = w->flag_list;
  P.sm_count = w->sm_count;
  P.vec16 = (((size_t)d_a & 15) == 0 && (lda & 1) == 0) ? 1 : 0;
  P.row0 = sh ? sh->row0 : 0; P.m_glob = m_glob; P.nranks = sh ? sh->nranks : 1;
  P.w_reduced = mg; P.mg_buf = w->mg_buf; P.mg_cnt = w->mg_cnt;
  { const char *dbg = getenv("QRDM_B200_DEBUG"); P.debug = dbg ? atoi(dbg) : 0; }
  P.thres0 = 5e-14; /* src/dgeqr2.c:40 */
  P.inv_scale = 1.0;
  /* fixed columns (already moved to the front, d_jpvt holds the matching permutation): factored first, as they stand,
   * in blocks of <= nb columns through the same kernels — the reference's dgeqrf + dormqr, src/dgeqrdm_work.c:612-635 */
  const int nfxd = nfxd_in < minmn ? nfxd_in : minmn; /* na = min(m, nfxd), :613 */
  P.nfxd = nfxd;
  P.keep_jpvt = nfxd_in > 0;
  P.vc_prev = w->vc + (size_t)w->ldv * 64;
  P.upd_flag = w->upd_marks; P.upd_eager = w->upd_marks + w->cap_n;
  /* Deferred ("lazy") trailing update, single GPU: pass 2 of a block is postponed and fused into pass 1
   * of the next block (k_fused) while the trailing matrix has at least lazy_min^2 elements; the columns
   * the next selection touches are completed eagerly.  QRDM_B200_LAZY=0 disables it,
   * QRDM_B200_LAZY_MIN overrides the threshold (tests force it on tiny matrices). */
  /* measured on B200 (16384^2): the fused kernel saves ~0.5 ms x (size/16384)^2 per iteration against
   * k_vtc + k_rankk, the eager completion costs ~0.09 ms: break-even near a 7000 x 7000 trailing matrix */
  /* with the look-ahead (below) pass 2 of small trailing matrices disappears into the selection / panel window
   * altogether, which moves the break-even down: 8192^2 84.4 -> 81.7 ms, 16384^2 324.4 -> 322.1 ms with 2048 */
  int lazy_min = (getenv("QRDM_B200_SIDE") && atoi(getenv("QRDM_B200_SIDE")) == 0) ? 7168 : 2048;
  { const char *e = getenv("QRDM_B200_LAZY_MIN"); if (e) lazy_min = atoi(e); if (lazy_min < 1) lazy_min = 1; }
  /* nb > 64: selection at full width, factorisation in micro-panels of 64 columns (k_wide.cu); eager schedule only */
  const int wide = nb > QRDM_KMAX;
  if (wide) {
    if (mg) {
      fprintf(stderr, "qrdm_b200: nb = %d > %d is not supported in the row-sharded entry point\n", nb, QRDM_KMAX);
      return QRDM_ERR_UNSUPPORTED;
    }
    rc = ws_ensure_wide(n);
    if (rc) return rc;
    CU(qrdm_rt_memset(w->wide_marks, 0, sizeof(int) * (size_t)n, stream));
  }
  const int lazy_on = !mg && !wide && !(getenv("QRDM_B200_LAZY") && atoi(getenv("QRDM_B200_LAZY")) == 0);
  int pending = 0, pend_j = 0, stamp = 0;
  double *vcbuf[2] = {w->vc, w->vc + (size_t)w->ldv * 64};
  /* Look-ahead.  DM selects on the norms AFTER the whole block has been applied, so the next panel cannot
   * start before pass 1 over the full trailing matrix; what can leave the critical path is pass 2 of the pending block:
   * as soon as the stamps of the eager set are written (after k_select + the eager completion), the side stream applies
   * it to the last columns of the matrix (k_rankk, side mode: many short CTAs at the least stream priority) while this
   * stream (greatest priority) syncs the mailbox and runs the next Gram / pick / permutation / panel — the side CTAs
   * fill whatever the latency-bound kernels of the chain leave idle.  The next k_fused then makes pass 1 only on those
   * columns (qrdm_prob::pre_col0).  The share is sized from the idle SM-time of the window:
   *   QRDM_B200_SIDE_US      microseconds of (nearly) the whole chip before the panel starts            [default 100]
   *   QRDM_B200_SIDE_COL_US  microseconds per panel column, on the SMs the panel does not occupy        [default 3.3]
   *   QRDM_B200_SIDE_EFF     side rate per SM relative to k_rankk's 24 TFLOP/s on the whole chip        [default 0.6]
   * (measured with the grouped panel, 16384^2: 0.6 -> 295.3 ms, 0.8 -> 298.9, 1.1 -> 304.6, look-ahead off 296.3: a side
   * update that is still running when the next k_fused is due costs more than the share it took off that kernel)
   * QRDM_B200_SIDE=0 switches the look-ahead off, QRDM_B200_SIDE_PANEL=0 ends the side update before the panel. */
  int side_on = lazy_on, side_over_panel = 1, side_upc = 1;
  double side_us = 100.0, side_col_us = 3.3, side_eff = 0.6;
  { const char *e = getenv("QRDM_B200_SIDE"); if (e && atoi(e) == 0) side_on = 0; }
  { const char *e = getenv("QRDM_B200_SIDE_US"); if (e) side_us = atof(e); }
  { const char *e = getenv("QRDM_B200_SIDE_PANEL"); if (e) side_over_panel = atoi(e); }
  { const char *e = getenv("QRDM_B200_SIDE_COL_US"); if (e) side_col_us = atof(e); }
  { const char *e = getenv("QRDM_B200_SIDE_EFF"); if (e) side_eff = atof(e); }
  { const char *e = getenv("QRDM_B200_SIDE_UPC"); if (e && atoi(e) >= 1) side_upc = atoi(e); }
  { const char *e = getenv("QRDM_B200_VT_WB"); P.vt_wb = e ? atoi(e) : 0; } /* experiment: cost of a pass-1-only unit (of 7) */
  int pend_pre = 0; /* first column the side stream owns for the pending block (0: none) */
  int fused_timed = 0, fused_prev_k = 0; /* bookkeeping of the separately timed k_fused launches (stats.fused_flops) */
  double fused_prev_cols = 0.0;
  if (w->side_busy) { /* a previous call left early with a side launch in flight: it still uses the workspace */
    CU(qrdm_rt_stream_wait_event(stream, w->ev_side[1]));
    w->side_busy = 0;
  }

  memset(&g_stats, 0, sizeof(g_stats));
  g_ev_used = 0;
  const long long launches0 = qrdm_rt_launch_count();
  CU(qrdm_rt_event_record(w->ev[0], stream));
  CU(qrdm_rt_memset(w->ctrl, 0, sizeof(qrdm_ctrl), stream));
  CU(qrdm_rt_memset(w->vc, 0, sizeof(double) * (size_t)w->ldv * 64 * 2, stream));
  CU(qrdm_rt_memset(w->upd_marks, 0, sizeof(int) * (size_t)w->cap_n * 2, stream));

  if (!mg) {
    STAGE(QRDM_STAGE_NORM_INIT, qrdm_k_colnorm(&P, 0, stream));   /* :672-682 */
  } else { /* partial sums of squares -> all-reduce -> sqrt */
    int nsplit = 1;
    CU(qrdm_k_colnorm_part(&P, 0, &nsplit, stream));
    if (mg_allreduce(P.nrm_part, (size_t)nsplit * n, stream)) return QRDM_ERR_COMM;
    CU(qrdm_k_colnorm_fin(&P, 0, nsplit, stream));
  }
  STAGE(QRDM_STAGE_SELECT, qrdm_k_select(&P, stream));
  rc = read_mailbox(&P, stream);
  if (rc) return rc;
  double in_scale = 1.0; /* power of two the input was multiplied by (1: the normal case) */
  if (!(w->mailbox->maxnrm <= 0x1p300) || w->mailbox->maxnrm < 0x1p-300) {
    rc = prescale_input(&P, mg, &in_scale, stream);
    if (rc) return rc;
    if (in_scale != 1.0) { /* norms and selection again, now on representable squares */
      P.thres0 = 5e-14 * in_scale;
      P.inv_scale = 1.0 / in_scale;
      wb = NULL; /* finished columns are only final after the unscaling at the end */
      CU(qrdm_rt_memset(w->ctrl, 0, sizeof(qrdm_ctrl), stream));
      if (!mg) {
        CU(qrdm_k_colnorm(&P, 0, stream));
      } else {
        int nsplit = 1;
        CU(qrdm_k_colnorm_part(&P, 0, &nsplit, stream));
        if (mg_allreduce(P.nrm_part, (size_t)nsplit * n, stream)) return QRDM_ERR_COMM;
        CU(qrdm_k_colnorm_fin(&P, 0, nsplit, stream));
      }
      CU(qrdm_k_select(&P, stream));
      rc = read_mailbox(&P, stream);
      if (rc) return rc;
    }
  }
  const double eta_factor = eta;
  eta *= w->mailbox->maxnrm; /* :684 (with fixed columns: re-evaluated on the free columns once they are done) */
  g_stats.stage_bytes[QRDM_STAGE_NORM_INIT] = 8.0 * (double)m * (double)n;

  int info = 0, it = 0, j = 0, sweep = 0; /* it: DM iterations (index into ncols); sweep: all iterations (V buffer parity) */
  while (j < minmn) { /* :694 */
    const int cols = n - j;
    int jr = j - P.row0;
    jr = jr < 0 ? 0 : (jr > m ? m : jr); /* first active local row */
    int lazy = 0;
    /* the mailbox read at the end of the previous iteration already holds this iteration's candidate count */
    if (w->mailbox->nc > 1) g_stats.stage_bytes[QRDM_STAGE_GRAM] += 8.0 * (double)(m - jr) * (double)w->mailbox->nc;
    if (wide) {
      const int kmax_h = nb < n - j ? (nb < m - j ? nb : m - j) : (n - j < m - j ? n - j : m - j);
      P.vc = vcbuf[0];
      P.vc_prev = vcbuf[1];
      STAGE(QRDM_STAGE_GRAM, qrdm_k_gram_wide(&P, w->wide_part, w->wide_gram, m - jr, stream));
      STAGE(QRDM_STAGE_PICK, qrdm_k_pick_wide(&P, w->wide_gram, w->wide_marks, w->wide_swaps, stream));
      for (int t = 0; t * QRDM_KMAX < kmax_h; ++t) { /* micro-panels; those behind a DM stop are no-ops on the device */
        const int jt = j + t * QRDM_KMAX;
        CU(qrdm_k_micro_begin(&P, t, stream));
        STAGE(QRDM_STAGE_PANEL, qrdm_k_panel(&P, jt, stream));
        STAGE(QRDM_STAGE_VTC, qrdm_k_trailing(&P, jt, stream));
      }
      CU(qrdm_k_micro_end(&P, stream));
      if (j < nfxd && j + (nb < nfxd - j ? nb : nfxd - j) >= nfxd) {
        STAGE(QRDM_STAGE_NORM_UPDATE, qrdm_k_norm_recompute_all(&P, j, stream));
      } else {
        STAGE(QRDM_STAGE_NORM_UPDATE, qrdm_k_norm_update(&P, j, stream));
      }
    } else if (!mg) {
      P.vc = vcbuf[sweep & 1];       /* V of this block; the pending block's V sits in the other buffer */
      P.vc_prev = vcbuf[(sweep & 1) ^ 1];
      STAGE(QRDM_STAGE_GRAM, qrdm_k_gram(&P, 0, m - jr, stream));
      STAGE(QRDM_STAGE_PICK, qrdm_k_pick(&P, stream));
      STAGE(QRDM_STAGE_PERMUTE, qrdm_k_permute(&P, stream));
      if (w->side_busy && !side_over_panel) { /* the panel needs every SM: the side update ends here */
        CU(qrdm_rt_stream_wait_event(stream, w->ev_side[1]));
        w->side_busy = 0;
      }
      STAGE(QRDM_STAGE_PANEL, qrdm_k_panel(&P, j, stream));
      /* both dimensions must be large as well: with few trailing columns (tall-skinny, C4) the eager set of
       * <= 128 columns is a large share of the matrix and the deferred schedule buys nothing */
      const int lazy_dim = lazy_min < 1024 ? lazy_min : 1024;
      lazy = lazy_on && j >= nfxd && n - j - QRDM_KMAX >= lazy_dim && m - j - QRDM_KMAX >= lazy_dim &&
             (double)(n - j - QRDM_KMAX) * (double)(m - j - QRDM_KMAX) >= (double)lazy_min * (double)lazy_min;
      if (!pending && !lazy) {
        STAGE(QRDM_STAGE_VTC, qrdm_k_trailing(&P, j, stream));
      } else {
        int vt_stride = 0, vt_grid = 0;
        long long lb = qrdm_rt_launch_count();
        /* the dominant kernel gets its own event pair (stage VTV): k_fused, or k_vtc for the first deferred block */
        CU(stage_begin(QRDM_STAGE_VTV, stream));
        const int bn = pending ? 64 : 128; /* tile width of the partial-W slots */
        if (w->side_busy) { /* columns >= pend_pre of the pending block come from the side stream */
          CU(qrdm_rt_stream_wait_event(stream, w->ev_side[1]));
          w->side_busy = 0;
        }
        P.pre_col0 = pending ? pend_pre : 0; /* k_fused / k_tinv / k_wapply share one unit partition, which depends on it */
        fused_prev_cols = 0.0;
        if (pending) { /* phase-A columns of this k_fused: up to the side stream's share, minus the eager set (>= 128 columns) */
          const int cend = pend_pre > 0 ? pend_pre : n;
          fused_prev_cols = (double)(cend - j - 128 > 0 ? cend - j - 128 : 0);
          fused_prev_k = w->mailbox->last_k;
        }
        fused_timed = 1;
        if (pending) CU(qrdm_k_fused(&P, j, &vt_stride, &vt_grid, stream)); /* pass 2 of block it-1 + pass 1 of block it */
        else CU(qrdm_k_vtc_only(&P, j, &vt_stride, &vt_grid, stream));
        CU(stage_end(QRDM_STAGE_VTV, lb, stream));
        lb = qrdm_rt_launch_count();
        CU(stage_begin(QRDM_STAGE_VTC, stream));
        pending = 0;
        /* lazy: k_wapply also finishes the k new R rows; everything below them waits for the next k_fused */
        if (vt_stride > 0) CU(qrdm_k_w2(&P, j, vt_grid, vt_stride, bn | (lazy ? 1 : 0), stream));
        P.pre_col0 = 0;
        pend_pre = 0;
        if (!lazy) CU(qrdm_k_rankk(&P, j, stream));
        CU(stage_end(QRDM_STAGE_VTC, lb, stream));
      }
      if (lazy) {
        P.stamp = ++stamp;
        STAGE(QRDM_STAGE_NORM_UPDATE, qrdm_k_norm_update_lazy(&P, j, stream));
      } else if (j < nfxd && j + (nb < nfxd - j ? nb : nfxd - j) >= nfxd) {
        /* last block of fixed columns: the free columns' norms from scratch, as :672-682 computes them */
        STAGE(QRDM_STAGE_NORM_UPDATE, qrdm_k_norm_recompute_all(&P, j, stream));
      } else {
        STAGE(QRDM_STAGE_NORM_UPDATE, qrdm_k_norm_update(&P, j, stream));
      }
    } else {
      /* every row sum is computed locally, all-reduced over NVLink, and consumed by a replicated
       * kernel: all ranks hold identical vn1/jpvt/ctrl and take identical decisions */
      const int kmax_h = nb < n - j ? (nb < m_glob - j ? nb : m_glob - j) : (n - j < m_glob - j ? n - j : m_glob - j);
      int vt_stride = 0, vt_grid = 0, nsplit = 1;
      long long lbm = qrdm_rt_launch_count();
      CU(stage_begin(QRDM_STAGE_GRAM, stream));
      CU(qrdm_k_gram_part(&P, m - jr > 0 ? m - jr : 1, stream));
      if (mg_allreduce(P.gram, 4096, stream)) return QRDM_ERR_COMM;
      CU(stage_end(QRDM_STAGE_GRAM, lbm, stream));
      STAGE(QRDM_STAGE_PICK, qrdm_k_pick(&P, stream));
      STAGE(QRDM_STAGE_PERMUTE, qrdm_k_permute(&P, stream));
      lbm = qrdm_rt_launch_count();
      CU(stage_begin(QRDM_STAGE_PANEL, stream));
      if (qrdm_rt_peer_available() == P.nranks && !getenv("QRDM_B200_MG_LEGACY")) {
        /* Peer memory open: every sharded panel runs blocked, 8-column sub-panels factored by ONE persistent kernel
         * each (k_panel_tall<true>) that exchanges the per-column vector [||x||^2, x'C_sub | pivot row] with the other
         * GPUs as LL packets over NVLink from inside the kernel — no launch and no collective call per column
         * (src/dgeqr2.c:148-189 is one tight loop upstream).  Each sub-panel's block reflector then updates the rest of
         * the panel: V'[V | C_p] partials, their sum over the GPUs inside the W2 kernel (LL packets again), C_p += V W2. */
        for (int sb = 0; sb < kmax_h; sb += QRDM_TALL_B) {
          qrdm_prob Ps = P;
          Ps.sub = sb + 1;
          int jrs = j + sb - P.row0;
          jrs = jrs < 0 ? 0 : (jrs > m ? m : jrs);
          CU(qrdm_k_panel_tall_mg(&Ps, j, stream));
          if (sb + QRDM_TALL_B < kmax_h) CU(qrdm_k_skinny_update_mg(&Ps, m - jrs > 0 ? m - jrs : 1, stream));
        }
      } else if ((m_glob - j) / P.nranks <= 16384) {
        /* legacy transport (NCCL only): one kernel + one 128-double all-reduce per panel column */
        CU(qrdm_k_panel_mg_init(&P, j, stream));
        if (qrdm_rt_allreduce(P.mg_buf, 128, stream)) return QRDM_ERR_COMM;
        for (int i = 0; i < kmax_h; ++i) {
          CU(qrdm_k_panel_mg_step(&P, j, i, stream));
          if (i + 1 < kmax_h && qrdm_rt_allreduce(P.mg_buf + ((i + 1) & 1) * 128, 128, stream)) return QRDM_ERR_COMM;
        }
        CU(qrdm_k_panel_mg_finish(&P, j, stream));
      } else {
        /* tall local slabs: blocked panel — 8-column sub-panels (per-column kernel + all-reduce on 8
         * columns only), each followed by the skinny update of the rest of the panel, whose 8 x 64
         * product V'[V | C_p] is all-reduced (512 doubles).  The choice depends on global sizes
         * only, so every rank issues the same sequence of collectives. */
        for (int sb = 0; sb < kmax_h; sb += QRDM_TALL_B) {
          qrdm_prob Ps = P;
          Ps.sub = sb + 1;
          const int wsub = kmax_h - sb < QRDM_TALL_B ? kmax_h - sb : QRDM_TALL_B;
          int jrs = j + sb - P.row0;
          jrs = jrs < 0 ? 0 : (jrs > m ? m : jrs);
          CU(qrdm_k_panel_mg_init(&Ps, j, stream));
          if (qrdm_rt_allreduce(P.mg_buf, 128, stream)) return QRDM_ERR_COMM;
          for (int i = 0; i < wsub; ++i) {
            CU(qrdm_k_panel_mg_step(&Ps, j, i, stream));
            if (i + 1 < wsub && qrdm_rt_allreduce(P.mg_buf + ((i + 1) & 1) * 128, 128, stream)) return QRDM_ERR_COMM;
          }
          CU(qrdm_k_panel_mg_finish(&Ps, j, stream));
          if (sb + QRDM_TALL_B < kmax_h) {
            CU(qrdm_k_skinny_part(&Ps, m - jrs > 0 ? m - jrs : 1, stream));
            if (qrdm_rt_allreduce(P.gram_part, 512, stream)) return QRDM_ERR_COMM;
            CU(qrdm_k_skinny_finish(&Ps, m - jrs > 0 ? m - jrs : 1, stream));
          }
        }
      }
      CU(stage_end(QRDM_STAGE_PANEL, lbm, stream));
      lbm = qrdm_rt_launch_count();
      CU(stage_begin(QRDM_STAGE_VTC, stream));
      CU(qrdm_k_vtc_only(&P, j, &vt_stride, &vt_grid, stream));
      if (vt_stride > 0) {
        CU(qrdm_k_wreduce(&P, j, vt_grid, vt_stride, stream));
        if (mg_allreduce(P.wp, (size_t)64 * vt_stride, stream)) return QRDM_ERR_COMM;
        CU(qrdm_k_trailing_finish(&P, j, vt_grid, vt_stride, stream));
      }
      CU(stage_end(QRDM_STAGE_VTC, lbm, stream));
      lbm = qrdm_rt_launch_count();
      CU(stage_begin(QRDM_STAGE_NORM_UPDATE, stream));
      if (n - j - 1 > 0) {
        CU(qrdm_k_norm_dpart(&P, j, stream));
        if (mg_allreduce(P.nrm_part, (size_t)n, stream)) return QRDM_ERR_COMM;
        CU(qrdm_k_norm_apply(&P, j, stream));
        CU(qrdm_k_colnorm_part(&P, 2, &nsplit, stream));
        if (mg_allreduce(P.nrm_part, (size_t)nsplit * n, stream)) return QRDM_ERR_COMM;
        CU(qrdm_k_colnorm_fin(&P, 2, nsplit, stream));
      }
      CU(stage_end(QRDM_STAGE_NORM_UPDATE, lbm, stream));
    }
    STAGE(QRDM_STAGE_SELECT, qrdm_k_select(&P, stream)); /* next iteration's prologue + max norm */
    if (lazy) { /* complete the update on everything the next Gram / pick / permutation / panel touches */
      STAGE(QRDM_STAGE_VTC, qrdm_k_colupd(&P, 0, j, stream));
      pending = 1;
      pend_j = j;
      pend_pre = 0;
      if (side_on) {
        /* columns for the side stream: the idle SM-time of the window at k_rankk's rate (24 TFLOP/s on the whole chip =
         * 187500 elements per us at k = 64, i.e. ~1267 per SM and us); never the leading 128 columns behind the block
         * (eager set, next panel) */
        const double rows_s = (double)(m - j - 1);
        const int kmax_next = nb < n - j - 1 ? nb : n - j - 1;
        double sm_us = side_us * (double)(w->sm_count > 4 ? w->sm_count - 4 : 1);
        if (side_over_panel) {
          const int free_sms = w->sm_count - qrdm_k_panel_ctas(&P, m - j - 1);
          if (free_sms > 0) sm_us += side_col_us * (double)kmax_next * (double)free_sms;
        }
        long long nside = rows_s > 0 ? (long long)(sm_us * side_eff * (187500.0 / 148.0) / rows_s) : 0;
        nside = nside / 32 * 32;
        int c0s = n - (nside > n ? n : (int)nside);
        if (c0s < j + 128) c0s = j + 128;
        if (n - c0s >= 64) {
          qrdm_prob Ps = P;
          Ps.side_col0 = c0s;
          CU(qrdm_rt_event_record(w->ev_side[0], stream));
          CU(qrdm_rt_stream_wait_event(w->side_stream, w->ev_side[0]));
          {
            long long lbs = qrdm_rt_launch_count();
            CU(stage_begin(QRDM_STAGE_RANKK, w->side_stream));
            CU(qrdm_k_side(&Ps, j, side_upc, w->side_stream));
            CU(stage_end(QRDM_STAGE_RANKK, lbs, w->side_stream));
          }
          CU(qrdm_rt_event_record(w->ev_side[1], w->side_stream));
          w->side_busy = 1;
          pend_pre = c0s;
          ++g_stats.side_launches;
        }
      }
    }
    {
      long long lb = qrdm_rt_launch_count();
      CU(stage_begin(QRDM_STAGE_SYNC, stream));
      rc = read_mailbox(&P, stream);
      if (rc) return rc;
      CU(stage_end(QRDM_STAGE_SYNC, lb, stream));
    }
    const qrdm_ctrl *mb = w->mailbox;
    const int k = mb->last_k;
    const int was_fixed = j < nfxd;
    ++sweep;
    if (!was_fixed) ncols[it++] = k; /* :740 (the reference counts DM iterations only: it starts after the fixed part) */
    g_stats.panel_cols += k;
    if (mb->err != 0) { info = mb->err; break; }
    if (k <= 0 || mb->j != j + k) {
      fprintf(stderr, "qrdm_b200: internal error, block size %d at j=%d (device j=%d)\n", k, j, mb->j);
      info = QRDM_ERR_INTERNAL;
      break;
    }
    g_stats.trailing_flops += 4.0 * (double)(m - jr) * (double)(cols - k) * (double)k;
    if (fused_timed) { /* FLOPs inside the separately timed k_fused / k_vtc launch: pass 1 of this block + pass 2 of the pending one */
      g_stats.fused_flops += 2.0 * (double)(m - jr) * (double)(cols - k) * (double)k
                             + 2.0 * (double)(m - jr) * fused_prev_cols * (double)fused_prev_k;
      ++g_stats.fused_launches;
      fused_timed = 0;
    }
    if (pend_pre > 0 && pending) /* this block's pass 2 on the columns >= pend_pre went to the side stream (stamped columns aside) */
      g_stats.side_flops += 2.0 * (double)(m - jr - k) * (double)(n - pend_pre) * (double)k;
    g_stats.stage_bytes[QRDM_STAGE_PANEL] += 16.0 * (double)(m - jr) * (double)k; /* >= 16 m_r k: k <= fjb columns read + written once */
    g_stats.stage_bytes[QRDM_STAGE_NORM_UPDATE] += 8.0 * (double)k * (double)(cols - k);
    g_stats.stage_bytes[QRDM_STAGE_VTC] += (lazy ? 16.0 : 24.0) * (double)(m - jr) * (double)(cols - k);
    j += k;
    if (wb && wb->io) {
      /* pageable host buffer: the same streaming through the pinned ring + drain thread of hostio.c; the ring
       * may be full, then the columns wait for a later push or for the final download */
      const int taken = qrdm_hostio_wb_push(wb->io, wb->done_cols, j);
      if (taken < 0) return QRDM_ERR_CUDA;
      wb->done_cols += taken;
    } else if (wb) {
      /* columns [done, j) are final (later iterations only touch columns >= j, and the mailbox
       * sync above means every kernel of this iteration has finished): stream them to the host
       * while the next iterations compute */
      CU(qrdm_rt_d2h_2d(wb->h_a + (size_t)wb->done_cols * wb->h_lda, sizeof(double) * wb->h_lda,
                        d_a + (size_t)wb->done_cols * lda, sizeof(double) * lda, sizeof(double) * m,
                        (size_t)(j - wb->done_cols), wb->copy_stream));
      wb->done_cols = j;
    }
    if (was_fixed) {
      if (j >= nfxd) eta = eta_factor * mb->maxnrm; /* :684 — max norm of the free columns below the fixed block */
      continue;
    }
    if (stop_mode && mb->maxnrm * sqrt((double)(cols - k)) <= eta) break; /* :782-785 */
  }
  if (w->side_busy) {
    CU(qrdm_rt_stream_wait_event(stream, w->ev_side[1]));
    w->side_busy = 0;
  }
  if (pending) { /* early exit (stop rule / error) with a block still deferred: finish it */
    P.vc_prev = vcbuf[(sweep - 1) & 1];
    P.pre_col0 = pend_pre; /* the side stream's columns are done */
    STAGE(QRDM_STAGE_VTC, qrdm_k_flush(&P, pend_j, stream));
    P.pre_col0 = 0;
  }
  if (in_scale != 1.0) CU(qrdm_k_scale(&P, 1.0 / in_scale, 1, j, stream)); /* R back to the caller's scale */
  CU(qrdm_rt_event_record(w->ev[1], stream));
  if (g_profile) CU(qrdm_rt_d2h(w->mailbox, w->ctrl, sizeof(qrdm_ctrl), stream)); /* device-side statistics (after the timed region) */
  CU(qrdm_rt_event_sync(w->ev[1]));
  if (g_profile) {
    CU(qrdm_rt_sync(stream));
    g_stats.stage_bytes[QRDM_STAGE_PERMUTE] = 16.0 * (double)m * (double)w->mailbox->stat_perm_cols;
  }
  g_stats.ms_total = qrdm_rt_event_ms(w->ev[0], w->ev[1]);
  for (int e = 0; e + 1 < g_ev_used; e += 2) /* light profile: sum the pooled event pairs */
    g_stats.ms_stage[g_ev_stage_of[e / 2]] += qrdm_rt_event_ms(g_ev_pool[e], g_ev_pool[e + 1]);
  g_stats.iterations = it;
  g_stats.rank = j;
  g_stats.launches = qrdm_rt_launch_count() - launches0;
  if (info == -8) fprintf(stderr, "LAPACK dlarft failed, info =%d \n", info);        /* :755-758 */
  else if (info == -13) fprintf(stderr, "LAPACK dlarfb failed, info =%d \n", info); /* :768-771 */
  return info;
}

int dgeqrdm_dev(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, int *ncols,
                const double *thres, int nb, void *stream) {
  int rc = check_args(QRDM_COL_MAJOR, m, n, lda, thres, nb);
  if (rc) return rc;
  API_BODY(factor_device(m, n, d_a, lda, d_jpvt, d_tau, ncols, thres, nb, stream, NULL, NULL, 0));
}

static int ensure_staging(int m, int n, int ldd) {
  qrdm_workspace *w = &g_ws;
  const int minmn = m < n ? m : n;
  const size_t a_bytes = sizeof(double) * (size_t)ldd * n;
  if (a_bytes > w->cap_a_bytes) {
    if (w->d_a) qrdm_rt_free(w->d_a);
    w->d_a = NULL;
    w->cap_a_bytes = 0;
    CU(qrdm_rt_malloc((void **)&w->d_a, a_bytes));
    w->cap_a_bytes = a_bytes;
  }
  if ((size_t)minmn > w->cap_tau) {
    if (w->d_tau) qrdm_rt_free(w->d_tau);
    w->d_tau = NULL;
    w->cap_tau = 0;
    CU(qrdm_rt_malloc((void **)&w->d_tau, sizeof(double) * minmn));
    w->cap_tau = minmn;
  }
  if ((size_t)n > w->cap_jpvt) {
    if (w->d_jpvt) qrdm_rt_free(w->d_jpvt);
    w->d_jpvt = NULL;
    w->cap_jpvt = 0;
    CU(qrdm_rt_malloc((void **)&w->d_jpvt, sizeof(int) * n));
    w->cap_jpvt = n;
  }
  return 0;
}

/* Host-pointer entry point.  Three transfer modes, chosen from the caller's buffer:
 *   pinned host memory      one async 2-D H2D; finished columns stream back D2H during the factorisation;
 *   pageable (NumPy) memory multi-threaded pinned bounce pipeline both ways + the same streamed write-back through a
 *                           pinned ring (hostio.c) — the reference's caller passes pageable arrays (QRDM_wrapper.c:155-159);
 *   QRDM_B200_NO_BOUNCE=1   plain cudaMemcpy2D from pageable memory (the round-1 behaviour, kept for A/B timing). */
static int dgeqrdm_work_locked(int m, int n, double *a, int lda, int *jpvt, double *tau, int *ncols, double *thres, int nb,
                               int nfxd, const int *src_col, const int *jpvt0) {
  qrdm_workspace *w = &g_ws;
  int rc = init_impl(-1);
  if (rc) return rc;
  void *stream = w->compute_stream;
  const int minmn = m < n ? m : n;
  const int ldd = (m + 1) & ~1; /* even leading dimension on the device: 16-byte aligned columns */
  const size_t a_bytes = sizeof(double) * (size_t)ldd * n;
  rc = ensure_staging(m, n, ldd);
  if (rc) return rc;
  const int pinned = qrdm_rt_is_pinned(a);
  const int bounce = !pinned && !getenv("QRDM_B200_NO_BOUNCE") && (size_t)m * n >= (size_t)1 << 16;
  if (bounce && !w->io && qrdm_hostio_create(&w->io, w->device)) return QRDM_ERR_CUDA;
  int info = QRDM_ERR_CUDA;
#define CUX(call)                                                                             \
  do {                                                                                        \
    int e__ = (call);                                                                         \
    if (e__ != 0) {                                                                           \
      fprintf(stderr, "qrdm_b200: CUDA error %d (%s) at %s:%d\n", e__, qrdm_rt_errstr(e__), \
              __FILE__, __LINE__);                                                            \
      info = QRDM_ERR_CUDA;                                                                   \
      goto fail;                                                                              \
    }                                                                                         \
  } while (0)
  CUX(qrdm_rt_event_record(w->ev[2], stream));
  if (ldd != m) CUX(qrdm_rt_memset(w->d_a, 0, a_bytes, stream));
  if (bounce) {
    CUX(qrdm_rt_sync(stream)); /* the memset of the padding row must not race with the workers' streams */
    if (qrdm_hostio_upload(w->io, w->d_a, ldd, a, lda, m, n)) { info = QRDM_ERR_CUDA; goto fail; }
  } else {
    CUX(qrdm_rt_h2d_2d(w->d_a, sizeof(double) * ldd, a, sizeof(double) * lda, sizeof(double) * m, n, stream));
  }
  CUX(qrdm_rt_h2d(w->d_tau, tau, sizeof(double) * minmn, stream)); /* entries >= rank stay as given */
  if (nfxd > 0) {
    /* fixed columns: the reference swaps them to the front on the host (src/dgeqrdm_work.c:592-607); here device column
     * pos simply receives the caller's column src_col[pos], and d_jpvt starts as the permutation those swaps produce */
    for (int pos = 0; pos < n; ++pos)
      if (src_col[pos] != pos)
        CUX(qrdm_rt_h2d(w->d_a + (size_t)pos * ldd, a + (size_t)src_col[pos] * lda, sizeof(double) * m, stream));
    CUX(qrdm_rt_h2d(w->d_jpvt, jpvt0, sizeof(int) * n, stream));
  }
  CUX(qrdm_rt_event_record(w->ev[3], stream));
  {
    /* overlap the D2H of finished columns with the rest of the factorisation */
    qrdm_writeback wb = {a, lda, w->copy_stream, 0, NULL};
    int overlap = (pinned || bounce) && !getenv("QRDM_B200_NO_OVERLAP");
    if (overlap && bounce) {
      const int r = qrdm_hostio_wb_begin(w->io, a, lda, w->d_a, ldd, m, w->copy_stream);
      if (r < 0) { info = QRDM_ERR_CUDA; goto fail; }
      if (r == 0) wb.io = w->io; else overlap = 0;
    }
    info = factor_device(m, n, w->d_a, ldd, w->d_jpvt, w->d_tau, ncols, thres, nb, stream, overlap ? &wb : NULL, NULL, nfxd);
    if (wb.io && qrdm_hostio_wb_end(wb.io) && info > QRDM_ERR_CUDA) info = QRDM_ERR_CUDA;
    if (info <= QRDM_ERR_CUDA) goto fail;
    const double ms_h2d = qrdm_rt_event_ms(w->ev[2], w->ev[3]);
    CUX(qrdm_rt_event_record(w->ev[2], stream));
    const int done = overlap ? wb.done_cols : 0; /* what is left: the unreduced tail (early stop / error), a full ring */
    if (done < n) {
      if (bounce) {
        if (qrdm_hostio_download(w->io, a, lda, w->d_a, ldd, m, done, n)) { info = QRDM_ERR_CUDA; goto fail; }
      } else {
        CUX(qrdm_rt_d2h_2d(a + (size_t)done * lda, sizeof(double) * lda, w->d_a + (size_t)done * ldd,
                           sizeof(double) * ldd, sizeof(double) * m, (size_t)(n - done), stream));
      }
    }
    if (overlap) CUX(qrdm_rt_sync(w->copy_stream));
    CUX(qrdm_rt_d2h(jpvt, w->d_jpvt, sizeof(int) * n, stream));
    CUX(qrdm_rt_d2h(tau, w->d_tau, sizeof(double) * minmn, stream));
    CUX(qrdm_rt_event_record(w->ev[3], stream));
    CUX(qrdm_rt_sync(stream));
    g_stats.ms_h2d = ms_h2d;
    g_stats.ms_d2h = qrdm_rt_event_ms(w->ev[2], w->ev[3]);
  }
  return info;
fail:
  /* nothing may still be writing into the caller's buffers (or reading ours) once we return */
  qrdm_rt_sync(w->copy_stream);
  qrdm_rt_sync(stream);
  return info;
#undef CUX
}

int dgeqrdm_work(int matrix_layout, int m, int n, double *a, int lda, int *jpvt, double *tau,
                 int *ncols, double *thres, int nb) {
  int rc = check_args(matrix_layout, m, n, lda, thres, nb);
  if (rc) return rc;
  /* Fixed columns, jpvt[j] != 0 on entry (LAPACK dgeqp3 convention): the reference moves them to the front with the
   * swap sequence of src/dgeqrdm_work.c:592-607 — replayed here on index arrays — factors them without pivoting and
   * runs DM on the rest.  (Its own continuation mis-indexes the auxiliary arrays for nfxd > 0; what is
   * implemented is what :592-635 sets out to do.)  ncols counts DM iterations only, as upstream's `it` does. */
  int nfxd = 0, any = 0;
  for (int c = 0; c < n; ++c) any |= (jpvt[c] != 0);
  if (!any) API_BODY(dgeqrdm_work_locked(m, n, a, lda, jpvt, tau, ncols, thres, nb, 0, NULL, NULL));
  int *src_col = (int *)malloc(sizeof(int) * (size_t)n * 2);
  if (!src_col) return QRDM_ERR_CUDA;
  int *jp = src_col + n;
  for (int c = 0; c < n; ++c) { src_col[c] = c; jp[c] = jpvt[c]; }
  for (int c = 0; c < n; ++c) { /* 0-based replay of :594-607 */
    if (jp[c] != 0) {
      if (c != nfxd) {
        const int t = src_col[c]; src_col[c] = src_col[nfxd]; src_col[nfxd] = t;
        jp[c] = jp[nfxd];
        jp[nfxd] = c + 1;
      } else {
        jp[c] = c + 1;
      }
      ++nfxd;
    } else {
      jp[c] = c + 1;
    }
  }
  api_lock();
  rc = dgeqrdm_work_locked(m, n, a, lda, jpvt, tau, ncols, thres, nb, nfxd, src_col, jp);
  api_unlock();
  free(src_col);
  return rc;
}

int dgeqrdm(int matrix_layout, int m, int n, double *a, int lda, int *jpvt, double *tau, int *ncols,
            double *thres, int nb) {
  return dgeqrdm_work(matrix_layout, m, n, a, lda, jpvt, tau, ncols, thres, nb);
}

/* ---- Q application: C <- Q C or Q' C with the reflectors of a factored matrix, in blocks of 64
 * through the trailing-update kernels (K6).  Replaces the LAPACKE_dormqr('L', .) call of the reference
 * wrapper's DORMQR (QRDM_wrapper.c:104-126) and makes auxil.checkQR (auxil.py:20-105) possible at sizes whose
 * m x m Q would never fit on the host.  Any partition of the k reflectors into blocks is valid: T is rebuilt per
 * block from V'V and tau (k_tinv), so this does not depend on the block sizes of the factorisation. ---- */
static int dormqr_dev_locked(char trans, int m, int n, int k, const double *d_a, int lda, const double *d_tau, double *d_c,
                             int ldc, void *stream) {
  const int transN = (trans == 'N' || trans == 'n');
  if (!transN && trans != 'T' && trans != 't') return bad_argument(1);
  if (m <= 0) return bad_argument(2);
  if (n <= 0) return bad_argument(3);
  if (k < 0 || k > m) return bad_argument(4);
  if (lda < m) return bad_argument(6);
  if (ldc < m) return bad_argument(9);
  int rc = ws_ensure(m, n + QRDM_KMAX);
  if (rc) return rc;
  qrdm_workspace *w = &g_ws;
  qrdm_prob P;
  memset(&P, 0, sizeof(P));
  P.m = m; P.lda = ldc; P.nb = QRDM_KMAX;
  P.tau = (double *)d_tau;
  P.vn1 = w->vn1; P.vn2 = w->vn2; P.ctrl = w->ctrl;
  P.gram_part = w->gram_part; P.gram = w->gram;
  P.vc = w->vc; P.vc_prev = w->vc; P.ldv = w->ldv;
  P.wp = w->wp; P.wp_elems = w->wp_elems; P.w2 = w->w2; P.ldw = w->ldw;
  P.nrm_part = w->nrm_part; P.nrm_splits = w->nrm_splits; P.flag_list = w->flag_list;
  P.upd_flag = w->upd_marks; P.upd_eager = w->upd_marks + w->cap_n;
  P.sm_count = w->sm_count;
  P.vec16 = (((size_t)d_c & 15) == 0 && (ldc & 1) == 0) ? 1 : 0;
  P.m_glob = m; P.nranks = 1;
  memset(&g_stats, 0, sizeof(g_stats));
  const long long launches0 = qrdm_rt_launch_count();
  CU(qrdm_rt_event_record(w->ev[0], stream));
  CU(qrdm_rt_memset(w->ctrl, 0, sizeof(qrdm_ctrl), stream));
  CU(qrdm_rt_memset(w->vc, 0, sizeof(double) * (size_t)w->ldv * 64, stream));
  const int nblk = (k + QRDM_KMAX - 1) / QRDM_KMAX;
  for (int bi = 0; bi < nblk; ++bi) {
    const int b = transN ? nblk - 1 - bi : bi; /* Q = H_0 H_1 ...: last block first; Q': first block first */
    const int j0 = b * QRDM_KMAX, kb = k - j0 < QRDM_KMAX ? k - j0 : QRDM_KMAX;
    int stride = 0, grid = 0;
    /* the kernels update the columns right of the block, A[:, j0+kb : P.n): make that range be C */
    P.n = j0 + kb + n;
    P.a = (double *)((size_t)d_c - sizeof(double) * (size_t)(j0 + kb) * (size_t)ldc);
    CU(qrdm_k_vc_build(&P, d_a, lda, j0, kb, stream));
    CU(qrdm_k_vtc_only(&P, j0, &stride, &grid, stream));
    if (stride > 0) {
      CU(qrdm_k_w2(&P, j0, grid, stride, 128 | (transN ? 2 : 0), stream));
      CU(qrdm_k_rankk(&P, j0, stream));
    }
    g_stats.trailing_flops += 4.0 * (double)(m - j0) * (double)n * (double)kb;
  }
  CU(qrdm_rt_event_record(w->ev[1], stream));
  rc = read_mailbox(&P, stream);
  if (rc) return rc;
  g_stats.ms_total = qrdm_rt_event_ms(w->ev[0], w->ev[1]);
  g_stats.iterations = nblk;
  g_stats.rank = k;
  g_stats.launches = qrdm_rt_launch_count() - launches0;
  return w->mailbox->err; /* 0, or -13 if a NaN went through (the screen of k_wapply) */
}

int qrdm_b200_dormqr_dev(char trans, int m, int n, int k, const double *d_a, int lda, const double *d_tau, double *d_c,
                         int ldc, void *stream) {
  API_BODY(dormqr_dev_locked(trans, m, n, k, d_a, lda, d_tau, d_c, ldc, stream));
}

/* ---- QR with classical column pivoting on the device: the comparator the reference exports as dgeqp3
 * (src/dgeqp3.c:39-93 -> LAPACKE_dgeqp3_work) and its wrapper calls as QP3 (QRDM_wrapper.c:15-41), here as LAPACK's
 * blocked algorithm (dlaqps) on the GPU (k_qp3.cu).  All columns are free: jpvt is output only (1-based). ---- */
static int dgeqp3_dev_locked(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, void *stream) {
  if (m <= 0) return bad_argument(2);
  if (n <= 0) return bad_argument(3);
  if (lda < m) return bad_argument(5);
  int rc = ws_ensure(m, n);
  if (rc) return rc;
  qrdm_workspace *w = &g_ws;
  qrdm_prob P;
  memset(&P, 0, sizeof(P));
  P.m = m; P.n = n; P.lda = lda; P.nb = QRDM_KMAX;
  P.a = d_a; P.jpvt = d_jpvt; P.tau = d_tau;
  P.vn1 = w->vn1; P.vn2 = w->vn2; P.ctrl = w->ctrl;
  P.gram_part = w->gram_part; P.gram = w->gram;
  P.vc = w->vc; P.vc_prev = w->vc; P.ldv = w->ldv;
  P.wp = w->wp; P.wp_elems = w->wp_elems; P.w2 = w->w2; P.ldw = w->ldw;
  P.nrm_part = w->nrm_part; P.nrm_splits = w->nrm_splits; P.flag_list = w->flag_list;
  P.upd_flag = w->upd_marks; P.upd_eager = w->upd_marks + w->cap_n;
  P.sm_count = w->sm_count;
  P.vec16 = (((size_t)d_a & 15) == 0 && (lda & 1) == 0) ? 1 : 0;
  P.m_glob = m; P.nranks = 1; P.inv_scale = 1.0; P.thres0 = 5e-14;
  memset(&g_stats, 0, sizeof(g_stats));
  const long long launches0 = qrdm_rt_launch_count();
  CU(qrdm_rt_event_record(w->ev[0], stream));
  CU(qrdm_rt_memset(w->vc, 0, sizeof(double) * (size_t)w->ldv * 64, stream));
  CU(qrdm_rt_memset(w->upd_marks, 0, sizeof(int) * (size_t)w->cap_n * 2, stream));
  CU(qrdm_qp3_dev(&P, w->mailbox, stream));
  CU(qrdm_rt_event_record(w->ev[1], stream));
  CU(qrdm_rt_event_sync(w->ev[1]));
  g_stats.ms_total = qrdm_rt_event_ms(w->ev[0], w->ev[1]);
  g_stats.rank = m < n ? m : n;
  g_stats.launches = qrdm_rt_launch_count() - launches0;
  return 0;
}
int qrdm_b200_dgeqp3_dev(int m, int n, double *d_a, int lda, int *d_jpvt, double *d_tau, void *stream) {
  API_BODY(dgeqp3_dev_locked(m, n, d_a, lda, d_jpvt, d_tau, stream));
}
static int dgeqp3_locked(int m, int n, double *a, int lda, int *jpvt, double *tau) {
  if (m <= 0) return bad_argument(2);
  if (n <= 0) return bad_argument(3);
  if (lda < (m > 1 ? m : 1)) return bad_argument(5);
  for (int c = 0; c < n; ++c)
    if (jpvt[c] != 0) {
      fprintf(stderr, "qrdm_b200_dgeqp3: fixed columns (jpvt[%d] != 0 on entry) are not supported\n", c);
      return QRDM_ERR_UNSUPPORTED;
    }
  int rc = init_impl(-1);
  if (rc) return rc;
  void *stream = g_ws.compute_stream;
  const int ldd = (m + 1) & ~1, minmn = m < n ? m : n;
  double *d_a = NULL, *d_tau = NULL;
  int *d_jpvt = NULL;
  int info = QRDM_ERR_CUDA;
#define CUX(call)                                                                             \
  do {                                                                                        \
    int e__ = (call);                                                                         \
    if (e__ != 0) {                                                                           \
      fprintf(stderr, "qrdm_b200: CUDA error %d (%s) at %s:%d\n", e__, qrdm_rt_errstr(e__), \
              __FILE__, __LINE__);                                                            \
      info = QRDM_ERR_CUDA;                                                                   \
      goto done;                                                                              \
    }                                                                                         \
  } while (0)
  CUX(qrdm_rt_malloc((void **)&d_a, sizeof(double) * (size_t)ldd * n));
  CUX(qrdm_rt_malloc((void **)&d_tau, sizeof(double) * minmn));
  CUX(qrdm_rt_malloc((void **)&d_jpvt, sizeof(int) * n));
  CUX(qrdm_rt_memset(d_a, 0, sizeof(double) * (size_t)ldd * n, stream));
  CUX(qrdm_rt_h2d_2d(d_a, sizeof(double) * ldd, a, sizeof(double) * lda, sizeof(double) * m, n, stream));
  info = dgeqp3_dev_locked(m, n, d_a, ldd, d_jpvt, d_tau, stream);
  if (info == 0) {
    CUX(qrdm_rt_d2h_2d(a, sizeof(double) * lda, d_a, sizeof(double) * ldd, sizeof(double) * m, n, stream));
    CUX(qrdm_rt_d2h(tau, d_tau, sizeof(double) * minmn, stream));
    CUX(qrdm_rt_d2h(jpvt, d_jpvt, sizeof(int) * n, stream));
    info = 0;
  }
done:
  qrdm_rt_sync(stream);
  if (d_a) qrdm_rt_free(d_a);
  if (d_tau) qrdm_rt_free(d_tau);
  if (d_jpvt) qrdm_rt_free(d_jpvt);
  return info;
#undef CUX
}
int qrdm_b200_dgeqp3(int m, int n, double *a, int lda, int *jpvt, double *tau) { API_BODY(dgeqp3_locked(m, n, a, lda, jpvt, tau)); }

/* Host-pointer variant: A (m x k reflector columns, as returned by dgeqrdm/dgeqrf), tau and C in host memory. */
static int dormqr_locked(char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc) {
  int rc = init_impl(-1);
  if (rc) return rc;
  void *stream = g_ws.compute_stream;
  const int ldd = (m + 1) & ~1;
  double *d_a = NULL, *d_c = NULL, *d_tau = NULL;
  const int kc = k > 0 ? k : 1;
  int info = QRDM_ERR_CUDA;
#define CUX(call)                                                                             \
  do {                                                                                        \
    int e__ = (call);                                                                         \
    if (e__ != 0) {                                                                           \
      fprintf(stderr, "qrdm_b200: CUDA error %d (%s) at %s:%d\n", e__, qrdm_rt_errstr(e__), \
              __FILE__, __LINE__);                                                            \
      info = QRDM_ERR_CUDA;                                                                   \
      goto done;                                                                              \
    }                                                                                         \
  } while (0)
  CUX(qrdm_rt_malloc((void **)&d_a, sizeof(double) * (size_t)ldd * kc));
  CUX(qrdm_rt_malloc((void **)&d_c, sizeof(double) * (size_t)ldd * n));
  CUX(qrdm_rt_malloc((void **)&d_tau, sizeof(double) * kc));
  CUX(qrdm_rt_memset(d_c, 0, sizeof(double) * (size_t)ldd * n, stream));
  if (k > 0) {
    CUX(qrdm_rt_h2d_2d(d_a, sizeof(double) * ldd, a, sizeof(double) * lda, sizeof(double) * m, k, stream));
    CUX(qrdm_rt_h2d(d_tau, tau, sizeof(double) * k, stream));
  }
  CUX(qrdm_rt_h2d_2d(d_c, sizeof(double) * ldd, c, sizeof(double) * ldc, sizeof(double) * m, n, stream));
  info = k > 0 ? dormqr_dev_locked(trans, m, n, k, d_a, ldd, d_tau, d_c, ldd, stream) : 0;
  if (info > QRDM_ERR_CUDA) {
    const int keep = info;
    CUX(qrdm_rt_d2h_2d(c, sizeof(double) * ldc, d_c, sizeof(double) * ldd, sizeof(double) * m, n, stream));
    info = keep;
  }
done:
  qrdm_rt_sync(stream); /* no copy may still touch the caller's buffers or ours */
  if (d_a) qrdm_rt_free(d_a);
  if (d_c) qrdm_rt_free(d_c);
  if (d_tau) qrdm_rt_free(d_tau);
  return info;
#undef CUX
}
int qrdm_b200_dormqr(char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc) {
  const int transN = (trans == 'N' || trans == 'n');
  if (!transN && trans != 'T' && trans != 't') return bad_argument(1);
  if (m <= 0) return bad_argument(2);
  if (n <= 0) return bad_argument(3);
  if (k < 0 || k > m) return bad_argument(4);
  if (lda < m) return bad_argument(6);
  if (ldc < m) return bad_argument(9);
  API_BODY(dormqr_locked(trans, m, n, k, a, lda, tau, c, ldc));
}

/* ---- 1-D block-row sharded entry point (one process per GPU; NCCL communicator set up through
 * qrdm_b200_comm_unique_id / qrdm_b200_comm_init, e.g. with the id broadcast by torch.distributed) ---- */
int qrdm_b200_comm_unique_id(char *out128) { API_BODY(qrdm_rt_comm_unique_id(out128) ? QRDM_ERR_COMM : 0); }
static int comm_init_locked(int rank, int nranks, const char *id128) {
  int rc = init_impl(-1);
  if (rc) return rc;
  return qrdm_rt_comm_init(rank, nranks, id128) ? QRDM_ERR_COMM : 0;
}
int qrdm_b200_comm_init(int rank, int nranks, const char *id128) { API_BODY(comm_init_locked(rank, nranks, id128)); }
void qrdm_b200_comm_destroy(void) { api_lock(); qrdm_rt_comm_destroy(); api_unlock(); }

/* Peer memory of the row-sharded path (SURVEY.md 8e "fusion with the collective"): every rank owns one receive
 * buffer in HBM, exports it with CUDA IPC and maps its peers' buffers, so that kernels exchange LL packets by
 * plain stores over NVLink (k_peer.cu).  The launcher moves the 64-byte handles between the processes
 * (qrdm_b200/sharded.py: torch.distributed all_gather) — the same division of labour as for the NCCL unique id. */
static int peer_handle_locked(char *out64) {
  int rc = init_impl(-1);
  if (rc) return rc;
  return qrdm_rt_peer_export(out64) ? QRDM_ERR_COMM : 0;
}
int qrdm_b200_peer_handle(char *out64) { API_BODY(peer_handle_locked(out64)); }
int qrdm_b200_peer_open(int rank, int nranks, const char *handles64) {
  API_BODY(qrdm_rt_peer_open(rank, nranks, handles64) ? QRDM_ERR_COMM : 0);
}
void qrdm_b200_peer_close(void) { api_lock(); qrdm_rt_peer_destroy(); api_unlock(); }

static int sharded_locked(int m_local, int m_global, int row0, int nranks, int n, double *d_a, int lda,
                          int *d_jpvt, double *d_tau, int *ncols, const double *thres, int nb, void *stream) {
  qrdm_shard sh = {row0, m_global, nranks};
  if (m_local > 0) return factor_device(m_local, n, d_a, lda, d_jpvt, d_tau, ncols, thres, nb, stream, NULL, &sh, 0);
  /* A rank that holds no rows at all (m_global < 32 * nranks or so) still has to take part in every exchange and to
   * keep its replicated jpvt / tau / ncols.  It runs the ordinary path on ONE internal all-zero row placed below the
   * matrix (row0 = m_global): a zero row contributes exact zeros to every sum, stays zero under every reflector
   * (v = 0), and no kernel ever sees a local matrix without rows. */
  double *dummy = NULL;
  CU(qrdm_rt_malloc((void **)&dummy, sizeof(double) * 2 * (size_t)n));
  int info = QRDM_ERR_CUDA;
  if (qrdm_rt_memset(dummy, 0, sizeof(double) * 2 * (size_t)n, stream) == 0) {
    sh.row0 = m_global;
    info = factor_device(1, n, dummy, 2, d_jpvt, d_tau, ncols, thres, nb, stream, NULL, &sh, 0);
  }
  qrdm_rt_sync(stream);
  qrdm_rt_free(dummy);
  return info;
}
int dgeqrdm_dev_sharded(int m_local, int m_global, int row0, int nranks, int n, double *d_a, int lda,
                        int *d_jpvt, double *d_tau, int *ncols, const double *thres, int nb, void *stream) {
  if (m_local < 0 || row0 < 0 || row0 + m_local > m_global || nranks < 1) return bad_argument(2);
  int rc = check_args(QRDM_COL_MAJOR, m_global, n, lda > m_global ? lda : m_global, thres, nb);
  if (rc) return rc;
  if (lda < (m_local > 1 ? m_local : 1)) return bad_argument(5);
  API_BODY(sharded_locked(m_local, m_global, row0, nranks, n, d_a, lda, d_jpvt, d_tau, ncols, thres, nb, stream));
}

/* ---- batched mode: `batch` independent m x n matrices, matrix b at
 * a + b*stride_a (column-major, lda), outputs at jpvt + b*n, tau + b*min(m,n), ncols + b*n;
 * multi-GPU = each rank passes its share (independent units, no collective).  Matrices with m, n <= 1024 run as
 * ONE launch with one CTA per matrix (k_small.cu); larger ones fall back to a loop over the one-matrix path.
 * thres[2] (the eta of stop mode 3) is read only when some matrix asks for mode 3 — the reference reads it under
 * `ncols[0] == 3` only (src/dgeqrdm_work.c:539-543) and its notebook passes a 2-element thres. ---- */
static int batched_dev_locked(int batch, int m, int n, double *d_a, int lda, long long stride_a, int *d_jpvt,
                              double *d_tau, int *d_ncols, int *d_infos, const double *thres, int nb, void *stream) {
  int rc = init_impl(-1);
  if (rc) return rc;
  if (!qrdm_k_small_supported(m, n)) return QRDM_ERR_UNSUPPORTED;
  qrdm_workspace *w = &g_ws;
  memset(&g_stats, 0, sizeof(g_stats));
  const long long launches0 = qrdm_rt_launch_count();
  CU(qrdm_rt_event_record(w->ev[0], stream));
  /* the stop modes live on the device here: thres must hold 3 elements (documented in include/qrdm_b200.h) */
  CU(qrdm_k_small(batch, m, n, d_a, lda, stride_a, d_jpvt, d_tau, d_ncols, d_infos, thres[0], thres[1], thres[2], nb,
                  stream));
  CU(qrdm_rt_event_record(w->ev[1], stream));
  CU(qrdm_rt_event_sync(w->ev[1]));
  g_stats.ms_total = qrdm_rt_event_ms(w->ev[0], w->ev[1]);
  g_stats.launches = qrdm_rt_launch_count() - launches0;
  return 0;
}
int dgeqrdm_batched_dev(int batch, int m, int n, double *d_a, int lda, long long stride_a, int *d_jpvt,
                        double *d_tau, int *d_ncols, int *d_infos, const double *thres, int nb, void *stream) {
  if (batch <= 0 || stride_a < (long long)lda * n) return bad_argument(1);
  int rc = check_args(QRDM_COL_MAJOR, m, n, lda, thres, nb);
  if (rc) return rc;
  API_BODY(batched_dev_locked(batch, m, n, d_a, lda, stride_a, d_jpvt, d_tau, d_ncols, d_infos, thres, nb, stream));
}

static int batched_loop(int batch, int m, int n, double *a, int lda, long long stride_a, int *jpvt, double *tau,
                        int *ncols, double *thres, int nb, int *infos);

#define QRDM_BATCH_CHUNK 592 /* matrices per pipeline stage: 4 waves of 148 CTAs */

#define CUX(call)                                                                             \
  do {                                                                                        \
    int e__ = (call);                                                                         \
    if (e__ != 0) {                                                                           \
      fprintf(stderr, "qrdm_b200: CUDA error %d (%s) at %s:%d\n", e__, qrdm_rt_errstr(e__), \
              __FILE__, __LINE__);                                                            \
      worst = QRDM_ERR_CUDA;                                                                  \
      goto done;                                                                              \
    }                                                                                         \
  } while (0)

static int batched_locked(int batch, int m, int n, double *a, int lda, long long stride_a, int *jpvt, double *tau,
                          int *ncols, double *thres, int nb, int *infos) {
  qrdm_workspace *w = &g_ws;
  int rc = init_impl(-1);
  if (rc) return rc;
  if (!qrdm_k_small_supported(m, n) || getenv("QRDM_B200_BATCH_LOOP"))
    return batched_loop(batch, m, n, a, lda, stride_a, jpvt, tau, ncols, thres, nb, infos);
  double eta3 = 0.0;
  for (int b = 0; b < batch; ++b)
    if (ncols[(size_t)n * b] == 3) { eta3 = thres[2]; break; }
  /* three streams: chunk c+1 is uploaded and chunk c-1 downloaded while chunk c is factored */
  if (!w->h2d_stream) CU(qrdm_rt_stream_create(&w->h2d_stream));
  void *s_up = w->h2d_stream, *s_k = w->compute_stream, *s_down = w->copy_stream;
  const int minmn = m < n ? m : n;
  const int ldd = (m + 1) & ~1;
  const size_t per = (size_t)ldd * n;
  const int chunk = batch < QRDM_BATCH_CHUNK ? batch : QRDM_BATCH_CHUNK;
  const int nbuf = batch > chunk ? 3 : 1; /* ring of device chunks */
  double *d_all = NULL, *d_tau = NULL;
  int *d_jpvt = NULL, *d_ncols = NULL, *d_infos = NULL, *h_infos = NULL;
  void *ev_up[3] = {0}, *ev_k[3] = {0}, *ev_down[3] = {0};
  int worst = 0;
  h_infos = (int *)malloc(sizeof(int) * (size_t)batch);
  if (!h_infos) return QRDM_ERR_CUDA;
  CUX(qrdm_rt_malloc((void **)&d_all, sizeof(double) * per * chunk * nbuf));
  CUX(qrdm_rt_malloc((void **)&d_tau, sizeof(double) * (size_t)minmn * chunk * nbuf));
  CUX(qrdm_rt_malloc((void **)&d_jpvt, sizeof(int) * (size_t)n * chunk * nbuf));
  CUX(qrdm_rt_malloc((void **)&d_ncols, sizeof(int) * (size_t)n * chunk * nbuf));
  CUX(qrdm_rt_malloc((void **)&d_infos, sizeof(int) * (size_t)chunk * nbuf));
  for (int q = 0; q < nbuf; ++q) {
    CUX(qrdm_rt_event_create(&ev_up[q]));
    CUX(qrdm_rt_event_create(&ev_k[q]));
    CUX(qrdm_rt_event_create(&ev_down[q]));
  }
  memset(&g_stats, 0, sizeof(g_stats));
  const long long launches0 = qrdm_rt_launch_count();
  CUX(qrdm_rt_event_record(w->ev[0], s_k));
  const int nchunks = (batch + chunk - 1) / chunk;
  for (int c = 0; c < nchunks; ++c) {
    const int q = c % nbuf, b0 = c * chunk, cnt = batch - b0 < chunk ? batch - b0 : chunk;
    double *da = d_all + per * chunk * q;
    double *dt = d_tau + (size_t)minmn * chunk * q;
    int *dj = d_jpvt + (size_t)n * chunk * q, *dn = d_ncols + (size_t)n * chunk * q, *di = d_infos + (size_t)chunk * q;
    if (c >= nbuf) CUX(qrdm_rt_stream_wait_event(s_up, ev_down[q])); /* ring slot free again */
    if (lda == m && ldd == m && stride_a == (long long)m * n) {
      CUX(qrdm_rt_h2d(da, a + (size_t)stride_a * b0, sizeof(double) * per * cnt, s_up));
    } else {
      for (int b = 0; b < cnt; ++b)
        CUX(qrdm_rt_h2d_2d(da + per * b, sizeof(double) * ldd, a + (size_t)stride_a * (b0 + b), sizeof(double) * lda,
                           sizeof(double) * m, n, s_up));
    }
    CUX(qrdm_rt_h2d(dt, tau + (size_t)minmn * b0, sizeof(double) * (size_t)minmn * cnt, s_up));
    CUX(qrdm_rt_h2d(dn, ncols + (size_t)n * b0, sizeof(int) * (size_t)n * cnt, s_up));
    CUX(qrdm_rt_event_record(ev_up[q], s_up));
    CUX(qrdm_rt_stream_wait_event(s_k, ev_up[q]));
    CUX(qrdm_k_small(cnt, m, n, da, ldd, (long long)per, dj, dt, dn, di, thres[0], thres[1], eta3, nb, s_k));
    CUX(qrdm_rt_event_record(ev_k[q], s_k));
    CUX(qrdm_rt_stream_wait_event(s_down, ev_k[q]));
    if (lda == m && ldd == m && stride_a == (long long)m * n) {
      CUX(qrdm_rt_d2h(a + (size_t)stride_a * b0, da, sizeof(double) * per * cnt, s_down));
    } else {
      for (int b = 0; b < cnt; ++b)
        CUX(qrdm_rt_d2h_2d(a + (size_t)stride_a * (b0 + b), sizeof(double) * lda, da + per * b, sizeof(double) * ldd,
                           sizeof(double) * m, n, s_down));
    }
    CUX(qrdm_rt_d2h(jpvt + (size_t)n * b0, dj, sizeof(int) * (size_t)n * cnt, s_down));
    CUX(qrdm_rt_d2h(tau + (size_t)minmn * b0, dt, sizeof(double) * (size_t)minmn * cnt, s_down));
    CUX(qrdm_rt_d2h(ncols + (size_t)n * b0, dn, sizeof(int) * (size_t)n * cnt, s_down));
    CUX(qrdm_rt_d2h(h_infos + b0, di, sizeof(int) * (size_t)cnt, s_down));
    CUX(qrdm_rt_event_record(ev_down[q], s_down));
  }
  CUX(qrdm_rt_event_record(w->ev[1], s_k));
  CUX(qrdm_rt_sync(s_down));
  CUX(qrdm_rt_sync(s_k));
  g_stats.ms_total = qrdm_rt_event_ms(w->ev[0], w->ev[1]);
  g_stats.launches = qrdm_rt_launch_count() - launches0;
  for (int b = 0; b < batch; ++b) {
    if (infos) infos[b] = h_infos[b];
    if (h_infos[b] != 0 && worst == 0) worst = h_infos[b];
  }
done:
  /* every stream idle before the buffers go away / the caller gets its arrays back */
  qrdm_rt_sync(s_up);
  qrdm_rt_sync(s_k);
  qrdm_rt_sync(s_down);
  free(h_infos);
  for (int q = 0; q < nbuf; ++q) { qrdm_rt_event_destroy(ev_up[q]); qrdm_rt_event_destroy(ev_k[q]); qrdm_rt_event_destroy(ev_down[q]); }
  if (d_all) qrdm_rt_free(d_all);
  if (d_tau) qrdm_rt_free(d_tau);
  if (d_jpvt) qrdm_rt_free(d_jpvt);
  if (d_ncols) qrdm_rt_free(d_ncols);
  if (d_infos) qrdm_rt_free(d_infos);
  return worst;
}

int dgeqrdm_batched(int batch, int m, int n, double *a, int lda, long long stride_a, int *jpvt, double *tau,
                    int *ncols, double *thres, int nb, int *infos) {
  if (batch <= 0 || stride_a < (long long)lda * n) return bad_argument(1);
  int rc = check_args(QRDM_COL_MAJOR, m, n, lda, thres, nb);
  if (rc) return rc;
  API_BODY(batched_locked(batch, m, n, a, lda, stride_a, jpvt, tau, ncols, thres, nb, infos));
}

/* matrices too large for the one-CTA kernel: one after the other through the one-matrix path */
static int batched_loop(int batch, int m, int n, double *a, int lda, long long stride_a, int *jpvt, double *tau,
                        int *ncols, double *thres, int nb, int *infos) {
  qrdm_workspace *w = &g_ws;
  void *stream = w->compute_stream;
  const int minmn = m < n ? m : n;
  const int ldd = (m + 1) & ~1;
  double *d_all = NULL, *d_tau = NULL;
  int *d_jpvt = NULL;
  const size_t per = (size_t)ldd * n;
  int worst = 0;
  double ms = 0.0;
  long long launches = 0;
  CUX(qrdm_rt_malloc((void **)&d_all, sizeof(double) * per * batch));
  CUX(qrdm_rt_malloc((void **)&d_tau, sizeof(double) * (size_t)minmn * batch));
  CUX(qrdm_rt_malloc((void **)&d_jpvt, sizeof(int) * (size_t)n * batch));
  for (int b = 0; b < batch; ++b)
    CUX(qrdm_rt_h2d_2d(d_all + per * b, sizeof(double) * ldd, a + (size_t)stride_a * b, sizeof(double) * lda,
                       sizeof(double) * m, n, stream));
  CUX(qrdm_rt_h2d(d_tau, tau, sizeof(double) * (size_t)minmn * batch, stream));
  for (int b = 0; b < batch; ++b) {
    int info = factor_device(m, n, d_all + per * b, ldd, d_jpvt + (size_t)n * b, d_tau + (size_t)minmn * b,
                             ncols + (size_t)n * b, thres, nb, stream, NULL, NULL, 0);
    if (infos) infos[b] = info;
    if (info <= QRDM_ERR_CUDA) { worst = info; goto done; }
    if (info != 0 && worst == 0) worst = info;
    ms += g_stats.ms_total;
    launches += g_stats.launches;
  }
  {
    const int keep = worst;
    for (int b = 0; b < batch; ++b)
      CUX(qrdm_rt_d2h_2d(a + (size_t)stride_a * b, sizeof(double) * lda, d_all + per * b, sizeof(double) * ldd,
                         sizeof(double) * m, n, stream));
    CUX(qrdm_rt_d2h(jpvt, d_jpvt, sizeof(int) * (size_t)n * batch, stream));
    CUX(qrdm_rt_d2h(tau, d_tau, sizeof(double) * (size_t)minmn * batch, stream));
    CUX(qrdm_rt_sync(stream));
    worst = keep;
  }
  g_stats.ms_total = ms;
  g_stats.launches = launches;
done:
  qrdm_rt_sync(stream);
  if (d_all) qrdm_rt_free(d_all);
  if (d_tau) qrdm_rt_free(d_tau);
  if (d_jpvt) qrdm_rt_free(d_jpvt);
  return worst;
}
#undef CUX
