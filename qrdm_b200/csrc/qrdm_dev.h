/* qrdm_dev.h — the thin CUDA C-ABI layer under the host C driver (dgeqrdm_host.c).
 * Every function launches one kernel (or does one runtime call) on the given stream and returns
 * a cudaError_t as int (0 = success).  No C++ types cross this boundary. */
#ifndef QRDM_DEV_H_
#define QRDM_DEV_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QRDM_KMAX 64       /* width of the panel / trailing-update tiles: a block of up to 64 columns per pass */
#define QRDM_CANDMAX 256   /* max nb = max candidates / selected columns per iteration (QRDM_NB_MAX); blocks wider than
                              QRDM_KMAX are factored in micro-panels of QRDM_KMAX columns (k_wide.cu) */
#define QRDM_WIDE_ROWCTAS 148 /* row chunks of the wide Gram kernel */
#define QRDM_MAXEX 160     /* max column exchanges planned per iteration (<= 2*KMAX, see k_pick) */
#define QRDM_MAXPOS 320    /* max column positions touched by those exchanges */
#define QRDM_SELCAP 1024   /* capacity of the top-k candidate list in k_select */
#define QRDM_GRAM_MAXCTA 296
#define QRDM_PANEL_MAXCTA 148
/* LL packet buffers of the panel kernels (16-byte packets): the per-column kernels use the first region of each buffer,
 * the grouped kernel (k_panel_grp: one exchange per QRDM_PANEL_SMAX columns at most) the second */
#define QRDM_PANEL_SMAX 4
#define QRDM_PANEL_PART_PKTS (2 * QRDM_PANEL_MAXCTA * 64)
#define QRDM_PANEL_GPART_PKTS (2 * QRDM_PANEL_MAXCTA * 64 * QRDM_PANEL_SMAX)
#define QRDM_PANEL_ROW_PKTS (2 * 128)
#define QRDM_PANEL_GROW_PKTS (2 * 64 * QRDM_PANEL_SMAX)
#define QRDM_ERR_INTERNAL (-103) /* planner overflow: cannot happen for nb <= QRDM_KMAX */
#define QRDM_TALL_B 8     /* sub-panel width of the blocked panel used when the slab does not fit on chip */
#define QRDM_ROWALIGN 32   /* row tiles of the trailing kernels start at multiples of this */

/* Device-resident control block.  The first QRDM_MAILBOX_BYTES are mirrored into a pinned host
 * mailbox once per iteration; that copy is the only host<->device synchronisation of the loop. */
typedef struct qrdm_ctrl {
  int j;        /* columns triangularised so far = first active row/column (0-based) */
  int last_k;   /* block size of the iteration that just finished (-> ncols[it]) */
  int kmax;     /* min(nb, rows, cols) of the iteration being prepared (0 = nothing left) */
  int nc;       /* candidates that passed the norm filter */
  int fjb;      /* columns selected by the greedy cosine pick */
  int fjb_cmp;  /* columns the panel actually triangularised (early stop) */
  int ncyc;     /* permutation cycles to apply to the columns of A */
  int nflag;    /* columns whose partial norm must be recomputed exactly */
  int err;      /* 0, or the reference's info code (-8 tau NaN, -6 V NaN, -13 C NaN) */
  int it;       /* iterations completed */
  int pad_[2];
  double maxnrm; /* max partial column norm of the active columns (stop rule) */
  double pad2_;
  /* ---- device-only part ---- */
  unsigned int panel_bar; /* (unused since the LL exchange) */
  int mg_k;               /* row-sharded panel: column at which the early stop fired, or -1 */
  int pad3_[2];
  double mg_thres2;       /* row-sharded panel: squared stop threshold carried across step kernels */
  /* tall (blocked) panel: state carried across the 8-column sub-panels */
  int sub_k;              /* reflectors produced by the current sub-panel */
  int tall_k;             /* reflectors produced so far by this panel */
  int tall_done;          /* 1 once the early stop fired: later sub-panels are no-ops */
  int tall_stop_s;        /* panel column at which the sub-panel that stopped early starts (valid when tall_done) */
  double tall_thres;      /* SQUARED stop threshold (set by the panel's first column), carried across sub-panels */
  int cand[QRDM_CANDMAX];        /* candidate column offsets (relative to j), by norm descending */
  double candnrm[QRDM_CANDMAX];  /* their partial norms */
  int sel[QRDM_CANDMAX];         /* accepted offsets, acceptance order */
  int cyc_start[QRDM_MAXPOS + 1];
  int cyc_pos[QRDM_MAXPOS];   /* cycle c: new[p_k] = old[p_{k+1}], new[p_last] = old[p_0] */
  /* deferred ("lazy") trailing update: block reflector whose pass 2 has not been applied to the bulk
   * of the trailing matrix yet (written by k_wapply in rows mode, read by k_fused / k_colupd / k_rankk<list> / the flush) */
  int pend_k;   /* reflectors of the pending block */
  int pend_c0;  /* first column the pending update applies to (= j + fjb of its iteration) */
  int pend_r0;  /* first row still to be updated (= j + k of its iteration: the k new R rows are done) */
  int forced;   /* 1: this iteration factors FIXED columns (jpvt != 0 on entry): the next columns as they stand, no DM
                   selection, no permutation, no early stop in the panel */
  long long stat_perm_cols; /* statistics: columns moved by k_permute since the start of the factorisation */
  /* nb > 64: the selected block is factored in micro-panels of QRDM_KMAX columns (k_wide.cu) */
  int w_j0, w_fjb;     /* j and fjb of the whole block */
  int w_k;             /* reflectors produced by the micro-panels finished so far */
  int w_stop;          /* a micro-panel stopped early: the block ends there */
  int w_pending;       /* a micro-panel has run and is not yet accounted for */
  int micro_t;         /* index of the current micro-panel (> 0: the panel continues a block — stop test from its first
                          column on, threshold carried in micro_thres2) */
  double micro_thres2; /* squared DM stop threshold carried across micro-panels */
} qrdm_ctrl;
#define QRDM_MAILBOX_BYTES 64

/* Problem + workspace descriptor passed by value to every launcher. */
typedef struct qrdm_prob {
  int m, n, lda, nb;
  double delta, tau_; /* thres[0] (cosine bound), thres[1] (norm fraction) */
  double *a;          /* m x n column-major, device */
  int *jpvt;          /* n, device, 1-based on exit */
  double *tau;        /* min(m,n), device */
  double *vn1, *vn2;  /* n each: partial norms / norms at last exact recompute */
  qrdm_ctrl *ctrl;
  double *gram_part;  /* [QRDM_GRAM_MAXCTA][64*64] partial Gram blocks */
  double *gram;       /* [64*64] reduced Gram (candidates, then V'V) */
  double *panel_part; /* [2][QRDM_PANEL_MAXCTA][64] 16-byte LL packets (partial column sums) */
  double *panel_row;  /* [2][128] LL packets: 64 column totals + 64 pivot-row entries */
  double *vc;         /* ldv x 64 clean copy of V (unit diagonal, zeros above), rows = global rows */
  int ldv;            /* multiple of QRDM_ROWALIGN, >= roundup(m, QRDM_ROWALIGN) */
  double *wp;         /* [splits][64][ldw] partial W = V'C */
  size_t wp_elems;
  double *w2;         /* [64][ldw]  T'W, row-major */
  int ldw;            /* multiple of 4, >= n */
  double *nrm_part;   /* [nsplit][n] partial sums of squares */
  int nrm_splits;
  int *flag_list;     /* n */
  int sm_count;
  int vec16;          /* 1 if a is 16-byte aligned and lda is even (fast cp.async path) */
  int debug;          /* QRDM_B200_DEBUG bit mask for timing experiments (results invalid when set) */
  /* 1-D block-row sharding: this rank holds global rows [row0, row0 + m) of all n
   * columns; m_glob = total rows.  Single GPU: row0 = 0, m_glob = m, nranks = 1. */
  int row0, m_glob, nranks;
  int w_reduced;      /* 1: partial-W slot 0 already holds the (all-reduced) sum over all slots */
  int sub;            /* tall blocked panel: s + 1 when the kernels work on the sub-panel starting at panel
                         column s (geometry from qrdm_sub_geom), 0 otherwise */
  double *mg_buf;     /* sharded panel: [2][128] send/recv vectors + [G][128] per-CTA partials */
  unsigned *mg_cnt;   /* sharded panel: arrival counter of the last-CTA reduction */
  /* deferred trailing update (k_fused): vc / vc_prev alternate between two ldv x 64 buffers */
  double *vc_prev;    /* clean V of the pending block */
  int *upd_flag;      /* [n] == stamp: column already brought up to date by k_colupd (flagged-norm list) */
  int *upd_eager;     /* [n] == stamp: likewise for the eager set (leading 64 positions + candidates) */
  int stamp;          /* id of the pending block (> 0) */
  int pend;           /* 1: kernels take their geometry from ctrl->pend_* (flush of a pending block) */
  int nfxd;           /* number of fixed columns, already moved to the front (src/dgeqrdm_work.c:592-607); 0 = none */
  int keep_jpvt;      /* 1: d_jpvt holds the caller's initial permutation, K1 must not reset it to the identity */
  double thres0;      /* the panel's initial absolute stop threshold 5e-14 (src/dgeqr2.c:40) times the power-of-two input
                         scale (1 unless the matrix was pre-scaled, see qrdm_k_scale) */
  /* look-ahead of the deferred update: while the selection / Gram / pick / permutation (and, on SMs the
   * panel leaves free, the panel) of the next block run, a second stream applies pass 2 of the pending block to the
   * columns >= side_col0 that nobody touches in that window (k_rankk in side mode skips stamped columns); the next
   * k_fused / k_tinv / k_wapply are told so through pre_col0: those columns need pass 1 only. */
  int side_col0;      /* side launch: first absolute column it owns (0 = not a side launch) */
  int pre_col0;       /* k_fused & co: columns >= pre_col0 already hold the pending update (0 = none) */
  int no_vtv;         /* 1: k_vtc skips its V'V tile (tile numbering starts at 1), k_tinv takes V'V from p->gram (qrdm_k_vtv): tall-skinny matrices */
  int vt_wb;          /* relative cost of a pass-1-only unit of k_fused against VT_WA = 7 for a full one (0: the default, 4) */
  double inv_scale;   /* 1 / that scale (MUST be 1.0, never 0, for an unscaled matrix): the norm downdate evaluates its
                         sum of squares in the CALLER's scale, where the reference's unscaled sum (src/dgeqrdm_work.c:81-86)
                         underflows to 0 for ~1e-200 entries (no downdate) and overflows for ~1e+200 (forced recompute) */
} qrdm_prob;

int qrdm_k_colnorm(const qrdm_prob *p, int use_flag_list, void *stream);   /* K1 / K2 recompute */
/* badly scaled inputs (max column norm beyond 2^+-300, where sums of squares leave the double range): max |a_ij| per
 * CTA -> out[0..*nparts), and A *= s (mode 0: every entry; mode 1: the R-like entries of a factored matrix of rank r —
 * rows <= c of the columns c < r, whole columns c >= r — the Householder vectors are scale invariant) */
int qrdm_k_amax(const qrdm_prob *p, double *out, int *nparts, void *stream);
int qrdm_k_scale(const qrdm_prob *p, double s, int mode, int r, void *stream);
int qrdm_k_select(const qrdm_prob *p, void *stream);                        /* K3a */
int qrdm_k_gram(const qrdm_prob *p, int of_v, int rows_hint, void *stream); /* K3b / K5 */
int qrdm_k_vtv(const qrdm_prob *p, int rows_hint, void *stream);            /* V'V of the block -> p->gram (TMA + DMMA), for p->no_vtv */
int qrdm_k_pick(const qrdm_prob *p, void *stream);                          /* K3c + plan */
int qrdm_k_permute(const qrdm_prob *p, void *stream);                       /* K3d */
int qrdm_k_panel(const qrdm_prob *p, int j_host, void *stream);             /* K4 */
int qrdm_k_panel_ctas(const qrdm_prob *p, int rows);  /* SMs the panel kernel will occupy for a panel of `rows` rows */
int qrdm_k_trailing(const qrdm_prob *p, int j_host, void *stream);          /* K6: vtc, wsolve, rankk */
int qrdm_k_norm_update(const qrdm_prob *p, int j_host, void *stream);       /* K2 */
int qrdm_k_norm_recompute_all(const qrdm_prob *p, int j_host, void *stream); /* exact norms of every column right of the block (end of the fixed-column phase, src/dgeqrdm_work.c:672-682) */
/* row-sharded variants: each stage is split at the point where the all-reduce sits */
int qrdm_k_colnorm_part(const qrdm_prob *p, int use_flag_list, int *nsplit_out, void *stream);
int qrdm_k_colnorm_fin(const qrdm_prob *p, int use_flag_list, int nsplit, void *stream);
int qrdm_k_gram_part(const qrdm_prob *p, int rows_hint, void *stream); /* partial + local reduce -> p->gram */
int qrdm_k_norm_dpart(const qrdm_prob *p, int j_host, void *stream);   /* d[c] partial -> nrm_part[0..n) */
int qrdm_k_norm_apply(const qrdm_prob *p, int j_host, void *stream);
int qrdm_k_vtc_only(const qrdm_prob *p, int j_host, int *stride_out, int *grid_out, void *stream);
int qrdm_k_wreduce(const qrdm_prob *p, int j_host, int vt_grid, int stride, void *stream);
int qrdm_k_trailing_finish(const qrdm_prob *p, int j_host, int vt_grid, int stride, void *stream);
/* deferred trailing update: pass 2 of the pending block fused into pass 1 of the current one */
int qrdm_k_fused(const qrdm_prob *p, int j_host, int *stride_out, int *grid_out, void *stream);
int qrdm_k_w2(const qrdm_prob *p, int j_host, int vt_grid, int stride, int bn_and_flags, void *stream); /* T', W2 = -T'W; bn | 1: + R rows, | 2: use T */
int qrdm_k_rankk(const qrdm_prob *p, int j_host, void *stream);    /* pass 2 alone */
int qrdm_k_colupd(const qrdm_prob *p, int mode, int j_host, void *stream); /* 0: eager set (DMMA, gathered), 1: flagged-norm list (FMA) */
int qrdm_k_norm_update_lazy(const qrdm_prob *p, int j_host, void *stream);
int qrdm_k_flush(const qrdm_prob *p, int j_host, void *stream);    /* apply a pending block to the whole trailing matrix */
/* look-ahead: pass 2 of the pending block on the columns >= p->side_col0 (unstamped ones), ~units_per_cta units per CTA */
int qrdm_k_side(const qrdm_prob *p, int j_host, int units_per_cta, void *stream);
int qrdm_k_vc_build(const qrdm_prob *p, const double *d_af, int ldf, int j0, int k, void *stream); /* Vc + ctrl of one block of a factored matrix */
int qrdm_k_skinny_update(const qrdm_prob *p, int rows_hint, void *stream); /* tall panel: sub-panel -> rest of panel */
int qrdm_k_skinny_part(const qrdm_prob *p, int rows_hint, void *stream);   /* row-sharded: before the all-reduce */
int qrdm_k_skinny_finish(const qrdm_prob *p, int rows_hint, void *stream); /* row-sharded: after it */
int qrdm_k_panel_mg_init(const qrdm_prob *p, int j_host, void *stream);
int qrdm_k_panel_mg_step(const qrdm_prob *p, int j_host, int step, void *stream);
int qrdm_k_panel_mg_finish(const qrdm_prob *p, int j_host, void *stream);
/* batched mode: one CTA per matrix, the whole factorisation in one launch (k_small.cu).  d_ncols is
 * [batch][n] with the stop-rule mode in [b][0] on entry; d_infos [batch] or NULL. */
int qrdm_k_small_supported(int m, int n);
int qrdm_k_small(int batch, int m, int n, double *d_a, int lda, long long stride_a, int *d_jpvt, double *d_tau,
                 int *d_ncols, int *d_infos, double delta, double tau_, double eta3, int nb, void *stream);
/* NCCL (dlopen'ed libnccl.so.2): in-place sum all-reduce of doubles on the stream */
int qrdm_rt_comm_unique_id(char *out128);
int qrdm_rt_comm_init(int rank, int nranks, const char *id128);
int qrdm_rt_comm_destroy(void);
int qrdm_rt_allreduce(double *buf, size_t count, void *stream);

/* runtime helpers so that the host driver stays plain C */
int qrdm_rt_malloc(void **ptr, size_t bytes);
int qrdm_rt_free(void *ptr);
int qrdm_rt_host_alloc(void **ptr, size_t bytes);
int qrdm_rt_host_free(void *ptr);
int qrdm_rt_memset(void *ptr, int v, size_t bytes, void *stream);
int qrdm_rt_h2d(void *dst, const void *src, size_t bytes, void *stream);
int qrdm_rt_d2h(void *dst, const void *src, size_t bytes, void *stream);
int qrdm_rt_h2d_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream);
int qrdm_rt_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream);
int qrdm_rt_sync(void *stream);
int qrdm_rt_stream_create(void **stream);
int qrdm_rt_stream_create_prio(void **stream, int prio); /* > 0: greatest priority of the device, <= 0: least */
int qrdm_rt_is_pinned(const void *ptr);
int qrdm_rt_stream_wait_event(void *stream, void *ev);
int qrdm_rt_event_create(void **ev);
int qrdm_rt_event_destroy(void *ev);
int qrdm_rt_event_record(void *ev, void *stream);
int qrdm_rt_event_sync(void *ev);
double qrdm_rt_event_ms(void *ev0, void *ev1);
int qrdm_rt_device_info(int *sm_count, size_t *free_bytes);
int qrdm_rt_set_device(int dev);
int qrdm_rt_get_device(int *dev);
int qrdm_rt_device_generation(void);
void qrdm_rt_new_device_generation(void);
int qrdm_rt_stream_destroy(void *stream);
double qrdm_rt_copy_gbs(size_t bytes, void *stream);
/* peer memory of the row-sharded path (k_peer.cu): CUDA-IPC receive buffers + the one-shot LL all-reduce */
int qrdm_rt_peer_export(char *out64);
int qrdm_rt_peer_open(int rank, int nranks, const char *handles64);
int qrdm_rt_peer_destroy(void);
int qrdm_rt_peer_available(void);              /* number of ranks of the open peer context, 0 = none */
int qrdm_k_peer_allreduce(double *buf, size_t count, void *stream);
int qrdm_k_panel_tall_mg(const qrdm_prob *p, int j_host, void *stream); /* sharded sub-panel, exchange inside the kernel */
int qrdm_k_skinny_update_mg(const qrdm_prob *p, int rows_hint, void *stream); /* sharded skinny update, cross-GPU sum inside k_sub_w2 */
/* nb > 64 (k_wide.cu): part = [pairs][QRDM_WIDE_ROWCTAS][4096], G = [QRDM_CANDMAX][QRDM_CANDMAX], marks = [n] zeroed once,
 * swaps = [2 * (2 * QRDM_CANDMAX + 4) + 1] */
int qrdm_k_gram_wide(const qrdm_prob *p, double *part, double *G, int rows_hint, void *stream);
int qrdm_k_pick_wide(const qrdm_prob *p, const double *G, int *marks, int *swaps, void *stream);
int qrdm_k_micro_begin(const qrdm_prob *p, int t, void *stream);
int qrdm_k_micro_end(const qrdm_prob *p, void *stream);
/* blocked QR with classical column pivoting (k_qp3.cu): the whole factorisation, one 64-byte mailbox read per panel */
int qrdm_qp3_dev(const qrdm_prob *p, void *mailbox, void *stream);
const char *qrdm_rt_errstr(int code);
long long qrdm_rt_launch_count(void);
double qrdm_rt_fp64_peak(int use_dmma, void *stream);

#ifdef __cplusplus
}
#endif
#endif
