/* hostio.c — host<->device transfers of the reference-facing entry point for PAGEABLE host buffers.
 *
 * The boundary's real caller hands over NumPy memory (reference QRDM_wrapper.c:89-96, 155-159: raw
 * `->data` pointers of pageable arrays).  A plain cudaMemcpy from pageable memory is staged by the driver
 * through one small pinned buffer on one thread and reaches a fraction of the PCIe rate, and it cannot
 * overlap with the factorisation.  Here:
 *
 *   upload / download   T worker threads; worker t owns tiles t, t+T, ... of the matrix, two pinned bounce
 *                       buffers, a stream and two events: memcpy(tile -> pinned[k&1]) -> async H2D, the memcpy
 *                       of its next tile overlapping the DMA of the previous one (and the other workers').
 *                       A tile is a group of whole columns, or a row segment of one column when a column is
 *                       longer than a bounce buffer (tall-skinny inputs).
 *   write-back stream   while the factorisation runs, finished column ranges are copied D2H into a pinned
 *                       ring on the copy stream; a drain thread waits for each copy's event and moves the data
 *                       into the caller's buffer — the pageable counterpart of the direct D2H used for pinned
 *                       buffers.  Whatever does not fit the ring is left to the final parallel download.
 *
 * Plain C + pthreads over the qrdm_rt_* wrappers, so tests/test_hostio.py can link the same file against a
 * mock runtime (memcpy "device") and run the tiling/threading logic on a CPU-only box.
 */
#include "hostio.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qrdm_dev.h"

#define IO_MAX_THREADS 8
#define IO_TILE_BYTES ((size_t)8 << 20)   /* one bounce buffer */
#define IO_RING_BYTES ((size_t)128 << 20) /* write-back ring */
#define IO_RING_JOBS 64

typedef struct {
  size_t off, bytes; /* ring region */
  int c0, c1;        /* columns [c0, c1) of the host matrix */
  void *ev;
} wb_job;

struct qrdm_hostio {
  int device, nthreads;
  void *pinned[IO_MAX_THREADS][2];
  void *stream[IO_MAX_THREADS];
  void *ev[IO_MAX_THREADS][2];
  /* write-back ring */
  char *ring;
  void *ring_ev[IO_RING_JOBS];
  wb_job jobs[IO_RING_JOBS];
  int job_head, job_tail;    /* producer writes head, drain thread advances tail */
  int wb_active, wb_stop, wb_error;
  pthread_t wb_thread;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  double *wb_h;
  const double *wb_d;
  int wb_ldh, wb_ldd, wb_m;
  void *wb_stream;
};

typedef struct {
  qrdm_hostio *io;
  int t, dir; /* dir 0: host -> device, 1: device -> host */
  double *d;
  double *h;
  int ldd, ldh, m, c0, n;
  int err;
} io_task;

static int io_threads_default(void) {
  const char *e = getenv("QRDM_B200_IO_THREADS");
  int t = e ? atoi(e) : 6;
  if (t < 1) t = 1;
  if (t > IO_MAX_THREADS) t = IO_MAX_THREADS;
  return t;
}

int qrdm_hostio_create(qrdm_hostio **out, int device) {
  qrdm_hostio *io = (qrdm_hostio *)calloc(1, sizeof(*io));
  if (!io) return -1;
  io->device = device;
  io->nthreads = io_threads_default();
  pthread_mutex_init(&io->mu, NULL);
  pthread_cond_init(&io->cv, NULL);
  for (int t = 0; t < io->nthreads; ++t) {
    if (qrdm_rt_stream_create(&io->stream[t])) goto fail;
    for (int k = 0; k < 2; ++k) {
      if (qrdm_rt_host_alloc(&io->pinned[t][k], IO_TILE_BYTES)) goto fail;
      if (qrdm_rt_event_create(&io->ev[t][k])) goto fail;
    }
  }
  *out = io;
  return 0;
fail:
  qrdm_hostio_destroy(io);
  return -1;
}

void qrdm_hostio_destroy(qrdm_hostio *io) {
  if (!io) return;
  for (int t = 0; t < IO_MAX_THREADS; ++t) {
    for (int k = 0; k < 2; ++k) {
      if (io->pinned[t][k]) qrdm_rt_host_free(io->pinned[t][k]);
      if (io->ev[t][k]) qrdm_rt_event_destroy(io->ev[t][k]);
    }
    if (io->stream[t]) qrdm_rt_stream_destroy(io->stream[t]);
  }
  if (io->ring) qrdm_rt_host_free(io->ring);
  for (int i = 0; i < IO_RING_JOBS; ++i)
    if (io->ring_ev[i]) qrdm_rt_event_destroy(io->ring_ev[i]);
  pthread_mutex_destroy(&io->mu);
  pthread_cond_destroy(&io->cv);
  free(io);
}

/* Tile geometry: a column of m doubles either fits a bounce buffer (tiles = groups of `cpt` whole columns) or it
 * does not (tiles = row segments of `rpt` rows of one column). */
typedef struct { int cpt, rpt, segs; long long ntiles; } io_geom;
static io_geom io_geometry(int m, int ncols) {
  io_geom g;
  const size_t colb = (size_t)m * sizeof(double);
  if (colb <= IO_TILE_BYTES) {
    g.cpt = (int)(IO_TILE_BYTES / colb);
    if (g.cpt < 1) g.cpt = 1;
    g.rpt = m; g.segs = 1;
    g.ntiles = ((long long)ncols + g.cpt - 1) / g.cpt;
  } else {
    g.cpt = 1;
    g.rpt = (int)(IO_TILE_BYTES / sizeof(double));
    g.segs = (m + g.rpt - 1) / g.rpt;
    g.ntiles = (long long)ncols * g.segs;
  }
  return g;
}

static void *io_worker(void *arg) {
  io_task *tk = (io_task *)arg;
  qrdm_hostio *io = tk->io;
  const int t = tk->t, m = tk->m;
  if (qrdm_rt_set_device(io->device)) { tk->err = -1; return NULL; }
  const io_geom g = io_geometry(m, tk->n - tk->c0);
  int k = 0, used[2] = {0, 0};
  /* pending host-side unpack of a download tile: done when its buffer comes round again / at the end */
  struct { int c, nc, r0, nr; } pend[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  for (long long ti = t; ti < g.ntiles; ti += io->nthreads, k ^= 1) {
    int c, nc, r0, nr;
    if (g.segs == 1) {
      c = tk->c0 + (int)ti * g.cpt;
      nc = tk->n - c < g.cpt ? tk->n - c : g.cpt;
      r0 = 0; nr = m;
    } else {
      c = tk->c0 + (int)(ti / g.segs);
      nc = 1;
      r0 = (int)(ti % g.segs) * g.rpt;
      nr = m - r0 < g.rpt ? m - r0 : g.rpt;
    }
    double *pb = (double *)io->pinned[t][k];
    if (used[k]) { /* the buffer's previous transfer must have finished */
      if (qrdm_rt_event_sync(io->ev[t][k])) { tk->err = -1; return NULL; }
      if (tk->dir == 1)
        for (int cc = 0; cc < pend[k].nc; ++cc)
          memcpy(tk->h + (size_t)(pend[k].c + cc) * tk->ldh + pend[k].r0, pb + (size_t)cc * pend[k].nr,
                 sizeof(double) * (size_t)pend[k].nr);
    }
    if (tk->dir == 0) {
      for (int cc = 0; cc < nc; ++cc)
        memcpy(pb + (size_t)cc * nr, tk->h + (size_t)(c + cc) * tk->ldh + r0, sizeof(double) * (size_t)nr);
      if (qrdm_rt_h2d_2d(tk->d + (size_t)c * tk->ldd + r0, sizeof(double) * (size_t)tk->ldd, pb, sizeof(double) * (size_t)nr,
                         sizeof(double) * (size_t)nr, (size_t)nc, io->stream[t])) { tk->err = -1; return NULL; }
    } else {
      if (qrdm_rt_d2h_2d(pb, sizeof(double) * (size_t)nr, tk->d + (size_t)c * tk->ldd + r0, sizeof(double) * (size_t)tk->ldd,
                         sizeof(double) * (size_t)nr, (size_t)nc, io->stream[t])) { tk->err = -1; return NULL; }
      pend[k].c = c; pend[k].nc = nc; pend[k].r0 = r0; pend[k].nr = nr;
    }
    if (qrdm_rt_event_record(io->ev[t][k], io->stream[t])) { tk->err = -1; return NULL; }
    used[k] = 1;
  }
  for (int q = 0; q < 2; ++q, k ^= 1) { /* drain, oldest buffer first */
    if (!used[k]) continue;
    if (qrdm_rt_event_sync(io->ev[t][k])) { tk->err = -1; return NULL; }
    if (tk->dir == 1) {
      double *pb = (double *)io->pinned[t][k];
      for (int cc = 0; cc < pend[k].nc; ++cc)
        memcpy(tk->h + (size_t)(pend[k].c + cc) * tk->ldh + pend[k].r0, pb + (size_t)cc * pend[k].nr,
               sizeof(double) * (size_t)pend[k].nr);
    }
  }
  return NULL;
}

static int io_run(qrdm_hostio *io, int dir, double *d, int ldd, double *h, int ldh, int m, int c0, int n) {
  if (m <= 0 || n <= c0) return 0;
  io_task tk[IO_MAX_THREADS];
  pthread_t th[IO_MAX_THREADS];
  int started = 0, err = 0;
  for (int t = 0; t < io->nthreads; ++t) {
    tk[t].io = io; tk[t].t = t; tk[t].dir = dir; tk[t].d = d; tk[t].h = h;
    tk[t].ldd = ldd; tk[t].ldh = ldh; tk[t].m = m; tk[t].c0 = c0; tk[t].n = n; tk[t].err = 0;
    if (pthread_create(&th[t], NULL, io_worker, &tk[t]) != 0) { err = -1; break; }
    ++started;
  }
  for (int t = 0; t < started; ++t) {
    pthread_join(th[t], NULL);
    if (tk[t].err) err = -1;
  }
  return err;
}

int qrdm_hostio_upload(qrdm_hostio *io, double *d, int ldd, const double *h, int ldh, int m, int n) {
  return io_run(io, 0, d, ldd, (double *)h, ldh, m, 0, n);
}
int qrdm_hostio_download(qrdm_hostio *io, double *h, int ldh, const double *d, int ldd, int m, int c0, int n) {
  return io_run(io, 1, (double *)d, ldd, h, ldh, m, c0, n);
}

/* ---- streamed write-back of finished columns ---- */
static void *wb_drain(void *arg) {
  qrdm_hostio *io = (qrdm_hostio *)arg;
  if (qrdm_rt_set_device(io->device)) { io->wb_error = 1; }
  for (;;) {
    pthread_mutex_lock(&io->mu);
    while (io->job_tail == io->job_head && !io->wb_stop) pthread_cond_wait(&io->cv, &io->mu);
    if (io->job_tail == io->job_head && io->wb_stop) { pthread_mutex_unlock(&io->mu); break; }
    wb_job jb = io->jobs[io->job_tail % IO_RING_JOBS];
    pthread_mutex_unlock(&io->mu);
    if (!io->wb_error && qrdm_rt_event_sync(jb.ev)) io->wb_error = 1;
    if (!io->wb_error) {
      const double *src = (const double *)(io->ring + jb.off);
      for (int c = jb.c0; c < jb.c1; ++c)
        memcpy(io->wb_h + (size_t)c * io->wb_ldh, src + (size_t)(c - jb.c0) * io->wb_m, sizeof(double) * (size_t)io->wb_m);
    }
    pthread_mutex_lock(&io->mu);
    ++io->job_tail;
    pthread_cond_broadcast(&io->cv);
    pthread_mutex_unlock(&io->mu);
  }
  return NULL;
}

int qrdm_hostio_wb_begin(qrdm_hostio *io, double *h, int ldh, const double *d, int ldd, int m, void *copy_stream) {
  if (io->wb_active) return -1;
  /* only worth it when an iteration's 64 columns are a small part of the ring */
  if ((size_t)m * sizeof(double) * 64 > IO_RING_BYTES / 4) return 1;
  if (!io->ring) {
    if (qrdm_rt_host_alloc((void **)&io->ring, IO_RING_BYTES)) { io->ring = NULL; return -1; }
    for (int i = 0; i < IO_RING_JOBS; ++i)
      if (qrdm_rt_event_create(&io->ring_ev[i])) return -1;
  }
  io->wb_h = h; io->wb_d = d; io->wb_ldh = ldh; io->wb_ldd = ldd; io->wb_m = m; io->wb_stream = copy_stream;
  io->job_head = io->job_tail = 0;
  io->wb_stop = 0; io->wb_error = 0;
  if (pthread_create(&io->wb_thread, NULL, wb_drain, io) != 0) return -1;
  io->wb_active = 1;
  return 0;
}

/* Columns [c0, c1) are final on the device (the caller has synchronised with the kernels that wrote them).
 * Returns the number of columns taken (a prefix of the range; 0 if the ring is full right now — never blocks
 * the factorisation loop), or < 0 on error. */
int qrdm_hostio_wb_push(qrdm_hostio *io, int c0, int c1) {
  if (!io->wb_active || c1 <= c0) return 0;
  const size_t colb = (size_t)io->wb_m * sizeof(double);
  const size_t slotb = colb * 64; /* the ring is cut into equal slots of 64 columns; job q lives in slot q % nslots */
  int nslots = (int)(IO_RING_BYTES / slotb);
  if (nslots > IO_RING_JOBS) nslots = IO_RING_JOBS;
  int taken = 0;
  while (c0 < c1) {
    pthread_mutex_lock(&io->mu);
    const int in_flight = io->job_head - io->job_tail;
    const int q = io->job_head;
    pthread_mutex_unlock(&io->mu);
    if (in_flight >= nslots) break; /* ring full right now: the rest goes with a later push or the final download */
    const int nc = c1 - c0 < 64 ? c1 - c0 : 64;
    wb_job *jb = &io->jobs[q % IO_RING_JOBS];
    jb->off = (size_t)(q % nslots) * slotb; jb->bytes = (size_t)nc * colb; jb->c0 = c0; jb->c1 = c0 + nc;
    jb->ev = io->ring_ev[q % IO_RING_JOBS];
    if (qrdm_rt_d2h_2d(io->ring + jb->off, colb, io->wb_d + (size_t)c0 * io->wb_ldd, sizeof(double) * (size_t)io->wb_ldd, colb,
                       (size_t)nc, io->wb_stream) ||
        qrdm_rt_event_record(jb->ev, io->wb_stream)) {
      io->wb_error = 1;
      return -1;
    }
    pthread_mutex_lock(&io->mu);
    ++io->job_head;
    pthread_cond_broadcast(&io->cv);
    pthread_mutex_unlock(&io->mu);
    c0 += nc;
    taken += nc;
  }
  return taken;
}

int qrdm_hostio_wb_end(qrdm_hostio *io) {
  if (!io->wb_active) return 0;
  pthread_mutex_lock(&io->mu);
  io->wb_stop = 1;
  pthread_cond_broadcast(&io->cv);
  pthread_mutex_unlock(&io->mu);
  pthread_join(io->wb_thread, NULL);
  io->wb_active = 0;
  return io->wb_error ? -1 : 0;
}
