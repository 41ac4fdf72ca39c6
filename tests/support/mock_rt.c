/* TEST INFRASTRUCTURE — mock of the qrdm_rt_* CUDA wrappers hostio.c uses, so that its tiling / threading /
 * ring logic runs on a CPU-only box.  "Device memory" is host memory; copies are DEFERRED: an async copy is only
 * queued on its stream and executed when somebody synchronises with an event recorded after it (or with the
 * stream).  A bounce buffer that is reused before its transfer was waited for therefore produces wrong data —
 * exactly the bug class the real asynchronous runtime would show. */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct op { void *dst; const void *src; size_t dpitch, spitch, width, height; struct op *next; long seq; } op;
typedef struct { pthread_mutex_t mu; op *head, *tail; long enq, done; } mstream;
typedef struct { mstream *s; long seq; } mevent;

static void run_until(mstream *s, long seq) {
  pthread_mutex_lock(&s->mu);
  while (s->head && s->head->seq <= seq) {
    op *o = s->head;
    for (size_t r = 0; r < o->height; ++r) memcpy((char *)o->dst + r * o->dpitch, (const char *)o->src + r * o->spitch, o->width);
    s->head = o->next;
    if (!s->head) s->tail = NULL;
    s->done = o->seq;
    free(o);
  }
  pthread_mutex_unlock(&s->mu);
}
static int enqueue(mstream *s, void *dst, size_t dp, const void *src, size_t sp, size_t w, size_t h) {
  op *o = (op *)calloc(1, sizeof(op));
  o->dst = dst; o->src = src; o->dpitch = dp; o->spitch = sp; o->width = w; o->height = h;
  pthread_mutex_lock(&s->mu);
  o->seq = ++s->enq;
  if (s->tail) s->tail->next = o; else s->head = o;
  s->tail = o;
  pthread_mutex_unlock(&s->mu);
  return 0;
}
int qrdm_rt_malloc(void **p, size_t b) { *p = malloc(b); return *p ? 0 : 2; }
int qrdm_rt_free(void *p) { free(p); return 0; }
int qrdm_rt_host_alloc(void **p, size_t b) { *p = malloc(b); return *p ? 0 : 2; }
int qrdm_rt_host_free(void *p) { free(p); return 0; }
int qrdm_rt_stream_create(void **s) {
  mstream *m = (mstream *)calloc(1, sizeof(mstream));
  pthread_mutex_init(&m->mu, NULL);
  *s = m;
  return 0;
}
int qrdm_rt_stream_destroy(void *s) { if (s) { run_until((mstream *)s, 1L << 60); free(s); } return 0; }
int qrdm_rt_event_create(void **e) { *e = calloc(1, sizeof(mevent)); return 0; }
int qrdm_rt_event_destroy(void *e) { free(e); return 0; }
int qrdm_rt_event_record(void *e, void *s) {
  mevent *ev = (mevent *)e; mstream *m = (mstream *)s;
  pthread_mutex_lock(&m->mu); ev->s = m; ev->seq = m->enq; pthread_mutex_unlock(&m->mu);
  return 0;
}
int qrdm_rt_event_sync(void *e) { mevent *ev = (mevent *)e; if (ev->s) run_until(ev->s, ev->seq); return 0; }
int qrdm_rt_sync(void *s) { run_until((mstream *)s, 1L << 60); return 0; }
int qrdm_rt_set_device(int d) { (void)d; return 0; }
int qrdm_rt_h2d_2d(void *dst, size_t dp, const void *src, size_t sp, size_t w, size_t h, void *s) { return enqueue((mstream *)s, dst, dp, src, sp, w, h); }
int qrdm_rt_d2h_2d(void *dst, size_t dp, const void *src, size_t sp, size_t w, size_t h, void *s) { return enqueue((mstream *)s, dst, dp, src, sp, w, h); }
