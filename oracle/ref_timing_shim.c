/* TEST / BENCH INFRASTRUCTURE ONLY — call-site timers for the per-stage CPU split of the reference
 * ("split via the call sites").  The reference's sources stay untouched: oracle/Makefile compiles
 * src/dgeqrdm_work.c of the timed variant with three more -D renames, so that its calls at
 *   src/dgeqrdm_work.c:735  dgeqr2_mia           (panel)
 *   src/dgeqrdm_work.c:751  LAPACKE_dlarft       (T factor)
 *   src/dgeqrdm_work.c:762  LAPACKE_dlarfb_mia   (trailing update)
 * land in the forwarding functions below.  DM_perm and norm_update are calls inside that translation unit and
 * cannot be separated this way: they are reported together as the remainder (total - the three). */
#include <time.h>

static double g_t[3];
static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void qt_reset(void) { g_t[0] = g_t[1] = g_t[2] = 0.0; }
void qt_get(double *out3) { out3[0] = g_t[0]; out3[1] = g_t[1]; out3[2] = g_t[2]; }

extern int dgeqr2_mia(void *m, void *n, void *a, void *lda, void *tau, void *work, int threschoice, void *thresnrm, void *info);
int qt_dgeqr2_mia(void *m, void *n, void *a, void *lda, void *tau, void *work, int threschoice, void *thresnrm, void *info) {
  const double t0 = now();
  const int r = dgeqr2_mia(m, n, a, lda, tau, work, threschoice, thresnrm, info);
  g_t[0] += now() - t0;
  return r;
}

extern int scipy_LAPACKE_dlarft(int matrix_layout, char direct, char storev, int n, int k, const double *v, int ldv,
                                const double *tau, double *t, int ldt);
int qt_dlarft(int matrix_layout, char direct, char storev, int n, int k, const double *v, int ldv, const double *tau,
              double *t, int ldt) {
  const double t0 = now();
  const int r = scipy_LAPACKE_dlarft(matrix_layout, direct, storev, n, k, v, ldv, tau, t, ldt);
  g_t[1] += now() - t0;
  return r;
}

extern int LAPACKE_dlarfb_mia(int matrix_layout, char side, char trans, char direct, char storev, int m, int n, int k,
                              const double *v, int ldv, const double *t, int ldt, double *c, int ldc);
int qt_dlarfb_mia(int matrix_layout, int side, int trans, int direct, int storev, int m, int n, int k, const double *v,
                  int ldv, const double *t, int ldt, double *c, int ldc) {
  const double t0 = now();
  const int r = LAPACKE_dlarfb_mia(matrix_layout, (char)side, (char)trans, (char)direct, (char)storev, m, n, k, v, ldv, t, ldt, c, ldc);
  g_t[2] += now() - t0;
  return r;
}
