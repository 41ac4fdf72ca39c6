/* TEST INFRASTRUCTURE ONLY — see qrdm_port.h.
 *
 * Plain-C, 0-based, BLAS-free restatement of the reference algorithm, written from the
 * behavioural spec in SURVEY.md §3.2.  Each function cites the reference lines it restates.
 * It additionally records how close every data-dependent decision was to flipping
 * ("margins"), which the parity tests use to apply the north-star rule "pivots and rank must
 * match wherever the reference's decisions are separated by more than 1e-12".
 */
#include "qrdm_port.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define A_(r, c) a[(size_t)(c) * (size_t)lda + (size_t)(r)]

typedef struct {
  double val;
  int idx;
} keyed_t;

static inline int imin(int x, int y) { return x < y ? x : y; }
static inline void note(double *slot, double v) {
  if (slot && v < *slot) *slot = v;
}

/* Euclidean norm as cblas_dnrm2 delivers it on x86-64 OpenBLAS (x87 extended-precision sum of
 * squares, one final sqrt): call sites src/dgeqrdm_work.c:96,673 and src/dlarfg.c:126,166. */
static double nrm2(int len, const double *x) {
  long double s = 0.0L;
  for (int i = 0; i < len; ++i) s += (long double)x[i] * (long double)x[i];
  return (double)sqrtl(s);
}

/* Stable descending merge sort = what glibc qsort + cmpStruct gives
 * (src/dgeqrdm_work.c:126-142, 345): ties keep ascending column index. */
static void sort_desc_stable(keyed_t *v, keyed_t *tmp, int len) {
  if (len < 2) return;
  int h = len / 2;
  sort_desc_stable(v, tmp, h);
  sort_desc_stable(v + h, tmp, len - h);
  int i = 0, j = h, k = 0;
  while (i < h && j < len) tmp[k++] = (v[j].val > v[i].val) ? v[j++] : v[i++];
  while (i < h) tmp[k++] = v[i++];
  while (j < len) tmp[k++] = v[j++];
  memcpy(v, tmp, (size_t)len * sizeof(keyed_t));
}

/* One exchange of permute_marked: flags, jpvt, vn1 and the FULL-height columns; vn2 is
 * deliberately not exchanged (src/dgeqrdm_work.c:166-183, 209-226). */
static void exchange(int m, double *a, int lda, int j0, int p, int q, int *marked, int *jpvt,
                     double *vn1) {
  int ti = marked[p]; marked[p] = marked[q]; marked[q] = ti;
  ti = jpvt[j0 + p]; jpvt[j0 + p] = jpvt[j0 + q]; jpvt[j0 + q] = ti;
  double td = vn1[j0 + p]; vn1[j0 + p] = vn1[j0 + q]; vn1[j0 + q] = td;
  double *cp = &A_(0, j0 + p), *cq = &A_(0, j0 + q);
  for (int r = 0; r < m; ++r) { td = cp[r]; cp[r] = cq[r]; cq[r] = td; }
}

/* permute_marked with nz = 0 (src/dgeqrdm_work.c:149-262). */
static void move_selected_to_front(int m, double *a, int lda, int j0, int cols, const int *sel,
                                   int fjb, int *marked, int *jpvt, double *vn1) {
  int jb = 0, jt = cols - 1;
  for (int s = 0; s < fjb; ++s) {
    int jc = sel[s];
    while (jb < jt && marked[jt] == 1) {
      exchange(m, a, lda, j0, jt, jb, marked, jpvt, vn1);
      while (jb < cols && marked[jb] == 1) ++jb;
      /* the reference's companion loop on jt only moves past flag value 2, which never occurs */
    }
    if (marked[jc] == 1) {
      while (jb < cols && marked[jb] == 1) ++jb;
      if (jc <= jb || jc < fjb) continue;
      if (marked[jb] == 0) {
        exchange(m, a, lda, j0, jc, jb, marked, jpvt, vn1);
        ++jb;
      }
    }
  }
}

/* DM_perm (src/dgeqrdm_work.c:280-418): order, norm filter, cosine Gram, greedy pick. */
static int dm_select(int m, int j0, int cols, int kmax, double *a, int lda, const double *vn1,
                     double tau_, double delta, keyed_t *keys, keyed_t *ktmp, int *marked, int *sel,
                     double *gram, double *mg) {
  int rows = m - j0;
  for (int c = 0; c < cols; ++c) { keys[c].val = vn1[j0 + c]; keys[c].idx = c; marked[c] = 0; }
  sort_desc_stable(keys, ktmp, cols);

  double top = keys[0].val, thr = tau_ * top;
  int nc = 0;
  while (nc < cols && nc < kmax && keys[nc].val > thr) ++nc; /* :348-351 */
  if (mg) {
    double sc = top > 0 ? top : 1.0;
    int lim = imin(cols, kmax + 1);
    for (int t = 0; t < lim; ++t) note(&mg[QRDM_PORT_M_FILTER], fabs(keys[t].val - thr) / sc);
    int upto = imin(nc, cols - 1); /* order among candidates and vs the first non-candidate */
    for (int t = 0; t < upto; ++t) note(&mg[QRDM_PORT_M_ORDER], (keys[t].val - keys[t + 1].val) / sc);
  }

  int fjb = 1; /* the largest column is always taken (:354-359) */
  sel[0] = keys[0].idx;
  marked[keys[0].idx] = 1;
  if (nc > 1) {
    /* cos(s,t) = (A_s/vn1_s)'(A_t/vn1_t) over the remaining rows (:365-379) */
    for (int t = 0; t < nc; ++t) {
      const double *ct = &A_(j0, j0 + keys[t].idx);
      double it = 1.0 / keys[t].val;
      for (int s = 0; s <= t; ++s) {
        const double *cs = &A_(j0, j0 + keys[s].idx);
        double is = 1.0 / keys[s].val, acc = 0.0;
        for (int r = 0; r < rows; ++r) acc += (is * cs[r]) * (it * ct[r]);
        gram[(size_t)t * nc + s] = acc; /* column t, row s: upper triangle, column-major */
      }
    }
    int *pos = (int *)ktmp; /* sorted positions of accepted candidates */
    pos[0] = 0;
    for (int t = 1; t < nc; ++t) { /* greedy (:382-403) */
      if (marked[keys[t].idx]) continue;
      double worst = 0.0;
      for (int s = 0; s < fjb; ++s) {
        double cs = fabs(gram[(size_t)t * nc + pos[s]]);
        if (worst < cs) worst = cs;
      }
      if (fjb < kmax) note(mg ? &mg[QRDM_PORT_M_COSINE] : NULL, fabs(worst - delta));
      if (worst < delta && fjb < kmax) {
        pos[fjb] = t;
        sel[fjb] = keys[t].idx;
        marked[keys[t].idx] = 1;
        ++fjb;
      }
    }
  }
  return fjb;
}

/* dgeqr2_mia + dlarfg_mia + dlarf_ (src/dgeqr2.c:148-191, src/dlarfg.c:120-185,
 * src/dlarf.c:133-185): unblocked Householder panel with the DM early stop.
 * Returns the number of columns actually triangularised (>= 1). */
static int panel_factor(int rows, int fjb, double *p, int lda, double *tau, double tau_, double *w,
                        double *mg) {
  const double safmin = DBL_MIN / (DBL_EPSILON * 0.5); /* dlamch('S')/dlamch('E') */
  double thres = 5e-14;                                 /* src/dgeqr2.c:40 */
  int k = imin(rows, fjb);
#define P_(r, c) p[(size_t)(c) * (size_t)lda + (size_t)(r)]
  for (int i = 0; i < k; ++i) {
    int len = rows - i; /* order of the reflector */
    double *alpha = &P_(i, i), *x = &P_(imin(i + 1, rows - 1), i);
    if (len <= 1) {
      tau[i] = 0.0; /* src/dlarfg.c:120-123 */
    } else {
      double xnorm = nrm2(len - 1, x);
      if (i > 0) {
        double sc = xnorm > thres ? xnorm : thres;
        note(mg ? &mg[QRDM_PORT_M_PANEL] : NULL, sc > 0 ? fabs(xnorm - thres) / sc : 0.0);
        if (xnorm < thres) return i; /* src/dlarfg.c:129-133, src/dgeqr2.c:165-169 */
      }
      if (xnorm == 0.0) {
        tau[i] = 0.0;
      } else {
        double beta = (*alpha >= 0.0) ? -hypot(*alpha, xnorm) : hypot(*alpha, xnorm); /* d_sign, src/dlarfg.c:21-26 */
        int knt = 0;
        if (fabs(beta) < safmin) { /* src/dlarfg.c:148-168 */
          double rs = 1.0 / safmin;
          do {
            ++knt;
            for (int r = 0; r < len - 1; ++r) x[r] *= rs;
            beta *= rs;
            *alpha *= rs;
          } while (fabs(beta) < safmin);
          xnorm = nrm2(len - 1, x);
          beta = (*alpha >= 0.0) ? -hypot(*alpha, xnorm) : hypot(*alpha, xnorm);
        }
        tau[i] = (beta - *alpha) / beta;
        double sc = 1.0 / (*alpha - beta);
        for (int r = 0; r < len - 1; ++r) x[r] *= sc;
        for (int q = 0; q < knt; ++q) beta *= safmin;
        *alpha = beta;
      }
    }
    if (i < fjb - 1) {
      double aii = P_(i, i);
      if (i == 0 && tau_ > 0.0) thres = tau_ * fabs(aii); /* src/dgeqr2.c:176-177 */
      if (tau[i] != 0.0) {
        P_(i, i) = 1.0;
        int nc = fjb - i - 1;
        for (int c = 0; c < nc; ++c) { /* w = C'v */
          const double *cc = &P_(i, i + 1 + c);
          const double *v = &P_(i, i);
          double acc = 0.0;
          for (int r = 0; r < len; ++r) acc += cc[r] * v[r];
          w[c] = acc;
        }
        for (int c = 0; c < nc; ++c) { /* C -= tau v w' */
          double *cc = &P_(i, i + 1 + c);
          const double *v = &P_(i, i);
          double f = tau[i] * w[c];
          for (int r = 0; r < len; ++r) cc[r] -= v[r] * f;
        }
        P_(i, i) = aii;
      }
    }
  }
#undef P_
  return k;
}

/* T of the compact-WY form, forward/columnwise = LAPACK dlarft('F','C')
 * (call at src/dgeqrdm_work.c:751-754); t is k x k, leading dimension ldt. */
static void form_t(int rows, int k, const double *v, int lda, const double *tau, double *t, int ldt) {
#define V_(r, c) v[(size_t)(c) * (size_t)lda + (size_t)(r)]
  for (int i = 0; i < k; ++i) {
    if (tau[i] == 0.0) {
      for (int q = 0; q <= i; ++q) t[(size_t)i * ldt + q] = 0.0;
      continue;
    }
    for (int q = 0; q < i; ++q) { /* t(0:i,i) = -tau_i * V(i:,0:i)' v_i, with v_i(i)=1 */
      double acc = V_(i, q);
      for (int r = i + 1; r < rows; ++r) acc += V_(r, q) * V_(r, i);
      t[(size_t)i * ldt + q] = -tau[i] * acc;
    }
    for (int q = 0; q < i; ++q) { /* t(0:i,i) = T(0:i,0:i) * t(0:i,i) (upper triangular mult) */
      double acc = 0.0;
      for (int s = q; s < i; ++s) acc += t[(size_t)s * ldt + q] * t[(size_t)i * ldt + s];
      t[(size_t)i * ldt + q] = acc;
    }
    t[(size_t)i * ldt + i] = tau[i];
  }
#undef V_
}

static int has_nan(int r, int c, const double *x, int ld) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i)
      if (x[(size_t)j * ld + i] != x[(size_t)j * ld + i]) return 1;
  return 0;
}

/* C <- (I - V T V')' C = C - V (T' (V'C))   (LAPACKE_dlarfb_mia 'L','T','F','C',
 * src/dlarfb.c:40-151; call at src/dgeqrdm_work.c:762-767).  w: k doubles. */
static void apply_block_reflector(int rows, int nc, int k, const double *v, int lda, const double *t,
                                  int ldt, double *c, double *w) {
#define V_(r, q) v[(size_t)(q) * (size_t)lda + (size_t)(r)]
  for (int col = 0; col < nc; ++col) {
    double *cc = c + (size_t)col * lda;
    for (int q = 0; q < k; ++q) {
      double acc = cc[q];
      for (int r = q + 1; r < rows; ++r) acc += V_(r, q) * cc[r];
      w[q] = acc;
    }
    for (int q = k - 1; q >= 0; --q) { /* w = T' w, T upper */
      double acc = 0.0;
      for (int s = 0; s <= q; ++s) acc += t[(size_t)q * ldt + s] * w[s];
      w[q] = acc;
    }
    for (int q = 0; q < k; ++q) {
      double f = w[q];
      cc[q] -= f;
      for (int r = q + 1; r < rows; ++r) cc[r] -= V_(r, q) * f;
    }
  }
#undef V_
}

/* norm_update (src/dgeqrdm_work.c:36-122).  Returns the new max norm. */
static double downdate_norms(int m, int n, int j0, int k, const double *a, int lda, double *vn1,
                             double *vn2, double tol3z) {
  double maxnrm = 0.0;
  int r1 = j0 + k;
  for (int c = r1; c < n; ++c) {
    if (vn1[c] == 0.0) continue;
    double d = 0.0;
    for (int r = j0; r < r1; ++r) d += A_(r, c) * A_(r, c);
    double t = sqrt(fabs(d)) / vn1[c];
    t = (t + 1.0) * (1.0 - t);
    t = (0.0 >= t) ? 0.0 : t; /* the reference's max(0,t) macro: a NaN stays NaN */
    double q = vn1[c] / vn2[c];
    double t2 = t * (q * q);
    if (t2 <= tol3z) {
      if (m - r1 > 0) {
        vn1[c] = nrm2(m - r1, &A_(r1, c));
        vn2[c] = vn1[c];
      } else {
        vn1[c] = vn2[c] = 0.0;
      }
    } else {
      vn1[c] *= sqrt(t);
    }
    if (maxnrm < vn1[c]) maxnrm = vn1[c];
  }
  return maxnrm;
}

int qrdm_port_dgeqrdm(int matrix_layout, int m, int n, double *a, int lda, int *jpvt, double *tau,
                      int *ncols, double *thres, int nb, double *margins) {
  const double eps = DBL_EPSILON * 0.5, tol3z = sqrt(eps);
  double delta = thres[0], tau_ = thres[1], eta = 0.0;
  int stop_mode = 0;
  if (ncols[0] == 1) { stop_mode = 1; eta = eps * n; }
  else if (ncols[0] == 2) { stop_mode = 2; eta = eps * sqrt((double)n); }
  else if (ncols[0] == 3) { stop_mode = 3; eta = thres[2]; }

  /* argument checks, src/dgeqrdm_work.c:559-589 (all failures return -1 after xerbla) */
  int bad = 0;
  if (matrix_layout != 102) bad = 1; /* 101 passes the reference's check but is indexed
                                        column-major and fails downstream (-11): unsupported */
  else if (m <= 0) bad = 2;
  else if (n <= 0) bad = 3;
  else if (lda < (m > 1 ? m : 1)) bad = 5;
  else if (delta < 0.0 || delta > 1.0) bad = 9;
  else if (tau_ < 0.0 || tau_ > 1.0) bad = 9;
  else if (nb <= 0) bad = 10;
  if (bad) {
    fprintf(stderr, "qrdm_port: parameter %d to DGEQRDM had an illegal value\n", bad);
    return -1;
  }
  for (int c = 0; c < n; ++c)
    if (jpvt[c] != 0) {
      fprintf(stderr, "qrdm_port: fixed columns (jpvt != 0) are out of scope (SURVEY 2a)\n");
      return -6;
    }
  for (int c = 0; c < n; ++c) jpvt[c] = c + 1;

  int minmn = imin(m, n), info = 0;
  double *vn1 = (double *)malloc(sizeof(double) * 2 * (size_t)n), *vn2 = vn1 + n;
  keyed_t *keys = (keyed_t *)malloc(sizeof(keyed_t) * 2 * (size_t)n), *ktmp = keys + n;
  int *marked = (int *)malloc(sizeof(int) * ((size_t)n + nb));
  int *sel = marked + n;
  double *gram = (double *)malloc(sizeof(double) * ((size_t)nb * nb * 2 + 2 * (size_t)nb));
  double *tmat = gram + (size_t)nb * nb, *w = tmat + (size_t)nb * nb;
  if (margins)
    for (size_t q = 0; q < (size_t)n * QRDM_PORT_NKINDS; ++q) margins[q] = INFINITY;

  double maxnrm = 0.0; /* :672-682 */
  for (int c = 0; c < n; ++c) {
    vn1[c] = vn2[c] = nrm2(m, &A_(0, c));
    if (vn1[c] > maxnrm) maxnrm = vn1[c];
  }
  eta *= maxnrm;

  int j0 = 0, it = -1;
  while (j0 < minmn) { /* :694-787 */
    ++it;
    int rows = m - j0, cols = n - j0, kmax = imin(imin(nb, rows), cols);
    double *mg = margins ? margins + (size_t)it * QRDM_PORT_NKINDS : NULL;

    int fjb = dm_select(m, j0, cols, kmax, a, lda, vn1, tau_, delta, keys, ktmp, marked, sel, gram, mg);
    move_selected_to_front(m, a, lda, j0, cols, sel, fjb, marked, jpvt, vn1);

    int k = panel_factor(rows, fjb, &A_(j0, j0), lda, tau + j0, tau_, w, mg);
    ncols[it] = k;

    /* LAPACKE_dlarft's own NaN screen (tau -> -8, V -> -6), then the _mia screen of
     * C (-13), T (-11), V (-9): src/dlarfb.c:73-86 */
    if (has_nan(k, 1, tau + j0, k)) { info = -8; break; }
    if (has_nan(rows, k, &A_(j0, j0), lda)) { info = -6; break; }
    form_t(rows, k, &A_(j0, j0), lda, tau + j0, tmat, nb);
    int ncc = cols - fjb;
    if (has_nan(rows, ncc, &A_(j0, j0 + fjb), lda)) { info = -13; break; }
    apply_block_reflector(rows, ncc, k, &A_(j0, j0), lda, tmat, nb, &A_(j0, j0 + fjb), w);

    maxnrm = downdate_norms(m, n, j0, k, a, lda, vn1, vn2, tol3z);
    j0 += k;
    if (stop_mode) { /* :782-785 */
      double lhs = maxnrm * sqrt((double)(cols - k));
      note(mg ? &mg[QRDM_PORT_M_STOP] : NULL, eta > 0 ? fabs(lhs - eta) / eta : fabs(lhs));
      if (lhs <= eta) break;
    }
  }
  free(gram);
  free(marked);
  free(keys);
  free(vn1);
  return info;
}
