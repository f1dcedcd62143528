/* halton_spline.h -- CPU restatement of the once-sampled noise table of the halton-spline mode. TEST INFRASTRUCTURE.
 *
 * Reference path (src/m3p2i_aip):
 *   planners/motion_planner/mppi.py:458-478     get_samples: knots [K, nu, n_knots] -> bspline per (sample, dimension)
 *   utils/mppi_utils.py:70-104                  generate_halton_samples / generate_gaussian_halton_samples
 *   utils/skill_utils.py:9-22                   bspline: si.splrep(linspace(0, m, m), cv, k=degree, s=0.5),
 *                                               si.splev(linspace(0, m, T), spl, ext=3)
 * Third-party algorithms restated here (neither is vendored in the reference):
 *   - ghalton 0.6.x GeneralizedHalton (pyproject.toml:15): digit-permuted radical inverse, dimension d uses the d-th
 *     prime as base and a permutation of its digits. The package's EA_PERMS table is not available offline; callers
 *     pass permutations (NULL = identity = the plain Halton sequence of mppi_utils.py:70-79, the reference's own
 *     use_ghalton=False branch).
 *   - scipy.interpolate.splrep = FITPACK curfit/fpcurf (P. Dierckx, "Curve and Surface Fitting with Splines", 1993;
 *     Dierckx, Computer Graphics and Image Processing 20 (1982) 171-184): knots added where the residuals are largest
 *     until the least-squares spline has fp <= s, then the smoothing parameter p with fp(p) = s by rational
 *     interpolation (fprati), at most 20 iterations, tolerance 0.001 s.
 * Pinned by tests/golden/halton_spline_T*.npz (written by tests/golden/make_halton_golden.py from the imported
 * reference functions with scipy) and by tests/test_halton_spline.py against scipy directly on random data. */
#ifndef ORACLE_HALTON_SPLINE_H
#define ORACLE_HALTON_SPLINE_H
#include <math.h>
#include <string.h>

#define HS_MAXM 32               /* data points (n_knots) per spline */
#define HS_MAXN (HS_MAXM + 8)    /* knots */

/* ---------------------------------------------------------------- quasi-random knots */
static inline int hs_nth_prime(int n) { /* n = 0 -> 2 (mppi_utils.py:50-68) */
  int count = 0, c = 1;
  for (;;) {
    ++c;
    int prime = 1;
    for (int j = 2; j * j <= c; ++j) if (c % j == 0) { prime = 0; break; }
    if (prime && count++ == n) return c;
  }
}

/* i-th point (i >= 1) of the van der Corput sequence in `base`, digits mapped through perm (or identity) */
static inline double hs_radical_inverse(unsigned long long i, int base, const unsigned short* perm) {
  double f = 1.0, r = 0.0;
  while (i > 0) {
    f /= (double)base;
    const int d = (int)(i % (unsigned long long)base);
    r += f * (double)(perm ? perm[d] : d);
    i /= (unsigned long long)base;
  }
  return r;
}

/* sqrt(2) * erfinv(2 u - 1) evaluated like torch does on float32 tensors (mppi_utils.py:99-103): the argument is
 * rounded to fp32 first; erfinv itself is computed in double and rounded (torch's fp32 kernel is within 1-2 ulp) */
static inline double hs_erfinv(double y) {
  if (y <= -1.0) return -INFINITY;
  if (y >= 1.0) return INFINITY;
  /* initial guess (Giles 2010), then Newton / Halley on erf */
  double w = -log((1.0 - y) * (1.0 + y)), x;
  if (w < 5.0) {
    w -= 2.5;
    x = 2.81022636e-08; x = 3.43273939e-07 + x * w; x = -3.5233877e-06 + x * w; x = -4.39150654e-06 + x * w;
    x = 0.00021858087 + x * w; x = -0.00125372503 + x * w; x = -0.00417768164 + x * w; x = 0.246640727 + x * w;
    x = 1.50140941 + x * w;
  } else {
    w = sqrt(w) - 3.0;
    x = -0.000200214257; x = 0.000100950558 + x * w; x = 0.00134934322 + x * w; x = -0.00367342844 + x * w;
    x = 0.00573950773 + x * w; x = -0.0076224613 + x * w; x = 0.00943887047 + x * w; x = 1.00167406 + x * w;
    x = 2.83297682 + x * w;
  }
  x *= y;
  for (int it = 0; it < 3; ++it) {
    const double e = erf(x) - y;
    x -= e / (1.1283791670955126 * exp(-x * x) - x * e);
  }
  return x;
}
static inline float hs_gaussian(double u) {
  const float uf = (float)u;                     /* torch.tensor(..., dtype=float32) */
  const float arg = 2.0f * uf - 1.0f;
  return 1.41421356237309515f * (float)hs_erfinv((double)arg);
}

/* ---------------------------------------------------------------- FITPACK curfit (iopt = 0, w = 1, xb = x[0], xe = x[m-1]) */
/* values of the k+1 B-splines of degree k that are non-zero at x, t[l] <= x < t[l+1] (fpbspl) */
static inline void hs_bspl(const double* t, int k, double x, int l, double* h) {
  double hh[6];
  h[0] = 1.0;
  for (int j = 1; j <= k; ++j) {
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
    for (int i = 0; i < j; ++i) {
      const int li = l + i + 1, lj = li - j;
      const double f = hh[i] / (t[li] - t[lj]);
      h[i] += f * (t[li] - x);
      h[i + 1] = f * (x - t[lj]);
    }
  }
}

/* least squares: min |A c - y|^2 over the rows (A, y) and, when nb > 0, the extra rows (B / p, 0). A is m x nc dense.
 * Givens QR with a non-negative diagonal (fpgivs); returns the trace of R in *trace (from the observation rows only
 * when want_trace_of_A). */
static inline void hs_givens(double piv, double* ww, double* c, double* s) {
  const double store = fabs(piv);
  double dd;
  if (store >= *ww) dd = store * sqrt(1.0 + (*ww / piv) * (*ww / piv));
  else dd = *ww * sqrt(1.0 + (piv / *ww) * (piv / *ww));
  *c = *ww / dd; *s = piv / dd; *ww = dd;
}

typedef struct {
  double R[HS_MAXM][HS_MAXM];   /* upper triangular, nc x nc */
  double z[HS_MAXM];            /* Q^T y */
} HsQR;

static inline void hs_qr_rotate_row(HsQR* q, int nc, double* row, double rhs, int first) {
  for (int j = first; j < nc; ++j) {
    const double piv = row[j];
    if (piv == 0.0) continue;
    double c, s;
    hs_givens(piv, &q->R[j][j], &c, &s);
    /* rotate the right-hand side and the rest of the row */
    const double zj = q->z[j];
    q->z[j] = c * zj + s * rhs;
    rhs = c * rhs - s * zj;
    for (int i = j + 1; i < nc; ++i) {
      const double rji = q->R[j][i], ri = row[i];
      q->R[j][i] = c * rji + s * ri;
      row[i] = c * ri - s * rji;
    }
  }
}

static inline void hs_backsolve(const HsQR* q, int nc, double* c) {
  for (int i = nc - 1; i >= 0; --i) {
    double s = q->z[i];
    for (int j = i + 1; j < nc; ++j) s -= q->R[i][j] * c[j];
    c[i] = s / q->R[i][i];
  }
}

/* observation matrix of the data on knots t (n knots, degree k): row i has k+1 entries starting at column col[i] */
static inline void hs_observe(const double* x, int m, const double* t, int n, int k, double A[][6], int* col) {
  const int nk1 = n - k - 1;
  int l = k;   /* t[l] <= x < t[l+1] */
  for (int i = 0; i < m; ++i) {
    while (x[i] >= t[l + 1] && l < nk1 - 1) ++l;
    hs_bspl(t, k, x[i], l, A[i]);
    col[i] = l - k;
  }
}

static inline double hs_lsq(const double* y, int m, int n, int k, double A[][6],
                            const int* col, const double B[][6], int nb, double p, double* c, double* res, double* trace) {
  const int nc = n - k - 1;
  HsQR q;
  memset(&q, 0, sizeof(q));
  double row[HS_MAXM];
  for (int i = 0; i < m; ++i) {
    memset(row, 0, sizeof(double) * nc);
    for (int j = 0; j <= k; ++j) row[col[i] + j] = A[i][j];
    hs_qr_rotate_row(&q, nc, row, y[i], col[i]);
  }
  if (trace) { *trace = 0.0; for (int i = 0; i < nc; ++i) *trace += q.R[i][i]; }
  for (int r = 0; r < nb; ++r) {
    memset(row, 0, sizeof(double) * nc);
    for (int j = 0; j <= k + 1; ++j) row[r + j] = B[r][j] / p;
    hs_qr_rotate_row(&q, nc, row, 0.0, r);
  }
  hs_backsolve(&q, nc, c);
  double fp = 0.0;
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int j = 0; j <= k; ++j) s += c[col[i] + j] * A[i][j];
    const double e = (s - y[i]) * (s - y[i]);
    if (res) res[i] = e;
    fp += e;
  }
  return fp;
}

/* fpdisc: jumps of the k-th derivative of the B-splines at the interior knots, scaled as FITPACK does */
static inline int hs_disc(const double* t, int n, int k, double B[][6]) {
  const int nrint = n - 2 * k - 1;
  const double fac = pow((t[n - k - 1] - t[k]) / (double)nrint, (double)k);
  for (int jj = 0; jj < nrint - 1; ++jj) {
    const int j = jj + k + 1;
    for (int ii = 0; ii < k + 2; ++ii) {
      const int i = jj + ii;
      double prod = 1.0;
      for (int s = 0; s < k + 2; ++s) if (i + s != j) prod *= t[j] - t[i + s];
      B[jj][ii] = (t[i + k + 1] - t[i]) / prod * fac;
    }
  }
  return nrint - 1;
}

static inline double hs_fprati(double* p1, double* f1, double p2, double f2, double* p3, double* f3) {
  double p;
  if (*p3 > 0.0) {
    const double h1 = *f1 * (f2 - *f3), h2 = f2 * (*f3 - *f1), h3 = *f3 * (*f1 - f2);
    p = -(*p1 * p2 * h3 + p2 * *p3 * h1 + *p3 * *p1 * h2) / (*p1 * h1 + p2 * h2 + *p3 * h3);
  } else {
    p = (*p1 * (*f1 - *f3) * f2 - p2 * (f2 - *f3) * *f1) / ((*f1 - f2) * *f3);
  }
  if (f2 < 0.0) { *p3 = p2; *f3 = f2; } else { *p1 = p2; *f1 = f2; }
  return p;
}

/* Smoothing spline of degree k through (x, y), smoothing factor s: knots t[0..n) and coefficients c[0..n-k-1).
 * Returns n. */
static inline int hs_curfit(const double* x, const double* y, int m, int k, double s, double* t, double* c) {
  const double tol = 0.001, acc = tol * s, con1 = 0.1, con9 = 0.9, con4 = 0.04;
  const int maxit = 20, k1 = k + 1, nmin = 2 * k1, nmax = m + k1, nest = nmax > 2 * k + 3 ? nmax : 2 * k + 3;
  const double xb = x[0], xe = x[m - 1];
  double A[HS_MAXM][6], B[HS_MAXM][6], res[HS_MAXM], fpint[HS_MAXN];
  int col[HS_MAXM], nrdata[HS_MAXN];
  int n = nmin, nplus = 0, first = 1;
  double fp = 0.0, fpold = 0.0, fp0 = 0.0, fpms = 0.0, trace = 0.0;
  for (int i = 0; i < k1; ++i) { t[i] = xb; t[n - 1 - i] = xe; }
  nrdata[0] = m - 2;
  for (int iter = 0; iter < m; ++iter) {
    const int nrint = n - nmin + 1, nk1 = n - k1;
    for (int i = 0; i < k1; ++i) { t[i] = xb; t[n - 1 - i] = xe; }
    hs_observe(x, m, t, n, k, A, col);
    fp = hs_lsq(y, m, n, k, A, col, B, 0, 1.0, c, res, &trace);
    if (n == nmin) fp0 = fp;
    fpms = fp - s;
    if (fabs(fpms) < acc) return n;
    if (fpms < 0.0) break;          /* the knots are accepted: part 2 */
    if (n == nmax || n == nest) return n;
    if (first) { nplus = 1; first = 0; }
    else {
      int npl1 = nplus * 2;
      if (fpold - fp > acc) npl1 = (int)((double)nplus * fpms / (fpold - fp));
      int a = npl1 > nplus / 2 ? npl1 : nplus / 2;
      if (a < 1) a = 1;
      nplus = nplus * 2 < a ? nplus * 2 : a;
    }
    fpold = fp;
    /* squared residuals per knot interval; a data point on a knot gives half to each side */
    {
      double fpart = 0.0;
      int i = 0, l = k + 1, neu = 0;
      for (int it = 0; it < m; ++it) {
        if (!(x[it] < t[l] || l + 1 > nk1)) { neu = 1; ++l; }
        const double term = res[it];
        fpart += term;
        if (neu) {
          const double store = term * 0.5;
          fpint[i++] = fpart - store;
          fpart = store;
          neu = 0;
        }
      }
      fpint[nrint - 1] = fpart;
    }
    int nri = nrint, to_interp = 0;
    for (int l = 0; l < nplus; ++l) {
      /* fpknot: new knot in the interval with the largest residual sum, on the data point in its middle */
      double fpmax = 0.0;
      int number = -1, maxpt = 0, maxbeg = 0, jbegin = 1;
      for (int j = 0; j < nri; ++j) {
        const int jpoint = nrdata[j];
        if (!(fpmax >= fpint[j] || jpoint == 0)) { fpmax = fpint[j]; number = j; maxpt = jpoint; maxbeg = jbegin; }
        jbegin += jpoint + 1;
      }
      if (number < 0) break;
      const int ihalf = maxpt / 2 + 1, nrx = maxbeg + ihalf - 1;   /* 0-based index of the data point */
      for (int j = nri; j > number + 1; --j) { fpint[j] = fpint[j - 1]; nrdata[j] = nrdata[j - 1]; }
      for (int j = n; j > number + 1 + k; --j) t[j] = t[j - 1];
      nrdata[number] = ihalf - 1;
      nrdata[number + 1] = maxpt - ihalf;
      fpint[number] = fpmax * (double)nrdata[number] / (double)maxpt;
      fpint[number + 1] = fpmax * (double)nrdata[number + 1] / (double)maxpt;
      t[number + 1 + k] = x[nrx];
      ++n; ++nri;
      if (n == nmax) { to_interp = 1; break; }
      if (n == nest) break;
    }
    if (to_interp) {
      /* knots as for interpolation: for even k midway between the data points */
      const int mk1 = m - k1;
      int i = k1, j = k / 2 + 1;
      for (int l = 0; l < mk1; ++l, ++i, ++j) t[i] = (k % 2 == 0) ? 0.5 * (x[j] + x[j - 1]) : x[j];
    }
  }
  if (n == nmin) return n;   /* the least-squares polynomial is smooth enough */
  /* part 2: smoothing parameter p with f(p) = fp(p) - s = 0 */
  {
    const int nk1 = n - k1;
    for (int i = 0; i < k1; ++i) { t[i] = xb; t[n - 1 - i] = xe; }
    hs_observe(x, m, t, n, k, A, col);
    const int nb = hs_disc(t, n, k, B);
    double p1 = 0.0, f1 = fp0 - s, p3 = -1.0, f3 = fpms, p = (double)nk1 / trace;
    int ich1 = 0, ich3 = 0;
    for (int iter = 1; iter <= maxit; ++iter) {
      fp = hs_lsq(y, m, n, k, A, col, B, nb, p, c, 0, 0);
      fpms = fp - s;
      if (fabs(fpms) < acc || iter == maxit) break;
      const double p2 = p, f2 = fpms;
      if (!ich3) {
        if (!(f2 - f3 > acc)) {
          p3 = p2; f3 = f2; p *= con4;
          if (p <= p1) p = p1 * con9 + p2 * con1;
          continue;
        }
        if (f2 < 0.0) ich3 = 1;
      }
      if (!ich1) {
        if (!(f1 - f2 > acc)) {
          p1 = p2; f1 = f2; p /= con4;
          if (p3 < 0.0) continue;
          if (p >= p3) p = p2 * con1 + p3 * con9;
          continue;
        }
        if (f2 > 0.0) ich1 = 1;
      }
      if (f2 >= f1 || f2 <= f3) break;   /* not monotone: FITPACK gives up with ier = 2 and keeps this spline */
      p = hs_fprati(&p1, &f1, p2, f2, &p3, &f3);
    }
  }
  return n;
}

/* splev with ext = 3: x clamped to the base interval */
static inline double hs_splev(const double* t, int n, const double* c, int k, double x) {
  const int nk1 = n - k - 1;
  if (x < t[k]) x = t[k];
  if (x > t[nk1]) x = t[nk1];
  int l = k;
  while (x >= t[l + 1] && l < nk1 - 1) ++l;
  double h[6];
  hs_bspl(t, k, x, l, h);
  double s = 0.0;
  for (int j = 0; j <= k; ++j) s += c[l - k + j] * h[j];
  return s;
}

/* skill_utils.py:9-22: smoothing spline through the m knot values, sampled at T points */
static inline void hs_bspline_samples(const float* cv, int m, int T, int degree, double smoothing, float* out) {
  double x[HS_MAXM], y[HS_MAXM], t[HS_MAXN + 2], c[HS_MAXN];
  for (int i = 0; i < m; ++i) { x[i] = m == 1 ? 0.0 : (double)m * (double)i / (double)(m - 1); y[i] = (double)cv[i]; }
  x[m - 1] = (double)m;
  const int n = hs_curfit(x, y, m, degree, smoothing, t, c);
  for (int j = 0; j < T; ++j) {
    double xx = T == 1 ? 0.0 : (double)m * (double)j / (double)(T - 1);
    if (j == T - 1) xx = (double)m;
    out[j] = (float)hs_splev(t, n, c, degree, xx);
  }
}

/* The noise table delta[K, T, nu] for GLOBAL samples offset .. offset + K (mppi.py:458-478): point i of the
 * (generalised) Halton sequence in n_knots * nu dimensions (sequence index i + 1), erfinv, splines.
 * perms: NULL or perm_stride entries per dimension (digit permutation of that dimension's base). */
static inline int hs_table(int K, int offset, int T, int nu, int knot_scale, int degree, double smoothing,
                           const unsigned short* perms, int perm_stride, float* out) {
  const int m = T / knot_scale, ndims = m * nu;
  if (m <= degree || m > HS_MAXM) return -1;
  int bases[HS_MAXM * 16];
  if (ndims > HS_MAXM * 16) return -1;
  for (int d = 0; d < ndims; ++d) bases[d] = hs_nth_prime(d);
#pragma omp parallel for schedule(static)
  for (int kk = 0; kk < K; ++kk) {
    float cv[HS_MAXM], smp[256];
    for (int j = 0; j < nu; ++j) {
      for (int q = 0; q < m; ++q) {
        const int d = j * m + q;   /* knot_points.view(K, nu, n_knots) */
        cv[q] = hs_gaussian(hs_radical_inverse((unsigned long long)(offset + kk) + 1ull, bases[d], perms ? perms + (size_t)d * perm_stride : 0));
      }
      hs_bspline_samples(cv, m, T, degree, smoothing, smp);
      for (int tt = 0; tt < T; ++tt) out[((size_t)kk * T + tt) * nu + j] = smp[tt];
    }
  }
  return 0;
}

#endif
