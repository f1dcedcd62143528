// halton_spline.cuh -- the once-sampled noise table of the halton-spline mode, built on the device.
//
// Replaces the host loop of the reference (mppi.py:458-478: K * nu calls of skill_utils.bspline -> scipy splrep /
// splev, 36 864 calls at K = 4096): one thread per (sample, action dimension) draws its n_knots points of the
// (generalised) Halton sequence (mppi_utils.py:70-97), maps them through sqrt(2) erfinv(2 u - 1) (mppi_utils.py:99-103),
// fits FITPACK's smoothing spline (skill_utils.py:9-22: k = degree, s = 0.5 on x = linspace(0, m, m)) and samples it at
// linspace(0, m, T) straight into the [T][nu][K] table the rollout kernel reads.
//
// The spline fit is FITPACK curfit / fpcurf for iopt = 0, unit weights (Dierckx 1982 / 1993): least-squares splines on
// a growing knot set (new knot on the data point in the middle of the knot interval with the largest residual sum, fpknot)
// until fp <= s, then the smoothing parameter p with fp(p) = s by rational interpolation (fprati; tolerance 0.001 s,
// at most 20 iterations). Everything in fp64 like FITPACK; the triangular factor is kept as a band of k + 2 columns
// (fpcurf's `g`). Arithmetic per spline: a few thousand fp64 operations; the whole table is a one-off of < 1 ms.
#pragma once
#include "common.cuh"

namespace m3 {
namespace hs {

constexpr int kMaxM = 16;          // data points per spline: T / knot_scale with T <= M3P2I_MAX_HORIZON, knot_scale >= 4
constexpr int kMaxN = kMaxM + 5;   // knots: at most m + degree + 1, degree <= 3
constexpr int kBand = 5;           // degree + 2 <= 5

struct Work {
  double x[kMaxM], y[kMaxM], t[kMaxN + 1], c[kMaxM];
  double A[kMaxM][4], B[kMaxM][kBand], R[kMaxM][kBand], z[kMaxM], res[kMaxM], fpint[kMaxN];
  int col[kMaxM], nrdata[kMaxN];
};

__device__ inline void bspl(const double* t, int k, double x, int l, double* h) {
  double hh[4];
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

// one row (entries row[0..len) for columns first..first+len) with right-hand side rhs rotated into the band R
__device__ inline void rotate_row(Work& w, int nc, int band, double* row, int len, double rhs, int first) {
  for (int j = first; j < nc && len > 0; ++j) {
    const double piv = row[0];
    if (piv != 0.0) {
      // fpgivs
      double& ww = w.R[j][0];
      const double store = fabs(piv);
      const double dd = store >= ww ? store * sqrt(1.0 + (ww / piv) * (ww / piv)) : ww * sqrt(1.0 + (piv / ww) * (piv / ww));
      const double cs = ww / dd, sn = piv / dd;
      ww = dd;
      const double zj = w.z[j];
      w.z[j] = cs * zj + sn * rhs;
      rhs = cs * rhs - sn * zj;
      for (int i = 1; i < band; ++i) {
        const double rji = w.R[j][i], ri = i < len ? row[i] : 0.0;
        w.R[j][i] = cs * rji + sn * ri;
        row[i] = cs * ri - sn * rji;   // fill-in stays inside the band
      }
      if (len < band) len = band;
    }
    // next column: drop the (now zero) leading entry
    for (int i = 0; i + 1 < len; ++i) row[i] = row[i + 1];
    --len;
    if (j + 1 + len > nc) len = nc - j - 1;
  }
}

__device__ inline void observe(Work& w, int m, int n, int k) {
  const int nk1 = n - k - 1;
  int l = k;
  for (int i = 0; i < m; ++i) {
    while (w.x[i] >= w.t[l + 1] && l < nk1 - 1) ++l;
    bspl(w.t, k, w.x[i], l, w.A[i]);
    w.col[i] = l - k;
  }
}

// least squares over the observation rows and nb rows of B / p; coefficients in w.c; returns fp = sum of squared residuals
__device__ inline double lsq(Work& w, int m, int n, int k, int nb, double p, bool want_res, double* trace) {
  const int nc = n - k - 1, band = k + 2;
  for (int i = 0; i < nc; ++i) {
    w.z[i] = 0.0;
    for (int j = 0; j < kBand; ++j) w.R[i][j] = 0.0;
  }
  double row[kBand + 1];
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j <= k; ++j) row[j] = w.A[i][j];
    for (int j = k + 1; j <= kBand; ++j) row[j] = 0.0;
    rotate_row(w, nc, band, row, k + 1, w.y[i], w.col[i]);
  }
  if (trace) { *trace = 0.0; for (int i = 0; i < nc; ++i) *trace += w.R[i][0]; }
  for (int r = 0; r < nb; ++r) {
    for (int j = 0; j <= k + 1; ++j) row[j] = w.B[r][j] / p;
    rotate_row(w, nc, band, row, k + 2, 0.0, r);
  }
  for (int i = nc - 1; i >= 0; --i) {
    double s = w.z[i];
    for (int j = 1; j < band && i + j < nc; ++j) s -= w.R[i][j] * w.c[i + j];
    w.c[i] = s / w.R[i][0];
  }
  double fp = 0.0;
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int j = 0; j <= k; ++j) s += w.c[w.col[i] + j] * w.A[i][j];
    const double e = (s - w.y[i]) * (s - w.y[i]);
    if (want_res) w.res[i] = e;
    fp += e;
  }
  return fp;
}

// fpdisc
__device__ inline int disc(Work& w, int n, int k) {
  const int nrint = n - 2 * k - 1;
  const double fac = pow((w.t[n - k - 1] - w.t[k]) / (double)nrint, (double)k);
  for (int jj = 0; jj < nrint - 1; ++jj) {
    const int j = jj + k + 1;
    for (int ii = 0; ii < k + 2; ++ii) {
      const int i = jj + ii;
      double prod = 1.0;
      for (int s = 0; s < k + 2; ++s) if (i + s != j) prod *= w.t[j] - w.t[i + s];
      w.B[jj][ii] = (w.t[i + k + 1] - w.t[i]) / prod * fac;
    }
  }
  return nrint - 1;
}

// fpcurf, iopt = 0: knots in w.t, coefficients in w.c; returns the number of knots
__device__ inline int curfit(Work& w, int m, int k, double s) {
  const double acc = 0.001 * s, con1 = 0.1, con9 = 0.9, con4 = 0.04;
  const int maxit = 20, k1 = k + 1, nmin = 2 * k1, nmax = m + k1, nest = nmax > 2 * k + 3 ? nmax : 2 * k + 3;
  const double xb = w.x[0], xe = w.x[m - 1];
  int n = nmin, nplus = 0;
  bool first = true;
  double fp = 0.0, fpold = 0.0, fp0 = 0.0, fpms = 0.0, trace = 0.0;
  w.nrdata[0] = m - 2;
  for (int iter = 0; iter < m; ++iter) {
    const int nrint = n - nmin + 1, nk1 = n - k1;
    for (int i = 0; i < k1; ++i) { w.t[i] = xb; w.t[n - 1 - i] = xe; }
    observe(w, m, n, k);
    fp = lsq(w, m, n, k, 0, 1.0, true, &trace);
    if (n == nmin) fp0 = fp;
    fpms = fp - s;
    if (fabs(fpms) < acc) return n;
    if (fpms < 0.0) break;
    if (n == nmax || n == nest) return n;
    if (first) { nplus = 1; first = false; }
    else {
      int npl1 = nplus * 2;
      if (fpold - fp > acc) npl1 = (int)((double)nplus * fpms / (fpold - fp));
      nplus = min(nplus * 2, max(max(npl1, nplus / 2), 1));
    }
    fpold = fp;
    {
      double fpart = 0.0;
      int i = 0, l = k + 1;
      bool neu = false;
      for (int it = 0; it < m; ++it) {
        if (!(w.x[it] < w.t[l] || l + 1 > nk1)) { neu = true; ++l; }
        const double term = w.res[it];
        fpart += term;
        if (neu) {
          const double store = term * 0.5;
          w.fpint[i++] = fpart - store;
          fpart = store;
          neu = false;
        }
      }
      w.fpint[nrint - 1] = fpart;
    }
    int nri = nrint;
    bool to_interp = false;
    for (int l = 0; l < nplus; ++l) {
      // fpknot
      double fpmax = 0.0;
      int number = -1, maxpt = 0, maxbeg = 0, jbegin = 1;
      for (int j = 0; j < nri; ++j) {
        const int jpoint = w.nrdata[j];
        if (!(fpmax >= w.fpint[j] || jpoint == 0)) { fpmax = w.fpint[j]; number = j; maxpt = jpoint; maxbeg = jbegin; }
        jbegin += jpoint + 1;
      }
      if (number < 0) break;
      const int ihalf = maxpt / 2 + 1, nrx = maxbeg + ihalf - 1;
      for (int j = nri; j > number + 1; --j) { w.fpint[j] = w.fpint[j - 1]; w.nrdata[j] = w.nrdata[j - 1]; }
      for (int j = n; j > number + 1 + k; --j) w.t[j] = w.t[j - 1];
      w.nrdata[number] = ihalf - 1;
      w.nrdata[number + 1] = maxpt - ihalf;
      w.fpint[number] = fpmax * (double)w.nrdata[number] / (double)maxpt;
      w.fpint[number + 1] = fpmax * (double)w.nrdata[number + 1] / (double)maxpt;
      w.t[number + 1 + k] = w.x[nrx];
      ++n; ++nri;
      if (n == nmax) { to_interp = true; break; }
      if (n == nest) break;
    }
    if (to_interp) {
      const int mk1 = m - k1;
      int i = k1, j = k / 2 + 1;
      for (int l = 0; l < mk1; ++l, ++i, ++j) w.t[i] = (k % 2 == 0) ? 0.5 * (w.x[j] + w.x[j - 1]) : w.x[j];
    }
  }
  if (n == nmin) return n;
  const int nk1 = n - k1;
  for (int i = 0; i < k1; ++i) { w.t[i] = xb; w.t[n - 1 - i] = xe; }
  observe(w, m, n, k);
  const int nb = disc(w, n, k);
  double p1 = 0.0, f1 = fp0 - s, p3 = -1.0, f3 = fpms, p = (double)nk1 / trace;
  bool ich1 = false, ich3 = false;
  for (int iter = 1; iter <= maxit; ++iter) {
    fp = lsq(w, m, n, k, nb, p, false, nullptr);
    fpms = fp - s;
    if (fabs(fpms) < acc || iter == maxit) break;
    const double p2 = p, f2 = fpms;
    if (!ich3) {
      if (!(f2 - f3 > acc)) {
        p3 = p2; f3 = f2; p *= con4;
        if (p <= p1) p = p1 * con9 + p2 * con1;
        continue;
      }
      if (f2 < 0.0) ich3 = true;
    }
    if (!ich1) {
      if (!(f1 - f2 > acc)) {
        p1 = p2; f1 = f2; p /= con4;
        if (p3 < 0.0) continue;
        if (p >= p3) p = p2 * con1 + p3 * con9;
        continue;
      }
      if (f2 > 0.0) ich1 = true;
    }
    if (f2 >= f1 || f2 <= f3) break;
    // fprati
    if (p3 > 0.0) {
      const double h1 = f1 * (f2 - f3), h2 = f2 * (f3 - f1), h3 = f3 * (f1 - f2);
      p = -(p1 * p2 * h3 + p2 * p3 * h1 + p3 * p1 * h2) / (p1 * h1 + p2 * h2 + p3 * h3);
    } else {
      p = (p1 * (f1 - f3) * f2 - p2 * (f2 - f3) * f1) / ((f1 - f2) * f3);
    }
    if (f2 < 0.0) { p3 = p2; f3 = f2; } else { p1 = p2; f1 = f2; }
  }
  return n;
}

__device__ inline double splev(const Work& w, int n, int k, double x) {
  const int nk1 = n - k - 1;
  x = fmin(fmax(x, w.t[k]), w.t[nk1]);   // ext = 3
  int l = k;
  while (x >= w.t[l + 1] && l < nk1 - 1) ++l;
  double h[4];
  bspl(w.t, k, x, l, h);
  double s = 0.0;
  for (int j = 0; j <= k; ++j) s += w.c[l - k + j] * h[j];
  return s;
}

// i-th point (i >= 1) of the van der Corput sequence in `base`, digits mapped through perm (nullptr: identity)
__device__ inline double radical_inverse(unsigned long long i, int base, const unsigned short* perm) {
  double f = 1.0, r = 0.0;
  while (i > 0) {
    f /= (double)base;
    const int d = (int)(i % (unsigned long long)base);
    r += f * (double)(perm ? perm[d] : d);
    i /= (unsigned long long)base;
  }
  return r;
}

// sqrt(2) erfinv(2 u - 1) with the argument rounded to fp32 as the reference's float32 tensors are; erfinv in fp64
// (polynomial start + three Newton-Halley steps on erf) and rounded, within 1 - 2 ulp of torch.erfinv
__device__ inline float gaussian(double u) {
  const float uf = (float)u;
  const double y = (double)(2.0f * uf - 1.0f);
  double x;
  if (y <= -1.0) x = -INFINITY;
  else if (y >= 1.0) x = INFINITY;
  else {
    double ww = -log((1.0 - y) * (1.0 + y));
    if (ww < 5.0) {
      ww -= 2.5;
      x = 2.81022636e-08; x = 3.43273939e-07 + x * ww; x = -3.5233877e-06 + x * ww; x = -4.39150654e-06 + x * ww;
      x = 0.00021858087 + x * ww; x = -0.00125372503 + x * ww; x = -0.00417768164 + x * ww; x = 0.246640727 + x * ww;
      x = 1.50140941 + x * ww;
    } else {
      ww = sqrt(ww) - 3.0;
      x = -0.000200214257; x = 0.000100950558 + x * ww; x = 0.00134934322 + x * ww; x = -0.00367342844 + x * ww;
      x = 0.00573950773 + x * ww; x = -0.0076224613 + x * ww; x = 0.00943887047 + x * ww; x = 1.00167406 + x * ww;
      x = 2.83297682 + x * ww;
    }
    x *= y;
    for (int it = 0; it < 3; ++it) {
      const double e = erf(x) - y;
      x -= e / (1.1283791670955126 * exp(-x * x) - x * e);
    }
  }
  return 1.41421356237309515f * (float)x;
}

}  // namespace hs

// grid over (sample, dimension), sample fastest: thread (k, j) writes out[(t * nu + j) * K + k], t = 0..T
// (K = samples of this launch = row stride of `out`), for the GLOBAL sample offset + k
__global__ void k_halton_spline(float* __restrict__ out, int K, int offset, int T, int nu, int m, int degree, double smoothing,
                                const int* __restrict__ bases, const unsigned short* __restrict__ perms, int perm_stride) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * nu) return;
  const int k = idx % K, j = idx / K;
  hs::Work w;
  for (int q = 0; q < m; ++q) {
    const int d = j * m + q;   // knot_points.view(K, nu, n_knots), mppi.py:474
    w.x[q] = (double)m * (double)q / (double)(m - 1);
    w.y[q] = (double)hs::gaussian(hs::radical_inverse((unsigned long long)(offset + k) + 1ull, bases[d],
                                                      perms ? perms + (size_t)d * perm_stride : nullptr));
  }
  w.x[m - 1] = (double)m;
  const int n = hs::curfit(w, m, degree, smoothing);
  for (int t = 0; t < T; ++t) {
    const double xx = t == T - 1 ? (double)m : (double)m * (double)t / (double)(T - 1);
    out[(size_t)(t * nu + j) * K + k] = (float)hs::splev(w, n, degree, xx);
  }
}

}  // namespace m3
