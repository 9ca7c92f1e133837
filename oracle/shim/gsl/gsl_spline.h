/* TEST INFRASTRUCTURE — stand-in for <gsl/gsl_spline.h> (GSL is not installed in this image;
 * the reference pins no GSL version: SLICER/CMakeLists.txt:2 `find_package(GSL REQUIRED)`).
 *
 * Restates the published algorithm of GSL's `gsl_interp_cspline` (natural cubic spline):
 * second derivatives c[0]=c[n-1]=0, interior c[] from the symmetric tridiagonal system
 *   diag_i = 2 (h_i + h_{i+1}),  offdiag_i = h_{i+1},  rhs_i = 3 (dy_{i+1}/h_{i+1} - dy_i/h_i)
 * solved by the LDL^T recurrence, and the evaluation
 *   b = dy/dx - dx (c_{i+1} + 2 c_i)/3,  d = (c_{i+1} - c_i)/(3 dx),
 *   y = y_i + t (b + t (c_i + t d)),  t = x - x_i,
 * with x_i found by bisection (largest i with x_i <= x, clipped to n-2).
 * Only the six entry points the reference calls are provided (slicer-v2.cpp:89-96,219,319-322;
 * densitymaps.cpp:23,84,95,97).  Parity with a real GSL build is UNPINNED at bit level.
 */
#ifndef SLICER_SHIM_GSL_SPLINE_H
#define SLICER_SHIM_GSL_SPLINE_H

#include <cstddef>
#include <vector>
#include <cmath>
#include <limits>

struct gsl_interp_type
{
  int kind;
};
static const gsl_interp_type shim_gsl_cspline_type = {1};
static const gsl_interp_type *const gsl_interp_cspline = &shim_gsl_cspline_type;

struct gsl_interp_accel
{
  size_t cache;
};

struct gsl_spline
{
  size_t size;
  std::vector<double> x, y, c;
};

inline gsl_interp_accel *gsl_interp_accel_alloc()
{
  gsl_interp_accel *a = new gsl_interp_accel;
  a->cache = 0;
  return a;
}
inline void gsl_interp_accel_free(gsl_interp_accel *a) { delete a; }

inline gsl_spline *gsl_spline_alloc(const gsl_interp_type *, size_t size)
{
  gsl_spline *s = new gsl_spline;
  s->size = size;
  s->x.resize(size);
  s->y.resize(size);
  s->c.assign(size, 0.0);
  return s;
}
inline void gsl_spline_free(gsl_spline *s) { delete s; }

inline int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size)
{
  s->size = size;
  for (size_t i = 0; i < size; i++)
  {
    s->x[i] = xa[i];
    s->y[i] = ya[i];
  }
  const size_t max_index = size - 1;
  const size_t sys_size = max_index - 1;
  s->c[0] = 0.0;
  s->c[max_index] = 0.0;
  if (size < 3)
    return 0;
  std::vector<double> g(sys_size), diag(sys_size), offdiag(sys_size);
  for (size_t i = 0; i < sys_size; i++)
  {
    const double h_i = xa[i + 1] - xa[i];
    const double h_ip1 = xa[i + 2] - xa[i + 1];
    const double ydiff_i = ya[i + 1] - ya[i];
    const double ydiff_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
    const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    offdiag[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
  }
  if (sys_size == 1)
  {
    s->c[1] = g[0] / diag[0];
    return 0;
  }
  /* symmetric tridiagonal solve, A = L D L^T */
  const size_t N = sys_size;
  std::vector<double> alpha(N), gamma(N), cc(N), z(N);
  alpha[0] = diag[0];
  gamma[0] = offdiag[0] / alpha[0];
  for (size_t i = 1; i + 1 < N; i++)
  {
    alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1];
    gamma[i] = offdiag[i] / alpha[i];
  }
  alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
  z[0] = g[0];
  for (size_t i = 1; i < N; i++)
    z[i] = g[i] - gamma[i - 1] * z[i - 1];
  for (size_t i = 0; i < N; i++)
    cc[i] = z[i] / alpha[i];
  s->c[N] = cc[N - 1];
  if (N >= 2)
  {
    for (size_t i = N - 2, j = 0; j <= N - 2; j++, i--)
      s->c[i + 1] = cc[i] - gamma[i] * s->c[i + 2];
  }
  return 0;
}

inline double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *)
{
  const size_t n = s->size;
  if (x < s->x[0] || x > s->x[n - 1])
    return std::numeric_limits<double>::quiet_NaN(); /* GSL: GSL_EDOM */
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1)
  {
    size_t mid = (hi + lo) / 2;
    if (s->x[mid] > x)
      hi = mid;
    else
      lo = mid;
  }
  const size_t index = lo;
  const double x_lo = s->x[index], x_hi = s->x[index + 1];
  const double dx = x_hi - x_lo;
  const double y_lo = s->y[index], y_hi = s->y[index + 1];
  const double dy = y_hi - y_lo;
  const double c_i = s->c[index], c_ip1 = s->c[index + 1];
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  const double delx = x - x_lo;
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}

#endif
