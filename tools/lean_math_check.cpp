// lean_math_check.cpp — error-budget check of the lean projection (slicer_b200/csrc/lean_math.h) against the reference's
// chain evaluated with this machine's libm (densitymaps.cpp:382-386, utilities.cpp:23-25).
//
//   g++ -O2 -mfma -ffp-contract=off -o lean_math_check tools/lean_math_check.cpp -lm
//   ./lean_math_check N seed fov_deg npix
//
// For N random particles inside and around the field it evaluates both chains; the MUFU seeds of the GPU are emulated by
// correctly rounded float reciprocals / reciprocal square roots perturbed by a random relative error of up to 2^-21.5
// (worse than the hardware's documented 2^-22.9 / 1 ulp).  Prints one JSON line:
//   accepted (reference), flagged, mismatches among the NOT flagged (decision or float bits; must be 0),
//   max |v_lean - v_ref| in units of 2^-53 (the guard assumes < 2^-47 = 64 units).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float g_seed_err = 0.f;
#define LEAN_SEED_ERR g_seed_err
#include "../slicer_b200/csrc/lean_math.h"

static uint64_t mix(uint64_t z)
{
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static double u01(uint64_t *s)
{
  *s += 0x9E3779B97F4A7C15ull;
  return (double)(mix(*s) >> 11) * 0x1p-53;
}

int main(int argc, char **argv)
{
  const long long N = argc > 1 ? atoll(argv[1]) : 10000000;
  uint64_t st = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
  const double fov_deg = argc > 3 ? atof(argv[3]) : 5.0;
  const int npix = argc > 4 ? atoi(argv[4]) : 2048;
  const double fov = (double)(float)fov_deg / 180. * M_PI; // data.cpp:29 parses fov with stof
  const double T = fov * (1. + 2. / npix) * 0.5;           // densitymaps.cpp:383
  LeanDev L;
  if (!lean_setup(&L, fov, T))
  {
    printf("{\"error\": \"field too wide for the series\"}\n");
    return 2;
  }
  long long acc_ref = 0, flagged = 0, mism = 0, rej_lean_only = 0;
  double max_dv = 0;
  const double tanT = tan(T);
  for (long long i = 0; i < N; i++)
  {
    // z log-uniform in [1e-3, 16); x, y within 1.15 field half-widths of the axis (clamped to the box)
    const double zd = exp(log(1e-3) + u01(&st) * (log(16.) - log(1e-3)));
    const float z = (float)zd;
    const double hw = zd * tanT * 1.15;
    double xd = 0.5 + (2 * u01(&st) - 1) * hw, yd = 0.5 + (2 * u01(&st) - 1) * hw;
    if (i % 7 == 0)
      xd = 0.5 + ((i & 8) ? 1 : -1) * zd * tan(T) / cos(atan((yd - 0.5) / zd)) * (1 + (u01(&st) - 0.5) * 1e-9); // hug the dec edge
    if (i % 11 == 0)
      yd = 0.5 + ((i & 16) ? 1 : -1) * zd * tanT * (1 + (u01(&st) - 0.5) * 1e-9); // hug the ra edge
    if (xd < 0 || xd > 1 || yd < 0 || yd > 1)
      continue;
    const float x = (float)xd, y = (float)yd;
    g_seed_err = (float)((2 * u01(&st) - 1) * 0x1p-21 * 0.7071);
    // ---- reference
    const double X = (double)x - 0.5, Y = (double)y - 0.5, Z = (double)z;
    const double d = sqrt(X * X + Y * Y + Z * Z);
    const double dec = asin(X / d), ra = atan2(Y, Z);
    const bool ok_ref = fabs(ra) <= T && fabs(dec) <= T;
    const double vx_ref = dec / fov + 0.5, vy_ref = ra / fov + 0.5;
    const float xs_ref = (float)vx_ref, ys_ref = (float)vy_ref;
    acc_ref += ok_ref;
    // ---- lean
    float xs = 0, ys = 0;
    const int c = lean_project(x, y, z, L, &xs, &ys);
    if (c == LEAN_FLAGGED)
    {
      flagged++;
      continue;
    }
    if (c == LEAN_REJECT)
    {
      if (ok_ref)
        mism++, rej_lean_only++;
      continue;
    }
    if (!ok_ref || memcmp(&xs, &xs_ref, 4) || memcmp(&ys, &ys_ref, 4))
      mism++;
    double s, t;
    lean_ratios(x, y, z, &s, &t);
    const double dvx = fabs(lean_series(s, L.cs, L.K, L.A) + 0.5 - vx_ref), dvy = fabs(lean_series(t, L.ct, L.K, L.A) + 0.5 - vy_ref);
    if (dvx > max_dv)
      max_dv = dvx;
    if (dvy > max_dv)
      max_dv = dvy;
  }
  printf("{\"n\": %lld, \"K\": %d, \"accepted_ref\": %lld, \"flagged\": %lld, \"mismatch_unflagged\": %lld, \"rejected_by_lean_only\": %lld, "
         "\"max_dv_units_2m53\": %.2f, \"eta_units_2m53\": %.0f}\n",
         N, L.K, acc_ref, flagged, mism, rej_lean_only, max_dv / 0x1p-53, LEAN_ETA / 0x1p-53);
  return mism ? 1 : 0;
}
