/* TEST INFRASTRUCTURE — CPU restatement of SLICER's light-cone mass-map hot path.
 *
 * This file is the parity ORACLE for the CUDA path.  It is not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it (oracle/oracle_bindings.py).
 * It restates, function by function, the arithmetic of the reference with the exact precision chain
 * (which operand is float, which is double, where values are narrowed), in plain C with
 * -ffp-contract=off so that, like the reference binary (no -march => no FMA), nothing is fused.
 *
 * PINNING: every function here is checked bit-for-bit against the reference's own compiled code
 * (oracle/_ref/libslicer_ref.so, built from the .cpp files under /root/reference/SLICER by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, and against the committed golden vectors in tests/golden/ that were
 * generated from that build (oracle/make_golden.py).  The reference ships no tests of its own.
 *
 * Reference citations are file:line into /root/reference/SLICER/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Randomised box transform — gadget2io.cpp:204-270 (inside readPos).
 * raw: AoS [n][3] float32 as stored in the POS block; outputs SoA x,y,z in box units, z piled by rcase.
 *   b = sgn * (raw / boxsize)      double division, narrowed to float on assignment (:204-206)
 *   wraps compare/adjust against double literals, narrowed (:209-220)
 *   axis permutation by face (:223-252), recentre with double x0 (:254-256), wraps (:258-269), z += rcase (:270)
 * ------------------------------------------------------------------------------------------- */
void orc_transform(const float *raw, long n, double boxsize, const int *sgn, int face, const double *centre,
                   float rcase, float *ox, float *oy, float *oz)
{
  for (long pp = 0; pp < n; pp++)
  {
    float num_float1 = raw[3 * pp + 0], num_float2 = raw[3 * pp + 1], num_float3 = raw[3 * pp + 2];
    float xb = (float)(sgn[0] * (num_float1 / boxsize));
    float yb = (float)(sgn[1] * (num_float2 / boxsize));
    float zb = (float)(sgn[2] * (num_float3 / boxsize));
    if (xb > 1.)
      xb = (float)(xb - 1.);
    if (yb > 1.)
      yb = (float)(yb - 1.);
    if (zb > 1.)
      zb = (float)(zb - 1.);
    if (xb < 0.)
      xb = (float)(1. + xb);
    if (yb < 0.)
      yb = (float)(1. + yb);
    if (zb < 0.)
      zb = (float)(1. + zb);
    float x = xb, y = yb, z = zb;
    switch (face)
    {
    case 1:
      break;
    case 2:
      x = xb, y = zb, z = yb;
      break;
    case 3:
      x = yb, y = zb, z = xb;
      break;
    case 4:
      x = yb, y = xb, z = zb;
      break;
    case 5:
      x = zb, y = xb, z = yb;
      break;
    case 6:
      x = zb, y = yb, z = xb;
      break;
    }
    x = (float)(x - centre[0]);
    y = (float)(y - centre[1]);
    z = (float)(z - centre[2]);
    if (x > 1.)
      x = (float)(x - 1.);
    if (y > 1.)
      y = (float)(y - 1.);
    if (z > 1.)
      z = (float)(z - 1.);
    if (x < 0.)
      x = (float)(1. + x);
    if (y < 0.)
      y = (float)(1. + y);
    if (z < 0.)
      z = (float)(1. + z);
    z += rcase;
    ox[pp] = x;
    oy[pp] = y;
    oz[pp] = z;
  }
}

/* getPolar(..., radec=true) — utilities.cpp:19-26 */
void orc_getpolar(double x, double y, double z, double *ra, double *dec, double *d)
{
  *d = sqrt(x * x + y * y + z * z);
  *dec = asin(x / (*d));
  *ra = atan2(y, z);
}

/* TSC kernel — utilities.cpp:4-16.  ixx, ixh float; dx double. */
float orc_weight(float ixx, float ixh, double dx)
{
  float DD = ixx - ixh;
  float x = (float)(fabsf(DD) / dx);
  float w;
  if (fabsf(DD) <= 0.5 * dx)
    w = (float)(3. / 4. - x * x);
  else if (fabsf(DD) > 0.5 * dx && fabsf(DD) <= 0.5 * 3.0 * dx)
    w = (float)(0.5 * ((3. / 2. - x) * (3. / 2. - x)));
  else
    w = 0.f;
  return w;
}

/* ---------------------------------------------------------------------------------------------
 * Shell selection + perpendicular replication + projection + FoV cut for ONE particle type —
 * densitymaps.cpp:346-402.
 *   mass: const_mass = float(massarr[i]) when per_particle == NULL (:372), else the file float with the
 *         MAX_M cut (:367-369; max_m <= 0 disables the cut, used to model the non-hydro branch)
 *   slab: float z against double minDist/maxDist (:346-347,374)
 *   replicas: x + ni is a FLOAT add (float + int), then - 0.5 in double (:382)
 *   accept: |ra|,|dec| <= fovradiants*(1+2/npix)*0.5 (:383);  xs = float(dec/fov+0.5), ys = float(ra/fov+0.5) (:385-386)
 *   degradation snopt>0: one libc rand() per accepted pair, in acceptance order (:393-396)
 * Outputs are appended at xs/ys/ms[0..cap); returns the number accepted (may exceed cap: nothing is
 * written past cap, the caller must retry with a larger buffer).
 * ------------------------------------------------------------------------------------------- */
long orc_select_project(const float *x, const float *y, const float *z, const float *per_particle, float const_mass,
                        double max_m, long n, double ld, double ld2, double boxsize, int nrepperp,
                        double fovradiants, int npix, int snopt, float *xs, float *ys, float *ms, long cap)
{
  const double POS_U = 1.0; /* gadget2io.h:14 */
  double minDist = ld / boxsize * 1.e+3 / POS_U;
  double maxDist = ld2 / boxsize * 1.e+3 / POS_U;
  long na = 0;
  for (long l = 0; l < n; l++)
  {
    float num_float1;
    if (per_particle)
    {
      num_float1 = per_particle[l];
      if (max_m > 0 && num_float1 > max_m)
        num_float1 = 0;
    }
    else
      num_float1 = const_mass;
    if (z[l] >= minDist && z[l] < maxDist)
    {
      for (int ni = -nrepperp; ni <= nrepperp; ni++)
        for (int nj = -nrepperp; nj <= nrepperp; nj++)
        {
          double rai, deci, dd;
          orc_getpolar((x[l] + ni) - 0.5, (y[l] + nj) - 0.5, z[l], &rai, &deci, &dd);
          if (fabs(rai) <= fovradiants * (1. + 2. / npix) * 0.5 && fabs(deci) <= fovradiants * (1. + 2. / npix) * 0.5)
          {
            float m;
            if (snopt == 0)
              m = num_float1;
            else
            {
              if (rand() / (float)RAND_MAX < 1. / pow(2, snopt))
                m = (float)(pow(2, snopt) * num_float1);
              else
                m = 0.f;
            }
            if (na < cap)
            {
              xs[na] = (float)(deci / fovradiants + 0.5);
              ys[na] = (float)(rai / fovradiants + 0.5);
              ms[na] = m;
            }
            na++;
          }
        }
    }
  }
  return na;
}

/* One TSC/NGP contribution list for a particle — utilities.cpp:66-94.
 * Returns the number of (cell, value) pairs written (<= 9); cells outside the grid are dropped
 * exactly as the reference drops them (:74,:91).  value is the float the reference adds. */
static int deposit_contribs(float px, float py, float w, int nn, int do_ngp, long *cell, float *val)
{
  double dl = 1. / (double)nn;
  int gx4 = (int)floor(px / dl);
  int gy4 = (int)floor(py / dl);
  int k = 0;
  if (do_ngp)
  {
    if (gx4 >= 0 && gx4 < nn && gy4 >= 0 && gy4 < nn)
    {
      cell[k] = gx4 + (long)nn * gy4;
      val[k++] = w;
    }
    return k;
  }
  for (int j = 0; j < 9; j++)
  {
    int gx = gx4 + (j % 3) - 1;
    int gy = gy4 + (j / 3) - 1;
    float posgridx = (float)((gx + 0.5) * dl);
    float posgridy = (float)((gy + 0.5) * dl);
    float wfx = sqrtf(w) * orc_weight(px, posgridx, dl);
    float wfy = sqrtf(w) * orc_weight(py, posgridy, dl);
    if (gx >= 0 && gx < nn && gy >= 0 && gy < nn)
    {
      cell[k] = gx + (long)nn * gy;
      val[k++] = wfx * wfy;
    }
  }
  return k;
}

/* gridist_w — utilities.cpp:36-97: float32 accumulation in particle order (bit-identical to the
 * reference's valarray<float>).  map must hold nn*nn floats and is ZEROED here, like `valarray<float> grxy(nn*nn)`. */
void orc_gridist_w(const float *x, const float *y, const float *w, long n, int nn, int do_ngp, float *map)
{
  memset(map, 0, sizeof(float) * (size_t)nn * nn);
  long cell[9];
  float val[9];
  for (long i = 0; i < n; i++)
  {
    int k = deposit_contribs(x[i], y[i], w[i], nn, do_ngp, cell, val);
    for (int j = 0; j < k; j++)
      map[cell[j]] = map[cell[j]] + val[j];
  }
}

/* Same float contributions, accumulated WITHOUT zeroing into double (the "exact sum of the same
 * contributions" that mass-conservation is judged against). */
void orc_gridist_w_f64(const float *x, const float *y, const float *w, long n, int nn, int do_ngp, double *map)
{
  long cell[9];
  float val[9];
  for (long i = 0; i < n; i++)
  {
    int k = deposit_contribs(x[i], y[i], w[i], nn, do_ngp, cell, val);
    for (int j = 0; j < k; j++)
      map[cell[j]] += (double)val[j];
  }
}

/* Same float contributions in the product's accumulator format: int64 fixed point,
 * q = llrint(value * 2^frac_bits) (round-to-nearest-even), summed exactly (order independent). */
void orc_gridist_w_fixed(const float *x, const float *y, const float *w, long n, int nn, int do_ngp, int frac_bits,
                         int64_t *map)
{
  long cell[9];
  float val[9];
  double scale = ldexp(1.0, frac_bits);
  for (long i = 0; i < n; i++)
  {
    int k = deposit_contribs(x[i], y[i], w[i], nn, do_ngp, cell, val);
    for (int j = 0; j < k; j++)
      map[cell[j]] += (int64_t)llrint((double)val[j] * scale);
  }
}

/* NGP cell index of one projected point, or -1 if outside the grid — utilities.cpp:69-76 */
long orc_ngp_cell(float px, float py, int nn)
{
  double dl = 1. / (double)nn;
  int gx = (int)floor(px / dl);
  int gy = (int)floor(py / dl);
  if (gx >= 0 && gx < nn && gy >= 0 && gy < nn)
    return gx + (long)nn * gy;
  return -1;
}

/* ---------------------------------------------------------------------------------------------
 * randomizeBox — densitymaps.cpp:166-248 (default build; fixed_vertex mirrors -DFixedPLCVertex :191-195).
 * Uses libc srand/rand like the reference (glibc TYPE_3 generator state is global).
 * ------------------------------------------------------------------------------------------- */
void orc_randomize_box(int seedcenter, int seedface, int seedsign, int nplanes, const int *randomize,
                       int lens_per_snap, int fixed_vertex, double *x0, double *y0, double *z0, int *face, int *sx,
                       int *sy, int *sz)
{
  for (int i = 0; i < nplanes; i++)
  {
    if (randomize[i])
    {
      srand(seedcenter + i / lens_per_snap * 13);
      if (!fixed_vertex)
      {
        x0[i] = rand() / (float)RAND_MAX;
        y0[i] = rand() / (float)RAND_MAX;
        z0[i] = rand() / (float)RAND_MAX;
      }
      else
      {
        x0[i] = 0.0;
        y0[i] = 0.0;
        z0[i] = 0.5;
      }
      face[i] = 7;
      srand(seedface + i / lens_per_snap * 5);
      while (face[i] > 6 || face[i] < 1)
        face[i] = (int)(1 + rand() / (float)RAND_MAX * 5. + 0.5);
      sx[i] = 2;
      srand(seedsign + i / lens_per_snap * 8);
      while (sx[i] > 1 || sx[i] < 0)
        sx[i] = (int)(rand() / (float)RAND_MAX + 0.5);
      sy[i] = 2;
      while (sy[i] > 1 || sy[i] < 0)
        sy[i] = (int)(rand() / (float)RAND_MAX + 0.5);
      sz[i] = 2;
      while (sz[i] > 1 || sz[i] < 0)
        sz[i] = (int)(rand() / (float)RAND_MAX + 0.5);
      if (sx[i] == 0)
        sx[i] = -1;
      if (sy[i] == 0)
        sy[i] = -1;
      if (sz[i] == 0)
        sz[i] = -1;
    }
    else
    {
      x0[i] = x0[i - 1];
      y0[i] = y0[i - 1];
      z0[i] = z0[i - 1];
      face[i] = face[i - 1];
      sx[i] = sx[i - 1];
      sy[i] = sy[i - 1];
      sz[i] = sz[i - 1];
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * w0waCDM::Hz / comovingDistance / transverseComovingDistance — w0waCDM.cpp:18-84, driven the way
 * main drives it (slicer-v2.cpp:79-86): n table points z_i = i (zs+1)/(n-1) in increasing order, wa = 0,
 * H0 = 100.  Reproduces the cache recurrence including its quirks: each point integrates
 * from the previous table point with dz = (z - lastZ)/100 and a `zi < z` loop, accumulating in `distance`
 * (unscaled, cached), and returns distance * CSPEEDOFLIGHT.
 * ------------------------------------------------------------------------------------------- */
static double hz(double z, double H0, double om, double ol, double w0, double wa)
{
  double rhoLambda = ol * pow(1 + z, 3 * (1 + w0 + wa)) * exp(-3 * wa * z / (1 + z));
  double rhoM = om * pow(1 + z, 3);
  double rhoTot = rhoLambda + rhoM + (1 - om - ol) * pow(1 + z, 2);
  return H0 * sqrt(rhoTot);
}

void orc_cosmo_table(double om, double ol, double w, double zs, int n, double *zl, double *dl)
{
  const double speedcunit = 2.99792458e+3; /* utilities.h:19 */
  const double CSPEEDOFLIGHT = speedcunit * 100;
  const double H0 = 100.0, wa = 0.0;
  /* the std::map cache holds (z -> unscaled distance); with strictly increasing queries the
   * `lower_bound(z)` predecessor is always the previous query */
  int have_prev = 0;
  double prev_z = 0, prev_d = 0;
  for (int i = 0; i < n; i++)
  {
    double z = i * (zs + 1.0) / (n - 1);
    zl[i] = z;
    double distance = 0, lastZ = 0, dz = 1e-4;
    if (have_prev && prev_z == z)
    { /* cache hit returns the unscaled value (w0waCDM.cpp:30-33); unreachable for increasing z */
      dl[i] = prev_d;
      continue;
    }
    if (have_prev)
    {
      distance = prev_d;
      lastZ = prev_z;
      dz = (z - lastZ) / 100;
    }
    for (double zi = lastZ; zi < z; zi += dz)
      distance += 0.5 * dz * (1.0 / hz(zi, H0, om, ol, w, wa) + 1.0 / hz(zi + dz, H0, om, ol, w, wa));
    prev_z = z;
    prev_d = distance;
    have_prev = 1;
    double D_C = distance * CSPEEDOFLIGHT;
    if (fabs(1 - om - ol) < 1e-5)
      dl[i] = D_C;
    else
    {
      double OmegaK = 1.0 - om - ol;
      double sqrtOmegaK = sqrt(fabs(OmegaK));
      if (OmegaK < 0)
        dl[i] = CSPEEDOFLIGHT / H0 / sqrtOmegaK * sinh(sqrtOmegaK * H0 / CSPEEDOFLIGHT * D_C);
      else
        dl[i] = CSPEEDOFLIGHT / H0 / sqrtOmegaK * sin(sqrtOmegaK * H0 / CSPEEDOFLIGHT * D_C);
    }
  }
}
