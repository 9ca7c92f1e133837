// plan.cpp — configuration and the light-cone plan (host, O(#planes)): readInput, readRedList, cosmology table,
// natural cubic spline, buildPlanes, randomizeBox, testFov, computeReplications.  The outputs are the parameters of
// the CUDA pass; they are reproduced with the reference's own expressions (including its numerical quirks, SURVEY.md
// App. D.7) so that a real InputParams.ini + snapshot list gives the reference's planes.
#include "slicer_host.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>

namespace slicer
{

static std::string sconv_int(int v)
{
  char b[64];
  snprintf(b, sizeof(b), "%i", v); // utilities.h:22 fINT
  return b;
}

// data.cpp:8-87 — 14 values on the even lines of a positional file; odd lines are labels and ignored
int readInput(InputParams &p, const std::string &name)
{
  std::ifstream fin(name.c_str());
  if (!fin.is_open())
  {
    std::cerr << " Params file " << name << " does not exist where you are running the code " << std::endl;
    std::cerr << " I will STOP here!!! " << std::endl;
    exit(1);
  }
  std::string str;
  auto value = [&](std::string &dst) {
    std::getline(fin, str);
    std::getline(fin, dst);
  };
  std::string v;
  value(v);
  p.npix = std::stoi(v); //  1. Number of Pixels
  value(v);
  p.zs = std::stof(v); //    2. Redshift Source   (float precision, data.cpp:26)
  value(v);
  p.fov = std::stof(v); //   3. Field of View     (float precision, data.cpp:29)
  value(p.filredshiftlist); // 4.
  value(p.pathsnap);        // 5.
  value(p.simulation);      // 6.
  value(v);
  p.seedcenter = std::stoi(v); // 7.
  value(v);
  p.seedface = std::stoi(v); //   8. (labelled "Pos. Reflec." but drives the axis permutation)
  value(v);
  p.seedsign = std::stoi(v); //   9. (labelled "Axis Sel." but drives the reflections)
  value(v);
  p.partinplanes = std::stoi(v) != 0; // 10.
  value(p.directory);                 // 11.
  value(p.suffix);                    // 12.
  value(v);
  p.snopt = std::stoi(v); // 13.
  value(v);
  p.w = std::stof(v); //     14.
  p.simType = p.npix == 0 ? "SubFind" : "Gadget";
  p.physical = p.npix < 0;
  if (!p.physical)
    p.snpix = sconv_int(p.npix);
  else
  {
    const int n = -p.npix;
    p.snpix = sconv_int(n) + "_kpc";
    p.rgrid = n;
  }
  if (p.snopt < 0)
  {
    std::cerr << "Impossible value for Shot-Noise option!" << std::endl;
    return 1;
  }
  return 0;
}

// gadget2io.cpp:613-661 — including its behaviour at the end of the list (the last name is re-used when the list
// runs out before a snapshot deeper than the source, App. C)
int readRedList(const std::string &filredshiftlist, std::vector<double> &snapred, std::vector<std::string> &snappath,
                std::vector<double> &snapbox, InputParams &p)
{
  std::ifstream redlist(filredshiftlist.c_str());
  double zlast = -999.9;
  if (!redlist.is_open())
  {
    std::cerr << " redshift list file redshift_list.txt does not " << std::endl;
    std::cerr << " exist in the Code dir ... check this out      " << std::endl;
    std::cerr << "    I will STOP here !!! " << std::endl;
    return 1;
  }
  std::string name;
  Header header;
  do
  {
    redlist >> name;
    snappath.push_back(name);
    if (readHeader(p.pathsnap + name + ".0", header))
    {
      std::cerr << name << " not found!" << std::endl;
      return 1;
    }
    if (header.redshift < zlast)
    {
      std::cerr << " Snapshots on " << filredshiftlist << " are not sorted!" << std::endl;
      return 1;
    }
    zlast = header.redshift;
    if (std::abs(zlast) < 1e-5)
      zlast = 0.0;
    snapred.push_back(zlast);
    snapbox.push_back(header.boxsize);
  } while ((header.redshift < p.zs) & (!redlist.eof()));
  return 0;
}

// gadget2io.cpp:34-48
void testHydro(InputParams &p, const Header &data)
{
  if (p.simType == "Gadget")
  {
    int dimmass0 = 0;
    for (int i = 0; i <= 5; i++)
      if (data.massarr[i] == 0)
        dimmass0 += data.npart[i];
    p.hydro = dimmass0 != 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GSL gsl_interp_cspline: natural boundary (c_0 = c_n = 0), interior second-derivative coefficients from the
// symmetric tridiagonal system diag_i = 2(h_i + h_{i+1}), offdiag_i = h_{i+1}, rhs_i = 3(dy_{i+1}/h_{i+1} - dy_i/h_i)
// solved by the L D L^T recurrence of GSL's solve_tridiag, evaluated as y_i + t(b + t(c_i + t d)).
// ---------------------------------------------------------------------------------------------------------------
void CubicSpline::init(const std::vector<double> &x, const std::vector<double> &y)
{
  x_ = x;
  y_ = y;
  const size_t n = x.size();
  c_.assign(n, 0.0);
  if (n < 3)
    return;
  const size_t N = n - 2;
  std::vector<double> g(N), diag(N), off(N);
  for (size_t i = 0; i < N; i++)
  {
    const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    const double dy_i = y[i + 1] - y[i], dy_ip1 = y[i + 2] - y[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (dy_ip1 * g_ip1 - dy_i * g_i);
  }
  if (N == 1)
  {
    c_[1] = g[0] / diag[0];
    return;
  }
  std::vector<double> alpha(N), gamma(N), cc(N), z(N);
  alpha[0] = diag[0];
  gamma[0] = off[0] / alpha[0];
  for (size_t i = 1; i + 1 < N; i++)
  {
    alpha[i] = diag[i] - off[i - 1] * gamma[i - 1];
    gamma[i] = off[i] / alpha[i];
  }
  alpha[N - 1] = diag[N - 1] - off[N - 2] * gamma[N - 2];
  z[0] = g[0];
  for (size_t i = 1; i < N; i++)
    z[i] = g[i] - gamma[i - 1] * z[i - 1];
  for (size_t i = 0; i < N; i++)
    cc[i] = z[i] / alpha[i];
  c_[N] = cc[N - 1];
  for (size_t i = N - 1; i-- > 0;)
    c_[i + 1] = cc[i] - gamma[i] * c_[i + 2];
}

double CubicSpline::eval(double x) const
{
  const size_t n = x_.size();
  if (n < 2 || x < x_[0] || x > x_[n - 1])
    return std::numeric_limits<double>::quiet_NaN();
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1)
  {
    const size_t mid = (hi + lo) / 2;
    if (x_[mid] > x)
      hi = mid;
    else
      lo = mid;
  }
  const double dx = x_[lo + 1] - x_[lo], dy = y_[lo + 1] - y_[lo];
  const double b = (dy / dx) - dx * (c_[lo + 1] + 2.0 * c_[lo]) / 3.0;
  const double d = (c_[lo + 1] - c_[lo]) / (3.0 * dx);
  const double t = x - x_[lo];
  return y_[lo] + t * (b + t * (c_[lo] + t * d));
}

// w0waCDM.cpp:18-84 as main drives it (slicer-v2.cpp:79-86): H0 = 100, wa = 0, z_i = i (zs+1)/(neval-1) ascending.
// comovingDistance integrates each table point from the previous cached one with dz = (z - lastZ)/100 and a
// `zi < z` loop that often takes a 101st step; the cached (unscaled) value accumulates.  Reproduced as is.
static double Hz(double z, double H0, double om, double ol, double w0, double wa)
{
  const double rhoLambda = ol * pow(1 + z, 3 * (1 + w0 + wa)) * exp(-3 * wa * z / (1 + z));
  const double rhoM = om * pow(1 + z, 3);
  const double rhoTot = rhoLambda + rhoM + (1 - om - ol) * pow(1 + z, 2);
  return H0 * sqrt(rhoTot);
}

void CosmoTable::build(double om0, double oml, double w, double zs)
{
  const double CSPEEDOFLIGHT = speedcunit * 100, H0 = 100.0, wa = 0.0;
  zl.assign(neval, 0.0);
  dl.assign(neval, 0.0);
  bool have_prev = false;
  double prev_z = 0, prev_d = 0;
  for (int i = 0; i < neval; i++)
  {
    const double z = i * (zs + 1.0) / (neval - 1);
    zl[i] = z;
    double distance = 0, lastZ = 0, dz = 1e-4;
    if (have_prev && prev_z == z)
    { // cache hit returns the UNSCALED value (w0waCDM.cpp:30-33); unreachable for strictly increasing z
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
      distance += 0.5 * dz * (1.0 / Hz(zi, H0, om0, oml, w, wa) + 1.0 / Hz(zi + dz, H0, om0, oml, w, wa));
    prev_z = z;
    prev_d = distance;
    have_prev = true;
    const double D_C = distance * CSPEEDOFLIGHT;
    if (fabs(1 - om0 - oml) < 1e-5)
      dl[i] = D_C;
    else
    {
      const double OmegaK = 1.0 - om0 - oml, s = sqrt(fabs(OmegaK));
      dl[i] = OmegaK < 0 ? CSPEEDOFLIGHT / H0 / s * sinh(s * H0 / CSPEEDOFLIGHT * D_C) : CSPEEDOFLIGHT / H0 / s * sin(s * H0 / CSPEEDOFLIGHT * D_C);
    }
  }
  getDl.init(zl, dl);
  getZl.init(dl, zl);
}

// densitymaps.cpp:9-32 — note the float `test` (the comparison is made on a float-rounded distance)
int getSnap(const std::vector<double> &zsnap, const CubicSpline &getDl, double dlens)
{
  if (zsnap.empty())
    return -1;
  unsigned pos = 0;
  double aux = 99999;
  for (size_t i = 0; i < zsnap.size(); i++)
  {
    const float test = std::abs(getDl.eval(zsnap[i]) - dlens);
    if (test < aux)
    {
      aux = test;
      pos = i;
    }
  }
  return pos;
}

// densitymaps.cpp:46-156
int buildPlanes(InputParams &p, Lens &lens, std::vector<double> &snapred, std::vector<std::string> &snappath,
                std::vector<double> &snapbox, const CubicSpline &getDl, const CubicSpline &getZl, int numOfLensPerSnap, int myid)
{
  const size_t nsnaps = snapred.size();
  int pos = 0, nrepi = 0, nrep = 0;
  double ldbut = 0.0;
  do
  {
    nrep++;
    nrepi++;
    double ztest = 9999;
    int pos_temp = pos;
    for (size_t i = pos_temp; i < nsnaps; i++)
    {
      const double dtest = ldbut + snapbox[i] / (1e3 / POS_U) / numOfLensPerSnap;
      const int itest = getSnap(snapred, getDl, dtest);
      if (itest == -1)
      {
        std::cerr << "snapred is an empty array!" << std::endl;
        std::cerr << "Check your snapshot list file." << std::endl;
        return 1;
      }
      const double dz = fabs(snapred[itest] - getZl.eval(dtest));
      if (dz < ztest)
        if (nrep == 1 || (!bool((nrep - 1) % numOfLensPerSnap) || snapbox[itest] == snapbox[pos]))
        {
          pos_temp = itest;
          ztest = dz;
        }
    }
    ldbut += snapbox[pos_temp] / (1e3 / POS_U) / numOfLensPerSnap;
    const double dlens = ldbut - 0.5 * snapbox[pos_temp] / (1e3 / POS_U) / numOfLensPerSnap;
    const double zlens = getZl.eval(dlens);
    pos_temp = getSnap(snapred, getDl, dlens);
    if (myid == 0)
      std::cout << " simulation snapshots = " << ldbut << "  " << getZl.eval(ldbut) << "  " << nrep << " from snap " << snappath[pos_temp]
                << "  " << zlens << std::endl;
    lens.ld.push_back(ldbut - snapbox[pos_temp] / (1e3 / POS_U) / numOfLensPerSnap);
    lens.ld2.push_back(ldbut);
    lens.zfromsnap.push_back(snapred[pos_temp]);
    if (nrep != 1 && pos_temp != pos)
    {
      for (int i = 0; i < nrepi - 1; i++)
        lens.replication.push_back(nrep - 1);
      nrepi = 1;
    }
    pos = pos_temp;
    lens.zsimlens.push_back(zlens);
    lens.fromsnap.push_back(snappath[pos]);
    lens.fromsnapi.push_back(pos);
    lens.randomize.push_back(nrep == 1 ? true : !((nrep - 1) % numOfLensPerSnap));
  } while (ldbut < p.Ds);
  for (int i = 0; i < nrepi + 1; i++)
    lens.replication.push_back(nrep);
  if (myid == 0)
  {
    std::cout << " Comoving Distance of the last plane " << p.Ds << std::endl;
    std::cout << " nsnaps = " << nsnaps << "\n" << std::endl;
  }
  std::ofstream planelist;
  const std::string planes_list = p.directory + "planes_list_" + p.suffix + ".txt";
  if (myid == 0)
    planelist.open(planes_list.c_str());
  for (size_t i = 0; i < lens.fromsnap.size(); i++)
  {
    if (myid == 0)
    {
      std::cout << lens.zsimlens[i] << " planes = " << lens.ld[i] << "  " << lens.ld2[i] << "  " << lens.replication[i] << " from snap "
                << lens.fromsnap[i] << std::endl;
      planelist << i << "   " << lens.zsimlens[i] << "   " << lens.ld[i] << "   " << lens.ld2[i] << "   " << lens.replication[i] << "   "
                << lens.fromsnap[i] << "   " << lens.zfromsnap[i] << "  " << lens.randomize[i] << std::endl;
    }
    lens.pll.push_back(i);
  }
  if (myid == 0)
    planelist.close();
  lens.nplanes = lens.replication.back();
  return 0;
}

void GlibcRand::seed(unsigned int s)
{
  if (s == 0)
    s = 1;
  r_[0] = (int32_t)s;
  for (int i = 1; i < 31; i++)
  {
    const long hi = r_[i - 1] / 127773, lo = r_[i - 1] % 127773;
    long word = 16807 * lo - 2836 * hi;
    if (word < 0)
      word += 2147483647;
    r_[i] = (int32_t)word;
  }
  f_ = 3;
  b_ = 0;
  for (int i = 0; i < 310; i++)
    next();
}

int GlibcRand::next()
{
  const uint32_t v = (uint32_t)r_[f_] + (uint32_t)r_[b_];
  r_[f_] = (int32_t)v;
  f_ = f_ + 1 == 31 ? 0 : f_ + 1;
  b_ = b_ + 1 == 31 ? 0 : b_ + 1;
  return (int)(v >> 1);
}

GlibcRand &sharedRand()
{
  static GlibcRand g;
  return g;
}

// densitymaps.cpp:166-248 — the reference uses libc srand/rand; GlibcRand is the same generator with private state
void randomizeBox(Random &random, const Lens &lens, const InputParams &p, int numOfLensPerSnap, int myid, bool fixedVertex)
{
  const size_t nrandom = lens.replication.back();
  random.x0.resize(nrandom);
  random.y0.resize(nrandom);
  random.z0.resize(nrandom);
  random.sgnX.resize(nrandom);
  random.sgnY.resize(nrandom);
  random.sgnZ.resize(nrandom);
  random.face.resize(nrandom);
  for (size_t i = 0; i < nrandom; i++)
  {
    if (lens.randomize[i])
    {
      GlibcRand &R = sharedRand();
      R.seed(p.seedcenter + i / numOfLensPerSnap * 13);
      if (!fixedVertex)
      {
        random.x0[i] = R.next() / float(RAND_MAX);
        random.y0[i] = R.next() / float(RAND_MAX);
        random.z0[i] = R.next() / float(RAND_MAX);
      }
      else
      { // -DFixedPLCVertex, densitymaps.cpp:191-195
        random.x0[i] = 0.0;
        random.y0[i] = 0.0;
        random.z0[i] = 0.5;
      }
      random.face[i] = 7;
      R.seed(p.seedface + i / numOfLensPerSnap * 5);
      while (random.face[i] > 6 || random.face[i] < 1)
        random.face[i] = int(1 + R.next() / float(RAND_MAX) * 5. + 0.5);
      R.seed(p.seedsign + i / numOfLensPerSnap * 8);
      int *sg[3] = {&random.sgnX[i], &random.sgnY[i], &random.sgnZ[i]};
      for (int k = 0; k < 3; k++)
      {
        *sg[k] = 2;
        while (*sg[k] > 1 || *sg[k] < 0)
          *sg[k] = int(R.next() / float(RAND_MAX) + 0.5);
        if (*sg[k] == 0)
          *sg[k] = -1;
      }
    }
    else
    {
      random.x0[i] = random.x0[i - 1];
      random.y0[i] = random.y0[i - 1];
      random.z0[i] = random.z0[i - 1];
      random.face[i] = random.face[i - 1];
      random.sgnX[i] = random.sgnX[i - 1];
      random.sgnY[i] = random.sgnY[i - 1];
      random.sgnZ[i] = random.sgnZ[i - 1];
    }
    if (myid == 0)
      std::cout << " plane " << i << " centre " << random.x0[i] << " " << random.y0[i] << " " << random.z0[i] << " face " << random.face[i]
                << " signs " << random.sgnX[i] << " " << random.sgnY[i] << " " << random.sgnZ[i] << std::endl;
  }
}

// densitymaps.cpp:255-269 — only rank 0 reports the failure (quirk iii of App. D.7 kept)
int testFov(double fov, double boxl, double Ds, int myid, double &fovradiants)
{
  fovradiants = fov / 180. * M_PI;
  if ((fovradiants)*Ds > boxl && myid == 0)
  {
    std::cerr << " !!Field view too large!!\n !!!I will STOP here!!! " << std::endl;
    std::cerr << " Value set is = " << fov << std::endl;
    std::cerr << " Maximum value allowed " << boxl / Ds * 180. / M_PI << " in degrees " << std::endl;
    std::cerr << " For the lens at " << Ds << std::endl;
    return 1;
  }
  return 0;
}

// densitymaps.cpp:275-283
void computeReplications(double fov, double boxl, double Ds, int, double &fovradiants, int &nrepperp)
{
  fovradiants = fov / 180. * M_PI;
  if (Ds * tan(fovradiants / 2.0) <= boxl / 2.0)
    nrepperp = 0;
  else
    nrepperp = ceil((Ds * tan(fovradiants / 2) - boxl / 2.0) / boxl);
}

} // namespace slicer
