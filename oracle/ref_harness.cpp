/* TEST INFRASTRUCTURE — C entry points over the UNMODIFIED reference functions.
 *
 * Compiled together with /root/reference/SLICER/{utilities,data,gadget2io,densitymaps,w0waCDM,
 * writeplc}.cpp (sources stay where they are; see oracle/Makefile) into oracle/_ref/libslicer_ref.so
 * so that tests and bench.py's cpu_baseline leg can call the reference's own code through ctypes.
 * Nothing in the product links or loads this.
 *
 * Each entry point only marshals arguments into the reference's structs (data.h:29-131) and calls:
 *   weight / getPolar / gridist_w            utilities.cpp:4,19,36
 *   randomizeBox / buildPlanes               densitymaps.cpp:166,46
 *   readHeader / readPos / fastforwardToBlock gadget2io.cpp:8,174,133
 *   mapParticles / createDensityMaps         densitymaps.cpp:297,419
 *   w0waCDM::transverseComovingDistance      w0waCDM.cpp:60
 */
#include "densitymaps.h"
#include "w0waCDM.h"
#include <cstring>

namespace
{
  void fill_params(InputParams &p, int npix, double fov_deg, int snopt, int hydro, int partinplanes)
  {
    p.npix = npix;
    p.zs = 0;
    p.Ds = 0;
    p.fov = fov_deg;
    p.hydro = hydro != 0;
    p.simType = "Gadget";
    p.rgrid = 0;
    p.seedcenter = p.seedface = p.seedsign = 0;
    p.partinplanes = partinplanes != 0;
    p.snopt = snopt;
    p.physical = false;
    p.w = -1;
    p.snpix = sconv(npix, fINT);
  }

  void fill_one_plane(Lens &lens, Random &random, const int *sgn, int face, const double *centre,
                      double ld, double ld2, int nrepperp)
  {
    lens.nplanes = 1;
    lens.ld.assign(1, ld);
    lens.ld2.assign(1, ld2);
    lens.nrepperp.assign(1, nrepperp);
    lens.randomize.assign(1, true);
    lens.replication.assign(1, 1);
    random.x0.assign(1, centre[0]);
    random.y0.assign(1, centre[1]);
    random.z0.assign(1, centre[2]);
    random.sgnX.assign(1, sgn[0]);
    random.sgnY.assign(1, sgn[1]);
    random.sgnZ.assign(1, sgn[2]);
    random.face.assign(1, face);
  }

  struct GadgetArrays
  {
    Gadget g;
    float *xx[6][3];
    explicit GadgetArrays(const Header &h)
    {
      vector<float> *v[6][3] = {{&g.xx0, &g.yy0, &g.zz0}, {&g.xx1, &g.yy1, &g.zz1}, {&g.xx2, &g.yy2, &g.zz2},
                                {&g.xx3, &g.yy3, &g.zz3}, {&g.xx4, &g.yy4, &g.zz4}, {&g.xx5, &g.yy5, &g.zz5}};
      for (int i = 0; i < 6; i++)
        for (int k = 0; k < 3; k++)
        {
          v[i][k]->resize(h.npart[i]);
          xx[i][k] = h.npart[i] > 0 ? &(*v[i][k])[0] : nullptr;
        }
    }
  };
}

extern "C"
{

  int ref_do_ngp() { return DO_NGP ? 1 : 0; }
  int ref_lens_per_snap() { return numberOfLensPerSnap; }
  double ref_max_m() { return MAX_M; }

  float ref_weight(float ixx, float ixh, double dx) { return weight(ixx, ixh, dx); }

  void ref_getpolar(double x, double y, double z, double *ra, double *dec, double *d)
  {
    getPolar(x, y, z, *ra, *dec, *d, true);
  }

  int ref_gridist_w(const float *x, const float *y, const float *w, long n, int nn, int do_ngp, float *out)
  {
    vector<float> vx(x, x + n), vy(y, y + n), vw(w, w + n);
    valarray<float> m = gridist_w(vx, vy, vw, nn, do_ngp != 0);
    for (long i = 0; i < (long)nn * nn; i++)
      out[i] = m[i];
    return 0;
  }

  void ref_srand(unsigned seed) { srand(seed); }

  int ref_randomize_box(int seedcenter, int seedface, int seedsign, int nplanes, const int *randomize,
                        double *x0, double *y0, double *z0, int *face, int *sx, int *sy, int *sz)
  {
    InputParams p;
    fill_params(p, 1, 1, 0, 0, 0);
    p.seedcenter = seedcenter;
    p.seedface = seedface;
    p.seedsign = seedsign;
    Lens lens;
    lens.replication.assign(1, nplanes);
    for (int i = 0; i < nplanes; i++)
      lens.randomize.push_back(randomize[i] != 0);
    Random random;
    randomizeBox(random, lens, p, numberOfLensPerSnap, 1);
    for (int i = 0; i < nplanes; i++)
    {
      x0[i] = random.x0[i];
      y0[i] = random.y0[i];
      z0[i] = random.z0[i];
      face[i] = random.face[i];
      sx[i] = random.sgnX[i];
      sy[i] = random.sgnY[i];
      sz[i] = random.sgnZ[i];
    }
    return 0;
  }

  /* slicer-v2.cpp:79-86 */
  int ref_cosmo_table(double om0, double oml, double w, double zs, int n, double *zl, double *dl)
  {
    w0waCDM cosmo(100.0, om0, oml, w, 0.0);
    for (int i = 0; i < n; i++)
    {
      zl[i] = i * (zs + 1.0) / (n - 1);
      dl[i] = cosmo.transverseComovingDistance(zl[i]);
    }
    return 0;
  }

  /* slicer-v2.cpp:79-100 on caller-supplied snapshot redshifts / box sizes (kpc/h). */
  int ref_plan(double om0, double oml, double w, double zs, int nsnaps, const double *snapred_in,
               const double *snapbox_in, const char *directory, const char *suffix, int cap, int *nplanes,
               double *Ds, double *ld, double *ld2, double *zsimlens, double *zfromsnap, int *fromsnapi,
               int *randomize, int *replication, int *nreplication)
  {
    const int n = 1000; /* neval, slicer-v2.cpp:5 */
    vector<double> zl(n), dl(n);
    ref_cosmo_table(om0, oml, w, zs, n, &zl[0], &dl[0]);
    gsl_interp_accel *accGetDl = gsl_interp_accel_alloc();
    gsl_interp_accel *accGetZl = gsl_interp_accel_alloc();
    gsl_spline *getDl = gsl_spline_alloc(gsl_interp_cspline, n);
    gsl_spline *getZl = gsl_spline_alloc(gsl_interp_cspline, n);
    gsl_spline_init(getDl, &zl[0], &dl[0], n);
    gsl_spline_init(getZl, &dl[0], &zl[0], n);
    InputParams p;
    fill_params(p, 1, 1, 0, 0, 0);
    p.zs = zs;
    p.directory = directory;
    p.suffix = suffix;
    p.Ds = gsl_spline_eval(getDl, p.zs, accGetDl);
    *Ds = p.Ds;
    vector<double> snapred(snapred_in, snapred_in + nsnaps), snapbox(snapbox_in, snapbox_in + nsnaps);
    vector<string> snappath;
    for (int i = 0; i < nsnaps; i++)
      snappath.push_back("snap_" + sconv(i, fINT));
    Lens lens;
    int rc = buildPlanes(p, lens, snapred, snappath, snapbox, getDl, accGetDl, getZl, accGetZl, numberOfLensPerSnap, 0);
    gsl_spline_free(getDl);
    gsl_spline_free(getZl);
    gsl_interp_accel_free(accGetDl);
    gsl_interp_accel_free(accGetZl);
    if (rc)
      return rc;
    *nplanes = lens.nplanes;
    *nreplication = (int)lens.replication.size();
    if ((int)lens.ld.size() > cap || (int)lens.replication.size() > cap)
      return 2;
    for (size_t i = 0; i < lens.ld.size(); i++)
    {
      ld[i] = lens.ld[i];
      ld2[i] = lens.ld2[i];
      zsimlens[i] = lens.zsimlens[i];
      zfromsnap[i] = lens.zfromsnap[i];
      fromsnapi[i] = lens.fromsnapi[i];
      randomize[i] = lens.randomize[i] ? 1 : 0;
    }
    for (size_t i = 0; i < lens.replication.size(); i++)
      replication[i] = lens.replication[i];
    return 0;
  }

  int ref_read_header(const char *file, int *npart, double *massarr, double *scalars /*time,z,box,om0,oml,h*/,
                      int *numfiles)
  {
    Header h;
    ifstream fin;
    if (readHeader(file, h, fin, true))
      return 1;
    for (int i = 0; i < 6; i++)
    {
      npart[i] = h.npart[i];
      massarr[i] = h.massarr[i];
    }
    scalars[0] = h.time;
    scalars[1] = h.redshift;
    scalars[2] = h.boxsize;
    scalars[3] = h.om0;
    scalars[4] = h.oml;
    scalars[5] = h.h;
    *numfiles = h.numfiles;
    return 0;
  }

  /* gadget2io.cpp:174 on one sub-file; outputs SoA, the six types concatenated in type order. */
  int ref_read_pos(const char *file, const int *sgn, int face, const double *centre, float rcase,
                   float *x, float *y, float *z)
  {
    Header data;
    ifstream fin;
    if (readHeader(file, data, fin, false))
      return 1;
    InputParams p;
    fill_params(p, 1, 1, 0, 0, 0);
    Lens lens;
    Random random;
    fill_one_plane(lens, random, sgn, face, centre, 0, 1, 0);
    GadgetArrays ga(data);
    readPos(fin, data, p, random, 0, ga.xx, rcase, 1);
    fin.close();
    size_t off = 0;
    for (int i = 0; i < 6; i++)
    {
      size_t n = data.npart[i];
      if (n)
      {
        memcpy(x + off, ga.xx[i][0], n * sizeof(float));
        memcpy(y + off, ga.xx[i][1], n * sizeof(float));
        memcpy(z + off, ga.xx[i][2], n * sizeof(float));
      }
      off += n;
    }
    return 0;
  }

  /* One sub-file through readPos + mapParticles exactly as createDensityMaps drives them
   * (densitymaps.cpp:435-509), but handing back what createDensityMaps loses: the per-type maps of
   * this sub-file AND the per-type accepted counts (densitymaps.cpp:497 shadows them).
   * maps: [6][npix*npix] float, counts: [6].  Returns mapParticles' return code. */
  int ref_map_subfile(const char *file, int npix, double fovradiants, int snopt, int hydro, const int *sgn,
                      int face, const double *centre, float rcase, double ld, double ld2, int nrepperp,
                      float *maps, int *counts)
  {
    Header data;
    ifstream fin;
    if (readHeader(file, data, fin, false))
      return 1;
    InputParams p;
    fill_params(p, npix, 0, snopt, hydro, 0);
    Lens lens;
    Random random;
    fill_one_plane(lens, random, sgn, face, centre, ld, ld2, nrepperp);
    GadgetArrays ga(data);
    readPos(fin, data, p, random, 0, ga.xx, rcase, 1);
    if (p.hydro)
      fastforwardToBlock(fin, "MASS", 1);
    else
    {
      fin.clear();
      fin.close();
    }
    valarray<float> mapxyi[6];
    int ntotxyi[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++)
      mapxyi[i].resize((size_t)npix * npix);
    int rc = mapParticles(fin, data, p, lens, ga.xx, fovradiants, 0, mapxyi, ntotxyi, 1);
    if (p.hydro)
    {
      fin.clear();
      fin.close();
    }
    for (int i = 0; i < 6; i++)
    {
      counts[i] = ntotxyi[i];
      for (size_t k = 0; k < (size_t)npix * npix; k++)
        maps[(size_t)i * npix * npix + k] = mapxyi[i][k];
    }
    return rc;
  }

  /* The reference's own createDensityMaps (densitymaps.cpp:419) over sub-files [ffmin, ffmax) of
   * <file_base>.<ff>.  maptot: [npix*npix]; mapsi: [6][npix*npix]. */
  int ref_create_density_maps(const char *file_base, unsigned ffmin, unsigned ffmax, int npix, double fovradiants,
                              int snopt, int hydro, const int *sgn, int face, const double *centre, double rcase,
                              double ld, double ld2, int nrepperp, float *maptot, float *mapsi)
  {
    InputParams p;
    fill_params(p, npix, 0, snopt, hydro, 0);
    Lens lens;
    Random random;
    fill_one_plane(lens, random, sgn, face, centre, ld, ld2, nrepperp);
    valarray<float> mapxytot, mapxytoti[6];
    int ntotxyi[6];
    int rc = createDensityMaps(p, lens, random, 0, ffmin, ffmax, file_base, fovradiants, rcase, nullptr, nullptr,
                               nullptr, nullptr, mapxytot, mapxytoti, ntotxyi, 1);
    if (rc)
      return rc;
    size_t np = (size_t)npix * npix;
    for (size_t k = 0; k < np; k++)
      maptot[k] = mapxytot[k];
    if (mapsi)
      for (int i = 0; i < 6; i++)
        for (size_t k = 0; k < np; k++)
          mapsi[i * np + k] = mapxytoti[i][k];
    return 0;
  }
}
