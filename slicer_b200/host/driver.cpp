// driver.cpp — createDensityMaps on the GPU and the plane driver of main() (slicer-v2.cpp:130-230), restructured:
// the reference walks the PLANES and re-reads + re-transforms a whole snapshot for each one (:138-207); here the
// loop is over SNAPSHOTS, every sub-file is read once into page-locked memory, copied once, and one CUDA pass
// deposits it into all the planes that use the snapshot.  Sub-files are dealt round-robin to the GPUs of the box
// (the reference deals them to MPI ranks, :162-175) and the planes are summed with ncclReduce instead of MPI_Reduce.
#include "slicer_host.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <thread>

namespace slicer
{

// wall-clock phase totals of the last runLightCone (seconds): printed when SLICER_B200_TIMING is set
static double t_read = 0, t_submit = 0, t_fetch = 0, t_write = 0, t_engine = 0;
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Engine
{
  std::vector<int> devices;
  std::vector<slicer_handle *> h;
  std::vector<SubFile> bufs; // two page-locked sub-file buffers per GPU
  std::vector<int> used;     // sub-files staged on each GPU since begin
  int npix_max = 0, mas = SLICER_MAS_TSC, deposit_mode = SLICER_DEPOSIT_AUTO;
  bool per_type = false;
  size_t capacity = 0;
  int max_planes = SLICER_MAX_PLANES; // accumulators per GPU: the most planes one snapshot feeds
  bool comm = false;
};

static int make_handles(Engine *e)
{
  for (auto *hh : e->h)
    slicer_destroy(hh);
  e->h.assign(e->devices.size(), nullptr);
  e->comm = false;
  // one thread per GPU: creating a CUDA context, loading the kernels and allocating the pools takes ~2 s per device, serially
  // that was most of an 8-GPU run's wall clock
  std::vector<std::string> errs(e->devices.size());
  auto make_one = [&](size_t g) {
    slicer_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = e->devices[g];
    cfg.mas = e->mas;
    cfg.max_m = MAX_M;
    cfg.max_planes = e->max_planes;
    cfg.npix_max = e->npix_max;
    cfg.per_type_maps = e->per_type;
    cfg.particle_capacity = e->capacity;
    cfg.mass_capacity = e->capacity;
    cfg.kernel = SLICER_KERNEL_AUTO;
    cfg.staging_buffers = 2;
    cfg.deposit_mode = e->deposit_mode;
    if (slicer_create(&cfg, &e->h[g]))
      errs[g] = slicer_last_error(); // (thread-local: read it on the thread that made the call)
  };
  if (e->devices.size() == 1)
    make_one(0);
  else
  {
    std::vector<std::thread> th;
    for (size_t g = 0; g < e->devices.size(); g++)
      th.emplace_back(make_one, g);
    for (auto &t : th)
      t.join();
  }
  for (size_t g = 0; g < e->devices.size(); g++)
    if (!e->h[g])
    {
      std::cerr << "slicer_create failed on GPU " << e->devices[g] << ": " << errs[g] << std::endl;
      return 1;
    }
  if (e->h.size() > 1)
  {
    if (slicer_comm_init_all(e->h.data(), (int)e->h.size()))
    {
      std::cerr << "NCCL communicator: " << slicer_last_error() << std::endl;
      return 1;
    }
    e->comm = true;
  }
  return 0;
}

Engine *engineCreate(const std::vector<int> &devices, int npix_max, int mas, bool per_type_maps, size_t particle_capacity, int deposit_mode,
                     int max_planes)
{
  Engine *e = new Engine;
  e->devices = devices;
  e->max_planes = std::min(std::max(max_planes, 1), (int)SLICER_MAX_PLANES);
  e->npix_max = npix_max;
  e->mas = mas;
  e->per_type = per_type_maps;
  e->capacity = particle_capacity;
  e->deposit_mode = deposit_mode;
  e->bufs = std::vector<SubFile>(2 * devices.size());
  e->used.assign(devices.size(), 0);
  if (make_handles(e))
  {
    engineDestroy(e);
    return nullptr;
  }
  return e;
}

void engineDestroy(Engine *e)
{
  if (!e)
    return;
  for (auto *hh : e->h)
    slicer_destroy(hh);
  delete e;
}

int engineGpuCount(const Engine *e) { return (int)e->h.size(); }

static slicer_plane_desc make_desc(const Lens &lens, const Random &random, int isnap, double rcase, double fovradiants, int npix)
{
  slicer_plane_desc d;
  memset(&d, 0, sizeof(d));
  d.sgn[0] = random.sgnX[isnap];
  d.sgn[1] = random.sgnY[isnap];
  d.sgn[2] = random.sgnZ[isnap];
  d.face = random.face[isnap];
  d.centre[0] = random.x0[isnap];
  d.centre[1] = random.y0[isnap];
  d.centre[2] = random.z0[isnap];
  d.rcase = (float)rcase; // createDensityMaps takes a double, readPos a float (densitymaps.h:162, gadget2io.h:123)
  d.ld = lens.ld[isnap];
  d.ld2 = lens.ld2[isnap];
  d.nrepperp = lens.nrepperp.empty() ? 0 : lens.nrepperp[isnap];
  d.fovradiants = fovradiants;
  d.npix = npix;
  return d;
}

static int fail_capi(const char *what)
{
  std::cerr << what << ": " << slicer_last_error() << std::endl;
  return 1;
}

// Stage every particle type of a sub-file that is already in `buf` (types in order: the order mapParticles walks them).
static int stage_subfile(slicer_handle *h, const SubFile &buf, bool hydro)
{
  size_t off = 0;
  for (int t = 0; t < 6; t++)
  {
    const size_t nt = (size_t)buf.header.npart[t];
    if (nt)
    {
      const bool pm = hydro && buf.header.massarr[t] == 0;
      if (slicer_stage_particles(h, t, buf.pos + 3 * off, SLICER_LAYOUT_AOS, pm ? buf.mass + off : nullptr, nt))
        return fail_capi("slicer_stage_particles");
    }
    off += nt;
  }
  return 0;
}

// The device staging pools must hold the largest sub-file of the snapshot (sub-files of one snapshot differ by tens of per cent
// with domain decomposition or gas).  The sub-file headers are scanned first (256 bytes each); if one needs more room than the
// handles have, the handles are rebuilt with that capacity before any work of this snapshot is queued.
static int ensure_capacity(Engine *e, const std::string &File, unsigned ffmin, unsigned ffmax)
{
  size_t largest = 0;
  for (unsigned ff = ffmin; ff < ffmax; ff++)
  {
    Header hd;
    if (readHeader(File + "." + std::to_string(ff), hd))
      return 1;
    size_t n = 0;
    for (int t = 0; t < 6; t++)
      n += (size_t)(hd.npart[t] > 0 ? hd.npart[t] : 0);
    largest = std::max(largest, n);
  }
  if (largest <= e->capacity)
    return 0;
  for (auto *hh : e->h)
    if (hh && slicer_synchronize(hh))
      return fail_capi("slicer_synchronize");
  e->capacity = largest + largest / 8 + 4096;
  return make_handles(e);
}

// `Part. Degradation` (snopt > 0, densitymaps.cpp:387-397).  The reference consumes one libc rand() per accepted
// (particle, replica) pair, plane after plane, sub-file after sub-file, type after type, particle after particle —
// one serial stream that continues from randomizeBox's last srand().  Sweep 1 counts the accepted pairs of every
// (plane, sub-file, type); the draws are then made here, in that order, with the same libc generator; sweep 2 applies
// them on the device by rank (slicer_count_accepted / slicer_deposit_degraded).
static int createDensityMapsDegraded(Engine *e, InputParams &p, Lens &lens, Random &random, const std::vector<PlaneJob> &jobs, unsigned ffmin,
                                     unsigned ffmax, const std::string &File, double fovradiants, std::vector<std::valarray<float>> &mapxytot,
                                     std::vector<std::valarray<float>> &mapxytoti, std::vector<long long> &ntotxyi, int myid)
{
  const int nj = (int)jobs.size();
  if (nj > e->max_planes)
  {
    std::cerr << "Part. Degradation: more than " << e->max_planes << " planes share one snapshot; not supported" << std::endl;
    return 1;
  }
  mapxytot.assign(nj, std::valarray<float>());
  mapxytoti.assign((size_t)nj * 6, std::valarray<float>());
  ntotxyi.assign((size_t)nj * 6, 0);
  const int ngpu = (int)e->h.size();
  const unsigned nff = ffmax - ffmin;
  std::vector<slicer_plane_desc> descs;
  for (int j = 0; j < nj; j++)
    descs.push_back(make_desc(lens, random, jobs[j].isnap, jobs[j].rcase, fovradiants, jobs[j].npix));
  if (ensure_capacity(e, File, ffmin, ffmax))
    return 1;
  // ---- sweep 1: accepted pairs per (plane, sub-file, type)
  std::vector<long long> cnt((size_t)nff * nj * 6, 0);
  for (unsigned ff = ffmin; ff < ffmax; ff++)
  {
    const int g = (int)((ff - ffmin) % ngpu);
    SubFile &buf = e->bufs[2 * g];
    if (slicer_synchronize(e->h[g]) || readSubFile(File + "." + std::to_string(ff), p.hydro, buf, true))
      return 1;
    if (buf.ntotal > e->capacity)
    {
      std::cerr << "sub-file " << ff << " holds " << buf.ntotal << " particles, more than the engine's capacity " << e->capacity << std::endl;
      return 1;
    }
    if (slicer_begin_snapshot(e->h[g], buf.header.boxsize, buf.header.massarr, p.hydro) || stage_subfile(e->h[g], buf, p.hydro))
      return fail_capi("staging");
    if (slicer_count_accepted(e->h[g], descs.data(), nj, &cnt[(size_t)(ff - ffmin) * nj * 6]))
      return fail_capi("slicer_count_accepted");
  }
  // ---- the draws, in the reference's order: plane-major, then sub-file, then type/particle/replica
  const double thr = 1. / pow(2, p.snopt);
  std::vector<std::vector<unsigned char>> keep((size_t)nj * nff);
  for (int j = 0; j < nj; j++)
    for (unsigned f = 0; f < nff; f++)
    {
      long long n = 0;
      for (int t = 0; t < 6; t++)
        n += cnt[((size_t)f * nj + j) * 6 + t];
      std::vector<unsigned char> &k = keep[(size_t)j * nff + f];
      k.resize((size_t)n);
      for (long long i = 0; i < n; i++)
        k[i] = (sharedRand().next() / float(RAND_MAX) < thr) ? 1 : 0; // densitymaps.cpp:393
    }
  // ---- sweep 2: deposit with the draws applied by rank
  std::fill(e->used.begin(), e->used.end(), 0);
  Header first_header;
  bool have_header = false;
  for (unsigned ff = ffmin; ff < ffmax; ff++)
  {
    const int g = (int)((ff - ffmin) % ngpu);
    SubFile &buf = e->bufs[2 * g];
    if (slicer_synchronize(e->h[g]) || readSubFile(File + "." + std::to_string(ff), p.hydro, buf, true))
      return 1;
    if (!have_header)
    {
      first_header = buf.header;
      have_header = true;
    }
    if (myid == 0)
      std::cout << " sub-file " << ff << ": " << buf.ntotal << " particles -> GPU " << e->devices[g] << " (degradation 2^-" << p.snopt << ")" << std::endl;
    if (slicer_begin_snapshot(e->h[g], buf.header.boxsize, buf.header.massarr, p.hydro) || stage_subfile(e->h[g], buf, p.hydro))
      return fail_capi("staging");
    std::vector<long long> again((size_t)nj * 6);
    if (slicer_count_accepted(e->h[g], descs.data(), nj, again.data()))
      return fail_capi("slicer_count_accepted");
    std::vector<const unsigned char *> kp(nj);
    for (int j = 0; j < nj; j++)
      kp[j] = keep[(size_t)j * nff + (ff - ffmin)].data();
    if (slicer_deposit_degraded(e->h[g], descs.data(), nj, p.snopt, kp.data(), e->used[g] > 0))
      return fail_capi("slicer_deposit_degraded");
    e->used[g]++;
  }
  for (int g = 0; g < ngpu; g++)
    if (e->used[g] == 0)
    {
      const double zero[6] = {0, 0, 0, 0, 0, 0};
      if (slicer_begin_snapshot(e->h[g], have_header ? first_header.boxsize : 1.0, have_header ? first_header.massarr : zero, p.hydro) ||
          slicer_deposit(e->h[g], descs.data(), nj))
        return fail_capi("slicer_deposit");
    }
  if (ngpu > 1 && slicer_reduce_all(e->h.data(), ngpu, nj, 0))
    return fail_capi("slicer_reduce_all");
  for (int j = 0; j < nj; j++)
  {
    const int npix = jobs[j].npix;
    long long counts[6];
    mapxytot[j].resize((size_t)npix * npix);
    if (slicer_fetch(e->h[0], j, -1, &mapxytot[j][0], counts, nullptr))
      return fail_capi("slicer_fetch");
    for (int t = 0; t < 6; t++)
    {
      ntotxyi[(size_t)j * 6 + t] = counts[t];
      if (e->per_type)
      {
        mapxytoti[(size_t)j * 6 + t].resize((size_t)npix * npix);
        if (slicer_fetch(e->h[0], j, t, &mapxytoti[(size_t)j * 6 + t][0], nullptr, nullptr))
          return fail_capi("slicer_fetch");
      }
    }
  }
  return 0;
}

int createDensityMapsMulti(Engine *e, InputParams &p, Lens &lens, Random &random, const std::vector<PlaneJob> &jobs, unsigned ffmin,
                           unsigned ffmax, const std::string &File, double fovradiants, std::vector<std::valarray<float>> &mapxytot,
                           std::vector<std::valarray<float>> &mapxytoti, std::vector<long long> &ntotxyi, int myid)
{
  if (p.snopt > 0)
    return createDensityMapsDegraded(e, p, lens, random, jobs, ffmin, ffmax, File, fovradiants, mapxytot, mapxytoti, ntotxyi, myid);
  const int njobs = (int)jobs.size();
  mapxytot.assign(njobs, std::valarray<float>());
  mapxytoti.assign((size_t)njobs * 6, std::valarray<float>());
  ntotxyi.assign((size_t)njobs * 6, 0);
  const int ngpu = (int)e->h.size();
  for (int j0 = 0; j0 < njobs; j0 += e->max_planes)
  {
    const int nj = std::min(njobs - j0, e->max_planes);
    std::vector<slicer_plane_desc> descs;
    for (int j = 0; j < nj; j++)
      descs.push_back(make_desc(lens, random, jobs[j0 + j].isnap, jobs[j0 + j].rcase, fovradiants, jobs[j0 + j].npix));
    std::fill(e->used.begin(), e->used.end(), 0);
    Header first_header;
    bool have_header = false;
    if (ffmax > ffmin && readHeader(File + "." + std::to_string(ffmin), first_header) == 0)
      have_header = true;
    // One host thread per GPU (the reference: one MPI rank per group of sub-files, slicer-v2.cpp:162-175, densitymaps.cpp:432-484):
    // GPU g takes sub-files ffmin + g, ffmin + g + ngpu, ...; it reads sub-file k+1 into its second page-locked buffer while the
    // copy and the pass of sub-file k run.  Messages are collected per thread and printed in sub-file order afterwards.
    std::vector<int> rcs(ngpu, 0);
    std::vector<std::string> logs(ngpu), errs(ngpu);
    std::vector<double> read_s(ngpu, 0.0);
    auto work = [&](int g) {
      std::ostringstream log, err;
      for (unsigned ff = ffmin + (unsigned)g; ff < ffmax && !rcs[g]; ff += (unsigned)ngpu)
      {
        SubFile &buf = e->bufs[2 * g + (e->used[g] & 1)];
        if (e->used[g] >= 2 && slicer_wait_staging(e->h[g])) // the copy that last read this host buffer must be done
        {
          err << "slicer_wait_staging: " << slicer_last_error() << "\n";
          rcs[g] = 1;
          break;
        }
        const double tr0 = now_s();
        if (readSubFile(File + "." + std::to_string(ff), p.hydro, buf, true))
        {
          rcs[g] = 1;
          break;
        }
        read_s[g] += now_s() - tr0;
        const Header &data = buf.header;
        if (buf.ntotal > e->capacity)
        { // cannot happen after the header scan of ensure_capacity() unless the file changed in between
          err << "sub-file " << ff << " holds " << buf.ntotal << " particles, more than the engine's capacity " << e->capacity << "\n";
          rcs[g] = 1;
          break;
        }
        log << " sub-file " << ff << ": " << buf.ntotal << " particles -> GPU " << e->devices[g] << "\n";
        int rc = e->used[g] == 0 ? slicer_begin_snapshot(e->h[g], data.boxsize, data.massarr, p.hydro) : slicer_next_batch(e->h[g]);
        if (!rc)
          rc = stage_subfile(e->h[g], buf, p.hydro);
        if (!rc)
          rc = e->used[g] == 0 ? slicer_deposit(e->h[g], descs.data(), nj) : slicer_deposit_accumulate(e->h[g], descs.data(), nj);
        if (rc)
        {
          err << "GPU " << e->devices[g] << ", sub-file " << ff << ": " << slicer_last_error() << "\n";
          rcs[g] = 1;
          break;
        }
        e->used[g]++;
      }
      logs[g] = log.str();
      errs[g] = err.str();
    };
    if (ensure_capacity(e, File, ffmin, ffmax))
      return 1;
    if (ngpu == 1)
      work(0);
    else
    {
      std::vector<std::thread> th;
      for (int g = 0; g < ngpu; g++)
        th.emplace_back(work, g);
      for (auto &t : th)
        t.join();
    }
    for (int g = 0; g < ngpu; g++)
    {
      t_read += read_s[g] / ngpu;
      if (myid == 0)
        std::cout << logs[g];
      std::cerr << errs[g];
      if (rcs[g])
        return 1;
    }
    for (int g = 0; g < ngpu; g++)
      if (e->used[g] == 0)
      { // a GPU without sub-files still contributes zero maps to the sum (slicer-v2.cpp:162-175: ranks with no files)
        const double zero[6] = {0, 0, 0, 0, 0, 0};
        if (slicer_begin_snapshot(e->h[g], have_header ? first_header.boxsize : 1.0, have_header ? first_header.massarr : zero, p.hydro) ||
            slicer_deposit(e->h[g], descs.data(), nj))
          return fail_capi("slicer_deposit");
      }
    if (ngpu > 1 && slicer_reduce_all(e->h.data(), ngpu, nj, 0)) // replaces the 7 MPI_Reduce of slicer-v2.cpp:214-217
      return fail_capi("slicer_reduce_all");
    const double tf0 = now_s();
    for (int j = 0; j < nj; j++)
    {
      const int npix = jobs[j0 + j].npix;
      long long counts[6];
      mapxytot[j0 + j].resize((size_t)npix * npix);
      if (slicer_fetch(e->h[0], j, -1, &mapxytot[j0 + j][0], counts, nullptr))
        return fail_capi("slicer_fetch");
      for (int t = 0; t < 6; t++)
      {
        ntotxyi[(size_t)(j0 + j) * 6 + t] = counts[t];
        if (e->per_type)
        {
          mapxytoti[(size_t)(j0 + j) * 6 + t].resize((size_t)npix * npix);
          if (slicer_fetch(e->h[0], j, t, &mapxytoti[(size_t)(j0 + j) * 6 + t][0], nullptr, nullptr))
            return fail_capi("slicer_fetch");
        }
      }
    }
    t_fetch += now_s() - tf0;
  }
  return 0;
}

int createDensityMaps(Engine *e, InputParams &p, Lens &lens, Random &random, int isnap, unsigned ffmin, unsigned ffmax,
                      const std::string &File, double fovradiants, double rcase, std::valarray<float> &mapxytot,
                      std::valarray<float> (&mapxytoti)[6], int (&ntotxyi)[6], int myid)
{
  std::vector<PlaneJob> jobs = {{isnap, rcase, p.npix}};
  std::vector<std::valarray<float>> tot, per;
  std::vector<long long> cnt;
  if (createDensityMapsMulti(e, p, lens, random, jobs, ffmin, ffmax, File, fovradiants, tot, per, cnt, myid))
    return 1;
  mapxytot = tot[0];
  for (int i = 0; i < 6; i++)
  {
    if (e->per_type)
      mapxytoti[i] = per[i];
    else
      mapxytoti[i].resize((size_t)p.npix * p.npix); // zero-filled, as the reference's resize() leaves them
    ntotxyi[i] = (int)cnt[i];
  }
  return 0;
}

static std::string plane_label(int pll)
{ // slicer-v2.cpp:154-159
  char b[32];
  snprintf(b, sizeof(b), "%03d", pll);
  return b;
}

int runLightCone(const std::string &inifile, const RunOptions &opt)
{
  const int myid = opt.quiet ? 1 : 0;
  const double t_begin = now_s();
  t_read = t_submit = t_fetch = t_write = t_engine = 0;
  InputParams p;
  if (readInput(p, inifile))
    return 1;
  if (p.simType == "SubFind")
  {
    std::cerr << "npix == 0 selects the SubFind halo catalogue branch, which this build does not provide" << std::endl;
    return 1;
  }
  std::vector<std::string> snappath;
  std::vector<double> snapred, snapbox;
  if (readRedList(p.filredshiftlist, snapred, snappath, snapbox, p))
    return 1;
  Header simdata;
  if (readHeader(p.pathsnap + snappath[0] + ".0", simdata))
    return 1;
  testHydro(p, simdata); // decided once from the first listed snapshot (slicer-v2.cpp:74-76)
  CosmoTable cosmo;
  cosmo.build(simdata.om0, simdata.oml, p.w, p.zs);
  p.Ds = cosmo.getDl.eval(p.zs);
  Lens lens;
  if (buildPlanes(p, lens, snapred, snappath, snapbox, cosmo.getDl, cosmo.getZl, numberOfLensPerSnap, 0)) // rank 0 writes planes_list
    return 1;
  double fovradiants = 0;
  for (size_t i = 0; i < lens.ld.size(); i++)
  {
    if (i == 0)
      lens.nrepperp.assign(lens.ld.size(), 0);
    const double boxl = snapbox[lens.fromsnapi[i]] / 1e3 * POS_U;
    if (!opt.replication)
    {
      if (testFov(p.fov, boxl, lens.ld2[i], 0, fovradiants))
        return 1;
    }
    else
      computeReplications(p.fov, boxl, lens.ld2[i], myid, fovradiants, lens.nrepperp[i]);
  }
  Random random;
  randomizeBox(random, lens, p, numberOfLensPerSnap, 1, opt.fixed_vertex);

  // per-plane rcase and npix, exactly as the sequential loop of slicer-v2.cpp:137-185 would see them
  std::vector<double> rcase(lens.nplanes, 0.0);
  std::vector<int> npix(lens.nplanes, p.npix);
  float rc = 0.0f;
  int npix_max = 1;
  for (int isnap = 0; isnap < lens.nplanes; isnap++)
  {
    if (p.physical)
      npix[isnap] = int((lens.ld2[isnap] + lens.ld[isnap]) / 2 * fovradiants / p.rgrid * 1e3 / POS_U) + 1;
    if (lens.randomize[isnap])
      rc = lens.ld[isnap] / snapbox[lens.fromsnapi[isnap]] * 1e3 / POS_U;
    rcase[isnap] = rc;
    npix_max = std::max(npix_max, npix[isnap]);
  }
  if (p.partinplanes && myid == 0)
    std::cout << "!It is not possible to resume a Gadget run with partinplanes == true!" << std::endl
              << "!!               Files on Output folder will be overwritten        !!" << std::endl;

  Engine *e = nullptr;
  int status = 0;
  // The planes of a snapshot are written by a background thread while the next snapshot is read and deposited (an 8192^2
  // plane is 256 MiB of byte-swapped FITS).  At most one snapshot's planes are in flight.
  std::thread writer;
  int writer_status = 0;
  double writer_seconds = 0;
  auto join_writer = [&]() {
    if (writer.joinable())
    {
      writer.join();
      t_write += writer_seconds;
      if (writer_status)
        status = 1;
    }
  };
  for (int s0 = 0; s0 < lens.nplanes && !status;)
  {
    // planes [s0, s1) use the same snapshot
    int s1 = s0 + 1;
    while (s1 < lens.nplanes && lens.fromsnapi[s1] == lens.fromsnapi[s0])
      s1++;
    std::vector<PlaneJob> jobs;
    for (int isnap = s0; isnap < s1; isnap++)
    {
      if (!p.partinplanes && std::ifstream(fileOutput(p, plane_label(lens.pll[isnap]))))
      { // resume: this plane was already written (slicer-v2.cpp:188-196)
        if (myid == 0)
          std::cout << fileOutput(p, plane_label(lens.pll[isnap])) << " Already exists" << std::endl;
        continue;
      }
      jobs.push_back({isnap, rcase[isnap], npix[isnap]});
    }
    const std::string File = p.pathsnap + lens.fromsnap[s0];
    Header snapdata;
    if (readHeader(File + ".0", snapdata))
    {
      status = 1;
      break;
    }
    if (!jobs.empty())
    {
      if (!e)
      {
        size_t tot = 0;
        for (int i = 0; i < 6; i++)
          tot += (size_t)snapdata.npartTotal[i] + ((size_t)(uint32_t)snapdata.nTotalHW[i] << 32);
        const size_t cap = tot / std::max(1, snapdata.numfiles) * 5 / 4 + 65536;
        const double te0 = now_s();
        int most = 1; // the most planes one snapshot feeds
        for (int a = 0; a < lens.nplanes;)
        {
          int b = a + 1;
          while (b < lens.nplanes && lens.fromsnapi[b] == lens.fromsnapi[a])
            b++;
          most = std::max(most, b - a);
          a = b;
        }
        e = engineCreate(opt.devices, npix_max, opt.mas, p.partinplanes, cap, opt.deposit_mode, most);
        t_engine += now_s() - te0;
        if (!e)
        {
          status = 1;
          break;
        }
      }
      std::vector<std::valarray<float>> tot, per;
      std::vector<long long> cnt;
      if (createDensityMapsMulti(e, p, lens, random, jobs, 0, snapdata.numfiles, File, fovradiants, tot, per, cnt, myid))
      {
        status = 1;
        break;
      }
      join_writer(); // the previous snapshot's planes
      if (status)
        break;
      struct Batch
      {
        std::vector<PlaneJob> jobs;
        std::vector<double> zsim;
        std::vector<std::valarray<float>> tot, per;
        std::vector<long long> cnt;
        Header snapdata;
      };
      auto batch = std::make_shared<Batch>();
      batch->jobs = jobs;
      for (size_t j = 0; j < jobs.size(); j++)
        batch->zsim.push_back(cosmo.getZl.eval((lens.ld2[jobs[j].isnap] + lens.ld[jobs[j].isnap]) / 2.0));
      batch->tot = std::move(tot);
      batch->per = std::move(per);
      batch->cnt = std::move(cnt);
      batch->snapdata = snapdata;
      writer_status = 0;
      writer_seconds = 0;
      writer = std::thread([batch, &p, &lens, &writer_status, &writer_seconds]() {
        const double tw0 = now_s();
        for (size_t j = 0; j < batch->jobs.size(); j++)
        {
          const int isnap = batch->jobs[j].isnap;
          InputParams pj = p;
          pj.npix = batch->jobs[j].npix;
          try
          {
            writeMaps(pj, batch->snapdata, lens, isnap, batch->zsim[j], plane_label(lens.pll[isnap]), batch->tot[j],
                      p.partinplanes ? &batch->per[j * 6] : nullptr, &batch->cnt[j * 6], 0);
          }
          catch (const SliceError &err)
          {
            std::cerr << "It was not possible to create the map: " << err.what << std::endl;
            writer_status = 1;
            break;
          }
        }
        writer_seconds = now_s() - tw0;
      });
    }
    s0 = s1;
  }
  join_writer();
  engineDestroy(e);
  if (getenv("SLICER_B200_TIMING"))
    std::cerr << "[timing] total " << now_s() - t_begin << " s: engine/CUDA set-up " << t_engine << ", sub-file reads " << t_read << ", fetch (waits for the GPU) " << t_fetch
              << ", FITS writes (background thread) " << t_write << std::endl;
  return status;
}

} // namespace slicer
