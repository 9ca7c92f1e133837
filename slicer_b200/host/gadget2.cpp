// gadget2.cpp — GADGET format-2 sub-files, as the reference reads them:
//   every block = tag [i32 8]["NAME"][i32 size+8][i32 8] + payload [i32 size][bytes][i32 size]
//   readHeader (gadget2io.cpp:8-31) skips 5 int32 and raw-reads the 256-byte header; `<file>` without its ".N"
//   suffix is tried when `<file>.N` does not open (:16)
//   POS: float32 xyz triplets, the six types concatenated (:189-202)
//   MASS: float32 of the types with massarr == 0 in type order (densitymaps.cpp:358-370); the type-5 part is
//   skipped and read from "BHMA" instead (:361-365)
// The reference reads 4 bytes at a time through an ifstream (57 % of its run time); here every block is one bulk
// read straight into (optionally page-locked) memory, ready for slicer_stage_particles.  Payloads of 8 MiB and more are
// read by several threads at once (pread on disjoint ranges): one thread copies ~4 GB/s out of the page cache, a
// PCIe 5 x16 link takes 54 GB/s, so the serial read is what bounds a snapshot's end-to-end time.  SLICER_B200_IO_THREADS
// sets the thread count (default: the hardware threads, at most 8; 1 = plain fread).
#include "slicer_host.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <thread>
#include <vector>

#include <unistd.h>

namespace slicer
{

static FILE *open_subfile(const std::string &file_in)
{
  FILE *f = fopen(file_in.c_str(), "rb");
  if (!f && file_in.size() > 2)
    f = fopen(file_in.substr(0, file_in.size() - 2).c_str(), "rb");
  return f;
}

int readHeader(const std::string &file_in, Header &header)
{
  FILE *f = open_subfile(file_in);
  if (!f)
  {
    std::cerr << "Error in opening the file: " << file_in << "!\n\a";
    return 1;
  }
  int32_t blockheader[5];
  const bool ok = fread(blockheader, sizeof(blockheader), 1, f) == 1 && fread(&header, sizeof(Header), 1, f) == 1;
  fclose(f);
  if (!ok)
  {
    std::cerr << "Error in reading the header of: " << file_in << "!\n";
    return 1;
  }
  return 0;
}

SubFile::~SubFile()
{
  if (pinned)
  {
    if (pos)
      slicer_free_pinned(pos);
    if (mass)
      slicer_free_pinned(mass);
  }
  else
  {
    free(pos);
    free(mass);
  }
}

static int ensure_capacity(SubFile &s, size_t n, bool need_mass, bool pinned)
{
  if (s.capacity >= n && s.pos && (!need_mass || s.mass) && s.pinned == pinned)
    return 0;
  if (s.pinned)
  {
    if (s.pos)
      slicer_free_pinned(s.pos);
    if (s.mass)
      slicer_free_pinned(s.mass);
  }
  else
  {
    free(s.pos);
    free(s.mass);
  }
  s.pos = s.mass = nullptr;
  s.pinned = pinned;
  const size_t cap = n + n / 16 + 1024; // sub-files of one snapshot differ slightly in size: avoid re-allocating
  if (pinned)
  {
    void *p = nullptr, *m = nullptr;
    if (slicer_alloc_pinned(cap * 3 * sizeof(float), &p))
      return 1;
    s.pos = (float *)p;
    if (need_mass)
    {
      if (slicer_alloc_pinned(cap * sizeof(float), &m))
        return 1;
      s.mass = (float *)m;
    }
  }
  else
  {
    s.pos = (float *)malloc(cap * 3 * sizeof(float));
    s.mass = need_mass ? (float *)malloc(cap * sizeof(float)) : nullptr;
    if (!s.pos || (need_mass && !s.mass))
      return 1;
  }
  s.capacity = cap;
  return 0;
}

static int io_threads()
{
  static const int n = [] {
    const char *e = getenv("SLICER_B200_IO_THREADS");
    int v = e ? atoi(e) : 0;
    if (v <= 0)
    {
      v = (int)std::thread::hardware_concurrency();
      v = v > 8 ? 8 : v;
    }
    return v < 1 ? 1 : v;
  }();
  return n;
}

// fread(dst, 1, bytes, f) from the stream's current position, by several threads when the payload is large.
// Leaves the stream positioned after the bytes.  Returns true when all bytes were read.
static bool bulk_read(FILE *f, void *dst, size_t bytes)
{
  const int nthreads = io_threads();
  constexpr size_t PARALLEL_MIN = (size_t)8 << 20;
  if (nthreads == 1 || bytes < PARALLEL_MIN)
    return bytes == 0 || fread(dst, 1, bytes, f) == bytes;
  const off_t base = ftello(f);
  if (base < 0)
    return false;
  const int fd = fileno(f);
  const size_t mib = (size_t)1 << 20;
  const size_t chunk = ((bytes + nthreads - 1) / nthreads + mib - 1) / mib * mib;
  std::atomic<bool> ok(true);
  auto work = [&](size_t lo) {
    size_t hi = lo + chunk < bytes ? lo + chunk : bytes;
    while (lo < hi)
    {
      const ssize_t got = pread(fd, (char *)dst + lo, hi - lo, base + (off_t)lo);
      if (got <= 0)
      {
        ok = false;
        return;
      }
      lo += (size_t)got;
    }
  };
  std::vector<std::thread> pool;
  for (size_t lo = chunk; lo < bytes; lo += chunk)
    pool.emplace_back(work, lo);
  work(0);
  for (auto &t : pool)
    t.join();
  return ok && fseeko(f, base + (off_t)bytes, SEEK_SET) == 0;
}

// Positions the stream at the payload bytes of block `name` (searching forward from the current tag) and returns
// the payload size, or -1.  Equivalent of fastforwardToBlock (gadget2io.cpp:133-165) without its straddling struct.
static long long seek_block(FILE *f, const char *name)
{
  for (;;)
  {
    int32_t tag[4]; // [8]["NAME"][size+8][8]
    if (fread(tag, sizeof(tag), 1, f) != 1)
      return -1;
    int32_t size = 0;
    if (fread(&size, sizeof(size), 1, f) != 1)
      return -1;
    if (memcmp(&tag[1], name, 4) == 0)
      return (uint32_t)size;
    if (fseek(f, (long)(uint32_t)size + 4, SEEK_CUR))
      return -1;
  }
}

int readSubFile(const std::string &file, bool hydro, SubFile &out, bool pinned)
{
  FILE *f = open_subfile(file);
  if (!f)
  {
    std::cerr << "Error in opening the file: " << file << "!\n\a";
    return 1;
  }
  int rc = 1;
  do
  {
    if (seek_block(f, "HEAD") != (long long)sizeof(Header) || fread(&out.header, sizeof(Header), 1, f) != 1 || fseek(f, 4, SEEK_CUR))
    {
      std::cerr << "Error in reading the header of: " << file << "!\n";
      break;
    }
    const Header &h = out.header;
    size_t n = 0;
    for (int i = 0; i < 6; i++)
      n += (size_t)h.npart[i];
    out.ntotal = n;
    if (ensure_capacity(out, n, hydro, pinned)) // the mass buffer only where per-particle masses can occur
    {
      std::cerr << "Out of (pinned) host memory for " << n << " particles\n";
      break;
    }
    const long long psz = seek_block(f, "POS ");
    if (psz < (long long)(n * 12) || !bulk_read(f, out.pos, n * 12) || fseek(f, psz - (long long)n * 12 + 4, SEEK_CUR))
    {
      std::cerr << "POS block of " << file << " is missing or too short\n";
      break;
    }
    if (hydro)
    {
      // MASS: concatenation over the types with massarr == 0 (densitymaps.cpp:358-370)
      size_t nm = 0;
      for (int i = 0; i < 6; i++)
        if (h.massarr[i] == 0)
          nm += (size_t)h.npart[i];
      memset(out.mass, 0, n * sizeof(float));
      if (nm)
      {
        const long long msz = seek_block(f, "MASS");
        if (msz < 0)
        {
          std::cerr << "MASS block of " << file << " is missing\n";
          break;
        }
        size_t off = 0, remaining = (size_t)msz / 4;
        bool bad = false;
        for (int i = 0; i < 6 && !bad; i++)
        {
          const size_t ni = (size_t)h.npart[i];
          if (h.massarr[i] == 0 && ni)
          {
            const size_t take = ni < remaining ? ni : remaining;
            if (i < 5)
              bad = take != ni || !bulk_read(f, out.mass + off, ni * 4);
            else
              bad = fseek(f, (long)take * 4, SEEK_CUR) != 0; // type 5: skipped, read from BHMA below (:361-365)
            remaining -= take;
          }
          off += ni;
        }
        if (bad || fseek(f, (long)remaining * 4 + 4, SEEK_CUR))
        {
          std::cerr << "MASS block of " << file << " is too short\n";
          break;
        }
        if (h.massarr[5] == 0 && h.npart[5] > 0)
        {
          const long long bsz = seek_block(f, "BHMA");
          const size_t n5 = (size_t)h.npart[5];
          if (bsz < (long long)(n5 * 4) || !bulk_read(f, out.mass + (n - n5), n5 * 4))
          {
            std::cerr << "BHMA block of " << file << " is missing or too short\n";
            break;
          }
        }
      }
    }
    rc = 0;
  } while (0);
  fclose(f);
  return rc;
}

} // namespace slicer
