// deposit_degrade.cuh — `Part. Degradation` (InputParams.snopt > 0), densitymaps.cpp:387-397.
//
// For every accepted (particle, replica) pair, in the order mapParticles meets them (type by type, particle by
// particle, ni then nj), the reference draws one libc rand(): the pair keeps mass m * 2^snopt if
// rand()/float(RAND_MAX) < 2^-snopt, else 0 (it is still counted and "deposited" with zero mass).  The draws are one
// serial stream, so they are made on the host (slicer_host: degradeDraws) and handed over as one byte per accepted
// pair; the device only has to know the RANK of each accepted pair in the reference's order:
//   degrade_kernel<MARK>     one thread per particle runs the exact chain and writes how many of its replicas each
//                            plane accepts (a byte per particle and plane)
//   scan_* kernels           exclusive prefix sum over the concatenated particles of the batch, per plane
//   degrade_kernel<DEPOSIT>  the same chain again; pair number r of particle i looks up keep[rank[i] + r]
// This mode trades speed for strict reproducibility of the reference's random stream (one thread per particle, map
// atomics straight to global memory); the fast pipelined passes are for snopt == 0.
#pragma once
#include "device_chain.cuh"

namespace degrade
{

struct Dev
{
  unsigned char *cnt;          // [nplanes][stride] accepted replicas per particle
  unsigned *rank;              // [nplanes][stride + 1] exclusive prefix sums of cnt
  const unsigned char *keep;   // all planes' keep bytes, plane q starts at keep_off[q]
  unsigned long long keep_off[SLICER_MAX_PLANES];
  unsigned long long stride;   // particles the arrays hold per plane
  unsigned long long seg_base; // index of this segment's first particle in the batch
  int snopt;
};

enum { MARK = 0, DEPOSIT = 1 };

template <int MAS, int MODE>
__global__ void __launch_bounds__(256) degrade_kernel(const __grid_constant__ PassParams P, const __grid_constant__ SegmentDev S,
                                                      const __grid_constant__ Dev G)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += stride)
  {
    float r0, r1, r2;
    if (S.layout == SLICER_LAYOUT_AOS)
    {
      const float *p = S.pos + 3ull * i;
      r0 = __ldg(p);
      r1 = __ldg(p + 1);
      r2 = __ldg(p + 2);
    }
    else
    {
      r0 = __ldg(S.pos + i);
      r1 = __ldg(S.pos + S.soa_stride + i);
      r2 = __ldg(S.pos + 2ull * S.soa_stride + i);
    }
    for (int t = 0; t < P.nxform; t++)
    {
      const XformDev &X = P.xf[t];
      const float z = chain::box_axis(2, r0, r1, r2, X);
      if (!(z >= X.zmin && z < X.zmax))
        continue;
      float x = 0.f, y = 0.f, m = 0.f;
      bool have_xy = false;
      for (int q = X.first_plane; q < X.first_plane + X.nplanes; q++)
      {
        const PlaneDev &L = P.pl[q];
        if (!chain::in_slab(z, L))
          continue;
        if (!have_xy)
        {
          x = chain::box_axis(0, r0, r1, r2, X);
          y = chain::box_axis(1, r0, r1, r2, X);
          m = chain::particle_mass(S, i);
          have_xy = true;
        }
        unsigned a = 0, g = 0;
        const unsigned long long slot = (unsigned long long)L.slot * (G.stride + (MODE == DEPOSIT ? 1 : 0)) + G.seg_base + i;
        const unsigned long long first = MODE == DEPOSIT ? G.keep_off[L.slot] + G.rank[slot] : 0ull;
        unsigned long long *map = L.acc + L.type_stride * (unsigned long long)S.type;
        for (int ni = -L.nrep; ni <= L.nrep; ni++)
          for (int nj = -L.nrep; nj <= L.nrep; nj++)
          {
            if (!chain::prefilter(x, y, z, ni, nj, L))
              continue;
            float xs, ys;
            if (chain::project_accept(x, y, z, ni, nj, L, xs, ys))
            {
              if (MODE == DEPOSIT)
              {
                // densitymaps.cpp:393-396: kept pairs weigh 2^snopt m, the others 0 (and add nothing to the map)
                const float mm = G.keep[first + a] ? ldexpf(m, G.snopt) : 0.f;
                const int gx = chain::grid_index(xs, L), gy = chain::grid_index(ys, L);
                if (gx >= 0 && gx < L.npix && gy >= 0 && gy < L.npix)
                  g++;
                if (mm != 0.f)
                  chain::deposit<MAS>(xs, ys, mm, L, map);
              }
              a++;
            }
          }
        if (MODE == MARK)
        {
          if (a)
            G.cnt[slot] = (unsigned char)(a > 255 ? 255 : a);
        }
        else
        {
          if (a)
            atomicAdd(L.counts + 2 * S.type, (unsigned long long)a);
          if (g)
            atomicAdd(L.counts + 2 * S.type + 1, (unsigned long long)g);
        }
      }
    }
  }
}

// ---- exclusive prefix sum of bytes -> u32, per plane (grid.y), three small kernels ---------------------------
constexpr int SCAN_BLOCK = 1024;
constexpr int SCAN_PER = 8; // elements per thread

__global__ void __launch_bounds__(SCAN_BLOCK) scan_block_sums(const unsigned char *cnt, unsigned long long stride, unsigned long long n, unsigned *block_sums,
                                                              unsigned long long nblocks)
{
  __shared__ unsigned wsum[32];
  const unsigned char *c = cnt + (unsigned long long)blockIdx.y * stride;
  const unsigned long long base = ((unsigned long long)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_PER;
  unsigned s = 0;
  for (int j = 0; j < SCAN_PER; j++)
    if (base + j < n)
      s += c[base + j];
  s = __reduce_add_sync(0xffffffffu, s);
  if ((threadIdx.x & 31) == 0)
    wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32)
  {
    const unsigned v = __reduce_add_sync(0xffffffffu, wsum[threadIdx.x]);
    if (threadIdx.x == 0)
      block_sums[(unsigned long long)blockIdx.y * nblocks + blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(1024) scan_of_block_sums(unsigned *block_sums, unsigned long long nblocks)
{
  __shared__ unsigned wsum[32];
  __shared__ unsigned carry_s;
  unsigned *b = block_sums + (unsigned long long)blockIdx.y * nblocks;
  if (threadIdx.x == 0)
    carry_s = 0;
  __syncthreads();
  for (unsigned long long base = 0; base < nblocks; base += 1024)
  {
    const unsigned long long i = base + threadIdx.x;
    const unsigned x = i < nblocks ? b[i] : 0u;
    unsigned incl = x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d)
        incl += y;
    }
    if (lane == 31)
      wsum[w] = incl;
    __syncthreads();
    if (w == 0)
    {
      const unsigned sv = wsum[lane];
      unsigned si = sv;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const unsigned y = __shfl_up_sync(0xffffffffu, si, d);
        if (lane >= d)
          si += y;
      }
      wsum[lane] = si - sv;
    }
    __syncthreads();
    const unsigned carry = carry_s;
    const unsigned excl = carry + wsum[w] + incl - x;
    if (i < nblocks)
      b[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023)
      carry_s = excl + x;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_finish(const unsigned char *cnt, unsigned long long stride, unsigned long long n, const unsigned *block_sums,
                                                          unsigned long long nblocks, unsigned *rank)
{
  __shared__ unsigned wsum[32];
  const unsigned char *c = cnt + (unsigned long long)blockIdx.y * stride;
  unsigned *r = rank + (unsigned long long)blockIdx.y * (stride + 1);
  const unsigned long long base = ((unsigned long long)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_PER;
  unsigned v[SCAN_PER], s = 0;
  for (int j = 0; j < SCAN_PER; j++)
  {
    v[j] = base + j < n ? c[base + j] : 0u;
    s += v[j];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const unsigned y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d)
      incl += y;
  }
  if (lane == 31)
    wsum[w] = incl;
  __syncthreads();
  if (w == 0)
  {
    const unsigned sv = wsum[lane];
    unsigned si = sv;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned y = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d)
        si += y;
    }
    wsum[lane] = si - sv;
  }
  __syncthreads();
  unsigned e = block_sums[(unsigned long long)blockIdx.y * nblocks + blockIdx.x] + wsum[w] + incl - s;
  for (int j = 0; j < SCAN_PER; j++)
  {
    if (base + j <= n) // element n holds the total
      r[base + j] = e;
    e += v[j];
  }
}

} // namespace degrade
