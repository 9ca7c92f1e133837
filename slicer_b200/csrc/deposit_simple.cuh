// deposit_simple.cuh — SLICER_KERNEL_SIMPLE: one thread per particle, grid-stride, plain global loads.
//
// The whole per-particle pipeline of the reference in one kernel:
//   gadget2io.cpp:204-270 (box transform) -> densitymaps.cpp:358-372 (mass) -> :374 (slab) ->
//   :377-383 (replicas, getPolar, FoV) -> utilities.cpp:66-94 (NGP / TSC deposit)
// It is the correctness baseline the pipelined kernel is checked against and the "before" of the ncu
// comparison in DESIGN.md; every warp pays for the rare accepted lanes (divergence), which the pipelined
// kernel removes by compaction.
#pragma once
#include "device_chain.cuh"

template <int MAS>
__global__ void __launch_bounds__(256) deposit_simple_kernel(const __grid_constant__ PassParams P,
                                                             const __grid_constant__ SegmentDev S, const __grid_constant__ DeferDev F)
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
        unsigned long long *map = L.acc + L.type_stride * (unsigned long long)S.type;
        unsigned long long *cnt = L.counts + 2 * S.type;
        for (int ni = -L.nrep; ni <= L.nrep; ni++)
          for (int nj = -L.nrep; nj <= L.nrep; nj++)
          {
            if (!chain::prefilter(x, y, z, ni, nj, L))
              continue;
            float xs, ys;
            bool amb = false;
            if (chain::project_accept(x, y, z, ni, nj, L, xs, ys, &amb))
            {
              atomicAdd(&cnt[0], 1ull); // mapParticles' totPartxyi (densitymaps.cpp:402)
              if (chain::deposit<MAS>(xs, ys, m, L, map))
                atomicAdd(&cnt[1], 1ull);
            }
            else if (amb) // within the rounding guard of a decision: the host's libm settles it
              chain::defer_push(F, __fadd_rn(x, (float)ni), __fadd_rn(y, (float)nj), z, m, q, S.type);
          }
      }
    }
  }
}
