// aux_kernels.cuh — small helper kernels around the deposit: accumulator read-out and synthetic inputs.
#pragma once
#include <stdint.h>

// int64 fixed point -> float32 map, summing `ntypes` per-type accumulators first (exact integer sum).
// Replaces the float adds of densitymaps.cpp:511-513 and the float MPI_Reduce of slicer-v2.cpp:214-217:
// one rounding per pixel instead of one per particle.
__global__ void finalize_map_kernel(const unsigned long long *__restrict__ acc, unsigned long long type_stride, int ntypes,
                                    unsigned long long npix2, double inv_scale, float *__restrict__ out)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix2; i += stride)
  {
    long long s = 0;
    for (int t = 0; t < ntypes; t++)
      s += (long long)acc[(unsigned long long)t * type_stride + i];
    out[i] = __double2float_rn(__dmul_rn((double)s, inv_scale));
  }
}

__global__ void sum_types_kernel(const unsigned long long *__restrict__ acc, unsigned long long type_stride, int ntypes,
                                 unsigned long long npix2, long long *__restrict__ out)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix2; i += stride)
  {
    long long s = 0;
    for (int t = 0; t < ntypes; t++)
      s += (long long)acc[(unsigned long long)t * type_stride + i];
    out[i] = s;
  }
}

// Counter-based generator: 24-bit uniform from splitmix64(seed, 3*i+k), scaled by the box (float multiply).
// slicer_b200/synth_hash.py restates it in numpy so hosts can regenerate any chunk.
__host__ __device__ inline uint64_t synth_mix(uint64_t seed, uint64_t idx)
{
  uint64_t z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void synth_positions_kernel(float *__restrict__ pos, unsigned long long n, unsigned long long soa_stride,
                                       int layout_soa, uint64_t seed, float boxf)
{
  const unsigned long long total = 3ull * n;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride)
  {
    unsigned long long i, k;
    if (layout_soa)
    {
      k = j / n;
      i = j - k * n;
    }
    else
    {
      i = j / 3ull;
      k = j - 3ull * i;
    }
    const float u = (float)(synth_mix(seed, 3ull * i + k) >> 40) * (1.0f / 16777216.0f);
    const float v = __fmul_rn(u, boxf);
    if (layout_soa)
      pos[k * soa_stride + i] = v;
    else
      pos[j] = v;
  }
}
