// tools/microbench_atoms.cu — measurement aid (not product): throughput of 32-bit shared-memory atomicAdd on a 168x168
// tile as the tile deposit kernel issues them (9 cells x {lo with return, hi without}), with the base cell of the 32
// lanes of a warp (a) random, (b) random but in 32 distinct banks.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mba tools/microbench_atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TW = 168, TCELLS = TW * TW;
__device__ inline uint64_t mix(uint64_t z){ z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
template <int MODE> __global__ void __launch_bounds__(1024, 1) k(unsigned *out, int iters)
{
  extern __shared__ unsigned sm[];
  unsigned *lo = sm, *hi = sm + TCELLS;
  for (int i = threadIdx.x; i < 2 * TCELLS; i += 1024) sm[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++)
  {
    uint64_t h = mix(((uint64_t)blockIdx.x * 1024 + threadIdx.x) * 0x9E3779B97F4A7C15ull + it);
    int lx = (int)((h & 0xffff) % 166u), ly = (int)(((h >> 16) & 0xffff) % 166u);
    int c = ly * TW + lx;
    if (MODE == 1) // force bank(c) == lane: move lx by < 32 (wraps inside the row; keeps the 3x3 inside the tile)
    {
      int d = (lane - (c & 31)) & 31;
      lx = lx + d < 166 ? lx + d : lx + d - 32 >= 0 ? lx + d - 32 : lx;
      c = ly * TW + lx;
    }
    const unsigned vl = (unsigned)(h >> 32), vh = (unsigned)(h >> 56);
#pragma unroll
    for (int jy = 0; jy < 3; jy++)
#pragma unroll
      for (int jx = 0; jx < 3; jx++)
      {
        const unsigned old = atomicAdd(lo + c + jy * TW + jx, vl);
        atomicAdd(hi + c + jy * TW + jx, vh + ((old + vl < old) ? 1u : 0u));
      }
  }
  __syncthreads();
  unsigned s = 0;
  for (int i = threadIdx.x; i < 2 * TCELLS; i += 1024) s += sm[i];
  if (s == 0xdeadbeef) out[0] = s;
}
template <int MODE> void run(const char *name, unsigned *out)
{
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCELLS * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 2000;
  k<MODE><<<148, 1024, TCELLS * 8>>>(out, 10);
  cudaEventRecord(a); k<MODE><<<148, 1024, TCELLS * 8>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double rec = 148.0 * 1024 * iters;
  printf("%-28s %8.3f ms  %7.2f G records/s  %6.2f cycles per warp-atomic (at 1.9 GHz)  %s\n", name, ms, rec / ms / 1e6,
         ms * 1e-3 * 1.9e9 / (iters * 32.0 * 18.0), cudaGetErrorString(cudaGetLastError()));
}
int main(){ unsigned *out; cudaMalloc(&out, 4); run<0>("random base cell", out); run<1>("bank-distinct base cells", out); return 0; }
