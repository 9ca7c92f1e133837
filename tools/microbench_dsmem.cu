// microbench_dsmem.cu — measurement aid for one design question: can the tile deposit (deposit_binned.cuh: two 32-bit limbs per
// cell, returning shared-memory atomics, carry from the old value) use a thread-block CLUSTER so that two CTAs own one large
// tile in distributed shared memory?  Half of the stencil atomics of a record would then land in the partner CTA's shared memory.
// The kernel below runs the tile kernel's inner loop (9 cells x 2 limbs per record, random cells of a 168x168 tile) with the
// cells in (mode 0) the CTA's own shared memory, (mode 1) the partner's, (mode 2) own or partner's by the cell's column.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/microbench_dsmem tools/microbench_dsmem.cu
//   prints records/s per SM and the ratio to mode 0.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

constexpr int TW = 168, TCELLS = TW * TW;

__device__ __forceinline__ unsigned mix(unsigned x)
{
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1) dsmem_tile_kernel(int records_per_thread, unsigned long long *sink)
{
  extern __shared__ unsigned tile[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  unsigned *mine = tile, *other = cluster.map_shared_rank(tile, rank ^ 1);
  for (int i = threadIdx.x; i < 2 * TCELLS; i += blockDim.x)
    tile[i] = 0;
  cluster.sync();
  unsigned h = mix(blockIdx.x * 1024u + threadIdx.x + 1u);
  for (int r = 0; r < records_per_thread; r++)
  {
    h = mix(h + r);
    const int lx = h % (TW - 2), ly = (h >> 12) % (TW - 2);
    const unsigned long long v = ((unsigned long long)(h & 0xffu) << 32) | mix(h);
#pragma unroll
    for (int jy = 0; jy < 3; jy++)
#pragma unroll
      for (int jx = 0; jx < 3; jx++)
      {
        const int c = (ly + jy) * TW + lx + jx;
        unsigned *base = MODE == 0 ? mine : (MODE == 1 ? other : ((lx + jx) < TW / 2 ? mine : other));
        const unsigned vl = (unsigned)v, vh = (unsigned)(v >> 32);
        const unsigned old = atomicAdd(base + c, vl);
        atomicAdd(base + TCELLS + c, vh + ((old + vl < old) ? 1u : 0u));
      }
  }
  cluster.sync();
  unsigned long long s = 0;
  for (int i = threadIdx.x; i < 2 * TCELLS; i += blockDim.x)
    s += tile[i];
  if (s == 0x123456789ull)
    *sink = s;
}

template <int MODE>
static double run(int rpt, unsigned long long *sink, int sms)
{
  cudaFuncSetAttribute(dsmem_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TCELLS * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int grid = sms / 2 * 2;
  dsmem_tile_kernel<MODE><<<grid, 1024, 2 * TCELLS * 4>>>(rpt / 8, sink);
  cudaEventRecord(a);
  dsmem_tile_kernel<MODE><<<grid, 1024, 2 * TCELLS * 4>>>(rpt, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  if (cudaGetLastError() != cudaSuccess)
    return -1;
  return (double)rpt * 1024 / (ms * 1e-3); // records per second per SM
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  unsigned long long *sink;
  cudaMalloc(&sink, 8);
  const int rpt = 4096;
  const double r0 = run<0>(rpt, sink, p.multiProcessorCount), r1 = run<1>(rpt, sink, p.multiProcessorCount), r2 = run<2>(rpt, sink, p.multiProcessorCount);
  printf("{\"records_per_s_per_sm\": {\"own_smem\": %.3e, \"partner_smem\": %.3e, \"half_and_half\": %.3e}, \"partner_over_own\": %.3f, "
         "\"half_over_own\": %.3f, \"sms\": %d}\n",
         r0, r1, r2, r1 / r0, r2 / r0, p.multiProcessorCount);
  return 0;
}
