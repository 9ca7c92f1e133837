// probe_h2d.cu — measurement aid: the box's ceiling for concurrent pinned-host -> device copies, the limit of the end-to-end
// path (12 B per particle over PCIe).  For n = 1, 2, 4, ... GPUs, n host threads (one per GPU) each copy a 1 GiB page-locked
// buffer to their GPU `reps` times with cudaMemcpyAsync; reports the aggregate and the slowest per-GPU rate as one JSON line.
//   nvcc -O2 -o tools/_build/probe_h2d tools/probe_h2d.cu && tools/_build/probe_h2d [reps=8]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
  const int reps = argc > 1 ? atoi(argv[1]) : 8;
  const size_t bytes = (size_t)1 << 30;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    printf("{\"error\": \"no CUDA device\"}\n");
    return 1;
  }
  std::vector<void *> hbuf(ndev, nullptr), dbuf(ndev, nullptr);
  std::vector<cudaStream_t> st(ndev);
  for (int g = 0; g < ndev; g++)
  {
    cudaSetDevice(g);
    if (cudaHostAlloc(&hbuf[g], bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&dbuf[g], bytes) != cudaSuccess)
    {
      printf("{\"error\": \"allocation failed on device %d\"}\n", g);
      return 1;
    }
    memset(hbuf[g], g + 1, bytes); // first touch by this thread
    cudaStreamCreate(&st[g]);
  }
  printf("{\"bytes_per_copy\": %zu, \"reps\": %d, \"runs\": [", bytes, reps);
  bool first = true;
  for (int n = 1; n <= ndev; n *= 2)
  {
    std::vector<double> secs(n, 0.0);
    std::atomic<int> ready(0);
    std::vector<std::thread> th;
    const double t0 = now();
    for (int g = 0; g < n; g++)
      th.emplace_back([&, g]() {
        cudaSetDevice(g);
        cudaMemcpyAsync(dbuf[g], hbuf[g], bytes, cudaMemcpyHostToDevice, st[g]); // warm-up
        cudaStreamSynchronize(st[g]);
        ready++;
        while (ready.load() < n)
          std::this_thread::yield();
        const double a = now();
        for (int r = 0; r < reps; r++)
          cudaMemcpyAsync(dbuf[g], hbuf[g], bytes, cudaMemcpyHostToDevice, st[g]);
        cudaStreamSynchronize(st[g]);
        secs[g] = now() - a;
      });
    for (auto &t : th)
      t.join();
    (void)t0;
    double slow = 0, agg = 0;
    for (int g = 0; g < n; g++)
    {
      slow = secs[g] > slow ? secs[g] : slow;
      agg += (double)bytes * reps / secs[g];
    }
    printf("%s{\"gpus\": %d, \"aggregate_GBps\": %.1f, \"slowest_gpu_GBps\": %.1f}", first ? "" : ", ", n, agg / 1e9, (double)bytes * reps / slow / 1e9);
    first = false;
  }
  printf("]}\n");
  return 0;
}
