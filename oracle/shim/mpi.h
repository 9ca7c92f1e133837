/* TEST INFRASTRUCTURE — stand-in for <mpi.h> so the unmodified reference sources
 * under /root/reference/SLICER compile in an image without MPI.
 *
 * Covers exactly the calls SLICER/slicer-v2.cpp makes (lines 28-30, 35..297, 214-217,
 * 324-325): Init, Comm_size, Comm_rank, Abort, Reduce(MPI_FLOAT, MPI_SUM, root 0),
 * Barrier, Finalize.
 *
 * Ranks are plain processes started by a launcher that sets
 *   SLICER_SHIM_RANK, SLICER_SHIM_SIZE, SLICER_SHIM_SHM (path of a pre-sized file in /dev/shm)
 * With SIZE==1 (or unset) no shared memory is touched.  With SIZE>1 MPI_Reduce sums the
 * ranks' buffers in rank order on rank 0 (deterministic), through the mmap'd file:
 *   [int32 barrier_count][int32 barrier_sense][pad to 64 B][size x capacity floats]
 * capacity (floats per rank) = (file_size-64)/4/size.
 */
#ifndef SLICER_SHIM_MPI_H
#define SLICER_SHIM_MPI_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <sched.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_FLOAT 1
#define MPI_SUM 1
#define MPI_SUCCESS 0

struct ShimMpiState
{
  int rank, size;
  char *base;
  size_t bytes;
  size_t capacity; /* floats per rank */
  int local_sense;
};

inline ShimMpiState &shim_mpi_state()
{
  static ShimMpiState s = {0, 1, nullptr, 0, 0, 0};
  return s;
}

inline int MPI_Init(int *, char ***)
{
  ShimMpiState &s = shim_mpi_state();
  const char *r = getenv("SLICER_SHIM_RANK");
  const char *n = getenv("SLICER_SHIM_SIZE");
  const char *f = getenv("SLICER_SHIM_SHM");
  s.rank = r ? atoi(r) : 0;
  s.size = n ? atoi(n) : 1;
  if (s.size > 1)
  {
    if (!f)
    {
      fprintf(stderr, "shim mpi: SLICER_SHIM_SHM not set\n");
      exit(2);
    }
    int fd = open(f, O_RDWR);
    if (fd < 0)
    {
      perror("shim mpi: open shm");
      exit(2);
    }
    struct stat st;
    fstat(fd, &st);
    s.bytes = (size_t)st.st_size;
    s.base = (char *)mmap(nullptr, s.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (s.base == (char *)MAP_FAILED)
    {
      perror("shim mpi: mmap");
      exit(2);
    }
    s.capacity = (s.bytes - 64) / sizeof(float) / (size_t)s.size;
  }
  return MPI_SUCCESS;
}

inline int MPI_Comm_size(MPI_Comm, int *n)
{
  *n = shim_mpi_state().size;
  return MPI_SUCCESS;
}

inline int MPI_Comm_rank(MPI_Comm, int *r)
{
  *r = shim_mpi_state().rank;
  return MPI_SUCCESS;
}

inline int MPI_Abort(MPI_Comm, int code)
{
  fprintf(stderr, "shim mpi: MPI_Abort(%d) on rank %d\n", code, shim_mpi_state().rank);
  exit(code ? 1 : 0);
  return MPI_SUCCESS;
}

inline int MPI_Barrier(MPI_Comm)
{
  ShimMpiState &s = shim_mpi_state();
  if (s.size <= 1)
    return MPI_SUCCESS;
  volatile int32_t *count = (volatile int32_t *)s.base;
  volatile int32_t *sense = (volatile int32_t *)(s.base + 4);
  s.local_sense = !s.local_sense;
  if (__sync_add_and_fetch(count, 1) == s.size)
  {
    *count = 0;
    __sync_synchronize();
    *sense = s.local_sense;
  }
  else
  {
    while (*sense != s.local_sense)
      sched_yield();
  }
  __sync_synchronize();
  return MPI_SUCCESS;
}

inline int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype, MPI_Op, int root, MPI_Comm comm)
{
  ShimMpiState &s = shim_mpi_state();
  if (s.size <= 1)
  {
    memcpy(recvbuf, sendbuf, (size_t)count * sizeof(float));
    return MPI_SUCCESS;
  }
  if ((size_t)count > s.capacity)
  {
    fprintf(stderr, "shim mpi: reduce of %d floats exceeds capacity %zu\n", count, s.capacity);
    exit(2);
  }
  float *slots = (float *)(s.base + 64);
  memcpy(slots + (size_t)s.rank * s.capacity, sendbuf, (size_t)count * sizeof(float));
  MPI_Barrier(comm);
  if (s.rank == root)
  {
    float *out = (float *)recvbuf;
    for (int i = 0; i < count; i++)
      out[i] = 0.f;
    for (int r = 0; r < s.size; r++)
    {
      const float *in = slots + (size_t)r * s.capacity;
      for (int i = 0; i < count; i++)
        out[i] += in[i];
    }
  }
  MPI_Barrier(comm);
  return MPI_SUCCESS;
}

inline int MPI_Finalize() { return MPI_SUCCESS; }

#endif
