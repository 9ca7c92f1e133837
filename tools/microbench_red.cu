// tools/microbench_red.cu — measurement aid (not product): throughput of the TSC 3x3 scatter as a function of
// the map footprint and atomic flavour.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb microbench_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__host__ __device__ inline uint64_t mix(uint64_t z){ z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
template<int MODE> __global__ void k(unsigned long long* map, int npix, unsigned long long nrec, int band_rows){
  unsigned long long stride=(unsigned long long)gridDim.x*blockDim.x;
  for(unsigned long long i=(unsigned long long)blockIdx.x*blockDim.x+threadIdx.x;i<nrec;i+=stride){
    uint64_t h=mix(i*0x9E3779B97F4A7C15ull+12345);
    int gx=1+(int)((h&0xffffffff)%(unsigned)(npix-2));
    int gy=1+(int)((h>>32)%(unsigned)(band_rows-2));
    long long q=(long long)(h>>20)|1;
    if(MODE==0){ // 9 x red.u64 row-major
      #pragma unroll
      for(int jy=-1;jy<=1;jy++)
      #pragma unroll
      for(int jx=-1;jx<=1;jx++)
        asm volatile("red.global.add.u64 [%0], %1;"::"l"(map+(size_t)(gy+jy)*npix+gx+jx),"l"(q):"memory");
    } else if(MODE==1){ // 1 x red.u64 (NGP)
        asm volatile("red.global.add.u64 [%0], %1;"::"l"(map+(size_t)gy*npix+gx),"l"(q):"memory");
    } else if(MODE==2){ // 9 x red.f32
      float* m=(float*)map; float v=(float)(h&0xff);
      #pragma unroll
      for(int jy=-1;jy<=1;jy++)
      #pragma unroll
      for(int jx=-1;jx<=1;jx++)
        asm volatile("red.global.add.f32 [%0], %1;"::"l"(m+(size_t)(gy+jy)*npix+gx+jx),"f"(v):"memory");
    } else if(MODE==3){ // tiled 4x4 int64 (128 B tiles)
      #pragma unroll
      for(int jy=-1;jy<=1;jy++)
      #pragma unroll
      for(int jx=-1;jx<=1;jx++){
        int x=gx+jx,y=gy+jy;
        size_t off=((size_t)(y>>2)*(npix>>2)+(x>>2))*16+((y&3)<<2)+(x&3);
        asm volatile("red.global.add.u64 [%0], %1;"::"l"(map+off),"l"(q):"memory");
      }
    } else if(MODE==4){ // 9 x red.u32
      unsigned* m=(unsigned*)map;
      #pragma unroll
      for(int jy=-1;jy<=1;jy++)
      #pragma unroll
      for(int jx=-1;jx<=1;jx++)
        asm volatile("red.global.add.u32 [%0], %1;"::"l"(m+(size_t)(gy+jy)*npix+gx+jx),"r"((unsigned)q):"memory");
    } else if(MODE==5){ // 3 x red.v4.f32 (one aligned quad per row)
      float* m=(float*)map; float v=(float)(h&0xff);
      #pragma unroll
      for(int jy=-1;jy<=1;jy++){
        size_t off=((size_t)(gy+jy)*npix+gx)&~(size_t)3;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"::"l"(m+off),"f"(v),"f"(v),"f"(v),"f"(v):"memory");
      }
    }
  }
}
template<int MODE> void run(const char* name,unsigned long long* map,int npix,int band_rows,unsigned long long nrec){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148*8,256>>>(map,npix,nrec/8,band_rows); 
  cudaEventRecord(a); k<MODE><<<148*8,256>>>(map,npix,nrec,band_rows); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  cudaError_t e=cudaGetLastError();
  printf("%-28s npix %5d band_rows %5d footprint %7.1f MiB  %8.3f ms  %7.2f Grec/s %s\n",name,npix,band_rows,(double)npix*band_rows*8/1048576.0,ms,nrec/ms/1e6,e?cudaGetErrorString(e):"");
}
int main(){
  unsigned long long* map; size_t bytes=(size_t)8192*8192*8*4; cudaMalloc(&map,bytes); cudaMemset(map,0,bytes);
  unsigned long long nrec=1ull<<28;
  int cfg[][2]={{256,256},{1024,1024},{2048,512},{2048,1024},{2048,2048},{4096,2048},{4096,4096},{8192,512},{8192,1024},{8192,8192},{16384,8192}};
  for(auto&c:cfg){
    run<0>("9x red.u64 rowmajor",map,c[0],c[1],nrec);
    run<3>("9x red.u64 tiled4x4",map,c[0],c[1],nrec);
    run<1>("1x red.u64 (NGP)",map,c[0],c[1],nrec);
    run<4>("9x red.u32",map,c[0],c[1],nrec);
    run<2>("9x red.f32",map,c[0],c[1],nrec);
    run<5>("3x red.v4.f32",map,c[0],c[1],nrec);
  }
  return 0;
}
