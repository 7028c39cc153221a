// smembench.cu -- throughput of shared-memory atomics on B200 (float CAS-loop vs native int / u64), per warp-instruction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template<int MODE> __global__ void k(const int* __restrict__ idx, float* out, int iters, int tile){
  extern __shared__ unsigned long long s64[];
  float* sf=(float*)s64; int* si=(int*)s64;
  for(int i=threadIdx.x;i<tile*2;i+=blockDim.x) si[i]=0;
  __syncthreads();
  int base=idx[blockIdx.x*blockDim.x+threadIdx.x];
  float v=1.0f+threadIdx.x*1e-3f;
  for(int it=0;it<iters;it++){
    int a=(base+it*37)%tile;
    if(MODE==0) atomicAdd(&sf[a],v);
    else if(MODE==1) atomicAdd(&si[a],(int)(v*1024));
    else if(MODE==2) atomicAdd(&s64[a],(unsigned long long)(v*1048576));
    else if(MODE==3){ sf[a]+=v; }   // racy plain RMW: lower bound
    else if(MODE==4){ asm volatile("red.shared.add.f32 [%0], %1;"::"r"((unsigned)__cvta_generic_to_shared(&sf[a])),"f"(v):"memory"); }
  }
  __syncthreads();
  float acc=0; for(int i=threadIdx.x;i<tile;i+=blockDim.x) acc+= (MODE==2)? (float)s64[i] : (MODE==1? (float)si[i]: sf[i]);
  if(acc==12345.f) out[0]=acc;
}
int main(){
  int tile=4096, iters=2000, blocks=148*4, threads=256;
  int n=blocks*threads; int* h=(int*)malloc(4*n); int* d; float* out; cudaMalloc(&d,4*n); cudaMalloc(&out,4);
  const char* names[5]={"float atomicAdd (CAS loop)","int atomicAdd (native)","u64 atomicAdd (native)","plain racy RMW","red.shared.add.f32 PTX"};
  for(int pattern=0;pattern<3;pattern++){
    for(int i=0;i<n;i++){ int t=i%threads; h[i]= pattern==0? t : pattern==1? (t*17+ (rand()%3)) : rand()%tile; }
    cudaMemcpy(d,h,4*n,cudaMemcpyHostToDevice);
    printf("--- pattern %s\n", pattern==0?"consecutive (conflict-free)":pattern==1?"stride 17 + small jitter":"random in 4096");
    for(int mode=0;mode<5;mode++){
      cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e9;
      for(int r=0;r<4;r++){ cudaEventRecord(e0);
        size_t sm=tile*8;
        if(mode==0) k<0><<<blocks,threads,sm>>>(d,out,iters,tile); else if(mode==1) k<1><<<blocks,threads,sm>>>(d,out,iters,tile);
        else if(mode==2) k<2><<<blocks,threads,sm>>>(d,out,iters,tile); else if(mode==3) k<3><<<blocks,threads,sm>>>(d,out,iters,tile); else k<4><<<blocks,threads,sm>>>(d,out,iters,tile);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      double warp_instr=(double)blocks*threads/32*iters; double per_sm=warp_instr/148;
      // cycles per warp-instruction per SM at ~1.9 GHz
      printf("%-28s %.3f ms  -> %.2f cycles/warp-instr/SM (@1.9GHz), %.1f G lane-ops/s chip\n",names[mode],best,best*1e-3*1.9e9/per_sm,(double)blocks*threads*iters/best/1e6);
    }
  }
  cudaError_t e=cudaGetLastError(); if(e) printf("err %s\n",cudaGetErrorString(e));
  return 0;
}
