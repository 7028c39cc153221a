// fftbench.cu -- cuFFT 256^3 R2C/C2R: planar batch-3 vs float4-interleaved (stride 4) layouts, and (de)interleave passes.
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)
#define CF(x) do{cufftResult r=(x); if(r!=CUFFT_SUCCESS){printf("cufft %d at %d\n",(int)r,__LINE__);exit(1);} }while(0)
__global__ void interleave3(const float* __restrict__ in, float4* __restrict__ out, long n){
  for(long i=blockIdx.x*(long)blockDim.x+threadIdx.x;i<n;i+=(long)gridDim.x*blockDim.x) out[i]=make_float4(in[i],in[n+i],in[2*n+i],0.f);
}
__global__ void deinterleave3(const float4* __restrict__ in, float* __restrict__ out, long n){
  for(long i=blockIdx.x*(long)blockDim.x+threadIdx.x;i<n;i+=(long)gridDim.x*blockDim.x){ float4 v=in[i]; out[i]=v.x; out[n+i]=v.y; out[2*n+i]=v.z; }
}
template<class F> float timeit(F f, float* flush, size_t flushn){
  float best=1e9; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int r=0;r<6;r++){ CK(cudaMemsetAsync(flush,0,flushn)); cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r>=1&&ms<best)best=ms; }
  return best;
}
int main(){
  int n=256, nzc=n/2+1; long N=(long)n*n*n, Nc=(long)n*n*nzc;
  float *r4,*flush; cufftComplex* c; size_t flushn=256u<<20;
  CK(cudaMalloc(&r4,16*N)); CK(cudaMalloc(&c,8*Nc*4)); CK(cudaMalloc(&flush,flushn)); CK(cudaMemset(r4,0,16*N)); CK(cudaMemset(c,0,8*Nc*4));
  float* r4b; CK(cudaMalloc(&r4b,16*N));
  int dims[3]={n,n,n}; int re[3]={n,n,n}, ce[3]={n,n,nzc};
  cufftHandle p;
  auto plan=[&](cufftType t,int batch,int rstride,int rdist,int cstride,int cdist){ cufftHandle h; size_t ws;
    if(t==CUFFT_R2C) CF(cufftPlanMany(&h,3,dims,re,rstride,rdist,ce,cstride,cdist,t,batch)); else CF(cufftPlanMany(&h,3,dims,ce,cstride,cdist,re,rstride,rdist,t,batch));
    return h; };
  struct Case{const char* name; cufftType t; int batch,rs,rd,cs,cd;};
  Case cases[]={ {"R2C planar b1",CUFFT_R2C,1,1,(int)N,1,(int)Nc}, {"R2C planar b3",CUFFT_R2C,3,1,(int)N,1,(int)Nc},
    {"R2C real-interleaved4 b3 (istride 4)",CUFFT_R2C,3,4,1,1,(int)Nc}, {"R2C real-interleaved4 b4",CUFFT_R2C,4,4,1,1,(int)Nc},
    {"R2C both interleaved b3 (cstride 3)",CUFFT_R2C,3,4,1,3,1},
    {"C2R planar b1",CUFFT_C2R,1,1,(int)N,1,(int)Nc}, {"C2R planar b3",CUFFT_C2R,3,1,(int)N,1,(int)Nc},
    {"C2R real-interleaved4 b3 (ostride 4)",CUFFT_C2R,3,4,1,1,(int)Nc}, {"C2R real-interleaved4 b4",CUFFT_C2R,4,4,1,1,(int)Nc},
    {"C2R real-interleaved4 b1 (one channel)",CUFFT_C2R,1,4,1,1,(int)Nc} };
  for(auto& cs:cases){ p=plan(cs.t,cs.batch,cs.rs,cs.rd,cs.cs,cs.cd);
    float t=timeit([&]{ if(cs.t==CUFFT_R2C) cufftExecR2C(p,r4,c); else cufftExecC2R(p,c,r4); },flush,flushn);
    printf("%-44s %.3f ms\n",cs.name,t); cufftDestroy(p); }
  float t=timeit([&]{interleave3<<<148*8,256>>>(r4,(float4*)r4b,N);},flush,flushn); printf("%-44s %.3f ms\n","interleave3 kernel (12N -> 16N)",t);
  t=timeit([&]{deinterleave3<<<148*8,256>>>((float4*)r4,r4b,N);},flush,flushn); printf("%-44s %.3f ms\n","deinterleave3 kernel (16N -> 12N)",t);
  cudaError_t e=cudaGetLastError(); if(e) printf("err %s\n",cudaGetErrorString(e));
  return 0;
}
