// redbench.cu -- what bounds red.global.add on B200 for a CIC scatter in lattice order?  Development microbenchmark.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)

struct P { int n; };
__device__ __forceinline__ void cic(const float* __restrict__ pos, long p, int n, int& ix,int& iy,int& iz,float& fx,float& fy,float& fz){
  float x=pos[3*p],y=pos[3*p+1],z=pos[3*p+2];
  float bx=floorf(x),by=floorf(y),bz=floorf(z); fx=x-bx;fy=y-by;fz=z-bz;
  ix=((int)bx%n+n)%n; iy=((int)by%n+n)%n; iz=((int)bz%n+n)%n;
}
// A: 8 scalar REDs
__global__ void kA(const float* __restrict__ pos, float* mesh, long np, int n, int ncorner){
  for(long p=blockIdx.x*(long)blockDim.x+threadIdx.x;p<np;p+=(long)gridDim.x*blockDim.x){
    int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
    int c=0;
    for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
      for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
        float* row=mesh+((long)ia*n+ib)*n;
        for(int d=0;d<2;d++){int id=iz+d; if(id>=n)id-=n; if(c<ncorner) atomicAdd(row+id, wb*(d?fz:1-fz)); c++; }}}
  }
}
// B: z-pair merged across lanes by shuffle when lane k+1 sits in the next z cell of the same row
__global__ void kB(const float* __restrict__ pos, float* mesh, long np, int n){
  for(long p0=blockIdx.x*(long)blockDim.x+threadIdx.x;p0<((np+31)/32)*32;p0+=(long)gridDim.x*blockDim.x){
    bool act=p0<np; long p=act?p0:np-1;
    int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
    int lane=threadIdx.x&31;
    // predecessor lane info
    int pix=__shfl_up_sync(0xffffffffu,ix,1), piy=__shfl_up_sync(0xffffffffu,iy,1), piz=__shfl_up_sync(0xffffffffu,iz,1);
    bool pact=__shfl_up_sync(0xffffffffu,(int)act,1);
    bool chain = lane>0 && pact && act && pix==ix && piy==iy && piz+1==iz;   // my z cell == pred's z+1 cell
    bool nextchain = __shfl_down_sync(0xffffffffu,(int)chain,1) && lane<31;  // successor merges my upper deposit
    for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
      for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
        float lo=wb*(1-fz), hi=wb*fz;
        float phi=__shfl_up_sync(0xffffffffu,hi,1);
        if(!act){lo=0;hi=0;}
        if(chain) lo+=phi;
        float* row=mesh+((long)ia*n+ib)*n;
        if(act) atomicAdd(row+iz, lo);
        if(act && !nextchain){int id=iz+1; if(id>=n)id-=n; atomicAdd(row+id, hi);}
      }}
  }
}
// C: red.v2 on the (z, z+1) pair when 8-byte aligned, scalar otherwise
__global__ void kC(const float* __restrict__ pos, float* mesh, long np, int n){
  for(long p=blockIdx.x*(long)blockDim.x+threadIdx.x;p<np;p+=(long)gridDim.x*blockDim.x){
    int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
    for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
      for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
        float* row=mesh+((long)ia*n+ib)*n;
        float lo=wb*(1-fz), hi=wb*fz;
        if((iz&1)==0){ atomicAdd((float2*)(row+iz), make_float2(lo,hi)); }
        else { atomicAdd(row+iz,lo); int id=iz+1; if(id>=n)id-=n; atomicAdd(row+id,hi);} }}
  }
}
// G: 3 channels, planar (24 scalar) vs interleaved float4 (8 x v4)
__global__ void kG_planar(const float* __restrict__ pos, const float* __restrict__ val, float* mesh, long np, int n){
  long plane=(long)n*n*n;
  for(long p=blockIdx.x*(long)blockDim.x+threadIdx.x;p<np;p+=(long)gridDim.x*blockDim.x){
    int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
    float v0=val[3*p],v1=val[3*p+1],v2=val[3*p+2];
    for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
      for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
        float* row=mesh+((long)ia*n+ib)*n;
        for(int d=0;d<2;d++){int id=iz+d; if(id>=n)id-=n; float w=wb*(d?fz:1-fz);
          atomicAdd(row+id,v0*w); atomicAdd(row+plane+id,v1*w); atomicAdd(row+2*plane+id,v2*w);}}}
  }
}
__global__ void kG_v4(const float* __restrict__ pos, const float* __restrict__ val, float4* mesh, long np, int n){
  for(long p=blockIdx.x*(long)blockDim.x+threadIdx.x;p<np;p+=(long)gridDim.x*blockDim.x){
    int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
    float v0=val[3*p],v1=val[3*p+1],v2=val[3*p+2];
    for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
      for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
        float4* row=mesh+((long)ia*n+ib)*n;
        for(int d=0;d<2;d++){int id=iz+d; if(id>=n)id-=n; float w=wb*(d?fz:1-fz);
          atomicAdd(row+id, make_float4(v0*w,v1*w,v2*w,0.f));}}}
  }
}
// H: 3 channels interleaved, z-pair merged by shuffle (4 x v4 + leftovers)
__global__ void kH_v4_merge(const float* __restrict__ pos, const float* __restrict__ val, float4* mesh, long np, int n){
  for(long p0=blockIdx.x*(long)blockDim.x+threadIdx.x;p0<((np+31)/32)*32;p0+=(long)gridDim.x*blockDim.x){
    bool act=p0<np; long p=act?p0:np-1;
    int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
    float v0=val[3*p],v1=val[3*p+1],v2=val[3*p+2];
    int lane=threadIdx.x&31;
    int pix=__shfl_up_sync(0xffffffffu,ix,1), piy=__shfl_up_sync(0xffffffffu,iy,1), piz=__shfl_up_sync(0xffffffffu,iz,1);
    bool pact=__shfl_up_sync(0xffffffffu,(int)act,1);
    bool chain = lane>0 && pact && act && pix==ix && piy==iy && piz+1==iz;
    bool nextchain = __shfl_down_sync(0xffffffffu,(int)chain,1) && lane<31;
    float pv0=__shfl_up_sync(0xffffffffu,v0,1),pv1=__shfl_up_sync(0xffffffffu,v1,1),pv2=__shfl_up_sync(0xffffffffu,v2,1);
    for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
      for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
        float lo=wb*(1-fz), hi=wb*fz;
        float phi=__shfl_up_sync(0xffffffffu,hi,1);
        float4 L=make_float4(v0*lo,v1*lo,v2*lo,0.f);
        if(chain){L.x+=pv0*phi;L.y+=pv1*phi;L.z+=pv2*phi;}
        float4* row=mesh+((long)ia*n+ib)*n;
        if(act) atomicAdd(row+iz, L);
        if(act && !nextchain){int id=iz+1; if(id>=n)id-=n; atomicAdd(row+id, make_float4(v0*hi,v1*hi,v2*hi,0.f));}
      }}
  }
}
template<class F> float timeit(F f, float* flush, size_t flushn){
  float best=1e9; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int r=0;r<7;r++){ CK(cudaMemsetAsync(flush,0,flushn)); cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r>=2&&ms<best)best=ms; }
  return best;
}
int main(int argc,char**argv){
  int n=argc>1?atoi(argv[1]):256; float sigma=argc>2?atof(argv[2]):1.5f; float jit=argc>3?atof(argv[3]):0.3f;
  long np=(long)n*n*n; std::vector<float> h(3*np); srand(1);
  double k=2*M_PI/n;
  for(long p=0;p<np;p++){ int i=p/((long)n*n), j=(p/n)%n, l=p%n;
    auto r=[&]{return jit*((rand()/(float)RAND_MAX)*2-1)*1.7f;};
    h[3*p]=i+sigma*(sin(k*3*j)+cos(k*5*l))+r(); h[3*p+1]=j+sigma*(sin(k*4*l)+cos(k*2*i))+r(); h[3*p+2]=l+sigma*(sin(k*3*i)+cos(k*6*j))+r(); }
  float *pos,*val,*mesh,*flush; size_t flushn=256u<<20;
  CK(cudaMalloc(&pos,12*np)); CK(cudaMalloc(&val,12*np)); CK(cudaMalloc(&mesh,16*np)); CK(cudaMalloc(&flush,flushn));
  CK(cudaMemcpy(pos,h.data(),12*np,cudaMemcpyHostToDevice)); CK(cudaMemcpy(val,h.data(),12*np,cudaMemcpyHostToDevice));
  int grids[3]={148*8,148*16,(int)((np+255)/256)};
  for(int gi=0;gi<3;gi++){ int g=grids[gi];
    printf("--- grid %d x 256, n=%d sigma=%.2f jitter=%.2f\n",g,n,sigma,jit);
    for(int nc=1;nc<=8;nc*=2){ float t=timeit([&]{cudaMemsetAsync(mesh,0,4*np); kA<<<g,256>>>(pos,mesh,np,n,nc);},flush,flushn); printf("A scalar, %d corner(s)      %.3f ms\n",nc,t);}
    { float t=timeit([&]{cudaMemsetAsync(mesh,0,4*np); kB<<<g,256>>>(pos,mesh,np,n);},flush,flushn); printf("B z-merge by shuffle       %.3f ms\n",t);}
    { float t=timeit([&]{cudaMemsetAsync(mesh,0,4*np); kC<<<g,256>>>(pos,mesh,np,n);},flush,flushn); printf("C red.v2 when aligned      %.3f ms\n",t);}
    { float t=timeit([&]{cudaMemsetAsync(mesh,0,12*np); kG_planar<<<g,256>>>(pos,val,mesh,np,n);},flush,flushn); printf("G 3ch planar (24 scalar)   %.3f ms\n",t);}
    { float t=timeit([&]{cudaMemsetAsync(mesh,0,16*np); kG_v4<<<g,256>>>(pos,val,(float4*)mesh,np,n);},flush,flushn); printf("G 3ch interleaved (8 v4)   %.3f ms\n",t);}
    { float t=timeit([&]{cudaMemsetAsync(mesh,0,16*np); kH_v4_merge<<<g,256>>>(pos,val,(float4*)mesh,np,n);},flush,flushn); printf("H 3ch v4 + z-merge         %.3f ms\n",t);}
  }
  // checksum of B vs A
  std::vector<float> ma(np), mb(np);
  cudaMemset(mesh,0,4*np); kA<<<148*8,256>>>(pos,mesh,np,n,8); CK(cudaMemcpy(ma.data(),mesh,4*np,cudaMemcpyDeviceToHost));
  cudaMemset(mesh,0,4*np); kB<<<148*8,256>>>(pos,mesh,np,n); CK(cudaMemcpy(mb.data(),mesh,4*np,cudaMemcpyDeviceToHost));
  double d=0,s=0; for(long i=0;i<np;i++){d+=fabs(ma[i]-mb[i]); s+=fabs(ma[i]);} printf("B vs A: sum|diff|/sum = %.3e\n",d/s);
  return 0;
}
