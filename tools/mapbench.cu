// mapbench.cu -- does the particle->thread mapping matter?  CTA of 256 threads covers a BXxBYxBZ brick of the particle
// lattice (lanes fastest along z, then y, then x), direct global atomics / loads, no shared memory.  Development tool.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)
struct Map { int bx,by,bz; int wx,wy,wz; };  // CTA brick dims and warp sub-brick dims (wx*wy*wz = 32)
__device__ __forceinline__ long pidx(const Map& m, int n, bool& valid){
  // bricks tile the lattice; within the CTA, warps tile the brick with sub-bricks wx x wy x wz
  int nbz=n/m.bz, nby=n/m.by;
  long b=blockIdx.x; int cz=b%nbz; b/=nbz; int cy=b%nby; int cx=b/nby;
  int warp=threadIdx.x>>5, lane=threadIdx.x&31;
  int swz=m.bz/m.wz, swy=m.by/m.wy;
  int sz=warp%swz, sy=(warp/swz)%swy, sx=warp/(swz*swy);
  int lz=lane%m.wz, ly=(lane/m.wz)%m.wy, lx=lane/(m.wz*m.wy);
  int i=cx*m.bx+sx*m.wx+lx, j=cy*m.by+sy*m.wy+ly, k=cz*m.bz+sz*m.wz+lz;
  valid = i<n && j<n && k<n;
  return ((long)i*n+j)*n+k;
}
__device__ __forceinline__ void cic(const float* __restrict__ pos, long p, int n, int& ix,int& iy,int& iz,float& fx,float& fy,float& fz){
  float x=pos[3*p],y=pos[3*p+1],z=pos[3*p+2];
  float bx=floorf(x),by=floorf(y),bz=floorf(z); fx=x-bx;fy=y-by;fz=z-bz;
  ix=((int)bx%n+n)%n; iy=((int)by%n+n)%n; iz=((int)bz%n+n)%n;
}
template<int NCH,bool V4> __global__ void __launch_bounds__(256) scat(Map m,const float* __restrict__ pos,const float* __restrict__ val,float* mesh,int n){
  bool valid; long p=pidx(m,n,valid); if(!valid) return;
  int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
  float v0=1.f,v1=0.f,v2=0.f; if(NCH==3){v0=val[3*p];v1=val[3*p+1];v2=val[3*p+2];}
  long plane=(long)n*n*n;
  for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
    for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
      long row=((long)ia*n+ib)*n;
      for(int d=0;d<2;d++){int id=iz+d; if(id>=n)id-=n; float w=wb*(d?fz:1-fz);
        if(NCH==1) atomicAdd(mesh+row+id,v0*w);
        else if(V4) atomicAdd(((float4*)mesh)+row+id, make_float4(v0*w,v1*w,v2*w,0.f));
        else { atomicAdd(mesh+row+id,v0*w); atomicAdd(mesh+plane+row+id,v1*w); atomicAdd(mesh+2*plane+row+id,v2*w);} }}}
}
template<bool V4> __global__ void __launch_bounds__(256) gath(Map m,const float* __restrict__ pos,const float* __restrict__ mesh,float* out,int n){
  bool valid; long p=pidx(m,n,valid); if(!valid) return;
  int ix,iy,iz;float fx,fy,fz; cic(pos,p,n,ix,iy,iz,fx,fy,fz);
  long plane=(long)n*n*n; float a0=0,a1=0,a2=0;
  for(int a=0;a<2;a++){int ia=ix+a; if(ia>=n)ia-=n; float wa=a?fx:1-fx;
    for(int b=0;b<2;b++){int ib=iy+b; if(ib>=n)ib-=n; float wb=wa*(b?fy:1-fy);
      long row=((long)ia*n+ib)*n;
      for(int d=0;d<2;d++){int id=iz+d; if(id>=n)id-=n; float w=wb*(d?fz:1-fz);
        if(V4){ float4 q=__ldg(((const float4*)mesh)+row+id); a0+=q.x*w;a1+=q.y*w;a2+=q.z*w; }
        else { a0+=__ldg(mesh+row+id)*w; a1+=__ldg(mesh+plane+row+id)*w; a2+=__ldg(mesh+2*plane+row+id)*w; } }}}
  out[3*p]=a0; out[3*p+1]=a1; out[3*p+2]=a2;
}
template<class F> float timeit(F f, float* flush, size_t flushn){
  float best=1e9; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int r=0;r<6;r++){ CK(cudaMemsetAsync(flush,0,flushn)); cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r>=1&&ms<best)best=ms; }
  return best;
}
int main(int argc,char**argv){
  int n=256; float sigma=argc>1?atof(argv[1]):1.5f; float jit=argc>2?atof(argv[2]):0.3f;
  long np=(long)n*n*n; std::vector<float> h(3*np); srand(1); double k=2*M_PI/n;
  for(long p=0;p<np;p++){ int i=p/((long)n*n), j=(p/n)%n, l=p%n; auto r=[&]{return jit*((rand()/(float)RAND_MAX)*2-1)*1.7f;};
    h[3*p]=i+sigma*(sin(k*3*j)+cos(k*5*l))+r(); h[3*p+1]=j+sigma*(sin(k*4*l)+cos(k*2*i))+r(); h[3*p+2]=l+sigma*(sin(k*3*i)+cos(k*6*j))+r(); }
  float *pos,*val,*mesh,*out,*flush; size_t flushn=256u<<20;
  CK(cudaMalloc(&pos,12*np)); CK(cudaMalloc(&val,12*np)); CK(cudaMalloc(&mesh,16*np)); CK(cudaMalloc(&out,12*np)); CK(cudaMalloc(&flush,flushn));
  CK(cudaMemcpy(pos,h.data(),12*np,cudaMemcpyHostToDevice)); CK(cudaMemcpy(val,h.data(),12*np,cudaMemcpyHostToDevice));
  Map maps[]={ {1,1,256,1,1,32}, {1,8,32,1,1,32}, {2,4,32,1,1,32}, {4,4,16,1,2,16}, {4,8,8,1,4,8}, {8,8,4,2,4,4}, {2,2,64,1,1,32}, {4,4,16,2,2,8}, {8,4,8,2,2,8}, {4,2,32,1,1,32} };
  printf("sigma=%.2f jitter=%.2f   (CTA brick / warp sub-brick)   paint1  paint3-planar  paint3-v4  gather3-planar  gather3-v4  [ms]\n",sigma,jit);
  for(auto m:maps){
    int g=(n/m.bx)*(n/m.by)*(n/m.bz);
    float t1=timeit([&]{cudaMemsetAsync(mesh,0,4*np); scat<1,false><<<g,256>>>(m,pos,val,mesh,n);},flush,flushn);
    float t3=timeit([&]{cudaMemsetAsync(mesh,0,12*np); scat<3,false><<<g,256>>>(m,pos,val,mesh,n);},flush,flushn);
    float t4=timeit([&]{cudaMemsetAsync(mesh,0,16*np); scat<3,true><<<g,256>>>(m,pos,val,mesh,n);},flush,flushn);
    float g3=timeit([&]{gath<false><<<g,256>>>(m,pos,mesh,out,n);},flush,flushn);
    float g4=timeit([&]{gath<true><<<g,256>>>(m,pos,mesh,out,n);},flush,flushn);
    printf("%dx%dx%-3d / %dx%dx%-2d   %.3f   %.3f   %.3f   %.3f   %.3f\n",m.bx,m.by,m.bz,m.wx,m.wy,m.wz,t1,t3,t4,g3,g4);
  }
  cudaError_t e=cudaGetLastError(); if(e) printf("err %s\n",cudaGetErrorString(e));
  return 0;
}
