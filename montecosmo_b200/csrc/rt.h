// rt.h -- thin runtime layer: CUDA (sm_100a) in the product build; plain C++ (OpenMP) loops when compiled with
// -DMCPM_HOSTEMU.  The host build exists ONLY as checker / CPU baseline: oracle/cpu_port.py compiles it into
// oracle/_build/ so that tests/ can exercise the engine's orchestration (buffer wiring, adjoint bookkeeping) without
// a GPU and bench.py can time the same algorithm on the host cores.  The montecosmo_b200 package never loads it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/mcpm.h"

namespace mcpm {

struct cfloat {
  float re, im;
};

void set_error(const std::string& msg);
void count_launch();  // api.cu: number of engine kernels launched by this process (bench.py's gpu_launches)

#ifdef MCPM_HOSTEMU
#define MCPM_HD inline
#define MCPM_LAMBDA
typedef void* stream_t;

inline stream_t as_stream(void* s) { return s; }

template <class F>
inline void launch_1d(stream_t, int64_t n, F f) {
  if (n > 0) count_launch();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) f(i);
}
template <int MINB, class F>
inline void launch_1d_occ(stream_t s, int64_t n, F f) {
  launch_1d(s, n, f);
}
MCPM_HD void atomic_add(float* p, float v) {
#pragma omp atomic
  *p += v;
}
MCPM_HD void atomic_add(double* p, double v) {
#pragma omp atomic
  *p += v;
}
inline int rt_memset(void* p, int v, size_t bytes, stream_t) {
  memset(p, v, bytes);
  return 0;
}
inline int rt_copy(void* d, const void* s, size_t bytes, stream_t) {
  memcpy(d, s, bytes);
  return 0;
}
inline int rt_malloc(void** p, size_t bytes) {
  *p = malloc(bytes ? bytes : 1);
  return *p ? 0 : 1;
}
inline void rt_free(void* p) { free(p); }
inline int rt_check(const char*) { return 0; }
inline int rt_get_device() { return 0; }
inline int rt_bind_device(int) { return 0; }

#else  // ---------------------------------------------------------------- CUDA
#include <cuda_runtime.h>
#define MCPM_HD __host__ __device__ __forceinline__
#define MCPM_LAMBDA __device__
typedef cudaStream_t stream_t;

inline stream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

template <class F>
__global__ void __launch_bounds__(256) k_launch_1d(int64_t n, F f) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

// Grid-stride launch sized in whole waves of the 148 SMs (16 CTAs of 256 threads per SM: two full-occupancy waves).
template <class F>
inline void launch_1d(stream_t s, int64_t n, F f) {
  if (n <= 0) return;
  int64_t blocks = (n + 255) / 256;
  const int64_t wave = (int64_t)kSMs * 16;
  if (blocks > wave) blocks = wave;
  count_launch();
  k_launch_1d<<<(unsigned)blocks, 256, 0, s>>>(n, f);
}

// Same, with a register cap chosen for MINB resident CTAs of 256 threads per SM (latency-bound gathers want occupancy).
template <int MINB, class F>
__global__ void __launch_bounds__(256, MINB) k_launch_1d_occ(int64_t n, F f) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <int MINB, class F>
inline void launch_1d_occ(stream_t s, int64_t n, F f) {
  if (n <= 0) return;
  int64_t blocks = (n + 255) / 256;
  const int64_t wave = (int64_t)kSMs * MINB * 2;
  if (blocks > wave) blocks = wave;
  count_launch();
  k_launch_1d_occ<MINB><<<(unsigned)blocks, 256, 0, s>>>(n, f);
}

__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }

inline int rt_memset(void* p, int v, size_t bytes, stream_t s) { return (int)cudaMemsetAsync(p, v, bytes, s); }
inline int rt_copy(void* d, const void* src, size_t bytes, stream_t s) {
  return (int)cudaMemcpyAsync(d, src, bytes, cudaMemcpyDeviceToDevice, s);
}
inline int rt_malloc(void** p, size_t bytes) { return (int)cudaMalloc(p, bytes ? bytes : 1); }
inline void rt_free(void* p) { cudaFree(p); }
inline int rt_get_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
// Make the engine's device (and its primary context) current on the calling host thread.  Callers such as torch's
// autograd workers or XLA's executor threads may arrive with no context bound, which cuFFT does not tolerate.
inline int rt_bind_device(int d) {
  if (cudaSetDevice(d) != cudaSuccess) {
    set_error("cudaSetDevice failed");
    cudaGetLastError();
    return MCPM_ECUDA;
  }
  return 0;
}
inline int rt_check(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cudaGetLastError();
    return MCPM_ECUDA;
  }
  return 0;
}
#endif

// FFT layer (fft.cu): batched 3-D R2C / C2R on contiguous planes.
struct FftPlans;
FftPlans* fft_create(int nx, int ny, int nz, size_t* work_bytes);
void fft_destroy(FftPlans*);
int fft_r2c(FftPlans*, stream_t, const float* in, cfloat* out, int batch);
int fft_c2r(FftPlans*, stream_t, cfloat* in, float* out, int batch);

}  // namespace mcpm
