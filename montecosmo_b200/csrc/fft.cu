// fft.cu -- batched 3-D R2C / C2R on contiguous planes.  Product build: cuFFT plans created once per engine, one
// shared work area, attached to the caller's stream per call (no allocation, no sync on the call path).
// Reference call sites: jnp.fft.rfftn / irfftn at nbody.py:589, 603, 620, 627, 630, 525 (unnormalised forward,
// 1/N inverse).
#include <map>
#include <mutex>

#include "engine.h"

#ifndef MCPM_HOSTEMU
#include <cufft.h>
#endif

namespace mcpm {

#ifdef MCPM_HOSTEMU
// Checker-only hooks: the host build has no FFT of its own; oracle/cpu_port.py registers scipy.fft (pocketfft) here.
typedef void (*fft_hook_t)(const void* in, void* out, int nx, int ny, int nz, int batch);
static fft_hook_t g_r2c = nullptr, g_c2r = nullptr;
extern "C" __attribute__((visibility("default"))) void mcpm_hostemu_set_fft(fft_hook_t r2c, fft_hook_t c2r) {
  g_r2c = r2c;
  g_c2r = c2r;
}
struct FftPlans {
  int nx, ny, nz;
};
FftPlans* fft_create(int nx, int ny, int nz, size_t* work_bytes) {
  if (work_bytes) *work_bytes = 0;
  return new FftPlans{nx, ny, nz};
}
void fft_destroy(FftPlans* p) { delete p; }
int fft_r2c(FftPlans* p, stream_t, const float* in, cfloat* out, int batch) {
  if (!g_r2c) {
    set_error("hostemu: no FFT hook registered");
    return MCPM_ECUFFT;
  }
  g_r2c(in, out, p->nx, p->ny, p->nz, batch);
  return 0;
}
int fft_c2r(FftPlans* p, stream_t, cfloat* in, float* out, int batch) {
  if (!g_c2r) {
    set_error("hostemu: no FFT hook registered");
    return MCPM_ECUFFT;
  }
  g_c2r(in, out, p->nx, p->ny, p->nz, batch);
  // cuFFT's C2R may overwrite its input: poison it so that orchestration bugs relying on it show up on CPU.
  size_t nc = (size_t)p->nx * p->ny * (p->nz / 2 + 1) * batch;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nc; ++i) in[i] = cfloat{NAN, NAN};
  return 0;
}

#else
struct FftPlans {
  int nx, ny, nz;
  void* work = nullptr;
  size_t work_bytes = 0;
  std::map<int, cufftHandle> r2c, c2r;  // keyed by batch
  std::mutex mu;
};

static int make_plan(FftPlans* p, cufftType type, int batch, cufftHandle* out, size_t* ws) {
  cufftHandle h;
  if (cufftCreate(&h) != CUFFT_SUCCESS) return MCPM_ECUFFT;
  cufftSetAutoAllocation(h, 0);
  int n[3] = {p->nx, p->ny, p->nz};
  int nzc = p->nz / 2 + 1;
  int rembed[3] = {p->nx, p->ny, p->nz};
  int cembed[3] = {p->nx, p->ny, nzc};
  int rdist = p->nx * p->ny * p->nz, cdist = p->nx * p->ny * nzc;
  cufftResult r;
  if (type == CUFFT_R2C)
    r = cufftMakePlanMany(h, 3, n, rembed, 1, rdist, cembed, 1, cdist, CUFFT_R2C, batch, ws);
  else
    r = cufftMakePlanMany(h, 3, n, cembed, 1, cdist, rembed, 1, rdist, CUFFT_C2R, batch, ws);
  if (r != CUFFT_SUCCESS) {
    cufftDestroy(h);
    set_error("cufftMakePlanMany failed with code " + std::to_string((int)r));
    return MCPM_ECUFFT;
  }
  *out = h;
  return 0;
}

// Plans for batch 1, 2, 3 and 6 in both directions, sharing one work area sized for the largest.
FftPlans* fft_create(int nx, int ny, int nz, size_t* work_bytes) {
  FftPlans* p = new FftPlans;
  p->nx = nx;
  p->ny = ny;
  p->nz = nz;
  if ((int64_t)nx * ny * nz >= (int64_t)1 << 31) {
    set_error("mesh too large for 32-bit cuFFT plan distances");
    delete p;
    return nullptr;
  }
  size_t maxws = 0;
  const int batches[4] = {1, 2, 3, 6};
  for (int b : batches) {
    size_t ws = 0;
    cufftHandle h;
    if (make_plan(p, CUFFT_R2C, b, &h, &ws)) {
      fft_destroy(p);
      return nullptr;
    }
    p->r2c[b] = h;
    maxws = ws > maxws ? ws : maxws;
    if (make_plan(p, CUFFT_C2R, b, &h, &ws)) {
      fft_destroy(p);
      return nullptr;
    }
    p->c2r[b] = h;
    maxws = ws > maxws ? ws : maxws;
  }
  if (cudaMalloc(&p->work, maxws ? maxws : 1) != cudaSuccess) {
    set_error("cuFFT work area allocation failed");
    cudaGetLastError();
    fft_destroy(p);
    return nullptr;
  }
  p->work_bytes = maxws;
  for (auto& kv : p->r2c) cufftSetWorkArea(kv.second, p->work);
  for (auto& kv : p->c2r) cufftSetWorkArea(kv.second, p->work);
  if (work_bytes) *work_bytes = maxws;
  return p;
}

void fft_destroy(FftPlans* p) {
  if (!p) return;
  for (auto& kv : p->r2c) cufftDestroy(kv.second);
  for (auto& kv : p->c2r) cufftDestroy(kv.second);
  if (p->work) cudaFree(p->work);
  delete p;
}

// A batch without its own plan is run as a sequence of the available ones (largest first).
template <class Exec>
static int run_batched(std::map<int, cufftHandle>& plans, int batch, Exec exec) {
  int done = 0;
  while (done < batch) {
    int take = 0;
    for (auto it = plans.rbegin(); it != plans.rend(); ++it)
      if (it->first <= batch - done) {
        take = it->first;
        break;
      }
    cufftResult r = exec(plans[take], done);
    if (r != CUFFT_SUCCESS) {
      set_error("cufftExec failed with code " + std::to_string((int)r));
      return MCPM_ECUFFT;
    }
    done += take;
  }
  return 0;
}

int fft_r2c(FftPlans* p, stream_t st, const float* in, cfloat* out, int batch) {
  std::lock_guard<std::mutex> lk(p->mu);
  const size_t rd = (size_t)p->nx * p->ny * p->nz, cd = (size_t)p->nx * p->ny * (p->nz / 2 + 1);
  return run_batched(p->r2c, batch, [&](cufftHandle h, int off) {
    cufftSetStream(h, st);
    return cufftExecR2C(h, const_cast<float*>(in) + off * rd, reinterpret_cast<cufftComplex*>(out + off * cd));
  });
}

int fft_c2r(FftPlans* p, stream_t st, cfloat* in, float* out, int batch) {
  std::lock_guard<std::mutex> lk(p->mu);
  const size_t rd = (size_t)p->nx * p->ny * p->nz, cd = (size_t)p->nx * p->ny * (p->nz / 2 + 1);
  return run_batched(p->c2r, batch, [&](cufftHandle h, int off) {
    cufftSetStream(h, st);
    return cufftExecC2R(h, reinterpret_cast<cufftComplex*>(in + off * cd), out + off * rd);
  });
}
#endif

}  // namespace mcpm

// ---------------------------------------------------------------------------------------------------------------
// Slab-decomposed FFT building blocks.  A distributed R2C is: 2-D R2C over (y,z) of the local x-planes,
// an all-to-all that trades x-planes for ky rows (done by the caller over NCCL), then a 1-D C2C along x of the local
// ky block [nx, kyl, nzc] (x is the slowest axis, so the transform is strided by kyl*nzc with unit distance between
// batches).  The inverse runs the same steps backwards.  All transforms are unnormalised.
// ---------------------------------------------------------------------------------------------------------------
namespace mcpm {

#ifdef MCPM_HOSTEMU
typedef void (*slab_hook_t)(int kind, const void* in, void* out, int nx, int ny, int nz, int nb, int xl, int kyl);
static slab_hook_t g_slab = nullptr;
extern "C" __attribute__((visibility("default"))) void mcpm_hostemu_set_slabfft(slab_hook_t h) { g_slab = h; }
struct SlabFft {
  int nx, ny, nz, parts, xl, kyl;
};
SlabFft* slabfft_create(int nx, int ny, int nz, int parts) {
  if (parts < 1 || nx % parts || ny % parts || (nz & 1)) {
    set_error("slabfft: nx and ny must be divisible by the number of slabs and nz even");
    return nullptr;
  }
  return new SlabFft{nx, ny, nz, parts, nx / parts, ny / parts};
}
void slabfft_destroy(SlabFft* p) { delete p; }
static int slab_call(SlabFft* p, int kind, const void* in, void* out, int nb) {
  if (!g_slab) {
    set_error("hostemu: no slab FFT hook registered");
    return MCPM_ECUFFT;
  }
  g_slab(kind, in, out, p->nx, p->ny, p->nz, nb, p->xl, p->kyl);
  return 0;
}
int slabfft_r2c_yz(SlabFft* p, stream_t, const float* in, cfloat* out, int nb) { return slab_call(p, 0, in, out, nb); }
int slabfft_c2r_yz(SlabFft* p, stream_t, cfloat* in, float* out, int nb) { return slab_call(p, 1, in, out, nb); }
int slabfft_c2c_x(SlabFft* p, stream_t, cfloat* data, int nb, int inverse) {
  return slab_call(p, inverse ? 3 : 2, data, data, nb);
}
#else
struct SlabFft {
  int nx, ny, nz, parts, xl, kyl;
  void* work = nullptr;
  std::map<int, cufftHandle> r2c, c2r;  // 2-D over (y,z), keyed by number of meshes nb (batch = nb * xl)
  cufftHandle c2c = 0;                  // 1-D along x, batch kyl*nzc, stride kyl*nzc
  size_t work_bytes = 0;
  std::mutex mu;
};

static int slab_plan2d(SlabFft* p, cufftType type, int nb, cufftHandle* out, size_t* ws) {
  cufftHandle h;
  if (cufftCreate(&h) != CUFFT_SUCCESS) return MCPM_ECUFFT;
  cufftSetAutoAllocation(h, 0);
  int n[2] = {p->ny, p->nz};
  int nzc = p->nz / 2 + 1;
  int re[2] = {p->ny, p->nz}, ce[2] = {p->ny, nzc};
  cufftResult r = type == CUFFT_R2C
                      ? cufftMakePlanMany(h, 2, n, re, 1, p->ny * p->nz, ce, 1, p->ny * nzc, CUFFT_R2C, nb * p->xl, ws)
                      : cufftMakePlanMany(h, 2, n, ce, 1, p->ny * nzc, re, 1, p->ny * p->nz, CUFFT_C2R, nb * p->xl, ws);
  if (r != CUFFT_SUCCESS) {
    cufftDestroy(h);
    set_error("slabfft: cufftMakePlanMany (2-D) failed with code " + std::to_string((int)r));
    return MCPM_ECUFFT;
  }
  *out = h;
  return 0;
}

SlabFft* slabfft_create(int nx, int ny, int nz, int parts) {
  if (parts < 1 || nx % parts || ny % parts || (nz & 1)) {
    set_error("slabfft: nx and ny must be divisible by the number of slabs and nz even");
    return nullptr;
  }
  SlabFft* p = new SlabFft;
  p->nx = nx;
  p->ny = ny;
  p->nz = nz;
  p->parts = parts;
  p->xl = nx / parts;
  p->kyl = ny / parts;
  size_t maxws = 0, ws = 0;
  // Batched plans for 1, 2, 3 and 6 meshes at once; a batch without its own plan runs as a sequence of the available ones
  // (run_batched).  The work area is shared and sized by the largest plan: for big local volumes (the 2048^3 paint mesh
  // of BASELINE C5: 256 planes of 2048^2 per rank) only the plans whose work area stays under 4 GiB are kept.
  const int nbs[4] = {1, 2, 3, 6};
  const size_t cap = (size_t)4 << 30;
  for (int nb : nbs) {
    cufftHandle h, h2;
    size_t ws2 = 0;
    if (slab_plan2d(p, CUFFT_R2C, nb, &h, &ws)) {
      slabfft_destroy(p);
      return nullptr;
    }
    if (slab_plan2d(p, CUFFT_C2R, nb, &h2, &ws2)) {
      cufftDestroy(h);
      slabfft_destroy(p);
      return nullptr;
    }
    if (nb > 1 && (ws > cap || ws2 > cap)) {
      cufftDestroy(h);
      cufftDestroy(h2);
      continue;
    }
    p->r2c[nb] = h;
    p->c2r[nb] = h2;
    maxws = ws > maxws ? ws : maxws;
    maxws = ws2 > maxws ? ws2 : maxws;
  }
  {
    cufftHandle h;
    cufftCreate(&h);
    cufftSetAutoAllocation(h, 0);
    int n[1] = {nx};
    int stride = p->kyl * (nz / 2 + 1);
    int embed[1] = {nx};
    cufftResult r = cufftMakePlanMany(h, 1, n, embed, stride, 1, embed, stride, 1, CUFFT_C2C, stride, &ws);
    if (r != CUFFT_SUCCESS) {
      cufftDestroy(h);
      set_error("slabfft: cufftMakePlanMany (1-D along x) failed with code " + std::to_string((int)r));
      slabfft_destroy(p);
      return nullptr;
    }
    p->c2c = h;
    maxws = ws > maxws ? ws : maxws;
  }
  if (cudaMalloc(&p->work, maxws ? maxws : 1) != cudaSuccess) {
    set_error("slabfft: work area allocation failed");
    cudaGetLastError();
    slabfft_destroy(p);
    return nullptr;
  }
  p->work_bytes = maxws;
  for (auto& kv : p->r2c) cufftSetWorkArea(kv.second, p->work);
  for (auto& kv : p->c2r) cufftSetWorkArea(kv.second, p->work);
  cufftSetWorkArea(p->c2c, p->work);
  return p;
}

void slabfft_destroy(SlabFft* p) {
  if (!p) return;
  for (auto& kv : p->r2c) cufftDestroy(kv.second);
  for (auto& kv : p->c2r) cufftDestroy(kv.second);
  if (p->c2c) cufftDestroy(p->c2c);
  if (p->work) cudaFree(p->work);
  delete p;
}

int slabfft_r2c_yz(SlabFft* p, stream_t st, const float* in, cfloat* out, int nb) {
  if (tune().yzfft && yzfft_supported(p->ny, p->nz)) {  // both passes in one kernel (yzfft.cu)
    const int r = yzfft_r2c(st, in, out, p->ny, p->nz, nb * p->xl);
    if (r != MCPM_EUNSUP) return r;
  }
  std::lock_guard<std::mutex> lk(p->mu);
  const size_t rd = (size_t)p->xl * p->ny * p->nz, cd = (size_t)p->xl * p->ny * (p->nz / 2 + 1);
  return run_batched(p->r2c, nb, [&](cufftHandle h, int off) {
    cufftSetStream(h, st);
    return cufftExecR2C(h, const_cast<float*>(in) + off * rd, reinterpret_cast<cufftComplex*>(out + off * cd));
  });
}

int slabfft_c2r_yz(SlabFft* p, stream_t st, cfloat* in, float* out, int nb) {
  if (tune().yzfft && yzfft_supported(p->ny, p->nz)) {
    const int r = yzfft_c2r(st, in, out, p->ny, p->nz, nb * p->xl);
    if (r != MCPM_EUNSUP) return r;
  }
  std::lock_guard<std::mutex> lk(p->mu);
  const size_t rd = (size_t)p->xl * p->ny * p->nz, cd = (size_t)p->xl * p->ny * (p->nz / 2 + 1);
  return run_batched(p->c2r, nb, [&](cufftHandle h, int off) {
    cufftSetStream(h, st);
    return cufftExecC2R(h, reinterpret_cast<cufftComplex*>(in + off * cd), out + off * rd);
  });
}

int slabfft_c2c_x(SlabFft* p, stream_t st, cfloat* data, int nb, int inverse) {
  std::lock_guard<std::mutex> lk(p->mu);
  const size_t cd = (size_t)p->nx * p->kyl * (p->nz / 2 + 1);
  cufftSetStream(p->c2c, st);
  for (int b = 0; b < nb; ++b) {
    cufftComplex* d = reinterpret_cast<cufftComplex*>(data + b * cd);
    cufftResult r = cufftExecC2C(p->c2c, d, d, inverse ? CUFFT_INVERSE : CUFFT_FORWARD);
    if (r != CUFFT_SUCCESS) {
      set_error("slabfft: cufftExecC2C failed with code " + std::to_string((int)r));
      return MCPM_ECUFFT;
    }
  }
  return 0;
}
#endif

}  // namespace mcpm
