// api.cu -- the extern "C" boundary declared in include/mcpm.h.  No exceptions cross it; errors are codes plus a
// thread-local message.
#include <atomic>
#include <cstdlib>
#include <exception>

#include "engine.h"

namespace mcpm {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static Tune g_default_tune;
static thread_local const Tune* t_tune = nullptr;
Tune& default_tune() { return g_default_tune; }
const Tune& tune() { return t_tune ? *t_tune : g_default_tune; }
TuneScope::TuneScope(const Tune* t) : prev(t_tune) { t_tune = t; }
TuneScope::~TuneScope() { t_tune = prev; }

int tune_set(Tune& t, const char* key, int value) {
  const std::string k(key ? key : "");
  auto bad = [&](const char* msg) {
    set_error(std::string("tune: ") + msg);
    return (int)MCPM_EINVAL;
  };
  if (k == "gather_minb") {
    if (value < 4 || value > 6) return bad("gather_minb must be 4, 5 or 6");
    t.gather_minb = value;
  } else if (k == "side_zero") {
    t.side_zero = value != 0;
  } else if (k == "gather_blocked") {
    t.gather_blocked = value != 0;
  } else if (k == "brick") {
    t.brick = value != 0;
  } else if (k == "brick_stream") {
    if (value != 0 && value != 1 && value != 44 && value != 48) return bad("brick_stream must be 0, 44 or 48");
    t.brick_stream = value == 1 ? 44 : value;
  } else if (k == "brick_stream1") {
    t.brick_stream1 = value != 0;
  } else if (k == "gather_tma") {
    t.gather_tma = value != 0;
  } else if (k == "yzfft") {
    t.yzfft = value != 0;
  } else if (k == "gather_brick") {
    t.gather_brick = value != 0;
  } else if (k == "gather_seg") {
    if (value != 32 && value != 64 && value != 128) return bad("gather_seg must be 32, 64 or 128");
    t.gather_seg = value;
  } else {
    return bad(("unknown key " + k).c_str());
  }
  return MCPM_OK;
}

// Make the engine's device current for the duration of an entry point and restore the caller's device afterwards
// (callers such as XLA's executor threads or torch's autograd workers own their current-device state).
struct DeviceScope {
  int prev = -1, err = 0;
  explicit DeviceScope(int d) {
#ifndef MCPM_HOSTEMU
    if (cudaGetDevice(&prev) != cudaSuccess) {
      prev = -1;
      cudaGetLastError();
    }
    // cudaSetDevice ALWAYS: a host thread that has not touched the runtime yet (torch's autograd workers, XLA's executor
    // threads) reports device 0 as current without a context bound to it, and cufftExec then fails
    err = rt_bind_device(d);
    if (prev == d) prev = -1;  // nothing to restore
#else
    (void)d;
#endif
  }
  ~DeviceScope() {
#ifndef MCPM_HOSTEMU
    if (prev >= 0) cudaSetDevice(prev);
#endif
  }
};
}  // namespace mcpm

using namespace mcpm;

struct mcpm_engine {
  Engine* e;
};

struct mcpm_slabfft {
  SlabFft* f;
  int device;
};

#define API_BEGIN try {
#define API_END                                  \
  }                                              \
  catch (const std::exception& ex) {             \
    set_error(std::string("exception: ") + ex.what()); \
    return MCPM_EINVAL;                          \
  }                                              \
  catch (...) {                                  \
    set_error("unknown exception");              \
    return MCPM_EINVAL;                          \
  }

#define BIND(eng)                               \
  DeviceScope _dev((eng)->e->device);           \
  if (_dev.err) return _dev.err;                \
  TuneScope _tune(&(eng)->e->tune)

#define NEED(cond, msg)   \
  do {                    \
    if (!(cond)) {        \
      set_error(msg);     \
      return MCPM_EINVAL; \
    }                     \
  } while (0)

static inline const cfloat* C(const void* p) { return reinterpret_cast<const cfloat*>(p); }
static inline cfloat* C(void* p) { return reinterpret_cast<cfloat*>(p); }

// public mcpm_frame -> internal Frame; returns false on an inconsistent descriptor
static bool to_frame(const mcpm_frame* in, int64_t np, Frame& f, const Frame*& out) {
  out = nullptr;
  if (!in || !in->relative) return true;
  if (in->px <= 0 || in->py <= 0 || in->pz <= 0 || in->sx <= 0 || in->sy <= 0 || in->sz <= 0 ||
      (int64_t)in->px * in->py * in->pz != np) {
    set_error("frame: lattice px*py*pz must equal np and the spans must be positive");
    return false;
  }
  f = make_rel_frame(in->px, in->py, in->pz, in->sx, in->sy, in->sz, in->ox, in->oy, in->oz);
  out = &f;
  return true;
}
#define FRAME(np)                                  \
  Frame _f;                                        \
  const Frame* fr;                                 \
  if (!to_frame(frame, (np), _f, fr)) return MCPM_EINVAL

extern "C" {
#pragma GCC visibility push(default)

int mcpm_version(void) { return MCPM_VERSION; }
const char* mcpm_last_error(void) { return g_last_error.c_str(); }

long long mcpm_launch_count(int reset) {
  long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int mcpm_engine_create(int nx, int ny, int nz, mcpm_engine** out) {
  API_BEGIN
  NEED(out, "engine_create: null out");
  Engine* e = engine_create(nx, ny, nz);
  if (!e) return MCPM_ENOMEM;
  *out = new mcpm_engine{e};
  return MCPM_OK;
  API_END
}

int mcpm_engine_destroy(mcpm_engine* eng) {
  if (!eng) return MCPM_OK;
  engine_destroy(eng->e);
  delete eng;
  return MCPM_OK;
}

int mcpm_engine_set_lattice(mcpm_engine* eng, int px, int py, int pz) {
  API_BEGIN
  NEED(eng, "null engine");
  NEED(px >= 0 && py >= 0 && pz >= 0, "set_lattice: negative lattice shape");
  Lattice L;
  if (px > 0 && py > 0 && pz > 0) {
    L.px = px;
    L.py = py;
    L.pz = pz;
  }
  eng->e->lat = L;
  return MCPM_OK;
  API_END
}

int mcpm_engine_set_relative(mcpm_engine* eng, int on) {
  API_BEGIN
  NEED(eng, "null engine");
  if (on) NEED(eng->e->lat.px > 0, "set_relative: set the particle lattice first (mcpm_engine_set_lattice)");
  eng->e->rel = on ? 1 : 0;
  return MCPM_OK;
  API_END
}

int mcpm_tune(const char* key, int value) {
  API_BEGIN
  NEED(key, "tune: null key");
  return tune_set(default_tune(), key, value);
  API_END
}

int mcpm_engine_tune(mcpm_engine* eng, const char* key, int value) {
  API_BEGIN
  NEED(eng && key, "engine_tune: null argument");
  return tune_set(eng->e->tune, key, value);
  API_END
}

int mcpm_engine_set_fused_fft(mcpm_engine* eng, int on) {
  API_BEGIN
  NEED(eng, "null engine");
  if (on && !eng->e->fft2d) {
    set_error("set_fused_fft: not available for this mesh shape (nx must be 64, 128, 256, 512 or 1024) or CPU build");
    return MCPM_EUNSUP;
  }
  eng->e->fused_fft = on ? 1 : 0;
  return MCPM_OK;
  API_END
}

int mcpm_paint_brick(void* stream, int px, int py, int pz, const float* pos, const float* weights, float wscalar,
                     float shift, int64_t np, int nx, int ny, int nz, float* mesh) {
  API_BEGIN
  NEED(pos && mesh, "paint_brick: null pointer");
#ifndef MCPM_HOSTEMU
  Lattice L;
  L.px = px;
  L.py = py;
  L.pz = pz;
  int r = brick_paint_cic(as_stream(stream), L, pos, weights, wscalar, shift, np, nx, ny, nz, mesh);
  if (r < 0) return MCPM_ECUDA;
  if (r == 1) return MCPM_OK;
#endif
  set_error("paint_brick: unsupported lattice / mesh geometry, or CPU build");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_paint3_brick(void* stream, int px, int py, int pz, const float* pos, float* vbar, const float* xbar,
                      float drift, float scale, int64_t np, int nx, int ny, int nz, float* mesh3) {
  API_BEGIN
  NEED(pos && vbar && mesh3, "paint3_brick: null pointer");
#ifndef MCPM_HOSTEMU
  Lattice L;
  L.px = px;
  L.py = py;
  L.pz = pz;
  if (xbar)  // vbar += xbar * drift: its own pass here (the engine's reverse sweep folds it into the preceding gather)
    if (int e = axpy3(as_stream(stream), vbar, xbar, drift, 3 * np, vbar)) return e;
  int r = brick_paint3_cic(as_stream(stream), L, pos, vbar, scale, np, nx, ny, nz, mesh3);
  if (r < 0) return MCPM_ECUDA;
  if (r == 1) return MCPM_OK;
#endif
  set_error("paint3_brick: unsupported lattice / mesh geometry, or CPU build");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_paint_lattice(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar,
                       int64_t np, float* mesh) {
  API_BEGIN
  NEED(eng && pos && mesh, "paint_lattice: null pointer");
  BIND(eng);
#ifndef MCPM_HOSTEMU
  Engine* E = eng->e;
  int r = brick_paint_cic(as_stream(stream), E->lat, pos, weights, wscalar, 0.0f, np, E->nx, E->ny, E->nz, mesh);
  if (r < 0) return MCPM_ECUDA;
  if (r == 1) return MCPM_OK;
#endif
  set_error("paint_lattice: no matching lattice hint on this engine (mcpm_engine_set_lattice), or CPU build");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_paint3_lattice(mcpm_engine* eng, void* stream, const float* pos, float* vbar, const float* xbar, float drift,
                        float scale, int64_t np, float* mesh3) {
  API_BEGIN
  NEED(eng && pos && vbar && mesh3, "paint3_lattice: null pointer");
  BIND(eng);
#ifndef MCPM_HOSTEMU
  Engine* E = eng->e;
  if (xbar)
    if (int e = axpy3(as_stream(stream), vbar, xbar, drift, 3 * np, vbar)) return e;
  int r = brick_paint3_cic(as_stream(stream), E->lat, pos, vbar, scale, np, E->nx, E->ny, E->nz, mesh3);
  if (r < 0) return MCPM_ECUDA;
  if (r == 1) return MCPM_OK;
#endif
  set_error("paint3_lattice: no matching lattice hint on this engine (mcpm_engine_set_lattice), or CPU build");
  return MCPM_EUNSUP;
  API_END
}

size_t mcpm_engine_scratch_bytes(const mcpm_engine* eng) { return eng ? eng->e->scratch_bytes : 0; }

int mcpm_paint(void* stream, const float* pos, const float* weights, float wscalar, int64_t np, int nx, int ny,
               int nz, int order, const float scale[3], float shift, float* mesh, int accumulate) {
  API_BEGIN
  return paint(as_stream(stream), pos, weights, wscalar, np, nx, ny, nz, order, scale, shift, mesh, accumulate);
  API_END
}

int mcpm_read(void* stream, const float* pos, const float* mesh, int nmesh, int64_t np, int nx, int ny, int nz,
              int order, const float scale[3], float shift, float* out) {
  API_BEGIN
  return read(as_stream(stream), pos, mesh, nmesh, np, nx, ny, nz, order, scale, shift, out);
  API_END
}

int mcpm_read_grad(void* stream, const float* pos, const float* mesh, int nmesh, const float* cot, int64_t np,
                   int nx, int ny, int nz, int order, const float scale[3], float shift, float* grad,
                   int accumulate) {
  API_BEGIN
  NEED(nmesh >= 1 && nmesh <= 4, "read_grad: nmesh must be 1..4");
  const int64_t plane = (int64_t)nx * ny * nz;
  const float* ms[4] = {mesh, mesh + plane, mesh + 2 * plane, mesh + 3 * plane};
  return read_grad(as_stream(stream), pos, ms, nmesh, cot, cot ? nmesh : 0, 1.0f, nullptr, np, nx, ny, nz, order,
                   scale, shift, grad, accumulate);
  API_END
}

int mcpm_paint_vjp(void* stream, const float* pos, const float* weights, float wscalar, const float* mesh_bar,
                   int64_t np, int nx, int ny, int nz, int order, const float scale[3], float shift, float* posbar,
                   float* weightsbar, int accumulate) {
  API_BEGIN
  return paint_vjp(as_stream(stream), pos, weights, wscalar, mesh_bar, np, nx, ny, nz, order, scale, shift, posbar,
                   weightsbar, accumulate);
  API_END
}

// ---- Kaiser-Bessel window (kernel_type = 'kaiser_bessel', nbody.py:280-312, 321-322, 383-384, 415-416) ---------------
int mcpm_paint_kb(void* stream, const float* pos, const float* weights, float wscalar, int64_t np, int nx, int ny,
                  int nz, int order, float kcut, const float scale[3], float shift, float* mesh, int accumulate) {
  API_BEGIN
  NEED(kcut > 0.0f, "paint_kb: kcut must be positive (optim_kcut(oversamp))");
  return paint(as_stream(stream), pos, weights, wscalar, np, nx, ny, nz, order, scale, shift, mesh, accumulate, kcut);
  API_END
}

int mcpm_read_kb(void* stream, const float* pos, const float* mesh, int nmesh, int64_t np, int nx, int ny, int nz,
                 int order, float kcut, const float scale[3], float shift, float* out) {
  API_BEGIN
  NEED(kcut > 0.0f, "read_kb: kcut must be positive (optim_kcut(oversamp))");
  return read(as_stream(stream), pos, mesh, nmesh, np, nx, ny, nz, order, scale, shift, out, kcut);
  API_END
}

int mcpm_read_grad_kb(void* stream, const float* pos, const float* mesh, int nmesh, const float* cot, int64_t np,
                      int nx, int ny, int nz, int order, float kcut, const float scale[3], float shift, float* grad,
                      int accumulate) {
  API_BEGIN
  NEED(nmesh >= 1 && nmesh <= 4, "read_grad_kb: nmesh must be 1..4");
  NEED(kcut > 0.0f, "read_grad_kb: kcut must be positive (optim_kcut(oversamp))");
  const int64_t plane = (int64_t)nx * ny * nz;
  const float* ms[4] = {mesh, mesh + plane, mesh + 2 * plane, mesh + 3 * plane};
  return read_grad(as_stream(stream), pos, ms, nmesh, cot, cot ? nmesh : 0, 1.0f, nullptr, np, nx, ny, nz, order,
                   scale, shift, grad, accumulate, kcut);
  API_END
}

int mcpm_paint_vjp_kb(void* stream, const float* pos, const float* weights, float wscalar, const float* mesh_bar,
                      int64_t np, int nx, int ny, int nz, int order, float kcut, const float scale[3], float shift,
                      float* posbar, float* weightsbar, int accumulate) {
  API_BEGIN
  NEED(kcut > 0.0f, "paint_vjp_kb: kcut must be positive (optim_kcut(oversamp))");
  return paint_vjp(as_stream(stream), pos, weights, wscalar, mesh_bar, np, nx, ny, nz, order, scale, shift, posbar,
                   weightsbar, accumulate, kcut);
  API_END
}

int mcpm_deconv_kb(void* stream, const void* in, void* out, int nx, int ny, int nz, int order, float kcut) {
  API_BEGIN
  NEED(in && out, "deconv_kb: null pointer");
  NEED(order >= 1 && order <= 4 && kcut > 0.0f, "deconv_kb: order must be 1..4 and kcut positive");
  return deconv(as_stream(stream), C(in), C(out), nx, ny, nz, order, kcut);
  API_END
}

int mcpm_paint3(void* stream, const float* pos, const float* vals3, float vscale, int64_t np, int nx, int ny, int nz,
                int order, float* mesh3, int accumulate) {
  API_BEGIN
  return paint3(as_stream(stream), pos, vals3, vscale, nullptr, 0.0f, np, nx, ny, nz, order, mesh3, accumulate);
  API_END
}

int mcpm_rfftn(mcpm_engine* eng, void* stream, const float* in, void* out_c64, int batch) {
  API_BEGIN
  NEED(eng && in && out_c64 && batch >= 1, "rfftn: bad arguments");
  BIND(eng);
  return fft_r2c(eng->e->fft, as_stream(stream), in, C(out_c64), batch);
  API_END
}

int mcpm_irfftn(mcpm_engine* eng, void* stream, void* in_c64, float* out, int batch) {
  API_BEGIN
  NEED(eng && in_c64 && out && batch >= 1, "irfftn: bad arguments");
  BIND(eng);
  stream_t st = as_stream(stream);
  // jnp.fft.irfftn semantics for ANY input: the Hermitian projection it applies implicitly, made explicit for cuFFT
  if (int e = hermitian_project(st, C(in_c64), eng->e->nx, eng->e->ny, eng->e->nz, batch)) return e;
  if (int e = fft_c2r(eng->e->fft, st, C(in_c64), out, batch)) return e;
  return scale_real(st, out, eng->e->invN, out, eng->e->N * batch);  // standalone irfftn pays one scaling pass
  API_END
}

int mcpm_force_spectra(void* stream, const void* delta_k, void* out3, int nx, int ny, int nz, int lap_fd,
                       int grad_fd, float kcut, int deconv_order) {
  API_BEGIN
  return force_spectra(as_stream(stream), C(delta_k), C(out3), nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, 1.0f);
  API_END
}

int mcpm_force_spectra_T(void* stream, const void* in3, void* out1, int nx, int ny, int nz, int lap_fd,
                         int grad_fd, float kcut, int deconv_order, int half_weights, int accumulate) {
  API_BEGIN
  return force_spectra_T(as_stream(stream), C(in3), C(out1), nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order,
                         half_weights, accumulate, 1.0f);
  API_END
}

int mcpm_hessian_spectra(void* stream, const void* delta_k, void* out6, int nx, int ny, int nz, int lap_fd,
                         int grad_fd) {
  API_BEGIN
  return hessian_spectra(as_stream(stream), C(delta_k), C(out6), nx, ny, nz, lap_fd, grad_fd, 1.0f);
  API_END
}

int mcpm_hessian_spectra_T(void* stream, const void* in6, void* out1, int nx, int ny, int nz, int lap_fd,
                           int grad_fd, int half_weights, int accumulate) {
  API_BEGIN
  return hessian_spectra_T(as_stream(stream), C(in6), C(out1), nx, ny, nz, lap_fd, grad_fd, half_weights, accumulate,
                           1.0f);
  API_END
}

int mcpm_lpt2_source(void* stream, const float* h6, float* d2, int64_t n) {
  API_BEGIN
  return lpt2_source(as_stream(stream), h6, d2, n);
  API_END
}

int mcpm_lpt2_source_vjp(void* stream, const float* h6, const float* d2bar, float* hbar6, int64_t n) {
  API_BEGIN
  return lpt2_source_vjp(as_stream(stream), h6, d2bar, hbar6, n);
  API_END
}

int mcpm_deconv(void* stream, const void* in, void* out, int nx, int ny, int nz, int order) {
  API_BEGIN
  return deconv(as_stream(stream), C(in), C(out), nx, ny, nz, order);
  API_END
}

int mcpm_interlace_combine(void* stream, const void* in_m, void* out, int m, int nx, int ny, int nz, float scale,
                           int deconv_order) {
  API_BEGIN
  return interlace_combine(as_stream(stream), C(in_m), C(out), m, nx, ny, nz, scale, deconv_order);
  API_END
}

int mcpm_interlace_combine_T(void* stream, const void* in, void* out_m, int m, int nx, int ny, int nz, float scale,
                             int deconv_order) {
  API_BEGIN
  return interlace_combine_T(as_stream(stream), C(in), C(out_m), m, nx, ny, nz, scale, deconv_order,
                             (float)((double)nx * ny * nz));
  API_END
}

int mcpm_interlace_combine_slab(void* stream, const void* in_m, void* out, int m, int nx, int ny, int nz, int ny_loc,
                                int y0, float scale, int deconv_order) {
  API_BEGIN
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny, "interlace_combine_slab: bad ky block");
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  return interlace_combine(as_stream(stream), C(in_m), C(out), m, nx, ny, nz, scale, deconv_order, sk);
  API_END
}

int mcpm_interlace_combine_T_slab(void* stream, const void* in, void* out_m, int m, int nx, int ny, int nz, int ny_loc,
                                  int y0, float scale, int deconv_order, int half_weights, float norm) {
  API_BEGIN
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny, "interlace_combine_T_slab: bad ky block");
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  return interlace_combine_T(as_stream(stream), C(in), C(out_m), m, nx, ny, nz, scale, deconv_order, norm, sk,
                             half_weights);
  API_END
}

int mcpm_chreshape_crop_xz_slab(void* stream, const void* in, int inx, int iny, int inz, int rows, int y0,
                                const void* nyq_plane, const float* row_w, void* out, int onx, int onz, float scale) {
  API_BEGIN
  NEED(in && out && rows >= 0 && y0 >= 0 && y0 + rows <= iny, "chreshape_crop_xz_slab: bad arguments");
  return chreshape_crop_xz_slab(as_stream(stream), C(in), inx, iny, inz, rows, y0, nyq_plane ? C(nyq_plane) : nullptr,
                                row_w, C(out), onx, onz, scale);
  API_END
}

int mcpm_chreshape_crop_xz_slab_vjp(void* stream, const void* outbar, int onx, int onz, int rows, int y0,
                                    const float* row_w, void* inbar, int inx, int iny, int inz, void* nyq_plane_bar,
                                    float scale) {
  API_BEGIN
  NEED(outbar && inbar && rows >= 0 && y0 >= 0 && y0 + rows <= iny, "chreshape_crop_xz_slab_vjp: bad arguments");
  return chreshape_crop_xz_slab_T(as_stream(stream), C(outbar), onx, onz, rows, y0, row_w, C(inbar), inx, iny, inz,
                                  nyq_plane_bar ? C(nyq_plane_bar) : nullptr, scale);
  API_END
}

int mcpm_hermitian_project(void* stream, void* data_c64, int nx, int ny, int nz, int batch) {
  API_BEGIN
  NEED(data_c64 && batch >= 1, "hermitian_project: bad arguments");
  return hermitian_project(as_stream(stream), C(data_c64), nx, ny, nz, batch);
  API_END
}

int mcpm_half_weight_axpy(void* stream, const void* in, void* out, int64_t nc, int nz, float a, int inverse,
                          int accumulate) {
  API_BEGIN
  NEED(in && out && nc >= 0, "half_weight_axpy: bad arguments");
  return half_weight_axpy(as_stream(stream), C(in), C(out), nc, nz, a, inverse, accumulate);
  API_END
}

// ---- slab-decomposed building blocks ----------------------------------------------------------------------------
int mcpm_slabfft_create(int nx, int ny, int nz, int parts, mcpm_slabfft** out) {
  API_BEGIN
  NEED(out, "slabfft_create: null out");
  SlabFft* f = slabfft_create(nx, ny, nz, parts);
  if (!f) return MCPM_ECUFFT;
  *out = new mcpm_slabfft{f, rt_get_device()};
  return MCPM_OK;
  API_END
}

int mcpm_slabfft_destroy(mcpm_slabfft* h) {
  if (!h) return MCPM_OK;
  slabfft_destroy(h->f);
  delete h;
  return MCPM_OK;
}

int mcpm_slabfft_r2c_yz(mcpm_slabfft* h, void* stream, const float* in, void* out_c64, int nb) {
  API_BEGIN
  NEED(h && in && out_c64 && nb >= 1, "slabfft_r2c_yz: bad arguments");
  if (int b = rt_bind_device(h->device)) return b;
  return slabfft_r2c_yz(h->f, as_stream(stream), in, C(out_c64), nb);
  API_END
}

int mcpm_slabfft_c2r_yz(mcpm_slabfft* h, void* stream, void* in_c64, float* out, int nb) {
  API_BEGIN
  NEED(h && in_c64 && out && nb >= 1, "slabfft_c2r_yz: bad arguments");
  if (int b = rt_bind_device(h->device)) return b;
  return slabfft_c2r_yz(h->f, as_stream(stream), C(in_c64), out, nb);
  API_END
}

int mcpm_slabfft_c2c_x(mcpm_slabfft* h, void* stream, void* data_c64, int nb, int inverse) {
  API_BEGIN
  NEED(h && data_c64 && nb >= 1, "slabfft_c2c_x: bad arguments");
  if (int b = rt_bind_device(h->device)) return b;
  return slabfft_c2c_x(h->f, as_stream(stream), C(data_c64), nb, inverse);
  API_END
}

int mcpm_force_spectra_slab(void* stream, const void* delta_k, void* out3, int nx, int ny, int nz, int ny_loc, int y0,
                            int lap_fd, int grad_fd, float kcut, int deconv_order, float norm) {
  API_BEGIN
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny, "force_spectra_slab: bad ky block");
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  return force_spectra(as_stream(stream), C(delta_k), C(out3), nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk);
  API_END
}

int mcpm_xfuse_supported(int nx) {
#ifndef MCPM_HOSTEMU
  return xfuse_supported(nx) ? 1 : 0;
#else
  (void)nx;
  return 0;
#endif
}

int mcpm_xfuse_force_slab(void* stream, const void* in, void* out3, int nx, int ny, int nz, int ny_loc, int y0,
                          int lap_fd, int grad_fd, float kcut, int deconv_order, float norm) {
  API_BEGIN
  NEED(in && out3, "xfuse_force_slab: null pointer");
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny && nz > 0 && !(nz & 1), "xfuse_force_slab: bad shape or ky block");
#ifndef MCPM_HOSTEMU
  if (xfuse_supported(nx)) {
    SlabK sk;
    sk.ny_loc = ny_loc;
    sk.y0 = y0;
    return xfuse_force(as_stream(stream), C(in), C(out3), nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk);
  }
#endif
  set_error("xfuse_force_slab: nx must be 64, 128, 256, 512 or 1024 (CUDA build only)");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_xfuse_peer(void* stream, int mode, const void* const* in_peers, void* const* out_peers, void* k_local, int npeer,
                    int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, int half_weights, int accumulate,
                    float norm) {
  API_BEGIN
  NEED(npeer >= 1 && npeer <= 8, "xfuse_peer: 1..8 ranks");
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny && nz > 0 && !(nz & 1), "xfuse_peer: bad shape or ky block");
#ifndef MCPM_HOSTEMU
  if (xfuse_supported(nx)) {
    const cfloat* ip[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cfloat* op[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const bool x_in = mode == 0 || mode == 1 || mode == 4 || mode == 5, x_out = mode >= 0 && mode <= 3;
    NEED(!x_in || in_peers, "xfuse_peer: this mode reads x-space planes from the peers");
    NEED(!x_out || out_peers, "xfuse_peer: this mode writes x-space planes to the peers");
    for (int r = 0; r < npeer; ++r) {
      if (x_in) {
        NEED(in_peers[r], "xfuse_peer: null peer buffer");
        ip[r] = C(in_peers[r]);
      }
      if (x_out) {
        NEED(out_peers[r], "xfuse_peer: null peer buffer");
        op[r] = C(out_peers[r]);
      }
    }
    return xfuse_peer_mode(as_stream(stream), mode, ip, op, C(k_local), npeer, nx, ny, nz, ny_loc, y0, lap_fd, grad_fd,
                           half_weights, accumulate, norm);
  }
#endif
  set_error("xfuse_peer: nx must be 64, 128, 256, 512 or 1024 (CUDA build only)");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_xfuse_force_peer(void* stream, const void* const* in_peers, void* const* out_peers, int npeer, int transpose,
                          int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, float kcut,
                          int deconv_order, float norm) {
  API_BEGIN
  NEED(in_peers && out_peers, "xfuse_force_peer: null pointer table");
  NEED(npeer >= 1 && npeer <= 8, "xfuse_force_peer: 1..8 ranks");
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny && nz > 0 && !(nz & 1), "xfuse_force_peer: bad shape or ky block");
#ifndef MCPM_HOSTEMU
  if (xfuse_supported(nx)) {
    const cfloat* ip[8];
    cfloat* op[8];
    for (int r = 0; r < npeer; ++r) {
      NEED(in_peers[r] && out_peers[r], "xfuse_force_peer: null peer buffer");
      ip[r] = C(in_peers[r]);
      op[r] = C(out_peers[r]);
    }
    return xfuse_force_peer(as_stream(stream), ip, op, npeer, transpose, nx, ny, nz, ny_loc, y0, lap_fd, grad_fd, kcut,
                            deconv_order, norm);
  }
#endif
  set_error("xfuse_force_peer: nx must be 64, 128, 256, 512 or 1024 (CUDA build only)");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_xfuse_force_T_slab(void* stream, const void* in3, void* out1, int nx, int ny, int nz, int ny_loc, int y0,
                            int lap_fd, int grad_fd, float kcut, int deconv_order, float norm) {
  API_BEGIN
  NEED(in3 && out1, "xfuse_force_T_slab: null pointer");
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny && nz > 0 && !(nz & 1), "xfuse_force_T_slab: bad shape or ky block");
#ifndef MCPM_HOSTEMU
  if (xfuse_supported(nx)) {
    SlabK sk;
    sk.ny_loc = ny_loc;
    sk.y0 = y0;
    return xfuse_force_T(as_stream(stream), C(in3), C(out1), nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk);
  }
#endif
  set_error("xfuse_force_T_slab: nx must be 64, 128, 256, 512 or 1024 (CUDA build only)");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_force_spectra_T_slab(void* stream, const void* in3, void* out1, int nx, int ny, int nz, int ny_loc, int y0,
                              int lap_fd, int grad_fd, float kcut, int deconv_order, int half_weights, int accumulate,
                              float norm) {
  API_BEGIN
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny, "force_spectra_T_slab: bad ky block");
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  return force_spectra_T(as_stream(stream), C(in3), C(out1), nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order,
                         half_weights, accumulate, norm, sk);
  API_END
}

int mcpm_hessian_spectra_slab(void* stream, const void* delta_k, void* out6, int nx, int ny, int nz, int ny_loc, int y0,
                              int lap_fd, int grad_fd, float norm) {
  API_BEGIN
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny, "hessian_spectra_slab: bad ky block");
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  return hessian_spectra(as_stream(stream), C(delta_k), C(out6), nx, ny, nz, lap_fd, grad_fd, norm, sk);
  API_END
}

int mcpm_hessian_spectra_T_slab(void* stream, const void* in6, void* out1, int nx, int ny, int nz, int ny_loc, int y0,
                                int lap_fd, int grad_fd, int half_weights, int accumulate, float norm) {
  API_BEGIN
  NEED(ny_loc > 0 && y0 >= 0 && y0 + ny_loc <= ny, "hessian_spectra_T_slab: bad ky block");
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  return hessian_spectra_T(as_stream(stream), C(in6), C(out1), nx, ny, nz, lap_fd, grad_fd, half_weights, accumulate,
                           norm, sk);
  API_END
}

int mcpm_chreshape(void* stream, const void* in, int inx, int iny, int inz, void* out, int onx, int ony, int onz) {
  API_BEGIN
  NEED(in && out && in != out, "chreshape: in and out must be distinct buffers");
  return chreshape(as_stream(stream), C(in), inx, iny, inz, C(out), onx, ony, onz);
  API_END
}

int mcpm_chreshape_vjp(void* stream, const void* outbar, int onx, int ony, int onz, void* inbar, int inx, int iny,
                       int inz) {
  API_BEGIN
  NEED(outbar && inbar && outbar != inbar, "chreshape_vjp: buffers must be distinct");
  return chreshape_T(as_stream(stream), C(outbar), onx, ony, onz, C(inbar), inx, iny, inz);
  API_END
}

int mcpm_spectrum_bins(void* stream, const void* m0_c64, const void* m1_c64, int nx, int ny, int nz, double box_x,
                       double box_y, double box_z, const double* kedges, int n_edges, int deconv0, int deconv1,
                       double* out) {
  API_BEGIN
  NEED(m0_c64 && kedges && out, "spectrum_bins: null pointer");
  NEED(box_x > 0 && box_y > 0 && box_z > 0, "spectrum_bins: box sizes must be positive");
  return spectrum_bins(as_stream(stream), C(m0_c64), m1_c64 ? C(m1_c64) : nullptr, nx, ny, nz, box_x, box_y, box_z,
                       kedges, n_edges, deconv0, deconv1, out);
  API_END
}

int mcpm_spectrum_bins_ell(void* stream, const void* m0_c64, const void* m1_c64, int nx, int ny, int nz, double box_x,
                           double box_y, double box_z, const double* kedges, int n_edges, int deconv0, int deconv1,
                           int ell, const double los[3], double* out) {
  API_BEGIN
  NEED(m0_c64 && kedges && out && los, "spectrum_bins_ell: null pointer");
  NEED(box_x > 0 && box_y > 0 && box_z > 0, "spectrum_bins_ell: box sizes must be positive");
  return spectrum_bins(as_stream(stream), C(m0_c64), m1_c64 ? C(m1_c64) : nullptr, nx, ny, nz, box_x, box_y, box_z,
                       kedges, n_edges, deconv0, deconv1, out, ell, los[0], los[1], los[2]);
  API_END
}

int mcpm_rg2cgh(void* stream, const float* mesh, void* out_c64, int nx, int ny, int nz, float scale,
                const float* transfer) {
  API_BEGIN
  NEED(mesh && out_c64, "rg2cgh: null pointer");
  return rg2cgh(as_stream(stream), mesh, C(out_c64), nx, ny, nz, scale, transfer);
  API_END
}

int mcpm_rg2cgh_vjp(void* stream, const void* outbar_c64, float* meshbar, int nx, int ny, int nz, float scale,
                    const float* transfer) {
  API_BEGIN
  NEED(outbar_c64 && meshbar, "rg2cgh_vjp: null pointer");
  return rg2cgh_vjp(as_stream(stream), C(outbar_c64), meshbar, nx, ny, nz, scale, transfer);
  API_END
}

int mcpm_cgh2rg(void* stream, const void* meshk_c64, float* mesh, int nx, int ny, int nz, float inv_scale) {
  API_BEGIN
  NEED(meshk_c64 && mesh, "cgh2rg: null pointer");
  return cgh2rg(as_stream(stream), C(meshk_c64), mesh, nx, ny, nz, inv_scale);
  API_END
}

int mcpm_hermitian_weights(void* stream, const void* in, void* out, int nx, int ny, int nz, int mode) {
  API_BEGIN
  return hermitian_weights(as_stream(stream), C(in), C(out), nx, ny, nz, mode);
  API_END
}

int mcpm_axpby(void* stream, const float* x, float a, const float* y, float b, float c, int64_t n, float* out) {
  API_BEGIN
  return axpby(as_stream(stream), x, a, y, b, c, n, out);
  API_END
}

int mcpm_dot(void* stream, const float* a, const float* b, int64_t n, double* out) {
  API_BEGIN
  return dot_accum(as_stream(stream), a, b, n, 1.0, out);
  API_END
}

int mcpm_yz_gradients(void* stream, void* buf3, int xl, int ny, int nz, int grad_fd, int transpose) {
  API_BEGIN
  NEED(buf3, "yz_gradients: null pointer");
  return yz_gradients(as_stream(stream), C(buf3), xl, ny, nz, grad_fd, transpose);
  API_END
}

int mcpm_absmax(void* stream, const float* x, int64_t n, int stride, float* out) {
  API_BEGIN
  NEED(out && (x || n == 0) && stride >= 1, "absmax: null pointer or stride < 1");
  return absmax_strided(as_stream(stream), x, n, stride, out);
  API_END
}

int mcpm_rsd_shift(void* stream, const float* pos, const float* vel, const float los[3], float coef, int64_t np,
                   float* pos_out) {
  API_BEGIN
  NEED(los, "rsd_shift: null los");
  return rsd_shift(as_stream(stream), pos, vel, los[0], los[1], los[2], coef, np, pos_out);
  API_END
}

int mcpm_rsd_shift_vjp(void* stream, const float* posbar, const float los[3], float coef, int64_t np, float* velbar,
                       int accumulate) {
  API_BEGIN
  NEED(los, "rsd_shift_vjp: null los");
  return rsd_shift_vjp(as_stream(stream), posbar, los[0], los[1], los[2], coef, np, velbar, accumulate);
  API_END
}

int mcpm_bias_spectra(void* stream, const void* delta_k, int nx, int ny, int nz, const float cpl[3],
                      const float* inv_transfer, void* out) {
  API_BEGIN
  NEED(delta_k && out && cpl, "bias_spectra: null pointer");
  return bias_spectra(as_stream(stream), C(delta_k), nx, ny, nz, cpl[0], cpl[1], cpl[2], inv_transfer, C(out));
  API_END
}

int mcpm_bias_spectra_vjp(void* stream, const void* outbar, int nx, int ny, int nz, const float cpl[3],
                          const float* inv_transfer, void* dkbar, int accumulate) {
  API_BEGIN
  NEED(outbar && dkbar && cpl, "bias_spectra_vjp: null pointer");
  return bias_spectra_T(as_stream(stream), C(outbar), nx, ny, nz, cpl[0], cpl[1], cpl[2], inv_transfer, C(dkbar), accumulate);
  API_END
}

int mcpm_shear_invariants(void* stream, const float* s5, int64_t n, float* out2) {
  API_BEGIN
  NEED(s5 && out2, "shear_invariants: null pointer");
  return shear_invariants(as_stream(stream), s5, n, out2);
  API_END
}

int mcpm_shear_invariants_vjp(void* stream, const float* s5, const float* out2bar, int64_t n, float* s5bar) {
  API_BEGIN
  NEED(s5 && out2bar && s5bar, "shear_invariants_vjp: null pointer");
  return shear_invariants_vjp(as_stream(stream), s5, out2bar, n, s5bar);
  API_END
}

static BiasCoef bias_coef(const float c[13]) {
  return BiasCoef{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12]};
}

int mcpm_bias_moments(void* stream, const float* vals, int K, float growth, const float* growth_arr, int64_t np,
                      double* mom) {
  API_BEGIN
  NEED(mom && (vals || np == 0), "bias_moments: null pointer");
  NEED(K == 7 || K == 9, "bias_moments: K must be 7 or 9");
  return bias_moments(as_stream(stream), vals, K, growth, growth_arr, np, mom);
  API_END
}

int mcpm_bias_weights(void* stream, const float* vals, int K, float growth, const float* growth_arr, const float coef[13],
                      const double* mom, int64_t np, float* weights, float* dvel) {
  API_BEGIN
  NEED(coef && mom && ((vals && weights) || np == 0), "bias_weights: null pointer");
  NEED(K == 7 || K == 9, "bias_weights: K must be 7 or 9");
  NEED(np > 0, "bias_weights: the particle means need at least one particle");
  return bias_weights(as_stream(stream), vals, K, growth, growth_arr, bias_coef(coef), mom, np, weights, dvel);
  API_END
}

int mcpm_bias_weights_vjp(void* stream, const float* vals, int K, float growth, const float* growth_arr,
                          const float coef[13], const double* mom, const float* wbar, const float* dvelbar, int64_t np,
                          double* msum, float* valsbar, double* coefbar, float* gbar_arr) {
  API_BEGIN
  NEED(coef && mom && msum && coefbar && vals && wbar && valsbar, "bias_weights_vjp: null pointer");
  NEED(K == 7 || K == 9, "bias_weights_vjp: K must be 7 or 9");
  NEED(np > 0, "bias_weights_vjp: the particle means need at least one particle");
  NEED(!growth_arr || gbar_arr, "bias_weights_vjp: per-particle growth needs gbar_arr");
  return bias_weights_vjp(as_stream(stream), vals, K, growth, growth_arr, bias_coef(coef), mom, wbar, dvelbar, np, msum,
                          valsbar, coefbar, growth_arr ? gbar_arr : nullptr);
  API_END
}

int mcpm_scale_spectrum(void* stream, const void* in, const float* t, void* out, int64_t nc) {
  API_BEGIN
  return scale_spectrum(as_stream(stream), C(in), t, C(out), nc);
  API_END
}

int mcpm_lpt_combine(void* stream, const float* pos, const float* f1, const float* f2, float d1, float d2,
                     float dv2, int64_t np, float* dpos, float* vel, float* pos_out) {
  API_BEGIN
  return lpt_combine(as_stream(stream), pos, f1, f2, d1, d2, dv2, np, dpos, vel, pos_out);
  API_END
}

int mcpm_kick_drift(void* stream, float* pos, float* vel, const float* fmesh3, int64_t np, int nx, int ny, int nz,
                    int order, float alpha, float beta, float drift, float* force_out) {
  API_BEGIN
  return kick_drift(as_stream(stream), pos, vel, fmesh3, np, nx, ny, nz, order, alpha, beta, drift, pos, vel,
                    force_out);
  API_END
}

int mcpm_interleave3(void* stream, const float* planar3, float* mesh4, int64_t n) {
  API_BEGIN
  return interleave3(as_stream(stream), planar3, mesh4, n);
  API_END
}

int mcpm_deinterleave3(void* stream, const float* mesh4, float* planar3, int64_t n) {
  API_BEGIN
  return deinterleave3(as_stream(stream), mesh4, planar3, n);
  API_END
}

int mcpm_kick_drift4(void* stream, float* pos, float* vel, const float* fmesh4, int64_t np, int nx, int ny, int nz,
                     float alpha, float beta, float drift) {
  API_BEGIN
  NEED(pos && vel && fmesh4, "kick_drift4: null pointer");
  return kick_drift4(as_stream(stream), pos, vel, fmesh4, np, nx, ny, nz, alpha, beta, drift, pos, vel);
  API_END
}

int mcpm_paint3v4(void* stream, const float* pos, float* vbar, const float* xbar, float drift, float scale, int64_t np,
                  int nx, int ny, int nz, float* mesh4) {
  API_BEGIN
  NEED(pos && vbar && mesh4, "paint3v4: null pointer");
  return paint3v4(as_stream(stream), pos, vbar, xbar, drift, xbar != nullptr, scale, np, nx, ny, nz, mesh4);
  API_END
}

int mcpm_read_grad4v(void* stream, const float* pos, const float* fmesh4, const float* rhobar, float* vbar,
                     float cscale, float alpha, int64_t np, int nx, int ny, int nz, float* xbar) {
  API_BEGIN
  NEED(pos && fmesh4 && rhobar && vbar && xbar, "read_grad4v: null pointer");
  return read_grad4v(as_stream(stream), pos, fmesh4, rhobar, vbar, cscale, 1, alpha, np, nx, ny, nz, xbar, 1);
  API_END
}

// ---- the same particle kernels with a position frame (mcpm_frame: absolute or lattice-relative positions) -----------
int mcpm_paint_f(void* stream, const mcpm_frame* frame, const float* pos, const float* weights, float wscalar,
                 int64_t np, int nx, int ny, int nz, int order, const float scale[3], float shift, float* mesh,
                 int accumulate) {
  API_BEGIN
  FRAME(np);
  return paint(as_stream(stream), pos, weights, wscalar, np, nx, ny, nz, order, scale, shift, mesh, accumulate, 0.0f,
               fr);
  API_END
}

int mcpm_read_f(void* stream, const mcpm_frame* frame, const float* pos, const float* mesh, int nmesh, int64_t np,
                int nx, int ny, int nz, int order, const float scale[3], float shift, float* out) {
  API_BEGIN
  FRAME(np);
  return read(as_stream(stream), pos, mesh, nmesh, np, nx, ny, nz, order, scale, shift, out, 0.0f, fr);
  API_END
}

int mcpm_read_grad_f(void* stream, const mcpm_frame* frame, const float* pos, const float* mesh, int nmesh,
                     const float* cot, int64_t np, int nx, int ny, int nz, int order, const float scale[3],
                     float shift, float* grad, int accumulate) {
  API_BEGIN
  NEED(nmesh >= 1 && nmesh <= 4, "read_grad: nmesh must be 1..4");
  FRAME(np);
  const int64_t plane = (int64_t)nx * ny * nz;
  const float* ms[4] = {mesh, mesh + plane, mesh + 2 * plane, mesh + 3 * plane};
  return read_grad(as_stream(stream), pos, ms, nmesh, cot, cot ? nmesh : 0, 1.0f, nullptr, np, nx, ny, nz, order,
                   scale, shift, grad, accumulate, 0.0f, fr);
  API_END
}

int mcpm_paint_vjp_f(void* stream, const mcpm_frame* frame, const float* pos, const float* weights, float wscalar,
                     const float* mesh_bar, int64_t np, int nx, int ny, int nz, int order, const float scale[3],
                     float shift, float* posbar, float* weightsbar, int accumulate) {
  API_BEGIN
  FRAME(np);
  return paint_vjp(as_stream(stream), pos, weights, wscalar, mesh_bar, np, nx, ny, nz, order, scale, shift, posbar,
                   weightsbar, accumulate, 0.0f, fr);
  API_END
}

int mcpm_paint3_f(void* stream, const mcpm_frame* frame, const float* pos, const float* vals3, float vscale, int64_t np,
                  int nx, int ny, int nz, int order, float* mesh3, int accumulate) {
  API_BEGIN
  FRAME(np);
  return paint3(as_stream(stream), pos, vals3, vscale, nullptr, 0.0f, np, nx, ny, nz, order, mesh3, accumulate, fr);
  API_END
}

int mcpm_kick_drift_f(void* stream, const mcpm_frame* frame, float* pos, float* vel, const float* fmesh3, int64_t np,
                      int nx, int ny, int nz, int order, float alpha, float beta, float drift, float* force_out) {
  API_BEGIN
  FRAME(np);
  return kick_drift(as_stream(stream), pos, vel, fmesh3, np, nx, ny, nz, order, alpha, beta, drift, pos, vel,
                    force_out, fr);
  API_END
}

int mcpm_kick_drift4_f(void* stream, const mcpm_frame* frame, float* pos, float* vel, const float* fmesh4, int64_t np,
                       int nx, int ny, int nz, float alpha, float beta, float drift) {
  API_BEGIN
  NEED(pos && vel && fmesh4, "kick_drift4: null pointer");
  FRAME(np);
  return kick_drift4(as_stream(stream), pos, vel, fmesh4, np, nx, ny, nz, alpha, beta, drift, pos, vel, nullptr, 0, fr);
  API_END
}

int mcpm_paint3v4_f(void* stream, const mcpm_frame* frame, const float* pos, float* vbar, const float* xbar,
                    float drift, float scale, int64_t np, int nx, int ny, int nz, float* mesh4) {
  API_BEGIN
  NEED(pos && vbar && mesh4, "paint3v4: null pointer");
  FRAME(np);
  return paint3v4(as_stream(stream), pos, vbar, xbar, drift, xbar != nullptr, scale, np, nx, ny, nz, mesh4, fr);
  API_END
}

int mcpm_read_grad4v_f(void* stream, const mcpm_frame* frame, const float* pos, const float* fmesh4,
                       const float* rhobar, float* vbar, float cscale, float alpha, int64_t np, int nx, int ny, int nz,
                       float* xbar) {
  API_BEGIN
  NEED(pos && fmesh4 && rhobar && vbar && xbar, "read_grad4v: null pointer");
  FRAME(np);
  return read_grad4v(as_stream(stream), pos, fmesh4, rhobar, vbar, cscale, 1, alpha, np, nx, ny, nz, xbar, 1, nullptr,
                     0, fr);
  API_END
}

int mcpm_read_grad4v_step_f(void* stream, const mcpm_frame* frame, const float* pos, const float* fmesh4,
                            const float* rhobar, float* vbar, float cscale, float alpha, float dnext, int64_t np, int nx,
                            int ny, int nz, float* xbar) {
  API_BEGIN
  NEED(pos && fmesh4 && rhobar && vbar && xbar, "read_grad4v_step: null pointer");
  FRAME(np);
  return read_grad4v(as_stream(stream), pos, fmesh4, rhobar, vbar, cscale, 1, alpha, np, nx, ny, nz, xbar, 1, nullptr,
                     0, fr, dnext);
  API_END
}

int mcpm_paint_brick_f(void* stream, const mcpm_frame* frame, int px, int py, int pz, const float* pos,
                       const float* weights, float wscalar, float shift, int64_t np, int nx, int ny, int nz,
                       float* mesh) {
  API_BEGIN
  NEED(pos && mesh, "paint_brick: null pointer");
#ifndef MCPM_HOSTEMU
  FRAME(np);
  Lattice L;
  L.px = px;
  L.py = py;
  L.pz = pz;
  int r = brick_paint_cic(as_stream(stream), L, pos, weights, wscalar, shift, np, nx, ny, nz, mesh, fr);
  if (r < 0) return MCPM_ECUDA;
  if (r == 1) return MCPM_OK;
#endif
  set_error("paint_brick: unsupported lattice / mesh / frame geometry, or CPU build");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_paint3_brick_f(void* stream, const mcpm_frame* frame, int px, int py, int pz, const float* pos, float* vbar,
                        const float* xbar, float drift, float scale, int64_t np, int nx, int ny, int nz,
                        float* mesh3) {
  API_BEGIN
  NEED(pos && vbar && mesh3, "paint3_brick: null pointer");
#ifndef MCPM_HOSTEMU
  FRAME(np);
  Lattice L;
  L.px = px;
  L.py = py;
  L.pz = pz;
  if (xbar)
    if (int e = axpy3(as_stream(stream), vbar, xbar, drift, 3 * np, vbar)) return e;
  int r = brick_paint3_cic(as_stream(stream), L, pos, vbar, scale, np, nx, ny, nz, mesh3, fr);
  if (r < 0) return MCPM_ECUDA;
  if (r == 1) return MCPM_OK;
#endif
  set_error("paint3_brick: unsupported lattice / mesh / frame geometry, or CPU build");
  return MCPM_EUNSUP;
  API_END
}

int mcpm_halo_reduce_peer(void* stream, float* own_ext, const float* prev_ext, const float* next_ext, int halo, int xl,
                          int64_t plane, int nlead, int active) {
  API_BEGIN
  NEED(own_ext && prev_ext && next_ext && halo >= 1 && halo <= xl && plane > 0 && nlead >= 1, "halo_reduce_peer: bad arguments");
  NEED(active >= 0 && active <= halo, "halo_reduce_peer: active planes must be in 0..halo (0 = all)");
  return halo_reduce_peer(as_stream(stream), own_ext, prev_ext, next_ext, halo, xl, plane, nlead, active ? active : halo);
  API_END
}

int mcpm_halo_gather_peer(void* stream, float* own_ext, const float* prev_ext, const float* next_ext, int halo, int xl,
                          int64_t plane, int nlead, int active) {
  API_BEGIN
  NEED(own_ext && prev_ext && next_ext && halo >= 1 && halo <= xl && plane > 0 && nlead >= 1, "halo_gather_peer: bad arguments");
  NEED(active >= 0 && active <= halo, "halo_gather_peer: active planes must be in 0..halo (0 = all)");
  return halo_gather_peer(as_stream(stream), own_ext, prev_ext, next_ext, halo, xl, plane, nlead, active ? active : halo);
  API_END
}

int mcpm_halo_gather4_peer(void* stream, float* fmesh4_ext, const float* f3_own, const float* f3_prev, const float* f3_next,
                           int halo, int xl, int64_t plane, int active) {
  API_BEGIN
  NEED(fmesh4_ext && f3_own && f3_prev && f3_next && halo >= 1 && halo <= xl && plane > 0, "halo_gather4_peer: bad arguments");
  NEED(active >= 0 && active <= halo, "halo_gather4_peer: active planes must be in 0..halo (0 = all)");
  return halo_gather4_peer(as_stream(stream), fmesh4_ext, f3_own, f3_prev, f3_next, halo, xl, plane, active ? active : halo);
  API_END
}

int mcpm_drift(void* stream, float* pos, const float* vel, float drift, int64_t np) {
  API_BEGIN
  return axpy3(as_stream(stream), pos, vel, drift, 3 * np, pos);
  API_END
}

int mcpm_pm_forces(mcpm_engine* eng, void* stream, const float* pos, int64_t np, int order, int paint_deconv,
                   int lap_fd, int grad_fd, float kcut, float* fmesh3, float* forces) {
  API_BEGIN
  NEED(eng, "null engine");
  BIND(eng);
  return pm_forces(eng->e, as_stream(stream), pos, np, order, paint_deconv, lap_fd, grad_fd, kcut, fmesh3, forces);
  API_END
}

int mcpm_pm_forces_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* fbar, const float* fmesh3,
                       int64_t np, int order, int paint_deconv, int lap_fd, int grad_fd, float kcut,
                       float* posbar, int accumulate) {
  API_BEGIN
  NEED(eng && fmesh3 && fbar && posbar, "pm_forces_vjp: null pointer");
  BIND(eng);
  return pm_forces_vjp(eng->e, as_stream(stream), pos, fbar, 1.0f, fmesh3, np, order, paint_deconv, lap_fd, grad_fd,
                       kcut, posbar, accumulate);
  API_END
}

int mcpm_pm_forces_mesh(mcpm_engine* eng, void* stream, const float* pos, const void* delta_k, int64_t np,
                        int order, int lap_fd, int grad_fd, float kcut, float* forces) {
  API_BEGIN
  NEED(eng && delta_k && forces, "pm_forces_mesh: null pointer");
  BIND(eng);
  return pm_forces_mesh(eng->e, as_stream(stream), pos, C(delta_k), np, order, lap_fd, grad_fd, kcut, forces);
  API_END
}

int mcpm_pm_forces2(mcpm_engine* eng, void* stream, const float* pos, const void* delta_k, int64_t np, int order,
                    int lap_fd, int grad_fd, float* forces, float* h6_out) {
  API_BEGIN
  NEED(eng && delta_k && forces, "pm_forces2: null pointer");
  BIND(eng);
  return pm_forces2(eng->e, as_stream(stream), pos, C(delta_k), np, order, lap_fd, grad_fd, forces, h6_out);
  API_END
}

int mcpm_lpt(mcpm_engine* eng, void* stream, const void* delta_k, const float* pos, int64_t np, int lpt_order,
             int read_order, int lap_fd, int grad_fd, float d1, float d2, float dv2, float* dpos, float* vel,
             float* f1, float* f2, float* h6) {
  API_BEGIN
  NEED(eng && delta_k, "lpt: null pointer");  // pos == NULL: the particles sit on the cells of the mesh
  BIND(eng);
  return lpt(eng->e, as_stream(stream), C(delta_k), pos, np, lpt_order, read_order, lap_fd, grad_fd, d1, d2, dv2,
             dpos, vel, f1, f2, h6);
  API_END
}

int mcpm_lpt_vjp(mcpm_engine* eng, void* stream, const float* pos, int64_t np, int lpt_order, int read_order,
                 int lap_fd, int grad_fd, float d1, float d2, float dv2, const float* dposbar, const float* velbar,
                 const float* f1, const float* f2, const float* h6, void* dkbar, double* coefbar, int accumulate) {
  API_BEGIN
  NEED(eng && dposbar && velbar && dkbar, "lpt_vjp: null pointer");
  BIND(eng);
  return lpt_vjp(eng->e, as_stream(stream), pos, np, lpt_order, read_order, lap_fd, grad_fd, d1, d2, dv2, dposbar,
                 velbar, f1, f2, h6, C(dkbar), coefbar, accumulate);
  API_END
}

int mcpm_nbody_steps(mcpm_engine* eng, void* stream, float* pos, float* vel, int64_t np, int n_steps,
                     const float* alpha, const float* beta, const float* drift_pre, const float* drift_post,
                     int order, int paint_deconv, int lap_fd, int grad_fd, float* xk, float* vk, float* fm) {
  API_BEGIN
  NEED(eng && pos && vel, "nbody_steps: null pointer");
  BIND(eng);
  return nbody_steps(eng->e, as_stream(stream), pos, vel, np, n_steps, alpha, beta, drift_pre, drift_post, order,
                     paint_deconv, lap_fd, grad_fd, xk, vk, fm);
  API_END
}

int mcpm_nbody_steps_vjp(mcpm_engine* eng, void* stream, float* posbar, float* velbar, int64_t np, int n_steps,
                         const float* alpha, const float* beta, const float* drift_pre, const float* drift_post,
                         int order, int paint_deconv, int lap_fd, int grad_fd, const float* xk, const float* vk,
                         const float* fm, const float* v0, double* coefbar) {
  API_BEGIN
  NEED(eng && posbar && velbar, "nbody_steps_vjp: null pointer");
  BIND(eng);
  return nbody_steps_vjp(eng->e, as_stream(stream), posbar, velbar, np, n_steps, alpha, beta, drift_pre, drift_post,
                         order, paint_deconv, lap_fd, grad_fd, xk, vk, fm, v0, coefbar);
  API_END
}

int mcpm_nufft(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar, int64_t np,
               const float scale[3], int paint_order, int interlace_order, int paint_deconv, void* out_k) {
  API_BEGIN
  NEED(eng && out_k && (np == 0 || pos), "nufft: null pointer");  // an empty particle set paints a zero mesh
  BIND(eng);
  return nufft(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order,
               paint_deconv, C(out_k));
  API_END
}

int mcpm_nufft_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar,
                   int64_t np, const float scale[3], int paint_order, int interlace_order, int paint_deconv,
                   const void* outbar_k, float* posbar, float* weightsbar) {
  API_BEGIN
  NEED(eng && outbar_k && (np == 0 || pos), "nufft_vjp: null pointer");
  BIND(eng);
  return nufft_vjp(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order,
                   paint_deconv, C(outbar_k), posbar, weightsbar);
  API_END
}

int mcpm_nufft_rsd(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const float los[3], float coef,
                   const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order,
                   int interlace_order, int paint_deconv, void* out_k) {
  API_BEGIN
  NEED(eng && out_k && los && (np == 0 || (pos && vel)), "nufft_rsd: null pointer");
  BIND(eng);
  const ObsShift obs = {vel, los[0], los[1], los[2], coef};
  return nufft(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order, paint_deconv,
               C(out_k), 0.0f, &obs);
  API_END
}

int mcpm_nufft_rsd_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const float los[3], float coef,
                       const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order,
                       int interlace_order, int paint_deconv, const void* outbar_k, float* posbar, float* velbar,
                       float* weightsbar) {
  API_BEGIN
  NEED(eng && outbar_k && los && (np == 0 || (pos && vel)), "nufft_rsd_vjp: null pointer");
  BIND(eng);
  const ObsShift obs = {vel, los[0], los[1], los[2], coef};
  return nufft_vjp(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order,
                   paint_deconv, C(outbar_k), posbar, weightsbar, 0.0f, &obs, velbar);
  API_END
}

// mcpm_obs (ABI) -> ObsGen (kernels)
static_assert(kObsSlots == MCPM_OBS_SLOTS, "obs.h and mcpm.h must agree");
static int make_obs_gen(const mcpm_obs* o, const float* vel, int64_t np, ObsGen& g) {
  if (!o) {
    set_error("nufft_obs: null observation descriptor");
    return MCPM_EINVAL;
  }
  if ((o->lightcone && o->rsd && !o->tab_gf) || (o->ap == 1 && !o->tab_ap) || ((o->lightcone || o->ap == 1) && (o->nt < 2 || !(o->dr > 0.0f)))) {
    set_error("nufft_obs: the light cone / ap_auto need their radius tables (nt >= 2, dr > 0)");
    return MCPM_EINVAL;
  }
  if (o->ap < 0 || o->ap > 2 || (o->rsd && !vel && np > 0)) {  // an empty particle set may come with null arrays
    set_error("nufft_obs: ap must be 0 | 1 | 2, and rsd needs velocities");
    return MCPM_EINVAL;
  }
  g.on = 1;
  g.curved = o->curved != 0;
  g.lightcone = o->lightcone != 0;
  g.ap = o->ap;
  g.rsd = o->rsd != 0;
  g.cx = o->cell[0], g.cy = o->cell[1], g.cz = o->cell[2];
  g.ox = o->origin[0], g.oy = o->origin[1], g.oz = o->origin[2];
  g.lx = o->los[0], g.ly = o->los[1], g.lz = o->los[2];
  g.gf = o->gf, g.a_par = o->a_par, g.a_perp = o->a_perp;
  g.par = o->par;
  g.r0 = o->r0;
  g.inv_dr = o->dr > 0.0f ? 1.0f / o->dr : 0.0f;
  g.nt = o->nt > 0 ? o->nt : 0;
  g.tab_gf = o->tab_gf, g.tab_ap = o->tab_ap;
  g.vel = vel, g.dvel = o->rsd ? o->dvel : nullptr;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) g.rt[3 * a + b] = o->rot[3 * b + a];
  return 0;
}

int mcpm_nufft_obs(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const mcpm_obs* obs,
                   const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order, float kcut,
                   int interlace_order, int paint_deconv, void* out_k) {
  API_BEGIN
  NEED(eng && out_k && (np == 0 || pos || eng->e->rel), "nufft_obs: null pointer");
  BIND(eng);
  ObsGen g;
  if (int e = make_obs_gen(obs, vel, np, g)) return e;
  ObsShift sh;
  sh.gen = &g;
  return nufft(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order, paint_deconv,
               C(out_k), kcut > 0.0f ? kcut : 0.0f, &sh);
  API_END
}

int mcpm_nufft_obs_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const mcpm_obs* obs,
                       const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order, float kcut,
                       int interlace_order, int paint_deconv, const void* outbar_k, float* posbar, float* velbar,
                       float* dvelbar, float* weightsbar, double* parbar) {
  API_BEGIN
  NEED(eng && outbar_k && (np == 0 || pos || eng->e->rel), "nufft_obs_vjp: null pointer");
  BIND(eng);
  ObsGen g;
  if (int e = make_obs_gen(obs, vel, np, g)) return e;
  stream_t st = as_stream(stream);
  const int64_t row = 3 + 2 * (int64_t)g.nt;
  g.dvelbar = g.dvel ? dvelbar : nullptr;
  g.parbar = parbar;
  if (parbar && rt_memset(parbar, 0, sizeof(double) * (size_t)(kObsSlots * row), st)) return MCPM_ECUDA;
  ObsShift sh;
  sh.gen = &g;
  if (int e = nufft_vjp(eng->e, st, pos, weights, wscalar, np, scale, paint_order, interlace_order, paint_deconv,
                        C(outbar_k), posbar, weightsbar, kcut > 0.0f ? kcut : 0.0f, &sh, velbar))
    return e;
  return parbar ? obs_reduce_slots(st, parbar, row) : 0;  // the slots' partial sums into row 0
  API_END
}

static int make_radial_geom(const mcpm_obs* o, ObsGen& g) {
  if (!o || o->nt < 2 || !(o->dr > 0.0f)) {
    set_error("radial_tables: needs the radius grid (nt >= 2, dr > 0)");
    return MCPM_EINVAL;
  }
  g.on = 1;
  g.curved = o->curved != 0;
  g.cx = o->cell[0], g.cy = o->cell[1], g.cz = o->cell[2];
  g.ox = o->origin[0], g.oy = o->origin[1], g.oz = o->origin[2];
  g.lx = o->los[0], g.ly = o->los[1], g.lz = o->los[2];
  g.r0 = o->r0;
  g.inv_dr = 1.0f / o->dr;
  g.nt = o->nt;
  return 0;
}

int mcpm_radial_tables(void* stream, const float* pos, int64_t np, const mcpm_obs* geom, int ntab, const float* tabs,
                       float* out) {
  API_BEGIN
  NEED(ntab >= 1 && tabs && (np == 0 || (pos && out)), "radial_tables: null pointer or no table");
  ObsGen g;
  if (int e = make_radial_geom(geom, g)) return e;
  return radial_tables(as_stream(stream), pos, np, g, ntab, tabs, out);
  API_END
}

int mcpm_radial_tables_vjp(void* stream, const float* pos, int64_t np, const mcpm_obs* geom, int ntab, const float* tabs,
                           const float* outbar, float* posbar, double* tabbar) {
  API_BEGIN
  NEED(ntab >= 1 && tabs && (np == 0 || (pos && outbar)), "radial_tables_vjp: null pointer or no table");
  ObsGen g;
  if (int e = make_radial_geom(geom, g)) return e;
  stream_t st = as_stream(stream);
  const int64_t row = (int64_t)ntab * g.nt;
  if (tabbar && rt_memset(tabbar, 0, sizeof(double) * (size_t)(kObsSlots * row), st)) return MCPM_ECUDA;
  if (int e = radial_tables_vjp(st, pos, np, g, ntab, tabs, outbar, posbar, tabbar)) return e;
  return tabbar ? obs_reduce_slots(st, tabbar, row) : 0;
  API_END
}

int mcpm_nufft_kb(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar, int64_t np,
                  const float scale[3], int paint_order, float kcut, int interlace_order, int paint_deconv,
                  void* out_k) {
  API_BEGIN
  NEED(eng && out_k && (np == 0 || pos), "nufft_kb: null pointer");
  NEED(kcut > 0.0f, "nufft_kb: kcut must be positive (optim_kcut(oversamp))");
  BIND(eng);
  return nufft(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order,
               paint_deconv, C(out_k), kcut);
  API_END
}

int mcpm_nufft_vjp_kb(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar,
                      int64_t np, const float scale[3], int paint_order, float kcut, int interlace_order,
                      int paint_deconv, const void* outbar_k, float* posbar, float* weightsbar) {
  API_BEGIN
  NEED(eng && outbar_k && (np == 0 || pos), "nufft_vjp_kb: null pointer");
  NEED(kcut > 0.0f, "nufft_vjp_kb: kcut must be positive (optim_kcut(oversamp))");
  BIND(eng);
  return nufft_vjp(eng->e, as_stream(stream), pos, weights, wscalar, np, scale, paint_order, interlace_order,
                   paint_deconv, C(outbar_k), posbar, weightsbar, kcut);
  API_END
}

#pragma GCC visibility pop
}  // extern "C"
