// mcpm_xla.cc -- XLA-FFI shim over the C ABI of include/mcpm.h: one typed custom-call handler per entry point that
// INTEGRATION.md section 1 maps to a callable of montecosmo/nbody.py (file:line cited per handler).  montecosmo_b200/jax_nbody.py
// registers these symbols with jax.ffi and wraps them in jax.custom_vjp rules under the reference's own names.
//
// STATUS: this file is NOT COMPILED in this repository's image -- jaxlib and its xla/ffi/api/ffi.h are absent
// and nothing here is on the tested path; the same ABI is exercised through ctypes (montecosmo_b200/_capi.py).  It is kept
// compile-guarded (the build needs -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__),
// 'include'))") and -lmcpm) and argument-checked: tests/test_xla_shim.py parses every mcpm_* call below and compares its
// argument count with the ctypes prototype table, so that the shim cannot drift from the header silently.
//
// Conventions
//   * Optional array arguments are passed as zero-size buffers and reach the ABI as NULL (`opt()`).
//   * Engines (cuFFT plans + scratch, mcpm_engine_create) live in a process-global map keyed by (device ordinal, mesh
//     shape), created on the first call of that shape -- never on a captured / command-buffer path, XLA runs each custom
//     call once eagerly before capturing it.
//   * `lattice` = {px, py, pz} (attribute, {0,0,0} = none) sets the particle-lattice hint; `relative` != 0 additionally
//     makes `pos` lattice-relative (mcpm_engine_set_relative).  Both are set per call, under the engine's mutex.
//   * Complex cotangents: the ABI returns dL/dRe + i dL/dIm; JAX's convention for a real loss is its complex conjugate,
//     which jax_nbody.py applies (jnp.conj) on the way out -- not here.
//   * In-place state update (mcpm_nbody_steps) is expressed on the Python side with input_output_aliases; the handler
//     copies nothing when XLA honoured the alias and falls back to one device-to-device copy when it did not.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include <array>
#include <complex>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "mcpm.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using C64 = ffi::Buffer<ffi::C64>;
using F64 = ffi::Buffer<ffi::F64>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RC64 = ffi::ResultBuffer<ffi::C64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using Ints = ffi::Span<const int32_t>;
using Floats = ffi::Span<const float>;

template <class B>
auto* opt(B& b) {  // zero-size buffer == absent
  return b.element_count() == 0 ? nullptr : b.typed_data();
}
template <class B>
auto* optr(B& b) {
  return b->element_count() == 0 ? nullptr : b->typed_data();
}
ffi::Error status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(mcpm_last_error()); }

struct EngineSlot {
  mcpm_engine* eng = nullptr;
  std::mutex mu;
};
std::mutex g_map_mu;
std::map<std::tuple<int, int, int, int>, EngineSlot*> g_engines;

EngineSlot* engine_for(int device, int nx, int ny, int nz) {
  std::lock_guard<std::mutex> lock(g_map_mu);
  auto key = std::make_tuple(device, nx, ny, nz);
  auto it = g_engines.find(key);
  if (it != g_engines.end()) return it->second;
  auto* slot = new EngineSlot;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  int rc = mcpm_engine_create(nx, ny, nz, &slot->eng);
  cudaSetDevice(prev);
  if (rc != 0) {
    delete slot;
    return nullptr;
  }
  g_engines[key] = slot;
  return slot;
}

// Engine of a real mesh shape with the lattice hint / relative mode of this call applied.
struct EngineCall {
  EngineSlot* slot;
  std::unique_lock<std::mutex> lock;
  ffi::Error err = ffi::Error::Success();
  EngineCall(int device, int nx, int ny, int nz, Ints lattice, int32_t relative) : slot(engine_for(device, nx, ny, nz)) {
    if (!slot) {
      err = ffi::Error::Internal(mcpm_last_error());
      return;
    }
    lock = std::unique_lock<std::mutex>(slot->mu);
    int px = lattice.size() == 3 ? lattice[0] : 0, py = lattice.size() == 3 ? lattice[1] : 0,
        pz = lattice.size() == 3 ? lattice[2] : 0;
    int rc = mcpm_engine_set_relative(slot->eng, 0);
    if (rc == 0) rc = mcpm_engine_set_lattice(slot->eng, px, py, pz);
    if (rc == 0 && relative) rc = mcpm_engine_set_relative(slot->eng, 1);
    if (rc != 0) err = ffi::Error::Internal(mcpm_last_error());
  }
  mcpm_engine* eng() const { return slot->eng; }
};

template <class B>
std::array<int, 3> mesh3(const B& b, int lead = 0) {  // last three dimensions
  auto d = b.dimensions();
  size_t n = d.size();
  (void)lead;
  return {(int)d[n - 3], (int)d[n - 2], (int)d[n - 1]};
}
std::array<int, 3> real_shape_of_spectrum(const C64& k) {  // [.., nx, ny, nz/2+1] -> (nx, ny, nz)
  auto d = k.dimensions();
  size_t n = d.size();
  return {(int)d[n - 3], (int)d[n - 2], 2 * ((int)d[n - 1] - 1)};
}
const float* scale3(Floats s, float* buf) {
  for (int i = 0; i < 3; ++i) buf[i] = s.size() == 3 ? s[i] : 1.0f;
  return buf;
}

// ---------------------------------------------------------------------------------------------------- mass assignment
// paint (nbody.py:365-396); kcut > 0 selects kernel_type='kaiser_bessel' (nbody.py:280-290, 383-384)
ffi::Error Paint(cudaStream_t st, F32 pos, F32 weights, float wscalar, int32_t order, float kcut, Floats scale, float shift,
                 RF32 mesh) {
  auto m = mesh3(*mesh);
  float sc[3];
  int64_t np = pos.dimensions()[0];
  if (kcut > 0.0f)
    return status(mcpm_paint_kb(st, pos.typed_data(), opt(weights), wscalar, np, m[0], m[1], m[2], order, kcut,
                                scale3(scale, sc), shift, mesh->typed_data(), 0));
  return status(mcpm_paint(st, pos.typed_data(), opt(weights), wscalar, np, m[0], m[1], m[2], order, scale3(scale, sc),
                           shift, mesh->typed_data(), 0));
}
// VJP of paint: mesh cotangent -> (posbar, weightsbar)
ffi::Error PaintVjp(cudaStream_t st, F32 pos, F32 weights, F32 mesh_bar, float wscalar, int32_t order, float kcut,
                    Floats scale, float shift, RF32 posbar, RF32 weightsbar) {
  auto m = mesh3(mesh_bar);
  float sc[3];
  int64_t np = pos.dimensions()[0];
  if (kcut > 0.0f)
    return status(mcpm_paint_vjp_kb(st, pos.typed_data(), opt(weights), wscalar, mesh_bar.typed_data(), np, m[0], m[1],
                                    m[2], order, kcut, scale3(scale, sc), shift, optr(posbar), optr(weightsbar), 0));
  return status(mcpm_paint_vjp(st, pos.typed_data(), opt(weights), wscalar, mesh_bar.typed_data(), np, m[0], m[1], m[2],
                               order, scale3(scale, sc), shift, optr(posbar), optr(weightsbar), 0));
}
// read (nbody.py:398-427): mesh [nmesh, nx, ny, nz] -> out [np, nmesh]
ffi::Error Read(cudaStream_t st, F32 pos, F32 mesh, int32_t order, float kcut, Floats scale, float shift, RF32 out) {
  auto m = mesh3(mesh);
  float sc[3];
  int64_t np = pos.dimensions()[0];
  int nmesh = mesh.dimensions().size() == 4 ? (int)mesh.dimensions()[0] : 1;
  if (kcut > 0.0f)
    return status(mcpm_read_kb(st, pos.typed_data(), mesh.typed_data(), nmesh, np, m[0], m[1], m[2], order, kcut,
                               scale3(scale, sc), shift, out->typed_data()));
  return status(mcpm_read(st, pos.typed_data(), mesh.typed_data(), nmesh, np, m[0], m[1], m[2], order, scale3(scale, sc),
                          shift, out->typed_data()));
}
// VJP of read w.r.t. pos (cot = out_bar [np, nmesh], zero-size = ones)
ffi::Error ReadGrad(cudaStream_t st, F32 pos, F32 mesh, F32 cot, int32_t order, float kcut, Floats scale, float shift,
                    RF32 grad) {
  auto m = mesh3(mesh);
  float sc[3];
  int64_t np = pos.dimensions()[0];
  int nmesh = mesh.dimensions().size() == 4 ? (int)mesh.dimensions()[0] : 1;
  if (kcut > 0.0f)
    return status(mcpm_read_grad_kb(st, pos.typed_data(), mesh.typed_data(), nmesh, opt(cot), np, m[0], m[1], m[2], order,
                                    kcut, scale3(scale, sc), shift, grad->typed_data(), 0));
  return status(mcpm_read_grad(st, pos.typed_data(), mesh.typed_data(), nmesh, opt(cot), np, m[0], m[1], m[2], order,
                               scale3(scale, sc), shift, grad->typed_data(), 0));
}
// VJP of a 3-mesh read w.r.t. the meshes
ffi::Error Paint3(cudaStream_t st, F32 pos, F32 vals3, float vscale, int32_t order, RF32 mesh3_out) {
  auto m = mesh3(*mesh3_out);
  return status(mcpm_paint3(st, pos.typed_data(), vals3.typed_data(), vscale, pos.dimensions()[0], m[0], m[1], m[2], order,
                            mesh3_out->typed_data(), 0));
}

// ---------------------------------------------------------------------------------------------------- FFT, Fourier passes
// jnp.fft.rfftn / irfftn call sites nbody.py:589, 603, 620, 627, 630
ffi::Error Rfftn(cudaStream_t st, int32_t dev, F32 in, RC64 out) {
  auto m = mesh3(in);
  EngineCall e(dev, m[0], m[1], m[2], Ints(), 0);
  if (e.err.failure()) return e.err;
  int batch = in.dimensions().size() == 4 ? (int)in.dimensions()[0] : 1;
  return status(mcpm_rfftn(e.eng(), st, in.typed_data(), out->typed_data(), batch));
}
// mcpm_irfftn overwrites its input: the Python side donates it (input_output_aliases) or XLA hands us a copy
ffi::Error Irfftn(cudaStream_t st, int32_t dev, C64 in, RF32 out) {
  auto m = mesh3(*out);
  EngineCall e(dev, m[0], m[1], m[2], Ints(), 0);
  if (e.err.failure()) return e.err;
  int batch = in.dimensions().size() == 4 ? (int)in.dimensions()[0] : 1;
  return status(mcpm_irfftn(e.eng(), st, in.typed_data(), out->typed_data(), batch));
}
// transposes of rfftn / irfftn in the real inner product (mode 0 / 1)
ffi::Error HermitianWeights(cudaStream_t st, C64 in, int32_t mode, RC64 out) {
  auto m = real_shape_of_spectrum(in);
  return status(mcpm_hermitian_weights(st, in.typed_data(), out->typed_data(), m[0], m[1], m[2], mode));
}
// deconv_paint on a half spectrum (nbody.py:315-334); kcut > 0: Kaiser-Bessel transform (293-312)
ffi::Error Deconv(cudaStream_t st, C64 in, int32_t order, float kcut, RC64 out) {
  auto m = real_shape_of_spectrum(in);
  if (kcut > 0.0f) return status(mcpm_deconv_kb(st, in.typed_data(), out->typed_data(), m[0], m[1], m[2], order, kcut));
  return status(mcpm_deconv(st, in.typed_data(), out->typed_data(), m[0], m[1], m[2], order));
}
// utils.chreshape (utils.py:975-1013) and its transpose
ffi::Error Chreshape(cudaStream_t st, C64 in, RC64 out) {
  auto a = real_shape_of_spectrum(in);
  auto b = real_shape_of_spectrum(*out);
  return status(mcpm_chreshape(st, in.typed_data(), a[0], a[1], a[2], out->typed_data(), b[0], b[1], b[2]));
}
ffi::Error ChreshapeVjp(cudaStream_t st, C64 outbar, RC64 inbar) {
  auto b = real_shape_of_spectrum(outbar);
  auto a = real_shape_of_spectrum(*inbar);
  return status(mcpm_chreshape_vjp(st, outbar.typed_data(), b[0], b[1], b[2], inbar->typed_data(), a[0], a[1], a[2]));
}
// utils.rg2cgh / cgh2rg (utils.py:785-921) with the fused transfer multiply of samp2base_mesh (bricks.py:305-309)
ffi::Error Rg2cgh(cudaStream_t st, F32 mesh, F32 transfer, float scale, RC64 out) {
  auto m = mesh3(mesh);
  return status(mcpm_rg2cgh(st, mesh.typed_data(), out->typed_data(), m[0], m[1], m[2], scale, opt(transfer)));
}
ffi::Error Rg2cghVjp(cudaStream_t st, C64 outbar, F32 transfer, float scale, RF32 meshbar) {
  auto m = mesh3(*meshbar);
  return status(mcpm_rg2cgh_vjp(st, outbar.typed_data(), meshbar->typed_data(), m[0], m[1], m[2], scale, opt(transfer)));
}
ffi::Error Cgh2rg(cudaStream_t st, C64 meshk, float inv_scale, RF32 mesh) {
  auto m = mesh3(*mesh);
  return status(mcpm_cgh2rg(st, meshk.typed_data(), mesh->typed_data(), m[0], m[1], m[2], inv_scale));
}
// white2lin (bricks.py:152-157): out = in * transfer
ffi::Error ScaleSpectrum(cudaStream_t st, C64 in, F32 transfer, RC64 out) {
  return status(mcpm_scale_spectrum(st, in.typed_data(), transfer.typed_data(), out->typed_data(),
                                    (int64_t)in.element_count()));
}
// metrics.spectrum binning (metrics.py:121-182): out [4, n_edges + 1] float64, zero-initialised by the caller
ffi::Error SpectrumBins(cudaStream_t st, C64 m0, C64 m1, F64 kedges, F64 acc_in, Floats box, Ints deconv, int32_t ell,
                        Floats los, RF64 out) {
  auto m = real_shape_of_spectrum(m0);
  if (out->typed_data() != acc_in.typed_data())
    cudaMemcpyAsync(out->typed_data(), acc_in.typed_data(), sizeof(double) * acc_in.element_count(),
                    cudaMemcpyDeviceToDevice, st);
  const double l3[3] = {los.size() == 3 ? los[0] : 0.0, los.size() == 3 ? los[1] : 0.0, los.size() == 3 ? los[2] : 0.0};
  return status(mcpm_spectrum_bins_ell(st, m0.typed_data(), opt(m1), m[0], m[1], m[2], (double)box[0], (double)box[1],
                                       (double)box[2], kedges.typed_data(), (int)kedges.element_count(), deconv[0],
                                       deconv[1], ell, l3, out->typed_data()));
}

// ---------------------------------------------------------------------------------------------------- composites
// nufft at the paint shape (nbody.py:532-577; interlace 513-529): pos in FINAL units, scale = paint / final
ffi::Error Nufft(cudaStream_t st, int32_t dev, F32 pos, F32 weights, float wscalar, Floats scale, int32_t paint_order,
                 float kcut, int32_t interlace_order, int32_t paint_deconv, Ints lattice, int32_t relative, RC64 out) {
  auto m = real_shape_of_spectrum(*out);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  float sc[3];
  int64_t np = pos.dimensions()[0];
  if (kcut > 0.0f)
    return status(mcpm_nufft_kb(e.eng(), st, pos.typed_data(), opt(weights), wscalar, np, scale3(scale, sc), paint_order,
                                kcut, interlace_order, paint_deconv, out->typed_data()));
  return status(mcpm_nufft(e.eng(), st, pos.typed_data(), opt(weights), wscalar, np, scale3(scale, sc), paint_order,
                           interlace_order, paint_deconv, out->typed_data()));
}
ffi::Error NufftVjp(cudaStream_t st, int32_t dev, F32 pos, F32 weights, C64 outbar, float wscalar, Floats scale,
                    int32_t paint_order, float kcut, int32_t interlace_order, int32_t paint_deconv, Ints lattice,
                    int32_t relative, RF32 posbar, RF32 weightsbar) {
  auto m = real_shape_of_spectrum(outbar);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  float sc[3];
  int64_t np = pos.dimensions()[0];
  if (kcut > 0.0f)
    return status(mcpm_nufft_vjp_kb(e.eng(), st, pos.typed_data(), opt(weights), wscalar, np, scale3(scale, sc),
                                    paint_order, kcut, interlace_order, paint_deconv, outbar.typed_data(), optr(posbar),
                                    optr(weightsbar)));
  return status(mcpm_nufft_vjp(e.eng(), st, pos.typed_data(), opt(weights), wscalar, np, scale3(scale, sc), paint_order,
                               interlace_order, paint_deconv, outbar.typed_data(), optr(posbar), optr(weightsbar)));
}
// nufft of redshift-space positions, the flat-sky shift applied inside the paint kernels (model.py:780-809 with
// bricks.py:781-792); its VJP returns posbar, velbar and weightsbar from one gather
ffi::Error NufftRsd(cudaStream_t st, int32_t dev, F32 pos, F32 vel, F32 weights, Floats los, float coef, float wscalar,
                    Floats scale, int32_t paint_order, int32_t interlace_order, int32_t paint_deconv, Ints lattice,
                    int32_t relative, RC64 out) {
  auto m = real_shape_of_spectrum(*out);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  float sc[3];
  return status(mcpm_nufft_rsd(e.eng(), st, pos.typed_data(), vel.typed_data(), los.begin(), coef, opt(weights), wscalar,
                               pos.dimensions()[0], scale3(scale, sc), paint_order, interlace_order, paint_deconv,
                               out->typed_data()));
}
ffi::Error NufftRsdVjp(cudaStream_t st, int32_t dev, F32 pos, F32 vel, F32 weights, C64 outbar, Floats los, float coef,
                       float wscalar, Floats scale, int32_t paint_order, int32_t interlace_order, int32_t paint_deconv,
                       Ints lattice, int32_t relative, RF32 posbar, RF32 velbar, RF32 weightsbar) {
  auto m = real_shape_of_spectrum(outbar);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  float sc[3];
  return status(mcpm_nufft_rsd_vjp(e.eng(), st, pos.typed_data(), vel.typed_data(), los.begin(), coef, opt(weights), wscalar,
                                   pos.dimensions()[0], scale3(scale, sc), paint_order, interlace_order, paint_deconv,
                                   outbar.typed_data(), optr(posbar), velbar->typed_data(), optr(weightsbar)));
}
// nufft of OBSERVED positions, the general observation chain (model.py:780-799) applied inside the paint kernels.  What
// depends on the cosmology arrives as traced arrays: par [3] = (D f at a_obs, a_par, a_perp) and the two radius tables;
// zero-size buffers stand for "absent" (vel without rsd, dvel, the tables).
static mcpm_obs make_obs(F32 par, F32 tab_gf, F32 tab_ap, F32 dvel, Ints flags, Floats geom, Floats rot) {
  mcpm_obs o = {};
  o.curved = flags[0], o.lightcone = flags[1], o.ap = flags[2], o.rsd = flags[3];
  for (int d = 0; d < 3; ++d) o.cell[d] = geom[d], o.origin[d] = geom[3 + d], o.los[d] = geom[6 + d];
  o.r0 = geom[9], o.dr = geom[10];
  o.nt = (int)(tab_gf.element_count() ? tab_gf.element_count() : tab_ap.element_count());
  o.tab_gf = opt(tab_gf), o.tab_ap = opt(tab_ap), o.dvel = opt(dvel);
  for (int k = 0; k < 9; ++k) o.rot[k] = rot[k];
  o.par = par.typed_data();
  return o;
}
ffi::Error NufftObs(cudaStream_t st, int32_t dev, F32 pos, F32 vel, F32 dvel, F32 weights, F32 par, F32 tab_gf, F32 tab_ap,
                    Ints flags, Floats geom, Floats rot, float wscalar, Floats scale, int32_t paint_order, float kcut,
                    int32_t interlace_order, int32_t paint_deconv, Ints lattice, int32_t relative, RC64 out) {
  auto m = real_shape_of_spectrum(*out);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  float sc[3];
  const mcpm_obs o = make_obs(par, tab_gf, tab_ap, dvel, flags, geom, rot);
  return status(mcpm_nufft_obs(e.eng(), st, pos.typed_data(), opt(vel), &o, opt(weights), wscalar, pos.dimensions()[0],
                               scale3(scale, sc), paint_order, kcut, interlace_order, paint_deconv, out->typed_data()));
}
// parbar: float64 [MCPM_OBS_SLOTS, 3 + 2 nt], row 0 holds the cotangents of (par, tab_gf, tab_ap) on return
ffi::Error NufftObsVjp(cudaStream_t st, int32_t dev, F32 pos, F32 vel, F32 dvel, F32 weights, F32 par, F32 tab_gf,
                       F32 tab_ap, C64 outbar, Ints flags, Floats geom, Floats rot, float wscalar, Floats scale,
                       int32_t paint_order, float kcut, int32_t interlace_order, int32_t paint_deconv, Ints lattice,
                       int32_t relative, RF32 posbar, RF32 velbar, RF32 dvelbar, RF32 weightsbar, RF64 parbar) {
  auto m = real_shape_of_spectrum(outbar);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  float sc[3];
  const mcpm_obs o = make_obs(par, tab_gf, tab_ap, dvel, flags, geom, rot);
  return status(mcpm_nufft_obs_vjp(e.eng(), st, pos.typed_data(), opt(vel), &o, opt(weights), wscalar, pos.dimensions()[0],
                                   scale3(scale, sc), paint_order, kcut, interlace_order, paint_deconv, outbar.typed_data(),
                                   optr(posbar), optr(velbar), optr(dvelbar), optr(weightsbar), parbar->typed_data()));
}
// functions of the comoving distance at the particles (the light cone): tabs [ntab, nt] on the radius grid of geom[9..10]
ffi::Error RadialTables(cudaStream_t st, F32 pos, F32 tabs, Ints flags, Floats geom, RF32 out) {
  mcpm_obs o = {};
  o.curved = flags[0];
  for (int d = 0; d < 3; ++d) o.cell[d] = geom[d], o.origin[d] = geom[3 + d], o.los[d] = geom[6 + d];
  o.r0 = geom[9], o.dr = geom[10], o.nt = (int)tabs.dimensions()[1];
  return status(mcpm_radial_tables(st, pos.typed_data(), pos.dimensions()[0], &o, (int)tabs.dimensions()[0],
                                   tabs.typed_data(), out->typed_data()));
}
ffi::Error RadialTablesVjp(cudaStream_t st, F32 pos, F32 tabs, F32 outbar, Ints flags, Floats geom, RF32 posbar,
                           RF64 tabbar) {
  mcpm_obs o = {};
  o.curved = flags[0];
  for (int d = 0; d < 3; ++d) o.cell[d] = geom[d], o.origin[d] = geom[3 + d], o.los[d] = geom[6 + d];
  o.r0 = geom[9], o.dr = geom[10], o.nt = (int)tabs.dimensions()[1];
  return status(mcpm_radial_tables_vjp(st, pos.typed_data(), pos.dimensions()[0], &o, (int)tabs.dimensions()[0],
                                       tabs.typed_data(), outbar.typed_data(), optr(posbar), tabbar->typed_data()));
}
// Lagrangian bias expansion (bricks.py:327-452) as fused passes: Fourier multipliers, shear invariants, polynomial
ffi::Error BiasSpectra(cudaStream_t st, C64 dk, F32 inv_transfer, Floats cells_per_len, RC64 out) {
  auto m = real_shape_of_spectrum(dk);
  return status(mcpm_bias_spectra(st, dk.typed_data(), m[0], m[1], m[2], cells_per_len.begin(), opt(inv_transfer),
                                  out->typed_data()));
}
ffi::Error BiasSpectraVjp(cudaStream_t st, C64 outbar, F32 inv_transfer, Floats cells_per_len, RC64 dkbar) {
  auto m = real_shape_of_spectrum(*dkbar);
  return status(mcpm_bias_spectra_vjp(st, outbar.typed_data(), m[0], m[1], m[2], cells_per_len.begin(), opt(inv_transfer),
                                      dkbar->typed_data(), 0));
}
ffi::Error ShearInvariants(cudaStream_t st, F32 s5, RF32 out2) {
  return status(mcpm_shear_invariants(st, s5.typed_data(), (int64_t)(s5.element_count() / 5), out2->typed_data()));
}
ffi::Error ShearInvariantsVjp(cudaStream_t st, F32 s5, F32 out2bar, RF32 s5bar) {
  return status(mcpm_shear_invariants_vjp(st, s5.typed_data(), out2bar.typed_data(), (int64_t)(s5.element_count() / 5),
                                          s5bar->typed_data()));
}
// vals [np, K]; growth_arr: zero-size for a scalar growth factor.  mom (float64 [2]) is a result the backward call takes.
ffi::Error BiasWeights(cudaStream_t st, F32 vals, F32 growth_arr, float growth, Floats coef, RF32 weights, RF32 dvel,
                       RF64 mom) {
  const int64_t np = vals.dimensions()[0];
  const int K = (int)vals.dimensions()[1];
  if (cudaMemsetAsync(mom->typed_data(), 0, 2 * sizeof(double), st) != cudaSuccess) return ffi::Error::Internal("memset");
  if (int rc = mcpm_bias_moments(st, vals.typed_data(), K, growth, opt(growth_arr), np, mom->typed_data())) return status(rc);
  return status(mcpm_bias_weights(st, vals.typed_data(), K, growth, opt(growth_arr), coef.begin(), mom->typed_data(), np,
                                  weights->typed_data(), dvel->typed_data()));
}
// coefbar: float64 [14] (13 coefficients + the scalar growth factor); gbar_arr: [np] (unused with a scalar growth factor)
ffi::Error BiasWeightsVjp(cudaStream_t st, F32 vals, F32 growth_arr, F64 mom, F32 wbar, F32 dvelbar, float growth,
                          Floats coef, RF32 valsbar, RF64 coefbar, RF32 gbar_arr, RF64 msum) {
  const int64_t np = vals.dimensions()[0];
  const int K = (int)vals.dimensions()[1];
  if (cudaMemsetAsync(msum->typed_data(), 0, 2 * sizeof(double), st) != cudaSuccess ||
      cudaMemsetAsync(coefbar->typed_data(), 0, 14 * sizeof(double), st) != cudaSuccess)
    return ffi::Error::Internal("memset");
  return status(mcpm_bias_weights_vjp(st, vals.typed_data(), K, growth, opt(growth_arr), coef.begin(), mom.typed_data(),
                                      wbar.typed_data(), opt(dvelbar), np, msum->typed_data(), valsbar->typed_data(),
                                      coefbar->typed_data(), gbar_arr->typed_data()));
}
// pm_forces with a painted mesh (nbody.py:583-604): -> forces [np, 3] and the three force meshes (residual of the VJP)
ffi::Error PmForces(cudaStream_t st, int32_t dev, F32 pos, int32_t order, int32_t paint_deconv, int32_t lap_fd,
                    int32_t grad_fd, float kcut, Ints lattice, int32_t relative, RF32 forces, RF32 fmesh3) {
  auto m = mesh3(*fmesh3);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  return status(mcpm_pm_forces(e.eng(), st, pos.typed_data(), pos.dimensions()[0], order, paint_deconv, lap_fd, grad_fd,
                               kcut, fmesh3->typed_data(), forces->typed_data()));
}
ffi::Error PmForcesVjp(cudaStream_t st, int32_t dev, F32 pos, F32 fbar, F32 fmesh3, int32_t order, int32_t paint_deconv,
                       int32_t lap_fd, int32_t grad_fd, float kcut, Ints lattice, int32_t relative, RF32 posbar) {
  auto m = mesh3(fmesh3);
  EngineCall e(dev, m[0], m[1], m[2], lattice, relative);
  if (e.err.failure()) return e.err;
  return status(mcpm_pm_forces_vjp(e.eng(), st, pos.typed_data(), fbar.typed_data(), fmesh3.typed_data(),
                                   pos.dimensions()[0], order, paint_deconv, lap_fd, grad_fd, kcut, posbar->typed_data(),
                                   0));
}
// pm_forces with a given spectrum (nbody.py:595-604) and pm_forces2 (607-631)
ffi::Error PmForcesMesh(cudaStream_t st, int32_t dev, F32 pos, C64 dk, int32_t order, int32_t lap_fd, int32_t grad_fd,
                        float kcut, RF32 forces) {
  auto m = real_shape_of_spectrum(dk);
  EngineCall e(dev, m[0], m[1], m[2], Ints(), 0);
  if (e.err.failure()) return e.err;
  return status(mcpm_pm_forces_mesh(e.eng(), st, pos.typed_data(), dk.typed_data(), pos.dimensions()[0], order, lap_fd,
                                    grad_fd, kcut, forces->typed_data()));
}
ffi::Error PmForces2(cudaStream_t st, int32_t dev, F32 pos, C64 dk, int32_t order, int32_t lap_fd, int32_t grad_fd,
                     RF32 forces, RF32 h6) {
  auto m = real_shape_of_spectrum(dk);
  EngineCall e(dev, m[0], m[1], m[2], Ints(), 0);
  if (e.err.failure()) return e.err;
  return status(mcpm_pm_forces2(e.eng(), st, pos.typed_data(), dk.typed_data(), pos.dimensions()[0], order, lap_fd,
                                grad_fd, forces->typed_data(), optr(h6)));
}
// lpt (nbody.py:634-667): growth coefficients d1 = a2g(a), d2 = a2g2(a), dv2 = a2dg2dg(a) are attributes computed in JAX
ffi::Error Lpt(cudaStream_t st, int32_t dev, C64 dk, F32 pos, int32_t lpt_order, int32_t read_order, int32_t lap_fd,
               int32_t grad_fd, float d1, float d2, float dv2, RF32 dpos, RF32 vel, RF32 f1, RF32 f2, RF32 h6) {
  auto m = real_shape_of_spectrum(dk);
  EngineCall e(dev, m[0], m[1], m[2], Ints(), 0);
  if (e.err.failure()) return e.err;
  return status(mcpm_lpt(e.eng(), st, dk.typed_data(), pos.typed_data(), pos.dimensions()[0], lpt_order, read_order, lap_fd,
                         grad_fd, d1, d2, dv2, dpos->typed_data(), vel->typed_data(), optr(f1), optr(f2), optr(h6)));
}
ffi::Error LptVjp(cudaStream_t st, int32_t dev, F32 pos, F32 dposbar, F32 velbar, F32 f1, F32 f2, F32 h6,
                  int32_t lpt_order, int32_t read_order, int32_t lap_fd, int32_t grad_fd, float d1, float d2, float dv2,
                  RC64 dkbar, RF64 coefbar) {
  auto m = real_shape_of_spectrum(*dkbar);
  EngineCall e(dev, m[0], m[1], m[2], Ints(), 0);
  if (e.err.failure()) return e.err;
  if (coefbar->element_count()) cudaMemsetAsync(coefbar->typed_data(), 0, sizeof(double) * coefbar->element_count(), st);
  return status(mcpm_lpt_vjp(e.eng(), st, pos.typed_data(), pos.dimensions()[0], lpt_order, read_order, lap_fd, grad_fd, d1,
                             d2, dv2, dposbar.typed_data(), velbar.typed_data(), opt(f1), opt(f2), opt(h6),
                             dkbar->typed_data(), optr(coefbar), 0));
}
// BullFrog loop (bullfrog_vf nbody.py:902-960 under diffeqsolve(Euler), 999): per-step coefficients as array attributes;
// pos / vel are updated in place (aliased to the results); tape buffers are results that XLA owns
ffi::Error NbodySteps(cudaStream_t st, int32_t dev, F32 pos, F32 vel, Ints mesh, Floats alpha, Floats beta, Floats pre,
                      Floats post, int32_t order, int32_t paint_deconv, int32_t lap_fd, int32_t grad_fd, Ints lattice,
                      int32_t relative, RF32 pos_out, RF32 vel_out, RF32 xk, RF32 vk, RF32 fm) {
  EngineCall e(dev, mesh[0], mesh[1], mesh[2], lattice, relative);
  if (e.err.failure()) return e.err;
  const size_t bytes = sizeof(float) * pos.element_count();
  if (pos_out->typed_data() != pos.typed_data())
    cudaMemcpyAsync(pos_out->typed_data(), pos.typed_data(), bytes, cudaMemcpyDeviceToDevice, st);
  if (vel_out->typed_data() != vel.typed_data())
    cudaMemcpyAsync(vel_out->typed_data(), vel.typed_data(), bytes, cudaMemcpyDeviceToDevice, st);
  return status(mcpm_nbody_steps(e.eng(), st, pos_out->typed_data(), vel_out->typed_data(), pos.dimensions()[0],
                                 (int)alpha.size(), alpha.begin(), beta.begin(), pre.begin(), post.begin(), order,
                                 paint_deconv, lap_fd, grad_fd, optr(xk), optr(vk), optr(fm)));
}
ffi::Error NbodyStepsVjp(cudaStream_t st, int32_t dev, F32 posbar, F32 velbar, F32 xk, F32 vk, F32 fm, F32 v0, Ints mesh,
                         Floats alpha, Floats beta, Floats pre, Floats post, int32_t order, int32_t paint_deconv,
                         int32_t lap_fd, int32_t grad_fd, Ints lattice, int32_t relative, RF32 posbar_out,
                         RF32 velbar_out, RF64 coefbar) {
  EngineCall e(dev, mesh[0], mesh[1], mesh[2], lattice, relative);
  if (e.err.failure()) return e.err;
  const size_t bytes = sizeof(float) * posbar.element_count();
  if (posbar_out->typed_data() != posbar.typed_data())
    cudaMemcpyAsync(posbar_out->typed_data(), posbar.typed_data(), bytes, cudaMemcpyDeviceToDevice, st);
  if (velbar_out->typed_data() != velbar.typed_data())
    cudaMemcpyAsync(velbar_out->typed_data(), velbar.typed_data(), bytes, cudaMemcpyDeviceToDevice, st);
  if (coefbar->element_count()) cudaMemsetAsync(coefbar->typed_data(), 0, sizeof(double) * coefbar->element_count(), st);
  return status(mcpm_nbody_steps_vjp(e.eng(), st, posbar_out->typed_data(), velbar_out->typed_data(),
                                     posbar.dimensions()[0], (int)alpha.size(), alpha.begin(), beta.begin(), pre.begin(),
                                     post.begin(), order, paint_deconv, lap_fd, grad_fd, xk.typed_data(), opt(vk),
                                     opt(fm), opt(v0), optr(coefbar)));
}

}  // namespace

#define MCPM_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define MCPM_STREAM_DEV MCPM_STREAM.Ctx<ffi::DeviceOrdinal>()
#define XF .Attr<Floats>("scale").Attr<float>("shift")

XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPaint, Paint,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Attr<float>("wscalar").Attr<int32_t>("order").Attr<float>("kcut") XF.Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPaintVjp, PaintVjp,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Attr<float>("wscalar").Attr<int32_t>("order").Attr<float>("kcut") XF
        .Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmRead, Read,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Attr<int32_t>("order").Attr<float>("kcut") XF.Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmReadGrad, ReadGrad,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Attr<int32_t>("order").Attr<float>("kcut") XF.Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPaint3, Paint3,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Attr<float>("vscale").Attr<int32_t>("order").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmRfftn, Rfftn, MCPM_STREAM_DEV.Arg<F32>().Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmIrfftn, Irfftn, MCPM_STREAM_DEV.Arg<C64>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmHermitianWeights, HermitianWeights,
    MCPM_STREAM.Arg<C64>().Attr<int32_t>("mode").Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmDeconv, Deconv,
    MCPM_STREAM.Arg<C64>().Attr<int32_t>("order").Attr<float>("kcut").Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmChreshape, Chreshape, MCPM_STREAM.Arg<C64>().Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmChreshapeVjp, ChreshapeVjp, MCPM_STREAM.Arg<C64>().Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmRg2cgh, Rg2cgh, MCPM_STREAM.Arg<F32>().Arg<F32>().Attr<float>("scale").Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmRg2cghVjp, Rg2cghVjp,
    MCPM_STREAM.Arg<C64>().Arg<F32>().Attr<float>("scale").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmCgh2rg, Cgh2rg, MCPM_STREAM.Arg<C64>().Attr<float>("inv_scale").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmScaleSpectrum, ScaleSpectrum, MCPM_STREAM.Arg<C64>().Arg<F32>().Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmSpectrumBins, SpectrumBins,
    MCPM_STREAM.Arg<C64>().Arg<C64>().Arg<F64>().Arg<F64>().Attr<Floats>("box").Attr<Ints>("deconv").Attr<int32_t>("ell")
        .Attr<Floats>("los").Ret<F64>());
#define LATTICE .Attr<Ints>("lattice").Attr<int32_t>("relative")
#define FD .Attr<int32_t>("lap_fd").Attr<int32_t>("grad_fd")
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNufft, Nufft,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Attr<float>("wscalar").Attr<Floats>("scale").Attr<int32_t>("paint_order")
        .Attr<float>("kcut").Attr<int32_t>("interlace_order").Attr<int32_t>("paint_deconv") LATTICE.Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNufftVjp, NufftVjp,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<C64>().Attr<float>("wscalar").Attr<Floats>("scale")
        .Attr<int32_t>("paint_order").Attr<float>("kcut").Attr<int32_t>("interlace_order").Attr<int32_t>("paint_deconv")
            LATTICE.Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNufftRsd, NufftRsd,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Attr<Floats>("los").Attr<float>("coef").Attr<float>("wscalar")
        .Attr<Floats>("scale").Attr<int32_t>("paint_order").Attr<int32_t>("interlace_order").Attr<int32_t>("paint_deconv")
            LATTICE.Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNufftRsdVjp, NufftRsdVjp,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Arg<C64>().Attr<Floats>("los").Attr<float>("coef")
        .Attr<float>("wscalar").Attr<Floats>("scale").Attr<int32_t>("paint_order").Attr<int32_t>("interlace_order")
        .Attr<int32_t>("paint_deconv") LATTICE.Ret<F32>().Ret<F32>().Ret<F32>());
#define OBS .Attr<Ints>("flags").Attr<Floats>("geom").Attr<Floats>("rot")
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNufftObs, NufftObs,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>() OBS.Attr<float>("wscalar")
        .Attr<Floats>("scale").Attr<int32_t>("paint_order").Attr<float>("kcut").Attr<int32_t>("interlace_order")
        .Attr<int32_t>("paint_deconv") LATTICE.Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNufftObsVjp, NufftObsVjp,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<C64>() OBS
        .Attr<float>("wscalar").Attr<Floats>("scale").Attr<int32_t>("paint_order").Attr<float>("kcut")
        .Attr<int32_t>("interlace_order").Attr<int32_t>("paint_deconv")
            LATTICE.Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmRadialTables, RadialTables,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Attr<Ints>("flags").Attr<Floats>("geom").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmRadialTablesVjp, RadialTablesVjp,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Attr<Ints>("flags").Attr<Floats>("geom").Ret<F32>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmBiasSpectra, BiasSpectra,
    MCPM_STREAM.Arg<C64>().Arg<F32>().Attr<Floats>("cells_per_len").Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmBiasSpectraVjp, BiasSpectraVjp,
    MCPM_STREAM.Arg<C64>().Arg<F32>().Attr<Floats>("cells_per_len").Ret<C64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmShearInvariants, ShearInvariants, MCPM_STREAM.Arg<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmShearInvariantsVjp, ShearInvariantsVjp, MCPM_STREAM.Arg<F32>().Arg<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmBiasWeights, BiasWeights,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Attr<float>("growth").Attr<Floats>("coef").Ret<F32>().Ret<F32>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmBiasWeightsVjp, BiasWeightsVjp,
    MCPM_STREAM.Arg<F32>().Arg<F32>().Arg<F64>().Arg<F32>().Arg<F32>().Attr<float>("growth").Attr<Floats>("coef")
        .Ret<F32>().Ret<F64>().Ret<F32>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPmForces, PmForces,
    MCPM_STREAM_DEV.Arg<F32>().Attr<int32_t>("order").Attr<int32_t>("paint_deconv") FD.Attr<float>("kcut")
        LATTICE.Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPmForcesVjp, PmForcesVjp,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Attr<int32_t>("order").Attr<int32_t>("paint_deconv")
        FD.Attr<float>("kcut") LATTICE.Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPmForcesMesh, PmForcesMesh,
    MCPM_STREAM_DEV.Arg<F32>().Arg<C64>().Attr<int32_t>("order") FD.Attr<float>("kcut").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmPmForces2, PmForces2,
    MCPM_STREAM_DEV.Arg<F32>().Arg<C64>().Attr<int32_t>("order") FD.Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmLpt, Lpt,
    MCPM_STREAM_DEV.Arg<C64>().Arg<F32>().Attr<int32_t>("lpt_order").Attr<int32_t>("read_order") FD.Attr<float>("d1")
        .Attr<float>("d2").Attr<float>("dv2").Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmLptVjp, LptVjp,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Attr<int32_t>("lpt_order")
        .Attr<int32_t>("read_order") FD.Attr<float>("d1").Attr<float>("d2").Attr<float>("dv2").Ret<C64>().Ret<F64>());
#define STEPS .Attr<Ints>("mesh").Attr<Floats>("alpha").Attr<Floats>("beta").Attr<Floats>("drift_pre") \
    .Attr<Floats>("drift_post").Attr<int32_t>("order").Attr<int32_t>("paint_deconv") FD LATTICE
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNbodySteps, NbodySteps,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>() STEPS.Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(McpmNbodyStepsVjp, NbodyStepsVjp,
    MCPM_STREAM_DEV.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>() STEPS.Ret<F32>().Ret<F32>()
        .Ret<F64>());

#else  // no jaxlib headers in this image: the translation unit is empty and says so
#pragma message("mcpm_xla.cc: xla/ffi/api/ffi.h not found -- the XLA-FFI shim is not built (see INTEGRATION.md)")
#endif
