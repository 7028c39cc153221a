// engine.cu -- composite operators of montecosmo/nbody.py and their adjoints, orchestrated on one stream with no
// allocation or synchronisation on the call path: pm_forces (583-604), pm_forces2 (607-631), lpt (634-667), the
// BullFrog drift-kick-drift loop (933-951, 999), nufft (532-577).
//
// FFT normalisation: cuFFT's C2R is unnormalised; the 1/N of irfftn is folded into the Fourier pass that precedes
// each C2R (`norm` argument), so no separate scaling pass ever touches a mesh.
#include "engine.h"

namespace mcpm {

Engine* engine_create(int nx, int ny, int nz) {
  if (nx < 4 || ny < 4 || nz < 4 || (nz & 1)) {
    set_error("engine: mesh sides must be >= 4 and nz even");
    return nullptr;
  }
  Engine* e = new Engine;
  e->nx = nx;
  e->ny = ny;
  e->nz = nz;
  e->nzc = nz / 2 + 1;
  e->N = (int64_t)nx * ny * nz;
  e->Nc = (int64_t)nx * ny * e->nzc;
  e->invN = (float)(1.0 / (double)e->N);
  e->device = rt_get_device();
  e->tune = default_tune();
  size_t fftws = 0;
  e->fft = fft_create(nx, ny, nz, &fftws);
  if (!e->fft) {
    delete e;
    return nullptr;
  }
  size_t rb = sizeof(float) * (size_t)e->N * Engine::kR, cb = sizeof(cfloat) * (size_t)e->Nc * Engine::kC;
  if (rt_malloc((void**)&e->rbuf, rb) || rt_malloc((void**)&e->cbuf, cb)) {
    set_error("engine: scratch allocation failed");
    engine_destroy(e);
    return nullptr;
  }
  e->scratch_bytes = rb + cb + fftws;
#ifndef MCPM_HOSTEMU
  if (xfuse_supported(nx)) {  // optional: without the 2-D plans the engine simply keeps the 3-D cuFFT path
    e->fft2d = slabfft_create(nx, ny, nz, 1);
    e->fused_fft = e->fft2d != nullptr;
  }
#endif
  return e;
}

void engine_destroy(Engine* e) {
  if (!e) return;
  if (e->rbuf) rt_free(e->rbuf);
  if (e->cbuf) rt_free(e->cbuf);
  if (e->fft) fft_destroy(e->fft);
  if (e->fft2d) slabfft_destroy(e->fft2d);
  delete e;
}

#define TRY(x)            \
  do {                    \
    if (int _e = (x)) return _e; \
  } while (0)

// CIC paint of pos * scale + shift into a fresh mesh: brick-tiled when the engine carries a matching lattice hint and
// the positions are not rescaled (CUDA build), generic otherwise
// tune().side_zero: clear the next step's scatter meshes inside the gather kernels (1) or with memsets (0); measured: no
// net gain (the gathers slow down by what the memsets cost), kept selectable

static bool brick_path(const Engine* E, int order, const float* scale) {
#ifndef MCPM_HOSTEMU
  const bool unit = !scale || (scale[0] == 1.0f && scale[1] == 1.0f && scale[2] == 1.0f);
  return order == 2 && E->lat.px == E->nx && E->lat.py == E->ny && E->lat.pz == E->nz && unit;
#else
  return false;
#endif
}
// `prezeroed`: the caller guarantees the mesh is already zero (cleared on the side by the previous step's kernel)
static int paint_fresh(Engine* E, stream_t st, const float* pos, const float* weights, float wscalar, int64_t np,
                       int order, const float* scale, float shift, float* mesh, bool prezeroed = false,
                       const ObsShift* obs = nullptr) {
  Frame f;
  const Frame* fr = E->frame(f);
#ifndef MCPM_HOSTEMU
  if (brick_path(E, order, scale) && pos && !(obs && obs->gen)) {  // the general observation transform: generic kernels
    if (!prezeroed && rt_memset(mesh, 0, sizeof(float) * (size_t)E->N, st)) return MCPM_ECUDA;
    int r = brick_paint_cic(st, E->lat, pos, weights, wscalar, shift, np, E->nx, E->ny, E->nz, mesh, fr, obs);
    if (r < 0) return MCPM_ECUDA;
    if (r == 1) return 0;
  }
#endif
  return paint(st, pos, weights, wscalar, np, E->nx, E->ny, E->nz, order, scale, shift, mesh, 0, 0.0f, fr, obs);
}
static int paint_density(Engine* E, stream_t st, const float* pos, int64_t np, int order, float* mesh,
                         bool prezeroed = false) {
  return paint_fresh(E, st, pos, nullptr, 1.0f, np, order, nullptr, 0.0f, mesh, prezeroed);
}

// Three-mesh read at `pos`, or -- pos == NULL -- at the cells of the mesh themselves (regular_pos(mesh_shape)): a layout
// change for NGP / CIC, the generic kernel on an identity frame for TSC / PCS.
static int read3_at(Engine* E, stream_t st, const float* pos, const float* m3, int64_t np, int order, float* out) {
  Frame f;
  if (pos) return read(st, pos, m3, 3, np, E->nx, E->ny, E->nz, order, nullptr, 0.0f, out, 0.0f, E->frame(f));
  if (np != E->N) {
    set_error("pos == NULL means the particles sit on the cells of the mesh: np must equal nx*ny*nz");
    return MCPM_EINVAL;
  }
  if (order <= 2) return read_sites3(st, m3, np, out);
  f = make_rel_frame(E->nx, E->ny, E->nz, E->nx, E->ny, E->nz);
  return read(st, nullptr, m3, 3, np, E->nx, E->ny, E->nz, order, nullptr, 0.0f, out, 0.0f, &f);
}
// its transpose: mesh3 = paint3(pos; ca A + cb B), fresh meshes
static int paint3_at(Engine* E, stream_t st, const float* pos, const float* A, float ca, const float* B, float cb,
                     int64_t np, int order, float* mesh3) {
  Frame f;
  if (pos) return paint3(st, pos, A, ca, B, cb, np, E->nx, E->ny, E->nz, order, mesh3, 0, E->frame(f));
  if (np != E->N) {
    set_error("pos == NULL means the particles sit on the cells of the mesh: np must equal nx*ny*nz");
    return MCPM_EINVAL;
  }
  if (order <= 2) return paint_sites3(st, A, ca, B, cb, np, mesh3, 0);
  f = make_rel_frame(E->nx, E->ny, E->nz, E->nx, E->ny, E->nz);
  return paint3(st, nullptr, A, ca, B, cb, np, E->nx, E->ny, E->nz, order, mesh3, 0, &f);
}

// delta_k (any spectrum, preserved) -> three real force meshes fm[3][N]  (nbody.py:595-603 up to the irfftn)
static int force_meshes_from_spectrum(Engine* E, stream_t st, const cfloat* dk, int lap_fd, int grad_fd, float kcut,
                                      int deconv_order, float* fm) {
#ifndef MCPM_HOSTEMU
  if (E->fused_fft) {  // multiply + the three inverse x-transforms in one kernel, then 2-D C2R per x-plane
    TRY(xfuse_force_k(st, dk, E->c(0), E->nx, E->ny, E->nz, lap_fd, grad_fd, kcut, deconv_order, E->invN));
    TRY(slabfft_c2r_yz(E->fft2d, st, E->c(0), fm, 3));
    return 0;
  }
#endif
  TRY(force_spectra(st, dk, E->c(0), E->nx, E->ny, E->nz, lap_fd, grad_fd, kcut, deconv_order, E->invN));
  TRY(fft_c2r(E->fft, st, E->c(0), fm, 3));
  return 0;
}

// pm_forces with a painted density (mesh given as a shape tuple, nbody.py:588-604)
int pm_forces(Engine* E, stream_t st, const float* pos, int64_t np, int order, int paint_deconv, int lap_fd,
              int grad_fd, float kcut, float* fmesh3, float* forces, bool rho_prezeroed) {
  float* fm = fmesh3 ? fmesh3 : E->r(0);
  float* rho = E->r(6);
  TRY(paint_density(E, st, pos, np, order, rho, rho_prezeroed));
#ifndef MCPM_HOSTEMU
  if (E->fused_fft) {  // 2-D (y,z) R2C per x-plane, one kernel for x-FFT + force kernel + 3 inverse x-FFTs, 2-D C2R
    TRY(slabfft_r2c_yz(E->fft2d, st, rho, E->c(6), 1));
    TRY(xfuse_force(st, E->c(6), E->c(0), E->nx, E->ny, E->nz, lap_fd, grad_fd, kcut, paint_deconv ? order : 0,
                    E->invN));
    TRY(slabfft_c2r_yz(E->fft2d, st, E->c(0), fm, 3));
  } else
#endif
  {
    TRY(fft_r2c(E->fft, st, rho, E->c(6), 1));
    TRY(force_meshes_from_spectrum(E, st, E->c(6), lap_fd, grad_fd, kcut, paint_deconv ? order : 0, fm));
  }
  Frame f;
  if (forces) TRY(read(st, pos, fm, 3, np, E->nx, E->ny, E->nz, order, nullptr, 0.0f, forces, 0.0f, E->frame(f)));
  return 0;
}

// three real meshes (preserved) -> rhobar = C2R(sum_j conj(m_j) R2C(mesh_j)), the density cotangent of pm_forces
static int density_cotangent(Engine* E, stream_t st, const float* mesh3, int lap_fd, int grad_fd, float kcut,
                             int deconv_order, float* rhobar) {
#ifndef MCPM_HOSTEMU
  if (E->fused_fft) {
    TRY(slabfft_r2c_yz(E->fft2d, st, mesh3, E->c(0), 3));
    TRY(xfuse_force_T(st, E->c(0), E->c(3), E->nx, E->ny, E->nz, lap_fd, grad_fd, kcut, deconv_order, E->invN));
    TRY(slabfft_c2r_yz(E->fft2d, st, E->c(3), rhobar, 1));
    return 0;
  }
#endif
  TRY(fft_r2c(E->fft, st, mesh3, E->c(0), 3));
  TRY(force_spectra_T(st, E->c(0), E->c(3), E->nx, E->ny, E->nz, lap_fd, grad_fd, kcut, deconv_order, 0, 0, E->invN));
  TRY(fft_c2r(E->fft, st, E->c(3), rhobar, 1));
  return 0;
}

// VJP of pm_forces w.r.t. pos.  With m_j the force kernel, F_j = read(x, C2R(m_j R2C(paint(x)))):
//   phibar_j = paint(x, fbar_j);  rhobar = C2R(sum_j conj(m_j) R2C(phibar_j));
//   xbar = sum_j fbar_j * dread(x, phi_j) + dread(x, rhobar)          (one fused 4-mesh gather)
// `fbar` enters as cscale * fbar so that the BullFrog backward can pass vbar with cscale = beta.
int pm_forces_vjp(Engine* E, stream_t st, const float* pos, const float* fbar, float cscale, const float* fmesh3,
                  int64_t np, int order, int paint_deconv, int lap_fd, int grad_fd, float kcut, float* posbar,
                  int accumulate) {
  Frame f;
  const Frame* fr = E->frame(f);
  TRY(paint3(st, pos, fbar, cscale, nullptr, 0.0f, np, E->nx, E->ny, E->nz, order, E->r(0), 0, fr));
  TRY(density_cotangent(E, st, E->r(0), lap_fd, grad_fd, kcut, paint_deconv ? order : 0, E->r(3)));
  const float* ms[4] = {fmesh3, fmesh3 + E->N, fmesh3 + 2 * E->N, E->r(3)};
  TRY(read_grad(st, pos, ms, 4, fbar, 3, cscale, nullptr, np, E->nx, E->ny, E->nz, order, nullptr, 0.0f, posbar,
                accumulate, 0.0f, fr));
  return 0;
}

int pm_forces_mesh(Engine* E, stream_t st, const float* pos, const cfloat* dk, int64_t np, int order, int lap_fd,
                   int grad_fd, float kcut, float* forces) {
  TRY(force_meshes_from_spectrum(E, st, dk, lap_fd, grad_fd, kcut, 0, E->r(0)));
  TRY(read3_at(E, st, pos, E->r(0), np, order, forces));
  return 0;
}

// pm_forces2 (nbody.py:607-631): six Hessian meshes -> 2LPT source -> its spectrum -> forces
int pm_forces2(Engine* E, stream_t st, const float* pos, const cfloat* dk, int64_t np, int order, int lap_fd,
               int grad_fd, float* forces, float* h6_out) {
  float* h6 = h6_out ? h6_out : E->r(0);
#ifndef MCPM_HOSTEMU
  if (E->fused_fft) {
    TRY(xfuse_hessian_k(st, dk, E->c(0), E->nx, E->ny, E->nz, lap_fd, grad_fd, E->invN));
    TRY(slabfft_c2r_yz(E->fft2d, st, E->c(0), h6, 6));
    TRY(lpt2_source(st, h6, E->r(6), E->N));
    TRY(slabfft_r2c_yz(E->fft2d, st, E->r(6), E->c(6), 1));
    TRY(xfuse_force(st, E->c(6), E->c(0), E->nx, E->ny, E->nz, lap_fd, grad_fd, 0.0f, 0, E->invN));
    TRY(slabfft_c2r_yz(E->fft2d, st, E->c(0), E->r(0), 3));
  } else
#endif
  {
    TRY(hessian_spectra(st, dk, E->c(0), E->nx, E->ny, E->nz, lap_fd, grad_fd, E->invN));
    TRY(fft_c2r(E->fft, st, E->c(0), h6, 6));
    TRY(lpt2_source(st, h6, E->r(6), E->N));
    TRY(fft_r2c(E->fft, st, E->r(6), E->c(6), 1));
    TRY(force_meshes_from_spectrum(E, st, E->c(6), lap_fd, grad_fd, 0.0f, 0, E->r(0)));
  }
  TRY(read3_at(E, st, pos, E->r(0), np, order, forces));
  return 0;
}

int lpt(Engine* E, stream_t st, const cfloat* dk, const float* pos, int64_t np, int lpt_order, int read_order,
        int lap_fd, int grad_fd, float d1, float d2, float dv2, float* dpos, float* vel, float* f1, float* f2,
        float* h6) {
  if (lpt_order != 1 && lpt_order != 2) {
    set_error("lpt_order must be 1 or 2");
    return MCPM_EINVAL;
  }
  if (!dpos || !vel) {
    set_error("lpt: dpos and vel are required");
    return MCPM_EINVAL;
  }
  // without a tape the force arrays live in the output buffers and are combined in place
  float* F1 = f1 ? f1 : vel;
  float* F2 = f2 ? f2 : dpos;
  TRY(pm_forces_mesh(E, st, pos, dk, np, read_order, lap_fd, grad_fd, 0.0f, F1));
  if (lpt_order == 2) TRY(pm_forces2(E, st, pos, dk, np, read_order, lap_fd, grad_fd, F2, h6));
  TRY(lpt_combine(st, nullptr, F1, lpt_order == 2 ? F2 : nullptr, d1, d2, dv2, np, dpos, vel, nullptr));
  return 0;
}

// VJP of lpt w.r.t. delta_k (and the growth coefficients).  F1bar = d1*dposbar + velbar; F2bar = -d2*dposbar - dv2*velbar.
int lpt_vjp(Engine* E, stream_t st, const float* pos, int64_t np, int lpt_order, int read_order, int lap_fd,
            int grad_fd, float d1, float d2, float dv2, const float* dposbar, const float* velbar, const float* f1,
            const float* f2, const float* h6, cfloat* dkbar, double* coefbar, int accumulate) {
  if (lpt_order == 2) {
    if (!h6) {
      set_error("lpt_vjp: the Hessian tape h6 is required for lpt_order 2");
      return MCPM_EINVAL;
    }
    TRY(paint3_at(E, st, pos, dposbar, -d2, velbar, -dv2, np, read_order, E->r(0)));
    TRY(density_cotangent(E, st, E->r(0), lap_fd, grad_fd, 0.0f, 0, E->r(6)));  // d2bar (real)
    TRY(lpt2_source_vjp(st, h6, E->r(6), E->r(0), E->N));
#ifndef MCPM_HOSTEMU
    if (E->fused_fft) {
      TRY(slabfft_r2c_yz(E->fft2d, st, E->r(0), E->c(0), 6));
      TRY(xfuse_hessian_tk(st, E->c(0), dkbar, E->nx, E->ny, E->nz, lap_fd, grad_fd, 1, accumulate, 1.0f));
    } else
#endif
    {
      TRY(fft_r2c(E->fft, st, E->r(0), E->c(0), 6));
      TRY(hessian_spectra_T(st, E->c(0), dkbar, E->nx, E->ny, E->nz, lap_fd, grad_fd, 1, accumulate, 1.0f));
    }
    accumulate = 1;
  }
  TRY(paint3_at(E, st, pos, dposbar, d1, velbar, 1.0f, np, read_order, E->r(0)));
#ifndef MCPM_HOSTEMU
  if (E->fused_fft) {
    TRY(slabfft_r2c_yz(E->fft2d, st, E->r(0), E->c(0), 3));
    TRY(xfuse_force_tk(st, E->c(0), dkbar, E->nx, E->ny, E->nz, lap_fd, grad_fd, 0.0f, 0, 1, accumulate, 1.0f));
  } else
#endif
  {
    TRY(fft_r2c(E->fft, st, E->r(0), E->c(0), 3));
    TRY(force_spectra_T(st, E->c(0), dkbar, E->nx, E->ny, E->nz, lap_fd, grad_fd, 0.0f, 0, 1, accumulate, 1.0f));
  }
  if (coefbar) {
    if (!f1 || (lpt_order == 2 && !f2)) {
      set_error("lpt_vjp: coefficient cotangents need the f1 / f2 tape");
      return MCPM_EINVAL;
    }
    TRY(dot_accum(st, dposbar, f1, 3 * np, 1.0, coefbar + 0));
    if (lpt_order == 2) {
      TRY(dot_accum(st, dposbar, f2, 3 * np, -1.0, coefbar + 1));
      TRY(dot_accum(st, velbar, f2, 3 * np, -1.0, coefbar + 2));
    }
  }
  return 0;
}

// BullFrog loop (nbody.py:946-951 per step).  The trailing half drift of step s and the leading half drift of step
// s+1 use the same velocity, so they run as one drift fused into the kick kernel.
int nbody_steps(Engine* E, stream_t st, float* pos, float* vel, int64_t np, int n_steps, const float* alpha,
                const float* beta, const float* drift_pre, const float* drift_post, int order, int paint_deconv,
                int lap_fd, int grad_fd, float* xk, float* vk, float* fm) {
  if (n_steps < 0 || (n_steps > 0 && (!alpha || !beta || !drift_pre || !drift_post))) {
    set_error("nbody_steps: bad step arrays");
    return MCPM_EINVAL;
  }
  if (n_steps == 0) return 0;
  const int64_t P3 = 3 * np;
  float* cur = xk ? xk : pos;  // position at kick time of the current step
  TRY(axpy3(st, pos, vel, drift_pre[0], P3, cur));
  const float* vin = vel;
  // Tape slot per step: 4N floats.  CIC (order 2): the force mesh as float4 {Fx, Fy, Fz, 0} per cell (cic4.cu);
  // other orders: three planar meshes in the first 3N floats.
  const bool cic = (order == 2);
  Frame f;
  const Frame* fr = E->frame(f);
  for (int s = 0; s < n_steps; ++s) {
    float* slot = fm ? fm + (int64_t)s * 4 * E->N : E->r(3);
    float* planar = cic ? E->r(0) : slot;
    // under the brick path the density mesh of step s+1 is cleared by the kick kernel of step s
    // (needs the tape: without `fm` the float4 slot r(3..6) overlaps the density mesh r(6) the kick kernel would clear)
    const bool side_zero = tune().side_zero && fm && cic && brick_path(E, order, nullptr);
    TRY(pm_forces(E, st, cur, np, order, paint_deconv, lap_fd, grad_fd, 0.0f, planar, nullptr, side_zero && s > 0));
    const bool last = (s == n_steps - 1);
    float dcomb = drift_post[s] + (last ? 0.0f : drift_pre[s + 1]);
    float* xout = (xk && !last) ? xk + (int64_t)(s + 1) * P3 : pos;
    float* vout = vk ? vk + (int64_t)s * P3 : vel;
    if (cic) {
      TRY(interleave3(st, planar, slot, E->N));
      TRY(kick_drift4(st, cur, vin, slot, np, E->nx, E->ny, E->nz, alpha[s], beta[s], dcomb, xout, vout,
                      side_zero && !last ? E->r(6) : nullptr, E->N, fr));
    } else {
      TRY(kick_drift(st, cur, vin, slot, np, E->nx, E->ny, E->nz, order, alpha[s], beta[s], dcomb, xout, vout,
                     nullptr, fr));
    }
    cur = xout;
    vin = vout;
  }
  if (vk) TRY(rt_copy(vel, vk + (int64_t)(n_steps - 1) * P3, sizeof(float) * P3, st) ? MCPM_ECUDA : 0);
  return 0;
}

// Reverse sweep.  Per step (x1 = kick-time position, v1 = alpha v0 + beta F(x1), x_out = x1 + v1 * dcomb):
//   vbar += xbar * dcomb ; xbar += pm_forces_vjp(x1, beta * vbar) ; [coef cotangents] ; vbar *= alpha
// and finally vbar += xbar * drift_pre[0].  The leading update of step s-1 (and the final one) uses the xbar the gather of
// step s has just produced, so it rides in that gather's epilogue (`dnext`): vbar = alpha vbar + dnext xbar.  Only the
// first step of the sweep needs its own pass, and the 3-channel scatter reads vbar without touching xbar.
// fm == NULL: the force meshes were not taped; each step recomputes its own from xk[s] (paint + the forward Fourier
// passes: one extra pm_forces per step instead of 16 bytes per cell and step of tape -- the trade diffrax's
// checkpointed adjoint makes in the reference, nbody.py:999).
int nbody_steps_vjp(Engine* E, stream_t st, float* posbar, float* velbar, int64_t np, int n_steps, const float* alpha,
                    const float* beta, const float* drift_pre, const float* drift_post, int order, int paint_deconv,
                    int lap_fd, int grad_fd, const float* xk, const float* vk, const float* fm, const float* v0,
                    double* coefbar) {
  if (n_steps <= 0) return 0;
  if (!xk) {
    set_error("nbody_steps_vjp: the tape xk from the forward pass is required");
    return MCPM_EINVAL;
  }
  if (coefbar && (!vk || !v0)) {
    set_error("nbody_steps_vjp: coefficient cotangents need vk and v0");
    return MCPM_EINVAL;
  }
  const int64_t P3 = 3 * np;
  const bool cic = (order == 2);
  const bool side_zero = tune().side_zero && fm && cic && brick_path(E, order, nullptr);
  Frame f;
  const Frame* fr = E->frame(f);
  auto dcomb_of = [&](int s) { return drift_post[s] + (s == n_steps - 1 ? 0.0f : drift_pre[s + 1]); };
  if (coefbar) {  // both half drifts that were fused into dcomb(n-1) see the xbar the sweep starts from
    const float* v1 = vk + (int64_t)(n_steps - 1) * P3;
    TRY(dot_accum(st, posbar, v1, P3, 1.0, coefbar + 4 * (n_steps - 1) + 3));
  }
  TRY(axpy3(st, velbar, posbar, dcomb_of(n_steps - 1), P3, velbar));
  for (int s = n_steps - 1; s >= 0; --s) {
    const bool last = (s == n_steps - 1);
    const float* x1 = xk + (int64_t)s * P3;
    const float* v1 = vk ? vk + (int64_t)s * P3 : nullptr;
    const float* vprev = s == 0 ? v0 : (vk ? vk + (int64_t)(s - 1) * P3 : nullptr);
    const float dnext = s > 0 ? dcomb_of(s - 1) : drift_pre[0];
    const float* slot;
    if (fm) {
      slot = fm + (int64_t)s * 4 * E->N;
    } else {  // recompute this step's force meshes; r(3..6) holds them (float4 for CIC) until the gather has read them
      float* planar = cic ? E->r(0) : E->r(3);
      TRY(pm_forces(E, st, x1, np, order, paint_deconv, lap_fd, grad_fd, 0.0f, planar, nullptr, false));
      if (cic) TRY(interleave3(st, planar, E->r(3), E->N));
      slot = E->r(3);
    }
    // phibar = paint(beta * vbar): three planar meshes.  With recomputed forces the scratch r(3..6) is taken, so the
    // scatter goes to r(0..2) and rhobar to the tail of the complex scratch.
    float* m3 = fm ? E->r(4) : E->r(0);
    float* rhobar = fm ? E->r(3) : reinterpret_cast<float*>(E->c(Engine::kC - 1));
    if (cic) {
      int handled = 0;
#ifndef MCPM_HOSTEMU
      if (E->lat.px > 0) {  // brick-tiled: accumulates in shared memory, flushes planar meshes directly
        if (!(side_zero && !last))  // cleared on the side by read_grad4v of the step before (in sweep order)
          TRY(rt_memset(m3, 0, sizeof(float) * 3 * (size_t)E->N, st) ? MCPM_ECUDA : 0);
        handled = brick_paint3_cic(st, E->lat, x1, velbar, beta[s], np, E->nx, E->ny, E->nz, m3, fr);
        if (handled < 0) return MCPM_ECUDA;
      }
#endif
      if (!handled) TRY(paint3(st, x1, velbar, beta[s], nullptr, 0.0f, np, E->nx, E->ny, E->nz, order, m3, 0, fr));
    } else {
      TRY(paint3(st, x1, velbar, beta[s], nullptr, 0.0f, np, E->nx, E->ny, E->nz, order, m3, 0, fr));
    }
    TRY(density_cotangent(E, st, m3, lap_fd, grad_fd, 0.0f, paint_deconv ? order : 0, rhobar));
    if (coefbar) {
      TRY(dot_accum(st, velbar, vprev, P3, 1.0, coefbar + 4 * s + 0));
      if (beta[s] != 0.0f) {  // betabar = <vbar, F>, F = (v1 - alpha v0) / beta
        TRY(dot_accum(st, velbar, v1, P3, 1.0 / beta[s], coefbar + 4 * s + 1));
        TRY(dot_accum(st, velbar, vprev, P3, -(double)alpha[s] / beta[s], coefbar + 4 * s + 1));
      }
    }
    if (cic) {
      // xbar += dread(x1; beta*vbar . F + rhobar), then vbar = alpha vbar + dnext xbar: one gather
      TRY(read_grad4v(st, x1, slot, rhobar, velbar, beta[s], 1, alpha[s], np, E->nx, E->ny, E->nz, posbar, 1,
                      side_zero && s > 0 ? E->r(4) : nullptr, 3 * E->N, fr, coefbar ? 0.0f : dnext));
    } else {
      const float* ms[4] = {slot, slot + E->N, slot + 2 * E->N, rhobar};
      TRY(read_grad(st, x1, ms, 4, velbar, 3, beta[s], nullptr, np, E->nx, E->ny, E->nz, order, nullptr, 0.0f, posbar,
                    1, 0.0f, fr));
      TRY(axpby(st, velbar, alpha[s], posbar, coefbar ? 0.0f : dnext, 0.0f, P3, velbar));
    }
    if (coefbar) {  // the drift cotangents need xbar before it is folded into vbar: take the unfused order
      if (s > 0) {
        const float* vprev1 = vk + (int64_t)(s - 1) * P3;
        TRY(dot_accum(st, posbar, vprev1, P3, 1.0, coefbar + 4 * (s - 1) + 3));
        TRY(dot_accum(st, posbar, vprev1, P3, 1.0, coefbar + 4 * s + 2));
      } else {
        TRY(dot_accum(st, posbar, v0, P3, 1.0, coefbar + 2));
      }
      TRY(axpy3(st, velbar, posbar, dnext, P3, velbar));
    }
  }
  return 0;
}

// nufft at the paint shape (nbody.py:569-574): m interlaced paints -> batched R2C -> one combine pass
// kb_kcut > 0: Kaiser-Bessel window (paint and its deconvolution, nbody.py:321-322, 383-384) instead of `rectangular`.
int nufft(Engine* E, stream_t st, const float* pos, const float* weights, float wscalar, int64_t np,
          const float* scale, int paint_order, int interlace_order, int paint_deconv, cfloat* out_k, float kb_kcut,
          const ObsShift* obs) {
  const int m = interlace_order;
  if (m < 1 || m > Engine::kR - 1) {
    set_error("nufft: interlace_order must be in 1..6");
    return MCPM_EINVAL;
  }
  float jac = scale ? scale[0] * scale[1] * scale[2] : 1.0f;
  const bool kb = kb_kcut > 0.0f;
  Frame f;
  const Frame* fr = E->frame(f);
  for (int i = 0; i < m; ++i) {
    if (kb)
      TRY(paint(st, pos, weights, wscalar, np, E->nx, E->ny, E->nz, paint_order, scale, (float)i / (float)m, E->r(i), 0,
                kb_kcut, fr, obs));
    else
      TRY(paint_fresh(E, st, pos, weights, wscalar, np, paint_order, scale, (float)i / (float)m, E->r(i), false, obs));
  }
  TRY(fft_r2c(E->fft, st, E->r(0), E->c(0), m));
  TRY(interlace_combine(st, E->c(0), out_k, m, E->nx, E->ny, E->nz, jac, paint_deconv && !kb ? paint_order : 0));
  if (kb && paint_deconv) TRY(deconv(st, out_k, out_k, E->nx, E->ny, E->nz, paint_order, kb_kcut));
  return 0;
}

int nufft_vjp(Engine* E, stream_t st, const float* pos, const float* weights, float wscalar, int64_t np,
              const float* scale, int paint_order, int interlace_order, int paint_deconv, const cfloat* outbar_k,
              float* posbar, float* weightsbar, float kb_kcut, const ObsShift* obs, float* velbar) {
  const int m = interlace_order;
  if (m < 1 || m > Engine::kR - 1) {
    set_error("nufft: interlace_order must be in 1..6");
    return MCPM_EINVAL;
  }
  float jac = scale ? scale[0] * scale[1] * scale[2] : 1.0f;
  const bool kb = kb_kcut > 0.0f;
  if (kb && paint_deconv) {  // the deconvolution is a real diagonal factor: its transpose is itself
    TRY(deconv(st, outbar_k, E->c(Engine::kC - 1), E->nx, E->ny, E->nz, paint_order, kb_kcut));
    outbar_k = E->c(Engine::kC - 1);
  }
  // out_i = conj(kernel_i) * outbar / w'  (raw C2R supplies the remaining factor N * 1/N)
  TRY(interlace_combine_T(st, outbar_k, E->c(0), m, E->nx, E->ny, E->nz, jac, paint_deconv && !kb ? paint_order : 0,
                          1.0f));
  // conj(phase_i) breaks Hermitian symmetry on the Nyquist planes exactly as the forward phase does: project (fourier.cu)
  TRY(hermitian_project(st, E->c(0), E->nx, E->ny, E->nz, m));
  TRY(fft_c2r(E->fft, st, E->c(0), E->r(0), m));
  Frame f;
  const Frame* fr = E->frame(f);
  // the m interlaced transposes in one gather over the particles (meshes r(0) .. r(m - 1) are contiguous)
  TRY(paint_vjp(st, pos, weights, wscalar, E->r(0), np, E->nx, E->ny, E->nz, paint_order, scale, 0.0f, posbar, weightsbar,
                0, kb_kcut, fr, obs, velbar, m, E->N));
  return 0;
}

}  // namespace mcpm
