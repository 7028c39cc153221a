// kspace.h -- wavevectors and Fourier-space kernel factors recomputed from element indices, shared by the streaming
// passes (fourier.cu) and the fused x-transform passes (xfft.cu).  Reference: rfftk (nbody.py:50-77), invlaplace_hat
// (109-133), gradient_hat (136-163), gaussian_hat (166-188), rectangular_hat (249-277).
#pragma once
#include "engine.h"

namespace mcpm {

// Local block of the half spectrum [nx, ny_loc, nz/2+1] holding global ky rows y0 .. y0 + ny_loc - 1 (slab-decomposed
// runs keep k-space split along ky; single-GPU: ny_loc = ny, y0 = 0).
struct KGrid {
  int nx, ny, nz, nzc;
  int ny_loc, y0;
  float tx, ty, tz;  // 2*pi / n
  int lap_fd, grad_fd;
};

static KGrid make_kgrid(int nx, int ny, int nz, int lap_fd = 0, int grad_fd = 0, SlabK sk = SlabK()) {
  KGrid g;
  g.nx = nx;
  g.ny = ny;
  g.nz = nz;
  g.nzc = nz / 2 + 1;
  g.ny_loc = sk.ny_loc > 0 ? sk.ny_loc : ny;
  g.y0 = sk.ny_loc > 0 ? sk.y0 : 0;
  g.tx = (float)(6.283185307179586476925 / nx);
  g.ty = (float)(6.283185307179586476925 / ny);
  g.tz = (float)(6.283185307179586476925 / nz);
  g.lap_fd = lap_fd;
  g.grad_fd = grad_fd;
  return g;
}

// numpy.fft.fftfreq index -> signed integer frequency (Nyquist negative); rfftfreq keeps it positive.
MCPM_HD int signed_freq(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

struct KVec {
  float kx, ky, kz;
  // Hermitian consistency.  jnp.fft.irfftn of a spectrum that violates Hermitian symmetry keeps only its Hermitian
  // projection on the self-conjugate planes kz = 0 and kz = Nyquist; cuFFT's C2R on such input is algorithm dependent
  // (its batched 2-D C2R differed by 2.5% at batch 16).  For kernel x Hermitian field the projection is zero wherever the
  // kernel is odd under k -> -k, i.e. carries an odd number of Nyquist-valued gradient factors (rfftk puts -pi on x, y
  // and +pi on z, nbody.py:72-76, and -k maps a Nyquist index onto itself).  sc: element lies on a self-conjugate plane;
  // nq*: that component is at its Nyquist index.
  bool sc, nqx, nqy, nqz;
};

MCPM_HD KVec kvec_at(const KGrid& g, int64_t e, int& l) {
  l = (int)(e % g.nzc);
  int64_t r = e / g.nzc;
  int j = (int)(r % g.ny_loc) + g.y0;
  int i = (int)(r / g.ny_loc);
  KVec k;
  k.kx = g.tx * (float)signed_freq(i, g.nx);
  k.ky = g.ty * (float)signed_freq(j, g.ny);
  k.kz = g.tz * (float)l;
  k.nqx = 2 * i == g.nx;
  k.nqy = 2 * j == g.ny;
  k.nqz = 2 * l == g.nz;
  k.sc = l == 0 || k.nqz;
  return k;
}

// invlaplace_hat: -1/kk with 0 at kk == 0 (safe_div, utils.py:21-29)
MCPM_HD float lap_term(float k, int fd) {
  if (fd == 2) return (cosf(k) - 1.0f) * 2.0f;
  if (fd == 4) return (cosf(2.0f * k) - 16.0f * cosf(k) + 15.0f) * (1.0f / 6.0f);
  return k * k;
}
MCPM_HD float invlaplace(const KGrid& g, const KVec& k) {
  float kk = lap_term(k.kx, g.lap_fd) + lap_term(k.ky, g.lap_fd) + lap_term(k.kz, g.lap_fd);
  return kk == 0.0f ? 0.0f : -1.0f / kk;
}
// gradient_hat / i
MCPM_HD float grad_term(float k, int fd) {
  if (fd == 2) return sinf(k);
  if (fd == 4) return (8.0f * sinf(k) - sinf(2.0f * k)) * (1.0f / 6.0f);
  return k;
}
// the three gradient factors, with the Hermitian projection applied when the result feeds a C2R (see KVec)
MCPM_HD void grad_terms(const KGrid& g, const KVec& k, float& gx, float& gy, float& gz, bool project = true) {
  gx = grad_term(k.kx, g.grad_fd);
  gy = grad_term(k.ky, g.grad_fd);
  gz = grad_term(k.kz, g.grad_fd);
  if (project && k.sc) {
    if (k.nqx) gx = 0.0f;
    if (k.nqy) gy = 0.0f;
    if (k.nqz) gz = 0.0f;
  }
}
// sinc(k / 2pi) = sin(k/2) / (k/2)
MCPM_HD float sinc_half(float k) {
  float x = 0.5f * k;
  return x == 0.0f ? 1.0f : sinf(x) / x;
}
MCPM_HD float powi(float x, int p) {
  float r = 1.0f;
  for (int t = 0; t < p; ++t) r *= x;
  return r;
}
MCPM_HD float window_hat(const KVec& k, int order) {
  return powi(sinc_half(k.kx) * sinc_half(k.ky) * sinc_half(k.kz), order);
}
// kaiser_bessel_hat (nbody.py:293-312), one factor: with k' = k*order/2, kc = kcut*order/2, d = sqrt|kc^2 - k'^2|,
// sinh(d)/d inside the cutoff and sin(d)/d beyond it, over sinh(kc)/kc; the removable 0/0 at d = 0 is taken as 1.
struct KbHat {
  float half_order, kc, inv_norm;  // order/2 ; kcut*order/2 ; kc / sinh(kc)
};
static inline KbHat make_kbhat(int order, float kcut) {
  double kc = 0.5 * (double)kcut * order;
  KbHat h;
  h.half_order = 0.5f * (float)order;
  h.kc = (float)kc;
  h.inv_norm = (float)(kc / std::sinh(kc));
  return h;
}
MCPM_HD float kb_hat_1d(const KbHat& h, float k) {
  float kp = k * h.half_order;
  float d2 = h.kc * h.kc - kp * kp;
  float d = sqrtf(fabsf(d2));
  float v = d == 0.0f ? 1.0f : (fabsf(kp) <= h.kc ? sinhf(d) / d : sinf(d) / d);
  return v * h.inv_norm;
}
MCPM_HD float kb_window_hat(const KbHat& h, const KVec& k) {
  return kb_hat_1d(h, k.kx) * kb_hat_1d(h, k.ky) * kb_hat_1d(h, k.kz);
}
// weight of a half-spectrum element in the real inner product: 1 on the self-conjugate planes kz = 0, Nyquist; else 2
MCPM_HD float half_weight(int l, int nz) { return (l == 0 || 2 * l == nz) ? 1.0f : 2.0f; }

// common real factor of the force kernel: invlaplace * [gaussian] * [1 / rectangular_hat^2]
MCPM_HD float force_scalar(const KGrid& g, const KVec& k, float rcut2_half, int deconv_order) {
  float c = invlaplace(g, k);
  if (rcut2_half > 0.0f) c *= expf(-(k.kx * k.kx + k.ky * k.ky + k.kz * k.kz) * rcut2_half);
  if (deconv_order > 0) {
    float w = window_hat(k, deconv_order);
    c /= (w * w);
  }
  return c;
}

static float rcut2_half_of(float kcut) {
  if (!(kcut > 0.0f) || std::isinf(kcut)) return 0.0f;
  double rcut = 6.283185307179586476925 / kcut;
  return (float)(0.5 * rcut * rcut);
}

static int check_dims(int nx, int ny, int nz) {
  if (nx <= 0 || ny <= 0 || nz <= 0 || (nz & 1)) {
    set_error("mesh dimensions must be positive and nz even");
    return MCPM_EINVAL;
  }
  return 0;
}
static int check_fd(int fd) {
  if (fd != MCPM_FD_INF && fd != MCPM_FD_2 && fd != MCPM_FD_4) {
    set_error("Only orders 2, 4, and inf are supported.");
    return MCPM_EINVAL;
  }
  return 0;
}

}  // namespace mcpm
