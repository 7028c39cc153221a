// window.h -- 1-D mass-assignment windows of montecosmo's `rectangular` (nbody.py:220-246) and their derivatives,
// plus the base-index rule of `paint` / `read` (nbody.py:375-376, 388).
#pragma once
#include "rt.h"

namespace mcpm {

// Python-style modulo (nbody.py:372-373): result in [0, n).
MCPM_HD int wrap_index(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}

// Same result; the common range [-n, 2n) costs two compares instead of an integer division.
MCPM_HD int wrap_fast(int i, int n) {
  if ((unsigned)i < (unsigned)n) return i;
  if (i < 0 && i >= -n) return i + n;
  if (i >= n && i < 2 * n) return i - n;
  return wrap_index(i, n);
}

// W(s), s = |u| >= 0, and dW/ds.
template <int ORDER>
MCPM_HD float window(float s) {
  if (ORDER == 1) return 1.0f;
  if (ORDER == 2) return 1.0f - s;
  if (ORDER == 3) return s <= 0.5f ? 0.75f - s * s : 0.5f * (1.5f - s) * (1.5f - s);
  // ORDER == 4
  return s <= 1.0f ? (4.0f - 6.0f * s * s + 3.0f * s * s * s) * (1.0f / 6.0f)
                   : (2.0f - s) * (2.0f - s) * (2.0f - s) * (1.0f / 6.0f);
}

template <int ORDER>
MCPM_HD float dwindow_ds(float s) {
  if (ORDER == 1) return 0.0f;
  if (ORDER == 2) return -1.0f;
  if (ORDER == 3) return s <= 0.5f ? -2.0f * s : -(1.5f - s);
  return s <= 1.0f ? (-2.0f * s + 1.5f * s * s) : -0.5f * (2.0f - s) * (2.0f - s);
}

// For transformed coordinate x: first neighbour index `first` (unwrapped) and the ORDER window values.
//   id0 = floor(x) for even ORDER, round-half-to-even for odd ORDER (jnp.round);  shifts = arange(ORDER) - (ORDER-1)//2.
//   u_t = (id0 + shift_t) - x is formed as shift_t - (x - id0): both subtractions are exact in float32.
template <int ORDER>
MCPM_HD void window_weights(float x, int& first, float* w) {
  float f0 = (ORDER & 1) ? rintf(x) : floorf(x);
  float fr = x - f0;
  const int s0 = -((ORDER - 1) / 2);
  first = (int)f0 + s0;
#pragma unroll
  for (int t = 0; t < ORDER; ++t) w[t] = window<ORDER>(fabsf((float)(s0 + t) - fr));
}

// Same plus d/dx of each weight:  d/dx W(|u|) = -sign(u) W'(|u|), u = idx - x, sign(0) = 0 (autodiff of jnp.abs).
template <int ORDER>
MCPM_HD void window_weights_grad(float x, int& first, float* w, float* dw) {
  float f0 = (ORDER & 1) ? rintf(x) : floorf(x);
  float fr = x - f0;
  const int s0 = -((ORDER - 1) / 2);
  first = (int)f0 + s0;
#pragma unroll
  for (int t = 0; t < ORDER; ++t) {
    float u = (float)(s0 + t) - fr;
    float s = fabsf(u);
    w[t] = window<ORDER>(s);
    float sg = u > 0.0f ? 1.0f : (u < 0.0f ? -1.0f : 0.0f);
    dw[t] = -sg * dwindow_ds<ORDER>(s);
  }
}

// ---- window families as functors, so that one kernel body serves both -------------------------------------------
struct RectWin {  // `rectangular` (nbody.py:220-246)
  template <int ORDER>
  MCPM_HD void weights(float x, int& first, float* w) const {
    window_weights<ORDER>(x, first, w);
  }
  template <int ORDER>
  MCPM_HD void weights_grad(float x, int& first, float* w, float* dw) const {
    window_weights_grad<ORDER>(x, first, w, dw);
  }
};

// Modified Bessel functions I0(z) and I1(z)/z, z >= 0 (CUDA's cyl_bessel_i*f on the device, libstdc++'s on the host).
MCPM_HD float bessel_i0(float z) {
#if defined(__CUDA_ARCH__)
  return cyl_bessel_i0f(z);
#else
  return std::cyl_bessel_i(0.0f, z);
#endif
}
MCPM_HD float bessel_i1_over_z(float z) {
  if (z < 1e-3f) return 0.5f + 0.0625f * z * z;
#if defined(__CUDA_ARCH__)
  return cyl_bessel_i1f(z) / z;
#else
  return std::cyl_bessel_i(1.0f, z) / z;
#endif
}

// `kaiser_bessel` (nbody.py:280-290): W(s) = I0(kc sqrt(1 - s'^2)) / (order sinh(kc) / kc), s' = 2 s / order,
// kc = kcut * order / 2, kcut = optim_kcut(oversamp) (nbody.py:357-363).  Same base index and shifts as the rectangular
// family (nbody.py:375-376); the window does not vanish at the edge of its support (s' = 1 gives I0(0) / norm).
// dW/ds = -inv_norm * I1(z)/z * kc^2 * s' * (2/order), finite at s' = 1 (where autodiff of sqrt gives inf).
struct KbWin {
  float kc, inv_norm, s_scale;  // kcut*order/2 ; kc / (order sinh kc) ; 2 / order
  template <int ORDER>
  MCPM_HD void weights(float x, int& first, float* w) const {
    float f0 = (ORDER & 1) ? rintf(x) : floorf(x);
    float fr = x - f0;
    const int s0 = -((ORDER - 1) / 2);
    first = (int)f0 + s0;
#pragma unroll
    for (int t = 0; t < ORDER; ++t) {
      float sp = fabsf((float)(s0 + t) - fr) * s_scale;
      w[t] = bessel_i0(kc * sqrtf(fmaxf(1.0f - sp * sp, 0.0f))) * inv_norm;
    }
  }
  template <int ORDER>
  MCPM_HD void weights_grad(float x, int& first, float* w, float* dw) const {
    float f0 = (ORDER & 1) ? rintf(x) : floorf(x);
    float fr = x - f0;
    const int s0 = -((ORDER - 1) / 2);
    first = (int)f0 + s0;
#pragma unroll
    for (int t = 0; t < ORDER; ++t) {
      float u = (float)(s0 + t) - fr;
      float sp = fabsf(u) * s_scale;
      float z = kc * sqrtf(fmaxf(1.0f - sp * sp, 0.0f));
      w[t] = bessel_i0(z) * inv_norm;
      float sg = u > 0.0f ? 1.0f : (u < 0.0f ? -1.0f : 0.0f);
      // d/dx W(|u|) = -sign(u) dW/ds
      dw[t] = sg * inv_norm * bessel_i1_over_z(z) * kc * kc * sp * s_scale;
    }
  }
};

static inline KbWin make_kbwin(int order, float kcut) {
  double kc = 0.5 * (double)kcut * order;
  KbWin w;
  w.kc = (float)kc;
  w.inv_norm = (float)(kc / (order * std::sinh(kc)));
  w.s_scale = 2.0f / (float)order;
  return w;
}

}  // namespace mcpm
