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

}  // namespace mcpm
