// bias.cu -- the fused passes of the Lagrangian bias expansion (montecosmo/bricks.py:327-452).
//
// The reference builds the weights from ten inverse FFTs of Fourier multiples of the linear field (delta, five
// tidal-shear components, laplacian, three gradients; twelve with primordial non-Gaussianity), the two shear invariants,
// a read of every field at the particles and a polynomial with two particle means in it.  Round 1 composed that from
// engine transforms with ~60 eager elementwise passes between them.  Here each stage is one streaming pass:
//   bias_spectra      1 spectrum -> 10 | 12 spectra (every multiplier recomputed from the element index, kspace.h)
//   shear_invariants  5 real meshes -> s^2, s^3 (traceless condition used for the sixth component)
//   bias_moments      sum of delta_pos^2 and phi_pos delta_pos over the particles (the two means), float64
//   bias_weights      the polynomial and the velocity term from the K values read at each particle
// each with its transpose (bias_weights_vjp also returns the cotangents of the 13 coefficients and of the growth
// factor, which NUTS samples).  The transforms between them are the engine's C2R / R2C, the reads its gather.
// Bandwidth bound; bytes per element are in DESIGN.md section 3.
#include "engine.h"
#include "kspace.h"

namespace mcpm {

namespace {
struct BiasK {
  float kx, ky, kz, k2, ik2;
  float gx, gy, gz;     // gradient factors with the Hermitian projection of a C2R input (kspace.h: KVec)
  float mxy, mxz, myz;  // k_i k_j / k^2, projected likewise
};
MCPM_HD BiasK bias_k(const KGrid& g, int64_t e, float cx, float cy, float cz, bool project) {
  int l;
  KVec k = kvec_at(g, e, l);
  k.sc = k.sc && project;
  BiasK b;
  b.kx = k.kx * cx;  // rad / cell -> h / Mpc
  b.ky = k.ky * cy;
  b.kz = k.kz * cz;
  b.k2 = b.kx * b.kx + b.ky * b.ky + b.kz * b.kz;
  b.ik2 = b.k2 == 0.0f ? 0.0f : 1.0f / b.k2;
  b.gx = (k.sc && k.nqx) ? 0.0f : b.kx;
  b.gy = (k.sc && k.nqy) ? 0.0f : b.ky;
  b.gz = (k.sc && k.nqz) ? 0.0f : b.kz;
  b.mxy = (k.sc && k.nqx != k.nqy) ? 0.0f : b.kx * b.ky * b.ik2;
  b.mxz = (k.sc && k.nqx != k.nqz) ? 0.0f : b.kx * b.kz * b.ik2;
  b.myz = (k.sc && k.nqy != k.nqz) ? 0.0f : b.ky * b.kz * b.ik2;
  return b;
}
}  // namespace

// out[m] for m = 0..9 (png: ..11): s00, s11, s01, s02, s12, delta, -k^2 delta, i kx delta, i ky delta, i kz delta,
// (phi = delta * inv_transfer, -k^2 phi) -- the five shear components first, so that their transforms land next to each
// other for shear_invariants and everything the particles read follows in one block.  bricks.py:342-348, 361-372, 397, 411-413, 436, 444-446.
int bias_spectra(stream_t st, const cfloat* dk, int nx, int ny, int nz, float cx, float cy, float cz,
                 const float* inv_transfer, cfloat* out) {
  if (int e = check_dims(nx, ny, nz)) return e;
  KGrid g = make_kgrid(nx, ny, nz);
  const int64_t nc = (int64_t)nx * ny * g.nzc;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    const BiasK b = bias_k(g, e, cx, cy, cz, true);
    const cfloat d = dk[e];
    const float third = 1.0f / 3.0f;
    const float m1 = b.kx * b.kx * b.ik2 - third, m2 = b.ky * b.ky * b.ik2 - third;
    out[e] = cfloat{m1 * d.re, m1 * d.im};
    out[nc + e] = cfloat{m2 * d.re, m2 * d.im};
    out[2 * nc + e] = cfloat{b.mxy * d.re, b.mxy * d.im};
    out[3 * nc + e] = cfloat{b.mxz * d.re, b.mxz * d.im};
    out[4 * nc + e] = cfloat{b.myz * d.re, b.myz * d.im};
    out[5 * nc + e] = d;
    out[6 * nc + e] = cfloat{-b.k2 * d.re, -b.k2 * d.im};
    out[7 * nc + e] = cfloat{-b.gx * d.im, b.gx * d.re};
    out[8 * nc + e] = cfloat{-b.gy * d.im, b.gy * d.re};
    out[9 * nc + e] = cfloat{-b.gz * d.im, b.gz * d.re};
    if (inv_transfer) {
      const float t = inv_transfer[e];
      out[10 * nc + e] = cfloat{t * d.re, t * d.im};
      out[11 * nc + e] = cfloat{-b.k2 * t * d.re, -b.k2 * t * d.im};
    }
  });
  return rt_check("bias_spectra");
}

// transpose: dkbar (+)= sum_m conj(multiplier_m) outbar[m]   (cotangents as dL/dRe + i dL/dIm)
int bias_spectra_T(stream_t st, const cfloat* outbar, int nx, int ny, int nz, float cx, float cy, float cz,
                   const float* inv_transfer, cfloat* dkbar, int accumulate) {
  if (int e = check_dims(nx, ny, nz)) return e;
  KGrid g = make_kgrid(nx, ny, nz);
  const int64_t nc = (int64_t)nx * ny * g.nzc;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    // unprojected multipliers: the cotangent of a free complex array is what autodiff returns (the transposed irfftn in
    // front of this pass -- R2C + Hermitian weights -- already carries the projection the forward C2R applies)
    const BiasK b = bias_k(g, e, cx, cy, cz, false);
    const float third = 1.0f / 3.0f;
    const float m[7] = {b.kx * b.kx * b.ik2 - third, b.ky * b.ky * b.ik2 - third, b.mxy, b.mxz, b.myz, 1.0f, -b.k2};
    float re = 0.0f, im = 0.0f;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const cfloat v = outbar[t * nc + e];
      re += m[t] * v.re;
      im += m[t] * v.im;
    }
    const float gr[3] = {b.gx, b.gy, b.gz};
#pragma unroll
    for (int t = 0; t < 3; ++t) {  // conj(i g) v = g (v.im - i v.re)
      const cfloat v = outbar[(7 + t) * nc + e];
      re += gr[t] * v.im;
      im -= gr[t] * v.re;
    }
    if (inv_transfer) {
      const float t = inv_transfer[e];
      const cfloat v = outbar[10 * nc + e], w = outbar[11 * nc + e];
      re += t * (v.re - b.k2 * w.re);
      im += t * (v.im - b.k2 * w.im);
    }
    cfloat o = cfloat{re, im};
    if (accumulate) {
      o.re += dkbar[e].re;
      o.im += dkbar[e].im;
    }
    dkbar[e] = o;
  });
  return rt_check("bias_spectra_T");
}

// s5 = [s00, s11, s01, s02, s12][n] -> out = [s^2, s^3][n]  (bricks.py:373-395; s22 = -(s00 + s11))
int shear_invariants(stream_t st, const float* s5, int64_t n, float* out2) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t i) {
    const float a = s5[i], b = s5[n + i], d = s5[2 * n + i], e = s5[3 * n + i], f = s5[4 * n + i];
    const float c = -(a + b);
    out2[i] = a * a + b * b + c * c + 2.0f * (d * d + e * e + f * f);
    out2[n + i] = 3.0f * (a * (b * c - f * f) - d * (d * c - e * f) + e * (d * f - b * e));
  });
  return rt_check("shear_invariants");
}

int shear_invariants_vjp(stream_t st, const float* s5, const float* out2bar, int64_t n, float* s5bar) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t i) {
    const float a = s5[i], b = s5[n + i], d = s5[2 * n + i], e = s5[3 * n + i], f = s5[4 * n + i];
    const float c = -(a + b);
    const float q = out2bar[i], t = 3.0f * out2bar[n + i];
    const float dc = a * b - d * d;  // d(s^3 / 3) / dc at fixed a, b
    s5bar[i] = q * 2.0f * (a - c) + t * ((b * c - f * f) - dc);
    s5bar[n + i] = q * 2.0f * (b - c) + t * ((a * c - e * e) - dc);
    s5bar[2 * n + i] = q * 4.0f * d + t * 2.0f * (e * f - c * d);
    s5bar[3 * n + i] = q * 4.0f * e + t * 2.0f * (d * f - b * e);
    s5bar[4 * n + i] = q * 4.0f * f + t * 2.0f * (d * e - a * f);
  });
  return rt_check("shear_invariants_vjp");
}

// ---- particle side.  vals = [np, K] read at the particles: delta, s^2, s^3, lap delta, grad delta (3), [phi, lap phi]
// (K = 7 | 9); growth: one value (garr = NULL) or one per particle (light cone).
namespace {
// Sum NV per-element values over [0, n) in float64 into out[0..NV): one warp per chunk of 4096 elements.
template <int NV, class F>
int reduce_sums(stream_t st, int64_t n, double* out, const char* what, F f) {
#ifdef MCPM_HOSTEMU
  double acc[NV];
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double t[NV];
    f(i, t);
    for (int v = 0; v < NV; ++v) acc[v] += t[v];
  }
  for (int v = 0; v < NV; ++v) out[v] += acc[v];
  (void)st;
  (void)what;
  return 0;
#else
  const int64_t chunk = 4096, nchunks = (n + chunk - 1) / chunk;
  launch_1d(st, nchunks * 32, [=] MCPM_LAMBDA(int64_t t) {
    const int64_t c = t >> 5;
    const int lane = (int)(t & 31);
    const int64_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.0;
    for (int64_t i = lo + lane; i < hi; i += 32) {
      double tv[NV];
      f(i, tv);
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] += tv[v];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double s = acc[v];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0 && s != 0.0) atomic_add(out + v, s);
    }
  });
  return rt_check(what);
#endif
}
}  // namespace

// mom[0] += sum (delta g)^2, mom[1] += sum phi delta g   (the means of bricks.py:354, 421 times np)
int bias_moments(stream_t st, const float* vals, int K, float gs, const float* garr, int64_t np, double* mom) {
  return reduce_sums<2>(st, np, mom, "bias_moments", [=] MCPM_LAMBDA(int64_t p, double* t) {
    const float g = garr ? garr[p] : gs;
    const float D = vals[p * K] * g;
    t[0] = (double)D * D;
    t[1] = K >= 9 ? (double)vals[p * K + 7] * D : 0.0;
  });
}

// weights[p], dvel[p, 3] (bricks.py:350-449)
int bias_weights(stream_t st, const float* vals, int K, float gs, const float* garr, BiasCoef c, const double* mom,
                 int64_t np, float* weights, float* dvel) {
  const double inv_np = 1.0 / (double)np;
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    const float sigma2 = (float)(mom[0] * inv_np), sigma_pd = (float)(mom[1] * inv_np);
    const float* v = vals + p * K;
    const float g = garr ? garr[p] : gs;
    const float D = v[0] * g, s2p = v[1] * g * g - (2.0f / 3.0f) * sigma2, T = v[2] * g * g * g, L = v[3] * g;
    const float D2 = D * D - sigma2;
    float w = 1.0f + c.b1 * D + c.b2 * D2 * 0.5f + c.bs2 * s2p + c.b3 * (D * D * D - 3.0f * sigma2 * D) * (1.0f / 6.0f) +
              c.bds2 * D * s2p + c.bs3 * T + c.bn2 * L;
    if (K >= 9) {
      const float P = v[7], LP = v[8];
      w += c.fbp * P + c.fbpd * (P * D - sigma_pd) + c.fbpd2 * (P * D2 - 2.0f * sigma_pd * D) + c.fbps2 * P * s2p +
           c.fbn2p * LP;
    }
    weights[p] = w;
    if (dvel) {
      dvel[3 * p] = c.bnpar * v[4] * g;
      dvel[3 * p + 1] = c.bnpar * v[5] * g;
      dvel[3 * p + 2] = c.bnpar * v[6] * g;
    }
  });
  return rt_check("bias_weights");
}

namespace {
struct BiasPoint {  // per-particle terms shared by the two reverse passes
  float g, D, s2p, T, L, P, LP, D2;
  float dwdD, dwdQ, dwdP, dwds2, dwdspd;
};
MCPM_HD BiasPoint bias_point(const float* v, int K, float g, const BiasCoef& c, float sigma2, float sigma_pd) {
  BiasPoint b;
  b.g = g;
  b.D = v[0] * g;
  b.s2p = v[1] * g * g - (2.0f / 3.0f) * sigma2;
  b.T = v[2] * g * g * g;
  b.L = v[3] * g;
  b.P = K >= 9 ? v[7] : 0.0f;
  b.LP = K >= 9 ? v[8] : 0.0f;
  b.D2 = b.D * b.D - sigma2;
  b.dwdD = c.b1 + c.b2 * b.D + c.b3 * 0.5f * b.D2 + c.bds2 * b.s2p + c.fbpd * b.P + 2.0f * c.fbpd2 * (b.P * b.D - sigma_pd);
  b.dwdQ = c.bs2 + c.bds2 * b.D + c.fbps2 * b.P;
  b.dwdP = c.fbp + c.fbpd * b.D + c.fbpd2 * b.D2 + c.fbps2 * b.s2p;
  b.dwds2 = -0.5f * c.b2 - (2.0f / 3.0f) * b.dwdQ - 0.5f * c.b3 * b.D - c.fbpd2 * b.P;
  b.dwdspd = -c.fbpd - 2.0f * c.fbpd2 * b.D;
  return b;
}
}  // namespace

// Reverse of bias_moments + bias_weights.  wbar [np], dvelbar [np, 3] or NULL -> valsbar [np, K]; coefbar[14] (float64,
// accumulated): the 13 coefficients in the order of BiasCoef, then the growth factor when it is one scalar; gbar_arr [np]
// receives the per-particle growth cotangent when garr is given.
// msum: float64 scratch [2], zeroed by the caller (the sums that carry the cotangent of the two means).
int bias_weights_vjp(stream_t st, const float* vals, int K, float gs, const float* garr, BiasCoef c, const double* mom,
                     const float* wbar, const float* dvelbar, int64_t np, double* msum, float* valsbar, double* coefbar,
                     float* gbar_arr) {
  const double inv_np = 1.0 / (double)np;
  int rc = reduce_sums<2>(st, np, msum, "bias_weights_vjp (means)", [=] MCPM_LAMBDA(int64_t p, double* t) {
    const float sigma2 = (float)(mom[0] * inv_np), sigma_pd = (float)(mom[1] * inv_np);
    const BiasPoint b = bias_point(vals + p * K, K, garr ? garr[p] : gs, c, sigma2, sigma_pd);
    t[0] = (double)wbar[p] * b.dwds2;
    t[1] = (double)wbar[p] * b.dwdspd;
  });
  if (rc) return rc;
  return reduce_sums<14>(st, np, coefbar, "bias_weights_vjp", [=] MCPM_LAMBDA(int64_t p, double* t) {
    const float sigma2 = (float)(mom[0] * inv_np), sigma_pd = (float)(mom[1] * inv_np);
    const float S2 = (float)(msum[0] * inv_np), Spd = (float)(msum[1] * inv_np);
    const float* v = vals + p * K;
    const BiasPoint b = bias_point(v, K, garr ? garr[p] : gs, c, sigma2, sigma_pd);
    const float wb = wbar[p], g = b.g;
    const float Dbar = wb * b.dwdD + 2.0f * b.D * S2 + b.P * Spd;
    const float Qbar = wb * b.dwdQ;
    float* o = valsbar + p * K;
    o[0] = Dbar * g;
    o[1] = Qbar * g * g;
    o[2] = wb * c.bs3 * g * g * g;
    o[3] = wb * c.bn2 * g;
    float dv[3] = {0.0f, 0.0f, 0.0f};
    if (dvelbar) {
      dv[0] = dvelbar[3 * p];
      dv[1] = dvelbar[3 * p + 1];
      dv[2] = dvelbar[3 * p + 2];
    }
    o[4] = dv[0] * c.bnpar * g;
    o[5] = dv[1] * c.bnpar * g;
    o[6] = dv[2] * c.bnpar * g;
    const float dvG = dv[0] * v[4] + dv[1] * v[5] + dv[2] * v[6];
    if (K >= 9) {
      o[7] = wb * b.dwdP + b.D * Spd;
      o[8] = wb * c.fbn2p;
    }
    const float gb = Dbar * v[0] + Qbar * 2.0f * g * v[1] + wb * c.bs3 * 3.0f * g * g * v[2] + wb * c.bn2 * v[3] + c.bnpar * dvG;
    if (gbar_arr) gbar_arr[p] = gb;
    t[0] = (double)wb * b.D;
    t[1] = (double)wb * 0.5 * b.D2;
    t[2] = (double)wb * b.s2p;
    t[3] = (double)wb * (b.D * b.D * b.D - 3.0f * sigma2 * b.D) * (1.0 / 6.0);
    t[4] = (double)wb * b.D * b.s2p;
    t[5] = (double)wb * b.T;
    t[6] = (double)wb * b.L;
    t[7] = (double)dvG * g;
    t[8] = (double)wb * b.P;
    t[9] = (double)wb * (b.P * b.D - sigma_pd);
    t[10] = (double)wb * (b.P * b.D2 - 2.0f * sigma_pd * b.D);
    t[11] = (double)wb * b.P * b.s2p;
    t[12] = (double)wb * b.LP;
    t[13] = gbar_arr ? 0.0 : (double)gb;
  });
}

}  // namespace mcpm
