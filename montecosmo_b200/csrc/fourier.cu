// fourier.cu -- Fourier-space passes around cuFFT, one fused kernel per role.  All kernels recompute the wavevector
// from the element index (rfftk, nbody.py:50-77) instead of reading baked kernel arrays: every pass is a pure stream
// over the half spectrum [nx, ny, nz/2+1] of interleaved complex64, one thread per element, coalesced along kz.
// Reference kernels: invlaplace_hat (nbody.py:109-133), gradient_hat (136-163), gaussian_hat (166-188),
// rectangular_hat (249-277), deconv_paint (315-334), interlace phase (523-526), chreshape (utils.py:924-1013).
#include "engine.h"
#include "kspace.h"

namespace mcpm {

// out_j = -(i g_j) * c * delta   (nbody.py:597-603)
int force_spectra(stream_t st, const cfloat* dk, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd,
                  float kcut, int deconv_order, float norm, SlabK sk) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (int e = check_fd(lap_fd) ? check_fd(lap_fd) : check_fd(grad_fd)) return e;
  KGrid g = make_kgrid(nx, ny, nz, lap_fd, grad_fd, sk);
  const int64_t nc = (int64_t)nx * g.ny_loc * g.nzc;
  const float r2 = rcut2_half_of(kcut);
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float c = force_scalar(g, k, r2, deconv_order) * norm;
    cfloat d = dk[e];
    // (-i)(a + ib) = b - ia
    float re = d.im * c, im = -d.re * c;
    float gx, gy, gz;
    grad_terms(g, k, gx, gy, gz);
    out3[e] = cfloat{gx * re, gx * im};
    out3[nc + e] = cfloat{gy * re, gy * im};
    out3[2 * nc + e] = cfloat{gz * re, gz * im};
  });
  return rt_check("force_spectra");
}

// transpose pass: out (+)= [w'/N] * sum_j conj(-(i g_j) c) in_j = [w'/N] * c * i * sum_j g_j in_j
int force_spectra_T(stream_t st, const cfloat* in3, cfloat* out1, int nx, int ny, int nz, int lap_fd, int grad_fd,
                    float kcut, int deconv_order, int half_weights, int accumulate, float norm, SlabK sk) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (int e = check_fd(lap_fd) ? check_fd(lap_fd) : check_fd(grad_fd)) return e;
  KGrid g = make_kgrid(nx, ny, nz, lap_fd, grad_fd, sk);
  const int64_t nc = (int64_t)nx * g.ny_loc * g.nzc;
  const float r2 = rcut2_half_of(kcut);
  const float invn = (float)(1.0 / ((double)nx * ny * nz));
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float c = force_scalar(g, k, r2, deconv_order) * norm;
    if (half_weights) c *= half_weight(l, g.nz) * invn;
    // half_weights: the output is the cotangent of a free complex array (as autodiff returns it), not a C2R input
    float gx, gy, gz;
    grad_terms(g, k, gx, gy, gz, !half_weights);
    cfloat a = in3[e], b = in3[nc + e], d = in3[2 * nc + e];
    float sre = gx * a.re + gy * b.re + gz * d.re;
    float sim = gx * a.im + gy * b.im + gz * d.im;
    // i (sre + i sim) = -sim + i sre
    cfloat o = cfloat{-sim * c, sre * c};
    if (accumulate) {
      cfloat p = out1[e];
      o.re += p.re;
      o.im += p.im;
    }
    out1[e] = o;
  });
  return rt_check("force_spectra_T");
}

// out_ij = (i g_i)(i g_j) * invlaplace * delta = -g_i g_j G delta ; order 00, 11, 22, 01, 02, 12  (nbody.py:611-627)
int hessian_spectra(stream_t st, const cfloat* dk, cfloat* out6, int nx, int ny, int nz, int lap_fd, int grad_fd,
                    float norm, SlabK sk) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (int e = check_fd(lap_fd) ? check_fd(lap_fd) : check_fd(grad_fd)) return e;
  KGrid g = make_kgrid(nx, ny, nz, lap_fd, grad_fd, sk);
  const int64_t nc = (int64_t)nx * g.ny_loc * g.nzc;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float c = -invlaplace(g, k) * norm;
    float gx = grad_term(k.kx, g.grad_fd), gy = grad_term(k.ky, g.grad_fd), gz = grad_term(k.kz, g.grad_fd);
    // mixed terms are odd under k -> -k when exactly one of the two factors sits at its Nyquist index (see KVec)
    float mxy = (k.sc && k.nqx != k.nqy) ? 0.0f : gx * gy;
    float mxz = (k.sc && k.nqx != k.nqz) ? 0.0f : gx * gz;
    float myz = (k.sc && k.nqy != k.nqz) ? 0.0f : gy * gz;
    cfloat d = dk[e];
    float re = d.re * c, im = d.im * c;
    out6[e] = cfloat{gx * gx * re, gx * gx * im};
    out6[nc + e] = cfloat{gy * gy * re, gy * gy * im};
    out6[2 * nc + e] = cfloat{gz * gz * re, gz * gz * im};
    out6[3 * nc + e] = cfloat{mxy * re, mxy * im};
    out6[4 * nc + e] = cfloat{mxz * re, mxz * im};
    out6[5 * nc + e] = cfloat{myz * re, myz * im};
  });
  return rt_check("hessian_spectra");
}

int hessian_spectra_T(stream_t st, const cfloat* in6, cfloat* out1, int nx, int ny, int nz, int lap_fd, int grad_fd,
                      int half_weights, int accumulate, float norm, SlabK sk) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (int e = check_fd(lap_fd) ? check_fd(lap_fd) : check_fd(grad_fd)) return e;
  KGrid g = make_kgrid(nx, ny, nz, lap_fd, grad_fd, sk);
  const int64_t nc = (int64_t)nx * g.ny_loc * g.nzc;
  const float invn = (float)(1.0 / ((double)nx * ny * nz));
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float c = -invlaplace(g, k) * norm;
    if (half_weights) c *= half_weight(l, g.nz) * invn;
    float gx = grad_term(k.kx, g.grad_fd), gy = grad_term(k.ky, g.grad_fd), gz = grad_term(k.kz, g.grad_fd);
    const bool pr = k.sc && !half_weights;  // project only when the output feeds a C2R
    float w[6] = {gx * gx, gy * gy, gz * gz, (pr && k.nqx != k.nqy) ? 0.0f : gx * gy,
                  (pr && k.nqx != k.nqz) ? 0.0f : gx * gz, (pr && k.nqy != k.nqz) ? 0.0f : gy * gz};
    float re = 0.0f, im = 0.0f;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      cfloat v = in6[t * nc + e];
      re += w[t] * v.re;
      im += w[t] * v.im;
    }
    cfloat o = cfloat{re * c, im * c};
    if (accumulate) {
      cfloat p = out1[e];
      o.re += p.re;
      o.im += p.im;
    }
    out1[e] = o;
  });
  return rt_check("hessian_spectra_T");
}

// The y and z force components from the potential, on spectra that are already in x-space: buf3 = [3][xl][ny][nzc]
// holds (F^_x, Phi^) of xfuse's FORCE2 in components 0, 1; on return components 1, 2 hold F^_y = -(i g_y) Phi^ and
// F^_z = -(i g_z) Phi^ (gradient_hat along y / z, nbody.py:136-163, with the Hermitian projection of kspace.h: KVec --
// the factors depend on (ky, kz) only, so they commute with the x-transform).  transpose = 1: components (C^_x, C^_y,
// C^_z) -> component 1 = g_y C^_y + g_z C^_z, the input of FORCE2_T.  In place, one streaming pass over local planes.
int yz_gradients(stream_t st, cfloat* buf3, int xl, int ny, int nz, int grad_fd, int transpose) {
  if (int e = check_fd(grad_fd)) return e;
  if (xl <= 0 || ny <= 0 || nz <= 0 || (nz & 1)) {
    set_error("yz_gradients: bad shape");
    return MCPM_EINVAL;
  }
  const int nzc = nz / 2 + 1;
  const int64_t nc = (int64_t)xl * ny * nzc;
  const float ty = (float)(6.283185307179586476925 / ny), tz = (float)(6.283185307179586476925 / nz);
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    const int l = (int)(e % nzc), j = (int)((e / nzc) % ny);
    const bool nqy = 2 * j == ny, nqz = 2 * l == nz, sc = l == 0 || nqz;
    const float gy = (sc && nqy) ? 0.0f : grad_term(ty * (float)signed_freq(j, ny), grad_fd);
    const float gz = (sc && nqz) ? 0.0f : grad_term(tz * (float)l, grad_fd);
    if (!transpose) {
      const cfloat p = buf3[nc + e];
      buf3[nc + e] = cfloat{gy * p.im, -gy * p.re};
      buf3[2 * nc + e] = cfloat{gz * p.im, -gz * p.re};
    } else {
      const cfloat a = buf3[nc + e], b = buf3[2 * nc + e];
      buf3[nc + e] = cfloat{gy * a.re + gz * b.re, gy * a.im + gz * b.im};
    }
  });
  return rt_check("yz_gradients");
}

// d2 = h00 h11 + h00 h22 + h11 h22 - h01^2 - h02^2 - h12^2   (nbody.py:615-627)
int lpt2_source(stream_t st, const float* h6, float* d2, int64_t n) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t e) {
    float a = h6[e], b = h6[n + e], c = h6[2 * n + e], d = h6[3 * n + e], f = h6[4 * n + e], g = h6[5 * n + e];
    d2[e] = a * b + c * (a + b) - d * d - f * f - g * g;
  });
  return rt_check("lpt2_source");
}

int lpt2_source_vjp(stream_t st, const float* h6, const float* d2bar, float* hbar6, int64_t n) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t e) {
    float a = h6[e], b = h6[n + e], c = h6[2 * n + e], d = h6[3 * n + e], f = h6[4 * n + e], g = h6[5 * n + e];
    float t = d2bar[e];
    hbar6[e] = t * (b + c);
    hbar6[n + e] = t * (a + c);
    hbar6[2 * n + e] = t * (a + b);
    hbar6[3 * n + e] = -2.0f * t * d;
    hbar6[4 * n + e] = -2.0f * t * f;
    hbar6[5 * n + e] = -2.0f * t * g;
  });
  return rt_check("lpt2_source_vjp");
}

int deconv(stream_t st, const cfloat* in, cfloat* out, int nx, int ny, int nz, int order, float kb_kcut) {
  if (int e = check_dims(nx, ny, nz)) return e;
  KGrid g = make_kgrid(nx, ny, nz);
  const int64_t nc = (int64_t)nx * ny * g.nzc;
  if (kb_kcut > 0.0f) {  // deconv_paint with kernel_type = 'kaiser_bessel' (nbody.py:321-322)
    KbHat h = make_kbhat(order, kb_kcut);
    launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
      int l;
      KVec k = kvec_at(g, e, l);
      float c = 1.0f / kb_window_hat(h, k);
      cfloat v = in[e];
      out[e] = cfloat{v.re * c, v.im * c};
    });
    return rt_check("deconv_kb");
  }
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float c = 1.0f / window_hat(k, order);
    cfloat v = in[e];
    out[e] = cfloat{v.re * c, v.im * c};
  });
  return rt_check("deconv");
}

// out = scale / W^p * (1/m) sum_i in_i exp(+i (i/m)(kx+ky+kz))      (nbody.py:523-526, 571-574)
int interlace_combine(stream_t st, const cfloat* in_m, cfloat* out, int m, int nx, int ny, int nz, float scale,
                      int deconv_order, SlabK sk) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (m < 1 || m > 8) {
    set_error("interlace order must be in 1..8");
    return MCPM_EINVAL;
  }
  KGrid g = make_kgrid(nx, ny, nz, 0, 0, sk);
  const int64_t nc = (int64_t)nx * g.ny_loc * g.nzc;
  const float invm = 1.0f / (float)m;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float ks = k.kx + k.ky + k.kz;
    float c = scale * invm;
    if (deconv_order > 0) c /= window_hat(k, deconv_order);
    float re = 0.0f, im = 0.0f;
    for (int i = 0; i < m; ++i) {
      cfloat v = in_m[i * nc + e];
      float sn, cs;
      sincosf((float)i * invm * ks, &sn, &cs);
      re += v.re * cs - v.im * sn;
      im += v.re * sn + v.im * cs;
    }
    out[e] = cfloat{re * c, im * c};
  });
  return rt_check("interlace_combine");
}

// transpose: out_i = (norm / w') * conj(kernel_i) * in.  The paint-mesh cotangent is irfftn(out_i) with norm = N,
// or a raw (unnormalised) C2R of out_i with norm = 1.
int interlace_combine_T(stream_t st, const cfloat* in, cfloat* out_m, int m, int nx, int ny, int nz, float scale,
                        int deconv_order, float norm, SlabK sk, int half_weights) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (m < 1 || m > 8) {
    set_error("interlace order must be in 1..8");
    return MCPM_EINVAL;
  }
  KGrid g = make_kgrid(nx, ny, nz, 0, 0, sk);
  const int64_t nc = (int64_t)nx * g.ny_loc * g.nzc;
  const float invm = 1.0f / (float)m;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec k = kvec_at(g, e, l);
    float ks = k.kx + k.ky + k.kz;
    float c = scale * invm * norm;
    if (half_weights) c /= half_weight(l, g.nz);
    if (deconv_order > 0) c /= window_hat(k, deconv_order);
    cfloat v = in[e];
    for (int i = 0; i < m; ++i) {
      float sn, cs;
      sincosf((float)i * invm * ks, &sn, &cs);
      // (cs - i sn)(a + ib)
      out_m[i * nc + e] = cfloat{(v.re * cs + v.im * sn) * c, (v.im * cs - v.re * sn) * c};
    }
  });
  return rt_check("interlace_combine_T");
}

int scale_spectrum(stream_t st, const cfloat* in, const float* t, cfloat* out, int64_t nc) {
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    cfloat v = in[e];
    float s = t[e];
    out[e] = cfloat{v.re * s, v.im * s};
  });
  return rt_check("scale_spectrum");
}

int scale_real(stream_t st, const float* in, float sc, float* out, int64_t n) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t e) { out[e] = in[e] * sc; });
  return rt_check("scale_real");
}

// Hermitian weights of the half spectrum in the real inner product (w' = 1 on kz = 0 / Nyquist, else 2):
//   mode 0: out = in * N / w'   (rfftn^T:  xbar = irfftn(out))      mode 1: out = in * w' / N   (irfftn^T: ybar = out of rfftn(xbar))
int hermitian_weights(stream_t st, const cfloat* in, cfloat* out, int nx, int ny, int nz, int mode) {
  if (int e = check_dims(nx, ny, nz)) return e;
  const int nzc = nz / 2 + 1;
  const int64_t nc = (int64_t)nx * ny * nzc;
  const float nn = (float)((double)nx * ny * nz);
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l = (int)(e % nzc);
    float w = half_weight(l, nz);
    float c = mode == 0 ? nn / w : w / nn;
    cfloat v = in[e];
    out[e] = cfloat{v.re * c, v.im * c};
  });
  return rt_check("hermitian_weights");
}

// Hermitian projection of `batch` half spectra, in place, on the two self-conjugate planes kz = 0 and kz = Nyquist:
//   A(i, j, l) <- (A(i, j, l) + conj A(-i, -j, l)) / 2.
// jnp.fft.irfftn returns the real part of the full inverse transform, i.e. it applies this projection implicitly
// (its last-axis C2R ignores the imaginary parts left on those planes after the x and y transforms).  cuFFT's C2R on
// inconsistent input is algorithm dependent instead: the even-length real transform folds Im X(0) and Im X(N/2) into its
// output.  Spectra that are not Hermitian by construction -- the interlaced sum (its shift phase breaks the symmetry on
// the Nyquist planes, nbody.py:524-525), cotangents of free complex arrays, whatever a caller hands mcpm_irfftn -- are
// therefore projected before their C2R.  One thread per conjugate pair: 2 * nx * ny elements per mesh, negligible.
int hermitian_project(stream_t st, cfloat* data, int nx, int ny, int nz, int batch) {
  if (int e = check_dims(nx, ny, nz)) return e;
  const int nzc = nz / 2 + 1;
  const int64_t plane = (int64_t)nx * ny;
  const int64_t nc = plane * nzc;
  launch_1d(st, (int64_t)batch * 2 * plane, [=] MCPM_LAMBDA(int64_t t) {
    int64_t r = t % plane;
    int64_t bw = t / plane;
    int l = (bw & 1) ? nz / 2 : 0;
    cfloat* A = data + (bw >> 1) * nc;
    int i = (int)(r / ny), j = (int)(r % ny);
    int ip = i ? nx - i : 0, jp = j ? ny - j : 0;
    int64_t rp = (int64_t)ip * ny + jp;
    if (rp < r) return;  // the partner's thread handles this pair
    cfloat a = A[r * nzc + l];
    if (rp == r) {
      A[r * nzc + l] = cfloat{a.re, 0.0f};
      return;
    }
    cfloat c = A[rp * nzc + l];
    cfloat h = cfloat{0.5f * (a.re + c.re), 0.5f * (a.im - c.im)};
    A[r * nzc + l] = h;
    A[rp * nzc + l] = cfloat{h.re, -h.im};
  });
  return rt_check("hermitian_project");
}

// out (+)= a * w'(kz) * in, or a / w' * in when inverse != 0, on any block of a half spectrum whose fastest axis is the
// full kz axis (nzc = nz/2+1 entries): the Hermitian weights of hermitian_weights above without the 1/N bookkeeping,
// for callers that keep their own normalisation (the slab-decomposed model, whose distributed FFTs are unnormalised).
int half_weight_axpy(stream_t st, const cfloat* in, cfloat* out, int64_t nc, int nz, float a, int inverse,
                     int accumulate) {
  if (nz <= 0 || (nz & 1)) {
    set_error("half_weight_axpy: nz must be positive and even");
    return MCPM_EINVAL;
  }
  const int nzc = nz / 2 + 1;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l = (int)(e % nzc);
    float w = half_weight(l, nz);
    float c = inverse ? a / w : a * w;
    cfloat v = in[e];
    if (accumulate) {
      cfloat o = out[e];
      out[e] = cfloat{o.re + v.re * c, o.im + v.im * c};
    } else {
      out[e] = cfloat{v.re * c, v.im * c};
    }
  });
  return rt_check("half_weight_axpy");
}

// ---------------------------------------------------------------------------------------------------- spectrum
// Binned auto / cross power of half spectra, monopole (metrics.py:121-182): for every element, bin = np.digitize(|k|,
// kedges) with k = 2 pi f / box_size, weight w' = 1 on kz = 0 / Nyquist else 2 (Hermitian double counting), and
//   out[0][bin] += w'   out[1][bin] += w' |k|   out[2][bin] += w' Re(m0 conj m1)   out[3][bin] += w' Im(m0 conj m1)
// after dividing m_i by rectangular_hat^deconv_i (cell units).  Wavenumbers and sums are float64 so that bin membership
// matches the float64 reference exactly; out = [4][n_edges + 1] float64, accumulated into (caller zeroes it).
struct Box3 {
  double x, y, z;
};

// Multipole ell along the unit line of sight `los` (metrics.py:165-166): the power sums carry (2 ell + 1) L_ell(mu),
// mu = k . los / |k| (0 at k = 0, safe_div); the count and the mean wavenumber do not.  ell = 0 is the monopole.
int spectrum_bins(stream_t st, const cfloat* m0, const cfloat* m1, int nx, int ny, int nz, double bx, double by,
                  double bz, const double* kedges, int n_edges, int deconv0, int deconv1, double* out, int ell,
                  double lx, double ly, double lz) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if (ell < 0 || ell > 64) {
    set_error("spectrum_bins: multipole order must be in 0..64");
    return MCPM_EINVAL;
  }
  const Box3 box = {bx, by, bz};
  const Box3 los = {lx, ly, lz};
  if (n_edges < 1) {
    set_error("spectrum_bins: at least one bin edge is required");
    return MCPM_EINVAL;
  }
  KGrid g = make_kgrid(nx, ny, nz);
  const int64_t nc = (int64_t)nx * ny * g.nzc;
  const int nb = n_edges + 1;
  const double twopi = 6.283185307179586476925;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    int l;
    KVec kc = kvec_at(g, e, l);  // cell units, for the deconvolution window
    const int64_t r = e / g.nzc;
    const int j = (int)(r % g.ny), i = (int)(r / g.ny);
    const double kx = twopi * signed_freq(i, g.nx) / box.x, ky = twopi * signed_freq(j, g.ny) / box.y,
                 kz = twopi * l / box.z;
    const double kk = sqrt(kx * kx + ky * ky + kz * kz);
    int lo = 0, hi = n_edges;  // number of edges <= kk
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (kedges[mid] <= kk) lo = mid + 1;
      else hi = mid;
    }
    const double w = half_weight(l, g.nz);
    double leg = 1.0;
    if (ell > 0) {  // Bonnet recurrence, float64
      const double mu = kk == 0.0 ? 0.0 : (kx * los.x + ky * los.y + kz * los.z) / kk;
      double p0 = 1.0, p1 = mu;
      for (int n = 1; n < ell; ++n) {
        const double p2 = ((2 * n + 1) * mu * p1 - n * p0) / (n + 1);
        p0 = p1;
        p1 = p2;
      }
      leg = (2 * ell + 1) * p1;
    }
    cfloat a = m0[e];
    cfloat b = m1 ? m1[e] : a;
    float ca = deconv0 > 0 ? 1.0f / window_hat(kc, deconv0) : 1.0f;
    float cb = deconv1 > 0 ? 1.0f / window_hat(kc, deconv1) : 1.0f;
    if (!m1) cb = ca;
    const double ar = (double)a.re * ca, ai = (double)a.im * ca, br = (double)b.re * cb, bi = (double)b.im * cb;
    atomic_add(out + lo, w);
    atomic_add(out + nb + lo, w * kk);
    atomic_add(out + 2 * nb + lo, w * leg * (ar * br + ai * bi));
    if (m1) atomic_add(out + 3 * nb + lo, w * leg * (ai * br - ar * bi));
  });
  return rt_check("spectrum_bins");
}

// ---------------------------------------------------------------------------------------------------- rg2cgh
// rg2cgh / cgh2rg (utils.py:785-921): a real Gaussian mesh [nx,ny,nz] <-> a complex Gaussian Hermitian half spectrum
// [nx,ny,nz/2+1] by permutation and reweighting, so that rg2cgh(N(0,I)) is distributed as rfftn(N(0,I)).
// With (hx,hy,hz) = shape / 2 the N real numbers are laid out as
//   planes 0 < z < hz: Re of (x,y,l=z)                 planes z > hz: Im of (x,y,l=z-hz)
//   on the planes z = 0, hz (l = z):  rows 0 < y < hy: Re of (x,y,l)        rows y > hy: Im of (x,y-hy,l)
//     on the rows y = 0, hy:  0 < x < hx: Re of (x,y,l)   x > hx: Im of (x-hx,y,l)   x = 0, hx: Re / sqrt2 (Im = 0)
// and the other half of each self-conjugate plane / row follows from Hermitian symmetry, F(-k) = conj F(k).
struct RgSrc {
  int64_t re, im;  // source offsets in the real mesh (im < 0: no imaginary part)
  float sre, sim;  // weights
};

MCPM_HD RgSrc rg_source(int i, int j, int l, int nx, int ny, int nz) {
  const int hx = nx / 2, hy = ny / 2, hz = nz / 2;
  RgSrc r;
  r.sre = 1.0f;
  r.sim = 1.0f;
  auto at = [&](int x, int y, int z) { return ((int64_t)x * ny + y) * nz + z; };
  if (l > 0 && l < hz) {
    r.re = at(i, j, l);
    r.im = at(i, j, hz + l);
    return r;
  }
  const int z = l;  // 0 or hz: self-conjugate plane
  if (j != 0 && j != hy) {
    if (j < hy) {
      r.re = at(i, j, z);
      r.im = at(i, hy + j, z);
    } else {  // conjugate of the partner (-i, -j)
      const int ip = i == 0 ? 0 : nx - i, jp = ny - j;
      r.re = at(ip, jp, z);
      r.im = at(ip, hy + jp, z);
      r.sim = -1.0f;
    }
    return r;
  }
  if (i != 0 && i != hx) {
    if (i < hx) {
      r.re = at(i, j, z);
      r.im = at(hx + i, j, z);
    } else {
      const int ip = nx - i;
      r.re = at(ip, j, z);
      r.im = at(hx + ip, j, z);
      r.sim = -1.0f;
    }
    return r;
  }
  r.re = at(i, j, z);  // one of the 8 real modes
  r.im = -1;
  r.sre = 1.41421356237309504880f;
  r.sim = 0.0f;
  return r;
}

// out = scale * [transfer *] rg2cgh_unit(mesh); scale = sqrt(N/2) (norm "backward"), 1/sqrt2 ("ortho"), 1/sqrt(2N) ("forward")
int rg2cgh(stream_t st, const float* mesh, cfloat* out, int nx, int ny, int nz, float scale, const float* transfer) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if ((nx & 1) || (ny & 1)) {
    set_error("rg2cgh: dimension lengths must be even.");
    return MCPM_EINVAL;
  }
  const int nzc = nz / 2 + 1;
  const int64_t nc = (int64_t)nx * ny * nzc;
  launch_1d(st, nc, [=] MCPM_LAMBDA(int64_t e) {
    const int l = (int)(e % nzc);
    const int64_t r = e / nzc;
    const int j = (int)(r % ny), i = (int)(r / ny);
    const RgSrc s = rg_source(i, j, l, nx, ny, nz);
    const float c = transfer ? scale * transfer[e] : scale;
    out[e] = cfloat{mesh[s.re] * (s.sre * c), s.im >= 0 ? mesh[s.im] * (s.sim * c) : 0.0f};
  });
  return rt_check("rg2cgh");
}

// Role of a real-mesh element: which spectrum element(s) it feeds.  part 0: real part, 1: imaginary part.
struct RgDst {
  int64_t a, b;  // spectrum elements fed (b < 0: only one); the second one is the Hermitian partner
  int part;
  float wa, wb;
};

MCPM_HD RgDst rg_dest(int x, int y, int z, int nx, int ny, int nz) {
  const int hx = nx / 2, hy = ny / 2, hz = nz / 2, nzc = hz + 1;
  auto at = [&](int i, int j, int l) { return ((int64_t)i * ny + j) * nzc + l; };
  RgDst d;
  d.b = -1;
  d.wa = 1.0f;
  d.wb = 0.0f;
  if (z != 0 && z != hz) {
    d.part = z > hz;
    d.a = at(x, y, z > hz ? z - hz : z);
    return d;
  }
  if (y != 0 && y != hy) {
    d.part = y > hy;
    const int j = y > hy ? y - hy : y;
    d.a = at(x, j, z);
    d.b = at(x == 0 ? 0 : nx - x, ny - j, z);
    d.wb = d.part ? -1.0f : 1.0f;
    return d;
  }
  if (x != 0 && x != hx) {
    d.part = x > hx;
    const int i = x > hx ? x - hx : x;
    d.a = at(i, y, z);
    d.b = at(nx - i, y, z);
    d.wb = d.part ? -1.0f : 1.0f;
    return d;
  }
  d.part = 0;
  d.a = at(x, y, z);
  d.wa = 1.41421356237309504880f;
  return d;
}

// VJP of rg2cgh w.r.t. the real mesh: meshbar = scale * sum over the (one or two) spectrum elements fed of
// weight * [transfer] * (Re or Im of outbar)   (cotangent convention dL/dRe + i dL/dIm).
int rg2cgh_vjp(stream_t st, const cfloat* outbar, float* meshbar, int nx, int ny, int nz, float scale,
               const float* transfer) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if ((nx & 1) || (ny & 1)) {
    set_error("rg2cgh: dimension lengths must be even.");
    return MCPM_EINVAL;
  }
  const int64_t n = (int64_t)nx * ny * nz;
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t e) {
    const int z = (int)(e % nz);
    const int64_t r = e / nz;
    const int y = (int)(r % ny), x = (int)(r / ny);
    const RgDst d = rg_dest(x, y, z, nx, ny, nz);
    cfloat va = outbar[d.a];
    float acc = (d.part ? va.im : va.re) * d.wa * (transfer ? transfer[d.a] : 1.0f);
    if (d.b >= 0) {
      cfloat vb = outbar[d.b];
      acc += (d.part ? vb.im : vb.re) * d.wb * (transfer ? transfer[d.b] : 1.0f);
    }
    meshbar[e] = acc * scale;
  });
  return rt_check("rg2cgh_vjp");
}

// cgh2rg: the inverse permutation, mesh = inv_scale * (Re or Im of the element it came from) / weight
int cgh2rg(stream_t st, const cfloat* meshk, float* mesh, int nx, int ny, int nz, float inv_scale) {
  if (int e = check_dims(nx, ny, nz)) return e;
  if ((nx & 1) || (ny & 1)) {
    set_error("cgh2rg: dimension lengths must be even.");
    return MCPM_EINVAL;
  }
  const int64_t n = (int64_t)nx * ny * nz;
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t e) {
    const int z = (int)(e % nz);
    const int64_t r = e / nz;
    const int y = (int)(r % ny), x = (int)(r / ny);
    const RgDst d = rg_dest(x, y, z, nx, ny, nz);
    // the reference keeps the value found at the Hermitian partner where there is one (utils.py:860-868)
    const int64_t src = d.b >= 0 ? d.b : d.a;
    const float w = d.b >= 0 ? d.wb : 1.0f / d.wa;
    cfloat v = meshk[src];
    mesh[e] = (d.part ? v.im : v.re) * w * inv_scale;
  });
  return rt_check("cgh2rg");
}

// ---------------------------------------------------------------------------------------------------- chreshape
// One thread per OUTPUT element gathers its (at most 4 x 2) sources.  Per axis (utils.py:975-1013):
//   crop  (s < ms), axes x,y: the new Nyquist row (freq -s/2) = (in[+s/2] + in[-s/2]) / sqrt2
//   pad   (s > ms), axes x,y: rows of freq -ms/2 and +ms/2 both = in[-ms/2] / sqrt2; |freq| > ms/2 is zero
//   crop on kz: the new Nyquist plane l = s-1 = (P + conj(P[-kx,-ky])) / sqrt2 evaluated on the INPUT grid
//   pad  on kz: the old Nyquist plane l = ms-1 is divided by sqrt2; l > ms-1 is zero
struct AxisSrc {
  int n;       // number of sources (0, 1 or 2)
  int idx[2];  // input indices
  float w;     // common weight
};

MCPM_HD AxisSrc axis_sources(int i_out, int s, int ms) {
  AxisSrc a;
  int f = signed_freq(i_out, s);
  const float r = 0.70710678118654752440f;
  if (s < ms) {
    if (2 * f == -s) {
      a.n = 2;
      a.idx[0] = s / 2;
      a.idx[1] = ms - s / 2;
      a.w = r;
    } else {
      a.n = 1;
      a.idx[0] = f < 0 ? f + ms : f;
      a.idx[1] = 0;
      a.w = 1.0f;
    }
  } else if (s > ms) {
    int h = ms / 2;
    if (f == -h || f == h) {
      a.n = 1;
      a.idx[0] = ms - h;
      a.idx[1] = 0;
      a.w = r;
    } else if (f > -h && f < h) {
      a.n = 1;
      a.idx[0] = f < 0 ? f + ms : f;
      a.idx[1] = 0;
      a.w = 1.0f;
    } else {
      a.n = 0;
      a.idx[0] = a.idx[1] = 0;
      a.w = 0.0f;
    }
  } else {
    a.n = 1;
    a.idx[0] = i_out;
    a.idx[1] = 0;
    a.w = 1.0f;
  }
  return a;
}

int chreshape(stream_t st, const cfloat* in, int inx, int iny, int inz, cfloat* out, int onx, int ony, int onz) {
  if (int e = check_dims(inx, iny, inz)) return e;
  if (int e = check_dims(onx, ony, onz)) return e;
  if ((inx & 1) || (iny & 1) || (onx & 1) || (ony & 1)) {
    set_error("chreshape: mesh sides must be even");
    return MCPM_EINVAL;
  }
  const int inzc = inz / 2 + 1, onzc = onz / 2 + 1;
  const int64_t nout = (int64_t)onx * ony * onzc;
  const float scale = (float)(((double)onx * ony * onz) / ((double)inx * iny * inz));
  launch_1d(st, nout, [=] MCPM_LAMBDA(int64_t e) {
    int l = (int)(e % onzc);
    int64_t r = e / onzc;
    int j = (int)(r % ony);
    int i = (int)(r / ony);
    AxisSrc ax = axis_sources(i, onx, inx);
    AxisSrc ay = axis_sources(j, ony, iny);
    const float rs = 0.70710678118654752440f;
    float wl = 1.0f;
    bool herm = false, zero = false;
    if (onzc < inzc) {
      herm = (l == onzc - 1);
      if (herm) wl = rs;
    } else if (onzc > inzc) {
      if (l == inzc - 1) wl = rs;
      zero = l > inzc - 1;
    }
    float re = 0.0f, im = 0.0f;
    if (!zero) {
      for (int a = 0; a < ax.n; ++a)
        for (int b = 0; b < ay.n; ++b) {
          int ii = ax.idx[a], jj = ay.idx[b];
          cfloat v = in[((int64_t)ii * iny + jj) * inzc + l];
          float vr = v.re, vi = v.im;
          if (herm) {
            int im_ = ii == 0 ? 0 : inx - ii, jm = jj == 0 ? 0 : iny - jj;
            cfloat u = in[((int64_t)im_ * iny + jm) * inzc + l];
            vr += u.re;
            vi -= u.im;
          }
          re += vr;
          im += vi;
        }
    }
    float w = ax.w * ay.w * wl * scale;
    out[e] = cfloat{re * w, im * w};
  });
  return rt_check("chreshape");
}

// Slab form of the crop (slab-decomposed final paint with a finer paint mesh, SURVEY 8e / BASELINE C5).  The spectrum
// is split along ky: `in` = [inx, rows, inz/2+1] holds the global rows y0 .. y0+rows-1 of an iny-row spectrum.  This
// pass crops x and kz (onx <= inx, onz <= inz) and leaves the rows alone; the caller then redistributes rows (the ky
// crop is a row selection plus one 1/sqrt2-weighted sum for the new Nyquist row).  Cropping kz folds the mirrored
// element conj in[-kx, -ky, l] into the new Nyquist plane l = onz/2; the mirrored row lives on another rank, so that
// one plane of the WHOLE input spectrum is passed separately (`nyq_plane` = [inx, iny], all-gathered by the caller).
// `row_w` (nullable) weights each local row (the caller's 1/sqrt2 on the two rows that merge into the new ky Nyquist).
int chreshape_crop_xz_slab(stream_t st, const cfloat* in, int inx, int iny, int inz, int rows, int y0,
                           const cfloat* nyq_plane, const float* row_w, cfloat* out, int onx, int onz, float scale) {
  if (int e = check_dims(inx, iny, inz)) return e;
  if (onx <= 0 || onz <= 0 || (onx & 1) || (onz & 1) || (inx & 1) || onx > inx || onz > inz) {
    set_error("chreshape_crop_xz_slab: even sides with onx <= inx and onz <= inz are required");
    return MCPM_EINVAL;
  }
  const int inzc = inz / 2 + 1, onzc = onz / 2 + 1;
  if (onzc < inzc && !nyq_plane) {
    set_error("chreshape_crop_xz_slab: cropping kz needs the gathered Nyquist plane");
    return MCPM_EINVAL;
  }
  const int64_t nout = (int64_t)onx * rows * onzc;
  launch_1d(st, nout, [=] MCPM_LAMBDA(int64_t e) {
    int l = (int)(e % onzc);
    int64_t r = e / onzc;
    int jr = (int)(r % rows);
    int i = (int)(r / rows);
    AxisSrc ax = axis_sources(i, onx, inx);
    const bool herm = onzc < inzc && l == onzc - 1;
    const float wl = herm ? 0.70710678118654752440f : 1.0f;
    float re = 0.0f, im = 0.0f;
    for (int a = 0; a < ax.n; ++a) {
      int ii = ax.idx[a];
      cfloat v = in[((int64_t)ii * rows + jr) * inzc + l];
      re += v.re;
      im += v.im;
      if (herm) {
        int im_ = ii == 0 ? 0 : inx - ii, j2 = y0 + jr, jm = j2 == 0 ? 0 : iny - j2;
        cfloat u = nyq_plane[(int64_t)im_ * iny + jm];
        re += u.re;
        im -= u.im;
      }
    }
    float w = ax.w * wl * scale * (row_w ? row_w[jr] : 1.0f);
    out[e] = cfloat{re * w, im * w};
  });
  return rt_check("chreshape_crop_xz_slab");
}

// Its transpose in the real inner product.  `inbar` [inx, rows, inzc] is zeroed here; the mirrored contributions of the
// new Nyquist plane are ACCUMULATED into `nyq_plane_bar` [inx, iny] (zeroed by the caller), which the caller sums over
// ranks and adds to plane l = onz/2 of its own rows.
int chreshape_crop_xz_slab_T(stream_t st, const cfloat* outbar, int onx, int onz, int rows, int y0, const float* row_w,
                             cfloat* inbar, int inx, int iny, int inz, cfloat* nyq_plane_bar, float scale) {
  if (int e = check_dims(inx, iny, inz)) return e;
  if (onx <= 0 || onz <= 0 || (onx & 1) || (onz & 1) || (inx & 1) || onx > inx || onz > inz) {
    set_error("chreshape_crop_xz_slab_T: even sides with onx <= inx and onz <= inz are required");
    return MCPM_EINVAL;
  }
  const int inzc = inz / 2 + 1, onzc = onz / 2 + 1;
  if (onzc < inzc && !nyq_plane_bar) {
    set_error("chreshape_crop_xz_slab_T: cropping kz needs the Nyquist-plane accumulator");
    return MCPM_EINVAL;
  }
  const int64_t nout = (int64_t)onx * rows * onzc;
  rt_memset(inbar, 0, sizeof(cfloat) * (size_t)inx * rows * inzc, st);
  float* ib = reinterpret_cast<float*>(inbar);
  float* pb = reinterpret_cast<float*>(nyq_plane_bar);
  launch_1d(st, nout, [=] MCPM_LAMBDA(int64_t e) {
    int l = (int)(e % onzc);
    int64_t r = e / onzc;
    int jr = (int)(r % rows);
    int i = (int)(r / rows);
    AxisSrc ax = axis_sources(i, onx, inx);
    const bool herm = onzc < inzc && l == onzc - 1;
    const float wl = herm ? 0.70710678118654752440f : 1.0f;
    cfloat v = outbar[e];
    float w = ax.w * wl * scale * (row_w ? row_w[jr] : 1.0f);
    float re = v.re * w, im = v.im * w;
    for (int a = 0; a < ax.n; ++a) {
      int ii = ax.idx[a];
      int64_t s = ((int64_t)ii * rows + jr) * inzc + l;
      atomic_add(ib + 2 * s, re);
      atomic_add(ib + 2 * s + 1, im);
      if (herm) {
        int im_ = ii == 0 ? 0 : inx - ii, j2 = y0 + jr, jm = j2 == 0 ? 0 : iny - j2;
        int64_t s2 = (int64_t)im_ * iny + jm;
        atomic_add(pb + 2 * s2, re);
        atomic_add(pb + 2 * s2 + 1, -im);
      }
    }
  });
  return rt_check("chreshape_crop_xz_slab_T");
}

// Transpose of chreshape in the real inner product (its VJP): every OUTPUT-side cotangent element scatters to the
// input elements it was gathered from; the Hermitian-mirrored source receives the conjugate.  `inbar` is zeroed here.
int chreshape_T(stream_t st, const cfloat* outbar, int onx, int ony, int onz, cfloat* inbar, int inx, int iny, int inz) {
  if (int e = check_dims(inx, iny, inz)) return e;
  if (int e = check_dims(onx, ony, onz)) return e;
  if ((inx & 1) || (iny & 1) || (onx & 1) || (ony & 1)) {
    set_error("chreshape: mesh sides must be even");
    return MCPM_EINVAL;
  }
  const int inzc = inz / 2 + 1, onzc = onz / 2 + 1;
  const int64_t nout = (int64_t)onx * ony * onzc;
  const float scale = (float)(((double)onx * ony * onz) / ((double)inx * iny * inz));
  rt_memset(inbar, 0, sizeof(cfloat) * (size_t)inx * iny * inzc, st);
  float* ib = reinterpret_cast<float*>(inbar);
  launch_1d(st, nout, [=] MCPM_LAMBDA(int64_t e) {
    int l = (int)(e % onzc);
    int64_t r = e / onzc;
    int j = (int)(r % ony);
    int i = (int)(r / ony);
    AxisSrc ax = axis_sources(i, onx, inx);
    AxisSrc ay = axis_sources(j, ony, iny);
    const float rs = 0.70710678118654752440f;
    float wl = 1.0f;
    bool herm = false, zero = false;
    if (onzc < inzc) {
      herm = (l == onzc - 1);
      if (herm) wl = rs;
    } else if (onzc > inzc) {
      if (l == inzc - 1) wl = rs;
      zero = l > inzc - 1;
    }
    if (zero) return;
    cfloat v = outbar[e];
    float w = ax.w * ay.w * wl * scale;
    float re = v.re * w, im = v.im * w;
    for (int a = 0; a < ax.n; ++a)
      for (int b = 0; b < ay.n; ++b) {
        int ii = ax.idx[a], jj = ay.idx[b];
        int64_t s = ((int64_t)ii * iny + jj) * inzc + l;
        atomic_add(ib + 2 * s, re);
        atomic_add(ib + 2 * s + 1, im);
        if (herm) {
          int im_ = ii == 0 ? 0 : inx - ii, jm = jj == 0 ? 0 : iny - jj;
          int64_t s2 = ((int64_t)im_ * iny + jm) * inzc + l;
          atomic_add(ib + 2 * s2, re);
          atomic_add(ib + 2 * s2 + 1, -im);
        }
      }
  });
  return rt_check("chreshape_T");
}

}  // namespace mcpm
