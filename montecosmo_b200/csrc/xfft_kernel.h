// xfft_kernel.h -- kernel templates of the fused x-transform (see xfft.cu for the design notes); included by xfft.cu
// (nx = 64, 128, 256) and xfft_large.cu (nx = 512, 1024) so that the two sets of instantiations compile in parallel.
#pragma once
#include "engine.h"
#include "kspace.h"

namespace mcpm {

namespace xf {
// (ky,kz) columns per CTA: 16 x 8 B = one 128-byte segment per x-plane; 8 for the 32-element-per-thread transforms
// (nx = 512, 1024), whose register and shared-memory footprint per column is twice as large
__host__ __device__ constexpr int columns_per_cta(int n) { return n >= 512 ? 8 : 16; }
// exchange buffer geometry (float2 units): element (column c, row k1, n2) at c * col_stride + k1 * row_stride + n2.
// A half-warp (one shared-memory phase of a 64-bit access) holds 16 columns of one t (CT = 16: any odd column stride is
// conflict free) or 8 columns of two t's (CT = 8: column stride = 2 mod 16 and row stride R2 + 1).
__host__ __device__ constexpr int row_stride(int r1, int r2) { return columns_per_cta(r1 * r2) == 16 ? r2 : r2 + 1; }
__host__ __device__ constexpr int col_stride(int r1, int r2) {
  if (columns_per_cta(r1 * r2) == 16) return r1 * r2 + 1;
  int s = r1 * row_stride(r1, r2);
  while (s % 16 != 2) ++s;
  return s;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// d * exp(sign * 2 pi i * idx / 32), idx in 0..15 (folds to immediates once the caller's loops are unrolled)
__device__ __forceinline__ float2 tw32(float2 d, int idx, int sign) {
  const float C[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                       0.f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                       -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  const float S[16] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                       0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                       1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
  if (idx == 0) return d;
  if (idx == 8) return sign < 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
  const float c = C[idx], s = sign < 0 ? -S[idx] : S[idx];
  return make_float2(d.x * c - d.y * s, d.x * s + d.y * c);
}

__host__ __device__ constexpr int brev(int k, int n) {
  int r = 0;
  for (int b = 1; b < n; b <<= 1) {
    r = (r << 1) | (k & 1);
    k >>= 1;
  }
  return r;
}

// in-register FFT of R (<= 32) points, natural order in and out: radix-2 decimation in frequency + bit reversal
template <int R, int SIGN>
__device__ __forceinline__ void fft_reg(float2* v) {
#pragma unroll
  for (int L = R; L >= 2; L >>= 1) {
#pragma unroll
    for (int b = 0; b < R; b += L) {
#pragma unroll
      for (int i = 0; i < L / 2; ++i) {
        const float2 p = v[b + i], q = v[b + i + L / 2];
        v[b + i] = make_float2(p.x + q.x, p.y + q.y);
        v[b + i + L / 2] = tw32(make_float2(p.x - q.x, p.y - q.y), i * (32 / L), SIGN);
      }
    }
  }
  float2 tmp[R];
#pragma unroll
  for (int k = 0; k < R; ++k) tmp[k] = v[brev(k, R)];
#pragma unroll
  for (int k = 0; k < R; ++k) v[k] = tmp[k];
}

// Length-N transform of one column spread over R2 threads.  In: v[m] = a[t + R2*m], m < R1.  Out: v[j*R2 + k2] =
// A[t + R2*j + R1*k2] -- as input to another col_fft that is element m = j + (R1/R2)*k2.  One __syncthreads.
template <int R1, int R2, int SIGN>
__device__ __forceinline__ void col_fft(float2 (&v)[R1], int t, float2* colbuf, const float2* tw) {
  constexpr int N = R1 * R2, J = R1 / R2, RS = row_stride(R1, R2);
  fft_reg<R1, SIGN>(v);
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) {
    float2 w = tw[(t * k1) & (N - 1)];
    if (SIGN > 0) w.y = -w.y;
    v[k1] = cmul(v[k1], w);
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) colbuf[k1 * RS + t] = v[k1];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) v[j * R2 + n2] = colbuf[(t + R2 * j) * RS + n2];
#pragma unroll
  for (int j = 0; j < J; ++j) fft_reg<R2, SIGN>(&v[j * R2]);
}

struct Args {
  const cfloat* in;   // `nin` arrays [nx][M], `cstride` elements apart
  cfloat* out;        // `nout` arrays, `cstride` apart
  int64_t cstride;
  int M;              // columns = ny_loc * nzc
  KGrid g;
  float r2, norm;
  int deconv_order;
  int half_weights;   // spectrum-out modes: cotangent of a free complex array (w'/N weights, no Hermitian projection)
  int accumulate;     // spectrum-out modes: out += result
  // Peer mode (FORCE / FORCE_T on a slab-decomposed mesh, npeer > 0): x-plane p lives on rank p / xl as plane p % xl of
  // that rank's buffer [ncomp][xl][ny][nzc]; this rank transforms the columns of its ky block y0 .. y0 + ny_loc - 1 and
  // reads / writes every plane in its owner's memory (NVLink peer pointers): the kernel IS the all-to-all.
  int npeer, xl;
  const cfloat* in_peer[8];
  cfloat* out_peer[8];
};

// What sits between the x-transforms (fourier.cu names) and which side is already / stays in k-space:
//   FORCE    out_j = IFFT_x[ -(i g_j) c FFT_x[in] ]                 1 -> 3   force_spectra between the x-passes
//   FORCE_T  out   = IFFT_x[ i c sum_j g_j FFT_x[in_j] ]            3 -> 1   force_spectra_T
//   FORCE_K  out_j = IFFT_x[ -(i g_j) c in ]                        1 -> 3   `in` is a full 3-D spectrum (lpt)
//   HESS_K   out_t = IFFT_x[ -G g_a g_b in ], t = 00 11 22 01 02 12 1 -> 6   hessian_spectra, `in` a 3-D spectrum
//   FORCE_TK out (+)= [w'/N] i c sum_j g_j FFT_x[in_j]              3 -> 1   force_spectra_T, `out` a 3-D spectrum
//   HESS_TK  out (+)= [w'/N] -G sum_t g_a g_b FFT_x[in_t]           6 -> 1   hessian_spectra_T, `out` a 3-D spectrum
//   FORCE2   out_0 = IFFT_x[ -(i g_x) c FFT_x[in] ], out_1 = IFFT_x[ c FFT_x[in] ]   1 -> 2   force along x + potential:
//            the y and z components follow from the potential on the (y,z) spectra of each x-plane (fourier.cu:
//            yz_gradients), so a slab-decomposed caller moves TWO fields back instead of three
//   FORCE2_T out   = IFFT_x[ i c (g_x FFT_x[in_0] + FFT_x[in_1]) ]                   2 -> 1   its transpose: in_1 holds
//            g_y in_y + g_z in_z, combined on the local planes before the exchange
enum Mode { FORCE = 0, FORCE_T = 1, FORCE_K = 2, HESS_K = 3, FORCE_TK = 4, HESS_TK = 5, FORCE2 = 6, FORCE2_T = 7 };
__host__ __device__ constexpr int n_in(int mode) {
  return mode == FORCE_T || mode == FORCE_TK ? 3 : (mode == HESS_TK ? 6 : (mode == FORCE2_T ? 2 : 1));
}
__host__ __device__ constexpr int n_out(int mode) {
  return mode == FORCE || mode == FORCE_K ? 3 : (mode == HESS_K ? 6 : (mode == FORCE2 ? 2 : 1));
}
__host__ __device__ constexpr bool fwd_x(int mode) { return mode != FORCE_K && mode != HESS_K; }
__host__ __device__ constexpr bool inv_x(int mode) { return mode != FORCE_TK && mode != HESS_TK; }
__host__ __device__ constexpr bool hessian(int mode) { return mode == HESS_K || mode == HESS_TK; }

// PLAIN: no long-range filter and no deconvolution (the BullFrog loop, lpt): the scalar factor is a table look-up
// Register budget: two resident CTAs of 256 threads per SM (128 registers) -- 80 registers for three was measured
// slower (profiles/r1_tune_gather_xfuse.txt); the 32-element-per-thread transforms take what they need (one CTA).
template <int R1, int R2, int MODE, bool PLAIN>
__global__ void __launch_bounds__(R2* columns_per_cta(R1* R2), (R1 >= 32 ? 1 : 512 / (R2 * columns_per_cta(R1 * R2))))
    xfuse_kernel(Args a) {
  constexpr int N = R1 * R2, J = R1 / R2, NIN = n_in(MODE), NOUT = n_out(MODE);
  constexpr int CT = columns_per_cta(N), CS = col_stride(R1, R2);
  constexpr bool HESS = hessian(MODE);
  extern __shared__ float2 xsm[];
  float2* tw = xsm;                // [N]   exp(-2 pi i n / N)
  float2* xk = xsm + N;            // [N]   per-x kernel pieces: (lap_term(kx), grad_term(kx))
  float2* buf0 = xsm + 2 * N;      // [CT][CS] x 2: exchange buffers, alternating between transforms
  float2* buf1 = buf0 + CT * CS;
  const int c = threadIdx.x % CT, t = threadIdx.x / CT;
  for (int n = threadIdx.x; n < N; n += R2 * CT) {
    float sn, cs;
    sincospif(-2.0f * (float)n / (float)N, &sn, &cs);
    tw[n] = make_float2(cs, sn);
    const float kx = a.g.tx * (float)signed_freq(n, a.g.nx);
    xk[n] = make_float2(lap_term(kx, a.g.lap_fd), grad_term(kx, a.g.grad_fd));
  }
  __syncthreads();
  const int m = blockIdx.x * CT + c;
  const bool active = m < a.M;
  const int mm = active ? m : a.M - 1;  // inactive lanes compute on a valid column and skip the stores
  const int l = mm % a.g.nzc, jy = mm / a.g.nzc + a.g.y0;
  float2* cb0 = buf0 + c * CS;
  float2* cb1 = buf1 + c * CS;
  const int64_t M = a.M;
  // thread t touches x-planes t + R2*n: pointers to plane t of this column, strides in whole planes
  const float2* in = reinterpret_cast<const float2*>(a.in) + mm + t * M;
  float2* out = reinterpret_cast<float2*>(a.out) + mm + t * M;
  const int64_t MR2 = M * R2, MR1 = M * R1;
  // x-space element addresses: plane t + R2*n1 on the load side, t + R2*j + R1*k2 (position e) on the store side
  const bool peer = a.npeer > 0;
  const int64_t pplane = (int64_t)a.g.ny * a.g.nzc;                    // peer buffers hold full-ky planes
  const int64_t pcol = (int64_t)a.g.y0 * a.g.nzc + mm;                  // this column inside such a plane
  const int64_t pcomp = (int64_t)a.xl * pplane;                         // component stride of a peer buffer
  auto x_in = [&](int comp, int plane) -> const float2* {
    if (!peer) return in + comp * a.cstride + (int64_t)(plane - t) * M;
    const int r = plane / a.xl;
    return reinterpret_cast<const float2*>(a.in_peer[r]) + comp * pcomp + (int64_t)(plane - r * a.xl) * pplane + pcol;
  };
  auto x_out = [&](int comp, int plane) -> float2* {
    if (!peer) return out + comp * a.cstride + (int64_t)(plane - t) * M;
    const int r = plane / a.xl;
    return reinterpret_cast<float2*>(a.out_peer[r]) + comp * pcomp + (int64_t)(plane - r * a.xl) * pplane + pcol;
  };
  // x-space elements of a thread: m-th is plane t + R2*m.  k-space elements: position e = j*R2 + k2 is kx index
  // t + R2*j + R1*k2, i.e. plane offset (e/R2)*R2 + (e%R2)*R1 from plane t; as col_fft input it is m = j + J*k2.
  auto kx_index = [&](int e) { return t + R2 * (e / R2) + R1 * (e % R2); };
  auto k_offset = [&](int e) { return (e / R2) * MR2 + (e % R2) * MR1; };

  // per-column wavevector pieces (kspace.h)
  KVec k0;
  k0.kx = 0.f;
  k0.ky = a.g.ty * (float)signed_freq(jy, a.g.ny);
  k0.kz = a.g.tz * (float)l;
  const bool nqy = 2 * jy == a.g.ny, nqz = 2 * l == a.g.nz;
  // Hermitian projection on the self-conjugate planes, only where the result feeds a C2R (kspace.h: KVec)
  const bool proj = (l == 0 || nqz) && !((MODE == FORCE_TK || MODE == HESS_TK) && a.half_weights);
  const float gy = grad_term(k0.ky, a.g.grad_fd), gz = grad_term(k0.kz, a.g.grad_fd);
  const float lapyz = lap_term(k0.ky, a.g.lap_fd) + lap_term(k0.kz, a.g.lap_fd);
  auto scalar_at = [&](int e) {  // invlaplace * [gaussian] * [1 / window^2] at element e
    if (PLAIN || HESS) {
      const float kk = xk[kx_index(e)].x + lapyz;
      return kk == 0.0f ? 0.0f : -1.0f / kk;
    }
    KVec k = k0;
    k.kx = a.g.tx * (float)signed_freq(kx_index(e), a.g.nx);
    return force_scalar(a.g, k, a.r2, a.deconv_order);
  };
  // weight of component `comp` at element e: force g_j (projected), Hessian g_a g_b (mixed terms projected)
  auto weight = [&](int comp, int e) {
    const int i = kx_index(e);
    const float gx = xk[i].y;
    const bool nqx = 2 * i == a.g.nx;
    if (MODE == FORCE2 || MODE == FORCE2_T) return comp == 0 ? ((proj && nqx) ? 0.f : gx) : 1.0f;
    if (!HESS) {
      if (comp == 0) return (proj && nqx) ? 0.f : gx;
      if (comp == 1) return (proj && nqy) ? 0.f : gy;
      return (proj && nqz) ? 0.f : gz;
    }
    switch (comp) {
      case 0: return gx * gx;
      case 1: return gy * gy;
      case 2: return gz * gz;
      case 3: return (proj && nqx != nqy) ? 0.f : gx * gy;
      case 4: return (proj && nqx != nqz) ? 0.f : gx * gz;
      default: return (proj && nqy != nqz) ? 0.f : gy * gz;
    }
  };
  float outw = a.norm;  // common output factor of the column
  if ((MODE == FORCE_TK || MODE == HESS_TK) && a.half_weights)
    outw *= half_weight(l, a.g.nz) * (float)(1.0 / ((double)a.g.nx * a.g.ny * a.g.nz));

  // ---- gather side: X = sum_comp weight * [FFT_x] in_comp   (a single input carries weight 1)
  float2 X[R1];
  int nfft = 0;  // transforms done so far: picks the exchange buffer
  if (NIN == 1) {
    if (fwd_x(MODE)) {
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) X[n1] = *x_in(0, t + R2 * n1);
      col_fft<R1, R2, -1>(X, t, cb0, tw);
      nfft = 1;
    } else {
#pragma unroll
      for (int e = 0; e < R1; ++e) X[e] = in[k_offset(e)];
    }
  } else {
#pragma unroll
    for (int e = 0; e < R1; ++e) X[e] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int comp = 0; comp < NIN; ++comp) {
      float2 v[R1];
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) v[n1] = *x_in(comp, t + R2 * n1);
      col_fft<R1, R2, -1>(v, t, (comp & 1) ? cb1 : cb0, tw);
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const float g = weight(comp, e);
        X[e].x += g * v[e].x;
        X[e].y += g * v[e].y;
      }
    }
    nfft = NIN;
  }
  // ---- scalar factor.  Force: -(i) c for 1 -> 3, (+i) c for 3 -> 1.  Hessian: -G both ways.
#pragma unroll
  for (int e = 0; e < R1; ++e) {
    const float cs = scalar_at(e) * outw;
    if (HESS) X[e] = make_float2(-X[e].x * cs, -X[e].y * cs);
    else if (NOUT >= 2) X[e] = make_float2(X[e].y * cs, -X[e].x * cs);
    else X[e] = make_float2(-X[e].y * cs, X[e].x * cs);
  }
  // ---- scatter side
  if (NOUT == 1) {
    if (inv_x(MODE)) {
      float2 w[R1];
#pragma unroll
      for (int e = 0; e < R1; ++e) w[(e / R2) + J * (e % R2)] = X[e];
      col_fft<R1, R2, +1>(w, t, (nfft & 1) ? cb1 : cb0, tw);
      if (active) {
#pragma unroll
        for (int e = 0; e < R1; ++e) *x_out(0, kx_index(e)) = w[e];  // x-space: same plane pattern as the k-space one
      }
    } else if (active) {
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        float2 o = X[e];
        if (a.accumulate) {
          const float2 p = out[k_offset(e)];
          o.x += p.x;
          o.y += p.y;
        }
        out[k_offset(e)] = o;
      }
    }
  } else {
#pragma unroll 1
    for (int comp = 0; comp < NOUT; ++comp) {
      float2 w[R1];
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const float g = weight(comp, e);
        if (MODE == FORCE2 && comp == 1)  // the potential: c X, i.e. i times the -(i) c X held in X
          w[(e / R2) + J * (e % R2)] = make_float2(-X[e].y, X[e].x);
        else
          w[(e / R2) + J * (e % R2)] = make_float2(X[e].x * g, X[e].y * g);  // inverse-transform input order
      }
      col_fft<R1, R2, +1>(w, t, ((nfft + comp) & 1) ? cb1 : cb0, tw);
      if (active) {
#pragma unroll
        for (int e = 0; e < R1; ++e) *x_out(comp, kx_index(e)) = w[e];
      }
    }
  }
}


// one launch of mode MODE for nx = R1 * R2 (PLAIN chosen from the arguments)
template <int R1, int R2, int MODE, bool PLAIN>
static void launch_one(stream_t st, const Args& a) {
  constexpr int N = R1 * R2, CT = columns_per_cta(N);
  const size_t smem = sizeof(float2) * (2 * N + 2 * CT * col_stride(R1, R2));
  cudaFuncSetAttribute(xfuse_kernel<R1, R2, MODE, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  xfuse_kernel<R1, R2, MODE, PLAIN><<<(a.M + CT - 1) / CT, R2 * CT, smem, st>>>(a);
}

template <int R1, int R2, int MODE>
static int launch(stream_t st, const Args& a) {
  count_launch();
  const bool plain = !(a.r2 > 0.f) && a.deconv_order <= 0;
  // filter / deconvolution variants exist only for the force operators of pm_forces (nbody.py:591-603)
  if constexpr (MODE == FORCE || MODE == FORCE_T || MODE == FORCE_K) {
    if (!plain) {
      launch_one<R1, R2, MODE, false>(st, a);
      return rt_check("xfuse") ? MCPM_ECUDA : 0;
    }
  } else if (!plain) {
    set_error("xfuse: this operator has no long-range filter / deconvolution variant");
    return MCPM_EINVAL;
  }
  launch_one<R1, R2, MODE, true>(st, a);
  return rt_check("xfuse") ? MCPM_ECUDA : 0;
}

template <int R1, int R2>
static int launch_mode(int mode, stream_t st, const Args& a) {
  switch (mode) {
    case FORCE: return launch<R1, R2, FORCE>(st, a);
    case FORCE_T: return launch<R1, R2, FORCE_T>(st, a);
    case FORCE_K: return launch<R1, R2, FORCE_K>(st, a);
    case HESS_K: return launch<R1, R2, HESS_K>(st, a);
    case FORCE_TK: return launch<R1, R2, FORCE_TK>(st, a);
    case HESS_TK: return launch<R1, R2, HESS_TK>(st, a);
    case FORCE2: return launch<R1, R2, FORCE2>(st, a);
    case FORCE2_T: return launch<R1, R2, FORCE2_T>(st, a);
  }
  set_error("xfuse: unknown mode");
  return MCPM_EINVAL;
}
}  // namespace xf

int xfuse_dispatch_large(int mode, stream_t st, const xf::Args& a);  // xfft_large.cu: nx = 512, 1024

}  // namespace mcpm
