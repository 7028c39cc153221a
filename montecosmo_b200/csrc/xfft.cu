// xfft.cu -- the x-direction transform of the 3-D FFT fused with the Fourier-space force kernel (sm_100a build only).
//
// pm_forces (nbody.py:583-604) is rfftn -> multiply by -(i k_j) G(k) -> 3 x irfftn.  cuFFT runs each 3-D transform as
// three passes over the mesh; the last pass of the forward transform and the first pass of each inverse transform are
// 1-D FFTs along x, with only the pointwise kernel between them.  Here those four x-passes and the multiply are ONE
// kernel: a CTA loads a tile of 16 (ky,kz) columns x all nx planes (128-byte segments), transforms along x in
// registers, applies the kernel, transforms back three times and stores the three spectra that the 2-D (y,z) C2R of
// each x-plane consumes.  Traffic 4N read + 12N written -- what the multiply alone used to move -- and four of the
// sixteen FFT passes of a force evaluation disappear.  The reverse-step operator (3 spectra -> 1) is the transpose.
//
// The 1-D transform of length N = R1*R2 is a two-pass Cooley-Tukey in registers: thread t of the R2 threads of a column
// holds the R1 elements {t + R2*m}; pass 1 is a radix-R1 FFT per thread, then twiddles and one exchange through
// shared memory, pass 2 is R1/R2 radix-R2 FFTs per thread.  The output lands on the same distribution (k = t mod R2),
// so the inverse transform starts from registers: one shared-memory exchange per transform, none between them.
#ifndef MCPM_HOSTEMU
#include "engine.h"
#include "kspace.h"

namespace mcpm {

namespace xf {
constexpr int CT = 16;  // (ky,kz) columns per CTA: 16 x 8 B = one 128-byte segment per x-plane

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// d * exp(sign * 2 pi i * idx / 16), idx in 0..7 (folds to immediates once the caller's loops are unrolled)
__device__ __forceinline__ float2 tw16(float2 d, int idx, int sign) {
  const float C[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                      0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
  const float S[8] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                      1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
  if (idx == 0) return d;
  if (idx == 4) return sign < 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
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

// in-register FFT of R (<= 16) points, natural order in and out: radix-2 decimation in frequency + bit reversal
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
        v[b + i + L / 2] = tw16(make_float2(p.x - q.x, p.y - q.y), i * (16 / L), SIGN);
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
  constexpr int N = R1 * R2, J = R1 / R2;
  fft_reg<R1, SIGN>(v);
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) {
    float2 w = tw[(t * k1) & (N - 1)];
    if (SIGN > 0) w.y = -w.y;
    v[k1] = cmul(v[k1], w);
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) colbuf[k1 * R2 + t] = v[k1];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) v[j * R2 + n2] = colbuf[(t + R2 * j) * R2 + n2];
#pragma unroll
  for (int j = 0; j < J; ++j) fft_reg<R2, SIGN>(&v[j * R2]);
}

struct Args {
  const cfloat* in;  // MODE 0: one spectrum [nx][M]; MODE 1: three, `cstride` elements apart
  cfloat* out;       // MODE 0: three spectra, `cstride` apart; MODE 1: one
  int64_t cstride;
  int M;             // columns = ny_loc * nzc
  KGrid g;
  float r2, norm;
  int deconv_order;
};

// MODE 0: out_j = IFFT_x[ -(i g_j) c FFT_x[in] ]            (force_spectra between the x-passes)
// MODE 1: out   = IFFT_x[ i c sum_j g_j FFT_x[in_j] ]        (force_spectra_T between the x-passes)
// PLAIN: no long-range filter and no deconvolution (the BullFrog loop): the scalar factor is a table look-up
// OCC: resident CTAs of 256 threads per SM the register budget is chosen for (2: 128 registers, 3: 80)
template <int R1, int R2, int MODE, bool PLAIN, int OCC>
__global__ void __launch_bounds__(R2 * CT, OCC * 256 / (R2 * CT)) xfuse_kernel(Args a) {
  constexpr int N = R1 * R2, J = R1 / R2;
  extern __shared__ float2 xsm[];
  float2* tw = xsm;                // [N]   exp(-2 pi i n / N)
  float2* xk = xsm + N;            // [N]   per-x kernel pieces: (lap_term(kx), grad_term(kx))
  float2* buf0 = xsm + 2 * N;      // [CT][N + 1] x 2: exchange buffers, alternating between transforms
  float2* buf1 = buf0 + CT * (N + 1);
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
  float2* cb0 = buf0 + c * (N + 1);
  float2* cb1 = buf1 + c * (N + 1);
  const int64_t M = a.M;
  // thread t touches x-planes t + R2*n: pointers to plane t of this column, strides in whole planes
  const float2* in = reinterpret_cast<const float2*>(a.in) + mm + t * M;
  float2* out = reinterpret_cast<float2*>(a.out) + mm + t * M;
  const int64_t MR2 = M * R2, MR1 = M * R1;

  // per-column wavevector pieces; position e = j*R2 + k2 of a thread's elements <-> kx index t + R2*j + R1*k2
  KVec k0;
  k0.kx = 0.f;
  k0.ky = a.g.ty * (float)signed_freq(jy, a.g.ny);
  k0.kz = a.g.tz * (float)l;
  const bool nqy = 2 * jy == a.g.ny, nqz = 2 * l == a.g.nz, sc = l == 0 || nqz;
  float gy = grad_term(k0.ky, a.g.grad_fd), gz = grad_term(k0.kz, a.g.grad_fd);
  if (sc && nqy) gy = 0.f;  // Hermitian projection on the self-conjugate planes (kspace.h: KVec)
  if (sc && nqz) gz = 0.f;
  auto kx_index = [&](int e) { return t + R2 * (e / R2) + R1 * (e % R2); };
  const float lapyz = lap_term(k0.ky, a.g.lap_fd) + lap_term(k0.kz, a.g.lap_fd);
  auto scalar_at = [&](int e) {  // invlaplace * [gaussian] * [1 / window^2] * norm at element e (kspace.h)
    if (PLAIN) {
      const float kk = xk[kx_index(e)].x + lapyz;
      return (kk == 0.0f ? 0.0f : -1.0f / kk) * a.norm;
    }
    KVec k = k0;
    k.kx = a.g.tx * (float)signed_freq(kx_index(e), a.g.nx);
    return force_scalar(a.g, k, a.r2, a.deconv_order) * a.norm;
  };
  auto gx_at = [&](int e) {
    const int i = kx_index(e);
    return (sc && 2 * i == a.g.nx) ? 0.f : xk[i].y;
  };

  float2 X[R1];
  if (MODE == 0) {
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) X[n1] = in[n1 * MR2];
    col_fft<R1, R2, -1>(X, t, cb0, tw);
#pragma unroll
    for (int e = 0; e < R1; ++e) {  // -(i g) c (re + i im) = g * [c (im - i re)]
      const float cs = scalar_at(e);
      X[e] = make_float2(X[e].y * cs, -X[e].x * cs);
    }
#pragma unroll 1
    for (int comp = 0; comp < 3; ++comp) {
      float2 w[R1];
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const float g = comp == 0 ? gx_at(e) : (comp == 1 ? gy : gz);
        w[(e / R2) + J * (e % R2)] = make_float2(X[e].x * g, X[e].y * g);  // inverse-transform input order
      }
      col_fft<R1, R2, +1>(w, t, (comp & 1) ? cb0 : cb1, tw);
      if (active) {
        float2* o = out + comp * a.cstride;
#pragma unroll
        for (int e = 0; e < R1; ++e) o[(e / R2) * MR2 + (e % R2) * MR1] = w[e];
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < R1; ++e) X[e] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int comp = 0; comp < 3; ++comp) {
      float2 v[R1];
      const float2* ip = in + comp * a.cstride;
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) v[n1] = ip[n1 * MR2];
      col_fft<R1, R2, -1>(v, t, (comp & 1) ? cb1 : cb0, tw);
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const float g = comp == 0 ? gx_at(e) : (comp == 1 ? gy : gz);
        X[e].x += g * v[e].x;
        X[e].y += g * v[e].y;
      }
    }
    float2 w[R1];
#pragma unroll
    for (int e = 0; e < R1; ++e) {  // i c (re + i im) = c (-im + i re)
      const float cs = scalar_at(e);
      w[(e / R2) + J * (e % R2)] = make_float2(-X[e].y * cs, X[e].x * cs);
    }
    col_fft<R1, R2, +1>(w, t, cb1, tw);
    if (active) {
#pragma unroll
      for (int e = 0; e < R1; ++e) out[(e / R2) * MR2 + (e % R2) * MR1] = w[e];
    }
  }
}

static int g_occ = 2;

template <int R1, int R2, int MODE, bool PLAIN, int OCC>
static void launch_one(stream_t st, const Args& a) {
  constexpr int N = R1 * R2;
  const size_t smem = sizeof(float2) * (2 * N + 2 * CT * (N + 1));
  cudaFuncSetAttribute(xfuse_kernel<R1, R2, MODE, PLAIN, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  xfuse_kernel<R1, R2, MODE, PLAIN, OCC><<<(a.M + CT - 1) / CT, R2 * CT, smem, st>>>(a);
}

template <int R1, int R2, int MODE>
static int launch(stream_t st, const Args& a) {
  count_launch();
  const bool plain = !(a.r2 > 0.f) && a.deconv_order <= 0;
  if (plain && g_occ == 3) launch_one<R1, R2, MODE, true, 3>(st, a);
  else if (plain) launch_one<R1, R2, MODE, true, 2>(st, a);
  else launch_one<R1, R2, MODE, false, 2>(st, a);
  return rt_check("xfuse") ? MCPM_ECUDA : 0;
}

template <int MODE>
static int dispatch(stream_t st, const Args& a) {
  switch (a.g.nx) {
    case 64: return launch<8, 8, MODE>(st, a);
    case 128: return launch<16, 8, MODE>(st, a);
    case 256: return launch<16, 16, MODE>(st, a);
  }
  set_error("xfuse: unsupported nx");
  return MCPM_EUNSUP;
}
}  // namespace xf

void set_xfuse_occ(int v) { xf::g_occ = v; }
bool xfuse_supported(int nx) { return nx == 64 || nx == 128 || nx == 256; }

static xf::Args make_args(const cfloat* in, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd, float kcut,
                          int deconv_order, float norm, SlabK sk) {
  xf::Args a;
  a.g = make_kgrid(nx, ny, nz, lap_fd, grad_fd, sk);
  a.in = in;
  a.out = out;
  a.M = a.g.ny_loc * a.g.nzc;
  a.cstride = (int64_t)nx * a.M;
  a.r2 = rcut2_half_of(kcut);
  a.norm = norm;
  a.deconv_order = deconv_order;
  return a;
}

// in: [nx, ny_loc, nzc] after the 2-D (y,z) R2C of every x-plane.  out3: three such arrays, inputs of the 2-D C2R.
int xfuse_force(stream_t st, const cfloat* in, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd,
                float kcut, int deconv_order, float norm, SlabK sk) {
  return xf::dispatch<0>(st, make_args(in, out3, nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk));
}

// in3: three arrays after the 2-D R2C.  out: one array, input of the 2-D C2R.
int xfuse_force_T(stream_t st, const cfloat* in3, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                  float kcut, int deconv_order, float norm, SlabK sk) {
  return xf::dispatch<1>(st, make_args(in3, out, nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk));
}

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
