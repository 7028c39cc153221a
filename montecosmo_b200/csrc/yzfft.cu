// yzfft.cu -- the (y,z) part of rfftn / irfftn as ONE kernel per direction: real plane in, half-spectrum plane out (and
// back), both 1-D passes on-chip.  sm_100a build only; square planes ny = nz = N in {64, 128, 256}.
//
// Why: the step loop transforms ~64 meshes per gradient evaluation along (y,z) (1 R2C + 3 C2R per force evaluation, the
// transpose in the reverse step); cuFFT's batched 2-D plans run them as two kernels -- a z pass and a y pass, 52 + 72 us
// per 256^3 mesh, each a full HBM round trip (16N bytes per transform).  They were 29 % of the evaluation after round
// 1's x-pass fusion (profiles/r2_launches_*.txt).  Here the plane stays in shared memory between the passes: 8N bytes.
//
// One x-plane of the half spectrum is N x (N/2 + 1) complex64 = 264 KB at N = 256: more than one SM's shared memory
// (227 KB).  Each plane is therefore shared by TWO CTAs, and the FIRST pass is done by both:
//   R2C:  every CTA transforms all N rows along z (real rows packed in pairs into one complex transform), keeps the
//         kz columns of its half -- columns 0 and N/2 of a real input are real, so they travel as ONE packed complex
//         column and each half owns exactly N/4 columns -- then transforms its columns along y and stores them.
//   C2R:  every CTA transforms all N/2 columns along y (packed column included), keeps the N/2 rows of its half, then
//         transforms its row pairs along z (conjugate-symmetric spectrum rebuilt on the fly) and stores real rows.
// The duplicated pass costs arithmetic and a second read of the input that hits L2 (the twin CTA runs at the same time:
// blockIdx.x selects the half); HBM sees every byte once.  The 1-D transforms are the register FFTs of the x-transform
// kernel (xfft_kernel.h: radix-2 DIF inside a thread, one exchange through shared memory between two passes).
// Unnormalised in both directions, like cuFFT; the C2R assumes its input Hermitian-consistent on kz = 0 / Nyquist (the
// engine projects every spectrum that is not, fourier.cu: hermitian_project).
#ifndef MCPM_HOSTEMU
#include "xfft_kernel.h"

namespace mcpm {
namespace yz {

using xf::cmul;
using xf::fft_reg;

// Length-N transform of one ROW spread over R2 threads that are NEIGHBOURING LANES (t = lane % R2): same algorithm and
// element distribution as xf::col_fft, with an odd row stride R2 + 1 in the exchange buffer so that both the
// consecutive-t writes and the stride-(R2+1) reads are bank-conflict free.  In: v[m] = a[t + R2 m].  Out:
// v[j R2 + k2] = A[t + R2 j + R1 k2].  One __syncthreads.
template <int R1, int R2, int SIGN>
__device__ __forceinline__ void row_fft(float2 (&v)[R1], int t, float2* buf, const float2* tw) {
  constexpr int N = R1 * R2, J = R1 / R2, RS = R2 + 1;
  fft_reg<R1, SIGN>(v);
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) {
    float2 w = tw[(t * k1) & (N - 1)];
    if (SIGN > 0) w.y = -w.y;
    v[k1] = cmul(v[k1], w);
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) buf[k1 * RS + t] = v[k1];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) v[j * R2 + n2] = buf[(t + R2 * j) * RS + n2];
#pragma unroll
  for (int j = 0; j < J; ++j) fft_reg<R2, SIGN>(&v[j * R2]);
}

template <int R1, int R2>
struct Geo {
  static constexpr int N = R1 * R2;
  static constexpr int NZC = N / 2 + 1;
  static constexpr int THREADS = 16 * R2;      // 16 rows (row pairs) or 16 columns in flight per pass
  static constexpr int ROWBUF = R1 * (R2 + 1);  // float2 per row pair in a row_fft exchange buffer
  static constexpr int COLBUF = N + 1;          // float2 per column in a col_fft exchange buffer (xf::col_stride)
  static constexpr int XBUF = 16 * (ROWBUF > COLBUF ? ROWBUF : COLBUF);  // one exchange buffer, either use
  // forward: S[N rows][N/4 columns] pitch N/4 + 1 (odd: conflict free for the column reads); Z[16][N + 1] raw row spectra
  static constexpr int FW_PITCH = N / 4 + 1;
  static constexpr size_t FW_SMEM = sizeof(float2) * ((size_t)N * FW_PITCH + 2 * XBUF + N);
  // inverse: S[N/2 rows][N/2 columns] pitch N/2 + 1
  static constexpr int BW_PITCH = N / 2 + 1;
  static constexpr size_t BW_SMEM = sizeof(float2) * ((size_t)(N / 2) * BW_PITCH + 2 * XBUF + N);
};

__device__ __forceinline__ void fill_twiddles(float2* tw, int n, int nthreads) {
  for (int i = threadIdx.x; i < n; i += nthreads) {
    float sn, cs;
    sincospif(-2.0f * (float)i / (float)n, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
}

// in: real planes [nplanes][N][N]; out: complex planes [nplanes][N][N/2 + 1].  grid (2, nplanes).
template <int R1, int R2>
__global__ void __launch_bounds__(16 * R2, 1) r2c_yz_kernel(const float* __restrict__ in, float2* __restrict__ out) {
  using G = Geo<R1, R2>;
  constexpr int N = G::N, NZC = G::NZC, Q = N / 4, P = G::FW_PITCH;
  extern __shared__ float2 ysm[];
  float2* S = ysm;                 // [N][P]
  float2* xb0 = S + (size_t)N * P;  // exchange buffers / raw row spectra, alternating
  float2* xb1 = xb0 + G::XBUF;
  float2* tw = xb1 + G::XBUF;
  const int half = blockIdx.x;
  const float* pin = in + (size_t)blockIdx.y * N * N;
  float2* pout = out + (size_t)blockIdx.y * N * NZC;
  fill_twiddles(tw, N, G::THREADS);
  __syncthreads();

  // ---- pass 1: all rows along z, two real rows per complex transform; keep this half's columns
  {
    const int t = threadIdx.x % R2, g = threadIdx.x / R2;  // 16 row pairs per sweep
    const int klo = half ? Q : 0, khi = half ? 2 * Q : Q;   // regular columns k in [klo, khi); half 0 also packs 0 & N/2
    // the next sweep's rows are loaded before the current sweep's transform (one CTA per SM: nothing else hides the latency)
    float2 nv[R1];
    {
      const float* pa = pin + (size_t)(2 * g) * N + t;
#pragma unroll
      for (int m = 0; m < R1; ++m) nv[m] = make_float2(pa[R2 * m], pa[N + R2 * m]);
    }
#pragma unroll 1
    for (int sweep = 0; sweep < N / 32; ++sweep) {
      const int ra = 2 * (sweep * 16 + g);
      float2 v[R1];
#pragma unroll
      for (int m = 0; m < R1; ++m) v[m] = nv[m];
      if (sweep + 1 < N / 32) {
        const float* pa = pin + (size_t)(ra + 32) * N + t;
#pragma unroll
        for (int m = 0; m < R1; ++m) nv[m] = make_float2(pa[R2 * m], pa[N + R2 * m]);
      }
      row_fft<R1, R2, -1>(v, t, xb0 + g * G::ROWBUF, tw);
      float2* Z = xb1 + g * (N + 1);  // raw spectrum C[k] of the pair, for the partner look-up C[N - k]
#pragma unroll
      for (int e = 0; e < R1; ++e) Z[t + R2 * (e / R2) + R1 * (e % R2)] = v[e];
      __syncthreads();
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const int k = t + R2 * (e / R2) + R1 * (e % R2);
        if (k >= klo && k < khi && k != 0) {
          const float2 c = v[e], p = Z[N - k];
          // A_k = (C_k + conj C_{N-k}) / 2 ;  B_k = -i (C_k - conj C_{N-k}) / 2
          S[(size_t)ra * P + (k - klo)] = make_float2(0.5f * (c.x + p.x), 0.5f * (c.y - p.y));
          S[(size_t)(ra + 1) * P + (k - klo)] = make_float2(0.5f * (c.y + p.y), -0.5f * (c.x - p.x));
        }
      }
      if (half == 0 && t == 0) {  // packed column: (A_0, A_{N/2}) are real; B likewise.  C_0 = A_0 + i B_0 etc.
        const float2 c0 = Z[0], cn = Z[N / 2];
        S[(size_t)ra * P] = make_float2(c0.x, cn.x);
        S[(size_t)(ra + 1) * P] = make_float2(c0.y, cn.y);
      }
      __syncthreads();  // Z and the exchange buffer are reused by the next sweep
    }
  }
  // ---- pass 2: this half's columns along y, straight to global memory
  {
    const int c = threadIdx.x % 16, t = threadIdx.x / 16;
#pragma unroll 1
    for (int blk = 0; blk < Q / 16; ++blk) {
      const int col = blk * 16 + c;
      float2 v[R1];
#pragma unroll
      for (int m = 0; m < R1; ++m) v[m] = S[(size_t)(t + R2 * m) * P + col];
      float2* xb = (blk & 1) ? xb1 : xb0;
      xf::col_fft<R1, R2, -1>(v, t, xb + c * G::COLBUF, tw);
      const bool packed = half == 0 && col == 0;
      if (!packed) {
        const int kz = (half ? Q : 0) + col;
#pragma unroll
        for (int e = 0; e < R1; ++e) pout[(size_t)(t + R2 * (e / R2) + R1 * (e % R2)) * NZC + kz] = v[e];
      }
      if (half == 0 && blk == 0) {  // unpack the packed column: G = FFT(a0 + i aN); Y0 = (G_k + conj G_-k)/2, YN = -i(..-..)/2
        float2* Gs = ((blk & 1) ? xb0 : xb1);  // the other exchange buffer is idle during this block
        if (c == 0) {
#pragma unroll
          for (int e = 0; e < R1; ++e) Gs[t + R2 * (e / R2) + R1 * (e % R2)] = v[e];
        }
        __syncthreads();
        if (c == 0) {
#pragma unroll
          for (int e = 0; e < R1; ++e) {
            const int ky = t + R2 * (e / R2) + R1 * (e % R2);
            const float2 a = v[e], p = Gs[(N - ky) & (N - 1)];
            pout[(size_t)ky * NZC] = make_float2(0.5f * (a.x + p.x), 0.5f * (a.y - p.y));
            pout[(size_t)ky * NZC + N / 2] = make_float2(0.5f * (a.y + p.y), -0.5f * (a.x - p.x));
          }
        }
        __syncthreads();  // Gs is the next block's exchange buffer
      }
    }
  }
}

// in: complex planes [nplanes][N][N/2 + 1] (Hermitian-consistent on kz = 0, N/2); out: real planes [nplanes][N][N].
// grid (2, nplanes): CTA `half` produces rows [half N/2, (half + 1) N/2).
template <int R1, int R2>
__global__ void __launch_bounds__(16 * R2, 1) c2r_yz_kernel(const float2* __restrict__ in, float* __restrict__ out) {
  using G = Geo<R1, R2>;
  constexpr int N = G::N, NZC = G::NZC, H = N / 2, P = G::BW_PITCH;
  extern __shared__ float2 ysm[];
  float2* S = ysm;                  // [N/2 rows][P]: column q = kz for 1 <= q < N/2, column 0 = packed (kz 0, kz N/2)
  float2* xb0 = S + (size_t)H * P;
  float2* xb1 = xb0 + G::XBUF;
  float2* tw = xb1 + G::XBUF;
  const int half = blockIdx.x;
  const float2* pin = in + (size_t)blockIdx.y * N * NZC;
  float* pout = out + (size_t)blockIdx.y * N * N;
  fill_twiddles(tw, N, G::THREADS);
  __syncthreads();

  // ---- pass 1: all N/2 columns (packed one included) along y; keep this half's rows
  {
    const int c = threadIdx.x % 16, t = threadIdx.x / 16;
    const int ylo = half * H;
    auto load_col = [&](int q, float2* v) {
      if (q == 0) {  // G_ky = Y0_ky + i YN_ky
#pragma unroll
        for (int m = 0; m < R1; ++m) {
          const float2 a = pin[(size_t)(t + R2 * m) * NZC], b = pin[(size_t)(t + R2 * m) * NZC + H];
          v[m] = make_float2(a.x - b.y, a.y + b.x);
        }
      } else {
#pragma unroll
        for (int m = 0; m < R1; ++m) v[m] = pin[(size_t)(t + R2 * m) * NZC + q];
      }
    };
    float2 nv[R1];  // next block's columns, in flight under the current block's transform
    load_col(c, nv);
#pragma unroll 1
    for (int blk = 0; blk < H / 16; ++blk) {
      const int q = blk * 16 + c;
      float2 v[R1];
#pragma unroll
      for (int m = 0; m < R1; ++m) v[m] = nv[m];
      if (blk + 1 < H / 16) load_col(q + 16, nv);
      xf::col_fft<R1, R2, +1>(v, t, ((blk & 1) ? xb1 : xb0) + c * G::COLBUF, tw);
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const int y = t + R2 * (e / R2) + R1 * (e % R2);
        if (y >= ylo && y < ylo + H) S[(size_t)(y - ylo) * P + q] = v[e];
      }
    }
  }
  __syncthreads();
  // ---- pass 2: this half's row pairs along z: C_k = A_k + i B_k, with A_{N-k} = conj A_k for the upper half
  {
    const int t = threadIdx.x % R2, g = threadIdx.x / R2;
#pragma unroll 1
    for (int sweep = 0; sweep < H / 32; ++sweep) {
      const int la = 2 * (sweep * 16 + g);
      const float2* Sa = S + (size_t)la * P;
      const float2* Sb = Sa + P;
      float2 v[R1];
#pragma unroll
      for (int m = 0; m < R1; ++m) {
        const int k = t + R2 * m;
        if (k == 0) {
          v[m] = make_float2(Sa[0].x, Sb[0].x);
        } else if (k == H) {
          v[m] = make_float2(Sa[0].y, Sb[0].y);
        } else if (k < H) {
          const float2 a = Sa[k], b = Sb[k];
          v[m] = make_float2(a.x - b.y, a.y + b.x);
        } else {
          const float2 a = Sa[N - k], b = Sb[N - k];
          v[m] = make_float2(a.x + b.y, b.x - a.y);
        }
      }
      row_fft<R1, R2, +1>(v, t, ((sweep & 1) ? xb1 : xb0) + g * G::ROWBUF, tw);
      float* oa = pout + (size_t)(half * H + la) * N;
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const int n = t + R2 * (e / R2) + R1 * (e % R2);
        oa[n] = v[e].x;
        oa[N + n] = v[e].y;
      }
    }
  }
}

template <int R1, int R2>
static int launch_r2c(stream_t st, const float* in, cfloat* out, int nplanes) {
  using G = Geo<R1, R2>;
  cudaFuncSetAttribute(r2c_yz_kernel<R1, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::FW_SMEM);
  count_launch();
  r2c_yz_kernel<R1, R2><<<dim3(2, nplanes), G::THREADS, G::FW_SMEM, st>>>(in, reinterpret_cast<float2*>(out));
  return rt_check("r2c_yz") ? MCPM_ECUDA : 0;
}
template <int R1, int R2>
static int launch_c2r(stream_t st, const cfloat* in, float* out, int nplanes) {
  using G = Geo<R1, R2>;
  cudaFuncSetAttribute(c2r_yz_kernel<R1, R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BW_SMEM);
  count_launch();
  c2r_yz_kernel<R1, R2><<<dim3(2, nplanes), G::THREADS, G::BW_SMEM, st>>>(reinterpret_cast<const float2*>(in), out);
  return rt_check("c2r_yz") ? MCPM_ECUDA : 0;
}
}  // namespace yz

bool yzfft_supported(int ny, int nz) { return ny == nz && (ny == 64 || ny == 128 || ny == 256); }

// nplanes = meshes x local x-planes; returns MCPM_EUNSUP for other shapes (the caller keeps cuFFT)
int yzfft_r2c(stream_t st, const float* in, cfloat* out, int ny, int nz, int nplanes) {
  if (!yzfft_supported(ny, nz) || nplanes <= 0 || nplanes > 65535) return MCPM_EUNSUP;
  switch (ny) {
    case 64: return yz::launch_r2c<8, 8>(st, in, out, nplanes);
    case 128: return yz::launch_r2c<16, 8>(st, in, out, nplanes);
    default: return yz::launch_r2c<16, 16>(st, in, out, nplanes);
  }
}
int yzfft_c2r(stream_t st, const cfloat* in, float* out, int ny, int nz, int nplanes) {
  if (!yzfft_supported(ny, nz) || nplanes <= 0 || nplanes > 65535) return MCPM_EUNSUP;
  switch (ny) {
    case 64: return yz::launch_c2r<8, 8>(st, in, out, nplanes);
    case 128: return yz::launch_c2r<16, 8>(st, in, out, nplanes);
    default: return yz::launch_c2r<16, 16>(st, in, out, nplanes);
  }
}

}  // namespace mcpm
#endif
