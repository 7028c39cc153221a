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
#include "xfft_kernel.h"

namespace mcpm {

namespace xf {
static int dispatch(int mode, stream_t st, const Args& a) {
  switch (a.g.nx) {
    case 64: return launch_mode<8, 8>(mode, st, a);
    case 128: return launch_mode<16, 8>(mode, st, a);
    case 256: return launch_mode<16, 16>(mode, st, a);
    case 512:
    case 1024: return xfuse_dispatch_large(mode, st, a);
  }
  set_error("xfuse: unsupported nx");
  return MCPM_EUNSUP;
}
}  // namespace xf

bool xfuse_supported(int nx) { return nx == 64 || nx == 128 || nx == 256 || nx == 512 || nx == 1024; }

static xf::Args make_args(const cfloat* in, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd, float kcut,
                          int deconv_order, float norm, SlabK sk, int half_weights = 0, int accumulate = 0) {
  xf::Args a;
  a.g = make_kgrid(nx, ny, nz, lap_fd, grad_fd, sk);
  a.in = in;
  a.out = out;
  a.M = a.g.ny_loc * a.g.nzc;
  a.cstride = (int64_t)nx * a.M;
  a.r2 = rcut2_half_of(kcut);
  a.norm = norm;
  a.deconv_order = deconv_order;
  a.half_weights = half_weights;
  a.accumulate = accumulate;
  a.npeer = 0;
  a.xl = nx;
  for (int r = 0; r < 8; ++r) {
    a.in_peer[r] = nullptr;
    a.out_peer[r] = nullptr;
  }
  return a;
}

// Peer-memory variant of xfuse_force (transpose = 0: 1 -> 3) and xfuse_force_T (transpose = 1: 3 -> 1) for a mesh
// slab-decomposed over `npeer` ranks: in_peers[r] / out_peers[r] are rank r's buffers [ncomp][nx/npeer][ny][nzc]
// (after / before the local 2-D (y,z) transforms), mapped into this process; this rank handles ky rows y0 .. y0+ny_loc-1.
int xfuse_force_peer(stream_t st, const cfloat* const* in_peers, cfloat* const* out_peers, int npeer, int transpose,
                     int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, float kcut, int deconv_order,
                     float norm) {
  if (npeer < 1 || npeer > 8 || nx % npeer) {
    set_error("xfuse_force_peer: 1..8 ranks, nx divisible by their number");
    return MCPM_EINVAL;
  }
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  xf::Args a = make_args(in_peers[0], out_peers[0], nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk);
  a.npeer = npeer;
  a.xl = nx / npeer;
  for (int r = 0; r < npeer; ++r) {
    a.in_peer[r] = in_peers[r];
    a.out_peer[r] = out_peers[r];
  }
  // transpose: 0 force (1 -> 3), 1 its transpose (3 -> 1), 2 / 3 the two-field forms (xfft_kernel.h: FORCE2, FORCE2_T)
  const int mode = transpose == 0 ? xf::FORCE : transpose == 1 ? xf::FORCE_T : transpose == 2 ? xf::FORCE2 : xf::FORCE2_T;
  return xf::dispatch(mode, st, a);
}

// Any of the six operators (xfft_kernel.h: Mode) with its x-space side in peer memory: modes whose INPUT is x-space
// (FORCE, FORCE_T, FORCE_TK, HESS_TK) load every x-plane from in_peers[owner]; modes whose OUTPUT is x-space (FORCE,
// FORCE_T, FORCE_K, HESS_K) store every plane at out_peers[owner]; the k-space side (a 3-D spectrum, this rank's ky block
// [nx, ny_loc, nzc]) is local: `k_local`.  This is what makes the slab-decomposed lpt / lpt_vjp free of all-to-alls.
int xfuse_peer_mode(stream_t st, int mode, const cfloat* const* in_peers, cfloat* const* out_peers, cfloat* k_local,
                    int npeer, int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, int half_weights,
                    int accumulate, float norm) {
  if (npeer < 1 || npeer > 8 || nx % npeer || mode < 0 || mode > xf::FORCE2_T) {
    set_error("xfuse_peer_mode: 1..8 ranks, nx divisible by their number, mode 0..7");
    return MCPM_EINVAL;
  }
  const bool k_in = mode == xf::FORCE_K || mode == xf::HESS_K, k_out = mode == xf::FORCE_TK || mode == xf::HESS_TK;
  if ((k_in || k_out) && !k_local) {
    set_error("xfuse_peer_mode: this mode needs the local spectrum");
    return MCPM_EINVAL;
  }
  SlabK sk;
  sk.ny_loc = ny_loc;
  sk.y0 = y0;
  xf::Args a = make_args(k_in ? k_local : in_peers[0], k_out ? k_local : out_peers[0], nx, ny, nz, lap_fd, grad_fd, 0.0f, 0,
                         norm, sk, half_weights, accumulate);
  a.npeer = npeer;
  a.xl = nx / npeer;
  for (int r = 0; r < npeer; ++r) {
    a.in_peer[r] = in_peers ? in_peers[r] : nullptr;
    a.out_peer[r] = out_peers ? out_peers[r] : nullptr;
  }
  return xf::dispatch(mode, st, a);
}

// in: [nx, ny_loc, nzc] after the 2-D (y,z) R2C of every x-plane.  out3: three such arrays, inputs of the 2-D C2R.
int xfuse_force(stream_t st, const cfloat* in, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd,
                float kcut, int deconv_order, float norm, SlabK sk) {
  return xf::dispatch(xf::FORCE, st, make_args(in, out3, nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk));
}

// in3: three arrays after the 2-D R2C.  out: one array, input of the 2-D C2R.
int xfuse_force_T(stream_t st, const cfloat* in3, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                  float kcut, int deconv_order, float norm, SlabK sk) {
  return xf::dispatch(xf::FORCE_T, st, make_args(in3, out, nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk));
}

// dk: a full 3-D half spectrum.  out3 / out6: inputs of the 2-D C2R (force components / Hessian set 00 11 22 01 02 12).
int xfuse_force_k(stream_t st, const cfloat* dk, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd,
                  float kcut, int deconv_order, float norm, SlabK sk) {
  return xf::dispatch(xf::FORCE_K, st, make_args(dk, out3, nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk));
}
int xfuse_hessian_k(stream_t st, const cfloat* dk, cfloat* out6, int nx, int ny, int nz, int lap_fd, int grad_fd,
                    float norm, SlabK sk) {
  return xf::dispatch(xf::HESS_K, st, make_args(dk, out6, nx, ny, nz, lap_fd, grad_fd, 0.0f, 0, norm, sk));
}

// in3 / in6: arrays after the 2-D R2C.  out: a full 3-D half spectrum (cotangent of delta_k), optionally accumulated.
int xfuse_force_tk(stream_t st, const cfloat* in3, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                   float kcut, int deconv_order, int half_weights, int accumulate, float norm, SlabK sk) {
  return xf::dispatch(xf::FORCE_TK, st, make_args(in3, out, nx, ny, nz, lap_fd, grad_fd, kcut, deconv_order, norm, sk,
                                                  half_weights, accumulate));
}
int xfuse_hessian_tk(stream_t st, const cfloat* in6, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                     int half_weights, int accumulate, float norm, SlabK sk) {
  return xf::dispatch(xf::HESS_TK, st, make_args(in6, out, nx, ny, nz, lap_fd, grad_fd, 0.0f, 0, norm, sk, half_weights,
                                                 accumulate));
}

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
