// xfft_large.cu -- instantiations of the fused x-transform for nx = 512 (32 x 16; 1024 = 32 x 32 lives in xfft_1024.cu):
// 32 elements per thread, 8 columns per CTA (xfft_kernel.h, design notes in xfft.cu).  Separate translation units only
// for build time.
#ifndef MCPM_HOSTEMU
#include "xfft_kernel.h"

namespace mcpm {

int xfuse_dispatch_1024(int mode, stream_t st, const xf::Args& a);  // xfft_1024.cu

int xfuse_dispatch_large(int mode, stream_t st, const xf::Args& a) {
  if (a.g.nx == 512) return xf::launch_mode<32, 16>(mode, st, a);
  if (a.g.nx == 1024) return xfuse_dispatch_1024(mode, st, a);
  set_error("xfuse: unsupported nx");
  return MCPM_EUNSUP;
}

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
