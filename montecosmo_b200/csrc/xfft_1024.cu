// xfft_1024.cu -- instantiations of the fused x-transform for nx = 1024 (32 x 32), see xfft_large.cu.
#ifndef MCPM_HOSTEMU
#include "xfft_kernel.h"

namespace mcpm {

int xfuse_dispatch_1024(int mode, stream_t st, const xf::Args& a) { return xf::launch_mode<32, 32>(mode, st, a); }

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
