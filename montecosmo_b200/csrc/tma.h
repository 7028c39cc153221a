// tma.h -- the few Blackwell/Hopper async-copy primitives the particle kernels use, as inline PTX (sm_100a build only):
// 1-D bulk copies between global and shared memory (cp.async.bulk, the descriptor-less form of TMA) completing on an
// mbarrier, and bulk stores / reductions from shared memory tracked by bulk groups.
//
// Why: particle arrays are [np, 3] float32 AoS at the ABI (the reference's layout).  A warp's three 4-byte loads of 32
// consecutive particles touch 12 sectors per request and its three stores write 12 partial sectors; ncu showed 12
// sectors / request, 1.2x DRAM over-fetch and the LSU pipe 70 % busy on every particle kernel (profiles/r1_ncu_v3).  A
// 32-particle row is 384 contiguous, 16-byte aligned bytes: one bulk copy moves it without occupying the LSU or any
// register, asynchronously (the next row's copy runs under the current row's arithmetic), and shared memory is then
// read with a conflict-free stride of 3 words.
#pragma once
#ifndef MCPM_HOSTEMU
#include <cstdint>
#include <cuda_runtime.h>

namespace mcpm {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (the copy engine) before anything is issued on them
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one arrival + `bytes` expected from the copies issued next
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for the phase of the given parity.  try_wait suspends the thread in hardware for a bounded time per call; the
// spin count is bounded too, so that a lost copy traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global as part of the calling thread's current bulk group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
// shared -> global float32 ADD (an L2 reduction per element), same tracking
__device__ __forceinline__ void bulk_red_add_f32(float* gmem_dst, const float* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the most recent N groups of this thread have finished READING shared memory (their source may be reused)
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order this thread's generic-proxy writes to shared memory before later async-proxy reads of them (bulk stores)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tma
}  // namespace mcpm
#endif
