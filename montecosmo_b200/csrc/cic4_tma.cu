// cic4_tma.cu -- the two CIC gather kernels of the BullFrog step (cic4.cu: kick_drift4, read_grad4v) with the particle
// arrays staged through shared memory by bulk async copies (tma.h).  sm_100a build only; same arithmetic, in the same
// association order, as the one-thread-per-particle kernels of cic4.cu, so results are bit-identical to them.
//
// Round 1's profile of those kernels (profiles/r1_ncu_v3_kernels.txt): DRAM traffic 1.20-1.24x the algorithmic bytes,
// l1tex 72 % busy, long_scoreboard 20 cycles per issue at 50 % occupancy.  A thread loaded its position with three 4-byte
// loads at a 12-byte stride (12 sectors per warp request), waited ~1 us, issued its 8 mesh loads, waited again, and
// stored six more strided scalars.  Here a warp works on SEGMENTS of SEG consecutive particles:
//   * lane 0 issues one cp.async.bulk per array for the NEXT segment (SEG * 12 contiguous, 16-byte aligned bytes) onto
//     the warp's mbarrier, then the warp waits for the CURRENT segment's bytes: the copy latency hides under the mesh
//     gather of the segment before; no LSU instruction and no register is spent on the particle arrays' global side;
//   * positions / velocities / cotangents are read from shared memory at a stride of 3 words (conflict-free);
//   * results go back through shared memory and one bulk store per array (full 32-byte sectors, asynchronous);
//   * the mesh is read as before, 8 x LDG.128 per particle through L1: the segments of a CTA are consecutive in the
//     particle order (Lagrangian order, z fastest), so the CTA's gathers stay inside a few mesh rows, and CTAs are
//     scheduled in particle order -- one front sweeps the mesh, every mesh line is fetched from HBM about once.
// The CTAs are PERSISTENT (one grid-sized wave; warp w of the grid takes segments w, w + W, w + 2W, ...): the first version
// gave every CTA four segments per warp and measured 0.52 ms where the kernel it replaced took 0.29 -- each short-lived
// CTA paid the pipeline fill (barrier set-up, first copy latency) and the drain of its last stores.  At any instant the
// grid works on W consecutive segments (a slab of ~2 lattice x-planes at 256^3), so the mesh footprint in flight stays
// inside L2.  Shared memory is kept small (SEG = 32: 0.4 KB per array and stage) because it is carved out of L1, which
// the mesh gathers need: 64- and 128-particle segments measured 1.5 and 7 ms slower per evaluation than 32.
// Works for absolute positions in any particle order, and for lattice-relative positions (frame.h) of spacing one cell
// whose pencils hold a whole number of segments; anything else takes the kernels of cic4.cu.
#ifndef MCPM_HOSTEMU
#include "engine.h"
#include "tma.h"
#include "window.h"

namespace mcpm {

namespace gtma {
constexpr int WARPS = 8;   // per CTA
constexpr int THREADS = WARPS * 32;
constexpr int NSTAGE = 3;  // input ring per warp: two segments in flight under the one being worked on
}  // namespace gtma

struct GatherArgs {
  const float* in[3];  // kick: pos, vel, -      grad: pos, cot, grad (grad only read when accumulate)
  float* out[2];       // kick: pos_out, vel_out  grad: cot, grad
  const float4* fm4;
  const float* rhobar;
  int nx, ny, nz;
  int rel, py, pz, ox, oy, oz;  // relative frame of spacing one cell (lattice py x pz pencils), else rel = 0
  // Brick-ordered work assignment (lat != 0: the particle order is a known px x py x pz lattice, relative or not): the 8
  // warps of a CTA take the 8 segments of a 2 (x) x 4 (y) patch of lattice rows at one z-segment, so every mesh line is
  // read by up to 8 rows of the same CTA at the same time instead of 2 (L1 hit rate ~50 % -> ~75 %).  spp = segments
  // per pencil, tiles = (px/2) * (py/4) * spp.  lat == 0: segments in linear order (any particle order).
  int lat, lpy, spp;
  unsigned ntiles;
  float alpha, beta, drift;     // kick
  float cscale, alpha_tail, dnext;  // grad
  int accumulate, scale_cot;
  int64_t nseg;
};

struct CicIdx {
  int i0, i1, j0, j1, k0, k1;
  float fx, fy, fz;
};

// the same split as cic4.cu's cic_setup: u = (remainder 0) + x, base = floor(u), cell = wrap(site + base)
__device__ __forceinline__ CicIdx cic_index(float x0, float x1, float x2, int sx, int sy, int sz, int nx, int ny, int nz) {
  CicIdx c;
  const float u0 = 0.0f + x0, u1 = 0.0f + x1, u2 = 0.0f + x2;
  const float bx = floorf(u0), by = floorf(u1), bz = floorf(u2);
  c.fx = u0 - bx;
  c.fy = u1 - by;
  c.fz = u2 - bz;
  c.i0 = wrap_fast(sx + (int)bx, nx);
  c.j0 = wrap_fast(sy + (int)by, ny);
  c.k0 = wrap_fast(sz + (int)bz, nz);
  c.i1 = c.i0 + 1 == nx ? 0 : c.i0 + 1;
  c.j1 = c.j0 + 1 == ny ? 0 : c.j0 + 1;
  c.k1 = c.k0 + 1 == nz ? 0 : c.k0 + 1;
  return c;
}

// MODE 0: kick_drift4.  MODE 1: read_grad4v.
template <int MODE, int SEG>
__global__ void __launch_bounds__(gtma::THREADS, MODE == 0 ? 4 : 3) gather_tma_kernel(GatherArgs a) {
  using namespace gtma;
  constexpr int NIN = MODE == 0 ? 2 : 3, NOUT = 2;
  constexpr int ROW = SEG * 3;                 // floats per array per segment
  constexpr uint32_t BYTES = ROW * sizeof(float);
  constexpr int SUB = SEG / 32;
  static_assert(SEG % 32 == 0 && BYTES % 16 == 0, "a segment is whole warps and whole 16-byte units");
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[WARPS][NSTAGE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sin = smem + (size_t)warp * ((NSTAGE * NIN + 2 * NOUT) * ROW);  // [NSTAGE][NIN][ROW]
  float* sout = sin + NSTAGE * NIN * ROW;                                // [2][NOUT][ROW]
  uint64_t* bar = bars[warp];
  if (lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) tma::mbar_init(&bar[i], 1);
    tma::fence_barrier_init();
  }
  __syncwarp();

  const int nload = (MODE == 1 && !a.accumulate) ? 2 : NIN;
  // 32-bit index arithmetic throughout: nseg < 2^31 (checked by the host side)
  const unsigned ntile = a.lat ? a.ntiles : (unsigned)((a.nseg + WARPS - 1) / WARPS);  // units of 8 segments
  const int nmine = blockIdx.x < ntile ? (int)((ntile - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const unsigned hj = (unsigned)a.lpy >> 2;  // y-patches per x-pair (lat mode)
  auto seg_of = [&](int t, unsigned& seg, unsigned& li, unsigned& lj, unsigned& ks) -> bool {
    const unsigned T = blockIdx.x + (unsigned)t * gridDim.x;
    if (a.lat) {  // tile T = (i2 * hj + j4) * spp + ks; warp (di, dj) = (warp >> 2, warp & 3)
      const unsigned ij = T / (unsigned)a.spp;
      ks = T - ij * (unsigned)a.spp;
      const unsigned i2 = ij / hj, j4 = ij - i2 * hj;
      li = 2 * i2 + (warp >> 2);
      lj = 4 * j4 + (warp & 3);
      seg = (li * (unsigned)a.lpy + lj) * (unsigned)a.spp + ks;
      return true;
    }
    seg = T * WARPS + warp;
    li = lj = ks = 0;
    return seg < (unsigned)a.nseg;
  };
  auto issue = [&](int t) {  // lane 0: bulk loads of this warp's t-th segment into stage t % NSTAGE
    unsigned seg, li, lj, ks;
    if (!seg_of(t, seg, li, lj, ks)) return;
    const int st = t % NSTAGE;
    tma::mbar_arrive_expect_tx(&bar[st], BYTES * nload);
    for (int m = 0; m < nload; ++m)
      tma::bulk_g2s(sin + (st * NIN + m) * ROW, a.in[m] + (size_t)seg * ROW, BYTES, &bar[st]);
  };
  if (lane == 0)
    for (int t = 0; t < NSTAGE - 1 && t < nmine; ++t) issue(t);

  const float4* __restrict__ fm = a.fm4;
  const int nx = a.nx, ny = a.ny, nz = a.nz;
  for (int t = 0; t < nmine; ++t) {
    const int st = t % NSTAGE;
    // stage (t + NSTAGE - 1) % NSTAGE was last read in iteration t - 1 (syncwarp since)
    if (lane == 0 && t + NSTAGE - 1 < nmine) issue(t + NSTAGE - 1);
    unsigned seg, li, lj, ks;
    if (!seg_of(t, seg, li, lj, ks)) break;  // linear mode: the last unit of 8 segments may be partial (warp-uniform)
    tma::mbar_wait(&bar[st], (uint32_t)(t / NSTAGE) & 1u);
    const size_t p0 = (size_t)seg * SEG;
    int sx = 0, sy = 0, sk = 0;
    if (a.rel) {  // lattice site of the segment's first particle: one pencil holds whole segments
      if (!a.lat) {
        const unsigned spp = (unsigned)a.pz / SEG, jk = seg / spp;
        ks = seg - jk * spp;
        li = jk / (unsigned)a.py;
        lj = jk - li * (unsigned)a.py;
      }
      sx = (int)li + a.ox;
      sy = (int)lj + a.oy;
      sk = (int)(ks * SEG) + a.oz;
    }
    const float* spos = sin + (st * NIN + 0) * ROW;
    const float* sb = sin + (st * NIN + 1) * ROW;
    const float* sc = sin + (st * NIN + (NIN - 1)) * ROW;
    if (lane == 0) tma::bulk_wait_read<1>();  // the stores of iteration t-2 have released sout[t & 1]
    __syncwarp();
    float* o0 = sout + ((t & 1) * NOUT + 0) * ROW;
    float* o1 = sout + ((t & 1) * NOUT + 1) * ROW;
#pragma unroll
    for (int s = 0; s < SUB; ++s) {
      const int q = (s * 32 + lane) * 3;
      const float x0 = spos[q], x1 = spos[q + 1], x2 = spos[q + 2];
      const CicIdx c = cic_index(x0, x1, x2, sx, sy, a.rel ? sk + s * 32 + lane : 0, nx, ny, nz);
      const int64_t r00 = ((int64_t)c.i0 * ny + c.j0) * nz, r01 = ((int64_t)c.i0 * ny + c.j1) * nz;
      const int64_t r10 = ((int64_t)c.i1 * ny + c.j0) * nz, r11 = ((int64_t)c.i1 * ny + c.j1) * nz;
      if (MODE == 0) {
        const float gx = 1.0f - c.fx, gy = 1.0f - c.fy, gz = 1.0f - c.fz;
        const float w00 = gx * gy, w01 = gx * c.fy, w10 = c.fx * gy, w11 = c.fx * c.fy;
        const float4 a0 = __ldg(fm + r00 + c.k0), a1 = __ldg(fm + r00 + c.k1);
        const float4 b0 = __ldg(fm + r01 + c.k0), b1 = __ldg(fm + r01 + c.k1);
        const float4 d0 = __ldg(fm + r10 + c.k0), d1 = __ldg(fm + r10 + c.k1);
        const float4 e0 = __ldg(fm + r11 + c.k0), e1 = __ldg(fm + r11 + c.k1);
        const float f0 = a0.x * (w00 * gz) + a1.x * (w00 * c.fz) + b0.x * (w01 * gz) + b1.x * (w01 * c.fz) +
                         d0.x * (w10 * gz) + d1.x * (w10 * c.fz) + e0.x * (w11 * gz) + e1.x * (w11 * c.fz);
        const float f1 = a0.y * (w00 * gz) + a1.y * (w00 * c.fz) + b0.y * (w01 * gz) + b1.y * (w01 * c.fz) +
                         d0.y * (w10 * gz) + d1.y * (w10 * c.fz) + e0.y * (w11 * gz) + e1.y * (w11 * c.fz);
        const float f2 = a0.z * (w00 * gz) + a1.z * (w00 * c.fz) + b0.z * (w01 * gz) + b1.z * (w01 * c.fz) +
                         d0.z * (w10 * gz) + d1.z * (w10 * c.fz) + e0.z * (w11 * gz) + e1.z * (w11 * c.fz);
        const float v0 = a.alpha * sb[q] + a.beta * f0;
        const float v1 = a.alpha * sb[q + 1] + a.beta * f1;
        const float v2 = a.alpha * sb[q + 2] + a.beta * f2;
        o1[q] = v0;
        o1[q + 1] = v1;
        o1[q + 2] = v2;
        o0[q] = x0 + v0 * a.drift;
        o0[q + 1] = x1 + v1 * a.drift;
        o0[q + 2] = x2 + v2 * a.drift;
      } else {
        const float* __restrict__ rb = a.rhobar;
        const float q0 = sb[q], q1 = sb[q + 1], q2 = sb[q + 2];
        const float c0 = a.cscale * q0, c1 = a.cscale * q1, c2 = a.cscale * q2;
        float4 tt;
        float u000, u001, u010, u011, u100, u101, u110, u111;
        tt = __ldg(fm + r00 + c.k0); u000 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r00 + c.k0);
        tt = __ldg(fm + r00 + c.k1); u001 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r00 + c.k1);
        tt = __ldg(fm + r01 + c.k0); u010 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r01 + c.k0);
        tt = __ldg(fm + r01 + c.k1); u011 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r01 + c.k1);
        tt = __ldg(fm + r10 + c.k0); u100 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r10 + c.k0);
        tt = __ldg(fm + r10 + c.k1); u101 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r10 + c.k1);
        tt = __ldg(fm + r11 + c.k0); u110 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r11 + c.k0);
        tt = __ldg(fm + r11 + c.k1); u111 = c0 * tt.x + c1 * tt.y + c2 * tt.z + __ldg(rb + r11 + c.k1);
        const float wx0 = 1.0f - c.fx, wx1 = c.fx, wy0 = 1.0f - c.fy, wy1 = c.fy, wz0 = 1.0f - c.fz, wz1 = c.fz;
        const float dx0 = c.fx > 0.0f ? -1.0f : 0.0f, dy0 = c.fy > 0.0f ? -1.0f : 0.0f, dz0 = c.fz > 0.0f ? -1.0f : 0.0f;
        const float g0 = (dx0 * u000 + u100) * (wy0 * wz0) + (dx0 * u001 + u101) * (wy0 * wz1) +
                         (dx0 * u010 + u110) * (wy1 * wz0) + (dx0 * u011 + u111) * (wy1 * wz1);
        const float g1 = (dy0 * u000 + u010) * (wx0 * wz0) + (dy0 * u001 + u011) * (wx0 * wz1) +
                         (dy0 * u100 + u110) * (wx1 * wz0) + (dy0 * u101 + u111) * (wx1 * wz1);
        const float g2 = (dz0 * u000 + u001) * (wx0 * wy0) + (dz0 * u010 + u011) * (wx0 * wy1) +
                         (dz0 * u100 + u101) * (wx1 * wy0) + (dz0 * u110 + u111) * (wx1 * wy1);
        float t0 = g0, t1 = g1, t2 = g2;
        if (a.accumulate) {
          t0 += sc[q];
          t1 += sc[q + 1];
          t2 += sc[q + 2];
        }
        o1[q] = t0;
        o1[q + 1] = t1;
        o1[q + 2] = t2;
        if (a.scale_cot) {
          o0[q] = a.alpha_tail * q0 + a.dnext * t0;
          o0[q + 1] = a.alpha_tail * q1 + a.dnext * t1;
          o0[q + 2] = a.alpha_tail * q2 + a.dnext * t2;
        }
      }
    }
    tma::fence_async_smem();  // every lane: its shared-memory writes before the bulk stores that read them
    __syncwarp();
    if (lane == 0) {
      const size_t off = p0 * 3;
      if (MODE == 0 || a.scale_cot) tma::bulk_s2g(a.out[0] + off, o0, BYTES);
      tma::bulk_s2g(a.out[1] + off, o1, BYTES);
      tma::bulk_commit();
    }
    __syncwarp();
  }
  if (lane == 0) tma::bulk_wait<0>();  // shared memory must outlive the last stores
}

// knobs: tune().gather_tma (1 = these kernels where they apply, 0 = cic4.cu's), tune().gather_seg (32 default | 64 | 128)

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Geometry check shared by both entry points; fills the frame part of the arguments.
static bool gather_tma_ok(GatherArgs& a, int64_t np, const Frame* fr, int seg) {
  if (!tune().gather_tma || np <= 0 || np % seg || np / seg >= ((int64_t)1 << 31)) return false;
  a.rel = 0;
  a.lat = 0;
  if (fr && fr->rel) {
    if (fr->nux != 1 || fr->nuy != 1 || fr->nuz != 1 || fr->dex != 1 || fr->dey != 1 || fr->dez != 1) return false;
    if (fr->pz % seg) return false;
    a.rel = 1;
    a.py = fr->py;
    a.pz = fr->pz;
    a.ox = fr->ox;
    a.oy = fr->oy;
    a.oz = fr->oz;
    // brick-ordered assignment when the lattice divides into 2 x 4 patches of rows
    if (fr->px % 2 == 0 && fr->py % 4 == 0 && tune().gather_brick) {
      a.lat = 1;
      a.lpy = fr->py;
      a.spp = fr->pz / seg;
      a.ntiles = (unsigned)((int64_t)(fr->px / 2) * (fr->py / 4) * a.spp);
    }
  }
  a.nseg = np / seg;
  return true;
}

template <int MODE, int SEG>
static void launch_gather_tma_seg(stream_t st, const GatherArgs& a) {
  using namespace gtma;
  constexpr int NIN = MODE == 0 ? 2 : 3, NOUT = 2;
  const size_t smem = (size_t)WARPS * (NSTAGE * NIN + 2 * NOUT) * SEG * 3 * sizeof(float);
  static int per_sm[16] = {0};  // resident CTAs per SM of this instantiation, per device (queried once)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (per_sm[dev] == 0) {
    cudaFuncSetAttribute(gather_tma_kernel<MODE, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gather_tma_kernel<MODE, SEG>, THREADS, smem) != cudaSuccess ||
        n < 1)
      n = 1;
    per_sm[dev] = n;
  }
  const int64_t want = a.lat ? (int64_t)a.ntiles : (a.nseg + WARPS - 1) / WARPS, wave = (int64_t)kSMs * per_sm[dev];
  const unsigned grid = (unsigned)(want < wave ? want : wave);  // persistent: at most one resident wave
  gather_tma_kernel<MODE, SEG><<<grid, THREADS, smem, st>>>(a);
}

template <int MODE>
static int launch_gather_tma(stream_t st, const GatherArgs& a, int seg) {
  count_launch();
  switch (seg) {
    case 64: launch_gather_tma_seg<MODE, 64>(st, a); break;
    case 128: launch_gather_tma_seg<MODE, 128>(st, a); break;
    default: launch_gather_tma_seg<MODE, 32>(st, a); break;
  }
  return rt_check(MODE == 0 ? "kick_drift4 (bulk-copy staged)" : "read_grad4v (bulk-copy staged)") ? -1 : 1;
}

// Return 1 if handled, 0 if the caller must take the cic4.cu kernel, < 0 on a launch error.
int kick_drift4_tma(stream_t st, const float* pos, const float* vel, const float* fmesh4, int64_t np, int nx, int ny,
                    int nz, float alpha, float beta, float drift, float* pos_out, float* vel_out, const Frame* fr) {
  const int gs = tune().gather_seg, seg = (gs == 64 || gs == 128) ? gs : 32;
  GatherArgs a = {};
  if (!gather_tma_ok(a, np, fr, seg)) return 0;
  if (!aligned16(pos) || !aligned16(vel) || !aligned16(pos_out) || !aligned16(vel_out) || !aligned16(fmesh4)) return 0;
  a.in[0] = pos;
  a.in[1] = vel;
  a.out[0] = pos_out;
  a.out[1] = vel_out;
  a.fm4 = reinterpret_cast<const float4*>(fmesh4);
  a.nx = nx;
  a.ny = ny;
  a.nz = nz;
  a.alpha = alpha;
  a.beta = beta;
  a.drift = drift;
  return launch_gather_tma<0>(st, a, seg);
}

int read_grad4v_tma(stream_t st, const float* pos, const float* fmesh4, const float* rhobar, float* cot, float cscale,
                    int scale_cot, float alpha_tail, int64_t np, int nx, int ny, int nz, float* grad, int accumulate,
                    const Frame* fr, float dnext) {
  const int gs = tune().gather_seg, seg = (gs == 64 || gs == 128) ? gs : 32;
  GatherArgs a = {};
  if (!gather_tma_ok(a, np, fr, seg)) return 0;
  if (!aligned16(pos) || !aligned16(cot) || !aligned16(grad) || !aligned16(fmesh4)) return 0;
  a.in[0] = pos;
  a.in[1] = cot;
  a.in[2] = grad;
  a.out[0] = cot;
  a.out[1] = grad;
  a.fm4 = reinterpret_cast<const float4*>(fmesh4);
  a.rhobar = rhobar;
  a.nx = nx;
  a.ny = ny;
  a.nz = nz;
  a.cscale = cscale;
  a.alpha_tail = alpha_tail;
  a.dnext = dnext;
  a.accumulate = accumulate;
  a.scale_cot = scale_cot;
  return launch_gather_tma<1>(st, a, seg);
}

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
