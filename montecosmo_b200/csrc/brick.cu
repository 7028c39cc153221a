// brick.cu -- brick-tiled CIC scatter for particles stored in Lagrangian-lattice order (sm_100a build only).
//
// The generic scatter sends 8 (density) or 8 x 16 B (reverse step) atomics per particle to L2, whose atomic units
// sustain ~1 lane-op per slice-clock (profiles/r1_atomics_microbench.txt): that, not HBM, bounds it.  Here one CTA
// takes a brick of 16 x 8 x 32 lattice-neighbour particles, accumulates them in a 26 x 18 x 44 shared-memory tile that
// follows the brick's mean displacement, and flushes only the touched 16-byte groups with red.global.add.v4.f32.
//
// Accumulation in shared memory: sm_100a has no native float atomics on shared memory (ATOMS.CAST.SPIN CAS loops,
// 9.5-16 cycles per warp-instruction) but int32 ATOMS.ADD is native, so contributions are accumulated in fixed point
// with a per-brick power-of-two scale S = 2^floor(log2(2^30 / sum_p |v_p|)): no cell of the tile can overflow, the
// quantum is <= 2^-18 of the brick's mean |v| per deposit (comparable to float32 rounding of the sums), and the result
// is independent of the order in which the brick's atomics land (bitwise reproducible inside the tile).
// Particles whose stencil leaves the tile use float atomics straight to global memory, so results never depend on the
// lattice hint being right.
//
// v2 of this kernel (22 x 14 x 40 tile) issued 600 warp-instructions per 32 particles and was issue-bound
// (profiles/r1_ncu_brick_v2.txt: 26% of them in the flush loop's index arithmetic, 38% around the ATOMS).  This version
// keeps per-particle state to 4 registers (tile cell + 3 fractions), works in float up to the single F2I of the tile
// cell, takes the tile origin from a 512-particle sample, looks the flush rows up in a shared table, and re-zeroes the
// tile inside the flush -- about a third of the instructions, with margins wide enough (5 / 5 / 6 cells each side) that
// fewer than 5% of the particles stray at a = 1 on the benchmark run (tools/stray_probe.py).
// Reference semantics: CIC paint of montecosmo/nbody.py:365-396 (same index rule and weights as cic4.cu / window.h).
#ifndef MCPM_HOSTEMU
#include "engine.h"
#include "tma.h"
#include "window.h"

namespace mcpm {

namespace brick {
constexpr int BX = 16, BY = 8, BZ = 32;  // lattice brick per CTA: 4096 particles, one warp per z-row
constexpr int TX = 26, TY = 18, TZ = 44;  // mesh tile (TZ % 4 == 0): 5.0 cells per particle, 82 KB of int32
constexpr int CELLS = TX * TY * TZ;       // 20592
constexpr int THREADS = 512;              // 16 warps, two CTAs per SM
constexpr int WARPS = THREADS / 32;
constexpr int PPT = BX * BY / WARPS;      // 8 z-rows (particles) per thread
constexpr int ZG = TZ / 4;                // 11 float4 groups per tile row
constexpr int TROWS = TX * TY;            // 468 tile rows
constexpr int GROUPS = TROWS * ZG;        // 5148 float4 groups per channel
constexpr size_t SMEM = sizeof(int) * (CELLS + TROWS) + sizeof(unsigned short) * PPT * THREADS;
static_assert(BX == 16 && BY == 8 && WARPS == 16, "row mapping below assumes a 16 x 8 x 32 brick and 16 warps");
}  // namespace brick

struct BrickArgs {
  int px, py, pz;  // particle lattice == mesh here (spacing 1 cell)
  int nx, ny, nz;
  float inx, iny, inz;  // 1 / n
  float shift;          // pos + shift is painted (interlacing); 0 otherwise
  int rel;              // 1: pos holds displacements from the lattice sites (frame.h), 0: absolute positions
  int ox, oy, oz;       // mesh cell of lattice site (0, 0, 0): the halo offset of a slab-decomposed rank
  const float* pos;
  // NCH = 1: value = (w ? w[p] : 1) * ws.        NCH = 3: value = s * A[p, c].
  const float* w;
  float ws;
  const float* A;
  float s;
  float* mesh;  // NCH planar meshes
  ObsShift obs;  // NCH = 1: optional redshift-space shift of the painted position (engine.h)
};

// thread (warp, lane), slot r -> lattice offsets inside the brick.  Slot 0 of the 16 warps samples all 16 x-planes and
// all 8 y-rows of the brick (the tile origin is taken from those 512 particles).
__device__ __forceinline__ int brick_qi(int warp, int r) { return ((warp >> 3) + 2 * ((warp & 7) + r)) & 15; }

template <int NCH, bool OBS = false>
__device__ __forceinline__ void stray_deposit(const BrickArgs& a, int si0, int sj0, int sk0);

// FULL: the lattice divides into whole bricks (no per-particle bounds checks).
// (A variant that handed each lane's four upper-z deposits to the next lane by shuffle -- 4 SHFL for 4 ATOMS -- was
// measured on a B200 in round 2 and removed: 29.98 vs 29.51 ms per evaluation, the shuffles cost what the atomics saved.)
// OBS: the painted position carries the fused redshift-space shift a.obs (a separate instantiation: the branch costs the
// plain density paint 8 % when it is a run-time test, 0.198 vs 0.183 ms at 256^3).
template <int NCH, bool FULL, bool OBS = false>
__global__ void __launch_bounds__(brick::THREADS, 2) brick_scatter_kernel(BrickArgs a) {
  using namespace brick;
  extern __shared__ __align__(16) int smem[];
  int* tile = smem;                 // [TX][TY][TZ] fixed point, reused channel after channel
  int* rowbase = smem + CELLS;      // [TX*TY] global offset of each tile row (wrapped)
  unsigned short* stray = reinterpret_cast<unsigned short*>(rowbase + TROWS);
  __shared__ float red[WARPS][4 + NCH];
  __shared__ float bc[4 + NCH];
  __shared__ int nstray;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) nstray = 0;
  int4* tile4 = reinterpret_cast<int4*>(tile);
  for (int i = tid; i < GROUPS; i += THREADS) tile4[i] = make_int4(0, 0, 0, 0);

  const int q0i = blockIdx.z * BX, q0j = blockIdx.y * BY, q0k = blockIdx.x * BZ;
  const int qj = q0j + (warp & 7), qk = q0k + lane;
  const bool colok = FULL || (qj < a.py && qk < a.pz);
  const int plane3 = 3 * a.py * a.pz;  // floats between consecutive lattice x-planes
  const int64_t pbase = 3 * (((int64_t)q0i * a.py + qj) * a.pz + qk);
  const float* xp = a.pos + pbase;

  // ---- pass 1: displacement u = x - site of the thread's 8 z-rows (exact when formed from an absolute position: the
  // site is a whole number below 2^24); for the reverse step also A += cb * B and sum |value|
  float x[PPT][3];
  unsigned valid = 0;
  const float sy = a.rel ? 0.f : (float)(qj + a.oy), sz = a.rel ? 0.f : (float)(qk + a.oz);
#pragma unroll
  for (int r = 0; r < PPT; ++r) {
    const int di = brick_qi(warp, r);
    if (FULL || (colok && q0i + di < a.px)) {
      valid |= 1u << r;
      const float* xr = xp + di * plane3;
      const float sx = a.rel ? 0.f : (float)(q0i + di + a.ox);
      float p0 = xr[0], p1 = xr[1], p2 = xr[2];
      if (OBS) {  // fused redshift-space shift, associated as rsd_shift (paint.cu) does
        const float* vr = a.obs.vel + pbase + di * plane3;
        const float sh = (vr[0] * a.obs.lx + vr[1] * a.obs.ly + vr[2] * a.obs.lz) * a.obs.coef;
        p0 = p0 + sh * a.obs.lx;
        p1 = p1 + sh * a.obs.ly;
        p2 = p2 + sh * a.obs.lz;
      }
      x[r][0] = (p0 + a.shift) - sx;
      x[r][1] = (p1 + a.shift) - sy;
      x[r][2] = (p2 + a.shift) - sz;
    } else {
      x[r][0] = x[r][1] = x[r][2] = 0.f;
    }
  }
  float l1[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) l1[c] = 0.f;
  if (NCH == 3) {  // sum |value| per channel (the fixed-point scale); the values are re-read channel by channel below
    const float* Ap = a.A + pbase;
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
      if (!FULL && !(valid >> r & 1)) continue;
      const int off = brick_qi(warp, r) * plane3;
      l1[0] += fabsf(a.s * Ap[off]);
      l1[NCH > 1 ? 1 : 0] += fabsf(a.s * Ap[off + 1]);
      l1[NCH > 2 ? 2 : 0] += fabsf(a.s * Ap[off + 2]);
    }
  } else if (a.w) {
    const float* wp = a.w + pbase / 3;
    const int plane1 = a.py * a.pz;
#pragma unroll
    for (int r = 0; r < PPT; ++r)
      if (FULL || (valid >> r & 1)) l1[0] += fabsf(wp[brick_qi(warp, r) * plane1] * a.ws);
  } else {
    l1[0] = fabsf(a.ws) * (float)__popc(valid);
  }
  // tile origin from slot 0: displacement off the lattice site, reduced to the nearest periodic image
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, cnt = 0.f;
  if (FULL || (valid & 1u)) {
    const float e0 = x[0][0], e1 = x[0][1], e2 = x[0][2];
    d0 = e0 - a.nx * rintf(e0 * a.inx);  // absolute positions may have been wrapped by the caller
    d1 = e1 - a.ny * rintf(e1 * a.iny);
    d2 = e2 - a.nz * rintf(e2 * a.inz);
    cnt = 1.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
    for (int c = 0; c < NCH; ++c) l1[c] += __shfl_xor_sync(0xffffffffu, l1[c], o);
  }
  if (lane == 0) {
    red[warp][0] = d0;
    red[warp][1] = d1;
    red[warp][2] = d2;
    red[warp][3] = cnt;
#pragma unroll
    for (int c = 0; c < NCH; ++c) red[warp][4 + c] = l1[c];
  }
  __syncthreads();  // also orders the zero fill and the stores to A before what follows
  if (tid < 4 + NCH) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[w][tid];
    bc[tid] = s;
  }
  __syncthreads();
  const float ninv = 1.0f / fmaxf(bc[3], 1.f);
  // origin (unwrapped cell coordinates, as floats: exact below 2^24) and its wrapped integer twin
  // (base cells of the brick span [q0 + mean, q0 + B - 1 + mean]; the tile admits base cells [o, o + T - 2])
  const float oxf = rintf((float)(q0i + a.ox) + bc[0] * ninv - 0.5f * (TX - BX));
  const float oyf = rintf((float)(q0j + a.oy) + bc[1] * ninv - 0.5f * (TY - BY));
  const float ozf = 4.0f * rintf(0.25f * ((float)(q0k + a.oz) + bc[2] * ninv - 0.5f * (TZ - 1 - BZ)));  // 16-byte groups along z
  const int oz = wrap_index((int)ozf, a.nz);
  if (tid < TROWS) {
    const int ix = tid / TY, jy = tid - ix * TY;
    int gx = wrap_index((int)oxf, a.nx) + ix, gy = wrap_index((int)oyf, a.ny) + jy;
    gx = gx >= a.nx ? gx - a.nx : gx;
    gy = gy >= a.ny ? gy - a.ny : gy;
    rowbase[tid] = (gx * a.ny + gy) * a.nz;
  }

  // ---- per-particle state: tile cell (or -1) and the three CIC fractions, held in place of x[r][.]
  int tcell[PPT];
  const float fnx = (float)a.nx, fny = (float)a.ny, fnz = (float)a.nz;
#pragma unroll
  for (int r = 0; r < PPT; ++r) {
    tcell[r] = -1;
    if (!FULL && !(valid >> r & 1)) continue;
    const float bx = floorf(x[r][0]), by = floorf(x[r][1]), bz = floorf(x[r][2]);
    x[r][0] -= bx;
    x[r][1] -= by;
    x[r][2] -= bz;
    // base cell (site + floor(u)) relative to the tile origin, modulo the mesh.  An inexact quotient can only misplace a
    // value onto +-n, i.e. out of the tile: the particle then strays, which is always correct.
    float tx = ((float)(q0i + brick_qi(warp, r) + a.ox) - oxf) + bx, ty = ((float)(qj + a.oy) - oyf) + by,
          tz = ((float)(qk + a.oz) - ozf) + bz;
    tx -= fnx * floorf(tx * a.inx);
    ty -= fny * floorf(ty * a.iny);
    tz -= fnz * floorf(tz * a.inz);
    const bool in = fabsf(tx - 0.5f * (TX - 2)) <= 0.5f * (TX - 2) && fabsf(ty - 0.5f * (TY - 2)) <= 0.5f * (TY - 2) &&
                    fabsf(tz - 0.5f * (TZ - 2)) <= 0.5f * (TZ - 2);
    if (in)
      tcell[r] = (int)((tx * TY + ty) * TZ + tz);  // exact: small integers
    else
      stray[atomicAdd(&nstray, 1)] = (unsigned short)(r * THREADS + tid);
  }
  __syncthreads();  // rowbase, nstray

  const int64_t plane = (int64_t)a.nx * a.ny * a.nz;
  bool all_stray = false;  // block-uniform
  // Per-particle values of the current channel, fetched ONE CHANNEL AHEAD: the loads of channel c + 1 are issued before
  // the flush of channel c and consumed after it.  (Round 1 re-read them inside the deposit loop, behind each particle's
  // branch: ncu put 45 % of this kernel's stall samples on the first use of those loads.)
  const float* valp0 = NCH == 1 ? (a.w ? a.w + pbase / 3 : nullptr) : a.A + pbase;
  const int vstride = NCH == 1 ? a.py * a.pz : plane3;
  float val[PPT];
#pragma unroll
  for (int r = 0; r < PPT; ++r)
    val[r] = (valp0 && (FULL || (valid >> r & 1))) ? valp0[brick_qi(warp, r) * vstride] : 1.0f;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const float l = bc[4 + c];
    // Fixed-point scale.  Safe: S = 2^floor(log2(2^30 / sum|v|)) cannot overflow whatever the clustering, but its quantum
    // (2^-18 of a unit weight for a 4096-particle brick) puts ~3e-6 of white noise on a unit-mean density mesh -- 50x
    // float32 rounding, and measured as a 4.5x larger displacement error than the float-atomic CPU port after 10 steps.
    // The density paint therefore deposits with 8 S (FINE = 3 bits: quantum 2^-21, float32-like) and verifies: the largest
    // true cell sum is then < 8 * 2^30, so a cell that left [-2^29, 2^29) -- possibly wrapped -- still reads outside that
    // band (an undetected alias would need a true sum beyond 2^32 - 2^29 > 8 * 2^30).  If any cell does (one cell holding
    // > 1/16 of the brick's weight: rare even in clusters) the tile is dropped and the whole brick goes through the stray
    // path below (float atomics to global memory).  The reverse-step channels keep the safe scale: their noise enters
    // the gradient linearly (~3e-6 relative), not through particles changing cells.
    constexpr int FINE = NCH == 1 ? 3 : 0;
    float e = l > 0.f ? floorf(log2f(1073741824.0f / l)) + (float)FINE : 0.f;
    e = fminf(fmaxf(e, -120.f), 120.f);
    const float S = exp2f(e), invS = exp2f(-e);
    const float scale = (NCH == 1 ? a.ws : a.s) * S;
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
      if (tcell[r] < 0) continue;
      const float vs = val[r] * scale;
      const float fx = x[r][0], fy = x[r][1], fz = x[r][2];
      const float gx = 1.f - fx, gy = 1.f - fy;
      const float vz1 = vs * fz, vz0 = vs - vz1;  // vs * (1 - fz) up to one rounding of the fixed-point product
      const float w00 = gx * gy, w01 = gx * fy, w10 = fx * gy, w11 = fx * fy;
      int* t = tile + tcell[r];
      atomicAdd(t, __float2int_rn(vz0 * w00));  // ATOMS.ADD, constant offsets
      atomicAdd(t + 1, __float2int_rn(vz1 * w00));
      atomicAdd(t + TZ, __float2int_rn(vz0 * w01));
      atomicAdd(t + TZ + 1, __float2int_rn(vz1 * w01));
      atomicAdd(t + TY * TZ, __float2int_rn(vz0 * w10));
      atomicAdd(t + TY * TZ + 1, __float2int_rn(vz1 * w10));
      atomicAdd(t + TY * TZ + TZ, __float2int_rn(vz0 * w11));
      atomicAdd(t + TY * TZ + TZ + 1, __float2int_rn(vz1 * w11));
    }
    __syncthreads();
    if (FINE > 0) {
      unsigned bad = 0;  // any cell outside [-2^29, 2^29)?
      for (int g = tid; g < GROUPS; g += THREADS) {
        const int4 q = tile4[g];
        bad |= ((unsigned)q.x + 0x20000000u) | ((unsigned)q.y + 0x20000000u) | ((unsigned)q.z + 0x20000000u) |
               ((unsigned)q.w + 0x20000000u);
      }
      if (__syncthreads_or((bad >> 30) != 0)) {  // drop the tile: every particle of the brick takes the stray path
        for (int i = tid; i < GROUPS; i += THREADS) tile4[i] = make_int4(0, 0, 0, 0);
        all_stray = true;
        __syncthreads();
      }
    }
    if (NCH > 1 && c + 1 < NCH) {  // next channel's values: in flight during the flush
#pragma unroll
      for (int r = 0; r < PPT; ++r)
        if (FULL || (valid >> r & 1)) val[r] = valp0[brick_qi(warp, r) * vstride + c + 1];
    }
    // flush the touched 16-byte groups (one red.global.add.v4.f32 each) and re-zero them for the next channel
    float* meshc = a.mesh + c * plane;
    for (int g = tid; g < GROUPS; g += THREADS) {
      const int4 q = tile4[g];
      if ((q.x | q.y | q.z | q.w) != 0) {
        if (c + 1 < NCH) tile4[g] = make_int4(0, 0, 0, 0);
        const int row = g / ZG, kz = g - row * ZG;  // compile-time divisor
        int gz = oz + 4 * kz;
        gz = gz >= a.nz ? gz - a.nz : gz;
        atomicAdd(reinterpret_cast<float4*>(meshc + (rowbase[row] + gz)),
                  make_float4(q.x * invS, q.y * invS, q.z * invS, q.w * invS));
      }
    }
    if (c + 1 < NCH) __syncthreads();
  }
  // strays: float atomics straight to global memory, one queued particle per thread (convergent); after a dropped tile,
  // every particle of the brick
  const int ns = all_stray ? PPT * THREADS : nstray;
  for (int si = tid; si < ns; si += THREADS) {
    const int code = all_stray ? si : stray[si], r = code / THREADS, t2 = code - r * THREADS;
    const int w2 = t2 >> 5;
    const int si0 = q0i + brick_qi(w2, r), sj0 = q0j + (w2 & 7), sk0 = q0k + (t2 & 31);
    if (!FULL && all_stray && (si0 >= a.px || sj0 >= a.py || sk0 >= a.pz)) continue;
    stray_deposit<NCH, OBS>(a, si0, sj0, sk0);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Streaming variant (round 2, "v4"): PERSISTENT CTAs, particle rows staged by bulk async copies.
//
// What the profile of the kernel above said (profiles/r2_ncu_kernels.txt, gpurun_out/r2a_full_brick.ncu-rep): 30-45 % of
// the stall samples sit on the first use of the strided 4-byte loads of pos / A (12 sectors per request, ~1 us of latency
// exposed at the start of every short-lived CTA), the LSU pipe is 63-70 % busy, and a full-tile zero fill, overflow check
// and flush sweep are paid per brick.  Here
//   * a CTA is resident for the whole launch and walks bricks b = blockIdx.x, + gridDim.x, ...;
//   * each z-row of the next brick (32 particles: 384 contiguous, 16-byte aligned bytes of pos, the same of A, 128 of w)
//     is fetched by ONE cp.async.bulk onto an mbarrier while the current brick deposits and flushes: no LSU
//     instruction, no register and no exposed latency on the global side of the particle arrays; shared memory is read
//     at a stride of 3 words (conflict free);
//   * the tile is zeroed once per launch: every flush re-zeroes what it wrote, and both the overflow check and the flush
//     visit only the bounding box of the brick's deposits (about half of the 18 x 18 rows of the tile);
//   * strays deposit straight to global memory where they are found (no queue).
// Brick 8 x 8 x 32 (2048 particles), tile 18 x 18 x 44: 57 KB + 48 KB of staging, two CTAs per SM.
// Same fixed-point tile arithmetic as above (per-brick scale from sum |v|; FINE bits with verified headroom for NCH = 1).
namespace bstream {
constexpr int BX = 8, BY = 8, BZ = 32, PPT = 4;
constexpr int TX = BX + 10, TY = BY + 10, TZ = 44, ZG = TZ / 4;
constexpr int ROWS = BX * BY, WARPS = ROWS / PPT, THREADS = WARPS * 32;  // 64 rows, 16 warps, 512 threads
constexpr int TROWS = TX * TY;
constexpr int PROW = 3 * BZ;  // floats per staged row of pos / A
template <int NCH, int TZS>
struct Smem {
  static constexpr int CELLS = TROWS * TZS;
  static constexpr int VROW = NCH == 3 ? PROW : BZ;
  static constexpr size_t TILE_B = sizeof(int) * CELLS;
  static constexpr size_t ROWBASE_B = sizeof(int) * ((TROWS + 3) & ~3);
  static constexpr size_t POS_B = sizeof(float) * ROWS * PROW;
  static constexpr size_t VAL_B = sizeof(float) * ROWS * VROW;
  static constexpr size_t STRAY_B = sizeof(unsigned short) * ROWS * BZ;
  static constexpr size_t BYTES = TILE_B + ROWBASE_B + POS_B + VAL_B + STRAY_B;
};
static_assert(WARPS % 8 == 0, "a warp keeps one lattice y-row");
}  // namespace bstream

// One stray particle (lattice site si0, sj0, sk0) as float atomics to global memory: the same u = x - site as the in-tile
// path, so that the fractions are bit-identical to it.
template <int NCH, bool OBS>
__device__ __forceinline__ void stray_deposit(const BrickArgs& a, int si0, int sj0, int sk0) {
  const int64_t plane = (int64_t)a.nx * a.ny * a.nz;
  const int64_t p3 = 3 * (((int64_t)si0 * a.py + sj0) * a.pz + sk0);
  float r0 = a.pos[p3], r1 = a.pos[p3 + 1], r2 = a.pos[p3 + 2];
  if (OBS) {
    const float* vr = a.obs.vel + p3;
    const float sh = (vr[0] * a.obs.lx + vr[1] * a.obs.ly + vr[2] * a.obs.lz) * a.obs.coef;
    r0 = r0 + sh * a.obs.lx;
    r1 = r1 + sh * a.obs.ly;
    r2 = r2 + sh * a.obs.lz;
  }
  const float px = (r0 + a.shift) - (a.rel ? 0.f : (float)(si0 + a.ox)),
              py = (r1 + a.shift) - (a.rel ? 0.f : (float)(sj0 + a.oy)),
              pz = (r2 + a.shift) - (a.rel ? 0.f : (float)(sk0 + a.oz));
  float vv[NCH];
  if (NCH == 1) {
    vv[0] = (a.w ? a.w[p3 / 3] : 1.0f) * a.ws;
  } else {
#pragma unroll
    for (int c = 0; c < NCH; ++c) vv[c] = a.s * a.A[p3 + c];
  }
  const float bx = floorf(px), by = floorf(py), bz = floorf(pz);
  const float fx = px - bx, fy = py - by, fz = pz - bz;
  const float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
  const float wxy[4] = {gx * gy, gx * fy, fx * gy, fx * fy};
  const int i0 = wrap_fast(si0 + a.ox + (int)bx, a.nx), j0 = wrap_fast(sj0 + a.oy + (int)by, a.ny),
            k0 = wrap_fast(sk0 + a.oz + (int)bz, a.nz);
  const int i1 = i0 + 1 == a.nx ? 0 : i0 + 1, j1 = j0 + 1 == a.ny ? 0 : j0 + 1, k1 = k0 + 1 == a.nz ? 0 : k0 + 1;
#pragma unroll
  for (int ab = 0; ab < 4; ++ab) {
    float* rowp = a.mesh + ((int64_t)((ab >> 1) ? i1 : i0) * a.ny + ((ab & 1) ? j1 : j0)) * a.nz;
    const float w0 = wxy[ab] * gz, w1 = wxy[ab] * fz;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      atomicAdd(rowp + c * plane + k0, vv[c] * w0);
      atomicAdd(rowp + c * plane + k1, vv[c] * w1);
    }
  }
}

// REL: lattice-relative positions (displacements; what the engine's step loop carries).  !REL: absolute positions,
// possibly wrapped by the caller by whole box lengths.
//
// Instruction diet (profiles/r2_ncu_brick_stream.txt): the first working version of this kernel issued 352 warp
// instructions per 32 particles and was ISSUE bound (67 % of the issue slots, 0.26 ms per 256^3 density paint against 0.16
// for the kernel above): 37 % of them in the check / flush loops (a runtime integer division per half-warp row), 8 % in
// the copy issue (two divisions per brick and thread), the rest float modulo arithmetic per particle.  Now: brick
// coordinates advance by carries, tile cells are formed in integers (one F2I.FLOOR per coordinate, a wrap only for the
// particles that need it), the origin comes from the first slot's 512 particles, flush items are decoded with
// multiply-shift divisions, and the unweighted density paint skips the sum |v| reduction.
template <int NCH, int TZS, bool REL>
__global__ void __launch_bounds__(bstream::THREADS, 2) brick_stream_kernel(BrickArgs a, int nby, int nbz, int nbricks) {
  using namespace bstream;
  using SM = Smem<NCH, TZS>;
  constexpr int ZGS = TZS / 4;  // float4 groups per tile row as stored (the first ZG are used)
  extern __shared__ __align__(128) unsigned char smraw[];
  int* tile = reinterpret_cast<int*>(smraw);
  int* rowbase = reinterpret_cast<int*>(smraw + SM::TILE_B);
  float* ps = reinterpret_cast<float*>(smraw + SM::TILE_B + SM::ROWBASE_B);  // [ROWS][PROW]
  float* vs = ps + ROWS * PROW;                                             // [ROWS][VROW]
  unsigned short* stray = reinterpret_cast<unsigned short*>(smraw + SM::TILE_B + SM::ROWBASE_B + SM::POS_B + SM::VAL_B);
  __shared__ int nstray;
  __shared__ float red[WARPS][4 + NCH];
  __shared__ int org[8];          // per brick, computed once and broadcast: tile origin (unwrapped x, y, z; wrapped x, y, z)
  __shared__ float scl[NCH][2];   // per brick and channel: value -> fixed point (sign of the scalar included), and back
  __shared__ int bb[4];  // bounding box of the brick's base cells in the tile: xmin, xmax, ymin, ymax
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int4* tile4 = reinterpret_cast<int4*>(tile);
  for (int i = tid; i < SM::CELLS / 4; i += THREADS) tile4[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    tma::mbar_init(&bar, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  const bool has_w = NCH == 1 && a.w != nullptr;
  const bool has_v = NCH == 3 || has_w;
  const uint32_t tx_bytes = (uint32_t)(SM::POS_B + (has_v ? SM::VAL_B : 0));
  // Fetch a brick: one bulk copy per array per z-row, issued by ALL warps (warp w: rows 4w .. 4w + 3, lanes 0-3 pos, lanes
  // 4-7 the values).  A first version issued the brick's 64-128 copies from warp 0 alone, lane after lane, with the whole
  // CTA waiting behind it at the next barrier.  Thread 0 posts the expected byte count; complete_tx of a copy may land
  // before it (the tx-count is signed, and the phase cannot complete before that one arrival).
  const int crow = PPT * warp + (lane & 3);  // the row this lane copies
  auto issue = [&](int bi, int bj, int bk) {
    if (lane < 8) {
      const int p0 = ((bi * BX + (crow >> 3)) * a.py + (bj * BY + (crow & 7))) * a.pz + bk * BZ;  // 3 np < 2^31 (brick_ok)
      if (lane < 4) tma::bulk_g2s(ps + crow * PROW, a.pos + 3 * (int64_t)p0, PROW * sizeof(float), &bar);
      else if (NCH == 3) tma::bulk_g2s(vs + crow * PROW, a.A + 3 * (int64_t)p0, PROW * sizeof(float), &bar);
      else if (has_w) tma::bulk_g2s(vs + crow * BZ, a.w + p0, BZ * sizeof(float), &bar);
    }
  };
  // brick coordinates advance by gridDim.x bricks with carries (no divisions in the loop)
  int bk = blockIdx.x % nbz, bj = (blockIdx.x / nbz) % nby, bi = blockIdx.x / nbz / nby;
  const int gk = gridDim.x % nbz, gj = (gridDim.x / nbz) % nby, gi = gridDim.x / nbz / nby;
  auto advance = [&](int& i, int& j, int& k) {
    k += gk;
    int c = k >= nbz;
    k -= c ? nbz : 0;
    j += gj + c;
    c = j >= nby;
    j -= c ? nby : 0;
    i += gi + c;
  };
  if ((int)blockIdx.x < nbricks) {
    if (tid == 0) tma::mbar_arrive_expect_tx(&bar, tx_bytes);
    issue(bi, bj, bk);
  }
  int ni = bi, nj = bj, nk = bk;
  advance(ni, nj, nk);

  const int64_t plane = (int64_t)a.nx * a.ny * a.nz;
  const int dj = warp & 7, di0 = warp >> 3;  // this warp's lattice row: (di0 + 2 r, dj), lane = z
  uint32_t parity = 0;
#pragma unroll 1
  for (int b = blockIdx.x; b < nbricks; b += gridDim.x, parity ^= 1) {
    const int q0i = bi * BX, q0j = bj * BY, q0k = bk * BZ;
    const bool more = b + (int)gridDim.x < nbricks;
    tma::mbar_wait(&bar, parity);
    if (tid == 0 && more) tma::mbar_arrive_expect_tx(&bar, tx_bytes);  // next phase: its copies follow barrier A

    // ---- pass 1: u = x (+ shift) - site, values; partial sums of the first slot's displacements and of |v|
    float u[PPT][3], val[NCH][PPT];
    float l1[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) l1[c] = 0.f;
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
      const int row = (di0 + 2 * r) * BY + dj;
      const float* xr = ps + row * PROW + 3 * lane;
      if (REL) {
        u[r][0] = xr[0] + a.shift;
        u[r][1] = xr[1] + a.shift;
        u[r][2] = xr[2] + a.shift;
      } else {  // exact: the site is a whole number below 2^24
        u[r][0] = (xr[0] + a.shift) - (float)(q0i + di0 + 2 * r + a.ox);
        u[r][1] = (xr[1] + a.shift) - (float)(q0j + dj + a.oy);
        u[r][2] = (xr[2] + a.shift) - (float)(q0k + lane + a.oz);
      }
      if (NCH == 3) {
        const float* vr = vs + row * PROW + 3 * lane;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          val[c][r] = vr[c];
          l1[c] += fabsf(val[c][r]);
        }
      } else if (has_w) {
        val[0][r] = vs[row * BZ + lane];
        l1[0] += fabsf(val[0][r]);
      } else {
        val[0][r] = 1.0f;
      }
    }
    float d0 = u[0][0], d1 = u[0][1], d2 = u[0][2];
    if (!REL) {  // nearest periodic image (absolute positions may have been wrapped by the caller)
      d0 -= (float)a.nx * rintf(d0 * a.inx);
      d1 -= (float)a.ny * rintf(d1 * a.iny);
      d2 -= (float)a.nz * rintf(d2 * a.inz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      if (has_v) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) l1[c] += __shfl_xor_sync(0xffffffffu, l1[c], o);
      }
    }
    if (lane == 0) {
      red[warp][0] = d0;
      red[warp][1] = d1;
      red[warp][2] = d2;
#pragma unroll
      for (int c = 0; c < NCH; ++c) red[warp][4 + c] = l1[c];
    }
    if (tid == 0) {
      bb[0] = TX;
      bb[1] = -1;
      bb[2] = TY;
      bb[3] = -1;
      nstray = 0;
    }
    __syncthreads();  // A: every thread has read the staged rows; red / bb are written
    if (more) issue(ni, nj, nk);  // the staging buffers are free: the next brick arrives under this one's deposits
    // Per-brick scalars are formed by ONE thread each and broadcast (the first version had all 16 warps redo the three
    // integer modulos, log2 / exp2 ... : ~150 of the 440 warp instructions a warp spent per brick between A' and B).
    // Threads 0-2: tile origin along x, y, z in unwrapped cell coordinates (base cells of the brick span
    // [q0 + mean, q0 + B - 1 + mean]; the tile admits base cells [o, o + T - 2]), along z a multiple of 4 (16-byte groups);
    // threads 4 .. 4 + NCH - 1: the channel's fixed-point scale.
    if (tid < 3) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) sum += red[w][tid];
      constexpr float ninv = 1.0f / (float)(WARPS * 32);  // the first slot: 512 particles spread over the whole brick
      const int q0 = tid == 0 ? q0i + a.ox : tid == 1 ? q0j + a.oy : q0k + a.oz;
      const int n = tid == 0 ? a.nx : tid == 1 ? a.ny : a.nz;
      const float slack = tid == 0 ? 0.5f * (TX - BX) : tid == 1 ? 0.5f * (TY - BY) : 0.5f * (TZ - 1 - BZ);
      const float of = (float)q0 + sum * ninv - slack;
      const int o = tid == 2 ? 4 * __float2int_rn(0.25f * of) : __float2int_rn(of);
      org[tid] = o;
      org[3 + tid] = wrap_index(o, n);
    } else if (tid >= 4 && tid < 4 + NCH) {
      const int c = tid - 4;
      float l;
      if (has_v) {
        l = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) l += red[w][tid];
        l *= fabsf(NCH == 1 ? a.ws : a.s);
      } else {
        l = fabsf(a.ws) * (float)(ROWS * BZ);
      }
      constexpr int FINE = NCH == 1 ? 3 : 0;  // see brick_scatter_kernel
      float e = l > 0.f ? floorf(log2f(1073741824.0f / l)) + (float)FINE : 0.f;
      e = fminf(fmaxf(e, -120.f), 120.f);
      scl[c][0] = (NCH == 1 ? a.ws : a.s) * exp2f(e);
      scl[c][1] = exp2f(-e);
    }
    __syncthreads();  // A'
    const int oxi = org[0], oyi = org[1], ozi = org[2], oxw = org[3], oyw = org[4], oz = org[5];
    if (tid < TROWS) {
      const int ix = tid / TY, jy = tid - ix * TY;
      int gx = oxw + ix, gy = oyw + jy;
      gx = gx >= a.nx ? gx - a.nx : gx;
      gy = gy >= a.ny ? gy - a.ny : gy;
      rowbase[tid] = (gx * a.ny + gy) * a.nz;
    }

    // ---- tile cell and fractions; strays are queued
    int tcell[PPT];
    int xmin = TX, xmax = -1, ymin = TY, ymax = -1;
    const int cxi = q0i + a.ox - oxi + di0, cyi = q0j + a.oy - oyi + dj, czi = q0k + a.oz - ozi + lane;
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
      const int ibx = __float2int_rd(u[r][0]), iby = __float2int_rd(u[r][1]), ibz = __float2int_rd(u[r][2]);
      u[r][0] -= (float)ibx;  // fractions: x - floor(x), as every CIC kernel of the engine forms them
      u[r][1] -= (float)iby;
      u[r][2] -= (float)ibz;
      int itx = ibx + cxi + 2 * r, ity = iby + cyi, itz = ibz + czi;
      if (!REL) {  // whole box lengths off (caller-wrapped positions): bring the cell back next to the tile
        if ((unsigned)itx > (unsigned)(TX - 2)) itx = wrap_fast(itx, a.nx);
        if ((unsigned)ity > (unsigned)(TY - 2)) ity = wrap_fast(ity, a.ny);
        if ((unsigned)itz > (unsigned)(TZ - 2)) itz = wrap_fast(itz, a.nz);
      }
      const bool in = (unsigned)itx <= (unsigned)(TX - 2) && (unsigned)ity <= (unsigned)(TY - 2) &&
                      (unsigned)itz <= (unsigned)(TZ - 2);
      if (in) {
        tcell[r] = (itx * TY + ity) * TZS + itz;
        xmin = min(xmin, itx);
        xmax = max(xmax, itx);
        ymin = min(ymin, ity);
        ymax = max(ymax, ity);
      } else {
        tcell[r] = -1;
        stray[atomicAdd(&nstray, 1)] = (unsigned short)(r * THREADS + tid);
      }
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin);
    xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin);
    ymax = __reduce_max_sync(0xffffffffu, ymax);
    if (lane == 0 && xmax >= 0) {
      atomicMin(&bb[0], xmin);
      atomicMax(&bb[1], xmax);
      atomicMin(&bb[2], ymin);
      atomicMax(&bb[3], ymax);
    }

    bool all_stray = false;  // block-uniform
    int nitem = 0, by0 = 0, nyb = 1, rbase = 0, ns = 0;
    unsigned mny = 0;
#pragma unroll  // val[c][r] stays in registers
    for (int c = 0; c < NCH; ++c) {
      constexpr int FINE = NCH == 1 ? 3 : 0;
      const float scale = scl[c][0], invS = scl[c][1];
#pragma unroll
      for (int r = 0; r < PPT; ++r) {
        if (tcell[r] < 0) continue;
        const float vsc = val[c][r] * scale;
        const float fx = u[r][0], fy = u[r][1], fz = u[r][2];
        const float gx = 1.f - fx, gy = 1.f - fy;
        const float vz1 = vsc * fz, vz0 = vsc - vz1;  // vsc * (1 - fz) up to one rounding of the fixed-point product
        const float w00 = gx * gy, w01 = gx * fy, w10 = fx * gy, w11 = fx * fy;
        int* t = tile + tcell[r];
        atomicAdd(t, __float2int_rn(vz0 * w00));  // ATOMS.ADD, constant offsets
        atomicAdd(t + 1, __float2int_rn(vz1 * w00));
        atomicAdd(t + TZS, __float2int_rn(vz0 * w01));
        atomicAdd(t + TZS + 1, __float2int_rn(vz1 * w01));
        atomicAdd(t + TY * TZS, __float2int_rn(vz0 * w10));
        atomicAdd(t + TY * TZS + 1, __float2int_rn(vz1 * w10));
        atomicAdd(t + TY * TZS + TZS, __float2int_rn(vz0 * w11));
        atomicAdd(t + TY * TZS + TZS + 1, __float2int_rn(vz1 * w11));
      }
      __syncthreads();  // B: the tile holds channel c (and bb, rowbase, nstray are final)
      if (c == 0) {  // read once: thread 0 re-initialises bb / nstray in the next brick's pass 1
        // rows [bb0, bb1 + 1] x [bb2, bb3 + 1] of the tile were touched: nrow x ZG groups, item q -> (row, kz) by
        // multiply-shift divisions (exact: q < 2^15 with divisor 11; row index < 324 with divisor <= 18)
        by0 = bb[2];
        nyb = bb[3] - by0 + 2;
        nitem = bb[1] < 0 ? 0 : (bb[1] - bb[0] + 2) * nyb * ZG;
        rbase = bb[0] * TY + by0;
        mny = (65536u + (unsigned)nyb - 1u) / (unsigned)nyb;
        ns = nstray;
      }
      auto group_of = [&](int q, int& row, int& kz) {
        const int rb = (int)(((unsigned)q * 5958u) >> 16);  // q / 11
        kz = q - rb * ZG;
        const int ixr = (int)(((unsigned)rb * mny) >> 16);  // rb / nyb
        row = rbase + ixr * TY + (rb - ixr * nyb);
      };
      static_assert(ZG == 11, "multiply-shift constant above");
      if (FINE > 0) {
        unsigned bad = 0;  // any cell outside [-2^29, 2^29)?
        for (int q = tid; q < nitem; q += THREADS) {
          int row, kz;
          group_of(q, row, kz);
          const int4 v = tile4[row * ZGS + kz];
          bad |= ((unsigned)v.x + 0x20000000u) | ((unsigned)v.y + 0x20000000u) | ((unsigned)v.z + 0x20000000u) |
                 ((unsigned)v.w + 0x20000000u);
        }
        if (__syncthreads_or((bad >> 30) != 0)) all_stray = true;  // drop the tile: the flush below only re-zeroes it
      }
      float* meshc = a.mesh + c * plane;
      for (int q = tid; q < nitem; q += THREADS) {
        int row, kz;
        group_of(q, row, kz);
        const int g = row * ZGS + kz;
        const int4 v = tile4[g];
        if ((v.x | v.y | v.z | v.w) != 0) {
          tile4[g] = make_int4(0, 0, 0, 0);
          if (!all_stray) {
            int gz = oz + 4 * kz;
            gz = gz >= a.nz ? gz - a.nz : gz;
            atomicAdd(reinterpret_cast<float4*>(meshc + (rowbase[row] + gz)),
                      make_float4(v.x * invS, v.y * invS, v.z * invS, v.w * invS));
          }
        }
      }
      if (c + 1 < NCH) __syncthreads();  // the tile is clean before the next channel's deposits
    }
    // strays: float atomics straight to global memory, one queued particle per thread (convergent); after a dropped tile
    // (NCH == 1 only), every particle of the brick
    if (all_stray) ns = PPT * THREADS;
    for (int si = tid; si < ns; si += THREADS) {
      const int code = all_stray ? si : stray[si], r = code / THREADS, t2 = code - r * THREADS, w2 = t2 >> 5;
      stray_deposit<NCH>(a, q0i + (w2 >> 3) + 2 * r, q0j + (w2 & 7), q0k + (t2 & 31));
    }
    // no barrier here: the next brick's pass 1 touches only the staging buffers, red and bb, and its barrier A orders
    // this flush before the next deposits
    bi = ni;
    bj = nj;
    bk = nk;
    advance(ni, nj, nk);
  }
}

// The lattice has a spacing of one cell; it need not span the mesh (a slab-decomposed rank paints its xl x ny x nz
// particles into a mesh extended by halo planes).
static bool brick_ok(const Lattice& L, int64_t np, int nx, int ny, int nz) {
  using namespace brick;
  if (!tune().brick) return false;  // mcpm_tune("brick", 0): A/B against the generic global-atomic kernels
  if (L.px <= 0 || L.py <= 0 || L.pz <= 0 || L.px > nx || L.py > ny || L.pz > nz) return false;
  if ((int64_t)L.px * L.py * L.pz != np) return false;
  if ((nz & 3) || nx < TX || ny < TY || nz < TZ) return false;         // a tile must not wrap onto itself
  if (3 * np >= ((int64_t)1 << 31)) return false;                  // 32-bit row offsets inside a brick
  return true;
}


template <int NCH, bool FULL, bool OBS>
static void launch_brick_obs(stream_t st, const BrickArgs& a, dim3 grid) {
  using namespace brick;
  cudaFuncSetAttribute(brick_scatter_kernel<NCH, FULL, OBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  brick_scatter_kernel<NCH, FULL, OBS><<<grid, THREADS, SMEM, st>>>(a);
}
template <int NCH, bool FULL>
static void launch_brick_variant(stream_t st, const BrickArgs& a, dim3 grid) {
  if (NCH == 1 && a.obs.vel) launch_brick_obs<1, FULL, true>(st, a, grid);
  else launch_brick_obs<NCH, FULL, false>(st, a, grid);
}

// Streaming variant: whole 8 x 8 x 32 bricks, 16-byte aligned rows (bulk copies), mesh at least one tile wide.
template <int NCH, int TZS, bool REL>
static int launch_stream(stream_t st, const BrickArgs& a) {
  using namespace bstream;
  using SM = Smem<NCH, TZS>;
  static int per_sm[16] = {0};  // resident CTAs per SM, per device (queried once)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (per_sm[dev] == 0) {
    cudaFuncSetAttribute(brick_stream_kernel<NCH, TZS, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, brick_stream_kernel<NCH, TZS, REL>, THREADS, SM::BYTES) !=
            cudaSuccess || n < 1)
      n = 1;
    per_sm[dev] = n;
  }
  const int nbx = a.px / BX, nby = a.py / BY, nbz = a.pz / BZ;
  const int64_t nb = (int64_t)nbx * nby * nbz, wave = (int64_t)kSMs * per_sm[dev];
  count_launch();
  brick_stream_kernel<NCH, TZS, REL><<<(unsigned)(nb < wave ? nb : wave), THREADS, SM::BYTES, st>>>(a, nby, nbz, (int)nb);
  return rt_check("brick_stream") ? -1 : 1;
}
static bool stream_ok(const BrickArgs& a) {
  using namespace bstream;
  if (!tune().brick_stream || a.obs.vel) return false;  // the fused redshift-space shift lives in the per-brick kernel
  if (a.px % BX || a.py % BY || a.pz % BZ) return false;
  if (a.nx < TX || a.ny < TY || a.nz < TZ) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(a.pos) && al16(a.A) && al16(a.w);
}

template <int NCH>
static int launch_brick(stream_t st, const BrickArgs& a) {
  using namespace brick;
  // Measured on a B200 at 256^3 (profiles/r2_brick_stream.txt): the 3-channel reverse scatter gains 10 % early in a run and
  // nothing at a = 1 (0.351 vs 0.392 ms, 0.435 vs 0.438); the density paint LOSES 6-15 % (0.173 vs 0.163, 0.210 vs 0.183:
  // one channel does not amortise the per-brick passes over a tile a third of whose rows it never touches) and keeps the
  // per-brick kernel unless mcpm_tune("brick_stream1", 1).
  if (stream_ok(a) && (NCH == 3 || tune().brick_stream1)) {
    const bool wide = tune().brick_stream == 48;
    if (a.rel) return wide ? launch_stream<NCH, 48, true>(st, a) : launch_stream<NCH, 44, true>(st, a);
    return wide ? launch_stream<NCH, 48, false>(st, a) : launch_stream<NCH, 44, false>(st, a);
  }
  dim3 grid((a.pz + BZ - 1) / BZ, (a.py + BY - 1) / BY, (a.px + BX - 1) / BX);
  count_launch();
  const bool full = a.px % BX == 0 && a.py % BY == 0 && a.pz % BZ == 0;
  if (full) launch_brick_variant<NCH, true>(st, a, grid);
  else launch_brick_variant<NCH, false>(st, a, grid);
  return rt_check("brick_scatter") ? -1 : 1;
}

// CIC density paint of pos + shift into a zeroed-or-accumulated planar mesh.  Returns 1 if handled, 0 -> generic path,
// < 0 error.
// `fr` (optional): relative frame of spacing one cell on lattice L -- pos holds displacements, fr->o the halo offset.
static bool frame_fits(const Lattice& L, const Frame* fr) {
  if (!fr || !fr->rel) return true;
  return fr->px == L.px && fr->py == L.py && fr->pz == L.pz && fr->nux == 1 && fr->nuy == 1 && fr->nuz == 1 &&
         fr->dex == 1 && fr->dey == 1 && fr->dez == 1;
}
static void set_frame(BrickArgs& a, const Frame* fr) {
  if (fr && fr->rel) {
    a.rel = 1;
    a.ox = fr->ox;
    a.oy = fr->oy;
    a.oz = fr->oz;
  }
}

int brick_paint_cic(stream_t st, const Lattice& L, const float* pos, const float* weights, float wscalar, float shift,
                    int64_t np, int nx, int ny, int nz, float* mesh, const Frame* fr, const ObsShift* obs) {
  if (!brick_ok(L, np, nx, ny, nz) || !frame_fits(L, fr)) return 0;
  BrickArgs a = {};
  set_frame(a, fr);
  if (obs) a.obs = *obs;
  a.px = L.px; a.py = L.py; a.pz = L.pz; a.nx = nx; a.ny = ny; a.nz = nz;
  a.inx = 1.0f / nx; a.iny = 1.0f / ny; a.inz = 1.0f / nz;
  a.shift = shift;
  a.pos = pos; a.w = weights; a.ws = wscalar; a.mesh = mesh;
  return launch_brick<1>(st, a);
}

// Reverse-step scatter: mesh3[c] += s * A[., c] * W, three planar meshes.
int brick_paint3_cic(stream_t st, const Lattice& L, const float* pos, const float* A, float s, int64_t np, int nx, int ny,
                     int nz, float* mesh3, const Frame* fr) {
  if (!brick_ok(L, np, nx, ny, nz) || !frame_fits(L, fr)) return 0;
  BrickArgs a = {};
  set_frame(a, fr);
  a.px = L.px; a.py = L.py; a.pz = L.pz; a.nx = nx; a.ny = ny; a.nz = nz;
  a.inx = 1.0f / nx; a.iny = 1.0f / ny; a.inz = 1.0f / nz;
  a.pos = pos; a.A = A; a.s = s; a.mesh = mesh3;
  return launch_brick<3>(st, a);
}

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
