// brick.cu -- brick-tiled CIC scatter for particles stored in Lagrangian-lattice order (sm_100a build only).
//
// The generic scatter sends 8 (density) or 8 x 16 B (reverse step) atomics per particle to L2, whose atomic units
// sustain ~1 lane-op per slice-clock (profiles/r1_atomics_microbench.txt): that, not HBM, bounds it.  Here one CTA
// takes a brick of 16 x 8 x 32 lattice-neighbour particles, accumulates them in a 22 x 14 x 40 shared-memory tile that
// follows the brick's mean displacement, and flushes only the touched 16-byte groups with red.global.add.v4.f32:
// ~0.7 L2 lane-ops per particle instead of 8.
//
// Accumulation in shared memory: sm_100a has no native float atomics on shared memory (ATOMS.CAST.SPIN CAS loops,
// 9.5-16 cycles per warp-instruction) but int32 ATOMS.ADD is native, so contributions are accumulated in fixed point
// with a per-brick power-of-two scale S = 2^floor(log2(2^30 / sum_p |v_p|)): no cell of the tile can overflow, the
// quantum is <= 2^-19 of the brick's mean |v| per deposit (comparable to float32 rounding of the sums), and the result
// is independent of the order in which the brick's atomics land (bitwise reproducible inside the tile).
// Particles whose stencil leaves the tile use float atomics straight to global memory, so results never depend on the
// lattice hint being right.  Geometry is compile-time (no runtime divisions in the zero / flush loops -- the v1 of this
// kernel lost to exactly that, tools/experiments/brick_tiled_v1.cu.txt).
// Reference semantics: CIC paint of montecosmo/nbody.py:365-396 (same index rule and weights as cic4.cu / window.h).
#ifndef MCPM_HOSTEMU
#include "engine.h"
#include "window.h"

namespace mcpm {

namespace brick {
constexpr int BX = 16, BY = 8, BZ = 32;         // lattice brick per CTA: 4096 particles
// mesh tile = brick + margin 6 / 6 / 8 (TZ % 4 == 0): 3 cells per particle.  Measured on the benchmark run at 256^3
// (tools/stray_probe.py): particles leaving a tile of margin M at a = 1: 46% (M=6), 16% (8), 4% (10); at a = 0.3: 0.4%.
// Margin 10 (26 x 18 x 44) was nevertheless slower (3-channel kernel 678 vs 577 us): zeroing and scanning the tile once
// per channel costs more than sending strays down the global-atomic path.
constexpr int TX = 22, TY = 14, TZ = 40;
constexpr int CELLS = TX * TY * TZ;             // 12320 (48 KB of int32: two CTAs per SM)
constexpr int THREADS = 512;                    // 16 warps
constexpr int ROWS = BX * BY;                   // 128 z-rows of 32 particles
constexpr int PPT = ROWS / 16;                  // 8 particles (z-rows) per thread: amortises zero / reduce / flush
constexpr int GROUPS = TX * TY * (TZ / 4);      // 3080 float4 groups per channel
}  // namespace brick

struct BrickArgs {
  int px, py, pz;  // particle lattice == mesh here (spacing 1 cell)
  int nx, ny, nz;
  float inx, iny, inz;  // 1 / n
  const float* pos;
  // NCH = 1: value = (w ? w[p] : 1) * ws.        NCH = 3: value = s * (A[p] + cb * B[p]), A updated in place if store.
  const float* w;
  float ws;
  float* A;
  const float* B;
  float cb, s;
  int store;
  float* mesh;  // NCH planar meshes
};

// d mod n into [-n/2, n/2): one conditional correction covers |d| < 1.5 n (always, unless the caller wrapped positions
// by several boxes); the integer division is only the rare fallback
__device__ __forceinline__ int fold(int d, int n) {
  const int h = (n + 1) >> 1;
  if (d >= h) d -= n;
  else if (d < -h) d += n;
  if (d >= h || d < -h) {
    int m = wrap_index(d, n);
    d = m >= h ? m - n : m;
  }
  return d;
}

template <int NCH>
__global__ void __launch_bounds__(brick::THREADS, 2) brick_scatter_kernel(BrickArgs a) {
  using namespace brick;
  extern __shared__ __align__(16) int tile[];  // [TX][TY][TZ] fixed point, reused channel after channel
  __shared__ float red[16][4 + NCH];
  __shared__ float bc[4 + NCH];
  __shared__ int stray[PPT * THREADS];
  __shared__ int nstray;
  if (threadIdx.x == 0) nstray = 0;
  for (int i = threadIdx.x; i < CELLS / 4; i += THREADS) reinterpret_cast<int4*>(tile)[i] = make_int4(0, 0, 0, 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0i = blockIdx.z * BX, q0j = blockIdx.y * BY, q0k = blockIdx.x * BZ;
  // PPT z-rows per thread: rows warp, warp + 16, ... of the brick's 128
  float x[PPT][3];
  unsigned validmask = 0;
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, cnt = 0.f, l1[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) l1[c] = 0.f;
#pragma unroll
  for (int r = 0; r < PPT; ++r) {
    const int row = warp + 16 * r;
    const int qi = q0i + row / BY, qj = q0j + row % BY, qk = q0k + lane;
    if (qi < a.px && qj < a.py && qk < a.pz) {
      validmask |= 1u << r;
      const int64_t p = ((int64_t)qi * a.py + qj) * a.pz + qk;
      x[r][0] = a.pos[3 * p];
      x[r][1] = a.pos[3 * p + 1];
      x[r][2] = a.pos[3 * p + 2];
      if (NCH == 1) {
        l1[0] += fabsf((a.w ? a.w[p] : 1.0f) * a.ws);
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          float t = a.A[3 * p + c];
          if (a.B) {
            t += a.cb * a.B[3 * p + c];
            if (a.store) a.A[3 * p + c] = t;  // re-read per channel pass below (L1 / L2 hit)
          }
          l1[c] += fabsf(a.s * t);
        }
      }
      float e0 = x[r][0] - qi, e1 = x[r][1] - qj, e2 = x[r][2] - qk;
      e0 -= a.nx * rintf(e0 * a.inx);  // positions may have been wrapped by the caller
      e1 -= a.ny * rintf(e1 * a.iny);
      e2 -= a.nz * rintf(e2 * a.inz);
      d0 += e0;
      d1 += e1;
      d2 += e2;
      cnt += 1.f;
    }
  }
  // block reduction: mean displacement of the brick and sum |v| per channel
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
  if (threadIdx.x < 4 + NCH) {
    float s = 0.f;
    for (int w = 0; w < 16; ++w) s += red[w][threadIdx.x];
    bc[threadIdx.x] = s;
  }
  __syncthreads();
  const float n = fmaxf(bc[3], 1.f);
  const int ox = (int)floorf(q0i + bc[0] / n - 0.5f * (TX - BX));
  const int oy = (int)floorf(q0j + bc[1] / n - 0.5f * (TY - BY));
  int oz = (int)floorf(q0k + bc[2] / n - 0.5f * (TZ - BZ));
  oz &= ~3;  // round down to a multiple of 4 (two's complement): 16-byte aligned groups along z

  // tile-local base cell of every particle (packed), strays queued once
  int tcell[PPT];
#pragma unroll
  for (int r = 0; r < PPT; ++r) {
    tcell[r] = -1;
    if (!(validmask >> r & 1)) continue;
    const int ti = fold((int)floorf(x[r][0]) - ox, a.nx), tj = fold((int)floorf(x[r][1]) - oy, a.ny),
              tk = fold((int)floorf(x[r][2]) - oz, a.nz);
    if ((unsigned)ti <= (unsigned)(TX - 2) && (unsigned)tj <= (unsigned)(TY - 2) && (unsigned)tk <= (unsigned)(TZ - 2))
      tcell[r] = (ti * TY + tj) * TZ + tk;
    else
      stray[atomicAdd(&nstray, 1)] = ((warp + 16 * r) << 5) | lane;
  }

  const int64_t plane = (int64_t)a.nx * a.ny * a.nz;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (c > 0) {  // the tile was flushed by the previous channel pass: clear it
      __syncthreads();
      for (int i = threadIdx.x; i < CELLS / 4; i += THREADS) reinterpret_cast<int4*>(tile)[i] = make_int4(0, 0, 0, 0);
      __syncthreads();
    }
    const float l = bc[4 + c];
    float e = l > 0.f ? floorf(log2f(1073741824.0f / l)) : 0.f;
    e = fminf(fmaxf(e, -120.f), 120.f);
    const float S = exp2f(e), invS = exp2f(-e);
#pragma unroll
    for (int r = 0; r < PPT; ++r) {
      if (tcell[r] < 0) continue;
      const int row = warp + 16 * r;
      const int64_t p = ((int64_t)(q0i + row / BY) * a.py + (q0j + row % BY)) * a.pz + q0k + lane;
      const float val = NCH == 1 ? (a.w ? a.w[p] : 1.0f) * a.ws : a.s * a.A[3 * p + c];
      const float fx = x[r][0] - floorf(x[r][0]), fy = x[r][1] - floorf(x[r][1]), fz = x[r][2] - floorf(x[r][2]);
      const float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
      const float vs = val * S, vz0 = vs * gz, vz1 = vs * fz;
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
    // flush the touched 16-byte groups: one red.global.add.v4.f32 per group
    float* meshc = a.mesh + c * plane;
    for (int g = threadIdx.x; g < GROUPS; g += THREADS) {
      const int kz = g % (TZ / 4), rr = g / (TZ / 4);  // compile-time divisors
      const int jy = rr % TY, ix = rr / TY;
      const int4 q = *reinterpret_cast<const int4*>(tile + (ix * TY + jy) * TZ + 4 * kz);
      if ((q.x | q.y | q.z | q.w) != 0) {
        const int gx = wrap_fast(ox + ix, a.nx), gy = wrap_fast(oy + jy, a.ny), gz = wrap_fast(oz + 4 * kz, a.nz);
        atomicAdd(reinterpret_cast<float4*>(meshc + ((int64_t)gx * a.ny + gy) * a.nz + gz),
                  make_float4(q.x * invS, q.y * invS, q.z * invS, q.w * invS));
      }
    }
  }
  // strays: float atomics straight to global memory, one queued particle per thread (convergent)
  for (int si = threadIdx.x; si < nstray; si += THREADS) {
    const int code = stray[si], row = code >> 5, ln = code & 31;
    const int qi = q0i + row / BY, qj = q0j + row % BY, qk = q0k + ln;
    const int64_t p = ((int64_t)qi * a.py + qj) * a.pz + qk;
    const float px = a.pos[3 * p], py = a.pos[3 * p + 1], pz = a.pos[3 * p + 2];
    float vv[NCH];
    if (NCH == 1) {
      vv[0] = (a.w ? a.w[p] : 1.0f) * a.ws;
    } else {
#pragma unroll
      for (int c = 0; c < NCH; ++c) vv[c] = a.s * a.A[3 * p + c];  // A already holds A + cb * B (stored above)
    }
    const float bx = floorf(px), by = floorf(py), bz = floorf(pz);
    const float fx = px - bx, fy = py - by, fz = pz - bz;
    const float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
    const float wxy[4] = {gx * gy, gx * fy, fx * gy, fx * fy};
    const int i0 = wrap_fast((int)bx, a.nx), j0 = wrap_fast((int)by, a.ny), k0 = wrap_fast((int)bz, a.nz);
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
}

static bool brick_ok(const Lattice& L, int64_t np, int nx, int ny, int nz) {
  using namespace brick;
  if (L.px != nx || L.py != ny || L.pz != nz) return false;        // lattice spacing of one cell only
  if ((int64_t)L.px * L.py * L.pz != np) return false;
  if ((nz & 3) || nx < 2 * TX || ny < 2 * TY || nz < 2 * TZ) return false;
  return true;
}

template <int NCH>
static int launch_brick(stream_t st, const BrickArgs& a) {
  using namespace brick;
  const size_t smem = sizeof(int) * CELLS;  // one channel at a time
  cudaFuncSetAttribute(brick_scatter_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device
  dim3 grid((a.pz + BZ - 1) / BZ, (a.py + BY - 1) / BY, (a.px + BX - 1) / BX);
  count_launch();
  brick_scatter_kernel<NCH><<<grid, THREADS, smem, st>>>(a);
  return rt_check("brick_scatter") ? -1 : 1;
}

// CIC density paint into a zeroed-or-accumulated planar mesh.  Returns 1 if handled, 0 -> generic path, < 0 error.
int brick_paint_cic(stream_t st, const Lattice& L, const float* pos, const float* weights, float wscalar, int64_t np,
                    int nx, int ny, int nz, float* mesh) {
  if (!brick_ok(L, np, nx, ny, nz)) return 0;
  BrickArgs a = {};
  a.px = L.px; a.py = L.py; a.pz = L.pz; a.nx = nx; a.ny = ny; a.nz = nz;
  a.inx = 1.0f / nx; a.iny = 1.0f / ny; a.inz = 1.0f / nz;
  a.pos = pos; a.w = weights; a.ws = wscalar; a.mesh = mesh;
  return launch_brick<1>(st, a);
}

// Reverse-step scatter: A += cb * B (stored when B != NULL), mesh3[c] += s * A[., c] * W, three planar meshes.
int brick_paint3_cic(stream_t st, const Lattice& L, const float* pos, float* A, const float* B, float cb, float s,
                     int64_t np, int nx, int ny, int nz, float* mesh3) {
  if (!brick_ok(L, np, nx, ny, nz)) return 0;
  BrickArgs a = {};
  a.px = L.px; a.py = L.py; a.pz = L.pz; a.nx = nx; a.ny = ny; a.nz = nz;
  a.inx = 1.0f / nx; a.iny = 1.0f / ny; a.inz = 1.0f / nz;
  a.pos = pos; a.A = A; a.B = B; a.cb = cb; a.s = s; a.store = B != nullptr; a.mesh = mesh3;
  return launch_brick<3>(st, a);
}

}  // namespace mcpm
#endif  // MCPM_HOSTEMU
