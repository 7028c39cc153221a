// cic4.cu -- CIC kernels of the BullFrog step on a float4-interleaved vector mesh ("mesh4": [nx,ny,nz] cells of
// {c0, c1, c2, c3}).  The three force components (forward) or the three cotangent channels (backward) of one cell sit
// in one 16-byte word, so a particle touches 8 x 16 B instead of 24 x 4 B:
//   * kick_drift4   : 8 LDG.128 trilinear readout of F + kick + drift                     (nbody.py:933-951)
//   * paint3v4      : 8 RED.128 (red.global.add.v4.f32) scatter of beta * vbar, fused with vbar += xbar * drift
//   * read_grad4v   : 8 LDG.128 + 8 LDG.32 gradient gather from (F, rhobar), fused with xbar += ..., vbar *= alpha
//   * interleave3 / deinterleave3: planar <-> mesh4 conversion around cuFFT (strided cuFFT plans measured 2.3x slower
//     than planar + a 0.08 ms conversion pass at 256^3, profiles/r1_layout_microbench.txt)
// Measured on B200 (tools/mapbench.cu): 3-channel scatter 0.57-0.70 ms as float4 vs 1.0-1.66 ms planar; 3-channel gather
// 0.33 ms vs 0.38-0.73 ms.  Same arithmetic as the generic kernels in paint.cu (window.h), CIC only.
#include "engine.h"
#include "window.h"

namespace mcpm {

struct alignas(16) f4 {
  float x, y, z, w;
};

#if defined(MCPM_HOSTEMU)
MCPM_HD void atomic_add4(f4* p, f4 v) {
  atomic_add(&p->x, v.x);
  atomic_add(&p->y, v.y);
  atomic_add(&p->z, v.z);
  atomic_add(&p->w, v.w);
}
MCPM_HD f4 load4(const f4* p) { return *p; }
#else
__device__ __forceinline__ void atomic_add4(f4* p, f4 v) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, v.w));  // RED.E.ADD.F32x4
}
__device__ __forceinline__ f4 load4(const f4* p) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return f4{v.x, v.y, v.z, v.w};
}
#endif

// Resident CTAs per SM the two gather kernels are compiled for (register cap 64 / 48 / 40 per thread).  They are bound by
// load latency at ~50% occupancy (profiles/r1_ncu_brick_v2.txt), so fewer registers can pay; mcpm_tune("gather_minb").
template <class F>
static void launch_gather(stream_t st, int64_t n, F f) {
#ifndef MCPM_HOSTEMU
  // blocked: one CTA per 256 consecutive particles, scheduled in order -- a single front sweeps the mesh, so every mesh
  // line is fetched from HBM about once; the grid-stride launch keeps ~27 fronts alive, more than L2 holds
  if (tune().gather_blocked && n > 0) {
    count_launch();
    k_launch_1d<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f);
    return;
  }
#endif
  if (tune().gather_minb == 5) launch_1d_occ<5>(st, n, f);
  else if (tune().gather_minb == 6) launch_1d_occ<6>(st, n, f);
  else launch_1d(st, n, f);
}

// CIC base cell and fractions; cheap wrap for the usual range, exact modulo otherwise (nbody.py:372-376, 388)
struct Cic {
  int i0, i1, j0, j1, k0, k1;
  float fx, fy, fz;
};

// `x` is the absolute position, or (relative frame, frame.h) the displacement from the particle's lattice site: the site
// enters as exact integers, the fraction comes from the small displacement alone.
MCPM_HD Cic cic_setup(const float* x, int64_t p, const Frame& fr, int nx, int ny, int nz) {
  Cic c;
  int sx, sy, sz;
  float rx, ry, rz;
  frame_site(fr, p, sx, sy, sz, rx, ry, rz);
  const float u0 = rx + x[0], u1 = ry + x[1], u2 = rz + x[2];
  float bx = floorf(u0), by = floorf(u1), bz = floorf(u2);
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

int interleave3(stream_t st, const float* planar3, float* mesh4, int64_t n) {
  f4* out = reinterpret_cast<f4*>(mesh4);
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t i) { out[i] = f4{planar3[i], planar3[n + i], planar3[2 * n + i], 0.0f}; });
  return rt_check("interleave3");
}

int deinterleave3(stream_t st, const float* mesh4, float* planar3, int64_t n) {
  const f4* in = reinterpret_cast<const f4*>(mesh4);
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t i) {
    f4 v = in[i];
    planar3[i] = v.x;
    planar3[n + i] = v.y;
    planar3[2 * n + i] = v.z;
  });
  return rt_check("deinterleave3");
}

// F = CIC read of mesh4.xyz at pos; vel' = alpha vel + beta F; pos' = pos + vel' drift  (pos_out/vel_out may alias)
// `zero` (optional): `nzero` floats cleared on the side, spread over the particle threads -- the density mesh the next
// step's scatter accumulates into; a few bytes per thread in a latency-bound kernel instead of a separate memset pass.
int kick_drift4(stream_t st, const float* pos, const float* vel, const float* fmesh4, int64_t np, int nx, int ny,
                int nz, float alpha, float beta, float drift, float* pos_out, float* vel_out, float* zero,
                int64_t nzero, const Frame* frp) {
  const f4* fm = reinterpret_cast<const f4*>(fmesh4);
#ifndef MCPM_HOSTEMU
  if (!zero) {  // particle arrays staged through shared memory by bulk async copies (cic4_tma.cu), where it applies
    const int r = kick_drift4_tma(st, pos, vel, fmesh4, np, nx, ny, nz, alpha, beta, drift, pos_out, vel_out, frp);
    if (r < 0) return MCPM_ECUDA;
    if (r == 1) return 0;
  }
#endif
  const Frame fr = frp ? *frp : Frame();
  launch_gather(st, np, [=] MCPM_LAMBDA(int64_t p) {
    float x[3] = {pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]};
    Cic c = cic_setup(x, p, fr, nx, ny, nz);
    const int64_t r00 = ((int64_t)c.i0 * ny + c.j0) * nz, r01 = ((int64_t)c.i0 * ny + c.j1) * nz;
    const int64_t r10 = ((int64_t)c.i1 * ny + c.j0) * nz, r11 = ((int64_t)c.i1 * ny + c.j1) * nz;
    const float gx = 1.0f - c.fx, gy = 1.0f - c.fy, gz = 1.0f - c.fz;
    const float w00 = gx * gy, w01 = gx * c.fy, w10 = c.fx * gy, w11 = c.fx * c.fy;
    f4 a0 = load4(fm + r00 + c.k0), a1 = load4(fm + r00 + c.k1);
    f4 b0 = load4(fm + r01 + c.k0), b1 = load4(fm + r01 + c.k1);
    f4 d0 = load4(fm + r10 + c.k0), d1 = load4(fm + r10 + c.k1);
    f4 e0 = load4(fm + r11 + c.k0), e1 = load4(fm + r11 + c.k1);
    // same association as the generic kernel: sum over corners of v * ((wx*wy)*wz)
    float f0 = a0.x * (w00 * gz) + a1.x * (w00 * c.fz) + b0.x * (w01 * gz) + b1.x * (w01 * c.fz) +
               d0.x * (w10 * gz) + d1.x * (w10 * c.fz) + e0.x * (w11 * gz) + e1.x * (w11 * c.fz);
    float f1 = a0.y * (w00 * gz) + a1.y * (w00 * c.fz) + b0.y * (w01 * gz) + b1.y * (w01 * c.fz) +
               d0.y * (w10 * gz) + d1.y * (w10 * c.fz) + e0.y * (w11 * gz) + e1.y * (w11 * c.fz);
    float f2 = a0.z * (w00 * gz) + a1.z * (w00 * c.fz) + b0.z * (w01 * gz) + b1.z * (w01 * c.fz) +
               d0.z * (w10 * gz) + d1.z * (w10 * c.fz) + e0.z * (w11 * gz) + e1.z * (w11 * c.fz);
    float v0 = alpha * vel[3 * p] + beta * f0;
    float v1 = alpha * vel[3 * p + 1] + beta * f1;
    float v2 = alpha * vel[3 * p + 2] + beta * f2;
    vel_out[3 * p] = v0;
    vel_out[3 * p + 1] = v1;
    vel_out[3 * p + 2] = v2;
    pos_out[3 * p] = x[0] + v0 * drift;
    pos_out[3 * p + 1] = x[1] + v1 * drift;
    pos_out[3 * p + 2] = x[2] + v2 * drift;
    if (zero)
      for (int64_t c = p; c < nzero; c += np) zero[c] = 0.0f;
  });
  return rt_check("kick_drift4");
}

// val = A + cb * B  (stored back into A when store != 0);  mesh4[cell] += scale * val * W  over the 8 CIC corners.
// Backward step: A = vbar, B = xbar, cb = drift, scale = beta   ->  vbar += xbar * drift ; phibar = paint(beta * vbar).
int paint3v4(stream_t st, const float* pos, float* A, const float* B, float cb, int store, float scale, int64_t np,
             int nx, int ny, int nz, float* mesh4, const Frame* frp) {
  f4* m = reinterpret_cast<f4*>(mesh4);
  const Frame fr = frp ? *frp : Frame();
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    float x[3] = {pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]};
    Cic c = cic_setup(x, p, fr, nx, ny, nz);
    float v0 = A[3 * p], v1 = A[3 * p + 1], v2 = A[3 * p + 2];
    if (B) {
      v0 += cb * B[3 * p];
      v1 += cb * B[3 * p + 1];
      v2 += cb * B[3 * p + 2];
      if (store) {
        A[3 * p] = v0;
        A[3 * p + 1] = v1;
        A[3 * p + 2] = v2;
      }
    }
    v0 *= scale;
    v1 *= scale;
    v2 *= scale;
    const int64_t r00 = ((int64_t)c.i0 * ny + c.j0) * nz, r01 = ((int64_t)c.i0 * ny + c.j1) * nz;
    const int64_t r10 = ((int64_t)c.i1 * ny + c.j0) * nz, r11 = ((int64_t)c.i1 * ny + c.j1) * nz;
    const float gx = 1.0f - c.fx, gy = 1.0f - c.fy, gz = 1.0f - c.fz;
    const float w00 = gx * gy, w01 = gx * c.fy, w10 = c.fx * gy, w11 = c.fx * c.fy;
    float w;
    w = w00 * gz;   atomic_add4(m + r00 + c.k0, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w00 * c.fz; atomic_add4(m + r00 + c.k1, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w01 * gz;   atomic_add4(m + r01 + c.k0, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w01 * c.fz; atomic_add4(m + r01 + c.k1, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w10 * gz;   atomic_add4(m + r10 + c.k0, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w10 * c.fz; atomic_add4(m + r10 + c.k1, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w11 * gz;   atomic_add4(m + r11 + c.k0, f4{v0 * w, v1 * w, v2 * w, 0.0f});
    w = w11 * c.fz; atomic_add4(m + r11 + c.k1, f4{v0 * w, v1 * w, v2 * w, 0.0f});
  });
  return rt_check("paint3v4");
}

// Backward force gather on (forces in mesh4.xyz, rhobar planar), CIC:
//   u(corner) = cscale * (cot . F(corner)) + rhobar(corner);   g_a = sum_corners u * dW_a * prod_{d != a} W_d
//   xbar += g ;  then (tail of the reverse step, `scale_cot`)  cot = alpha_tail * cot + dnext * xbar_new : the kick's
//   vbar *= alpha and the NEXT reverse step's leading vbar += xbar * drift in one pass, while both are in registers.
int read_grad4v(stream_t st, const float* pos, const float* fmesh4, const float* rhobar, float* cot, float cscale,
                int scale_cot, float alpha_tail, int64_t np, int nx, int ny, int nz, float* grad, int accumulate,
                float* zero, int64_t nzero, const Frame* frp, float dnext) {
  const f4* fm = reinterpret_cast<const f4*>(fmesh4);
#ifndef MCPM_HOSTEMU
  if (!zero) {
    const int r = read_grad4v_tma(st, pos, fmesh4, rhobar, cot, cscale, scale_cot, alpha_tail, np, nx, ny, nz, grad,
                                  accumulate, frp, dnext);
    if (r < 0) return MCPM_ECUDA;
    if (r == 1) return 0;
  }
#endif
  const Frame fr = frp ? *frp : Frame();
  launch_gather(st, np, [=] MCPM_LAMBDA(int64_t p) {
    float x[3] = {pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]};
    Cic c = cic_setup(x, p, fr, nx, ny, nz);
    const float q0 = cot[3 * p], q1 = cot[3 * p + 1], q2 = cot[3 * p + 2];
    const float c0 = cscale * q0, c1 = cscale * q1, c2 = cscale * q2;
    const int64_t r00 = ((int64_t)c.i0 * ny + c.j0) * nz, r01 = ((int64_t)c.i0 * ny + c.j1) * nz;
    const int64_t r10 = ((int64_t)c.i1 * ny + c.j0) * nz, r11 = ((int64_t)c.i1 * ny + c.j1) * nz;
    f4 t;
    float u000, u001, u010, u011, u100, u101, u110, u111;
    t = load4(fm + r00 + c.k0); u000 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r00 + c.k0];
    t = load4(fm + r00 + c.k1); u001 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r00 + c.k1];
    t = load4(fm + r01 + c.k0); u010 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r01 + c.k0];
    t = load4(fm + r01 + c.k1); u011 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r01 + c.k1];
    t = load4(fm + r10 + c.k0); u100 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r10 + c.k0];
    t = load4(fm + r10 + c.k1); u101 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r10 + c.k1];
    t = load4(fm + r11 + c.k0); u110 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r11 + c.k0];
    t = load4(fm + r11 + c.k1); u111 = c0 * t.x + c1 * t.y + c2 * t.z + rhobar[r11 + c.k1];
    // window values and derivatives: W = (1-f, f);  dW/dx = (-(f > 0), +1)   (sign(0) = 0 rule of jnp.abs, window.h)
    const float wx0 = 1.0f - c.fx, wx1 = c.fx, wy0 = 1.0f - c.fy, wy1 = c.fy, wz0 = 1.0f - c.fz, wz1 = c.fz;
    const float dx0 = c.fx > 0.0f ? -1.0f : 0.0f, dy0 = c.fy > 0.0f ? -1.0f : 0.0f, dz0 = c.fz > 0.0f ? -1.0f : 0.0f;
    float g0 = (dx0 * u000 + u100) * (wy0 * wz0) + (dx0 * u001 + u101) * (wy0 * wz1) + (dx0 * u010 + u110) * (wy1 * wz0) +
               (dx0 * u011 + u111) * (wy1 * wz1);
    float g1 = (dy0 * u000 + u010) * (wx0 * wz0) + (dy0 * u001 + u011) * (wx0 * wz1) + (dy0 * u100 + u110) * (wx1 * wz0) +
               (dy0 * u101 + u111) * (wx1 * wz1);
    float g2 = (dz0 * u000 + u001) * (wx0 * wy0) + (dz0 * u010 + u011) * (wx0 * wy1) + (dz0 * u100 + u101) * (wx1 * wy0) +
               (dz0 * u110 + u111) * (wx1 * wy1);
    float* g = grad + 3 * p;
    if (accumulate) {
      g0 += g[0];
      g1 += g[1];
      g2 += g[2];
    }
    g[0] = g0;
    g[1] = g1;
    g[2] = g2;
    if (scale_cot) {
      cot[3 * p] = alpha_tail * q0 + dnext * g0;
      cot[3 * p + 1] = alpha_tail * q1 + dnext * g1;
      cot[3 * p + 2] = alpha_tail * q2 + dnext * g2;
    }
    if (zero)  // the three meshes the next reverse step's scatter accumulates into (see kick_drift4)
      for (int64_t c = p; c < nzero; c += np) zero[c] = 0.0f;
  });
  return rt_check("read_grad4v");
}

}  // namespace mcpm
