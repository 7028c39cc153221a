// paint.cu -- mass assignment: paint (scatter-add), read (gather), read_grad (gather of the window gradient), and
// the fused readout + BullFrog kick + drift.  Reference: montecosmo/nbody.py:365-427 (paint/read), 933-951 (kick/drift).
//
// Layout: pos/vel [np,3] float32 AoS; mesh [nx,ny,nz] float32, z fastest.  One thread per particle.  Particles of a
// PM run are in Lagrangian-lattice order (z fastest), so the 32 lanes of a warp hold z-neighbours: their loads and
// their red.global.add.f32 land in a handful of consecutive 128-byte lines, which keeps both the LSU and the L2
// atomic units at line granularity instead of element granularity.
#include <limits>

#include "engine.h"
#include "window.h"

namespace mcpm {

struct MeshDims {
  int nx, ny, nz;
};

struct PosXform {
  float sx, sy, sz, shift;
  Frame fr;  // absolute or lattice-relative positions (frame.h)
  ObsShift obs;  // optional redshift-space shift added to the position before scale / shift (engine.h)
  ObsGen gen;    // ... or the general observation transform (obs.h)
};

// Transformed coordinate of particle p as an exact integer part b (its lattice site; 0 for absolute frames) plus a small
// float part u:  x' = b + u,  u = site remainder + pos * scale + shift.  pos may be NULL in a relative frame (particles
// on their sites).
// The particle's own part of that: integer site b, site remainder r, and position d (observation shift included).
MCPM_HD void load_site_disp(const float* pos, int64_t p, const PosXform& xf, int* b, float* r, float* d) {
  frame_site(xf.fr, p, b[0], b[1], b[2], r[0], r[1], r[2]);
  d[0] = pos ? pos[3 * p] : 0.0f;
  d[1] = pos ? pos[3 * p + 1] : 0.0f;
  d[2] = pos ? pos[3 * p + 2] : 0.0f;
  if (xf.obs.vel) {  // the same association as rsd_shift: pos + ((v . los) * coef) * los
    const float* v = xf.obs.vel + 3 * p;
    const float sh = (v[0] * xf.obs.lx + v[1] * xf.obs.ly + v[2] * xf.obs.lz) * xf.obs.coef;
    d[0] = d[0] + sh * xf.obs.lx;
    d[1] = d[1] + sh * xf.obs.ly;
    d[2] = d[2] + sh * xf.obs.lz;
  }
}
// Absolute position of particle p in the units of `pos` (the geometry of the observation transform needs it).
MCPM_HD void abs_pos(const PosXform& xf, const int* b, const float* r, const float* d, float* x) {
  x[0] = ((float)b[0] + r[0]) / xf.sx + d[0];  // relative frames give the site in target-mesh cells
  x[1] = ((float)b[1] + r[1]) / xf.sy + d[1];
  x[2] = ((float)b[2] + r[2]) / xf.sz + d[2];
}
// GEN: the general observation transform is compiled in (only the paint and its transpose are instantiated with it, so
// every other kernel -- and the paints of the step loop -- keep the register budget they had without it)
template <bool GEN = false>
MCPM_HD void load_pos(const float* pos, int64_t p, const PosXform& xf, int* b, float* u) {
  float r[3], d[3];
  load_site_disp(pos, p, xf, b, r, d);
  if (GEN && xf.gen.on) {
    float x[3], delta[3];
    ObsState s;
    abs_pos(xf, b, r, d, x);
    obs_forward(xf.gen, x, p, s, delta);
    d[0] += delta[0];
    d[1] += delta[1];
    d[2] += delta[2];
  }
  u[0] = r[0] + (d[0] * xf.sx + xf.shift);
  u[1] = r[1] + (d[1] * xf.sy + xf.shift);
  u[2] = r[2] + (d[2] * xf.sz + xf.shift);
}

template <int ORDER, class WIN = RectWin, bool GEN = false>
static void paint_impl(stream_t st, const float* pos, const float* weights, float wscalar, int64_t np, MeshDims n,
                       PosXform xf, float* mesh, WIN win = WIN()) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    int sb[3], fx, fy, fz;
    float su[3];
    load_pos<GEN>(pos, p, xf, sb, su);
    float wx[ORDER], wy[ORDER], wz[ORDER];
    win.template weights<ORDER>(su[0], fx, wx);
    win.template weights<ORDER>(su[1], fy, wy);
    win.template weights<ORDER>(su[2], fz, wz);
    fx = wrap_fast(fx + sb[0], n.nx);
    fy = wrap_fast(fy + sb[1], n.ny);
    fz = wrap_fast(fz + sb[2], n.nz);
    float wp = weights ? weights[p] * wscalar : wscalar;
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
      int ia = fx + a;
      ia = ia >= n.nx ? ia - n.nx : ia;
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        int ib = fy + b;
        ib = ib >= n.ny ? ib - n.ny : ib;
        float wab = wx[a] * wy[b];
        float* row = mesh + ((int64_t)ia * n.ny + ib) * n.nz;
#pragma unroll
        for (int c = 0; c < ORDER; ++c) {
          int ic = fz + c;
          ic = ic >= n.nz ? ic - n.nz : ic;
          atomic_add(row + ic, wp * (wab * wz[c]));
        }
      }
    }
  });
}

template <int ORDER, int NM, class WIN = RectWin>
static void read_impl(stream_t st, const float* pos, const float* mesh, int64_t np, MeshDims n, PosXform xf,
                      float* out, WIN win = WIN()) {
  const int64_t plane = (int64_t)n.nx * n.ny * n.nz;
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    int sb[3], fx, fy, fz;
    float su[3];
    load_pos(pos, p, xf, sb, su);
    float wx[ORDER], wy[ORDER], wz[ORDER];
    win.template weights<ORDER>(su[0], fx, wx);
    win.template weights<ORDER>(su[1], fy, wy);
    win.template weights<ORDER>(su[2], fz, wz);
    fx = wrap_fast(fx + sb[0], n.nx);
    fy = wrap_fast(fy + sb[1], n.ny);
    fz = wrap_fast(fz + sb[2], n.nz);
    float acc[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[m] = 0.0f;
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
      int ia = fx + a;
      ia = ia >= n.nx ? ia - n.nx : ia;
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        int ib = fy + b;
        ib = ib >= n.ny ? ib - n.ny : ib;
        float wab = wx[a] * wy[b];
        int64_t row = ((int64_t)ia * n.ny + ib) * n.nz;
#pragma unroll
        for (int c = 0; c < ORDER; ++c) {
          int ic = fz + c;
          ic = ic >= n.nz ? ic - n.nz : ic;
          float w = wab * wz[c];
#pragma unroll
          for (int m = 0; m < NM; ++m) acc[m] += mesh[m * plane + row + ic] * w;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) out[p * NM + m] = acc[m];
  });
}

struct MeshPtrs {
  const float* p[4];
};

// grad[p,a] (+)= gscale * scale_a * sum_m c[p,m] * sum_nbr mesh_m * dW_a * prod_{d!=a} W_d,
// c[p,m] = cscale * cot[p*ncot+m] for m < ncot, 1 otherwise;  gscale_p = gw ? gw[p] : 1.
// One gather serves the position-VJP of read (cot = out_bar), of paint (mesh = mesh_bar, gw = weights) and the
// backward force (3 force meshes with cot = beta*vbar, plus the density-cotangent mesh with c = 1).
template <int ORDER, class WIN = RectWin>
static void read_grad_impl(stream_t st, const float* pos, MeshPtrs ms, int nmesh, const float* cot, int ncot,
                           float cscale, const float* gw, int64_t np, MeshDims n, PosXform xf, float* grad,
                           int accumulate, WIN win = WIN()) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    int sb[3], fx, fy, fz;
    float su[3];
    load_pos(pos, p, xf, sb, su);
    float wx[ORDER], wy[ORDER], wz[ORDER], dx[ORDER], dy[ORDER], dz[ORDER];
    win.template weights_grad<ORDER>(su[0], fx, wx, dx);
    win.template weights_grad<ORDER>(su[1], fy, wy, dy);
    win.template weights_grad<ORDER>(su[2], fz, wz, dz);
    fx = wrap_fast(fx + sb[0], n.nx);
    fy = wrap_fast(fy + sb[1], n.ny);
    fz = wrap_fast(fz + sb[2], n.nz);
    float c4[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) c4[m] = (m < ncot && cot) ? cscale * cot[p * ncot + m] : 1.0f;
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
      int ia = fx + a;
      ia = ia >= n.nx ? ia - n.nx : ia;
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        int ib = fy + b;
        ib = ib >= n.ny ? ib - n.ny : ib;
        int64_t row = ((int64_t)ia * n.ny + ib) * n.nz;
#pragma unroll
        for (int c = 0; c < ORDER; ++c) {
          int ic = fz + c;
          ic = ic >= n.nz ? ic - n.nz : ic;
          float v = 0.0f;
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (m < nmesh) v += c4[m] * ms.p[m][row + ic];
          g0 += v * (dx[a] * wy[b] * wz[c]);
          g1 += v * (wx[a] * dy[b] * wz[c]);
          g2 += v * (wx[a] * wy[b] * dz[c]);
        }
      }
    }
    float gs = gw ? gw[p] : 1.0f;
    g0 *= xf.sx * gs;
    g1 *= xf.sy * gs;
    g2 *= xf.sz * gs;
    float* g = grad + 3 * p;
    if (accumulate) {
      g[0] += g0;
      g[1] += g1;
      g[2] += g2;
    } else {
      g[0] = g0;
      g[1] = g1;
      g[2] = g2;
    }
  });
}

// Three-channel scatter for the adjoints: mesh_m[...] += (ca*A[p,m] + cb*B[p,m]) * W, m = 0..2 (B may be NULL).
// The index / weight computation is shared by the three channels.
template <int ORDER>
static void paint3_impl(stream_t st, const float* pos, const float* A, float ca, const float* B, float cb,
                        int64_t np, MeshDims n, PosXform xf, float* mesh3) {
  const int64_t plane = (int64_t)n.nx * n.ny * n.nz;
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    int sb[3], fx, fy, fz;
    float su[3];
    load_pos(pos, p, xf, sb, su);
    float wx[ORDER], wy[ORDER], wz[ORDER];
    window_weights<ORDER>(su[0], fx, wx);
    window_weights<ORDER>(su[1], fy, wy);
    window_weights<ORDER>(su[2], fz, wz);
    fx = wrap_fast(fx + sb[0], n.nx);
    fy = wrap_fast(fy + sb[1], n.ny);
    fz = wrap_fast(fz + sb[2], n.nz);
    float v0 = ca * A[3 * p], v1 = ca * A[3 * p + 1], v2 = ca * A[3 * p + 2];
    if (B) {
      v0 += cb * B[3 * p];
      v1 += cb * B[3 * p + 1];
      v2 += cb * B[3 * p + 2];
    }
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
      int ia = fx + a;
      ia = ia >= n.nx ? ia - n.nx : ia;
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        int ib = fy + b;
        ib = ib >= n.ny ? ib - n.ny : ib;
        float wab = wx[a] * wy[b];
        float* row = mesh3 + ((int64_t)ia * n.ny + ib) * n.nz;
#pragma unroll
        for (int c = 0; c < ORDER; ++c) {
          int ic = fz + c;
          ic = ic >= n.nz ? ic - n.nz : ic;
          float w = wab * wz[c];
          atomic_add(row + ic, v0 * w);
          atomic_add(row + plane + ic, v1 * w);
          atomic_add(row + 2 * plane + ic, v2 * w);
        }
      }
    }
  });
}

// VJP of paint w.r.t. weights and positions from the mesh cotangent, one gather:
//   wbar[p] (+)= wscalar * sum mesh_bar * W ;  posbar[p,a] (+)= w_p * scale_a * sum mesh_bar * dW_a ...
// nshift > 1: the transposes of `nshift` interlaced paints in ONE gather -- mesh cotangent t = mbar + t * mesh_stride was
// painted at shift xf.shift + t / nshift (nufft, nbody.py:524) -- so the particle arrays are read once and every output
// is written once (two passes over the particles cost 0.77 ms of a 256^3 evaluation, with accumulating outputs).
template <int ORDER, class WIN = RectWin, bool GEN = false>
static void paint_vjp_impl(stream_t st, const float* pos, const float* weights, float wscalar, const float* mbar,
                           int64_t np, MeshDims n, PosXform xf, float* posbar, float* wbar, int accumulate,
                           float* velbar, int nshift, int64_t mesh_stride, WIN win = WIN()) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    int sb[3];
    float sr[3], sd[3];
    load_site_disp(pos, p, xf, sb, sr, sd);
    ObsState os;
    if (GEN && xf.gen.on) {  // general observation transform: keep its state for the transpose below
      float x[3], delta[3];
      abs_pos(xf, sb, sr, sd, x);
      obs_forward(xf.gen, x, p, os, delta);
      sd[0] += delta[0];
      sd[1] += delta[1];
      sd[2] += delta[2];
    }
    float r = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
    for (int t = 0; t < nshift; ++t) {
      const float sh = xf.shift + (float)t / (float)nshift;  // exactly the shifts the forward paints used (engine.cu: nufft)
      const float su[3] = {sr[0] + (sd[0] * xf.sx + sh), sr[1] + (sd[1] * xf.sy + sh), sr[2] + (sd[2] * xf.sz + sh)};
      const float* mb = mbar + (int64_t)t * mesh_stride;
      int fx, fy, fz;
      float wx[ORDER], wy[ORDER], wz[ORDER], dx[ORDER], dy[ORDER], dz[ORDER];
      win.template weights_grad<ORDER>(su[0], fx, wx, dx);
      win.template weights_grad<ORDER>(su[1], fy, wy, dy);
      win.template weights_grad<ORDER>(su[2], fz, wz, dz);
      fx = wrap_fast(fx + sb[0], n.nx);
      fy = wrap_fast(fy + sb[1], n.ny);
      fz = wrap_fast(fz + sb[2], n.nz);
#pragma unroll
      for (int a = 0; a < ORDER; ++a) {
        int ia = fx + a;
        ia = ia >= n.nx ? ia - n.nx : ia;
#pragma unroll
        for (int b = 0; b < ORDER; ++b) {
          int ib = fy + b;
          ib = ib >= n.ny ? ib - n.ny : ib;
          int64_t row = ((int64_t)ia * n.ny + ib) * n.nz;
#pragma unroll
          for (int c = 0; c < ORDER; ++c) {
            int ic = fz + c;
            ic = ic >= n.nz ? ic - n.nz : ic;
            float v = mb[row + ic];
            r += v * (wx[a] * wy[b] * wz[c]);
            g0 += v * (dx[a] * wy[b] * wz[c]);
            g1 += v * (wx[a] * dy[b] * wz[c]);
            g2 += v * (wx[a] * wy[b] * dz[c]);
          }
        }
      }
    }
    float wp = weights ? weights[p] * wscalar : wscalar;
    if (wbar) wbar[p] = (accumulate ? wbar[p] : 0.0f) + r * wscalar;
    if (GEN && xf.gen.on) {  // cotangent of the observed position -> positions, velocities, velocity bias, parameters
      const float gd[3] = {g0 * xf.sx * wp, g1 * xf.sy * wp, g2 * xf.sz * wp};
      float xadd[3], vb[3];
      obs_transpose(xf.gen, os, p, gd, xadd, vb, accumulate != 0);
      if (posbar) {
        float* g = posbar + 3 * p;
        g[0] = (accumulate ? g[0] : 0.0f) + (gd[0] + xadd[0]);
        g[1] = (accumulate ? g[1] : 0.0f) + (gd[1] + xadd[1]);
        g[2] = (accumulate ? g[2] : 0.0f) + (gd[2] + xadd[2]);
      }
      if (velbar) {
        float* o = velbar + 3 * p;
        o[0] = (accumulate ? o[0] : 0.0f) + vb[0];
        o[1] = (accumulate ? o[1] : 0.0f) + vb[1];
        o[2] = (accumulate ? o[2] : 0.0f) + vb[2];
      }
      return;
    }
    if (posbar) {
      float* g = posbar + 3 * p;
      g[0] = (accumulate ? g[0] : 0.0f) + g0 * xf.sx * wp;
      g[1] = (accumulate ? g[1] : 0.0f) + g1 * xf.sy * wp;
      g[2] = (accumulate ? g[2] : 0.0f) + g2 * xf.sz * wp;
    }
    if (velbar) {  // transpose of the fused redshift-space shift: velbar (+)= coef * (xbar . los) * los
      const float sh = (g0 * xf.sx * wp * xf.obs.lx + g1 * xf.sy * wp * xf.obs.ly + g2 * xf.sz * wp * xf.obs.lz) * xf.obs.coef;
      float* o = velbar + 3 * p;
      o[0] = (accumulate ? o[0] : 0.0f) + sh * xf.obs.lx;
      o[1] = (accumulate ? o[1] : 0.0f) + sh * xf.obs.ly;
      o[2] = (accumulate ? o[2] : 0.0f) + sh * xf.obs.lz;
    }
  });
}

// F = read(pos, fmesh[3]); vel = alpha*vel + beta*F; pos += vel*drift   (nbody.py:933-951)
template <int ORDER>
static void kick_drift_impl(stream_t st, const float* pos, const float* vel, const float* fmesh, int64_t np,
                            MeshDims n, PosXform xf, float alpha, float beta, float drift, float* pos_out,
                            float* vel_out, float* force_out) {
  const int64_t plane = (int64_t)n.nx * n.ny * n.nz;
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    const float* x = pos + 3 * p;
    const float* v = vel + 3 * p;
    float x0 = x[0], x1 = x[1], x2 = x[2];  // absolute position, or displacement from the lattice site (updated as such)
    int sb[3], fx, fy, fz;
    float su[3];
    load_pos(pos, p, xf, sb, su);
    float wx[ORDER], wy[ORDER], wz[ORDER];
    window_weights<ORDER>(su[0], fx, wx);
    window_weights<ORDER>(su[1], fy, wy);
    window_weights<ORDER>(su[2], fz, wz);
    fx = wrap_fast(fx + sb[0], n.nx);
    fy = wrap_fast(fy + sb[1], n.ny);
    fz = wrap_fast(fz + sb[2], n.nz);
    float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
      int ia = fx + a;
      ia = ia >= n.nx ? ia - n.nx : ia;
#pragma unroll
      for (int b = 0; b < ORDER; ++b) {
        int ib = fy + b;
        ib = ib >= n.ny ? ib - n.ny : ib;
        float wab = wx[a] * wy[b];
        int64_t row = ((int64_t)ia * n.ny + ib) * n.nz;
#pragma unroll
        for (int c = 0; c < ORDER; ++c) {
          int ic = fz + c;
          ic = ic >= n.nz ? ic - n.nz : ic;
          float w = wab * wz[c];
          f0 += fmesh[row + ic] * w;
          f1 += fmesh[plane + row + ic] * w;
          f2 += fmesh[2 * plane + row + ic] * w;
        }
      }
    }
    float v0 = alpha * v[0] + beta * f0;
    float v1 = alpha * v[1] + beta * f1;
    float v2 = alpha * v[2] + beta * f2;
    vel_out[3 * p] = v0;
    vel_out[3 * p + 1] = v1;
    vel_out[3 * p + 2] = v2;
    pos_out[3 * p] = x0 + v0 * drift;
    pos_out[3 * p + 1] = x1 + v1 * drift;
    pos_out[3 * p + 2] = x2 + v2 * drift;
    if (force_out) {
      force_out[3 * p] = f0;
      force_out[3 * p + 1] = f1;
      force_out[3 * p + 2] = f2;
    }
  });
}

// ------------------------------------------------------------------------------------------------ dispatch
static int check_mesh(int nx, int ny, int nz, int order) {
  if (nx <= 0 || ny <= 0 || nz <= 0) {
    set_error("mesh dimensions must be positive");
    return MCPM_EINVAL;
  }
  if (order < 1 || order > 4) {
    set_error("assignment order must be 1 (NGP), 2 (CIC), 3 (TSC) or 4 (PCS)");
    return MCPM_EINVAL;
  }
  if (nx < order || ny < order || nz < order) {
    set_error("mesh sides must be at least the assignment order");
    return MCPM_EINVAL;
  }
  return 0;
}

static PosXform make_xform(const float* scale, float shift, const Frame* fr = nullptr, const ObsShift* obs = nullptr) {
  PosXform xf = {1.0f, 1.0f, 1.0f, shift, fr ? *fr : Frame(), obs ? *obs : ObsShift(), ObsGen()};
  if (obs && obs->gen) {
    xf.gen = *obs->gen;
    xf.obs = ObsShift();
  }
  if (scale) {
    xf.sx = scale[0];
    xf.sy = scale[1];
    xf.sz = scale[2];
  }
  return xf;
}

int paint(stream_t st, const float* pos, const float* weights, float wscalar, int64_t np, int nx, int ny, int nz,
          int order, const float* scale, float shift, float* mesh, int accumulate, float kb_kcut, const Frame* fr,
          const ObsShift* obs) {
  if (int e = check_mesh(nx, ny, nz, order)) return e;
  if (!mesh || (np > 0 && !pos && !(fr && fr->rel))) {
    set_error("paint: null pointer");
    return MCPM_EINVAL;
  }
  MeshDims n = {nx, ny, nz};
  PosXform xf = make_xform(scale, shift, fr, obs);
  if (!accumulate) rt_memset(mesh, 0, sizeof(float) * (size_t)nx * ny * nz, st);
  if (xf.gen.on) {  // the instantiations with the general observation transform
    if (kb_kcut > 0.0f) {
      KbWin kb = make_kbwin(order, kb_kcut);
      switch (order) {
        case 1: paint_impl<1, KbWin, true>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
        case 2: paint_impl<2, KbWin, true>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
        case 3: paint_impl<3, KbWin, true>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
        default: paint_impl<4, KbWin, true>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
      }
    } else {
      switch (order) {
        case 1: paint_impl<1, RectWin, true>(st, pos, weights, wscalar, np, n, xf, mesh); break;
        case 2: paint_impl<2, RectWin, true>(st, pos, weights, wscalar, np, n, xf, mesh); break;
        case 3: paint_impl<3, RectWin, true>(st, pos, weights, wscalar, np, n, xf, mesh); break;
        default: paint_impl<4, RectWin, true>(st, pos, weights, wscalar, np, n, xf, mesh); break;
      }
    }
    return rt_check("paint");
  }
  if (kb_kcut > 0.0f) {
    KbWin kb = make_kbwin(order, kb_kcut);
    switch (order) {
      case 1: paint_impl<1>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
      case 2: paint_impl<2>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
      case 3: paint_impl<3>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
      default: paint_impl<4>(st, pos, weights, wscalar, np, n, xf, mesh, kb); break;
    }
    return rt_check("paint");
  }
  switch (order) {
    case 1: paint_impl<1>(st, pos, weights, wscalar, np, n, xf, mesh); break;
    case 2: paint_impl<2>(st, pos, weights, wscalar, np, n, xf, mesh); break;
    case 3: paint_impl<3>(st, pos, weights, wscalar, np, n, xf, mesh); break;
    default: paint_impl<4>(st, pos, weights, wscalar, np, n, xf, mesh); break;
  }
  return rt_check("paint");
}

template <int ORDER, class WIN = RectWin>
static int read_nm(stream_t st, const float* pos, const float* mesh, int nmesh, int64_t np, MeshDims n, PosXform xf,
                   float* out, WIN win = WIN()) {
  switch (nmesh) {
    case 1: read_impl<ORDER, 1>(st, pos, mesh, np, n, xf, out, win); return 0;
    case 2: read_impl<ORDER, 2>(st, pos, mesh, np, n, xf, out, win); return 0;
    case 3: read_impl<ORDER, 3>(st, pos, mesh, np, n, xf, out, win); return 0;
    case 4: read_impl<ORDER, 4>(st, pos, mesh, np, n, xf, out, win); return 0;
    case 7: read_impl<ORDER, 7>(st, pos, mesh, np, n, xf, out, win); return 0;  // the fields of lagrangian_bias (bias.cu)
    case 9: read_impl<ORDER, 9>(st, pos, mesh, np, n, xf, out, win); return 0;  // ... with primordial non-Gaussianity
  }
  set_error("read: nmesh must be 1..4, 7 or 9");
  return MCPM_EINVAL;
}

int read(stream_t st, const float* pos, const float* mesh, int nmesh, int64_t np, int nx, int ny, int nz, int order,
         const float* scale, float shift, float* out, float kb_kcut, const Frame* fr) {
  if (int e = check_mesh(nx, ny, nz, order)) return e;
  if (np > 0 && ((!pos && !(fr && fr->rel)) || !mesh || !out)) {
    set_error("read: null pointer");
    return MCPM_EINVAL;
  }
  MeshDims n = {nx, ny, nz};
  PosXform xf = make_xform(scale, shift, fr);
  int e;
  if (kb_kcut > 0.0f) {
    KbWin kb = make_kbwin(order, kb_kcut);
    switch (order) {
      case 1: e = read_nm<1>(st, pos, mesh, nmesh, np, n, xf, out, kb); break;
      case 2: e = read_nm<2>(st, pos, mesh, nmesh, np, n, xf, out, kb); break;
      case 3: e = read_nm<3>(st, pos, mesh, nmesh, np, n, xf, out, kb); break;
      default: e = read_nm<4>(st, pos, mesh, nmesh, np, n, xf, out, kb); break;
    }
    return e ? e : rt_check("read");
  }
  switch (order) {
    case 1: e = read_nm<1>(st, pos, mesh, nmesh, np, n, xf, out); break;
    case 2: e = read_nm<2>(st, pos, mesh, nmesh, np, n, xf, out); break;
    case 3: e = read_nm<3>(st, pos, mesh, nmesh, np, n, xf, out); break;
    default: e = read_nm<4>(st, pos, mesh, nmesh, np, n, xf, out); break;
  }
  return e ? e : rt_check("read");
}

int read_grad(stream_t st, const float* pos, const float* const* meshes, int nmesh, const float* cot, int ncot,
              float cscale, const float* gw, int64_t np, int nx, int ny, int nz, int order, const float* scale,
              float shift, float* grad, int accumulate, float kb_kcut, const Frame* fr) {
  if (int e = check_mesh(nx, ny, nz, order)) return e;
  if (nmesh < 1 || nmesh > 4 || ncot < 0 || ncot > nmesh) {
    set_error("read_grad: nmesh must be 1..4 and ncot <= nmesh");
    return MCPM_EINVAL;
  }
  if (np > 0 && ((!pos && !(fr && fr->rel)) || !meshes || !grad)) {
    set_error("read_grad: null pointer");
    return MCPM_EINVAL;
  }
  MeshDims n = {nx, ny, nz};
  PosXform xf = make_xform(scale, shift, fr);
  MeshPtrs ms = {{nullptr, nullptr, nullptr, nullptr}};
  for (int m = 0; m < nmesh; ++m) ms.p[m] = meshes[m];
  if (kb_kcut > 0.0f) {
    KbWin kb = make_kbwin(order, kb_kcut);
    switch (order) {
      case 1: read_grad_impl<1>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate, kb); break;
      case 2: read_grad_impl<2>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate, kb); break;
      case 3: read_grad_impl<3>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate, kb); break;
      default: read_grad_impl<4>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate, kb); break;
    }
    return rt_check("read_grad");
  }
  switch (order) {
    case 1: read_grad_impl<1>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate); break;
    case 2: read_grad_impl<2>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate); break;
    case 3: read_grad_impl<3>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate); break;
    default: read_grad_impl<4>(st, pos, ms, nmesh, cot, ncot, cscale, gw, np, n, xf, grad, accumulate); break;
  }
  return rt_check("read_grad");
}

int paint3(stream_t st, const float* pos, const float* A, float ca, const float* B, float cb, int64_t np, int nx,
           int ny, int nz, int order, float* mesh3, int accumulate, const Frame* fr) {
  if (int e = check_mesh(nx, ny, nz, order)) return e;
  if (!mesh3 || (np > 0 && ((!pos && !(fr && fr->rel)) || !A))) {
    set_error("paint3: null pointer");
    return MCPM_EINVAL;
  }
  MeshDims n = {nx, ny, nz};
  PosXform xf = make_xform(nullptr, 0.0f, fr);
  if (!accumulate) rt_memset(mesh3, 0, sizeof(float) * 3 * (size_t)nx * ny * nz, st);
  switch (order) {
    case 1: paint3_impl<1>(st, pos, A, ca, B, cb, np, n, xf, mesh3); break;
    case 2: paint3_impl<2>(st, pos, A, ca, B, cb, np, n, xf, mesh3); break;
    case 3: paint3_impl<3>(st, pos, A, ca, B, cb, np, n, xf, mesh3); break;
    default: paint3_impl<4>(st, pos, A, ca, B, cb, np, n, xf, mesh3); break;
  }
  return rt_check("paint3");
}

int paint_vjp(stream_t st, const float* pos, const float* weights, float wscalar, const float* mbar, int64_t np,
              int nx, int ny, int nz, int order, const float* scale, float shift, float* posbar, float* wbar,
              int accumulate, float kb_kcut, const Frame* fr, const ObsShift* obs, float* velbar, int nshift,
              int64_t mesh_stride) {
  if (int e = check_mesh(nx, ny, nz, order)) return e;
  if (np > 0 && ((!pos && !(fr && fr->rel)) || !mbar)) {
    set_error("paint_vjp: null pointer");
    return MCPM_EINVAL;
  }
  MeshDims n = {nx, ny, nz};
  PosXform xf = make_xform(scale, shift, fr, obs);
  if (!(obs && (obs->vel || (obs->gen && obs->gen->rsd)))) velbar = nullptr;
  if (nshift < 1) nshift = 1;
  if (xf.gen.on) {  // the instantiations with the general observation transform
    if (kb_kcut > 0.0f) {
      KbWin kb = make_kbwin(order, kb_kcut);
      switch (order) {
        case 1: paint_vjp_impl<1, KbWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
        case 2: paint_vjp_impl<2, KbWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
        case 3: paint_vjp_impl<3, KbWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
        default: paint_vjp_impl<4, KbWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
      }
    } else {
      switch (order) {
        case 1: paint_vjp_impl<1, RectWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
        case 2: paint_vjp_impl<2, RectWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
        case 3: paint_vjp_impl<3, RectWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
        default: paint_vjp_impl<4, RectWin, true>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
      }
    }
    return rt_check("paint_vjp");
  }
  if (kb_kcut > 0.0f) {
    KbWin kb = make_kbwin(order, kb_kcut);
    switch (order) {
      case 1: paint_vjp_impl<1>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
      case 2: paint_vjp_impl<2>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
      case 3: paint_vjp_impl<3>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
      default: paint_vjp_impl<4>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride, kb); break;
    }
    return rt_check("paint_vjp");
  }
  switch (order) {
    case 1: paint_vjp_impl<1>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
    case 2: paint_vjp_impl<2>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
    case 3: paint_vjp_impl<3>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
    default: paint_vjp_impl<4>(st, pos, weights, wscalar, mbar, np, n, xf, posbar, wbar, accumulate, velbar, nshift, mesh_stride); break;
  }
  return rt_check("paint_vjp");
}

int kick_drift(stream_t st, const float* pos, const float* vel, const float* fmesh3, int64_t np, int nx, int ny,
               int nz, int order, float alpha, float beta, float drift, float* pos_out, float* vel_out,
               float* force_out, const Frame* fr) {
  if (int e = check_mesh(nx, ny, nz, order)) return e;
  if (np > 0 && (!pos || !vel || !fmesh3 || !pos_out || !vel_out)) {
    set_error("kick_drift: null pointer");
    return MCPM_EINVAL;
  }
  MeshDims n = {nx, ny, nz};
  PosXform xf = make_xform(nullptr, 0.0f, fr);
  switch (order) {
    case 1: kick_drift_impl<1>(st, pos, vel, fmesh3, np, n, xf, alpha, beta, drift, pos_out, vel_out, force_out); break;
    case 2: kick_drift_impl<2>(st, pos, vel, fmesh3, np, n, xf, alpha, beta, drift, pos_out, vel_out, force_out); break;
    case 3: kick_drift_impl<3>(st, pos, vel, fmesh3, np, n, xf, alpha, beta, drift, pos_out, vel_out, force_out); break;
    default: kick_drift_impl<4>(st, pos, vel, fmesh3, np, n, xf, alpha, beta, drift, pos_out, vel_out, force_out); break;
  }
  return rt_check("kick_drift");
}

// ---- particles on the cells of the mesh (regular_pos(mesh_shape), bricks.py:593-603): NGP and CIC reads return the cell
// value itself (a particle at an integer position has weight 1 on its base cell), so lpt's reads at q (nbody.py:984-985
// pass read_order = 1) and their transposes are layout changes between planar meshes and the [np, 3] particle array.
int read_sites3(stream_t st, const float* planar3, int64_t n, float* out) {
  launch_1d(st, 3 * n, [=] MCPM_LAMBDA(int64_t i) {
    const int64_t p = i / 3;
    out[i] = planar3[(i - 3 * p) * n + p];
  });
  return rt_check("read_sites3");
}
// mesh3[c][p] (+)= ca * A[p, c] + cb * B[p, c]
int paint_sites3(stream_t st, const float* A, float ca, const float* B, float cb, int64_t n, float* mesh3, int accumulate) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t p) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = ca * A[3 * p + c] + (B ? cb * B[3 * p + c] : 0.0f);
      mesh3[c * n + p] = accumulate ? mesh3[c * n + p] + v : v;
    }
  });
  return rt_check("paint_sites3");
}

// ---- small particle-array kernels ------------------------------------------------------------------------------
// out = a + b * s (drift, nbody.py:942-944); out may alias a
int axpy3(stream_t st, const float* a, const float* b, float s, int64_t n3, float* out) {
  launch_1d(st, n3, [=] MCPM_LAMBDA(int64_t i) { out[i] = a[i] + b[i] * s; });
  return rt_check("axpy3");
}

// lpt combine (nbody.py:656-665).  dpos/vel may alias f2/f1.
int lpt_combine(stream_t st, const float* pos, const float* f1, const float* f2, float d1, float d2, float dv2,
                int64_t np, float* dpos, float* vel, float* pos_out) {
  launch_1d(st, 3 * np, [=] MCPM_LAMBDA(int64_t i) {
    float a = f1[i], b = f2 ? f2[i] : 0.0f;
    float dp = d1 * a - d2 * b;
    float v = a - dv2 * b;
    if (dpos) dpos[i] = dp;
    if (vel) vel[i] = v;
    if (pos_out) pos_out[i] = pos[i] + dp;
  });
  return rt_check("lpt_combine");
}

// out = a*x + b*y + c  (y may be NULL)
int axpby(stream_t st, const float* x, float a, const float* y, float b, float c, int64_t n, float* out) {
  launch_1d(st, n, [=] MCPM_LAMBDA(int64_t i) { out[i] = a * x[i] + (y ? b * y[i] : 0.0f) + c; });
  return rt_check("axpby");
}

// Functions of the comoving distance at the particles -- the light cone (bricks.py:747-766 followed by a2g, a2f, ...):
// out[p, k] = tab_k(r_p), r_p = |q_p| (curved sky) or |q_p . l| (flat), q = pos * cell + origin; tabs [ntab][nt].
int radial_tables(stream_t st, const float* pos, int64_t np, ObsGen g, int ntab, const float* tabs, float* out) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    ObsTabCell c;
    const float r = obs_radius(g, pos + 3 * p, nullptr, nullptr);
    for (int k = 0; k < ntab; ++k) out[p * ntab + k] = obs_interp(tabs + (int64_t)k * g.nt, g, r, c);
  });
  return rt_check("radial_tables");
}
// transpose: tabbar [kObsSlots][ntab][nt] float64 (partial rows, zeroed by the caller, summed by obs_reduce_slots);
// posbar [np,3] (nullable) = sum_k outbar[p,k] tab_k'(r) dr/dpos
int radial_tables_vjp(stream_t st, const float* pos, int64_t np, ObsGen g, int ntab, const float* tabs, const float* outbar,
                      float* posbar, double* tabbar) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    ObsTabCell c;
    float dir[3], sgn;
    const float r = obs_radius(g, pos + 3 * p, dir, &sgn);
    double* row = tabbar ? tabbar + (size_t)(p & (kObsSlots - 1)) * (size_t)ntab * g.nt : nullptr;
    float rb = 0.0f;
    for (int k = 0; k < ntab; ++k) {
      obs_interp(tabs + (int64_t)k * g.nt, g, r, c);
      const float ob = outbar[p * ntab + k];
      rb += ob * c.slope;
      if (row) {
        atomic_add(row + (size_t)k * g.nt + c.i, (double)((1.0f - c.w) * ob));
        atomic_add(row + (size_t)k * g.nt + c.i + 1, (double)(c.w * ob));
      }
    }
    if (posbar) {
      posbar[3 * p] = rb * sgn * dir[0] * g.cx;
      posbar[3 * p + 1] = rb * sgn * dir[1] * g.cy;
      posbar[3 * p + 2] = rb * sgn * dir[2] * g.cz;
    }
  });
  return rt_check("radial_tables_vjp");
}

// parameter cotangents of the general observation transform: the kObsSlots partial rows summed into row 0
int obs_reduce_slots(stream_t st, double* parbar, int64_t row) {
  launch_1d(st, row, [=] MCPM_LAMBDA(int64_t k) {
    double acc = parbar[k];
    for (int s = 1; s < kObsSlots; ++s) acc += parbar[(int64_t)s * row + k];
    parbar[k] = acc;
  });
  return rt_check("obs_reduce_slots");
}

// flat-sky redshift-space shift in cell units (bricks.py:781-792): pos_out = pos + (vel . los) * coef * los
int rsd_shift(stream_t st, const float* pos, const float* vel, float lx, float ly, float lz, float coef, int64_t np,
              float* pos_out) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    const float* v = vel + 3 * p;
    float s = (v[0] * lx + v[1] * ly + v[2] * lz) * coef;
    pos_out[3 * p] = pos[3 * p] + s * lx;
    pos_out[3 * p + 1] = pos[3 * p + 1] + s * ly;
    pos_out[3 * p + 2] = pos[3 * p + 2] + s * lz;
  });
  return rt_check("rsd_shift");
}

// its VJP w.r.t. vel: velbar (+)= coef * (posbar . los) * los   (posbar passes through unchanged)
int rsd_shift_vjp(stream_t st, const float* posbar, float lx, float ly, float lz, float coef, int64_t np,
                  float* velbar, int accumulate) {
  launch_1d(st, np, [=] MCPM_LAMBDA(int64_t p) {
    const float* g = posbar + 3 * p;
    float s = (g[0] * lx + g[1] * ly + g[2] * lz) * coef;
    float* o = velbar + 3 * p;
    o[0] = (accumulate ? o[0] : 0.0f) + s * lx;
    o[1] = (accumulate ? o[1] : 0.0f) + s * ly;
    o[2] = (accumulate ? o[2] : 0.0f) + s * lz;
  });
  return rt_check("rsd_shift_vjp");
}

// out[0] = max(out[0], max_i |x[i * stride]|); out[0] >= 0 on entry.  The halo guard of a slab-decomposed rank (the largest
// x-displacement of its particles, every step): as torch ops on the strided column it cost a 95 us elementwise pass and a
// reduction per step, 2.6 ms per 256^3 evaluation (profiles/r2_slab_one_rank.txt).
int absmax_strided(stream_t st, const float* x, int64_t n, int stride, float* out) {
#ifdef MCPM_HOSTEMU
  float m = out[0];
  int nan = m != m;
#pragma omp parallel for reduction(max : m) reduction(| : nan) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float a = std::fabs(x[i * stride]);
    if (a != a) nan |= 1;
    else if (a > m) m = a;
  }
  out[0] = nan ? std::numeric_limits<float>::quiet_NaN() : m;
  (void)st;
  return 0;
#else
  const int64_t chunk = 4096, nchunks = (n + chunk - 1) / chunk;  // one warp per chunk
  launch_1d(st, nchunks * 32, [=] MCPM_LAMBDA(int64_t t) {
    const int64_t c = t >> 5;
    const int lane = (int)(t & 31);
    const int64_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
    float m = 0.0f;
    for (int64_t i = lo + lane; i < hi; i += 32) {  // fmaxf would drop a NaN: keep it (a != a), and keep it once held
      const float a = fabsf(x[i * stride]);
      m = (a > m || a != a) ? a : m;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float a = __shfl_down_sync(0xffffffffu, m, o);
      m = (a > m || a != a) ? a : m;
    }
    // non-negative floats order like their bit patterns; NaN (0x7fc00000) compares above every finite value and sticks
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
  });
  return rt_check("absmax_strided");
#endif
}

// out[0] += sum_i a[i] * b[i] in float64 (coefficient cotangents)
int dot_accum(stream_t st, const float* a, const float* b, int64_t n, double scale, double* out) {
#ifdef MCPM_HOSTEMU
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += (double)a[i] * (double)b[i];
  *out += scale * s;
  return 0;
#else
  const int64_t chunk = 4096;  // one warp-reduced double atomic per 4096 products
  const int64_t nchunks = (n + chunk - 1) / chunk;
  launch_1d(st, nchunks * 32, [=] MCPM_LAMBDA(int64_t t) {
    int64_t c = t >> 5;
    int lane = (int)(t & 31);
    int64_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
    double s = 0.0;
    for (int64_t i = lo + lane; i < hi; i += 32) s += (double)a[i] * (double)b[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) atomic_add(out, scale * s);
  });
  return rt_check("dot_accum");
#endif
}

}  // namespace mcpm
