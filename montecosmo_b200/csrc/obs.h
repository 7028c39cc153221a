// obs.h -- the observation transform of the general model applied INSIDE the final paint: cell -> physical position,
// line of sight and scale factor, redshift-space distortion, Alcock-Paczynski rescaling, physical -> cell
// (model.py:780-799 with bricks.py:628-660, 747-766, 781-792, 795-813, 847-856), as one per-particle function and its
// hand-derived transpose.  The transformed positions are never written to memory.
//
// Everything is evaluated in the UNROTATED box frame, the trick radius_mesh uses (bricks.py:676-687): with R the box
// rotation, c the box centre and y = pos * cell - box / 2 the reference forms p = R y + c; then |p| = |y + R^T c| and
// l . p = (R^T l) . (y + R^T c), so with
//     q = pos * cell + o,   o = R^T c - box / 2
// the physical chain reads (l~ = R^T l, u~ = R^T u)
//     curved sky:  r = |q|,  l~ = q / r            flat sky:  l~ = R^T c / |c|,  r = |q . l~|
//     u~ = (vel * cell) * gf(r) + R^T dvel         gf = D f at the particle's scale factor (light cone: a table in r)
//     q' = q + (u~ . l~) l~                        redshift-space distortion
//     q'' = alpha(r') q'  |  alpha_iso q'  |  a_par (q' . l~) l~ + a_perp (q' - (q' . l~) l~)      Alcock-Paczynski
// and phys2cell of R q'' is pos + (q'' - q) / cell: only the DIFFERENCE q'' - q is formed (in float32 the round trip
// through absolute Mpc/h coordinates would cost 1e-4 cell a Gpc from the observer), and the rotation is needed for the
// velocity-bias term alone.  The functions of the comoving distance that carry the cosmology -- D f along the light
// cone, and alpha - 1 = chi_fid(a(chi)) / chi - 1 of ap_auto -- come as tables on a uniform radius grid, built by the
// caller from its (differentiable) cosmology; their cotangents are returned node by node.
#pragma once
#include "rt.h"

namespace mcpm {

#ifdef MCPM_HOSTEMU
#define MCPM_DEV inline
#else
#define MCPM_DEV __device__ __forceinline__
#endif

constexpr int kObsSlots = 32;  // replicas of the parameter-cotangent row (spreads the atomics; summed by the entry point)

struct ObsGen {
  int on = 0;
  int curved = 0;     // 1: lines of sight from the observer (bricks.py:754-756); 0: the fixed direction l
  int lightcone = 0;  // 1: D f from tab_gf at the particle's distance (a_obs None, bricks.py:761-762); 0: the scalar gf
  int ap = 0;         // 0 none | 1 ap_auto: alpha - 1 from tab_ap | 2 ap_param: a_par (= alpha_iso on a curved sky), a_perp
  int rsd = 1;        // 0: no redshift-space term
  float cx = 1.f, cy = 1.f, cz = 1.f;  // Mpc/h per unit of pos
  float ox = 0.f, oy = 0.f, oz = 0.f;  // q = pos * c + o
  float lx = 0.f, ly = 0.f, lz = 0.f;  // flat-sky line of sight in the box frame
  float gf = 0.f, a_par = 1.f, a_perp = 1.f;
  const float* par = nullptr;  // the same three numbers in device memory (a caller whose scalars are traced values)
  float r0 = 0.f, inv_dr = 0.f;  // table nodes r0 + k / inv_dr, k = 0 .. nt - 1
  int nt = 0;
  const float* tab_gf = nullptr;
  const float* tab_ap = nullptr;
  const float* vel = nullptr;   // [np,3] growth-time velocities in units of pos
  const float* dvel = nullptr;  // [np,3] velocity bias in the observer's frame, Mpc/h (nullable)
  float rt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // R^T, row-major
  // transpose only (each nullable)
  float* dvelbar = nullptr;  // [np,3]
  double* parbar = nullptr;  // [kObsSlots][3 + 2 nt]: gf, a_par, a_perp, tab_gf nodes, tab_ap nodes
};

struct ObsTabCell {
  int i;
  float w, slope;  // value = tab[i] + w (tab[i+1] - tab[i]); slope = d value / d r (0 outside the table: clamped)
};
MCPM_HD float obs_interp(const float* tab, const ObsGen& o, float r, ObsTabCell& c) {
  float t = (r - o.r0) * o.inv_dr;
  const float tmax = (float)(o.nt - 1);
  const bool inside = t >= 0.0f && t <= tmax;
  t = t < 0.0f ? 0.0f : (t > tmax ? tmax : t);
  c.i = (int)t;
  if (c.i > o.nt - 2) c.i = o.nt - 2;
  c.w = t - (float)c.i;
  const float a = tab[c.i], b = tab[c.i + 1];
  c.slope = inside ? (b - a) * o.inv_dr : 0.0f;
  return a + c.w * (b - a);
}

// Distance of a particle at x (units of pos) from the observer (curved sky) or along the line of sight (flat).
// dir (nullable): d r / d q up to the sign `sgn` (1 on a curved sky).
MCPM_HD float obs_radius(const ObsGen& o, const float* x, float* dir, float* sgn) {
  const float q0 = x[0] * o.cx + o.ox, q1 = x[1] * o.cy + o.oy, q2 = x[2] * o.cz + o.oz;
  if (o.curved) {
    const float r = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);
    if (dir) {
      const float ir = r > 0.0f ? 1.0f / r : 0.0f;
      dir[0] = q0 * ir, dir[1] = q1 * ir, dir[2] = q2 * ir;
      *sgn = 1.0f;
    }
    return r;
  }
  const float t = q0 * o.lx + q1 * o.ly + q2 * o.lz;
  if (dir) {
    dir[0] = o.lx, dir[1] = o.ly, dir[2] = o.lz;
    *sgn = t > 0.0f ? 1.0f : (t < 0.0f ? -1.0f : 0.0f);
  }
  return fabsf(t);
}

struct ObsState {
  float q[3], l[3], u[3], q1[3];
  float r, sgn, s, gf, r1, sgn1, t1, am1, a_par, a_perp;
  ObsTabCell cg, ca;
};

// delta (in units of pos) such that the particle is observed at pos + delta.  x: absolute position in units of pos.
MCPM_HD void obs_forward(const ObsGen& o, const float* x, int64_t p, ObsState& s, float* delta) {
  const float c[3] = {o.cx, o.cy, o.cz};
  s.q[0] = x[0] * o.cx + o.ox;
  s.q[1] = x[1] * o.cy + o.oy;
  s.q[2] = x[2] * o.cz + o.oz;
  s.sgn = 0.0f;
  s.a_par = o.par ? o.par[1] : o.a_par;
  s.a_perp = o.par ? o.par[2] : o.a_perp;
  if (o.curved) {
    s.r = sqrtf(s.q[0] * s.q[0] + s.q[1] * s.q[1] + s.q[2] * s.q[2]);
    const float ir = s.r > 0.0f ? 1.0f / s.r : 0.0f;  // safe_div
    s.l[0] = s.q[0] * ir;
    s.l[1] = s.q[1] * ir;
    s.l[2] = s.q[2] * ir;
  } else {
    s.l[0] = o.lx;
    s.l[1] = o.ly;
    s.l[2] = o.lz;
    const float t = s.q[0] * o.lx + s.q[1] * o.ly + s.q[2] * o.lz;
    s.r = fabsf(t);
    s.sgn = t > 0.0f ? 1.0f : (t < 0.0f ? -1.0f : 0.0f);
  }
  float d[3] = {0.0f, 0.0f, 0.0f};  // q'' - q
  s.s = 0.0f;
  s.gf = 0.0f;
  s.u[0] = s.u[1] = s.u[2] = 0.0f;
  if (o.rsd) {
    s.gf = o.lightcone ? obs_interp(o.tab_gf, o, s.r, s.cg) : (o.par ? o.par[0] : o.gf);
    const float* v = o.vel + 3 * p;
    for (int a = 0; a < 3; ++a) s.u[a] = v[a] * c[a] * s.gf;
    if (o.dvel) {
      const float* dv = o.dvel + 3 * p;
      for (int a = 0; a < 3; ++a) s.u[a] += o.rt[3 * a] * dv[0] + o.rt[3 * a + 1] * dv[1] + o.rt[3 * a + 2] * dv[2];
    }
    s.s = s.u[0] * s.l[0] + s.u[1] * s.l[1] + s.u[2] * s.l[2];
    for (int a = 0; a < 3; ++a) d[a] = s.s * s.l[a];
  }
  for (int a = 0; a < 3; ++a) s.q1[a] = s.q[a] + d[a];
  s.t1 = s.q1[0] * s.l[0] + s.q1[1] * s.l[1] + s.q1[2] * s.l[2];
  s.am1 = 0.0f;
  s.sgn1 = 0.0f;
  s.r1 = 0.0f;
  if (o.ap == 1) {
    if (o.curved) {
      s.r1 = sqrtf(s.q1[0] * s.q1[0] + s.q1[1] * s.q1[1] + s.q1[2] * s.q1[2]);
    } else {
      s.r1 = fabsf(s.t1);
      s.sgn1 = s.t1 > 0.0f ? 1.0f : (s.t1 < 0.0f ? -1.0f : 0.0f);
    }
    s.am1 = obs_interp(o.tab_ap, o, s.r1, s.ca);
    for (int a = 0; a < 3; ++a) d[a] += s.am1 * s.q1[a];
  } else if (o.ap == 2) {
    if (o.curved) {
      for (int a = 0; a < 3; ++a) d[a] += (s.a_par - 1.0f) * s.q1[a];
    } else {
      for (int a = 0; a < 3; ++a) d[a] += (s.a_par - 1.0f) * s.t1 * s.l[a] + (s.a_perp - 1.0f) * (s.q1[a] - s.t1 * s.l[a]);
    }
  }
  for (int a = 0; a < 3; ++a) delta[a] = d[a] / c[a];
}

// Transpose at the state `s` of obs_forward: g = cotangent of delta.  xadd = J_pos^T g (to be ADDED to the identity
// part g of the position cotangent), vbar = J_vel^T g; dvelbar and the parameter cotangents are written through `o`.
MCPM_DEV void obs_transpose(const ObsGen& o, const ObsState& s, int64_t p, const float* g, float* xadd, float* vbar,
                            bool accumulate) {
  const float c[3] = {o.cx, o.cy, o.cz};
  double* par = o.parbar ? o.parbar + (size_t)(p & (kObsSlots - 1)) * (size_t)(3 + 2 * o.nt) : nullptr;
  float h[3], q1b[3], qb[3] = {0.0f, 0.0f, 0.0f}, lb[3] = {0.0f, 0.0f, 0.0f};
  for (int a = 0; a < 3; ++a) h[a] = g[a] / c[a];
  const float hl = h[0] * s.l[0] + h[1] * s.l[1] + h[2] * s.l[2];
  const float hq1 = h[0] * s.q1[0] + h[1] * s.q1[1] + h[2] * s.q1[2];
  for (int a = 0; a < 3; ++a) q1b[a] = 0.0f;
  if (o.ap == 1) {
    const float r1b = hq1 * s.ca.slope;
    for (int a = 0; a < 3; ++a) q1b[a] = s.am1 * h[a];
    if (o.curved) {
      if (s.r1 > 0.0f)
        for (int a = 0; a < 3; ++a) q1b[a] += r1b * s.q1[a] / s.r1;
    } else {
      for (int a = 0; a < 3; ++a) q1b[a] += r1b * s.sgn1 * s.l[a];
    }
    if (par) {
      atomic_add(par + 3 + o.nt + s.ca.i, (double)((1.0f - s.ca.w) * hq1));
      atomic_add(par + 3 + o.nt + s.ca.i + 1, (double)(s.ca.w * hq1));
    }
  } else if (o.ap == 2) {
    if (o.curved) {
      for (int a = 0; a < 3; ++a) q1b[a] = (s.a_par - 1.0f) * h[a];
      if (par) atomic_add(par + 1, (double)hq1);
    } else {
      for (int a = 0; a < 3; ++a) q1b[a] = (s.a_perp - 1.0f) * h[a] + (s.a_par - s.a_perp) * hl * s.l[a];
      if (par) {
        atomic_add(par + 1, (double)(s.t1 * hl));
        atomic_add(par + 2, (double)(hq1 - s.t1 * hl));
      }
    }
  }
  // q1 = q + s l  and the direct term s l of the displacement
  for (int a = 0; a < 3; ++a) qb[a] = q1b[a];
  vbar[0] = vbar[1] = vbar[2] = 0.0f;
  float rb = 0.0f;
  if (o.rsd) {
    const float q1bl = q1b[0] * s.l[0] + q1b[1] * s.l[1] + q1b[2] * s.l[2];
    const float sb = hl + q1bl;
    float ub[3];
    for (int a = 0; a < 3; ++a) {
      lb[a] = s.s * (h[a] + q1b[a]) + sb * s.u[a];
      ub[a] = sb * s.l[a];
    }
    const float* v = o.vel + 3 * p;
    float gfb = 0.0f;
    for (int a = 0; a < 3; ++a) {
      vbar[a] = c[a] * ub[a] * s.gf;
      gfb += ub[a] * v[a] * c[a];
    }
    if (o.dvelbar) {
      float* db = o.dvelbar + 3 * p;
      for (int a = 0; a < 3; ++a) {  // R ub = (R^T)^T ub
        const float val = o.rt[a] * ub[0] + o.rt[3 + a] * ub[1] + o.rt[6 + a] * ub[2];
        db[a] = (accumulate ? db[a] : 0.0f) + val;
      }
    }
    if (o.lightcone) {
      rb = gfb * s.cg.slope;
      if (par) {
        atomic_add(par + 3 + s.cg.i, (double)((1.0f - s.cg.w) * gfb));
        atomic_add(par + 3 + s.cg.i + 1, (double)(s.cg.w * gfb));
      }
    } else if (par) {
      atomic_add(par, (double)gfb);
    }
  } else if (o.dvelbar) {
    float* db = o.dvelbar + 3 * p;
    for (int a = 0; a < 3; ++a) db[a] = accumulate ? db[a] : 0.0f;
  }
  if (o.curved) {
    if (s.r > 0.0f) {
      const float lbl = lb[0] * s.l[0] + lb[1] * s.l[1] + lb[2] * s.l[2];
      for (int a = 0; a < 3; ++a) qb[a] += (lb[a] - lbl * s.l[a]) / s.r + rb * s.l[a];
    }
  } else {
    for (int a = 0; a < 3; ++a) qb[a] += rb * s.sgn * s.l[a];
  }
  for (int a = 0; a < 3; ++a) xadd[a] = c[a] * qb[a];
}

}  // namespace mcpm
