// frame.h -- how a position array maps to mesh coordinates.
//
// The reference keeps absolute float64 positions `pos = q + displacement` (nbody.py:984-985: lpt's dpos is added to the
// lattice sites once and the sum is carried through the BullFrog loop).  In float32 an absolute coordinate near 256
// resolves 1.5e-5 of a cell (6e-5 at 1024): enough to put particles on the wrong side of a cell face, where the CIC
// derivative jumps -- the cause of the gradient error measured in round 1.  The engine therefore also accepts
// LATTICE-RELATIVE positions: the array holds d = x - site(p), the displacement of particle p from its own lattice
// site, and every kernel forms   cell = site + floor(d),  fraction = d - floor(d)   with `site` in exact integer
// arithmetic.  |d| stays of order 10 cells, so the fraction resolves ~1e-7..1e-6 of a cell at any mesh size.
//
//   rel == 0:  x_a = pos[p, a]                                                (the ABI's historical meaning)
//   rel == 1:  x_a = o_a + i_a * num_a / den_a + pos[p, a],  (i_x, i_y, i_z) = unravel(p, (px, py, pz)), C order
//              (regular_pos of bricks.py:593-603 has sites i_a * mesh_a / ptcl_a: num = mesh side, den = lattice side;
//               o_a is the halo offset of a slab-decomposed rank's local mesh)
// then, as before,  x'_a = x_a * scale_a + shift  (nufft's final -> paint units, nbody.py:569; interlace's shift, :524).
// With rel == 1 the site is given directly in the units of the TARGET mesh (num = target side), and only the
// displacement is multiplied by scale_a.
#pragma once
#include "rt.h"

namespace mcpm {

struct Frame {
  int rel = 0;
  int px = 0, py = 0, pz = 0;     // particle lattice (C order)
  int ox = 0, oy = 0, oz = 0;     // site origin, target-mesh cells
  int nux = 1, nuy = 1, nuz = 1;  // site spacing = nu / de target-mesh cells
  int dex = 1, dey = 1, dez = 1;
};

// Integer base cell and fractional remainder (in [0, 1)) of the lattice site of particle p.  Absolute frames return 0.
MCPM_HD void frame_site(const Frame& f, int64_t p, int& bx, int& by, int& bz, float& rx, float& ry, float& rz) {
  bx = by = bz = 0;
  rx = ry = rz = 0.0f;
  if (!f.rel) return;
  unsigned ix, iy, iz;
  if (p < ((int64_t)1 << 32)) {  // 32-bit divisions: every realistic particle count per device
    const unsigned q = (unsigned)p, jk = q / (unsigned)f.pz;
    iz = q - jk * (unsigned)f.pz;
    ix = jk / (unsigned)f.py;
    iy = jk - ix * (unsigned)f.py;
  } else {
    const int64_t jk = p / f.pz;
    iz = (unsigned)(p - jk * f.pz);
    ix = (unsigned)(jk / f.py);
    iy = (unsigned)(jk - (int64_t)ix * f.py);
  }
  if (f.dex == 1 && f.dey == 1 && f.dez == 1) {  // spacing a whole number of cells (lattice == mesh, or coarser)
    bx = f.ox + (int)ix * f.nux;
    by = f.oy + (int)iy * f.nuy;
    bz = f.oz + (int)iz * f.nuz;
    return;
  }
  const int64_t tx = (int64_t)ix * f.nux, ty = (int64_t)iy * f.nuy, tz = (int64_t)iz * f.nuz;
  const int qx = (int)(tx / f.dex), qy = (int)(ty / f.dey), qz = (int)(tz / f.dez);
  bx = f.ox + qx;
  by = f.oy + qy;
  bz = f.oz + qz;
  rx = (float)(tx - (int64_t)qx * f.dex) / (float)f.dex;
  ry = (float)(ty - (int64_t)qy * f.dey) / (float)f.dey;
  rz = (float)(tz - (int64_t)qz * f.dez) / (float)f.dez;
}

// Reduce num / den (site spacing mesh_side / lattice_side) to lowest terms.
inline void frame_ratio(int num, int den, int& nu, int& de) {
  int a = num, b = den;
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  nu = num / (a ? a : 1);
  de = den / (a ? a : 1);
}

// Frame of a px x py x pz lattice that spans an nx x ny x nz target mesh (offset by o cells).
inline Frame make_rel_frame(int px, int py, int pz, int nx, int ny, int nz, int ox = 0, int oy = 0, int oz = 0) {
  Frame f;
  f.rel = 1;
  f.px = px;
  f.py = py;
  f.pz = pz;
  f.ox = ox;
  f.oy = oy;
  f.oz = oz;
  frame_ratio(nx, px, f.nux, f.dex);
  frame_ratio(ny, py, f.nuy, f.dey);
  frame_ratio(nz, pz, f.nuz, f.dez);
  return f;
}

}  // namespace mcpm
