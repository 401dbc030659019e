// TEST INFRASTRUCTURE ONLY -- CPU oracle (see fv3_oracle.hpp header).
// Restates model/sw_core.F90 of the reference: c_sw (:79-488), d_sw (:494-1606),
// del6_vt_flux (:1608-1737), divergence_corner (:1740-1845), xtp_u (:2154-2521),
// ytp_v (:2524-2998), d2a2c_vect (:3006-3345), edge_interpolate4 (:3348-3359),
// fill2_4corners (:3434-3494), fill_4corners (:3496-3555); plus
// tools/fv_mp_mod.F90 fill_corners_2d BGRID (:1031-1062) and
// fill_corners_dgrid (:1249-1281).  Non-SW_DYNAMICS, non-USE_SG, non-ROT3,
// non-ONE_SIDE, non-GFS_PHYS build (the CI "solo nh 64bit repro" variant).
#include "fv3_oracle.hpp"

namespace fv3o {

// sw_core.F90:36-70
static const double r3 = 1. / 3.;
static const double s11 = 11. / 14., s14 = 4. / 7., s15 = 3. / 14.;
static const double near_zero = 1.E-9;
static const double big_number = 1.E30;
static const double p1 = 7. / 12.;
static const double p2 = -1. / 12.;
static const double a1 = 0.5625;
static const double a2 = -0.0625;
static const double c1 = -2. / 14.;
static const double c2 = 11. / 14.;
static const double c3 = 5. / 14.;

// sw_core.F90:3496-3555
void fill_4corners(V2 q, int dir, const Bd& bd) {
  const int npx = bd.npx, npy = bd.npy;
  if (dir == 1) {
    if (bd.sw_corner) { q(-1, 0) = q(0, 2); q(0, 0) = q(0, 1); }
    if (bd.se_corner) { q(npx + 1, 0) = q(npx, 2); q(npx, 0) = q(npx, 1); }
    if (bd.nw_corner) { q(0, npy) = q(0, npy - 1); q(-1, npy) = q(0, npy - 2); }
    if (bd.ne_corner) { q(npx, npy) = q(npx, npy - 1); q(npx + 1, npy) = q(npx, npy - 2); }
  } else if (dir == 2) {
    if (bd.sw_corner) { q(0, 0) = q(1, 0); q(0, -1) = q(2, 0); }
    if (bd.se_corner) { q(npx, 0) = q(npx - 1, 0); q(npx, -1) = q(npx - 2, 0); }
    if (bd.nw_corner) { q(0, npy) = q(1, npy); q(0, npy + 1) = q(2, npy); }
    if (bd.ne_corner) { q(npx, npy) = q(npx - 1, npy); q(npx, npy + 1) = q(npx - 2, npy); }
  }
}
// sw_core.F90:3434-3494
void fill2_4corners(V2 q1, V2 q2, int dir, const Bd& bd) {
  fill_4corners(q1, dir, bd);
  fill_4corners(q2, dir, bd);
}

// fv_mp_mod.F90:1031-1062 (BGRID branch; default == XDir)
void fill_corners_bgrid(V2 q, int npx, int npy, int ng, int fill_dir) {
  if (fill_dir == 2) {
    for (int j = 1; j <= ng; j++)
      for (int i = 1; i <= ng; i++) {
        q(1 - j, 1 - i) = q(i + 1, 1 - j);
        q(1 - j, npy + i) = q(i + 1, npy + j);
        q(npx + j, 1 - i) = q(npx - i, 1 - j);
        q(npx + j, npy + i) = q(npx - i, npy + j);
      }
  } else {
    for (int j = 1; j <= ng; j++)
      for (int i = 1; i <= ng; i++) {
        q(1 - i, 1 - j) = q(1 - j, i + 1);
        q(1 - i, npy + j) = q(1 - j, npy - i);
        q(npx + i, 1 - j) = q(npx + j, i + 1);
        q(npx + i, npy + j) = q(npx + j, npy - i);
      }
  }
}
// fv_mp_mod.F90:1249-1281
void fill_corners_dgrid_vec(V2 x, V2 y, int npx, int npy, int ng, double mySign) {
  for (int j = 1; j <= ng; j++)
    for (int i = 1; i <= ng; i++) {
      x(1 - i, 1 - j) = mySign * y(1 - j, i);
      x(1 - i, npy + j) = y(1 - j, npy - i);
      x(npx - 1 + i, 1 - j) = y(npx + j, i);
      x(npx - 1 + i, npy + j) = mySign * y(npx + j, npy - i);
    }
  for (int j = 1; j <= ng; j++)
    for (int i = 1; i <= ng; i++) {
      y(1 - i, 1 - j) = mySign * x(j, 1 - i);
      y(1 - i, npy - 1 + j) = x(j, npy + i);
      y(npx + i, 1 - j) = x(npx - j, 1 - i);
      y(npx + i, npy - 1 + j) = mySign * x(npx - j, npy + i);
    }
}

// sw_core.F90:3348-3359 (ua, dxa are 4-element windows)
static inline double edge_interpolate4(double ua1, double ua2, double ua3, double ua4, double d1, double d2,
                                       double d3, double d4) {
  double t1 = d1 + d2;
  double t2 = d3 + d4;
  return 0.5 * (((t1 + d2) * ua2 - d2 * ua1) / t1 + ((t2 + d3) * ua3 - d3 * ua4) / t2);
}

// sw_core.F90:3006-3345
void d2a2c_vect(V2 u, V2 v, V2 ua, V2 va, V2 uc, V2 vc, V2 ut, V2 vt, bool dord4, const Grid& g, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const int npx = bd.npx, npy = bd.npy, grid_type = bd.grid_type;
  const bool bounded = bd.bounded_domain;
  L2 utmp(isd, ied, jsd, jed, big_number), vtmp(isd, ied, jsd, jed, big_number);
  const int id = dord4 ? 1 : 0;
  const int npt = (grid_type < 3 && !bounded) ? 4 : -2;

  // Interior
  for (int j = std::max(npt, js - 1); j <= std::min(npy - npt, je + 1); j++)
    for (int i = std::max(npt, isd); i <= std::min(npx - npt, ied); i++)
      utmp(i, j) = a2 * (u(i, j - 1) + u(i, j + 2)) + a1 * (u(i, j) + u(i, j + 1));
  for (int j = std::max(npt, jsd); j <= std::min(npy - npt, jed); j++)
    for (int i = std::max(npt, is - 1); i <= std::min(npx - npt, ie + 1); i++)
      vtmp(i, j) = a2 * (v(i - 1, j) + v(i + 2, j)) + a1 * (v(i, j) + v(i + 1, j));
  // edges
  if (grid_type < 3) {
    if (js == 1 || jsd < npt)
      for (int j = jsd; j <= npt - 1; j++)
        for (int i = isd; i <= ied; i++) {
          utmp(i, j) = 0.5 * (u(i, j) + u(i, j + 1));
          vtmp(i, j) = 0.5 * (v(i, j) + v(i + 1, j));
        }
    if ((je + 1) == npy || jed >= (npy - npt))
      for (int j = npy - npt + 1; j <= jed; j++)
        for (int i = isd; i <= ied; i++) {
          utmp(i, j) = 0.5 * (u(i, j) + u(i, j + 1));
          vtmp(i, j) = 0.5 * (v(i, j) + v(i + 1, j));
        }
    if (is == 1 || isd < npt)
      for (int j = std::max(npt, jsd); j <= std::min(npy - npt, jed); j++)
        for (int i = isd; i <= npt - 1; i++) {
          utmp(i, j) = 0.5 * (u(i, j) + u(i, j + 1));
          vtmp(i, j) = 0.5 * (v(i, j) + v(i + 1, j));
        }
    if ((ie + 1) == npx || ied >= (npx - npt))
      for (int j = std::max(npt, jsd); j <= std::min(npy - npt, jed); j++)
        for (int i = npx - npt + 1; i <= ied; i++) {
          utmp(i, j) = 0.5 * (u(i, j) + u(i, j + 1));
          vtmp(i, j) = 0.5 * (v(i, j) + v(i + 1, j));
        }
  }
  // Contra-variant components at cell center
  for (int j = js - 1 - id; j <= je + 1 + id; j++)
    for (int i = is - 1 - id; i <= ie + 1 + id; i++) {
      ua(i, j) = (utmp(i, j) - vtmp(i, j) * g.cosa_s(i, j)) * g.rsin2(i, j);
      va(i, j) = (vtmp(i, j) - utmp(i, j) * g.cosa_s(i, j)) * g.rsin2(i, j);
    }
  // A -> C, Xdir: fix the edges
  if (bd.sw_corner) for (int i = -2; i <= 0; i++) utmp(i, 0) = -vtmp(0, 1 - i);
  if (bd.se_corner) for (int i = 0; i <= 2; i++) utmp(npx + i, 0) = vtmp(npx, i + 1);
  if (bd.ne_corner) for (int i = 0; i <= 2; i++) utmp(npx + i, npy) = -vtmp(npx, je - i);
  if (bd.nw_corner) for (int i = -2; i <= 0; i++) utmp(i, npy) = vtmp(0, je + i);

  int ifirst, ilast;
  if (grid_type < 3 && !bounded) { ifirst = std::max(3, is - 1); ilast = std::min(npx - 2, ie + 2); }
  else { ifirst = is - 1; ilast = ie + 2; }
  for (int j = js - 1; j <= je + 1; j++)
    for (int i = ifirst; i <= ilast; i++) {
      uc(i, j) = a2 * (utmp(i - 2, j) + utmp(i + 1, j)) + a1 * (utmp(i - 1, j) + utmp(i, j));
      ut(i, j) = (uc(i, j) - v(i, j) * g.cosa_u(i, j)) * g.rsin_u(i, j);
    }
  if (grid_type < 3) {
    if (bd.sw_corner) { ua(-1, 0) = -va(0, 2); ua(0, 0) = -va(0, 1); }
    if (bd.se_corner) { ua(npx, 0) = va(npx, 1); ua(npx + 1, 0) = va(npx, 2); }
    if (bd.ne_corner) { ua(npx, npy) = -va(npx, npy - 1); ua(npx + 1, npy) = -va(npx, npy - 2); }
    if (bd.nw_corner) { ua(-1, npy) = va(0, npy - 2); ua(0, npy) = va(0, npy - 1); }

    if (is == 1 && !bounded) {
      for (int j = js - 1; j <= je + 1; j++) {
        uc(0, j) = c1 * utmp(-2, j) + c2 * utmp(-1, j) + c3 * utmp(0, j);
        ut(1, j) = edge_interpolate4(ua(-1, j), ua(0, j), ua(1, j), ua(2, j), g.dxa(-1, j), g.dxa(0, j),
                                     g.dxa(1, j), g.dxa(2, j));
        if (ut(1, j) > 0.) uc(1, j) = ut(1, j) * g.sin_sg(0, j, 3);
        else uc(1, j) = ut(1, j) * g.sin_sg(1, j, 1);
        uc(2, j) = c1 * utmp(3, j) + c2 * utmp(2, j) + c3 * utmp(1, j);
        ut(0, j) = (uc(0, j) - v(0, j) * g.cosa_u(0, j)) * g.rsin_u(0, j);
        ut(2, j) = (uc(2, j) - v(2, j) * g.cosa_u(2, j)) * g.rsin_u(2, j);
      }
    }
    if ((ie + 1) == npx && !bounded) {
      for (int j = js - 1; j <= je + 1; j++) {
        uc(npx - 1, j) = c1 * utmp(npx - 3, j) + c2 * utmp(npx - 2, j) + c3 * utmp(npx - 1, j);
        ut(npx, j) = edge_interpolate4(ua(npx - 2, j), ua(npx - 1, j), ua(npx, j), ua(npx + 1, j),
                                       g.dxa(npx - 2, j), g.dxa(npx - 1, j), g.dxa(npx, j), g.dxa(npx + 1, j));
        if (ut(npx, j) > 0.) uc(npx, j) = ut(npx, j) * g.sin_sg(npx - 1, j, 3);
        else uc(npx, j) = ut(npx, j) * g.sin_sg(npx, j, 1);
        uc(npx + 1, j) = c3 * utmp(npx, j) + c2 * utmp(npx + 1, j) + c1 * utmp(npx + 2, j);
        ut(npx - 1, j) = (uc(npx - 1, j) - v(npx - 1, j) * g.cosa_u(npx - 1, j)) * g.rsin_u(npx - 1, j);
        ut(npx + 1, j) = (uc(npx + 1, j) - v(npx + 1, j) * g.cosa_u(npx + 1, j)) * g.rsin_u(npx + 1, j);
      }
    }
  }
  // Ydir
  if (bd.sw_corner) for (int j = -2; j <= 0; j++) vtmp(0, j) = -utmp(1 - j, 0);
  if (bd.nw_corner) for (int j = 0; j <= 2; j++) vtmp(0, npy + j) = utmp(j + 1, npy);
  if (bd.se_corner) for (int j = -2; j <= 0; j++) vtmp(npx, j) = utmp(ie + j, 0);
  if (bd.ne_corner) for (int j = 0; j <= 2; j++) vtmp(npx, npy + j) = -utmp(ie - j, npy);
  if (bd.sw_corner) { va(0, -1) = -ua(2, 0); va(0, 0) = -ua(1, 0); }
  if (bd.se_corner) { va(npx, 0) = ua(npx - 1, 0); va(npx, -1) = ua(npx - 2, 0); }
  if (bd.ne_corner) { va(npx, npy) = -ua(npx - 1, npy); va(npx, npy + 1) = -ua(npx - 2, npy); }
  if (bd.nw_corner) { va(0, npy) = ua(1, npy); va(0, npy + 1) = ua(2, npy); }

  if (grid_type < 3) {
    for (int j = js - 1; j <= je + 2; j++) {
      // Fortran precedence: .and. binds tighter than .or. (sw_core.F90:3308,3313)
      if (j == 1 && !bounded) {
        for (int i = is - 1; i <= ie + 1; i++) {
          vt(i, j) = edge_interpolate4(va(i, -1), va(i, 0), va(i, 1), va(i, 2), g.dya(i, -1), g.dya(i, 0),
                                       g.dya(i, 1), g.dya(i, 2));
          if (vt(i, j) > 0.) vc(i, j) = vt(i, j) * g.sin_sg(i, j - 1, 4);
          else vc(i, j) = vt(i, j) * g.sin_sg(i, j, 2);
        }
      } else if (j == 0 || (j == (npy - 1) && !bounded)) {
        for (int i = is - 1; i <= ie + 1; i++) {
          vc(i, j) = c1 * vtmp(i, j - 2) + c2 * vtmp(i, j - 1) + c3 * vtmp(i, j);
          vt(i, j) = (vc(i, j) - u(i, j) * g.cosa_v(i, j)) * g.rsin_v(i, j);
        }
      } else if (j == 2 || (j == (npy + 1) && !bounded)) {
        for (int i = is - 1; i <= ie + 1; i++) {
          vc(i, j) = c1 * vtmp(i, j + 1) + c2 * vtmp(i, j) + c3 * vtmp(i, j - 1);
          vt(i, j) = (vc(i, j) - u(i, j) * g.cosa_v(i, j)) * g.rsin_v(i, j);
        }
      } else if (j == npy && !bounded) {
        for (int i = is - 1; i <= ie + 1; i++) {
          vt(i, j) = edge_interpolate4(va(i, j - 2), va(i, j - 1), va(i, j), va(i, j + 1), g.dya(i, j - 2),
                                       g.dya(i, j - 1), g.dya(i, j), g.dya(i, j + 1));
          if (vt(i, j) > 0.) vc(i, j) = vt(i, j) * g.sin_sg(i, j - 1, 4);
          else vc(i, j) = vt(i, j) * g.sin_sg(i, j, 2);
        }
      } else {
        for (int i = is - 1; i <= ie + 1; i++) {
          vc(i, j) = a2 * (vtmp(i, j - 2) + vtmp(i, j + 1)) + a1 * (vtmp(i, j - 1) + vtmp(i, j));
          vt(i, j) = (vc(i, j) - u(i, j) * g.cosa_v(i, j)) * g.rsin_v(i, j);
        }
      }
    }
  } else {
    for (int j = js - 1; j <= je + 2; j++)
      for (int i = is - 1; i <= ie + 1; i++) {
        vc(i, j) = a2 * (vtmp(i, j - 2) + vtmp(i, j + 1)) + a1 * (vtmp(i, j - 1) + vtmp(i, j));
        vt(i, j) = vc(i, j);
      }
  }
}

// sw_core.F90:1740-1845
void divergence_corner(V2 u, V2 v, V2 ua, V2 va, V2 divg_d, const Grid& g, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, npx = bd.npx, npy = bd.npy;
  L2 uf(is - 2, ie + 2, js - 1, je + 2), vf(is - 1, ie + 2, js - 2, je + 2);
  int is2, ie1;
  if (bd.bounded_domain) { is2 = is; ie1 = ie + 1; }
  else { is2 = std::max(2, is); ie1 = std::min(npx - 1, ie + 1); }
  if (bd.grid_type > 3) {
    for (int j = js - 1; j <= je + 2; j++) for (int i = is - 2; i <= ie + 2; i++) uf(i, j) = u(i, j) * g.dyc(i, j);
    for (int j = js - 2; j <= je + 2; j++) for (int i = is - 1; i <= ie + 2; i++) vf(i, j) = v(i, j) * g.dxc(i, j);
    for (int j = js - 1; j <= je + 2; j++)
      for (int i = is - 1; i <= ie + 2; i++)
        divg_d(i, j) = g.rarea_c(i, j) * (vf(i, j - 1) - vf(i, j) + uf(i - 1, j) - uf(i, j));
  } else {
    for (int j = js; j <= je + 1; j++) {
      if (j == 1 || j == npy) {
        for (int i = is - 1; i <= ie + 1; i++)
          uf(i, j) = u(i, j) * g.dyc(i, j) * 0.5 * (g.sin_sg(i, j - 1, 4) + g.sin_sg(i, j, 2));
      } else {
        for (int i = is - 1; i <= ie + 1; i++)
          uf(i, j) = (u(i, j) - 0.25 * (va(i, j - 1) + va(i, j)) * (g.cos_sg(i, j - 1, 4) + g.cos_sg(i, j, 2))) *
                     g.dyc(i, j) * 0.5 * (g.sin_sg(i, j - 1, 4) + g.sin_sg(i, j, 2));
      }
    }
    for (int j = js - 1; j <= je + 1; j++) {
      for (int i = is2; i <= ie1; i++)
        vf(i, j) = (v(i, j) - 0.25 * (ua(i - 1, j) + ua(i, j)) * (g.cos_sg(i - 1, j, 3) + g.cos_sg(i, j, 1))) *
                   g.dxc(i, j) * 0.5 * (g.sin_sg(i - 1, j, 3) + g.sin_sg(i, j, 1));
      if (is == 1) vf(1, j) = v(1, j) * g.dxc(1, j) * 0.5 * (g.sin_sg(0, j, 3) + g.sin_sg(1, j, 1));
      if ((ie + 1) == npx) vf(npx, j) = v(npx, j) * g.dxc(npx, j) * 0.5 * (g.sin_sg(npx - 1, j, 3) + g.sin_sg(npx, j, 1));
    }
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) divg_d(i, j) = vf(i, j - 1) - vf(i, j) + uf(i - 1, j) - uf(i, j);
    if (bd.sw_corner) divg_d(1, 1) = divg_d(1, 1) - vf(1, 0);
    if (bd.se_corner) divg_d(npx, 1) = divg_d(npx, 1) - vf(npx, 0);
    if (bd.ne_corner) divg_d(npx, npy) = divg_d(npx, npy) + vf(npx, npy);
    if (bd.nw_corner) divg_d(1, npy) = divg_d(1, npy) + vf(1, npy);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) divg_d(i, j) = g.rarea_c(i, j) * divg_d(i, j);
  }
}

// sw_core.F90:79-488
void c_sw(V2 delpc, V2 delp, V2 ptc, V2 pt, V2 u, V2 v, V2 w, V2 uc, V2 vc, V2 ua, V2 va, V2 wc, V2 ut,
          V2 vt, V2 divg_d, int nord, double dt2, bool hydrostatic, bool dord4, const Bd& bd, const Grid& g) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, npx = bd.npx, npy = bd.npy;
  const bool bounded = bd.bounded_domain;
  const int grid_type = bd.grid_type;
  L2 vort(is - 1, ie + 1, js - 1, je + 1), ke(is - 1, ie + 1, js - 1, je + 1);
  L2 fx(is - 1, ie + 2, js - 1, je + 1), fx1(is - 1, ie + 2, js - 1, je + 1), fx2(is - 1, ie + 2, js - 1, je + 1);
  L2 fy(is - 1, ie + 1, js - 1, je + 2), fy1(is - 1, ie + 1, js - 1, je + 2), fy2(is - 1, ie + 1, js - 1, je + 2);
  const int iep1 = ie + 1, jep1 = je + 1;

  d2a2c_vect(u, v, ua, va, uc, vc, ut, vt, dord4, g, bd);
  if (nord > 0) divergence_corner(u, v, ua, va, divg_d, g, bd);

  for (int j = js - 1; j <= jep1; j++)
    for (int i = is - 1; i <= iep1 + 1; i++) {
      if (ut(i, j) > 0.) ut(i, j) = dt2 * ut(i, j) * g.dy(i, j) * g.sin_sg(i - 1, j, 3);
      else ut(i, j) = dt2 * ut(i, j) * g.dy(i, j) * g.sin_sg(i, j, 1);
    }
  for (int j = js - 1; j <= je + 2; j++)
    for (int i = is - 1; i <= iep1; i++) {
      if (vt(i, j) > 0.) vt(i, j) = dt2 * vt(i, j) * g.dx(i, j) * g.sin_sg(i, j - 1, 4);
      else vt(i, j) = dt2 * vt(i, j) * g.dx(i, j) * g.sin_sg(i, j, 2);
    }

  // Transport delp: Xdir
  if (grid_type < 3 && !bounded) fill2_4corners(delp, pt, 1, bd);
  if (hydrostatic) {
    for (int j = js - 1; j <= jep1; j++)
      for (int i = is - 1; i <= ie + 2; i++) {
        if (ut(i, j) > 0.) { fx1(i, j) = delp(i - 1, j); fx(i, j) = pt(i - 1, j); }
        else { fx1(i, j) = delp(i, j); fx(i, j) = pt(i, j); }
        fx1(i, j) = ut(i, j) * fx1(i, j);
        fx(i, j) = fx1(i, j) * fx(i, j);
      }
  } else {
    if (grid_type < 3) fill_4corners(w, 1, bd);
    for (int j = js - 1; j <= je + 1; j++)
      for (int i = is - 1; i <= ie + 2; i++) {
        if (ut(i, j) > 0.) { fx1(i, j) = delp(i - 1, j); fx(i, j) = pt(i - 1, j); fx2(i, j) = w(i - 1, j); }
        else { fx1(i, j) = delp(i, j); fx(i, j) = pt(i, j); fx2(i, j) = w(i, j); }
        fx1(i, j) = ut(i, j) * fx1(i, j);
        fx(i, j) = fx1(i, j) * fx(i, j);
        fx2(i, j) = fx1(i, j) * fx2(i, j);
      }
  }
  // Ydir
  if (grid_type < 3 && !bounded) fill2_4corners(delp, pt, 2, bd);
  if (hydrostatic) {
    for (int j = js - 1; j <= jep1 + 1; j++)
      for (int i = is - 1; i <= iep1; i++) {
        if (vt(i, j) > 0.) { fy1(i, j) = delp(i, j - 1); fy(i, j) = pt(i, j - 1); }
        else { fy1(i, j) = delp(i, j); fy(i, j) = pt(i, j); }
        fy1(i, j) = vt(i, j) * fy1(i, j);
        fy(i, j) = fy1(i, j) * fy(i, j);
      }
    for (int j = js - 1; j <= jep1; j++)
      for (int i = is - 1; i <= iep1; i++) {
        delpc(i, j) = delp(i, j) + (fx1(i, j) - fx1(i + 1, j) + fy1(i, j) - fy1(i, j + 1)) * g.rarea(i, j);
        ptc(i, j) = (pt(i, j) * delp(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * g.rarea(i, j)) / delpc(i, j);
      }
  } else {
    if (grid_type < 3) fill_4corners(w, 2, bd);
    for (int j = js - 1; j <= je + 2; j++)
      for (int i = is - 1; i <= ie + 1; i++) {
        if (vt(i, j) > 0.) { fy1(i, j) = delp(i, j - 1); fy(i, j) = pt(i, j - 1); fy2(i, j) = w(i, j - 1); }
        else { fy1(i, j) = delp(i, j); fy(i, j) = pt(i, j); fy2(i, j) = w(i, j); }
        fy1(i, j) = vt(i, j) * fy1(i, j);
        fy(i, j) = fy1(i, j) * fy(i, j);
        fy2(i, j) = fy1(i, j) * fy2(i, j);
      }
    for (int j = js - 1; j <= je + 1; j++)
      for (int i = is - 1; i <= ie + 1; i++) {
        delpc(i, j) = delp(i, j) + (fx1(i, j) - fx1(i + 1, j) + fy1(i, j) - fy1(i, j + 1)) * g.rarea(i, j);
        ptc(i, j) = (pt(i, j) * delp(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * g.rarea(i, j)) / delpc(i, j);
        wc(i, j) = (w(i, j) * delp(i, j) + (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * g.rarea(i, j)) / delpc(i, j);
      }
  }

  // Compute KE
  if (bounded || grid_type >= 3) {
    for (int j = js - 1; j <= jep1; j++)
      for (int i = is - 1; i <= iep1; i++) ke(i, j) = (ua(i, j) > 0.) ? uc(i, j) : uc(i + 1, j);
    for (int j = js - 1; j <= jep1; j++)
      for (int i = is - 1; i <= iep1; i++) vort(i, j) = (va(i, j) > 0.) ? vc(i, j) : vc(i, j + 1);
  } else {
    for (int j = js - 1; j <= jep1; j++)
      for (int i = is - 1; i <= iep1; i++) {
        if (ua(i, j) > 0.) {
          if (i == 1) ke(1, j) = uc(1, j) * g.sin_sg(1, j, 1) + v(1, j) * g.cos_sg(1, j, 1);
          else if (i == npx) ke(i, j) = uc(npx, j) * g.sin_sg(npx, j, 1) + v(npx, j) * g.cos_sg(npx, j, 1);
          else ke(i, j) = uc(i, j);
        } else {
          if (i == 0) ke(0, j) = uc(1, j) * g.sin_sg(0, j, 3) + v(1, j) * g.cos_sg(0, j, 3);
          else if (i == (npx - 1)) ke(i, j) = uc(npx, j) * g.sin_sg(npx - 1, j, 3) + v(npx, j) * g.cos_sg(npx - 1, j, 3);
          else ke(i, j) = uc(i + 1, j);
        }
      }
    for (int j = js - 1; j <= jep1; j++)
      for (int i = is - 1; i <= iep1; i++) {
        if (va(i, j) > 0.) {
          if (j == 1) vort(i, 1) = vc(i, 1) * g.sin_sg(i, 1, 2) + u(i, 1) * g.cos_sg(i, 1, 2);
          else if (j == npy) vort(i, j) = vc(i, npy) * g.sin_sg(i, npy, 2) + u(i, npy) * g.cos_sg(i, npy, 2);
          else vort(i, j) = vc(i, j);
        } else {
          if (j == 0) vort(i, 0) = vc(i, 1) * g.sin_sg(i, 0, 4) + u(i, 1) * g.cos_sg(i, 0, 4);
          else if (j == (npy - 1)) vort(i, j) = vc(i, npy) * g.sin_sg(i, npy - 1, 4) + u(i, npy) * g.cos_sg(i, npy - 1, 4);
          else vort(i, j) = vc(i, j + 1);
        }
      }
  }
  const double dt4 = 0.5 * dt2;
  for (int j = js - 1; j <= jep1; j++)
    for (int i = is - 1; i <= iep1; i++) ke(i, j) = dt4 * (ua(i, j) * ke(i, j) + va(i, j) * vort(i, j));

  // circulation on C grid
  for (int j = js - 1; j <= je + 1; j++)
    for (int i = is; i <= ie + 1; i++) fx(i, j) = uc(i, j) * g.dxc(i, j);
  for (int j = js; j <= je + 1; j++)
    for (int i = is - 1; i <= ie + 1; i++) fy(i, j) = vc(i, j) * g.dyc(i, j);
  for (int j = js; j <= je + 1; j++)
    for (int i = is; i <= ie + 1; i++) vort(i, j) = fx(i, j - 1) - fx(i, j) - fy(i - 1, j) + fy(i, j);
  if (bd.sw_corner) vort(1, 1) = vort(1, 1) + fy(0, 1);
  if (bd.se_corner) vort(npx, 1) = vort(npx, 1) - fy(npx, 1);
  if (bd.ne_corner) vort(npx, npy) = vort(npx, npy) - fy(npx, npy);
  if (bd.nw_corner) vort(1, npy) = vort(1, npy) + fy(0, npy);
  for (int j = js; j <= je + 1; j++)
    for (int i = is; i <= ie + 1; i++) vort(i, j) = g.fC(i, j) + g.rarea_c(i, j) * vort(i, j);

  // Transport absolute vorticity
  if (bounded || grid_type >= 3) {
    for (int j = js; j <= je; j++)
      for (int i = is; i <= iep1; i++) {
        fy1(i, j) = dt2 * (v(i, j) - uc(i, j) * g.cosa_u(i, j)) / g.sina_u(i, j);
        fy(i, j) = (fy1(i, j) > 0.) ? vort(i, j) : vort(i, j + 1);
      }
    for (int j = js; j <= jep1; j++)
      for (int i = is; i <= ie; i++) {
        fx1(i, j) = dt2 * (u(i, j) - vc(i, j) * g.cosa_v(i, j)) / g.sina_v(i, j);
        fx(i, j) = (fx1(i, j) > 0.) ? vort(i, j) : vort(i + 1, j);
      }
  } else {
    for (int j = js; j <= je; j++)
      for (int i = is; i <= iep1; i++) {
        if (i == 1 || i == npx) fy1(i, j) = dt2 * v(i, j);
        else fy1(i, j) = dt2 * (v(i, j) - uc(i, j) * g.cosa_u(i, j)) / g.sina_u(i, j);
        fy(i, j) = (fy1(i, j) > 0.) ? vort(i, j) : vort(i, j + 1);
      }
    for (int j = js; j <= jep1; j++) {
      if (j == 1 || j == npy) {
        for (int i = is; i <= ie; i++) {
          fx1(i, j) = dt2 * u(i, j);
          fx(i, j) = (fx1(i, j) > 0.) ? vort(i, j) : vort(i + 1, j);
        }
      } else {
        for (int i = is; i <= ie; i++) {
          fx1(i, j) = dt2 * (u(i, j) - vc(i, j) * g.cosa_v(i, j)) / g.sina_v(i, j);
          fx(i, j) = (fx1(i, j) > 0.) ? vort(i, j) : vort(i + 1, j);
        }
      }
    }
  }
  // Update time-centered winds on the C-Grid
  for (int j = js; j <= je; j++)
    for (int i = is; i <= iep1; i++)
      uc(i, j) = uc(i, j) + fy1(i, j) * fy(i, j) + g.rdxc(i, j) * (ke(i - 1, j) - ke(i, j));
  for (int j = js; j <= jep1; j++)
    for (int i = is; i <= ie; i++)
      vc(i, j) = vc(i, j) - fx1(i, j) * fx(i, j) + g.rdyc(i, j) * (ke(i, j - 1) - ke(i, j));
}

// sw_core.F90:1608-1737 (non-USE_SG, no damp_Km)
void del6_vt_flux(int nord, int npx, int npy, double damp, V2 q, V2 d2, V2 fx2, V2 fy2, const Grid& g, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
  const bool bounded = bd.bounded_domain;
  const int i1 = is - 1 - nord, i2 = ie + 1 + nord, j1 = js - 1 - nord, j2 = je + 1 + nord;
  for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) d2(i, j) = damp * q(i, j);
  if (nord > 0 && !bounded) copy_corners(d2, npx, npy, 1, bd);
  for (int j = js - nord; j <= je + nord; j++)
    for (int i = is - nord; i <= ie + nord + 1; i++) fx2(i, j) = g.del6_v(i, j) * (d2(i - 1, j) - d2(i, j));
  if (nord > 0 && !bounded) copy_corners(d2, npx, npy, 2, bd);
  for (int j = js - nord; j <= je + nord + 1; j++)
    for (int i = is - nord; i <= ie + nord; i++) fy2(i, j) = g.del6_u(i, j) * (d2(i, j - 1) - d2(i, j));
  if (nord > 0) {
    for (int n = 1; n <= nord; n++) {
      const int nt = nord - n;
      for (int j = js - nt - 1; j <= je + nt + 1; j++)
        for (int i = is - nt - 1; i <= ie + nt + 1; i++)
          d2(i, j) = (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * g.rarea(i, j);
      if (!bounded) copy_corners(d2, npx, npy, 1, bd);
      for (int j = js - nt; j <= je + nt; j++)
        for (int i = is - nt; i <= ie + nt + 1; i++) fx2(i, j) = g.del6_v(i, j) * (d2(i, j) - d2(i - 1, j));
      if (!bounded) copy_corners(d2, npx, npy, 2, bd);
      for (int j = js - nt; j <= je + nt + 1; j++)
        for (int i = is - nt; i <= ie + nt; i++) fy2(i, j) = g.del6_u(i, j) * (d2(i, j) - d2(i, j - 1));
    }
  }
}

// sw_core.F90:2154-2521
void xtp_u(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, V2 c, V2 u, V2 v, V2 flux,
           int iord, V2 dx, V2 rdx, int npx, int npy, int grid_type, bool bounded_domain, double lim_fac) {
  (void)isd; (void)ied; (void)jsd; (void)jed; (void)v;
  L1 bl(is - 1, ie + 1), br(is - 1, ie + 1), b0(is - 1, ie + 1);
  LB1 smt5(is - 1, ie + 1), smt6(is - 1, ie + 1), hi5(is, ie + 1), hi6(is, ie + 1);
  L1 fx0(is, ie + 1), al(is - 1, ie + 2), dm(is - 2, ie + 2), dq(is - 3, ie + 2);
  int is3, ie3;
  if (bounded_domain || grid_type > 3) { is3 = is - 1; ie3 = ie + 1; }
  else { is3 = std::max(3, is - 1); ie3 = std::min(npx - 3, ie + 1); }
  const bool cube = (!bounded_domain) && grid_type < 3;

  if (iord < 8) {
    for (int j = js; j <= je + 1; j++) {
      for (int i = is3; i <= ie3 + 1; i++) al(i) = p1 * (u(i - 1, j) + u(i, j)) + p2 * (u(i - 2, j) + u(i + 1, j));
      for (int i = is3; i <= ie3; i++) { bl(i) = al(i) - u(i, j); br(i) = al(i + 1) - u(i, j); }
      if (cube) {
        if (is == 1) {
          double xt = c3 * u(1, j) + c2 * u(2, j) + c1 * u(3, j);
          br(1) = xt - u(1, j);
          bl(2) = xt - u(2, j);
          br(2) = al(3) - u(2, j);
          if (j == 1 || j == npy) {
            bl(0) = 0.; br(0) = 0.; bl(1) = 0.; br(1) = 0.;
          } else {
            bl(0) = c1 * u(-2, j) + c2 * u(-1, j) + c3 * u(0, j) - u(0, j);
            xt = 0.5 * (((2. * dx(0, j) + dx(-1, j)) * (u(0, j)) - dx(0, j) * u(-1, j)) / (dx(0, j) + dx(-1, j)) +
                        ((2. * dx(1, j) + dx(2, j)) * (u(1, j)) - dx(1, j) * u(2, j)) / (dx(1, j) + dx(2, j)));
            br(0) = xt - u(0, j);
            bl(1) = xt - u(1, j);
          }
        }
        if ((ie + 1) == npx) {
          bl(npx - 2) = al(npx - 2) - u(npx - 2, j);
          double xt = c1 * u(npx - 3, j) + c2 * u(npx - 2, j) + c3 * u(npx - 1, j);
          br(npx - 2) = xt - u(npx - 2, j);
          bl(npx - 1) = xt - u(npx - 1, j);
          if (j == 1 || j == npy) {
            bl(npx - 1) = 0.; br(npx - 1) = 0.; bl(npx) = 0.; br(npx) = 0.;
          } else {
            xt = 0.5 * (((2. * dx(npx - 1, j) + dx(npx - 2, j)) * u(npx - 1, j) - dx(npx - 1, j) * u(npx - 2, j)) /
                            (dx(npx - 1, j) + dx(npx - 2, j)) +
                        ((2. * dx(npx, j) + dx(npx + 1, j)) * u(npx, j) - dx(npx, j) * u(npx + 1, j)) /
                            (dx(npx, j) + dx(npx + 1, j)));
            br(npx - 1) = xt - u(npx - 1, j);
            bl(npx) = xt - u(npx, j);
            br(npx) = c3 * u(npx, j) + c2 * u(npx + 1, j) + c1 * u(npx + 2, j) - u(npx, j);
          }
        }
      }
      for (int i = is - 1; i <= ie + 1; i++) b0(i) = bl(i) + br(i);

      if (iord == 1) {
        for (int i = is - 1; i <= ie + 1; i++) smt5(i) = std::fabs(lim_fac * b0(i)) < std::fabs(bl(i) - br(i));
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdx(i - 1, j); fx0(i) = (1. - cfl) * (br(i - 1) - cfl * b0(i - 1)); flux(i, j) = u(i - 1, j); }
          else { cfl = c(i, j) * rdx(i, j); fx0(i) = (1. + cfl) * (bl(i) + cfl * b0(i)); flux(i, j) = u(i, j); }
          if (smt5(i - 1) || smt5(i)) flux(i, j) = flux(i, j) + fx0(i);
        }
      } else if (iord == 2) {
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdx(i - 1, j); flux(i, j) = u(i - 1, j) + (1. - cfl) * (br(i - 1) - cfl * b0(i - 1)); }
          else { cfl = c(i, j) * rdx(i, j); flux(i, j) = u(i, j) + (1. + cfl) * (bl(i) + cfl * b0(i)); }
        }
      } else if (iord == 3) {
        for (int i = is - 1; i <= ie + 1; i++) {
          double x0 = std::fabs(b0(i)), x1 = std::fabs(bl(i) - br(i));
          smt5(i) = x0 < x1; smt6(i) = 3. * x0 < x1;
        }
        for (int i = is; i <= ie + 1; i++) {
          fx0(i) = 0.;
          hi5(i) = smt5(i - 1) && smt5(i);
          hi6(i) = smt6(i - 1) || smt6(i);
        }
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) {
            cfl = c(i, j) * rdx(i - 1, j);
            if (hi6(i)) fx0(i) = br(i - 1) - cfl * b0(i - 1);
            else if (hi5(i)) fx0(i) = fsign(std::min(std::fabs(bl(i - 1)), std::fabs(br(i - 1))), br(i - 1));
            flux(i, j) = u(i - 1, j) + (1. - cfl) * fx0(i);
          } else {
            cfl = c(i, j) * rdx(i, j);
            if (hi6(i)) fx0(i) = bl(i) + cfl * b0(i);
            else if (hi5(i)) fx0(i) = fsign(std::min(std::fabs(bl(i)), std::fabs(br(i))), bl(i));
            flux(i, j) = u(i, j) + (1. + cfl) * fx0(i);
          }
        }
      } else if (iord == 4) {
        for (int i = is - 1; i <= ie + 1; i++) {
          double x0 = std::fabs(b0(i)), x1 = std::fabs(bl(i) - br(i));
          smt5(i) = x0 < x1; smt6(i) = 3. * x0 < x1;
        }
        for (int i = is; i <= ie + 1; i++) {
          hi5(i) = smt5(i - 1) && smt5(i);
          hi6(i) = smt6(i - 1) || smt6(i);
          hi5(i) = hi5(i) || hi6(i);
        }
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdx(i - 1, j); fx0(i) = (1. - cfl) * (br(i - 1) - cfl * b0(i - 1)); flux(i, j) = u(i - 1, j); }
          else { cfl = c(i, j) * rdx(i, j); fx0(i) = (1. + cfl) * (bl(i) + cfl * b0(i)); flux(i, j) = u(i, j); }
          if (hi5(i)) flux(i, j) = flux(i, j) + fx0(i);
        }
      } else {
        if (iord == 5) {
          for (int i = is - 1; i <= ie + 1; i++) smt5(i) = bl(i) * br(i) < 0.;
        } else {
          for (int i = is - 1; i <= ie + 1; i++) smt5(i) = 3. * std::fabs(b0(i)) < std::fabs(bl(i) - br(i));
          if (cube) {
            if (is == 1) { smt5(0) = bl(0) * br(0) < 0.; smt5(1) = bl(1) * br(1) < 0.; }
            if ((ie + 1) == npx) { smt5(npx - 1) = bl(npx - 1) * br(npx - 1) < 0.; smt5(npx) = bl(npx) * br(npx) < 0.; }
          }
        }
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdx(i - 1, j); fx0(i) = (1. - cfl) * (br(i - 1) - cfl * b0(i - 1)); flux(i, j) = u(i - 1, j); }
          else { cfl = c(i, j) * rdx(i, j); fx0(i) = (1. + cfl) * (bl(i) + cfl * b0(i)); flux(i, j) = u(i, j); }
          if (smt5(i - 1) || smt5(i)) flux(i, j) = flux(i, j) + fx0(i);
        }
      }
    }
  } else {
    for (int j = js; j <= je + 1; j++) {
      for (int i = is - 2; i <= ie + 2; i++) {
        double xt = 0.25 * (u(i + 1, j) - u(i - 1, j));
        dm(i) = fsign(std::min(std::min(std::fabs(xt), max3(u(i - 1, j), u(i, j), u(i + 1, j)) - u(i, j)),
                               u(i, j) - min3(u(i - 1, j), u(i, j), u(i + 1, j))), xt);
      }
      for (int i = is - 3; i <= ie + 2; i++) dq(i) = u(i + 1, j) - u(i, j);

      if (grid_type < 3) {
        for (int i = is3; i <= ie3 + 1; i++) al(i) = 0.5 * (u(i - 1, j) + u(i, j)) + r3 * (dm(i - 1) - dm(i));
        if (iord == 8) {
          for (int i = is3; i <= ie3; i++) {
            double xt = 2. * dm(i);
            bl(i) = -fsign(std::min(std::fabs(xt), std::fabs(al(i) - u(i, j))), xt);
            br(i) = fsign(std::min(std::fabs(xt), std::fabs(al(i + 1) - u(i, j))), xt);
          }
        } else if (iord == 9) {
          for (int i = is3; i <= ie3; i++) {
            double pmp_1 = -2. * dq(i), lac_1 = pmp_1 + 1.5 * dq(i + 1);
            bl(i) = std::min(max3(0., pmp_1, lac_1), std::max(al(i) - u(i, j), min3(0., pmp_1, lac_1)));
            double pmp_2 = 2. * dq(i - 1), lac_2 = pmp_2 - 1.5 * dq(i - 2);
            br(i) = std::min(max3(0., pmp_2, lac_2), std::max(al(i + 1) - u(i, j), min3(0., pmp_2, lac_2)));
          }
        } else if (iord == 10) {
          for (int i = is3; i <= ie3; i++) {
            bl(i) = al(i) - u(i, j);
            br(i) = al(i + 1) - u(i, j);
            if (std::fabs(dm(i)) < near_zero) {
              if (std::fabs(dm(i - 1)) + std::fabs(dm(i + 1)) < near_zero) { bl(i) = 0.; br(i) = 0.; }
            } else if (std::fabs(3. * (bl(i) + br(i))) > std::fabs(bl(i) - br(i))) {
              double pmp_1 = -2. * dq(i), lac_1 = pmp_1 + 1.5 * dq(i + 1);
              bl(i) = std::min(max3(0., pmp_1, lac_1), std::max(bl(i), min3(0., pmp_1, lac_1)));
              double pmp_2 = 2. * dq(i - 1), lac_2 = pmp_2 - 1.5 * dq(i - 2);
              br(i) = std::min(max3(0., pmp_2, lac_2), std::max(br(i), min3(0., pmp_2, lac_2)));
            }
          }
        } else {
          for (int i = is3; i <= ie3; i++) { bl(i) = al(i) - u(i, j); br(i) = al(i + 1) - u(i, j); }
        }
        // fix the edges
        if (is == 1 && !bounded_domain) {
          br(2) = al(3) - u(2, j);
          double xt = s15 * u(1, j) + s11 * u(2, j) - s14 * dm(2);
          bl(2) = xt - u(2, j);
          br(1) = xt - u(1, j);
          if (j == 1 || j == npy) {
            bl(0) = 0.; br(0) = 0.; bl(1) = 0.; br(1) = 0.;
          } else {
            bl(0) = s14 * dm(-1) - s11 * dq(-1);
            double x0L = 0.5 * ((2. * dx(0, j) + dx(-1, j)) * (u(0, j)) - dx(0, j) * (u(-1, j))) / (dx(0, j) + dx(-1, j));
            double x0R = 0.5 * ((2. * dx(1, j) + dx(2, j)) * (u(1, j)) - dx(1, j) * (u(2, j))) / (dx(1, j) + dx(2, j));
            xt = x0L + x0R;
            br(0) = xt - u(0, j);
            bl(1) = xt - u(1, j);
          }
          pert_ppm(1, &u(2, j), &bl(2), &br(2), -1);
        }
        if ((ie + 1) == npx && !bounded_domain) {
          bl(npx - 2) = al(npx - 2) - u(npx - 2, j);
          double xt = s15 * u(npx - 1, j) + s11 * u(npx - 2, j) + s14 * dm(npx - 2);
          br(npx - 2) = xt - u(npx - 2, j);
          bl(npx - 1) = xt - u(npx - 1, j);
          if (j == 1 || j == npy) {
            bl(npx - 1) = 0.; br(npx - 1) = 0.; bl(npx) = 0.; br(npx) = 0.;
          } else {
            br(npx) = s11 * dq(npx) - s14 * dm(npx + 1);
            double x0L = 0.5 * ((2. * dx(npx - 1, j) + dx(npx - 2, j)) * (u(npx - 1, j)) - dx(npx - 1, j) * (u(npx - 2, j))) /
                         (dx(npx - 1, j) + dx(npx - 2, j));
            double x0R = 0.5 * ((2. * dx(npx, j) + dx(npx + 1, j)) * (u(npx, j)) - dx(npx, j) * (u(npx + 1, j))) /
                         (dx(npx, j) + dx(npx + 1, j));
            xt = x0L + x0R;
            br(npx - 1) = xt - u(npx - 1, j);
            bl(npx) = xt - u(npx, j);
          }
          pert_ppm(1, &u(npx - 2, j), &bl(npx - 2), &br(npx - 2), -1);
        }
      } else {
        for (int i = is - 1; i <= ie + 2; i++) al(i) = 0.5 * (u(i - 1, j) + u(i, j)) + r3 * (dm(i - 1) - dm(i));
        for (int i = is - 1; i <= ie + 1; i++) {
          double pmp = -2. * dq(i), lac = pmp + 1.5 * dq(i + 1);
          bl(i) = std::min(max3(0., pmp, lac), std::max(al(i) - u(i, j), min3(0., pmp, lac)));
          pmp = 2. * dq(i - 1); lac = pmp - 1.5 * dq(i - 2);
          br(i) = std::min(max3(0., pmp, lac), std::max(al(i + 1) - u(i, j), min3(0., pmp, lac)));
        }
      }
      for (int i = is; i <= ie + 1; i++) {
        double cfl;
        if (c(i, j) > 0.) {
          cfl = c(i, j) * rdx(i - 1, j);
          flux(i, j) = u(i - 1, j) + (1. - cfl) * (br(i - 1) - cfl * (bl(i - 1) + br(i - 1)));
        } else {
          cfl = c(i, j) * rdx(i, j);
          flux(i, j) = u(i, j) + (1. + cfl) * (bl(i) + cfl * (bl(i) + br(i)));
        }
      }
    }
  }
}

// sw_core.F90:2524-2998
void ytp_v(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, V2 c, V2 u, V2 v, V2 flux,
           int jord, V2 dy, V2 rdy, int npx, int npy, int grid_type, bool bounded_domain, double lim_fac) {
  (void)isd; (void)ied; (void)jsd; (void)jed; (void)u;
  LB2 smt5(is, ie + 1, js - 1, je + 1), smt6(is, ie + 1, js - 1, je + 1);
  LB1 hi5(is, ie + 1), hi6(is, ie + 1);
  L1 fx0(is, ie + 1);
  L2 dm(is, ie + 1, js - 2, je + 2), al(is, ie + 1, js - 1, je + 2);
  L2 bl(is, ie + 1, js - 1, je + 1), br(is, ie + 1, js - 1, je + 1), b0(is, ie + 1, js - 1, je + 1);
  L2 dq(is, ie + 1, js - 3, je + 2);
  int js3, je3;
  if (bounded_domain || grid_type > 3) { js3 = js - 1; je3 = je + 1; }
  else { js3 = std::max(3, js - 1); je3 = std::min(npy - 3, je + 1); }
  const bool cube = (!bounded_domain) && grid_type < 3;

  if (jord < 8) {
    for (int j = js3; j <= je3 + 1; j++)
      for (int i = is; i <= ie + 1; i++) al(i, j) = p1 * (v(i, j - 1) + v(i, j)) + p2 * (v(i, j - 2) + v(i, j + 1));
    for (int j = js3; j <= je3; j++)
      for (int i = is; i <= ie + 1; i++) { bl(i, j) = al(i, j) - v(i, j); br(i, j) = al(i, j + 1) - v(i, j); }
    if (cube) {
      if (js == 1) {
        for (int i = is; i <= ie + 1; i++) {
          bl(i, 0) = c1 * v(i, -2) + c2 * v(i, -1) + c3 * v(i, 0) - v(i, 0);
          double xt = 0.5 * (((2. * dy(i, 0) + dy(i, -1)) * v(i, 0) - dy(i, 0) * v(i, -1)) / (dy(i, 0) + dy(i, -1)) +
                             ((2. * dy(i, 1) + dy(i, 2)) * v(i, 1) - dy(i, 1) * v(i, 2)) / (dy(i, 1) + dy(i, 2)));
          br(i, 0) = xt - v(i, 0);
          bl(i, 1) = xt - v(i, 1);
          xt = c3 * v(i, 1) + c2 * v(i, 2) + c1 * v(i, 3);
          br(i, 1) = xt - v(i, 1);
          bl(i, 2) = xt - v(i, 2);
          br(i, 2) = al(i, 3) - v(i, 2);
        }
        if (is == 1) { bl(1, 0) = 0.; br(1, 0) = 0.; bl(1, 1) = 0.; br(1, 1) = 0.; }
        if ((ie + 1) == npx) { bl(npx, 0) = 0.; br(npx, 0) = 0.; bl(npx, 1) = 0.; br(npx, 1) = 0.; }
      }
      if ((je + 1) == npy) {
        for (int i = is; i <= ie + 1; i++) {
          bl(i, npy - 2) = al(i, npy - 2) - v(i, npy - 2);
          double xt = c1 * v(i, npy - 3) + c2 * v(i, npy - 2) + c3 * v(i, npy - 1);
          br(i, npy - 2) = xt - v(i, npy - 2);
          bl(i, npy - 1) = xt - v(i, npy - 1);
          xt = 0.5 * (((2. * dy(i, npy - 1) + dy(i, npy - 2)) * v(i, npy - 1) - dy(i, npy - 1) * v(i, npy - 2)) /
                          (dy(i, npy - 1) + dy(i, npy - 2)) +
                      ((2. * dy(i, npy) + dy(i, npy + 1)) * v(i, npy) - dy(i, npy) * v(i, npy + 1)) /
                          (dy(i, npy) + dy(i, npy + 1)));
          br(i, npy - 1) = xt - v(i, npy - 1);
          bl(i, npy) = xt - v(i, npy);
          br(i, npy) = c3 * v(i, npy) + c2 * v(i, npy + 1) + c1 * v(i, npy + 2) - v(i, npy);
        }
        if (is == 1) { bl(1, npy - 1) = 0.; br(1, npy - 1) = 0.; bl(1, npy) = 0.; br(1, npy) = 0.; }
        if ((ie + 1) == npx) { bl(npx, npy - 1) = 0.; br(npx, npy - 1) = 0.; bl(npx, npy) = 0.; br(npx, npy) = 0.; }
      }
    }
    for (int j = js - 1; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) b0(i, j) = bl(i, j) + br(i, j);

    // the flux stage shared by jord 1,4,5,6 once the smoothness mask is known
    auto flux_masked = [&](LB2& mask, bool and5_or6) {
      (void)and5_or6;
      for (int j = js; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdy(i, j - 1); fx0(i) = (1. - cfl) * (br(i, j - 1) - cfl * b0(i, j - 1)); flux(i, j) = v(i, j - 1); }
          else { cfl = c(i, j) * rdy(i, j); fx0(i) = (1. + cfl) * (bl(i, j) + cfl * b0(i, j)); flux(i, j) = v(i, j); }
          if (mask(i, j - 1) || mask(i, j)) flux(i, j) = flux(i, j) + fx0(i);
        }
    };

    if (jord == 1) {
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) smt5(i, j) = std::fabs(lim_fac * b0(i, j)) < std::fabs(bl(i, j) - br(i, j));
      flux_masked(smt5, false);
    } else if (jord == 2) {
      for (int j = js; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdy(i, j - 1); flux(i, j) = v(i, j - 1) + (1. - cfl) * (br(i, j - 1) - cfl * b0(i, j - 1)); }
          else { cfl = c(i, j) * rdy(i, j); flux(i, j) = v(i, j) + (1. + cfl) * (bl(i, j) + cfl * b0(i, j)); }
        }
    } else if (jord == 3) {
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) {
          double x0 = std::fabs(b0(i, j)), x1 = std::fabs(bl(i, j) - br(i, j));
          smt5(i, j) = x0 < x1; smt6(i, j) = 3. * x0 < x1;
        }
      for (int j = js; j <= je + 1; j++) {
        for (int i = is; i <= ie + 1; i++) {
          fx0(i) = 0.;
          hi5(i) = smt5(i, j - 1) && smt5(i, j);
          hi6(i) = smt6(i, j - 1) || smt6(i, j);
        }
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) {
            cfl = c(i, j) * rdy(i, j - 1);
            if (hi6(i)) fx0(i) = br(i, j - 1) - cfl * b0(i, j - 1);
            else if (hi5(i)) fx0(i) = fsign(std::min(std::fabs(bl(i, j - 1)), std::fabs(br(i, j - 1))), br(i, j - 1));
            flux(i, j) = v(i, j - 1) + (1. - cfl) * fx0(i);
          } else {
            cfl = c(i, j) * rdy(i, j);
            if (hi6(i)) fx0(i) = bl(i, j) + cfl * b0(i, j);
            else if (hi5(i)) fx0(i) = fsign(std::min(std::fabs(bl(i, j)), std::fabs(br(i, j))), bl(i, j));
            flux(i, j) = v(i, j) + (1. + cfl) * fx0(i);
          }
        }
      }
    } else if (jord == 4) {
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) {
          double x0 = std::fabs(b0(i, j)), x1 = std::fabs(bl(i, j) - br(i, j));
          smt5(i, j) = x0 < x1; smt6(i, j) = 3. * x0 < x1;
        }
      for (int j = js; j <= je + 1; j++) {
        for (int i = is; i <= ie + 1; i++) {
          fx0(i) = 0.;
          hi5(i) = smt5(i, j - 1) && smt5(i, j);
          hi6(i) = smt6(i, j - 1) || smt6(i, j);
          hi5(i) = hi5(i) || hi6(i);
        }
        for (int i = is; i <= ie + 1; i++) {
          double cfl;
          if (c(i, j) > 0.) { cfl = c(i, j) * rdy(i, j - 1); fx0(i) = (1. - cfl) * (br(i, j - 1) - cfl * b0(i, j - 1)); flux(i, j) = v(i, j - 1); }
          else { cfl = c(i, j) * rdy(i, j); fx0(i) = (1. + cfl) * (bl(i, j) + cfl * b0(i, j)); flux(i, j) = v(i, j); }
          if (hi5(i)) flux(i, j) = flux(i, j) + fx0(i);
        }
      }
    } else {
      if (jord == 5) {
        for (int j = js - 1; j <= je + 1; j++)
          for (int i = is; i <= ie + 1; i++) smt5(i, j) = bl(i, j) * br(i, j) < 0.;
        flux_masked(smt5, false);
      } else {
        for (int j = js - 1; j <= je + 1; j++)
          for (int i = is; i <= ie + 1; i++) smt6(i, j) = 3. * std::fabs(b0(i, j)) < std::fabs(bl(i, j) - br(i, j));
        if (cube) {
          if (js == 1)
            for (int i = is; i <= ie + 1; i++) { smt6(i, 0) = bl(i, 0) * br(i, 0) < 0.; smt6(i, 1) = bl(i, 1) * br(i, 1) < 0.; }
          if ((je + 1) == npy)
            for (int i = is; i <= ie + 1; i++) {
              smt6(i, npy - 1) = bl(i, npy - 1) * br(i, npy - 1) < 0.;
              smt6(i, npy) = bl(i, npy) * br(i, npy) < 0.;
            }
        }
        flux_masked(smt6, false);
      }
    }
  } else {
    for (int j = js - 2; j <= je + 2; j++)
      for (int i = is; i <= ie + 1; i++) {
        double xt = 0.25 * (v(i, j + 1) - v(i, j - 1));
        dm(i, j) = fsign(std::min(std::min(std::fabs(xt), max3(v(i, j - 1), v(i, j), v(i, j + 1)) - v(i, j)),
                                  v(i, j) - min3(v(i, j - 1), v(i, j), v(i, j + 1))), xt);
      }
    for (int j = js - 3; j <= je + 2; j++)
      for (int i = is; i <= ie + 1; i++) dq(i, j) = v(i, j + 1) - v(i, j);

    if (grid_type < 3) {
      for (int j = js3; j <= je3 + 1; j++)
        for (int i = is; i <= ie + 1; i++) al(i, j) = 0.5 * (v(i, j - 1) + v(i, j)) + r3 * (dm(i, j - 1) - dm(i, j));
      if (jord == 8) {
        for (int j = js3; j <= je3; j++)
          for (int i = is; i <= ie + 1; i++) {
            double xt = 2. * dm(i, j);
            bl(i, j) = -fsign(std::min(std::fabs(xt), std::fabs(al(i, j) - v(i, j))), xt);
            br(i, j) = fsign(std::min(std::fabs(xt), std::fabs(al(i, j + 1) - v(i, j))), xt);
          }
      } else if (jord == 9) {
        for (int j = js3; j <= je3; j++)
          for (int i = is; i <= ie + 1; i++) {
            double pmp_1 = -2. * dq(i, j), lac_1 = pmp_1 + 1.5 * dq(i, j + 1);
            bl(i, j) = std::min(max3(0., pmp_1, lac_1), std::max(al(i, j) - v(i, j), min3(0., pmp_1, lac_1)));
            double pmp_2 = 2. * dq(i, j - 1), lac_2 = pmp_2 - 1.5 * dq(i, j - 2);
            br(i, j) = std::min(max3(0., pmp_2, lac_2), std::max(al(i, j + 1) - v(i, j), min3(0., pmp_2, lac_2)));
          }
      } else if (jord == 10) {
        for (int j = js3; j <= je3; j++)
          for (int i = is; i <= ie + 1; i++) {
            bl(i, j) = al(i, j) - v(i, j);
            br(i, j) = al(i, j + 1) - v(i, j);
            if (std::fabs(dm(i, j)) < near_zero) {
              if (std::fabs(dm(i, j - 1)) + std::fabs(dm(i, j + 1)) < near_zero) { bl(i, j) = 0.; br(i, j) = 0.; }
            } else if (std::fabs(3. * (bl(i, j) + br(i, j))) > std::fabs(bl(i, j) - br(i, j))) {
              double pmp_1 = -2. * dq(i, j), lac_1 = pmp_1 + 1.5 * dq(i, j + 1);
              bl(i, j) = std::min(max3(0., pmp_1, lac_1), std::max(bl(i, j), min3(0., pmp_1, lac_1)));
              double pmp_2 = 2. * dq(i, j - 1), lac_2 = pmp_2 - 1.5 * dq(i, j - 2);
              br(i, j) = std::min(max3(0., pmp_2, lac_2), std::max(br(i, j), min3(0., pmp_2, lac_2)));
            }
          }
      } else {
        for (int j = js3; j <= je3; j++)
          for (int i = is; i <= ie + 1; i++) { bl(i, j) = al(i, j) - v(i, j); br(i, j) = al(i, j + 1) - v(i, j); }
      }
      // fix the edges
      if (js == 1 && !bounded_domain) {
        for (int i = is; i <= ie + 1; i++) {
          br(i, 2) = al(i, 3) - v(i, 2);
          double xt = s15 * v(i, 1) + s11 * v(i, 2) - s14 * dm(i, 2);
          br(i, 1) = xt - v(i, 1);
          bl(i, 2) = xt - v(i, 2);
          bl(i, 0) = s14 * dm(i, -1) - s11 * dq(i, -1);
          double x0L = 0.5 * ((2. * dy(i, 0) + dy(i, -1)) * (v(i, 0)) - dy(i, 0) * (v(i, -1))) / (dy(i, 0) + dy(i, -1));
          double x0R = 0.5 * ((2. * dy(i, 1) + dy(i, 2)) * (v(i, 1)) - dy(i, 1) * (v(i, 2))) / (dy(i, 1) + dy(i, 2));
          xt = x0L + x0R;
          bl(i, 1) = xt - v(i, 1);
          br(i, 0) = xt - v(i, 0);
        }
        if (is == 1) { bl(1, 0) = 0.; br(1, 0) = 0.; bl(1, 1) = 0.; br(1, 1) = 0.; }
        if ((ie + 1) == npx) { bl(npx, 0) = 0.; br(npx, 0) = 0.; bl(npx, 1) = 0.; br(npx, 1) = 0.; }
        for (int i = is; i <= ie + 1; i++) pert_ppm(1, &v(i, 2), &bl(i, 2), &br(i, 2), -1);
      }
      if ((je + 1) == npy && !bounded_domain) {
        for (int i = is; i <= ie + 1; i++) {
          bl(i, npy - 2) = al(i, npy - 2) - v(i, npy - 2);
          double xt = s15 * v(i, npy - 1) + s11 * v(i, npy - 2) + s14 * dm(i, npy - 2);
          br(i, npy - 2) = xt - v(i, npy - 2);
          bl(i, npy - 1) = xt - v(i, npy - 1);
          br(i, npy) = s11 * dq(i, npy) - s14 * dm(i, npy + 1);
          double x0L = 0.5 * ((2. * dy(i, npy - 1) + dy(i, npy - 2)) * (v(i, npy - 1)) - dy(i, npy - 1) * (v(i, npy - 2))) /
                       (dy(i, npy - 1) + dy(i, npy - 2));
          double x0R = 0.5 * ((2. * dy(i, npy) + dy(i, npy + 1)) * (v(i, npy)) - dy(i, npy) * (v(i, npy + 1))) /
                       (dy(i, npy) + dy(i, npy + 1));
          xt = x0L + x0R;
          br(i, npy - 1) = xt - v(i, npy - 1);
          bl(i, npy) = xt - v(i, npy);
        }
        if (is == 1) { bl(1, npy - 1) = 0.; br(1, npy - 1) = 0.; bl(1, npy) = 0.; br(1, npy) = 0.; }
        if ((ie + 1) == npx) { bl(npx, npy - 1) = 0.; br(npx, npy - 1) = 0.; bl(npx, npy) = 0.; br(npx, npy) = 0.; }
        for (int i = is; i <= ie + 1; i++) pert_ppm(1, &v(i, npy - 2), &bl(i, npy - 2), &br(i, npy - 2), -1);
      }
    } else {
      for (int j = js - 1; j <= je + 2; j++)
        for (int i = is; i <= ie + 1; i++) al(i, j) = 0.5 * (v(i, j - 1) + v(i, j)) + r3 * (dm(i, j - 1) - dm(i, j));
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) {
          double pmp = 2. * dq(i, j - 1), lac = pmp - 1.5 * dq(i, j - 2);
          br(i, j) = std::min(max3(0., pmp, lac), std::max(al(i, j + 1) - v(i, j), min3(0., pmp, lac)));
          pmp = -2. * dq(i, j); lac = pmp + 1.5 * dq(i, j + 1);
          bl(i, j) = std::min(max3(0., pmp, lac), std::max(al(i, j) - v(i, j), min3(0., pmp, lac)));
        }
    }
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) {
        double cfl;
        if (c(i, j) > 0.) {
          cfl = c(i, j) * rdy(i, j - 1);
          flux(i, j) = v(i, j - 1) + (1. - cfl) * (br(i, j - 1) - cfl * (bl(i, j - 1) + br(i, j - 1)));
        } else {
          cfl = c(i, j) * rdy(i, j);
          flux(i, j) = v(i, j) + (1. + cfl) * (bl(i, j) + cfl * (bl(i, j) + br(i, j)));
        }
      }
  }
}

// sw_core.F90:494-1606 (non-SW_DYNAMICS; inline_q=F; flagstruct%do_f3d only without ROT3)
// + the SW_DYNAMICS / test_case == 1 pure-advection branch (a.sw_test_case == 1, BASELINE config 1a)
void d_sw(V2 delpc, V2 delp, V2 ptc, V2 pt, V2 u, V2 v, V2 w, V2 uc, V2 vc, V2 ua, V2 va, V2 divg_d,
          V2 xflux, V2 yflux, V2 cx, V2 cy, V2 crx_adv, V2 cry_adv, V2 xfx_adv, V2 yfx_adv, V2 q_con,
          V2 z_rat, V2 heat_source, V2 diss_est, const DswArgs& a, const Grid& g, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const int npx = bd.npx, npy = bd.npy, ng = bd.ng, grid_type = bd.grid_type;
  const bool bounded = bd.bounded_domain;
  const double dt = a.dt;
  L2 ut(isd, ied + 1, jsd, jed), vt(isd, ied, jsd, jed + 1);
  L2 fx2(isd, ied + 1, jsd, jed), fy2(isd, ied, jsd, jed + 1);
  L2 dw(is, ie, js, je);
  L2 ub(is, ie + 1, js, je + 1), vb(is, ie + 1, js, je + 1);
  L2 wk(isd, ied, jsd, jed), ke(isd, ied + 1, jsd, jed + 1), vort(isd, ied, jsd, jed);
  L2 fx(is, ie + 1, js, je), fy(is, ie, js, je + 1);
  L2 ra_x(is, ie, jsd, jed), ra_y(isd, ied, js, je);
  L2 gx(is, ie + 1, js, je), gy(is, ie, js, je + 1);
  double damp, damp2, damp4, dd8;
  int is2, ie1, js2, je1;

  if (a.sw_test_case == 1) {
    // sw_core.F90:626-651 (SW_DYNAMICS, test_case == 1): Courant numbers and area fluxes from the prescribed C-grid winds
    for (int j = jsd; j <= jed; j++)
      for (int i = is; i <= ie + 1; i++) {
        xfx_adv(i, j) = dt * uc(i, j) / g.sina_u(i, j);
        if (xfx_adv(i, j) > 0.) crx_adv(i, j) = xfx_adv(i, j) * g.rdxa(i - 1, j);
        else crx_adv(i, j) = xfx_adv(i, j) * g.rdxa(i, j);
        xfx_adv(i, j) = g.dy(i, j) * xfx_adv(i, j) * g.sina_u(i, j);
      }
    for (int j = js; j <= je + 1; j++)
      for (int i = isd; i <= ied; i++) {
        yfx_adv(i, j) = dt * vc(i, j) / g.sina_v(i, j);
        if (yfx_adv(i, j) > 0.) cry_adv(i, j) = yfx_adv(i, j) * g.rdya(i, j - 1);
        else cry_adv(i, j) = yfx_adv(i, j) * g.rdya(i, j);
        yfx_adv(i, j) = g.dx(i, j) * yfx_adv(i, j) * g.sina_v(i, j);
      }
    // :909-940 common part: ra_x, ra_y, transport of delp, flux capacitors
    for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie; i++) ra_x(i, j) = g.area(i, j) + xfx_adv(i, j) - xfx_adv(i + 1, j);
    for (int j = js; j <= je; j++) for (int i = isd; i <= ied; i++) ra_y(i, j) = g.area(i, j) + yfx_adv(i, j) - yfx_adv(i, j + 1);
    fv_tp_2d(delp, crx_adv, cry_adv, npx, npy, a.hord_dp, fx, fy, xfx_adv, yfx_adv, g, bd, ra_x, ra_y, a.lim_fac,
             nullptr, nullptr, nullptr, true, a.nord_v, a.damp_v);
    for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie + 1; i++) cx(i, j) = cx(i, j) + crx_adv(i, j);
    for (int j = js; j <= je; j++) for (int i = is; i <= ie + 1; i++) xflux(i, j) = xflux(i, j) + fx(i, j);
    for (int j = js; j <= je + 1; j++) {
      for (int i = isd; i <= ied; i++) cy(i, j) = cy(i, j) + cry_adv(i, j);
      for (int i = is; i <= ie; i++) yflux(i, j) = yflux(i, j) + fy(i, j);
    }
    // :1055-1066 with SW_DYNAMICS: only delp is updated (pt untouched); :1069, :1602: the momentum part is skipped
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++)
        delp(i, j) = delp(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * g.rarea(i, j);
    return;
  }

  if (grid_type < 3) {
    if (bounded) {
      for (int j = jsd; j <= jed; j++)
        for (int i = is; i <= ie + 1; i++)
          ut(i, j) = (uc(i, j) - 0.25 * g.cosa_u(i, j) * (vc(i - 1, j) + vc(i, j) + vc(i - 1, j + 1) + vc(i, j + 1))) * g.rsin_u(i, j);
      for (int j = js; j <= je + 1; j++)
        for (int i = isd; i <= ied; i++)
          vt(i, j) = (vc(i, j) - 0.25 * g.cosa_v(i, j) * (uc(i, j - 1) + uc(i + 1, j - 1) + uc(i, j) + uc(i + 1, j))) * g.rsin_v(i, j);
    } else {
      for (int j = jsd; j <= jed; j++)
        if (j != 0 && j != 1 && j != (npy - 1) && j != npy)
          for (int i = is - 1; i <= ie + 2; i++)
            ut(i, j) = (uc(i, j) - 0.25 * g.cosa_u(i, j) * (vc(i - 1, j) + vc(i, j) + vc(i - 1, j + 1) + vc(i, j + 1))) * g.rsin_u(i, j);
      for (int j = js - 1; j <= je + 2; j++)
        if (j != 1 && j != npy)
          for (int i = isd; i <= ied; i++)
            vt(i, j) = (vc(i, j) - 0.25 * g.cosa_v(i, j) * (uc(i, j - 1) + uc(i + 1, j - 1) + uc(i, j) + uc(i + 1, j))) * g.rsin_v(i, j);
    }
    if (!bounded) {
      if (is == 1) {  // West edge
        for (int j = jsd; j <= jed; j++) {
          if (uc(1, j) * dt > 0.) ut(1, j) = uc(1, j) / g.sin_sg(0, j, 3);
          else ut(1, j) = uc(1, j) / g.sin_sg(1, j, 1);
        }
        for (int j = std::max(3, js); j <= std::min(npy - 2, je + 1); j++) {
          vt(0, j) = vc(0, j) - 0.25 * g.cosa_v(0, j) * (ut(0, j - 1) + ut(1, j - 1) + ut(0, j) + ut(1, j));
          vt(1, j) = vc(1, j) - 0.25 * g.cosa_v(1, j) * (ut(1, j - 1) + ut(2, j - 1) + ut(1, j) + ut(2, j));
        }
      }
      if ((ie + 1) == npx) {  // East edge
        for (int j = jsd; j <= jed; j++) {
          if (uc(npx, j) * dt > 0.) ut(npx, j) = uc(npx, j) / g.sin_sg(npx - 1, j, 3);
          else ut(npx, j) = uc(npx, j) / g.sin_sg(npx, j, 1);
        }
        for (int j = std::max(3, js); j <= std::min(npy - 2, je + 1); j++) {
          vt(npx - 1, j) = vc(npx - 1, j) - 0.25 * g.cosa_v(npx - 1, j) * (ut(npx - 1, j - 1) + ut(npx, j - 1) + ut(npx - 1, j) + ut(npx, j));
          vt(npx, j) = vc(npx, j) - 0.25 * g.cosa_v(npx, j) * (ut(npx, j - 1) + ut(npx + 1, j - 1) + ut(npx, j) + ut(npx + 1, j));
        }
      }
      if (js == 1) {  // South edge
        for (int i = isd; i <= ied; i++) {
          if (vc(i, 1) * dt > 0.) vt(i, 1) = vc(i, 1) / g.sin_sg(i, 0, 4);
          else vt(i, 1) = vc(i, 1) / g.sin_sg(i, 1, 2);
        }
        for (int i = std::max(3, is); i <= std::min(npx - 2, ie + 1); i++) {
          ut(i, 0) = uc(i, 0) - 0.25 * g.cosa_u(i, 0) * (vt(i - 1, 0) + vt(i, 0) + vt(i - 1, 1) + vt(i, 1));
          ut(i, 1) = uc(i, 1) - 0.25 * g.cosa_u(i, 1) * (vt(i - 1, 1) + vt(i, 1) + vt(i - 1, 2) + vt(i, 2));
        }
      }
      if ((je + 1) == npy) {  // North edge
        for (int i = isd; i <= ied; i++) {
          if (vc(i, npy) * dt > 0.) vt(i, npy) = vc(i, npy) / g.sin_sg(i, npy - 1, 4);
          else vt(i, npy) = vc(i, npy) / g.sin_sg(i, npy, 2);
        }
        for (int i = std::max(3, is); i <= std::min(npx - 2, ie + 1); i++) {
          ut(i, npy - 1) = uc(i, npy - 1) - 0.25 * g.cosa_u(i, npy - 1) * (vt(i - 1, npy - 1) + vt(i, npy - 1) + vt(i - 1, npy) + vt(i, npy));
          ut(i, npy) = uc(i, npy) - 0.25 * g.cosa_u(i, npy) * (vt(i - 1, npy) + vt(i, npy) + vt(i - 1, npy + 1) + vt(i, npy + 1));
        }
      }
      // 2x2 corner systems, sw_core.F90:772-844
      if (bd.sw_corner) {
        damp = 1. / (1. - 0.0625 * g.cosa_u(2, 0) * g.cosa_v(1, 0));
        ut(2, 0) = (uc(2, 0) - 0.25 * g.cosa_u(2, 0) * (vt(1, 1) + vt(2, 1) + vt(2, 0) + vc(1, 0) -
                    0.25 * g.cosa_v(1, 0) * (ut(1, 0) + ut(1, -1) + ut(2, -1)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(0, 1) * g.cosa_v(0, 2));
        vt(0, 2) = (vc(0, 2) - 0.25 * g.cosa_v(0, 2) * (ut(1, 1) + ut(1, 2) + ut(0, 2) + uc(0, 1) -
                    0.25 * g.cosa_u(0, 1) * (vt(0, 1) + vt(-1, 1) + vt(-1, 2)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(2, 1) * g.cosa_v(1, 2));
        ut(2, 1) = (uc(2, 1) - 0.25 * g.cosa_u(2, 1) * (vt(1, 1) + vt(2, 1) + vt(2, 2) + vc(1, 2) -
                    0.25 * g.cosa_v(1, 2) * (ut(1, 1) + ut(1, 2) + ut(2, 2)))) * damp;
        vt(1, 2) = (vc(1, 2) - 0.25 * g.cosa_v(1, 2) * (ut(1, 1) + ut(1, 2) + ut(2, 2) + uc(2, 1) -
                    0.25 * g.cosa_u(2, 1) * (vt(1, 1) + vt(2, 1) + vt(2, 2)))) * damp;
      }
      if (bd.se_corner) {
        damp = 1. / (1. - 0.0625 * g.cosa_u(npx - 1, 0) * g.cosa_v(npx - 1, 0));
        ut(npx - 1, 0) = (uc(npx - 1, 0) - 0.25 * g.cosa_u(npx - 1, 0) * (vt(npx - 1, 1) + vt(npx - 2, 1) + vt(npx - 2, 0) + vc(npx - 1, 0) -
                          0.25 * g.cosa_v(npx - 1, 0) * (ut(npx, 0) + ut(npx, -1) + ut(npx - 1, -1)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(npx + 1, 1) * g.cosa_v(npx, 2));
        vt(npx, 2) = (vc(npx, 2) - 0.25 * g.cosa_v(npx, 2) * (ut(npx, 1) + ut(npx, 2) + ut(npx + 1, 2) + uc(npx + 1, 1) -
                      0.25 * g.cosa_u(npx + 1, 1) * (vt(npx, 1) + vt(npx + 1, 1) + vt(npx + 1, 2)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(npx - 1, 1) * g.cosa_v(npx - 1, 2));
        ut(npx - 1, 1) = (uc(npx - 1, 1) - 0.25 * g.cosa_u(npx - 1, 1) * (vt(npx - 1, 1) + vt(npx - 2, 1) + vt(npx - 2, 2) + vc(npx - 1, 2) -
                          0.25 * g.cosa_v(npx - 1, 2) * (ut(npx, 1) + ut(npx, 2) + ut(npx - 1, 2)))) * damp;
        vt(npx - 1, 2) = (vc(npx - 1, 2) - 0.25 * g.cosa_v(npx - 1, 2) * (ut(npx, 1) + ut(npx, 2) + ut(npx - 1, 2) + uc(npx - 1, 1) -
                          0.25 * g.cosa_u(npx - 1, 1) * (vt(npx - 1, 1) + vt(npx - 2, 1) + vt(npx - 2, 2)))) * damp;
      }
      if (bd.ne_corner) {
        damp = 1. / (1. - 0.0625 * g.cosa_u(npx - 1, npy) * g.cosa_v(npx - 1, npy + 1));
        ut(npx - 1, npy) = (uc(npx - 1, npy) - 0.25 * g.cosa_u(npx - 1, npy) * (vt(npx - 1, npy) + vt(npx - 2, npy) + vt(npx - 2, npy + 1) + vc(npx - 1, npy + 1) -
                            0.25 * g.cosa_v(npx - 1, npy + 1) * (ut(npx, npy) + ut(npx, npy + 1) + ut(npx - 1, npy + 1)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(npx + 1, npy - 1) * g.cosa_v(npx, npy - 1));
        vt(npx, npy - 1) = (vc(npx, npy - 1) - 0.25 * g.cosa_v(npx, npy - 1) * (ut(npx, npy - 1) + ut(npx, npy - 2) + ut(npx + 1, npy - 2) + uc(npx + 1, npy - 1) -
                            0.25 * g.cosa_u(npx + 1, npy - 1) * (vt(npx, npy) + vt(npx + 1, npy) + vt(npx + 1, npy - 1)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(npx - 1, npy - 1) * g.cosa_v(npx - 1, npy - 1));
        ut(npx - 1, npy - 1) = (uc(npx - 1, npy - 1) - 0.25 * g.cosa_u(npx - 1, npy - 1) * (vt(npx - 1, npy) + vt(npx - 2, npy) + vt(npx - 2, npy - 1) + vc(npx - 1, npy - 1) -
                                0.25 * g.cosa_v(npx - 1, npy - 1) * (ut(npx, npy - 1) + ut(npx, npy - 2) + ut(npx - 1, npy - 2)))) * damp;
        vt(npx - 1, npy - 1) = (vc(npx - 1, npy - 1) - 0.25 * g.cosa_v(npx - 1, npy - 1) * (ut(npx, npy - 1) + ut(npx, npy - 2) + ut(npx - 1, npy - 2) + uc(npx - 1, npy - 1) -
                                0.25 * g.cosa_u(npx - 1, npy - 1) * (vt(npx - 1, npy) + vt(npx - 2, npy) + vt(npx - 2, npy - 1)))) * damp;
      }
      if (bd.nw_corner) {
        damp = 1. / (1. - 0.0625 * g.cosa_u(2, npy) * g.cosa_v(1, npy + 1));
        ut(2, npy) = (uc(2, npy) - 0.25 * g.cosa_u(2, npy) * (vt(1, npy) + vt(2, npy) + vt(2, npy + 1) + vc(1, npy + 1) -
                      0.25 * g.cosa_v(1, npy + 1) * (ut(1, npy) + ut(1, npy + 1) + ut(2, npy + 1)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(0, npy - 1) * g.cosa_v(0, npy - 1));
        vt(0, npy - 1) = (vc(0, npy - 1) - 0.25 * g.cosa_v(0, npy - 1) * (ut(1, npy - 1) + ut(1, npy - 2) + ut(0, npy - 2) + uc(0, npy - 1) -
                          0.25 * g.cosa_u(0, npy - 1) * (vt(0, npy) + vt(-1, npy) + vt(-1, npy - 1)))) * damp;
        damp = 1. / (1. - 0.0625 * g.cosa_u(2, npy - 1) * g.cosa_v(1, npy - 1));
        ut(2, npy - 1) = (uc(2, npy - 1) - 0.25 * g.cosa_u(2, npy - 1) * (vt(1, npy) + vt(2, npy) + vt(2, npy - 1) + vc(1, npy - 1) -
                          0.25 * g.cosa_v(1, npy - 1) * (ut(1, npy - 1) + ut(1, npy - 2) + ut(2, npy - 2)))) * damp;
        vt(1, npy - 1) = (vc(1, npy - 1) - 0.25 * g.cosa_v(1, npy - 1) * (ut(1, npy - 1) + ut(1, npy - 2) + ut(2, npy - 2) + uc(2, npy - 1) -
                          0.25 * g.cosa_u(2, npy - 1) * (vt(1, npy) + vt(2, npy) + vt(2, npy - 1)))) * damp;
      }
    }
  } else {
    for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie + 1; i++) ut(i, j) = uc(i, j);
    for (int j = js; j <= je + 1; j++) for (int i = isd; i <= ied; i++) vt(i, j) = vc(i, j);
  }

  for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie + 1; i++) xfx_adv(i, j) = dt * ut(i, j);
  for (int j = js; j <= je + 1; j++) for (int i = isd; i <= ied; i++) yfx_adv(i, j) = dt * vt(i, j);
  for (int j = jsd; j <= jed; j++)
    for (int i = is; i <= ie + 1; i++) {
      if (xfx_adv(i, j) > 0.) {
        crx_adv(i, j) = xfx_adv(i, j) * g.rdxa(i - 1, j);
        xfx_adv(i, j) = g.dy(i, j) * xfx_adv(i, j) * g.sin_sg(i - 1, j, 3);
      } else {
        crx_adv(i, j) = xfx_adv(i, j) * g.rdxa(i, j);
        xfx_adv(i, j) = g.dy(i, j) * xfx_adv(i, j) * g.sin_sg(i, j, 1);
      }
    }
  for (int j = js; j <= je + 1; j++)
    for (int i = isd; i <= ied; i++) {
      if (yfx_adv(i, j) > 0.) {
        cry_adv(i, j) = yfx_adv(i, j) * g.rdya(i, j - 1);
        yfx_adv(i, j) = g.dx(i, j) * yfx_adv(i, j) * g.sin_sg(i, j - 1, 4);
      } else {
        cry_adv(i, j) = yfx_adv(i, j) * g.rdya(i, j);
        yfx_adv(i, j) = g.dx(i, j) * yfx_adv(i, j) * g.sin_sg(i, j, 2);
      }
    }
  for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie; i++) ra_x(i, j) = g.area(i, j) + xfx_adv(i, j) - xfx_adv(i + 1, j);
  for (int j = js; j <= je; j++) for (int i = isd; i <= ied; i++) ra_y(i, j) = g.area(i, j) + yfx_adv(i, j) - yfx_adv(i, j + 1);

  fv_tp_2d(delp, crx_adv, cry_adv, npx, npy, a.hord_dp, fx, fy, xfx_adv, yfx_adv, g, bd, ra_x, ra_y, a.lim_fac,
           nullptr, nullptr, nullptr, true, a.nord_v, a.damp_v);

  // flux capacitors
  for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie + 1; i++) cx(i, j) = cx(i, j) + crx_adv(i, j);
  for (int j = js; j <= je; j++) for (int i = is; i <= ie + 1; i++) xflux(i, j) = xflux(i, j) + fx(i, j);
  for (int j = js; j <= je + 1; j++) {
    for (int i = isd; i <= ied; i++) cy(i, j) = cy(i, j) + cry_adv(i, j);
    for (int i = is; i <= ie; i++) yflux(i, j) = yflux(i, j) + fy(i, j);
  }
  for (int j = js; j <= je; j++) for (int i = is; i <= ie; i++) { heat_source(i, j) = 0.; diss_est(i, j) = 0.; }

  if (!a.hydrostatic) {
    if (a.damp_w > 1.E-5) {
      dd8 = a.kgb * std::fabs(dt);
      damp4 = std::pow(a.damp_w * g.da_min_c, (double)(a.nord_w + 1));
      del6_vt_flux(a.nord_w, npx, npy, damp4, w, wk, fx2, fy2, g, bd);
      if (a.prevent_diss_cooling) {
        for (int j = js; j <= je; j++)
          for (int i = is; i <= ie; i++) {
            dw(i, j) = (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * g.rarea(i, j);
            double tmp = dw(i, j) * (w(i, j) + 0.5 * dw(i, j));
            heat_source(i, j) = dd8 - std::min(0., tmp);
            if (a.do_diss_est) diss_est(i, j) = dd8 - tmp;
          }
      } else {
        for (int j = js; j <= je; j++)
          for (int i = is; i <= ie; i++) {
            dw(i, j) = (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * g.rarea(i, j);
            heat_source(i, j) = dd8 - dw(i, j) * (w(i, j) + 0.5 * dw(i, j));
            if (a.do_diss_est) diss_est(i, j) = heat_source(i, j);
          }
      }
    }
    fv_tp_2d(w, crx_adv, cry_adv, npx, npy, a.hord_vt, gx, gy, xfx_adv, yfx_adv, g, bd, ra_x, ra_y, a.lim_fac,
             &fx, &fy, nullptr, false, 0, 0.);
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++)
        w(i, j) = delp(i, j) * w(i, j) + (gx(i, j) - gx(i + 1, j) + gy(i, j) - gy(i, j + 1)) * g.rarea(i, j);
  }
  if (a.use_cond) {
    fv_tp_2d(q_con, crx_adv, cry_adv, npx, npy, a.hord_dp, gx, gy, xfx_adv, yfx_adv, g, bd, ra_x, ra_y, a.lim_fac,
             &fx, &fy, &delp, true, a.nord_t, a.damp_t);
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++)
        q_con(i, j) = delp(i, j) * q_con(i, j) + (gx(i, j) - gx(i + 1, j) + gy(i, j) - gy(i, j + 1)) * g.rarea(i, j);
  }
  // AM4 variant (no GFS_PHYS): pt damping uses nord_t, damp_t  (sw_core.F90:1014-1016)
  fv_tp_2d(pt, crx_adv, cry_adv, npx, npy, a.hord_tm, gx, gy, xfx_adv, yfx_adv, g, bd, ra_x, ra_y, a.lim_fac,
           &fx, &fy, &delp, true, a.nord_t, a.damp_t);
  for (int j = js; j <= je; j++)
    for (int i = is; i <= ie; i++) {
      pt(i, j) = pt(i, j) * delp(i, j) + (gx(i, j) - gx(i + 1, j) + gy(i, j) - gy(i, j + 1)) * g.rarea(i, j);
      delp(i, j) = delp(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * g.rarea(i, j);
      pt(i, j) = pt(i, j) / delp(i, j);
    }

  // Kinetic energy fluxes
  const double dt5 = 0.5 * dt, dt4 = 0.25 * dt;
  if (bounded) { is2 = is; ie1 = ie + 1; js2 = js; je1 = je + 1; }
  else { is2 = std::max(2, is); ie1 = std::min(npx - 1, ie + 1); js2 = std::max(2, js); je1 = std::min(npy - 1, je + 1); }

  if (grid_type < 3) {
    if (bounded) {
      for (int j = js2; j <= je1; j++)
        for (int i = is2; i <= ie1; i++)
          vb(i, j) = dt5 * (vc(i - 1, j) + vc(i, j) - (uc(i, j - 1) + uc(i, j)) * g.cosa(i, j)) * g.rsina(i, j);
    } else {
      if (js == 1) for (int i = is; i <= ie + 1; i++) vb(i, 1) = dt5 * (vt(i - 1, 1) + vt(i, 1));
      for (int j = js2; j <= je1; j++) {
        for (int i = is2; i <= ie1; i++)
          vb(i, j) = dt5 * (vc(i - 1, j) + vc(i, j) - (uc(i, j - 1) + uc(i, j)) * g.cosa(i, j)) * g.rsina(i, j);
        if (is == 1) vb(1, j) = dt4 * (-vt(-1, j) + 3. * (vt(0, j) + vt(1, j)) - vt(2, j));
        if ((ie + 1) == npx) vb(npx, j) = dt4 * (-vt(npx - 2, j) + 3. * (vt(npx - 1, j) + vt(npx, j)) - vt(npx + 1, j));
      }
      if ((je + 1) == npy) for (int i = is; i <= ie + 1; i++) vb(i, npy) = dt5 * (vt(i - 1, npy) + vt(i, npy));
    }
  } else {
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) vb(i, j) = dt5 * (vc(i - 1, j) + vc(i, j));
  }
  ytp_v(is, ie, js, je, isd, ied, jsd, jed, vb, u, v, ub, a.hord_mt, g.dy, g.rdy, npx, npy, grid_type, bounded, a.lim_fac);
  for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) ke(i, j) = vb(i, j) * ub(i, j);

  if (grid_type < 3) {
    if (bounded) {
      for (int j = js; j <= je + 1; j++)
        for (int i = is2; i <= ie1; i++)
          ub(i, j) = dt5 * (uc(i, j - 1) + uc(i, j) - (vc(i - 1, j) + vc(i, j)) * g.cosa(i, j)) * g.rsina(i, j);
    } else {
      if (is == 1) for (int j = js; j <= je + 1; j++) ub(1, j) = dt5 * (ut(1, j - 1) + ut(1, j));
      for (int j = js; j <= je + 1; j++) {
        if (j == 1 || j == npy) {
          for (int i = is2; i <= ie1; i++) ub(i, j) = dt4 * (-ut(i, j - 2) + 3. * (ut(i, j - 1) + ut(i, j)) - ut(i, j + 1));
        } else {
          for (int i = is2; i <= ie1; i++)
            ub(i, j) = dt5 * (uc(i, j - 1) + uc(i, j) - (vc(i - 1, j) + vc(i, j)) * g.cosa(i, j)) * g.rsina(i, j);
        }
      }
      if ((ie + 1) == npx) for (int j = js; j <= je + 1; j++) ub(npx, j) = dt5 * (ut(npx, j - 1) + ut(npx, j));
    }
  } else {
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) ub(i, j) = dt5 * (uc(i, j - 1) + uc(i, j));
  }
  xtp_u(is, ie, js, je, isd, ied, jsd, jed, ub, u, v, vb, a.hord_mt, g.dx, g.rdx, npx, npy, grid_type, bounded, a.lim_fac);
  for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) ke(i, j) = 0.5 * (ke(i, j) + ub(i, j) * vb(i, j));

  // Fix KE at the 4 corners of the face
  if (!bounded) {
    const double dt6 = dt / 6.;
    if (bd.sw_corner)
      ke(1, 1) = dt6 * ((ut(1, 1) + ut(1, 0)) * u(1, 1) + (vt(1, 1) + vt(0, 1)) * v(1, 1) + (ut(1, 1) + vt(1, 1)) * u(0, 1));
    if (bd.se_corner) {
      int i = npx;
      ke(i, 1) = dt6 * ((ut(i, 1) + ut(i, 0)) * u(i - 1, 1) + (vt(i, 1) + vt(i - 1, 1)) * v(i, 1) + (ut(i, 1) - vt(i - 1, 1)) * u(i, 1));
    }
    if (bd.ne_corner) {
      int i = npx, j = npy;
      ke(i, j) = dt6 * ((ut(i, j) + ut(i, j - 1)) * u(i - 1, j) + (vt(i, j) + vt(i - 1, j)) * v(i, j - 1) + (ut(i, j - 1) + vt(i - 1, j)) * u(i, j));
    }
    if (bd.nw_corner) {
      int j = npy;
      ke(1, j) = dt6 * ((ut(1, j) + ut(1, j - 1)) * u(1, j) + (vt(1, j) + vt(0, j)) * v(1, j - 1) + (ut(1, j - 1) - vt(1, j)) * u(0, j));
    }
  }

  // Compute vorticity
  for (int j = jsd; j <= jed + 1; j++) for (int i = isd; i <= ied; i++) vt(i, j) = u(i, j) * g.dx(i, j);
  for (int j = jsd; j <= jed; j++) for (int i = isd; i <= ied + 1; i++) ut(i, j) = v(i, j) * g.dy(i, j);
  for (int j = jsd; j <= jed; j++)
    for (int i = isd; i <= ied; i++) wk(i, j) = g.rarea(i, j) * (vt(i, j) - vt(i, j + 1) - ut(i, j) + ut(i + 1, j));

  if (!a.hydrostatic) {
    for (int j = js; j <= je; j++) for (int i = is; i <= ie; i++) w(i, j) = w(i, j) / delp(i, j);
    if (a.damp_w > 1.E-5)
      for (int j = js; j <= je; j++) for (int i = is; i <= ie; i++) w(i, j) = w(i, j) + dw(i, j);
  }
  if (a.use_cond)
    for (int j = js; j <= je; j++) for (int i = is; i <= ie; i++) q_con(i, j) = q_con(i, j) / delp(i, j);

  // Divergence damping
  if (a.nord == 0) {
    if (bounded) {
      for (int j = js; j <= je + 1; j++)
        for (int i = is - 1; i <= ie + 1; i++)
          ptc(i, j) = (u(i, j) - 0.5 * (va(i, j - 1) + va(i, j)) * g.cosa_v(i, j)) * g.dyc(i, j) * g.sina_v(i, j);
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = is2; i <= ie1; i++)
          vort(i, j) = (v(i, j) - 0.5 * (ua(i - 1, j) + ua(i, j)) * g.cosa_u(i, j)) * g.dxc(i, j) * g.sina_u(i, j);
    } else {
      for (int j = js; j <= je + 1; j++) {
        if (j == 1 || j == npy) {
          for (int i = is - 1; i <= ie + 1; i++) {
            if (vc(i, j) > 0) ptc(i, j) = u(i, j) * g.dyc(i, j) * g.sin_sg(i, j - 1, 4);
            else ptc(i, j) = u(i, j) * g.dyc(i, j) * g.sin_sg(i, j, 2);
          }
        } else {
          for (int i = is - 1; i <= ie + 1; i++)
            ptc(i, j) = (u(i, j) - 0.5 * (va(i, j - 1) + va(i, j)) * g.cosa_v(i, j)) * g.dyc(i, j) * g.sina_v(i, j);
        }
      }
      for (int j = js - 1; j <= je + 1; j++) {
        for (int i = is2; i <= ie1; i++)
          vort(i, j) = (v(i, j) - 0.5 * (ua(i - 1, j) + ua(i, j)) * g.cosa_u(i, j)) * g.dxc(i, j) * g.sina_u(i, j);
        if (is == 1) {
          if (uc(1, j) > 0) vort(1, j) = v(1, j) * g.dxc(1, j) * g.sin_sg(0, j, 3);
          else vort(1, j) = v(1, j) * g.dxc(1, j) * g.sin_sg(1, j, 1);
        }
        if ((ie + 1) == npx) {
          if (uc(npx, j) > 0) vort(npx, j) = v(npx, j) * g.dxc(npx, j) * g.sin_sg(npx - 1, j, 3);
          else vort(npx, j) = v(npx, j) * g.dxc(npx, j) * g.sin_sg(npx, j, 1);
        }
      }
    }
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) delpc(i, j) = vort(i, j - 1) - vort(i, j) + ptc(i - 1, j) - ptc(i, j);
    if (bd.sw_corner) delpc(1, 1) = delpc(1, 1) - vort(1, 0);
    if (bd.se_corner) delpc(npx, 1) = delpc(npx, 1) - vort(npx, 0);
    if (bd.ne_corner) delpc(npx, npy) = delpc(npx, npy) + vort(npx, npy);
    if (bd.nw_corner) delpc(1, npy) = delpc(1, npy) + vort(1, npy);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) {
        delpc(i, j) = g.rarea_c(i, j) * delpc(i, j);
        damp = g.da_min_c * std::max(a.d2_bg, std::min(0.20, a.dddmp * std::fabs(delpc(i, j) * dt)));
        vort(i, j) = damp * delpc(i, j);
        ke(i, j) = ke(i, j) + vort(i, j);
      }
  } else {
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) delpc(i, j) = divg_d(i, j);
    const int n2 = a.nord + 1;
    for (int n = 1; n <= a.nord; n++) {
      const int nt = a.nord - n;
      const bool fill_c = (nt != 0) && (grid_type < 3) &&
                          (bd.sw_corner || bd.se_corner || bd.ne_corner || bd.nw_corner) && !bounded;
      if (fill_c) fill_corners_bgrid(divg_d, npx, npy, ng, 1);
      for (int j = js - nt; j <= je + 1 + nt; j++)
        for (int i = is - 1 - nt; i <= ie + 1 + nt; i++) vc(i, j) = (divg_d(i + 1, j) - divg_d(i, j)) * g.divg_u(i, j);
      if (fill_c) fill_corners_bgrid(divg_d, npx, npy, ng, 2);
      for (int j = js - 1 - nt; j <= je + 1 + nt; j++)
        for (int i = is - nt; i <= ie + 1 + nt; i++) uc(i, j) = (divg_d(i, j + 1) - divg_d(i, j)) * g.divg_v(i, j);
      if (fill_c) fill_corners_dgrid_vec(vc, uc, npx, npy, ng, -1.0);
      for (int j = js - nt; j <= je + 1 + nt; j++)
        for (int i = is - nt; i <= ie + 1 + nt; i++) divg_d(i, j) = uc(i, j - 1) - uc(i, j) + vc(i - 1, j) - vc(i, j);
      if (bd.sw_corner) divg_d(1, 1) = divg_d(1, 1) - uc(1, 0);
      if (bd.se_corner) divg_d(npx, 1) = divg_d(npx, 1) - uc(npx, 0);
      if (bd.ne_corner) divg_d(npx, npy) = divg_d(npx, npy) + uc(npx, npy);
      if (bd.nw_corner) divg_d(1, npy) = divg_d(1, npy) + uc(1, npy);
      if (!bd.stretched_grid)
        for (int j = js - nt; j <= je + 1 + nt; j++)
          for (int i = is - nt; i <= ie + 1 + nt; i++) divg_d(i, j) = divg_d(i, j) * g.rarea_c(i, j);
    }
    if (a.dddmp < 1.E-5) {
      for (int j = jsd; j <= jed; j++) for (int i = isd; i <= ied; i++) vort(i, j) = 0.;
    } else {
      // grid_type<3 only (smag_corner for doubly-periodic is not restated)
      a2b_ord4(wk, vort, g, bd, false);
      for (int j = js; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++)
          vort(i, j) = std::fabs(dt) * std::sqrt(delpc(i, j) * delpc(i, j) + vort(i, j) * vort(i, j));
    }
    if (bd.stretched_grid) dd8 = g.da_min * std::pow(a.d4_bg, (double)n2);
    else dd8 = std::pow(g.da_min_c * a.d4_bg, (double)n2);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) {
        damp2 = g.da_min_c * std::max(a.d2_bg, std::min(0.20, a.dddmp * vort(i, j)));
        vort(i, j) = damp2 * delpc(i, j) + dd8 * divg_d(i, j);
        ke(i, j) = ke(i, j) + vort(i, j);
      }
  }

  if (a.d_con > 1.e-5 || a.do_diss_est) {
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie; i++) ub(i, j) = vort(i, j) - vort(i + 1, j);
    for (int j = js; j <= je; j++) for (int i = is; i <= ie + 1; i++) vb(i, j) = vort(i, j) - vort(i, j + 1);
  }

  // Vorticity transport
  if (!a.hydrostatic && a.do_f3d) {
    for (int j = jsd; j <= jed; j++) for (int i = isd; i <= ied; i++) vort(i, j) = wk(i, j) + g.f0(i, j) * z_rat(i, j);
  } else {
    for (int j = jsd; j <= jed; j++) for (int i = isd; i <= ied; i++) vort(i, j) = wk(i, j) + g.f0(i, j);
  }
  fv_tp_2d(vort, crx_adv, cry_adv, npx, npy, a.hord_vt, fx, fy, xfx_adv, yfx_adv, g, bd, ra_x, ra_y, a.lim_fac,
           nullptr, nullptr, nullptr, false, 0, 0.);
  for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie; i++) u(i, j) = vt(i, j) + ke(i, j) - ke(i + 1, j) + fy(i, j);
  for (int j = js; j <= je; j++) for (int i = is; i <= ie + 1; i++) v(i, j) = ut(i, j) + ke(i, j) - ke(i, j + 1) - fx(i, j);

  // damping applied to relative vorticity
  if (a.damp_v > 1.E-5) {
    damp4 = std::pow(a.damp_v * g.da_min_c, (double)(a.nord_v + 1));
    del6_vt_flux(a.nord_v, npx, npy, damp4, wk, vort, ut, vt, g, bd);
  } else if (a.do_diss_est) {
    ut.fill(0.);
    vt.fill(0.);
  }

  if (a.d_con > 1.e-5 || a.do_diss_est) {
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) {
        ub(i, j) = (ub(i, j) + vt(i, j)) * g.rdx(i, j);
        fy(i, j) = u(i, j) * g.rdx(i, j);
        gy(i, j) = fy(i, j) * ub(i, j);
      }
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) {
        vb(i, j) = (vb(i, j) - ut(i, j)) * g.rdy(i, j);
        fx(i, j) = v(i, j) * g.rdy(i, j);
        gx(i, j) = fx(i, j) * vb(i, j);
      }
    damp = 0.25 * a.d_con;
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++) {
        double u2 = fy(i, j) + fy(i, j + 1);
        double du2 = ub(i, j) + ub(i, j + 1);
        double v2 = fx(i, j) + fx(i + 1, j);
        double dv2 = vb(i, j) + vb(i + 1, j);
        double tmp = g.rsin2(i, j) * ((ub(i, j) * ub(i, j) + ub(i, j + 1) * ub(i, j + 1) + vb(i, j) * vb(i, j) + vb(i + 1, j) * vb(i + 1, j)) +
                                      2. * (gy(i, j) + gy(i, j + 1) + gx(i, j) + gx(i + 1, j)) -
                                      g.cosa_s(i, j) * (u2 * dv2 + v2 * du2 + du2 * dv2));
        if (a.prevent_diss_cooling) {
          if (a.d_con > 1.e-5) heat_source(i, j) = delp(i, j) * (heat_source(i, j) - damp * std::min(0., tmp));
          if (a.do_diss_est) diss_est(i, j) = diss_est(i, j) - tmp;
        } else {
          // sw_core.F90:1573-1576: damp*rsin2*(...) -- same factors, association (damp*rsin2)*(...)
          double inner = ((ub(i, j) * ub(i, j) + ub(i, j + 1) * ub(i, j + 1) + vb(i, j) * vb(i, j) + vb(i + 1, j) * vb(i + 1, j)) +
                          2. * (gy(i, j) + gy(i, j + 1) + gx(i, j) + gx(i + 1, j)) -
                          g.cosa_s(i, j) * (u2 * dv2 + v2 * du2 + du2 * dv2));
          heat_source(i, j) = delp(i, j) * (heat_source(i, j) - damp * g.rsin2(i, j) * inner);
          if (a.do_diss_est) diss_est(i, j) = diss_est(i, j) - g.rsin2(i, j) * inner;
        }
      }
  }
  if (a.damp_v > 1.E-5) {
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie; i++) u(i, j) = u(i, j) + vt(i, j);
    for (int j = js; j <= je; j++) for (int i = is; i <= ie + 1; i++) v(i, j) = v(i, j) - ut(i, j);
  }
}

}  // namespace fv3o
