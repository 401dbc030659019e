// 1-D PPM reconstruction + upwind flux device functions (sm_100a).
//
// What they compute: model/tp_core.F90 xppm (:324-712) / yppm (:715-1152) and
// model/sw_core.F90 xtp_u (:2154-2521) / ytp_v (:2524-2998) of the reference.
// How: the reference sweeps whole rows building al/bl/br arrays; here every thread owns ONE
// flux interface, picks the UPWIND cell from the sign of the Courant number first and
// reconstructs only that cell's (bl, br) from a 5-7 point register window -- no row
// temporaries, no redundant reconstruction for the monotone family.  The x and y operators
// share one implementation through a strided accessor.  Cube-edge one-sided formulas
// (tp_core.F90:643-681, sw_core.F90:2446-2490) are taken only by the threads whose upwind
// cell is one of {0,1,2,n-2,n-1,n}.
//
// Supported schemes: 5, 6, -5 (unlimited family) and 8, 10 (monotone family); the host
// rejects the others (no silent fallback).
#pragma once
#include <cstring>
#include "fv3_ctx.hpp"

// The pure-arithmetic routines below are also compiled for the host (PPM_HD) so that tests/host_ppm_test.cu can check the very
// expressions the kernels run against the oracle on a machine without a GPU (tests/test_host_device_math.py).
#define PPM_HD __host__ __device__
namespace ppm {
PPM_HD __forceinline__ int hi_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  long long b; memcpy(&b, &x, sizeof b); return (int)(b >> 32);
#endif
}

PPM_HD __forceinline__ double fsign(double a, double b) { return copysign(fabs(a), b); }
// compare-and-select min/max: 3 instructions (DSETP + 2 SEL) where fmin/fmax cost 6-7 on sm_100a because of their
// NaN-quieting fix-up (seen in the SASS); the path never sees NaNs and +-0 order is irrelevant to the limiters
PPM_HD __forceinline__ double mn(double a, double b) { return a < b ? a : b; }
PPM_HD __forceinline__ double mx(double a, double b) { return a > b ? a : b; }
PPM_HD __forceinline__ double min3(double a, double b, double c) { return mn(mn(a, b), c); }
PPM_HD __forceinline__ double max3(double a, double b, double c) { return mx(mx(a, b), c); }

// tp_core.F90:35-70
constexpr double r3 = 1. / 3.;
constexpr double r12 = 1. / 12.;
constexpr double s11 = 11. / 14., s14 = 4. / 7., s15 = 3. / 14.;
constexpr double c1 = -2. / 14., c2 = 11. / 14., c3 = 5. / 14.;
constexpr double p1 = 7. / 12., p2 = -1. / 12.;
constexpr double near_zero_tp = 1.E-25;   // tp_core.F90:37
constexpr double near_zero_sw = 1.E-9;    // sw_core.F90:39

// strided 1-D view: value at sweep index s is p[base + s*stride]
struct Acc {
  const double* p; long long base; long long stride;
  PPM_HD __forceinline__ double operator()(int s) const {
#ifdef __CUDA_ARCH__
    return __ldg(p + base + (long long)s * stride);
#else
    return p[base + (long long)s * stride];
#endif
  }
};

// scalar field with the copy_corners(dir=1) view, x sweep along row j (tp_core.F90:257-286)
struct QAccX {
  const double* q; Lay L; int j;
  __device__ __forceinline__ double operator()(int i) const {
    int ii = i, jj = j;
    if (L.cube) {
      if (j <= 0) {
        if (i <= 0) { ii = j; jj = 1 - i; }
        else if (i >= L.npx) { ii = L.npy - j; jj = i - L.npx + 1; }
      } else if (j >= L.npy) {
        if (i >= L.npx) { ii = j; jj = 2 * L.npx - 1 - i; }
        else if (i <= 0) { ii = L.npy - j; jj = i - 1 + L.npx; }
      }
    }
    return __ldg(q + LIDX(L, ii, jj));
  }
};
// copy_corners(dir=2) view, y sweep along column i (tp_core.F90:288-318)
struct QAccY {
  const double* q; Lay L; int i;
  __device__ __forceinline__ double operator()(int j) const {
    int ii = i, jj = j;
    if (L.cube) {
      if (i <= 0) {
        if (j <= 0) { ii = 1 - j; jj = i; }
        else if (j >= L.npy) { ii = j + 1 - L.npx; jj = L.npy - i; }
      } else if (i >= L.npx) {
        if (j <= 0) { ii = L.npy + j - 1; jj = L.npx - i; }
        else if (j >= L.npy) { ii = 2 * L.npy - 1 - j; jj = i; }
      }
    }
    return __ldg(q + LIDX(L, ii, jj));
  }
};

template <class Q>
PPM_HD __forceinline__ double dm_at(const Q& q, int i) {  // tp_core.F90:570-574
  const double qm = q(i - 1), q0 = q(i), qp = q(i + 1);
  const double xt = 0.25 * (qp - qm);
  return fsign(mn(mn(fabs(xt), max3(qm, q0, qp) - q0), q0 - min3(qm, q0, qp)), xt);
}

PPM_HD __forceinline__ void pert_std(double& al, double& ar) {  // pert_ppm iv/=0, tp_core.F90:1245-1261
  if (al * ar < 0.) {
    const double da1 = al - ar, da2 = da1 * da1, a6da = 3. * (al + ar) * da1;
    if (a6da < -da2) ar = -2. * al;
    else if (a6da > da2) al = -2. * ar;
  } else { al = 0.; ar = 0.; }
}

// pert_ppm with iv == 0 for one cell (positive-definite constraint of iord 9 / 13, tp_core.F90:1222-1244)
PPM_HD __forceinline__ void pert_pd(double a0, double& al, double& ar) {
  if (a0 <= 0.) { al = 0.; ar = 0.; return; }
  const double a4 = -3. * (ar + al), da1 = ar - al;
  if (fabs(da1) < -a4) {
    const double fmin_ = a0 + 0.25 / a4 * (da1 * da1) + a4 * r12;
    if (fmin_ < 0.) {
      if (ar > 0. && al > 0.) { ar = 0.; al = 0.; }
      else if (da1 > 0.) ar = -2. * al;
      else al = -2. * ar;
    }
  }
}
// Out of line: the rarely used schemes must not grow the hot 8 / 10 / 5 / 6 loop bodies (instruction cache, tp_tile.cuh).
// (bl, br) of an ordinary interior cell in the dm family for the schemes beyond 8 and 10: 11 (van Leer emulation, ppm_fac = 1.5,
// tp_core.F90:598-604), 12 (Lin & Rood positive definite, :605-627), 9 / 13 (unconstrained + pert_ppm(iv=0), :628-635)
static PPM_HD __noinline__ void mono_blbr_other(double q0, double al0, double al1, double dm0, int iord, double& bl, double& br) {
  if (iord == 11) {
    const double xt = 1.5 * dm0;
    bl = -fsign(mn(fabs(xt), fabs(al0 - q0)), xt);
    br = fsign(mn(fabs(xt), fabs(al1 - q0)), xt);
  } else if (iord == 12) {
    bl = al0 - q0; br = al1 - q0;
    const double a4 = -3. * (bl + br), da1 = br - bl;
    const bool ext5 = br * bl > 0., ext6 = fabs(da1) < -a4;
    if (ext6) {
      if (q0 + 0.25 / a4 * (da1 * da1) + a4 * r12 < 0.) {
        if (ext5) { br = 0.; bl = 0.; }
        else if (da1 > 0.) br = -2. * bl;
        else bl = -2. * br;
      }
    }
  } else {   // 9, 13
    bl = al0 - q0; br = al1 - q0;
    pert_pd(q0, bl, br);
  }
}
// al family, mord = |iord| in 1..4 (tp_core.F90:401-486): flux through the face between cells A (value qa) and B (qb) from
// al at the low face of A (alm), the shared face (al0) and the high face of B (alp).  lim_fac == 1 (checked on the host).
static PPM_HD __noinline__ double flux_al_low(double qa, double qb, double alm, double al0, double alp, double c, int mord) {
  if (mord == 2) {
    if (c > 0.) return qa + (1. - c) * (al0 - qa - c * (alm + al0 - (qa + qa)));
    return qb + (1. + c) * (al0 - qb + c * (al0 + alp - (qb + qb)));
  }
  const double Abl = alm - qa, Abr = al0 - qa, Ab0 = Abl + Abr;
  const double Bbl = al0 - qb, Bbr = alp - qb, Bb0 = Bbl + Bbr;
  const double Ax0 = fabs(Ab0), Axt = fabs(Abl - Abr), Bx0 = fabs(Bb0), Bxt = fabs(Bbl - Bbr);
  if (mord == 1) {
    const bool As = Ax0 < Axt, Bs = Bx0 < Bxt;
    double fx1, fl;
    if (c > 0.) { fx1 = (1. - c) * (Abr - c * Ab0); fl = qa; }
    else { fx1 = (1. + c) * (Bbl + c * Bb0); fl = qb; }
    if (As || Bs) fl = fl + fx1;
    return fl;
  }
  const bool A5 = Ax0 < Axt, A6 = 3. * Ax0 < Axt, B5 = Bx0 < Bxt, B6 = 3. * Bx0 < Bxt;
  if (mord == 3) {
    if (c > 0.) return (A5 || B6) ? qa + (1. - c) * (Abr - c * Ab0) : qa;
    return (A6 || B5) ? qb + (1. + c) * (Bbl + c * Bb0) : qb;
  }
  // mord == 4
  const bool hi5 = (A5 && B5) || (A6 || B6);
  double fx1, fl;
  if (c > 0.) { fx1 = (1. - c) * (Abr - c * Ab0); fl = qa; }
  else { fx1 = (1. + c) * (Bbl + c * Bb0); fl = qb; }
  if (hi5) fl = fl + fx1;
  return fl;
}

// iord == 7 (tp_core.F90:683-695): the flux tests smt5 = bl*br < 0 of BOTH cells of the face, so it needs the (positive-definite
// limited, :605-627) bl, br of cell A (low side, value qa) and of cell B (qb)
PPM_HD __forceinline__ double flux_pd7_from_cells(double qa, double qb, double Abl, double Abr, double Bbl, double Bbr, double c) {
  const double Ab0 = Abl + Abr, Bb0 = Bbl + Bbr;
  const bool As = Abl * Abr < 0., Bs = Bbl * Bbr < 0.;
  double fx1, fl;
  if (c > 0.) { fx1 = (1. - c) * (Abr - c * Ab0); fl = qa; }
  else { fx1 = (1. + c) * (Bbl + c * Bb0); fl = qb; }
  if (As || Bs) fl = fl + fx1;
  return fl;
}

// dxa-weighted two-sided edge value (tp_core.F90:376-377 / :647-648); e = first cell inside
// the face on the high side of the edge (e = 1 for the west/south edge, e = n for east/north)
template <class Q, class D>
PPM_HD __forceinline__ double edge_avg(const Q& q, const D& d, int e) {
  return 0.5 * (((2. * d(e - 1) + d(e - 2)) * q(e - 1) - d(e - 1) * q(e - 2)) / (d(e - 2) + d(e - 1)) +
                ((2. * d(e) + d(e + 1)) * q(e) - d(e) * q(e + 1)) / (d(e) + d(e + 1)));
}

// ------------------------------------------------------------------ scalar transport (xppm/yppm)
// monotone family: (bl, br) of cell i.  n = npx (or npy).  tp_core.F90:563-681
// RARE = false compiles the schemes beyond 8 / 10 (dm family) and 5 / 6 / -5 (al family) out: the hot instantiations of the
// transport kernels must not carry their code or their out-of-line calls (measured: +4.6 % on d_sw when they did)
template <bool RARE = true, class Q, class D>
PPM_HD __forceinline__ void cell_mono(const Q& q, const D& dxa, int i, int iord, int n, bool cube, double& bl, double& br) {
  if (!cube || (i >= 3 && i <= n - 3)) {
    const double qm1 = q(i - 1), q0 = q(i), qp1 = q(i + 1);
    const double dmm = dm_at(q, i - 1), dm0 = dm_at(q, i), dmp = dm_at(q, i + 1);
    const double al0 = 0.5 * (qm1 + q0) + r3 * (dmm - dm0);
    const double al1 = 0.5 * (q0 + qp1) + r3 * (dm0 - dmp);
    if (iord == 8) {
      const double xt = 2. * dm0;
      bl = -fsign(mn(fabs(xt), fabs(al0 - q0)), xt);
      br = fsign(mn(fabs(xt), fabs(al1 - q0)), xt);
    } else if (iord == 10 || !RARE) {
      bl = al0 - q0; br = al1 - q0;
      if (fabs(dmm) + fabs(dm0) + fabs(dmp) < near_zero_tp) { bl = 0.; br = 0.; }
      else if (fabs(3. * (bl + br)) > fabs(bl - br)) {
        const double dqm2 = 2. * (qm1 - q(i - 2)), dqm1 = 2. * (q0 - qm1), dq0 = 2. * (qp1 - q0), dqp1 = 2. * (q(i + 2) - qp1);
        const double pmp_2 = dqm1, lac_2 = pmp_2 - 0.75 * dqm2;
        br = mn(max3(0., pmp_2, lac_2), mx(br, min3(0., pmp_2, lac_2)));
        const double pmp_1 = -dq0, lac_1 = pmp_1 + 0.75 * dqp1;
        bl = mn(max3(0., pmp_1, lac_1), mx(bl, min3(0., pmp_1, lac_1)));
      }
    } else if (RARE) mono_blbr_other(q0, al0, al1, dm0, iord, bl, br);
    return;
  }
  // cube-edge cells
  if (i <= 1) {  // cells 0 and 1 share the clamped two-sided edge value
    double xt = edge_avg(q, dxa, 1);
    const double qa = q(-1), qb = q(0), qc = q(1), qd = q(2);
    xt = mx(xt, mn(mn(qa, qb), mn(qc, qd)));
    xt = mn(xt, mx(mx(qa, qb), mx(qc, qd)));
    if (i == 0) { bl = s14 * dm_at(q, -1) + s11 * (qa - qb); br = xt - qb; }
    else { bl = xt - qc; br = (s15 * qc + s11 * qd - s14 * dm_at(q, 2)) - qc; }
  } else if (i == 2) {
    const double q1 = q(1), q2 = q(2), q3 = q(3);
    const double dm2 = dm_at(q, 2);
    bl = (s15 * q1 + s11 * q2 - s14 * dm2) - q2;
    br = (0.5 * (q2 + q3) + r3 * (dm2 - dm_at(q, 3))) - q2;
  } else if (i == n - 2) {
    const double qa = q(n - 3), qb = q(n - 2), qc = q(n - 1);
    const double dmb = dm_at(q, n - 2);
    bl = (0.5 * (qa + qb) + r3 * (dm_at(q, n - 3) - dmb)) - qb;
    br = (s15 * qc + s11 * qb + s14 * dmb) - qb;
  } else {  // n-1, n
    double xt = edge_avg(q, dxa, n);
    const double qa = q(n - 2), qb = q(n - 1), qc = q(n), qd = q(n + 1);
    xt = mx(xt, mn(mn(qa, qb), mn(qc, qd)));
    xt = mn(xt, mx(mx(qa, qb), mx(qc, qd)));
    if (i == n - 1) { bl = (s15 * qb + s11 * qa + s14 * dm_at(q, n - 2)) - qb; br = xt - qb; }
    else { bl = xt - qc; br = s11 * (qd - qc) - s14 * dm_at(q, n + 1); }
  }
  pert_std(bl, br);   // pert_ppm(3, ..., 1), tp_core.F90:660,679
}

// unlimited family: edge value al(i).  tp_core.F90:369-392
template <class Q, class D>
PPM_HD __forceinline__ double al_unlim(const Q& q, const D& dxa, int i, int iord, int n, bool cube) {
  double al;
  if (!cube || (i >= 3 && i <= n - 2)) al = p1 * (q(i - 1) + q(i)) + p2 * (q(i - 2) + q(i + 1));
  else if (i == 0) al = c1 * q(-2) + c2 * q(-1) + c3 * q(0);
  else if (i == 1) al = edge_avg(q, dxa, 1);
  else if (i == 2) al = c3 * q(1) + c2 * q(2) + c1 * q(3);
  else if (i == n - 1) al = c1 * q(n - 3) + c2 * q(n - 2) + c3 * q(n - 1);
  else if (i == n) al = edge_avg(q, dxa, n);
  else al = c3 * q(n) + c2 * q(n + 1) + c1 * q(n + 2);  // n+1
  if (iord < 0) al = mx(0., al);
  return al;
}

struct CellU { double bl, br, b0; bool smt; };

// unlimited family cell (5, 6, -5).  tp_core.F90:491-546
template <class Q, class D>
PPM_HD __forceinline__ CellU cell_unlim(const Q& q, const D& dxa, int i, int iord, int n, bool cube) {
  CellU c;
  const double q0 = q(i);
  c.bl = al_unlim(q, dxa, i, iord, n, cube) - q0;
  c.br = al_unlim(q, dxa, i + 1, iord, n, cube) - q0;
  c.b0 = c.bl + c.br;
  if (iord == 5) { c.smt = c.bl * c.br < 0.; return c; }
  if (iord == -5) {
    c.smt = c.bl * c.br < 0.;
    const double da1 = c.br - c.bl, a4 = -3. * c.b0;
    if (fabs(da1) < -a4) {
      if (q0 + 0.25 / a4 * (da1 * da1) + a4 * r12 < 0.) {
        if (!c.smt) { c.br = 0.; c.bl = 0.; c.b0 = 0.; }
        else if (da1 > 0.) { c.br = -2. * c.bl; c.b0 = -c.bl; }
        else { c.bl = -2. * c.br; c.b0 = -c.br; }
      }
    }
  } else {
    c.smt = 3. * fabs(c.b0) < fabs(c.bl - c.br);
  }
  if (cube && (i == 0 || i == 1 || i == n - 1 || i == n)) c.smt = c.bl * c.br < 0.;   // tp_core.F90:536-545
  return c;
}

// flux through interface i for Courant number c.  tp_core.F90:549-558, :701-707
template <bool RARE = true, class Q, class D>
PPM_HD __forceinline__ double flux_scalar(const Q& q, const D& dxa, int i, double c, int iord, int n, bool cube) {
  if (RARE && iord == 7) {   // the limiter of 7 is the one of 12 (:605); cube-edge cells get the same one-sided overrides
    double Abl, Abr, Bbl, Bbr;
    cell_mono<true>(q, dxa, i - 1, 12, n, cube, Abl, Abr);
    cell_mono<true>(q, dxa, i, 12, n, cube, Bbl, Bbr);
    return flux_pd7_from_cells(q(i - 1), q(i), Abl, Abr, Bbl, Bbr, c);
  }
  if (iord >= 8) {
    const int iu = (c > 0.) ? i - 1 : i;
    double bl, br;
    cell_mono<RARE>(q, dxa, iu, iord, n, cube, bl, br);
    const double qu = q(iu);
    return (c > 0.) ? qu + (1. - c) * (br - c * (bl + br)) : qu + (1. + c) * (bl + c * (bl + br));
  }
  if (RARE && iord >= 1 && iord <= 4)   // no smt override at the cube-edge cells for these (tp_core.F90:536-545 is inside the 5 / -5 / 6 branch)
    return flux_al_low(q(i - 1), q(i), al_unlim(q, dxa, i - 1, iord, n, cube), al_unlim(q, dxa, i, iord, n, cube),
                       al_unlim(q, dxa, i + 1, iord, n, cube), c, iord);
  const CellU a = cell_unlim(q, dxa, i - 1, iord, n, cube);
  const CellU b = cell_unlim(q, dxa, i, iord, n, cube);
  double fx1, fl;
  if (c > 0.) { fx1 = (1. - c) * (a.br - c * a.b0); fl = q(i - 1); }
  else { fx1 = (1. + c) * (b.bl + c * b.b0); fl = q(i); }
  if (a.smt || b.smt) fl = fl + fx1;
  return fl;
}

// winds, al family, iord 1..4 (sw_core.F90:2246-2337): face between cells A (value ua, edge-aware bl/br) and B; c is a distance,
// rm / r0 = rdx of A / B.  lim_fac == 1 (host check).  Out of line: general instantiation of k_dsw_ke only.
static PPM_HD __noinline__ double flux_wind_low(double ua, double ub, double Abl, double Abr, double Bbl, double Bbr, double c,
                                                double rm, double r0, int iord) {
  const double Ab0 = Abl + Abr, Bb0 = Bbl + Bbr;
  if (iord == 2) {
    if (c > 0.) { const double cfl = c * rm; return ua + (1. - cfl) * (Abr - cfl * Ab0); }
    const double cfl = c * r0; return ub + (1. + cfl) * (Bbl + cfl * Bb0);
  }
  const double Ax0 = fabs(Ab0), Ax1 = fabs(Abl - Abr), Bx0 = fabs(Bb0), Bx1 = fabs(Bbl - Bbr);
  const bool A5 = Ax0 < Ax1, B5 = Bx0 < Bx1;
  if (iord == 1) {
    double fx0, fl;
    if (c > 0.) { const double cfl = c * rm; fx0 = (1. - cfl) * (Abr - cfl * Ab0); fl = ua; }
    else { const double cfl = c * r0; fx0 = (1. + cfl) * (Bbl + cfl * Bb0); fl = ub; }
    if (A5 || B5) fl = fl + fx0;
    return fl;
  }
  const bool A6 = 3. * Ax0 < Ax1, B6 = 3. * Bx0 < Bx1;
  const bool hi5 = A5 && B5, hi6 = A6 || B6;
  if (iord == 3) {
    double fx0 = 0.;
    if (c > 0.) {
      const double cfl = c * rm;
      if (hi6) fx0 = Abr - cfl * Ab0;
      else if (hi5) fx0 = fsign(mn(fabs(Abl), fabs(Abr)), Abr);
      return ua + (1. - cfl) * fx0;
    }
    const double cfl = c * r0;
    if (hi6) fx0 = Bbl + cfl * Bb0;
    else if (hi5) fx0 = fsign(mn(fabs(Bbl), fabs(Bbr)), Bbl);
    return ub + (1. + cfl) * fx0;
  }
  // iord == 4
  double fx0, fl;
  if (c > 0.) { const double cfl = c * rm; fx0 = (1. - cfl) * (Abr - cfl * Ab0); fl = ua; }
  else { const double cfl = c * r0; fx0 = (1. + cfl) * (Bbl + cfl * Bb0); fl = ub; }
  if (hi5 || hi6) fl = fl + fx0;
  return fl;
}
// winds, dm family, iord 9 (pmp-limited, sw_core.F90:2403-2411) and 11 / else (unlimited, :2434-2439), ordinary interior cell
static PPM_HD __noinline__ void wind_blbr_other(double um2, double um1, double u0, double up1, double up2, double al0, double al1, int iord,
                                                double& bl, double& br) {
  if (iord == 9) {
    const double dq0 = up1 - u0, dqp1 = up2 - up1, dqm1 = u0 - um1, dqm2 = um1 - um2;
    const double pmp_1 = -2. * dq0, lac_1 = pmp_1 + 1.5 * dqp1;
    bl = mn(max3(0., pmp_1, lac_1), mx(al0 - u0, min3(0., pmp_1, lac_1)));
    const double pmp_2 = 2. * dqm1, lac_2 = pmp_2 - 1.5 * dqm2;
    br = mn(max3(0., pmp_2, lac_2), mx(al1 - u0, min3(0., pmp_2, lac_2)));
  } else { bl = al0 - u0; br = al1 - u0; }
}

// ------------------------------------------------------------------ momentum (xtp_u / ytp_v)
// zero = the row/column of this sweep is a face edge line (j==1||j==npy for xtp_u),
// where bl=br=0 at the two cells touching the face corner (sw_core.F90:2206-2210,2451-2455)
template <bool RARE = true, class Q, class D>
PPM_HD __forceinline__ void cell_wind_mono(const Q& u, const D& dx, int i, int iord, int n, bool cube, bool zero,
                                               double& bl, double& br) {
  if (!cube) {   // "Other grids" branch, sw_core.F90:2494-2505
    const double um1 = u(i - 1), u0 = u(i), up1 = u(i + 1);
    const double dmm = dm_at(u, i - 1), dm0 = dm_at(u, i), dmp = dm_at(u, i + 1);
    const double al0 = 0.5 * (um1 + u0) + r3 * (dmm - dm0), al1 = 0.5 * (u0 + up1) + r3 * (dm0 - dmp);
    double pmp = -2. * (up1 - u0), lac = pmp + 1.5 * (u(i + 2) - up1);
    bl = mn(max3(0., pmp, lac), mx(al0 - u0, min3(0., pmp, lac)));
    pmp = 2. * (u0 - um1); lac = pmp - 1.5 * (um1 - u(i - 2));
    br = mn(max3(0., pmp, lac), mx(al1 - u0, min3(0., pmp, lac)));
    return;
  }
  if (i >= 3 && i <= n - 3) {
    const double um1 = u(i - 1), u0 = u(i), up1 = u(i + 1);
    const double dmm = dm_at(u, i - 1), dm0 = dm_at(u, i), dmp = dm_at(u, i + 1);
    const double al0 = 0.5 * (um1 + u0) + r3 * (dmm - dm0), al1 = 0.5 * (u0 + up1) + r3 * (dm0 - dmp);
    if (iord == 8) {
      const double xt = 2. * dm0;
      bl = -fsign(mn(fabs(xt), fabs(al0 - u0)), xt);
      br = fsign(mn(fabs(xt), fabs(al1 - u0)), xt);
    } else if (RARE && iord != 10) {
      wind_blbr_other(u(i - 2), um1, u0, up1, u(i + 2), al0, al1, iord, bl, br);
    } else {  // 10, sw_core.F90:2414-2433
      bl = al0 - u0; br = al1 - u0;
      if (fabs(dm0) < near_zero_sw) {
        if (fabs(dmm) + fabs(dmp) < near_zero_sw) { bl = 0.; br = 0.; }
      } else if (fabs(3. * (bl + br)) > fabs(bl - br)) {
        const double dq0 = up1 - u0, dqp1 = u(i + 2) - up1, dqm1 = u0 - um1, dqm2 = um1 - u(i - 2);
        const double pmp_1 = -2. * dq0, lac_1 = pmp_1 + 1.5 * dqp1;
        bl = mn(max3(0., pmp_1, lac_1), mx(bl, min3(0., pmp_1, lac_1)));
        const double pmp_2 = 2. * dqm1, lac_2 = pmp_2 - 1.5 * dqm2;
        br = mn(max3(0., pmp_2, lac_2), mx(br, min3(0., pmp_2, lac_2)));
      }
    }
    return;
  }
  // edges, sw_core.F90:2446-2490
  if (i == 2) {
    const double u1 = u(1), u2 = u(2), u3 = u(3);
    const double dm2 = dm_at(u, 2);
    br = (0.5 * (u2 + u3) + r3 * (dm2 - dm_at(u, 3))) - u2;
    bl = (s15 * u1 + s11 * u2 - s14 * dm2) - u2;
    pert_std(bl, br);
  } else if (i == n - 2) {
    const double ua = u(n - 3), ub = u(n - 2), uc = u(n - 1);
    const double dmb = dm_at(u, n - 2);
    bl = (0.5 * (ua + ub) + r3 * (dm_at(u, n - 3) - dmb)) - ub;
    br = (s15 * uc + s11 * ub + s14 * dmb) - ub;
    pert_std(bl, br);
  } else if (zero) {
    bl = 0.; br = 0.;
  } else if (i <= 1) {
    const double x0L = 0.5 * ((2. * dx(0) + dx(-1)) * (u(0)) - dx(0) * (u(-1))) / (dx(0) + dx(-1));
    const double x0R = 0.5 * ((2. * dx(1) + dx(2)) * (u(1)) - dx(1) * (u(2))) / (dx(1) + dx(2));
    const double xt = x0L + x0R;
    if (i == 0) { bl = s14 * dm_at(u, -1) - s11 * (u(0) - u(-1)); br = xt - u(0); }
    else { bl = xt - u(1); br = (s15 * u(1) + s11 * u(2) - s14 * dm_at(u, 2)) - u(1); }
  } else {  // n-1, n
    const double x0L = 0.5 * ((2. * dx(n - 1) + dx(n - 2)) * (u(n - 1)) - dx(n - 1) * (u(n - 2))) / (dx(n - 1) + dx(n - 2));
    const double x0R = 0.5 * ((2. * dx(n) + dx(n + 1)) * (u(n)) - dx(n) * (u(n + 1))) / (dx(n) + dx(n + 1));
    const double xt = x0L + x0R;
    if (i == n - 1) { bl = (s15 * u(n - 1) + s11 * u(n - 2) + s14 * dm_at(u, n - 2)) - u(n - 1); br = xt - u(n - 1); }
    else { bl = xt - u(n); br = s11 * (u(n + 1) - u(n)) - s14 * dm_at(u, n + 1); }
  }
}

// iord < 8 family (5, 6[,7]) for the winds.  sw_core.F90:2187-2377
template <class Q, class D>
PPM_HD __forceinline__ CellU cell_wind_unlim(const Q& u, const D& dx, int i, int iord, int n, bool cube, bool zero) {
  CellU c;
  auto alg = [&](int m) { return p1 * (u(m - 1) + u(m)) + p2 * (u(m - 2) + u(m + 1)); };
  if (!cube || (i >= 3 && i <= n - 3)) {
    c.bl = alg(i) - u(i); c.br = alg(i + 1) - u(i);
  } else if (i == 2) {
    c.bl = (c3 * u(1) + c2 * u(2) + c1 * u(3)) - u(2);
    c.br = alg(3) - u(2);
  } else if (i == n - 2) {
    c.bl = alg(n - 2) - u(n - 2);
    c.br = (c1 * u(n - 3) + c2 * u(n - 2) + c3 * u(n - 1)) - u(n - 2);
  } else if (zero) {
    c.bl = 0.; c.br = 0.;
  } else if (i <= 1) {
    const double xt = 0.5 * (((2. * dx(0) + dx(-1)) * (u(0)) - dx(0) * u(-1)) / (dx(0) + dx(-1)) +
                             ((2. * dx(1) + dx(2)) * (u(1)) - dx(1) * u(2)) / (dx(1) + dx(2)));
    if (i == 0) { c.bl = c1 * u(-2) + c2 * u(-1) + c3 * u(0) - u(0); c.br = xt - u(0); }
    else { c.bl = xt - u(1); c.br = (c3 * u(1) + c2 * u(2) + c1 * u(3)) - u(1); }
  } else {
    const double xt = 0.5 * (((2. * dx(n - 1) + dx(n - 2)) * u(n - 1) - dx(n - 1) * u(n - 2)) / (dx(n - 1) + dx(n - 2)) +
                             ((2. * dx(n) + dx(n + 1)) * u(n) - dx(n) * u(n + 1)) / (dx(n) + dx(n + 1)));
    if (i == n - 1) { c.bl = (c1 * u(n - 3) + c2 * u(n - 2) + c3 * u(n - 1)) - u(n - 1); c.br = xt - u(n - 1); }
    else { c.bl = xt - u(n); c.br = c3 * u(n) + c2 * u(n + 1) + c1 * u(n + 2) - u(n); }
  }
  c.b0 = c.bl + c.br;
  if (iord == 5) c.smt = c.bl * c.br < 0.;
  else {
    c.smt = 3. * fabs(c.b0) < fabs(c.bl - c.br);
    if (cube && (i == 0 || i == 1 || i == n - 1 || i == n)) c.smt = c.bl * c.br < 0.;
  }
  return c;
}

// flux of the wind itself through interface i; c is a DISTANCE, cfl = c * rdx(upwind)
template <bool RARE = true, class Q, class D>
PPM_HD __forceinline__ double flux_wind(const Q& u, const D& dx, const D& rdx, int i, double c, int iord, int n,
                                            bool cube, bool zero) {
  if (iord >= 8) {
    const int iu = (c > 0.) ? i - 1 : i;
    double bl, br;
    cell_wind_mono<RARE>(u, dx, iu, iord, n, cube, zero, bl, br);
    const double cfl = c * rdx(iu);
    const double uu = u(iu);
    return (c > 0.) ? uu + (1. - cfl) * (br - cfl * (bl + br)) : uu + (1. + cfl) * (bl + cfl * (bl + br));
  }
  const CellU a = cell_wind_unlim(u, dx, i - 1, iord, n, cube, zero);
  const CellU b = cell_wind_unlim(u, dx, i, iord, n, cube, zero);
  if (RARE && iord >= 1 && iord <= 4) return flux_wind_low(u(i - 1), u(i), a.bl, a.br, b.bl, b.br, c, rdx(i - 1), rdx(i), iord);
  double fx0, fl;
  if (c > 0.) { const double cfl = c * rdx(i - 1); fx0 = (1. - cfl) * (a.br - cfl * a.b0); fl = u(i - 1); }
  else { const double cfl = c * rdx(i); fx0 = (1. + cfl) * (b.bl + cfl * b.b0); fl = u(i); }
  if (a.smt || b.smt) fl = fl + fx0;
  return fl;
}

// ------------------------------------------------------------------ interior fast paths
// For an interface i with 4 <= i <= n-3 every cell the flux can touch is an ordinary interior
// cell (no cube-edge formula, no corner remap): the whole operator is register arithmetic on
// the 6-point window q(i-3..i+2), loaded once with plain strided loads.  Operation order is
// identical to cell_mono / cell_unlim / cell_wind_* above (the parity tests compare both).
PPM_HD __forceinline__ double dm3(double qm, double q0, double qp) {
  const double xt = 0.25 * (qp - qm);
  return fsign(mn(mn(fabs(xt), max3(qm, q0, qp) - q0), q0 - min3(qm, q0, qp)), xt);
}

// the operator on an explicit window a0..a5 = q(i-3..i+2) (registers; the caller loads from global or shared memory)
__device__ __forceinline__ double flux_scalar_win(double a0, double a1, double a2, double a3, double a4, double a5, double c, int iord) {
  if (iord >= 8) {
    const bool up = c > 0.;
    const double qm2 = up ? a0 : a1, qm1 = up ? a1 : a2, q0 = up ? a2 : a3, qp1 = up ? a3 : a4, qp2 = up ? a4 : a5;
    const double dmm = dm3(qm2, qm1, q0), dm0 = dm3(qm1, q0, qp1), dmp = dm3(q0, qp1, qp2);
    const double al0 = 0.5 * (qm1 + q0) + r3 * (dmm - dm0);
    const double al1 = 0.5 * (q0 + qp1) + r3 * (dm0 - dmp);
    double bl, br;
    if (iord == 8) {
      const double xt = 2. * dm0;
      bl = -fsign(mn(fabs(xt), fabs(al0 - q0)), xt);
      br = fsign(mn(fabs(xt), fabs(al1 - q0)), xt);
    } else {
      bl = al0 - q0; br = al1 - q0;
      if (fabs(dmm) + fabs(dm0) + fabs(dmp) < near_zero_tp) { bl = 0.; br = 0.; }
      else if (fabs(3. * (bl + br)) > fabs(bl - br)) {
        const double dqm2 = 2. * (qm1 - qm2), dqm1 = 2. * (q0 - qm1), dq0 = 2. * (qp1 - q0), dqp1 = 2. * (qp2 - qp1);
        const double pmp_2 = dqm1, lac_2 = pmp_2 - 0.75 * dqm2;
        br = mn(max3(0., pmp_2, lac_2), mx(br, min3(0., pmp_2, lac_2)));
        const double pmp_1 = -dq0, lac_1 = pmp_1 + 0.75 * dqp1;
        bl = mn(max3(0., pmp_1, lac_1), mx(bl, min3(0., pmp_1, lac_1)));
      }
    }
    return up ? q0 + (1. - c) * (br - c * (bl + br)) : q0 + (1. + c) * (bl + c * (bl + br));
  }
  double alm = p1 * (a1 + a2) + p2 * (a0 + a3);
  double al0 = p1 * (a2 + a3) + p2 * (a1 + a4);
  double alp = p1 * (a3 + a4) + p2 * (a2 + a5);
  if (iord < 0) { alm = mx(0., alm); al0 = mx(0., al0); alp = mx(0., alp); }
  auto cell = [&](double q0, double l, double r) {
    CellU cu; cu.bl = l - q0; cu.br = r - q0; cu.b0 = cu.bl + cu.br;
    if (iord == 5) cu.smt = cu.bl * cu.br < 0.;
    else if (iord == -5) {
      cu.smt = cu.bl * cu.br < 0.;
      const double da1 = cu.br - cu.bl, a4_ = -3. * cu.b0;
      if (fabs(da1) < -a4_) {
        if (q0 + 0.25 / a4_ * (da1 * da1) + a4_ * r12 < 0.) {
          if (!cu.smt) { cu.br = 0.; cu.bl = 0.; cu.b0 = 0.; }
          else if (da1 > 0.) { cu.br = -2. * cu.bl; cu.b0 = -cu.bl; }
          else { cu.bl = -2. * cu.br; cu.b0 = -cu.br; }
        }
      }
    } else cu.smt = 3. * fabs(cu.b0) < fabs(cu.bl - cu.br);
    return cu;
  };
  const CellU A = cell(a2, alm, al0), B = cell(a3, al0, alp);
  double fx1, fl;
  if (c > 0.) { fx1 = (1. - c) * (A.br - c * A.b0); fl = a2; }
  else { fx1 = (1. + c) * (B.bl + c * B.b0); fl = a3; }
  if (A.smt || B.smt) fl = fl + fx1;
  return fl;
}

__device__ __forceinline__ double flux_scalar_fast(const double* __restrict__ p, int s, double c, int iord) {
  return flux_scalar_win(__ldg(p - 3 * s), __ldg(p - 2 * s), __ldg(p - s), __ldg(p), __ldg(p + s), __ldg(p + 2 * s), c, iord);
}

// winds: c is a distance; rm, r0 = rdx of cells i-1 and i
__device__ __forceinline__ double flux_wind_fast(const double* __restrict__ p, int s, double c, double rm, double r0, int iord) {
  const double a0 = __ldg(p - 3 * s), a1 = __ldg(p - 2 * s), a2 = __ldg(p - s), a3 = __ldg(p), a4 = __ldg(p + s), a5 = __ldg(p + 2 * s);
  if (iord >= 8) {
    const bool up = c > 0.;
    const double um2 = up ? a0 : a1, um1 = up ? a1 : a2, u0 = up ? a2 : a3, up1 = up ? a3 : a4, up2 = up ? a4 : a5;
    const double dmm = dm3(um2, um1, u0), dm0 = dm3(um1, u0, up1), dmp = dm3(u0, up1, up2);
    const double al0 = 0.5 * (um1 + u0) + r3 * (dmm - dm0), al1 = 0.5 * (u0 + up1) + r3 * (dm0 - dmp);
    double bl, br;
    if (iord == 8) {
      const double xt = 2. * dm0;
      bl = -fsign(mn(fabs(xt), fabs(al0 - u0)), xt);
      br = fsign(mn(fabs(xt), fabs(al1 - u0)), xt);
    } else {
      bl = al0 - u0; br = al1 - u0;
      if (fabs(dm0) < near_zero_sw) {
        if (fabs(dmm) + fabs(dmp) < near_zero_sw) { bl = 0.; br = 0.; }
      } else if (fabs(3. * (bl + br)) > fabs(bl - br)) {
        const double dq0 = up1 - u0, dqp1 = up2 - up1, dqm1 = u0 - um1, dqm2 = um1 - um2;
        const double pmp_1 = -2. * dq0, lac_1 = pmp_1 + 1.5 * dqp1;
        bl = mn(max3(0., pmp_1, lac_1), mx(bl, min3(0., pmp_1, lac_1)));
        const double pmp_2 = 2. * dqm1, lac_2 = pmp_2 - 1.5 * dqm2;
        br = mn(max3(0., pmp_2, lac_2), mx(br, min3(0., pmp_2, lac_2)));
      }
    }
    const double cfl = c * (up ? rm : r0);
    return up ? u0 + (1. - cfl) * (br - cfl * (bl + br)) : u0 + (1. + cfl) * (bl + cfl * (bl + br));
  }
  const double alm = p1 * (a1 + a2) + p2 * (a0 + a3);
  const double al0 = p1 * (a2 + a3) + p2 * (a1 + a4);
  const double alp = p1 * (a3 + a4) + p2 * (a2 + a5);
  const double Abl = alm - a2, Abr = al0 - a2, Ab0 = Abl + Abr;
  const double Bbl = al0 - a3, Bbr = alp - a3, Bb0 = Bbl + Bbr;
  bool As, Bs;
  if (iord == 5) { As = Abl * Abr < 0.; Bs = Bbl * Bbr < 0.; }
  else { As = 3. * fabs(Ab0) < fabs(Abl - Abr); Bs = 3. * fabs(Bb0) < fabs(Bbl - Bbr); }
  double fx0, fl;
  if (c > 0.) { const double cfl = c * rm; fx0 = (1. - cfl) * (Abr - cfl * Ab0); fl = a2; }
  else { const double cfl = c * r0; fx0 = (1. + cfl) * (Bbl + cfl * Bb0); fl = a3; }
  if (As || Bs) fl = fl + fx0;
  return fl;
}

PPM_HD __forceinline__ double ldg_(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
// every scheme (general instantiation of k_dsw_ke; host-callable for tests/host_ppm_test.cu)
PPM_HD __forceinline__ double flux_wind_fast_g(const double* __restrict__ p, int s, double c, double rm, double r0, int iord) {
  const double a0 = ldg_(p - 3 * s), a1 = ldg_(p - 2 * s), a2 = ldg_(p - s), a3 = ldg_(p), a4 = ldg_(p + s), a5 = ldg_(p + 2 * s);
  if (iord >= 8) {
    const bool up = c > 0.;
    const double um2 = up ? a0 : a1, um1 = up ? a1 : a2, u0 = up ? a2 : a3, up1 = up ? a3 : a4, up2 = up ? a4 : a5;
    const double dmm = dm3(um2, um1, u0), dm0 = dm3(um1, u0, up1), dmp = dm3(u0, up1, up2);
    const double al0 = 0.5 * (um1 + u0) + r3 * (dmm - dm0), al1 = 0.5 * (u0 + up1) + r3 * (dm0 - dmp);
    double bl, br;
    if (iord == 8) {
      const double xt = 2. * dm0;
      bl = -fsign(mn(fabs(xt), fabs(al0 - u0)), xt);
      br = fsign(mn(fabs(xt), fabs(al1 - u0)), xt);
    } else if (iord != 10) {
      wind_blbr_other(um2, um1, u0, up1, up2, al0, al1, iord, bl, br);
    } else {
      bl = al0 - u0; br = al1 - u0;
      if (fabs(dm0) < near_zero_sw) {
        if (fabs(dmm) + fabs(dmp) < near_zero_sw) { bl = 0.; br = 0.; }
      } else if (fabs(3. * (bl + br)) > fabs(bl - br)) {
        const double dq0 = up1 - u0, dqp1 = up2 - up1, dqm1 = u0 - um1, dqm2 = um1 - um2;
        const double pmp_1 = -2. * dq0, lac_1 = pmp_1 + 1.5 * dqp1;
        bl = mn(max3(0., pmp_1, lac_1), mx(bl, min3(0., pmp_1, lac_1)));
        const double pmp_2 = 2. * dqm1, lac_2 = pmp_2 - 1.5 * dqm2;
        br = mn(max3(0., pmp_2, lac_2), mx(br, min3(0., pmp_2, lac_2)));
      }
    }
    const double cfl = c * (up ? rm : r0);
    return up ? u0 + (1. - cfl) * (br - cfl * (bl + br)) : u0 + (1. + cfl) * (bl + cfl * (bl + br));
  }
  const double alm = p1 * (a1 + a2) + p2 * (a0 + a3);
  const double al0 = p1 * (a2 + a3) + p2 * (a1 + a4);
  const double alp = p1 * (a3 + a4) + p2 * (a2 + a5);
  const double Abl = alm - a2, Abr = al0 - a2, Ab0 = Abl + Abr;
  const double Bbl = al0 - a3, Bbr = alp - a3, Bb0 = Bbl + Bbr;
  if (iord >= 1 && iord <= 4) return flux_wind_low(a2, a3, Abl, Abr, Bbl, Bbr, c, rm, r0, iord);
  bool As, Bs;
  if (iord == 5) { As = Abl * Abr < 0.; Bs = Bbl * Bbr < 0.; }
  else { As = 3. * fabs(Ab0) < fabs(Abl - Abr); Bs = 3. * fabs(Bb0) < fabs(Bbl - Bbr); }
  double fx0, fl;
  if (c > 0.) { const double cfl = c * rm; fx0 = (1. - cfl) * (Abr - cfl * Ab0); fl = a2; }
  else { const double cfl = c * r0; fx0 = (1. + cfl) * (Bbl + cfl * Bb0); fl = a3; }
  if (As || Bs) fl = fl + fx0;
  return fl;
}

// ------------------------------------------------------------------ two-pass form for shared-memory tiles
// Pass 1 stores one auxiliary value per point of a line, pass 2 evaluates the flux from it:
//   monotone family (8, 10):  aux(c) = dm of cell c                      (tp_core.F90:570-574)
//   unlimited family (5,6,-5): aux(c) = al at the low-side face of cell c (tp_core.F90:369-373)
// so the limiter / edge value is computed once per point instead of three times per flux.
// dm2 == dm3 up to the sign of a zero result: max3-q0 and q0-min3 are |q0-qm| and |qp-q0| when the
// two one-sided differences have the same sign, and one of them is 0 otherwise.
PPM_HD __forceinline__ double dm2(double qm, double q0, double qp) {
  const double a = q0 - qm, b = qp - q0, xt = 0.25 * (qp - qm);
  const bool same = ((hi_word(a) ^ hi_word(b)) >= 0);
  const double m = mn(mn(fabs(xt), fabs(a)), fabs(b));
  return same ? copysign(m, xt) : 0.;
}
PPM_HD __forceinline__ double aux_point(bool mono, int iord, double qm2, double qm1, double q0, double qp1) {
  if (mono) return dm2(qm1, q0, qp1);
  const double al = p1 * (qm1 + q0) + p2 * (qm2 + qp1);
  return iord < 0 ? mx(0., al) : al;
}
// iord == 7 on an interior line: a0..a5 = q(i-3..i+2) around the face between cells a2 and a3 (out of line: general instantiation only)
static PPM_HD __noinline__ double flux_pd7_line(double a0, double a1, double a2, double a3, double a4, double a5, double c) {
  const double d1 = dm2(a0, a1, a2), d2 = dm2(a1, a2, a3), d3 = dm2(a2, a3, a4), d4 = dm2(a3, a4, a5);
  const double alA = 0.5 * (a1 + a2) + r3 * (d1 - d2), alS = 0.5 * (a2 + a3) + r3 * (d2 - d3), alB = 0.5 * (a3 + a4) + r3 * (d3 - d4);
  double Abl, Abr, Bbl, Bbr;
  mono_blbr_other(a2, alA, alS, d2, 12, Abl, Abr);
  mono_blbr_other(a3, alS, alB, d3, 12, Bbl, Bbr);
  return flux_pd7_from_cells(a2, a3, Abl, Abr, Bbl, Bbr, c);
}
// monotone flux from the upwind cell's neighbourhood: q(iu-2..iu+2), dm(iu-1..iu+1)
template <bool RARE = true>
PPM_HD __forceinline__ double flux_mono_aux(double qm2, double qm1, double q0, double qp1, double qp2, double dmm, double dm0,
                                                double dmp, double c, int iord) {
  const double al0 = 0.5 * (qm1 + q0) + r3 * (dmm - dm0);
  const double al1 = 0.5 * (q0 + qp1) + r3 * (dm0 - dmp);
  double bl, br;
  if (iord == 8) {
    const double xt = 2. * dm0;
    bl = -fsign(mn(fabs(xt), fabs(al0 - q0)), xt);
    br = fsign(mn(fabs(xt), fabs(al1 - q0)), xt);
  } else if (iord == 10 || !RARE) {
    bl = al0 - q0; br = al1 - q0;
    if (fabs(dmm) + fabs(dm0) + fabs(dmp) < near_zero_tp) { bl = 0.; br = 0.; }
    else if (fabs(3. * (bl + br)) > fabs(bl - br)) {
      const double dqm2 = 2. * (qm1 - qm2), dqm1 = 2. * (q0 - qm1), dq0 = 2. * (qp1 - q0), dqp1 = 2. * (qp2 - qp1);
      const double pmp_2 = dqm1, lac_2 = pmp_2 - 0.75 * dqm2;
      br = mn(max3(0., pmp_2, lac_2), mx(br, min3(0., pmp_2, lac_2)));
      const double pmp_1 = -dq0, lac_1 = pmp_1 + 0.75 * dqp1;
      bl = mn(max3(0., pmp_1, lac_1), mx(bl, min3(0., pmp_1, lac_1)));
    }
  } else mono_blbr_other(q0, al0, al1, dm0, iord, bl, br);
  return (c > 0.) ? q0 + (1. - c) * (br - c * (bl + br)) : q0 + (1. + c) * (bl + c * (bl + br));
}
// unlimited-family flux through the face between cells A (low side, value qa) and B (qb); alm, al0, alp = al at the
// low face of A, the shared face, the high face of B
template <bool RARE = true>
PPM_HD __forceinline__ double flux_unlim_aux(double qa, double qb, double alm, double al0, double alp, double c, int iord) {
  if (RARE && iord >= 1 && iord <= 4) return flux_al_low(qa, qb, alm, al0, alp, c, iord);
  auto cell = [&](double q0, double l, double r) {
    CellU cu; cu.bl = l - q0; cu.br = r - q0; cu.b0 = cu.bl + cu.br;
    if (iord == 5) cu.smt = cu.bl * cu.br < 0.;
    else if (iord == -5) {
      cu.smt = cu.bl * cu.br < 0.;
      const double da1 = cu.br - cu.bl, a4_ = -3. * cu.b0;
      if (fabs(da1) < -a4_) {
        if (q0 + 0.25 / a4_ * (da1 * da1) + a4_ * r12 < 0.) {
          if (!cu.smt) { cu.br = 0.; cu.bl = 0.; cu.b0 = 0.; }
          else if (da1 > 0.) { cu.br = -2. * cu.bl; cu.b0 = -cu.bl; }
          else { cu.bl = -2. * cu.br; cu.b0 = -cu.br; }
        }
      }
    } else cu.smt = 3. * fabs(cu.b0) < fabs(cu.bl - cu.br);
    return cu;
  };
  const CellU A = cell(qa, alm, al0), B = cell(qb, al0, alp);
  double fx1, fl;
  if (c > 0.) { fx1 = (1. - c) * (A.br - c * A.b0); fl = qa; }
  else { fx1 = (1. + c) * (B.bl + c * B.b0); fl = qb; }
  if (A.smt || B.smt) fl = fl + fx1;
  return fl;
}

// every member of tp_valid_schemes (tp_core.F90:78); hord = 1 reads lim_fac (:404), only the default lim_fac = 1 is built
__host__ inline bool hord_supported(int h, double lim_fac = 1.0) {
  if (h == 1) return lim_fac == 1.0;
  return h == -5 || (h >= 2 && h <= 13);
}
// schemes outside the common five run in the general (FAM = 2) instantiations of the tile kernels
__host__ inline bool hord_is_rare(int h) { return !(h == 5 || h == 6 || h == -5 || h == 8 || h == 10); }
// xtp_u / ytp_v (sw_core.F90:2154-2998): 1..4, 5, 6 (= 7), 8, 9, 10, 11; hord_mt = 1 reads lim_fac (only 1 is built)
__host__ inline bool hord_wind_supported(int h, double lim_fac = 1.0) {
  if (h == 1) return lim_fac == 1.0;
  return h >= 2 && h <= 11;
}
__host__ inline bool hord_wind_is_rare(int h) { return !(h == 5 || h == 6 || h == 8 || h == 10); }

}  // namespace ppm
