// Column operators of the vertical remap (fv_operators.F90: scalar_profile / cs_profile :546-1300, cs_limiters :1303-1378,
// ppm_profile / ppm_limiters :1382-1723; fv_fill.F90: fillz :34-139; the
// mapping loop of map_scalar / map1_ppm / map1_q2 / mapn_tracer :88-132, 183-227, 276-336, 399-441) for ONE column whose
// reconstruction arrays are strided (level stride = plane).  __host__ __device__: remap.cu runs them one thread per column; the
// CPU suite runs the same source on the host against the oracle (tests/host_remap_test.cu, tests/test_host_remap.py).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define RMP_HD __host__ __device__
#else
#define RMP_HD
#endif

namespace rmp {

constexpr double r3 = 1. / 3., r23 = 2. / 3., r12 = 1. / 12.;

RMP_HD inline double dmin(double a, double b) { return a < b ? a : b; }
RMP_HD inline double dmax(double a, double b) { return a > b ? a : b; }
RMP_HD inline double dmin3(double a, double b, double c) { return dmin(dmin(a, b), c); }
RMP_HD inline double dmax3(double a, double b, double c) { return dmax(dmax(a, b), c); }

// the scratch of one column: pointers already offset to the column, level k (1-based) at [(k - 1) * plane]
struct Col {
  double *a1, *a2, *a3, *a4, *q, *gam;
  int* fl;   // bit 0 extm, bit 1 ext5, bit 2 ext6
  long long plane;
};
#define LV(p, k) (p)[(long long)((k) - 1) * C.plane]
#define A1(k) LV(C.a1, k)
#define A2(k) LV(C.a2, k)
#define A3(k) LV(C.a3, k)
#define A4(k) LV(C.a4, k)
#define QI(k) LV(C.q, k)
#define GAM(k) LV(C.gam, k)
#define FL(k) LV(C.fl, k)

// fv_operators.F90:1303-1378 for one layer, on registers
RMP_HD inline void cs_limiters(double a1, double& a2, double& a3, double& a4, bool extm, int iv) {
  if (iv == 0) {
    if (a1 <= 0.) { a2 = a1; a3 = a1; a4 = 0.; }
    else if (fabs(a3 - a2) < -a4) {
      if ((a1 + 0.25 * ((a3 - a2) * (a3 - a2)) / a4 + a4 * r12) < 0.) {
        if (a1 < a3 && a1 < a2) { a3 = a1; a2 = a1; a4 = 0.; }
        else if (a3 > a2) { a4 = 3. * (a2 - a1); a3 = a2 - a4; }
        else { a4 = 3. * (a3 - a1); a2 = a3 - a4; }
      }
    }
  } else {
    const bool flat = iv == 1 ? ((a1 - a2) * (a1 - a3) >= 0.) : extm;
    if (flat) { a2 = a1; a3 = a1; a4 = 0.; }
    else {
      const double da1 = a3 - a2, da2 = da1 * da1, a6da = a4 * da1;
      if (a6da < -da2) { a4 = 3. * (a2 - a1); a3 = a2 - a4; }
      else if (a6da > da2) { a4 = 3. * (a3 - a1); a2 = a3 - a4; }
    }
  }
}

// scalar_profile (scalar: with the q_min tests, :546-916) / cs_profile (:919-1300); A1 holds the layer means, P1(k) the source
// interface pressures (k = 1..km+1).  Every level access is an L2 round trip (the columns of a block do not fit L1), so each sweep
// keeps its loop-carried and neighbouring values in registers and touches a scratch element once; the operations and their
// order are those of the Fortran.
// RB levels are handled per batch: every sweep first issues the loads of RB levels (independent of the values the sweep carries from
// level to level), then does the arithmetic and the stores.  A thread otherwise has one or two loads in flight per level -- the
// sweeps are chains of L2 / DRAM round trips, bound by occupancy x memory-level parallelism (remap.cu header) -- and the compiler
// cannot hoist the loads itself (it cannot prove that the strided stores do not alias them).  No load of a batch reads an element
// an earlier iteration of the same batch stores (checked sweep by sweep below).
constexpr int RB = 4;

template <class P1>
RMP_HD inline void profile(const Col& C, int km, const P1& pe1, double qs, int iv, int ak, double qmin, bool scalar) {
  // ---- interface values: tridiagonal solve (:570-622 / :941-1013).  Loads: pe1, A1, then QI / GAM at other levels than stored.
  {
    double plast = pe1(3);
    double dpm, dpk;                       // delp(k - 1), delp(k)
    { const double p1 = pe1(1), p2 = pe1(2); dpm = p2 - p1; dpk = plast - p2; }
    double a_prev = A1(1), a_cur = A1(2);  // A1(k - 1), A1(k)
    if (iv == -2) {   // lower boundary condition q(km+1) = qs
      double g_cur = 0.5, q_prev = 1.5 * a_prev;
      GAM(2) = g_cur;
      QI(1) = q_prev;
      for (int k = 2; k <= km - 1;) {
        const int nb = km - k < RB ? km - k : RB;   // levels k .. k + nb - 1 <= km - 1
        double pn[RB], an[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) { pn[j] = pe1(k + j + 2); an[j] = A1(k + j + 1); }
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) {
          const double grat = dpm / dpk;
          const double bet = 2. + grat + grat - g_cur;
          const double qk = (3. * (a_prev + a_cur) - q_prev) / bet;
          QI(k + j) = qk;
          g_cur = grat / bet;
          GAM(k + j + 1) = g_cur;
          q_prev = qk;
          dpm = dpk; dpk = pn[j] - plast; plast = pn[j];
          a_prev = a_cur; a_cur = an[j];
        }
        k += nb;
      }
      const double grat = dpm / dpk;       // delp(km - 1) / delp(km)
      double q_next = (3. * (a_prev + a_cur) - grat * qs - q_prev) / (2. + grat + grat - g_cur);
      QI(km) = q_next;
      QI(km + 1) = qs;
      for (int k = km - 1; k >= 1;) {
        const int nb = k < RB ? k : RB;
        double qv[RB], gv[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) { qv[j] = QI(k - j); gv[j] = GAM(k - j + 1); }
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) { q_next = qv[j] - gv[j] * q_next; QI(k - j) = q_next; }
        k -= nb;
      }
    } else {
      double d4 = 0., g_prev, q_prev;
      {
        const double grat = dpk / dpm;     // delp(2) / delp(1)
        const double bet = grat * (grat + 0.5);
        q_prev = ((grat + grat) * (grat + 1.) * a_prev + a_cur) / bet;
        g_prev = (1. + grat * (grat + 1.5)) / bet;
        QI(1) = q_prev; GAM(1) = g_prev;
      }
      for (int k = 2; k <= km;) {
        const int nb = km - k + 1 < RB ? km - k + 1 : RB;   // levels k .. k + nb - 1 <= km
        double pn[RB], an[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb && k + j < km) { pn[j] = pe1(k + j + 2); an[j] = A1(k + j + 1); }
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) {
          d4 = dpm / dpk;
          const double bet = 2. + d4 + d4 - g_prev;
          q_prev = (3. * (a_prev + d4 * a_cur) - q_prev) / bet;
          g_prev = d4 / bet;
          QI(k + j) = q_prev; GAM(k + j) = g_prev;
          if (k + j < km) { dpm = dpk; dpk = pn[j] - plast; plast = pn[j]; a_prev = a_cur; a_cur = an[j]; }
        }
        k += nb;
      }
      const double a_bot = 1. + d4 * (d4 + 1.5);
      double q_next = (2. * d4 * (d4 + 1.) * a_cur + a_prev - a_bot * q_prev) / (d4 * (d4 + 0.5) - a_bot * g_prev);
      QI(km + 1) = q_next;
      for (int k = km; k >= 1;) {
        const int nb = k < RB ? k : RB;
        double qv[RB], gv[RB];
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) { qv[j] = QI(k - j); gv[j] = GAM(k - j); }
#pragma unroll
        for (int j = 0; j < RB; j++) if (j < nb) { q_next = qv[j] - gv[j] * q_next; QI(k - j) = q_next; }
        k -= nb;
      }
    }
  }
  // ---- differences gam(k) = a1(k) - a1(k-1), large-scale constraints on the interface values (:639-682 / :1034-1073) and the
  //      continuous first-guess edge values a2(k) = q(k), a3(k) = q(k+1): one sweep, q(k-1) is finished when gam(k) is known.
  //      Loads: A1, QI; stores: GAM, A2, A3.
  {
    double am2 = A1(1), am1 = A1(2);       // a1(k - 2), a1(k - 1)
    A2(1) = QI(1);
    {
      double q2 = QI(2);
      q2 = dmin(q2, dmax(am2, am1));
      q2 = dmax(q2, dmin(am2, am1));
      A2(2) = q2; A3(1) = q2;
    }
    double g_pp = 0., g_prev = am1 - am2;  // gam(k - 2), gam(k - 1)
    GAM(2) = g_prev;
    for (int k = 3; k <= km;) {
      const int nb = km - k + 1 < RB ? km - k + 1 : RB;
      double av[RB], qv[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) if (j < nb) { av[j] = A1(k + j); qv[j] = k + j >= 4 ? QI(k + j - 1) : 0.; }
#pragma unroll
      for (int j = 0; j < RB; j++) if (j < nb) {
        const double a_k = av[j], g_k = a_k - am1;
        GAM(k + j) = g_k;
        if (k + j >= 4) {                  // interface k + j - 1 (3 .. km - 1): gam(k - 2), gam(k), a1(k - 2), a1(k - 1)
          const double lo = dmin(am2, am1), hi = dmax(am2, am1);
          double qk = qv[j];
          if (ak >= 14 || g_pp * g_k > 0.) { qk = dmin(qk, hi); qk = dmax(qk, lo); }
          else if (g_pp > 0.) qk = dmax(qk, lo);
          else { qk = dmin(qk, hi); if (iv == 0) qk = dmax(0., qk); }
          A2(k + j - 1) = qk; A3(k + j - 2) = qk;
        }
        g_pp = g_prev; g_prev = g_k; am2 = am1; am1 = a_k;
      }
      k += nb;
    }
    {                                      // interface km (am2 = a1(km - 1), am1 = a1(km))
      double qk = QI(km);
      qk = dmin(qk, dmax(am2, am1));
      qk = dmax(qk, dmin(am2, am1));
      A2(km) = qk; A3(km - 1) = qk;
    }
    A3(km) = QI(km + 1);
  }
  // ---- extremum flags (:695-715 / :1082-1102).  Loads: A1, A3, GAM; stores: A4, FL.
  {
    double a2 = A2(1), g_k = 0.;
    for (int k = 1; k <= km;) {
      const int nb = km - k + 1 < RB ? km - k + 1 : RB;
      double av[RB], a3v[RB], gv[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) if (j < nb) { av[j] = A1(k + j); a3v[j] = A3(k + j); gv[j] = k + j < km ? GAM(k + j + 1) : 0.; }
#pragma unroll
      for (int j = 0; j < RB; j++) if (j < nb) {
        const int kk = k + j;
        const double a1 = av[j], a3 = a3v[j], g_k1 = gv[j];
        int f;
        if (kk == 1 || kk == km) f = ((a2 - a1) * (a3 - a1) > 0.) ? 1 : 0;
        else f = (g_k * g_k1 < 0.) ? 1 : 0;
        if (ak > 9) {
          const double x0 = 2. * a1 - (a2 + a3), x1 = fabs(a2 - a3);
          const double a4 = 3. * x0;
          A4(kk) = a4;
          if (fabs(x0) > x1) f |= 2;
          if (fabs(a4) > x1) f |= 4;
        }
        FL(kk) = f;
        a2 = a3;                           // the first-guess profile is continuous: a2(k + 1) = a3(k)
        g_k = g_k1;
      }
      k += nb;
    }
  }
  // ---- top two layers (:721-754 / :1109-1140)
  {
    const double a1 = A1(1);
    double a2 = A2(1), a3 = A3(1), a4 = 0.;
    if (iv == 0) a2 = dmax(0., a2);
    else if (iv == -1) { if (a2 * a1 <= 0.) a2 = 0.; }
    else if (iv == 2) { a2 = a1; a3 = a1; a4 = 0.; }
    if (iv != 2) {
      a4 = 3. * (2. * a1 - (a2 + a3));
      cs_limiters(a1, a2, a3, a4, FL(1) & 1, 1);
    }
    A2(1) = a2; A3(1) = a3; A4(1) = a4;
  }
  {
    const double a1 = A1(2);
    double a2 = A2(2), a3 = A3(2), a4 = 3. * (2. * a1 - (a2 + a3));
    cs_limiters(a1, a2, a3, a4, FL(2) & 1, 2);
    A2(2) = a2; A3(2) = a3; A4(2) = a4;
  }
  // ---- Huynh's second constraint in the interior (:759-893 / :1142-1276).  Loads: FL, GAM, A1 .. A4 at the batch's own levels
  //      (the stores of an iteration go to its own level only).
  if (km >= 5) {
    int fm = FL(2), f0 = FL(3);
    double gm = GAM(2), g0 = GAM(3), g1 = GAM(4);      // gam(k - 1), gam(k), gam(k + 1)
    // one level on registers; called with compile-time batch slots so that the batch arrays below stay in registers
    auto level = [&](int k, int fp, double g2, double a1, double a2, double a3, double a4) {
      const bool small = scalar && a1 < qmin;
      auto huynh = [&]() {
        const double pmp_1 = a1 - 2. * g1, lac_1 = pmp_1 + 1.5 * g2;
        a2 = dmin(dmax(a2, dmin3(a1, pmp_1, lac_1)), dmax3(a1, pmp_1, lac_1));
        const double pmp_2 = a1 + 2. * g0, lac_2 = pmp_2 - 1.5 * gm;
        a3 = dmin(dmax(a3, dmin3(a1, pmp_2, lac_2)), dmax3(a1, pmp_2, lac_2));
      };
      auto flat = [&]() { a2 = a1; a3 = a1; a4 = 0.; };
      auto a6_a = [&]() { return 3. * (2. * a1 - (a2 + a3)); };
      auto a6_b = [&]() { return 6. * a1 - 3. * (a2 + a3); };
      switch (ak) {
        case 8:
          huynh();
          a4 = a6_a();
          break;
        case 9:
          if ((f0 & 1) && ((fm & 1) || (fp & 1) || small)) flat();
          else {
            a4 = scalar ? a6_a() : a6_b();
            if (fabs(a4) > fabs(a2 - a3)) {
              huynh();
              a4 = scalar ? a6_a() : a6_b();
            }
          }
          break;
        case 10:
          if (f0 & 1) {
            if (small || (fm & 1) || (fp & 1)) flat();
            else a4 = a6_b();
          } else {
            a4 = a6_b();
            if (fabs(a4) > fabs(a2 - a3)) {
              huynh();
              a4 = a6_b();
            }
          }
          break;
        case 11:
          if ((f0 & 2) && ((fm & 2) || (fp & 2) || small)) flat();
          else a4 = a6_a();
          break;
        case 12:
          if (f0 & 2) {
            if ((fm & 2) || (fp & 2)) { a2 = a1; a3 = a1; }
            else if ((fm & 4) || (fp & 4)) huynh();
          } else if (f0 & 4) {
            if ((fm & 2) || (fp & 2)) huynh();
          }
          a4 = a6_a();
          break;
        case 13:
          a4 = a6_a();
          break;
        case 14:   // strict monotonicity constraint (a4 as the flag sweep left it)
          cs_limiters(a1, a2, a3, a4, f0 & 1, 2);
          break;
        default:   // 15
          cs_limiters(a1, a2, a3, a4, f0 & 1, 1);
          break;
      }
      if (iv == 0 && ak <= 13) cs_limiters(a1, a2, a3, a4, f0 & 1, 0);
      A2(k) = a2; A3(k) = a3; A4(k) = a4;
      fm = f0; f0 = fp; gm = g0; g0 = g1; g1 = g2;
    };
    static_assert(RB == 4, "the batch slots below are written out for RB = 4");
    for (int kb = 3; kb <= km - 2;) {
      const int nb = km - 2 - kb + 1 < RB ? km - 2 - kb + 1 : RB;
      int fpv[RB];
      double g2v[RB], a1v[RB], a2v[RB], a3v[RB], a4v[RB];
#pragma unroll
      for (int j = 0; j < RB; j++) if (j < nb) {
        fpv[j] = FL(kb + j + 1); g2v[j] = GAM(kb + j + 2); a1v[j] = A1(kb + j); a2v[j] = A2(kb + j); a3v[j] = A3(kb + j);
        a4v[j] = ak >= 14 ? A4(kb + j) : 0.;
      }
      level(kb, fpv[0], g2v[0], a1v[0], a2v[0], a3v[0], a4v[0]);
      if (nb > 1) level(kb + 1, fpv[1], g2v[1], a1v[1], a2v[1], a3v[1], a4v[1]);
      if (nb > 2) level(kb + 2, fpv[2], g2v[2], a1v[2], a2v[2], a3v[2], a4v[2]);
      if (nb > 3) level(kb + 3, fpv[3], g2v[3], a1v[3], a2v[3], a3v[3], a4v[3]);
      kb += nb;
    }
  }
  // ---- bottom two layers (:898-914 / :1281-1298)
  for (int k = km - 1; k <= km; k++) {
    const double a1 = A1(k);
    double a2 = A2(k), a3 = A3(k);
    if (k == km) {
      if (iv == 0) a3 = dmax(0., a3);
      else if (iv == -1) { if (a3 * a1 <= 0.) a3 = 0.; }
    }
    double a4 = 3. * (2. * a1 - (a2 + a3));
    cs_limiters(a1, a2, a3, a4, FL(k) & 1, k == km - 1 ? 2 : 1);
    A2(k) = a2; A3(k) = a3; A4(k) = a4;
  }
}

// ---- ppm_profile (fv_operators.F90:1382-1639) and ppm_limiters (:1642-1723): the piecewise parabolic reconstruction of the
// schemes kord <= 7.  Scratch use: A1 layer means, A2 / A3 / A4 the parabola, GAM the limited slopes dc, QI Huynh's h2.  The layer
// thicknesses are differences of the source interface pressures, as the callers form dp1 (:80-84).
RMP_HD inline double dsign(double a, double b) { return copysign(fabs(a), b); }   // Fortran sign(a, b)

RMP_HD inline void ppm_limiters(double dm, double a1, double& a2, double& a3, double& a4, int lmt) {
  if (lmt == 0) {          // standard PPM constraint
    if (dm == 0.) { a2 = a1; a3 = a1; a4 = 0.; }
    else {
      const double da1 = a3 - a2, da2 = da1 * da1, a6da = a4 * da1;
      if (a6da < -da2) { a4 = 3. * (a2 - a1); a3 = a2 - a4; }
      else if (a6da > da2) { a4 = 3. * (a3 - a1); a2 = a3 - a4; }
    }
  } else if (lmt == 1) {   // improved full monotonicity constraint
    const double qmp = 2. * dm;
    const double n2 = a1 - dsign(dmin(fabs(qmp), fabs(a2 - a1)), qmp);
    const double n3 = a1 + dsign(dmin(fabs(qmp), fabs(a3 - a1)), qmp);
    a2 = n2; a3 = n3;
    a4 = 3. * (2. * a1 - (a2 + a3));
  } else if (lmt == 2) {   // positive definite constraint
    if (fabs(a3 - a2) < -a4) {
      const double fmin = a1 + 0.25 * ((a3 - a2) * (a3 - a2)) / a4 + a4 * r12;
      if (fmin < 0.) {
        if (a1 < a3 && a1 < a2) { a3 = a1; a2 = a1; a4 = 0.; }
        else if (a3 > a2) { a4 = 3. * (a2 - a1); a3 = a2 - a4; }
        else { a4 = 3. * (a3 - a1); a2 = a3 - a4; }
      }
    }
  }                        // lmt == 3: nothing
}

template <class P1>
RMP_HD inline void ppm_profile(const Col& C, int km, const P1& pe1, int iv, int kord) {
  const int km1 = km - 1;
  // ---- limited slopes dc(2 .. km-1) (:1429-1439) and the 4th-order provisional edge values a2(3 .. km-1) (:1445-1454): one sweep,
  //      thicknesses and means of the neighbouring layers carried in registers
  {
    double plast = pe1(4);
    double dm2 = 0., dm1, d0, dp1;         // delp(k - 2), delp(k - 1), delp(k), delp(k + 1)
    { const double p1 = pe1(1), p2 = pe1(2), p3 = pe1(3); dm1 = p2 - p1; d0 = p3 - p2; dp1 = plast - p3; }
    double am1 = A1(1), a0 = A1(2), ap1 = A1(3);
    double dcm = 0.;                       // dc(k - 1)
    for (int k = 2; k <= km1; k++) {
      const double delq_k = ap1 - a0, delq_m = a0 - am1;
      const double d4_k = dm1 + d0, d4_p = d0 + dp1;
      const double c1 = (dm1 + 0.5 * d0) / d4_p;
      const double c2 = (dp1 + 0.5 * d0) / d4_k;
      const double df2 = d0 * (c1 * delq_k + c2 * delq_m) / (d4_k + dp1);
      const double dc = dsign(dmin(dmin(fabs(df2), dmax3(am1, a0, ap1) - a0), a0 - dmin3(am1, a0, ap1)), df2);
      GAM(k) = dc;
      if (k >= 3) {
        const double d4_m = dm2 + dm1;
        const double e1 = delq_m * dm1 / d4_k;
        const double b1 = d4_m / (d4_k + dm1);
        const double b2 = d4_p / (d4_k + d0);
        A2(k) = am1 + e1 + 2. / (d4_m + d4_p) * (d0 * (e1 * (b1 - b2) + b2 * dcm) - dm1 * b1 * dc);
      }
      if (k < km1) {
        const double pn = pe1(k + 3);
        dm2 = dm1; dm1 = d0; d0 = dp1; dp1 = pn - plast; plast = pn;
        am1 = a0; a0 = ap1; ap1 = A1(k + 2);
        dcm = dc;
      }
    }
  }
  // ---- top: area preserving cubic with zero second derivative at the boundary (:1460-1496)
  {
    const double p1 = pe1(1), p2 = pe1(2), p3 = pe1(3);
    const double d1 = p2 - p1, d2 = p3 - p2;
    const double a11 = A1(1), a12 = A1(2);
    const double qm = (d2 * a11 + d1 * a12) / (d1 + d2);
    const double dq = 2. * (a12 - a11) / (d1 + d2);
    const double c1 = 4. * (A2(3) - qm - d2 * dq) / (d2 * (2. * d2 * d2 + d1 * (d2 + 3. * d1)));
    const double c3 = dq - 0.5 * c1 * (d2 * (5. * d1 + d2) - 3. * d1 * d1);
    double a22 = qm - 0.25 * c1 * d1 * d2 * (d2 + 3. * d1);
    double a21 = d1 * (2. * c1 * (d1 * d1) - c3) + a22;
    a22 = dmax(a22, dmin(a11, a12));
    a22 = dmin(a22, dmax(a11, a12));
    GAM(1) = 0.5 * (a22 - a11);
    if (iv == 0) { a21 = dmax(0., a21); a22 = dmax(0., a22); }
    else if (iv == -1) { if (a21 * a11 <= 0.) a21 = 0.; }
    else if (iv == 2 || iv == -2) a21 = a11;   // (a3(1) = a1(1) of :1494 is overwritten by a3(1) = a2(2) below)
    A2(1) = a21; A2(2) = a22;
  }
  // ---- bottom (:1500-1548)
  {
    const double p1 = pe1(km1), p2 = pe1(km), p3 = pe1(km + 1);
    const double d1 = p3 - p2, d2 = p2 - p1;
    const double a1m = A1(km), a1n = A1(km1);
    const double qm = (d2 * a1m + d1 * a1n) / (d1 + d2);
    const double dq = 2. * (a1n - a1m) / (d1 + d2);
    const double c1 = (A2(km1) - qm - d2 * dq) / (d2 * (2. * d2 * d2 + d1 * (d2 + 3. * d1)));
    const double c3 = dq - 2.0 * c1 * (d2 * (5. * d1 + d2) - 3. * d1 * d1);
    double a2m = qm - c1 * d1 * d2 * (d2 + 3. * d1);
    double a3m = d1 * (8. * c1 * (d1 * d1) - c3) + a2m;
    a2m = dmax(a2m, dmin(a1m, a1n));
    a2m = dmin(a2m, dmax(a1m, a1n));
    GAM(km) = 0.5 * (a1m - a2m);
    if (iv == 0) { a2m = dmax(0., a2m); a3m = dmax(0., a3m); }
    else if (iv < 0) { if (a1m * a3m <= 0.) a3m = 0.; }
    A2(km) = a2m; A3(km) = a3m;
  }
  // ---- Huynh's h2(2 .. km-1) (:1572-1583), only for kord >= 7
  if (kord >= 7) {
    double plast = pe1(3);
    double dm1, d0;
    { const double p1 = pe1(1), p2 = pe1(2); dm1 = p2 - p1; d0 = plast - p2; }
    double dcm = GAM(1), dc0 = GAM(2);
    for (int k = 2; k <= km1; k++) {
      const double pn = pe1(k + 2);
      const double dp1 = pn - plast, dcp = GAM(k + 1);
      QI(k) = 2. * (dcp / dp1 - dcm / dm1) / (d0 + 0.5 * (dm1 + dp1)) * (d0 * d0);
      dm1 = d0; d0 = dp1; plast = pn; dcm = dc0; dc0 = dcp;
    }
  }
  // ---- the parabolas, top to bottom (:1551-1637).  a3(k) = a2(k + 1) (:1551-1555) is read from the first-guess a2 of the next
  //      layer, which no earlier layer has modified; the top two and bottom two layers always use the standard constraint
  int lmt = kord - 3;
  lmt = lmt > 0 ? lmt : 0;
  if (iv == 0) lmt = lmt < 2 ? lmt : 2;
  double a2n = A2(1);                      // first-guess a2 of the layer at hand
  for (int k = 1; k <= km; k++) {
    const double a1 = A1(k), dc = GAM(k);
    double a2 = a2n, a3, a4 = 0.;
    if (k < km) { a2n = A2(k + 1); a3 = a2n; } else a3 = A3(km);
    if (k <= 2 || k >= km1) {
      a4 = 3. * (2. * a1 - (a2 + a3));
      ppm_limiters(dc, a1, a2, a3, a4, 0);
    } else if (kord >= 7) {
      const double pmp = 2. * dc;
      double qmp = a1 + pmp;
      double lac = a1 + 1.5 * QI(k - 1) + dc;
      a3 = dmin(dmax(a3, dmin3(a1, qmp, lac)), dmax3(a1, qmp, lac));
      qmp = a1 - pmp;
      lac = a1 + 1.5 * QI(k + 1) - dc;
      a2 = dmin(dmax(a2, dmin3(a1, qmp, lac)), dmax3(a1, qmp, lac));
      a4 = 3. * (2. * a1 - (a2 + a3));
      if (iv == 0) ppm_limiters(dc, a1, a2, a3, a4, 2);
    } else {
      if (kord != 4) a4 = 3. * (2. * a1 - (a2 + a3));
      if (kord != 6) ppm_limiters(dc, a1, a2, a3, a4, lmt);
    }
    A2(k) = a2; A3(k) = a3; A4(k) = a4;
  }
}

// the conservative mapping loop (fv_operators.F90:88-132 = 183-227 = 399-441): P1 source, P2 target interface pressures;
// out(k, value) stores layer k.  div_dp2: map1_q2 divides by the tabulated target thickness -- the same difference here.
// mapn: the operation order of mapn_tracer (:276-336), which fv_mapz uses for nq > 5 tracers
template <class P1, class P2, class Out>
RMP_HD inline void map_column(const Col& C, int km, const P1& pe1, const P2& pe2, Out&& out, bool mapn = false) {
  int k0 = 1;
  for (int k = 1; k <= km; k++) {
    const double t = pe2(k), b = pe2(k + 1);
    double qsum = 0.;
    bool done = false;
    for (int l = k0; l <= km; l++) {
      const double p0 = pe1(l), p1 = pe1(l + 1);
      if (t >= p0 && t <= p1) {
        const double dpl = p1 - p0;
        const double pl = (t - p0) / dpl;
        const double a2 = A2(l), a3 = A3(l), a4 = A4(l);
        if (b <= p1) {
          const double pr = (b - p0) / dpl;
          if (mapn) {
            double fac1 = pr + pl;
            const double fac2 = r3 * (pr * fac1 + pl * pl);
            fac1 = 0.5 * fac1;
            out(k, a2 + (a4 + a3 - a2) * fac1 - a4 * fac2);
          } else out(k, a2 + 0.5 * (a4 + a3 - a2) * (pr + pl) - a4 * r3 * (pr * (pr + pl) + pl * pl));
          k0 = l;
          done = true;
        } else {
          if (mapn) {
            double fac1 = 1. + pl;
            const double fac2 = r3 * (1. + pl * fac1);
            fac1 = 0.5 * fac1;
            qsum = (p1 - t) * (a2 + (a4 + a3 - a2) * fac1 - a4 * fac2);
          } else qsum = (p1 - t) * (a2 + 0.5 * (a4 + a3 - a2) * (1. + pl) - a4 * (r3 * (1. + pl * (1. + pl))));
          for (int m = l + 1; m <= km; m++) {
            const double m0 = pe1(m), m1 = pe1(m + 1);
            if (b > m1) qsum = qsum + (m1 - m0) * A1(m);
            else {
              const double dp = b - m0, esl = dp / (m1 - m0);
              if (mapn) { const double fac1 = 0.5 * esl, fac2 = 1. - r23 * esl; qsum = qsum + dp * (A2(m) + fac1 * (A3(m) - A2(m) + A4(m) * fac2)); }
              else qsum = qsum + dp * (A2(m) + 0.5 * esl * (A3(m) - A2(m) + A4(m) * (1. - r23 * esl)));
              k0 = m;
              break;
            }
          }
        }
        break;
      }
    }
    if (!done) out(k, qsum / (b - t));
  }
}

// map_scalar / map1_ppm / map1_q2 (mapn: in the operation order of mapn_tracer) of one column, in place on fld (level k at
// fld[(k-1)*plane]); scalar: scalar_profile (with the q_min tests), else cs_profile.  PPM: the instantiation that also holds
// ppm_profile for abs(kord) <= 7 (:86-90, 181-185, 394-398); the kernels of the schemes 8..15 are compiled without it
template <bool PPM, class P1, class P2>
RMP_HD inline void remap_field(const Col& C, int km, const P1& pe1, const P2& pe2, double* fld, double qs, int iv, int kord, double qmin,
                               bool scalar, bool mapn = false) {
  for (int k = 1; k <= km;) {   // batched like the sweeps of profile()
    const int nb = km - k + 1 < RB ? km - k + 1 : RB;
    double v[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) if (j < nb) v[j] = LV(fld, k + j);
#pragma unroll
    for (int j = 0; j < RB; j++) if (j < nb) A1(k + j) = v[j];
    k += nb;
  }
  if constexpr (PPM) {
    const int ak = kord < 0 ? -kord : kord;
    if (ak <= 7) ppm_profile(C, km, pe1, iv, ak);
    else profile(C, km, pe1, qs, iv, ak, qmin, scalar);
  } else
  profile(C, km, pe1, qs, iv, kord < 0 ? -kord : kord, qmin, scalar);
  map_column(C, km, pe1, pe2, [&](int k, double v) { LV(fld, k) = v; }, mapn);
}

// fillz (fv_fill.F90:34-139; the default branch, DEV_GFS_PHYS not defined) for one column of one tracer, in place on q with the
// layer thicknesses dp (level k at [(k-1)*plane]): negative mixing ratios borrow mass from the layer above, then below; columns
// that needed it get the non-local rescaling of :113-135 (dm(k) = q(k) dp(k) is formed again in each of its three passes instead of
// being kept in a column array: the same product).  The three layers a step works on are carried in registers.
RMP_HD inline void fillz_column(int km, double* q, const double* dp, long long plane) {
  auto Q = [&](int k) -> double& { return q[(long long)(k - 1) * plane]; };
  auto DP = [&](int k) { return dp[(long long)(k - 1) * plane]; };
  auto pos = [](double x) { return x > 0. ? x : 0.; };   // max(0., x)
  double dpm = DP(1), dp0 = DP(2);         // dp(k - 1), dp(k)
  double qm = Q(1), q0 = Q(2);             // q(k - 1), q(k)
  if (qm < 0.) { q0 = q0 + qm * dpm / dp0; qm = 0.; }
  bool zfix = false;
  for (int k = 2; k <= km - 1; k++) {
    const double dpp = DP(k + 1);
    double qp = Q(k + 1);
    if (q0 < 0.) {
      zfix = true;
      if (qm > 0.) {                       // borrow from above
        const double dq = dmin(qm * dpm, -q0 * dp0);
        qm = qm - dq / dpm;
        q0 = q0 + dq / dp0;
      }
      if (q0 < 0. && qp > 0.) {            // borrow from below
        const double dq = dmin(qp * dpp, -q0 * dp0);
        qp = qp - dq / dpp;
        q0 = q0 + dq / dp0;
      }
    }
    Q(k - 1) = qm;
    qm = q0; q0 = qp; dpm = dp0; dp0 = dpp;
  }
  if (q0 < 0. && qm > 0.) {                // bottom layer (qm = q(km - 1), q0 = q(km))
    zfix = true;
    const double qup = qm * dpm, qly = -q0 * dp0;
    const double dup = dmin(qly, qup);
    qm = qm - dup / dpm;
    q0 = q0 + dup / dp0;
  }
  Q(km - 1) = qm; Q(km) = q0;
  if (zfix) {
    double sum0 = 0.;
    for (int k = 2; k <= km; k++) sum0 = sum0 + Q(k) * DP(k);
    if (sum0 > 0.) {
      double sum1 = 0.;
      for (int k = 2; k <= km; k++) sum1 = sum1 + pos(Q(k) * DP(k));
      const double fac = sum0 / sum1;
      for (int k = 2; k <= km; k++) { const double d = DP(k); Q(k) = pos(fac * (Q(k) * d) / d); }
    }
  }
}

}  // namespace rmp
