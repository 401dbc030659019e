// Vertical remapping on device-resident state (the step between the k_split iterations of fv_dynamics).
//
// Reference semantics: model/fv_mapz.F90:56-845 (Lagrangian_to_Eulerian) and the column operators of model/fv_operators.F90:
// map_scalar (:40-134), map1_ppm (:137-229), map1_q2 (:352-443), scalar_profile (:546-916), cs_profile (:919-1300),
// cs_limiters (:1303-1378).  Built: remap_te = F, moist_kappa = F, consv = 0 (no energy fixer), dry air (the last-step T_v -> T
// conversion is the identity), abs(kord) in 8..15, kord_wz > 0 (iv = -2), the tracers of the context's table (no fillz); anything
// else is an error (-2), never a silent fall-back.
//
// Design: column-parallel like the vertical solvers -- one thread per column, consecutive threads on consecutive i, so every
// level access is a coalesced row segment of the [k][NJ][NI] arrays.  The reconstruction (a4(1:4), the interface values, the
// differences and the extremum flags) lives in seven scratch planes of the context instead of per-thread local arrays
// (km = 79..127 levels x 7 arrays would be 4-7 KB of local memory per thread); each sweep keeps its loop-carried and neighbouring
// values in registers.  Measured inside an fv_dynamics step at C384L79 (profiles/prof_fv_dynamics.py, ncu launch list
// profiles/r2/r2_remap_launches.csv): 21 ms per face for six fields (pt, tracer, w, delz, u, v) against 58 ms for the eight
// acoustic substeps it follows.  The kernels are bound by occupancy x memory-level parallelism, not by bytes (1.5 TB/s): every
// sweep is a chain of L2 / DRAM round trips with one or two loads in flight per thread, so the launch bounds trade registers for
// resident warps (64 -> 40 and 124 -> 64 registers bought 17 %); the next step is to batch the loads of several levels per
// thread (software pipelining) and to split k_remap_cells per field.
#include "fv3_ctx.hpp"
#include <cmath>
#include <string>
#include <vector>

namespace {

constexpr double r3 = 1. / 3., r23 = 2. / 3., r12 = 1. / 12.;
constexpr int CB = 128;

__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin3(double a, double b, double c) { return dmin(dmin(a, b), c); }
__device__ __forceinline__ double dmax3(double a, double b, double c) { return dmax(dmax(a, b), c); }

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
__device__ __forceinline__ void cs_limiters(double a1, double& a2, double& a3, double& a4, bool extm, int iv) {
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
template <class P1>
__device__ void profile(const Col& C, int km, const P1& pe1, double qs, int iv, int ak, double qmin, bool scalar) {
  // ---- interface values: tridiagonal solve (:570-622 / :941-1013)
  {
    double plast = pe1(3);
    double dpm, dpk;                       // delp(k - 1), delp(k)
    { const double p1 = pe1(1), p2 = pe1(2); dpm = p2 - p1; dpk = plast - p2; }
    double a_prev = A1(1), a_cur = A1(2);  // A1(k - 1), A1(k)
    if (iv == -2) {   // lower boundary condition q(km+1) = qs
      double g_cur = 0.5, q_prev = 1.5 * a_prev;
      GAM(2) = g_cur;
      QI(1) = q_prev;
      for (int k = 2; k <= km - 1; k++) {
        const double grat = dpm / dpk;
        const double bet = 2. + grat + grat - g_cur;
        const double qk = (3. * (a_prev + a_cur) - q_prev) / bet;
        QI(k) = qk;
        g_cur = grat / bet;
        GAM(k + 1) = g_cur;
        q_prev = qk;
        dpm = dpk; { const double pn = pe1(k + 2); dpk = pn - plast; plast = pn; }
        a_prev = a_cur; a_cur = A1(k + 1);
      }
      const double grat = dpm / dpk;       // delp(km - 1) / delp(km)
      double q_next = (3. * (a_prev + a_cur) - grat * qs - q_prev) / (2. + grat + grat - g_cur);
      QI(km) = q_next;
      QI(km + 1) = qs;
      for (int k = km - 1; k >= 1; k--) { q_next = QI(k) - GAM(k + 1) * q_next; QI(k) = q_next; }
    } else {
      double d4, g_prev, q_prev;
      {
        const double grat = dpk / dpm;     // delp(2) / delp(1)
        const double bet = grat * (grat + 0.5);
        q_prev = ((grat + grat) * (grat + 1.) * a_prev + a_cur) / bet;
        g_prev = (1. + grat * (grat + 1.5)) / bet;
        QI(1) = q_prev; GAM(1) = g_prev;
      }
      for (int k = 2;; k++) {
        d4 = dpm / dpk;
        const double bet = 2. + d4 + d4 - g_prev;
        q_prev = (3. * (a_prev + d4 * a_cur) - q_prev) / bet;
        g_prev = d4 / bet;
        QI(k) = q_prev; GAM(k) = g_prev;
        if (k == km) break;
        dpm = dpk; { const double pn = pe1(k + 2); dpk = pn - plast; plast = pn; }
        a_prev = a_cur; a_cur = A1(k + 1);
      }
      const double a_bot = 1. + d4 * (d4 + 1.5);
      double q_next = (2. * d4 * (d4 + 1.) * a_cur + a_prev - a_bot * q_prev) / (d4 * (d4 + 0.5) - a_bot * g_prev);
      QI(km + 1) = q_next;
      for (int k = km; k >= 1; k--) { q_next = QI(k) - GAM(k) * q_next; QI(k) = q_next; }
    }
  }
  // ---- differences gam(k) = a1(k) - a1(k-1), large-scale constraints on the interface values (:639-682 / :1034-1073) and the
  //      continuous first-guess edge values a2(k) = q(k), a3(k) = q(k+1): one sweep, q(k-1) is finished when gam(k) is known
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
    for (int k = 3; k <= km; k++) {
      const double a_k = A1(k), g_k = a_k - am1;
      GAM(k) = g_k;
      if (k >= 4) {                        // interface k - 1 (3 .. km - 1): gam(k - 2), gam(k), a1(k - 2), a1(k - 1)
        const double lo = dmin(am2, am1), hi = dmax(am2, am1);
        double qk = QI(k - 1);
        if (ak >= 14 || g_pp * g_k > 0.) { qk = dmin(qk, hi); qk = dmax(qk, lo); }
        else if (g_pp > 0.) qk = dmax(qk, lo);
        else { qk = dmin(qk, hi); if (iv == 0) qk = dmax(0., qk); }
        A2(k - 1) = qk; A3(k - 2) = qk;
      }
      g_pp = g_prev; g_prev = g_k; am2 = am1; am1 = a_k;
    }
    {                                      // interface km (am2 = a1(km - 1), am1 = a1(km))
      double qk = QI(km);
      qk = dmin(qk, dmax(am2, am1));
      qk = dmax(qk, dmin(am2, am1));
      A2(km) = qk; A3(km - 1) = qk;
    }
    A3(km) = QI(km + 1);
  }
  // ---- extremum flags (:695-715 / :1082-1102)
  {
    double a2 = A2(1), g_k = 0.;
    for (int k = 1; k <= km; k++) {
      const double a1 = A1(k), a3 = A3(k);
      const double g_k1 = k < km ? GAM(k + 1) : 0.;
      int f;
      if (k == 1 || k == km) f = ((a2 - a1) * (a3 - a1) > 0.) ? 1 : 0;
      else f = (g_k * g_k1 < 0.) ? 1 : 0;
      if (ak > 9) {
        const double x0 = 2. * a1 - (a2 + a3), x1 = fabs(a2 - a3);
        const double a4 = 3. * x0;
        A4(k) = a4;
        if (fabs(x0) > x1) f |= 2;
        if (fabs(a4) > x1) f |= 4;
      }
      FL(k) = f;
      a2 = a3;                             // the first-guess profile is continuous: a2(k + 1) = a3(k)
      g_k = g_k1;
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
  // ---- Huynh's second constraint in the interior (:759-893 / :1142-1276)
  if (km >= 5) {
    int fm = FL(2), f0 = FL(3);
    double gm = GAM(2), g0 = GAM(3), g1 = GAM(4);      // gam(k - 1), gam(k), gam(k + 1)
    for (int k = 3; k <= km - 2; k++) {
      const int fp = FL(k + 1);
      const double g2 = GAM(k + 2);
      const double a1 = A1(k);
      double a2 = A2(k), a3 = A3(k), a4 = ak >= 14 ? A4(k) : 0.;
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

// the conservative mapping loop (fv_operators.F90:88-132 = 183-227 = 399-441): P1 source, P2 target interface pressures;
// out(k, value) stores layer k.  div_dp2: map1_q2 divides by the tabulated target thickness -- the same difference here.
// mapn: the operation order of mapn_tracer (:276-336), which fv_mapz uses for nq > 5 tracers
template <class P1, class P2, class Out>
__device__ void map_column(const Col& C, int km, const P1& pe1, const P2& pe2, Out&& out, bool mapn = false) {
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

struct Scr { double *a1, *a2, *a3, *a4, *q, *gam, *fl, *pe4; };
__device__ __forceinline__ Col col_of(const Scr& S, long long o, long long plane) {
  return Col{S.a1 + o, S.a2 + o, S.a3 + o, S.a4 + o, S.q + o, S.gam + o, reinterpret_cast<int*>(S.fl) + o, plane};
}

// map_scalar / map1_ppm / map1_q2 of one column, in place on fld (level k at fld[(k-1)*plane])
template <class P1, class P2>
__device__ void remap_field(const Col& C, int km, const P1& pe1, const P2& pe2, double* fld, double qs, int iv, int kord, double qmin,
                            bool scalar, bool mapn = false) {
  for (int k = 1; k <= km; k++) A1(k) = LV(fld, k);
  profile(C, km, pe1, qs, iv, kord < 0 ? -kord : kord, qmin, scalar);
  map_column(C, km, pe1, pe2, [&](int k, double v) { LV(fld, k) = v; }, mapn);
}

struct L2E {
  double *pt, *delp, *delz, *w, *u, *v, *pk, *pkz, *omga, *qtr, *pe, *peln;
  double* const* qtrs;   // device table of the tracer arrays (use_tracer entries)
  const double *ws, *ak, *bk;   // ak, bk: device tables (km + 1)
  double akap, k1k, rrg, ptop, t_min, r_vir;
  int sphum;             // index of the specific-humidity tracer in qtrs, or -1
  int hydrostatic, last_step, kord_mt, kord_wz, kord_tm, use_tracer, kord_tr;
};

// steps 0 - 3.3 of fv_mapz.F90 (:188-526) for the cell columns
__global__ void __launch_bounds__(CB, 8) k_remap_cells(Lay L, L2E a, Scr S) {
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const int t = blockIdx.x * CB + threadIdx.x;
  if (t >= nx * ny) return;
  const int i = L.is + t % nx, j = L.js + t / nx, km = L.npz;
  const long long o = LIDX(L, i, j);
  const Col C = col_of(S, o, L.plane);
  double *pt = a.pt + o, *delp = a.delp + o, *delz = a.delz + o, *pe = a.pe + o, *peln = a.peln + o, *pk = a.pk + o, *pkz = a.pkz + o;
  double* pe4 = S.pe4 + o;
  const double ps = LV(pe, km + 1);
  auto pe1 = [&](int k) { return LV(pe, k); };
  auto pe2 = [&](int k) { return k == 1 ? a.ptop : k == km + 1 ? ps : a.ak[k - 1] + a.bk[k - 1] * ps; };
  if (a.kord_tm < 0) {   // theta_v -> T_v (:200-230)
    if (a.hydrostatic) for (int k = 1; k <= km; k++) LV(pt, k) = LV(pt, k) * (LV(pk, k + 1) - LV(pk, k)) / (a.akap * (LV(peln, k + 1) - LV(peln, k)));
    else for (int k = 1; k <= km; k++) { const double p = LV(pt, k); LV(pt, k) = p * exp(a.k1k * log(a.rrg * LV(delp, k) / LV(delz, k) * p)); }
  }
  if (!a.hydrostatic) for (int k = 1; k <= km; k++) LV(delz, k) = -LV(delz, k) / LV(delp, k);   // :297-303
  for (int k = 1; k <= km; k++) LV(delp, k) = pe2(k + 1) - pe2(k);                              // :318-331
  // pn2 = log(pe2) is recomputed where it is needed (the temperature map, then the peln / pk update): no column array is kept
  auto pn1 = [&](int k) { return LV(peln, k); };
  auto pn2 = [&](int k) { return (k == 1 || k == km + 1) ? LV(peln, k) : log(pe2(k)); };
  if (a.kord_tm < 0) remap_field(C, km, pn1, pn2, pt, 0., 1, a.kord_tm, a.t_min, true);          // :373-386
  else remap_field(C, km, pe1, pe2, pt, 0., 1, a.kord_tm, 0., false);
  for (int iq = 0; iq < a.use_tracer; iq++) remap_field(C, km, pe1, pe2, a.qtrs[iq] + o, 0., 0, a.kord_tr, 0., true, a.use_tracer > 5);   // :390-408 (map1_q2; mapn_tracer for nq > 5)
  if (!a.hydrostatic) {                                                                          // :411-433
    remap_field(C, km, pe1, pe2, a.w + o, a.ws[o], -2, a.kord_wz, 0., false);
    remap_field(C, km, pe1, pe2, delz, 0., 1, a.kord_tm, 0., false);
    for (int k = 1; k <= km; k++) LV(delz, k) = -LV(delz, k) * (pe2(k + 1) - pe2(k));
  }
  // 3.1 / 3.2: pk, peln, pkz (:436-506); the old peln (pe0) and omega (pe3) of the last step are parked in the a1 / a2 scratch
  if (a.last_step) {
    A2(1) = 0.;
    for (int k = 2; k <= km + 1; k++) LV(C.a2, k) = LV(a.omga + o, k - 1);
    for (int k = 1; k <= km + 1; k++) LV(C.a1, k) = LV(peln, k);
  }
  for (int k = 2; k <= km; k++) { const double pn = log(pe2(k)); LV(peln, k) = pn; LV(pk, k) = exp(a.akap * pn); }
  if (a.hydrostatic) for (int k = 1; k <= km; k++) LV(pkz, k) = (LV(pk, k + 1) - LV(pk, k)) / (a.akap * (LV(peln, k + 1) - LV(peln, k)));
  else {
    const double ex = a.kord_tm < 0 ? a.akap : a.k1k;
    for (int k = 1; k <= km; k++) LV(pkz, k) = exp(ex * log(a.rrg * LV(delp, k) / LV(delz, k) * LV(pt, k)));
  }
  if (a.kord_tm > 0) for (int k = 1; k <= km; k++) LV(pt, k) = LV(pt, k) * LV(pkz, k);
  if (a.last_step) {   // 3.3 omega to the new layer centres (:509-526)
    double* om = a.omga + o;
    int k_next = 1;
    for (int n = 1; n <= km; n++) {
      const double pc = 0.5 * (LV(peln, n) + LV(peln, n + 1));
      for (int k = k_next; k <= km; k++) {
        const double l0 = LV(C.a1, k), l1 = LV(C.a1, k + 1);
        if (pc <= l1 && pc >= l0) {
          const double e0 = LV(C.a2, k), e1 = LV(C.a2, k + 1);
          LV(om, n) = e0 + (e1 - e0) * (pc - l0) / (l1 - l0);
          k_next = k;
          break;
        }
      }
    }
  }
  for (int k = 2; k <= km; k++) LV(pe4, k) = pe2(k);   // :652-668, stored until every wind column has read the old pe
}

// 4.1 / 4.2 (:535-571): u on the south faces (DIR = 0: i in is..ie, j in js..je+1), v on the west faces (DIR = 1)
template <int DIR>
__global__ void __launch_bounds__(CB, 12) k_remap_wind(Lay L, L2E a, Scr S) {
  const int nx = L.ie - L.is + 1 + DIR, ny = L.je - L.js + 1 + (1 - DIR);
  const int t = blockIdx.x * CB + threadIdx.x;
  if (t >= nx * ny) return;
  const int i = L.is + t % nx, j = L.js + t / nx, km = L.npz;
  const long long o = LIDX(L, i, j), om = DIR == 0 ? o - L.NI : o - 1;   // the neighbour column the pressures are averaged with
  const Col C = col_of(S, o, L.plane);
  const double *pa = a.pe + o, *pb = a.pe + om;
  const double ps2 = LV(pb, km + 1) + LV(pa, km + 1);
  auto pe0 = [&](int k) { return k == 1 ? LV(pa, 1) : 0.5 * (LV(pb, k) + LV(pa, k)); };
  auto pe3 = [&](int k) { return (DIR == 1 && k == 1) ? a.ak[0] : a.ak[k - 1] + 0.5 * a.bk[k - 1] * ps2; };
  remap_field(C, km, pe0, pe3, (DIR == 0 ? a.u : a.v) + o, 0., -1, a.kord_mt, 0., false);
}

// the new interface pressures (:661-668) and, between k_split steps, T_v back to theta_v (:833-843)
__global__ void __launch_bounds__(CB) k_remap_finish(Lay L, L2E a, Scr S) {
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const int t = blockIdx.x * CB + threadIdx.x;
  if (t >= nx * ny) return;
  const int i = L.is + t % nx, j = L.js + t / nx, km = L.npz;
  const long long o = LIDX(L, i, j);
  const Col C{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, L.plane};
  for (int k = 2; k <= km; k++) LV(a.pe + o, k) = LV(S.pe4 + o, k);
  if (!a.last_step) for (int k = 1; k <= km; k++) LV(a.pt + o, k) = LV(a.pt + o, k) / LV(a.pkz + o, k);
  else if (a.sphum >= 0) {   // T_v -> T (:792-822 with dtmp = 0, no condensates)
    const double* qv = a.qtrs[a.sphum] + o;
    for (int k = 1; k <= km; k++) LV(a.pt + o, k) = (LV(a.pt + o, k) + 0. * LV(a.pkz + o, k)) / (1. + a.r_vir * LV(qv, k));
  }
}

// stand-alone column operator on FV3_WORK_Q (parity of the profiles for every scheme / iv): pe1 = FV3_PE, pe2 = the hybrid levels
__global__ void __launch_bounds__(CB) k_remap_work_q(Lay L, L2E a, Scr S, int mode, int iv, int kord, double qmin) {
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const int t = blockIdx.x * CB + threadIdx.x;
  if (t >= nx * ny) return;
  const int i = L.is + t % nx, j = L.js + t / nx, km = L.npz;
  const long long o = LIDX(L, i, j);
  const Col C = col_of(S, o, L.plane);
  const double* pe = a.pe + o;
  const double ps = LV(pe, km + 1);
  auto pe1 = [&](int k) { return LV(pe, k); };
  auto pe2 = [&](int k) { return k == 1 ? a.ptop : k == km + 1 ? ps : a.ak[k - 1] + a.bk[k - 1] * ps; };
  remap_field(C, km, pe1, pe2, a.qtr + o, iv == -2 ? a.ws[o] : 0., iv, kord, qmin, mode != 1);
}

int check_kord(fv3_ctx* c, int kord, const char* what) {
  const int ak = kord < 0 ? -kord : kord;
  if (ak < 8 || ak > 15) return fv3_fail(c, -2, std::string("remap: ") + what + " outside 8..15 (ppm_profile, kord <= 7, is not built)");
  return 0;
}

int fill(fv3_ctx* c, L2E& a, Scr& S) {
  const fv3_flags_t& f = c->f;
  const int km = c->L.npz;
  if (!c->d_akbk) {
    std::vector<double> h(2 * (km + 1));
    for (int k = 0; k <= km; k++) { h[k] = c->ak[k]; h[km + 1 + k] = c->bk[k]; }
    FV3_CUDA(c, cudaMalloc(&c->d_akbk, h.size() * sizeof(double)));
    FV3_CUDA(c, cudaMemcpy(c->d_akbk, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  a.pt = c->fld[FV3_PT]; a.delp = c->fld[FV3_DELP]; a.delz = c->fld[FV3_DELZ]; a.w = c->fld[FV3_W]; a.u = c->fld[FV3_U]; a.v = c->fld[FV3_V];
  a.pk = c->fld[FV3_PK]; a.pkz = c->fld[FV3_PKZ]; a.omga = c->fld[FV3_OMGA]; a.qtr = c->fld[FV3_WORK_Q]; a.pe = c->fld[FV3_PE];
  a.peln = c->fld[FV3_PELN]; a.ws = c->fld[FV3_WS]; a.ak = c->d_akbk; a.bk = c->d_akbk + km + 1;
  const double cv_air = f.cp_air - f.rdgas;
  a.akap = f.kappa; a.k1k = f.rdgas / cv_air; a.rrg = -f.rdgas / f.grav; a.ptop = f.ptop; a.t_min = 184.;   // fv_mapz.F90:43
  a.hydrostatic = f.hydrostatic;
  S = Scr{c->scr[0], c->scr[1], c->scr[2], c->scr[3], c->scr[4], c->scr[5], c->scr[6], c->scr[7]};
  return 0;
}

}  // namespace

int stage_remap_work_q(fv3_ctx* c, int mode, int iv, int kord, double qmin) {
  StageScope ts(c, "REMAP_OP");
  if (mode < 0 || mode > 2 || iv < -2 || iv > 2) return fv3_fail(c, -1, "remap_work_q: mode in 0..2, iv in -2..2");
  int rc = check_kord(c, kord, "kord"); if (rc) return rc;
  L2E a{}; Scr S{};
  rc = fill(c, a, S); if (rc) return rc;
  const int n = (c->L.ie - c->L.is + 1) * (c->L.je - c->L.js + 1);
  k_remap_work_q<<<(n + CB - 1) / CB, CB, 0, c->stream>>>(c->L, a, S, mode, iv, kord, qmin);
  c->launches++;
  FV3_CUDA(c, cudaGetLastError());
  return 0;
}

int stage_lagrangian_to_eulerian(fv3_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum, double r_vir) {
  StageScope ts(c, "REMAP");
  const fv3_flags_t& f = c->f;
  if (f.moist_kappa) return fv3_fail(c, -2, "remap: moist_kappa not supported");
  if (kord_wz < 0) return fv3_fail(c, -2, "remap: kord_wz < 0 (iv = -3 lower boundary condition) not supported");
  if (kord_mt < 0) return fv3_fail(c, -2, "remap: kord_mt must be positive");
  int rc = check_kord(c, kord_mt, "kord_mt"); if (rc) return rc;
  rc = check_kord(c, kord_tm, "kord_tm"); if (rc) return rc;
  if (!f.hydrostatic) { rc = check_kord(c, kord_wz, "kord_wz"); if (rc) return rc; }
  if (use_tracer < 0 || use_tracer > 64) return fv3_fail(c, -1, "remap: use_tracer (the number of tracers to remap) in 0..64");
  if (use_tracer) { rc = check_kord(c, kord_tr, "kord_tr"); if (rc) return rc; if (kord_tr < 0) return fv3_fail(c, -2, "remap: kord_tr must be positive"); }
  L2E a{}; Scr S{};
  rc = fill(c, a, S); if (rc) return rc;
  if (use_tracer) {   // the first use_tracer tracers of the context's table (fv3_set_num_tracers; one tracer: FV3_WORK_Q's own array)
    if (c->tracers.empty()) c->tracers.push_back(c->fld[FV3_WORK_Q]);
    if (use_tracer > (int)c->tracers.size()) return fv3_fail(c, -1, "remap: use_tracer exceeds the number of tracers of the context");
    if (!c->d_qtr_tab) FV3_CUDA(c, cudaMalloc(&c->d_qtr_tab, 64 * sizeof(double*)));
    FV3_CUDA(c, cudaMemcpyAsync(c->d_qtr_tab, c->tracers.data(), use_tracer * sizeof(double*), cudaMemcpyHostToDevice, c->stream));
    a.qtrs = c->d_qtr_tab;
  }
  if (sphum >= use_tracer) return fv3_fail(c, -1, "remap: sphum must be one of the remapped tracers (or -1)");
  if (sphum >= 0 && c->f.use_cond) return fv3_fail(c, -2, "remap: specific humidity with use_cond (condensates, moist_cv) not supported");
  a.sphum = sphum < 0 ? -1 : sphum; a.r_vir = r_vir;
  a.last_step = last_step; a.kord_mt = kord_mt; a.kord_wz = kord_wz; a.kord_tm = kord_tm; a.use_tracer = use_tracer; a.kord_tr = kord_tr;
  const Lay& L = c->L;
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  k_remap_cells<<<(nx * ny + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
  k_remap_wind<0><<<(nx * (ny + 1) + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
  k_remap_wind<1><<<((nx + 1) * ny + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
  k_remap_finish<<<(nx * ny + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
  c->launches += 4;
  FV3_CUDA(c, cudaGetLastError());
  return 0;
}
