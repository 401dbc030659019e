// TEST INFRASTRUCTURE (CPU oracle) -- never linked by the product.
//
// Vertical remapping: restatement of model/fv_mapz.F90:56-845 (Lagrangian_to_Eulerian) and of the column operators it calls in
// model/fv_operators.F90: map_scalar (:40-134), map1_ppm (:137-229), map1_q2 (:352-443), scalar_profile (:546-916),
// cs_profile (:919-1300), cs_limiters (:1303-1378), ppm_profile (:1382-1639), ppm_limiters (:1642-1723).  Scope (everything else is refused with -2 by the C entry point):
//   remap_te = F, moist_kappa = F, use_cond = F, consv = 0 (no energy fixer), no intermediate physics; the last-step conversion
//   T_v -> T divides by 1 + r_vir q_v when a specific-humidity tracer is named, else it is the identity (`adiabatic`), abs(kord) in 8..15 (the cs / scalar profiles)
//   or 1..7 (ppm_profile :1382-1639, ppm_limiters :1642-1723), kord_wz > 0 (iv = -2; the iv = -3 branch of cs_profile reads an unset
//   gam(km), :969-985), tracers with map1_q2 or, for nq > 5, with the operation order of mapn_tracer (iv = 0).
// flagstruct%fill: fillz (fv_fill.F90:34-139, default branch) after each tracer.
// The Fortran vectorises every loop over i; here one column is processed at a time (same operations on the same operands in the
// same order for every element).  Parity unpinned: the reference holds no test or golden vector for these routines.
#include "fv3_oracle.hpp"
#include <cmath>
#include <vector>

namespace fv3o {
namespace {

constexpr double r3 = 1. / 3., r23 = 2. / 3., r12 = 1. / 12.;

struct A4 {   // a4(1:4, k), k = 1..km (1-based)
  std::vector<double> v[4];
  explicit A4(int km) { for (auto& x : v) x.assign(km + 2, 0.); }
  double& operator()(int n, int k) { return v[n - 1][k]; }
};

// fv_operators.F90:1303-1378, one element
void cs_limiters(bool extm, A4& a4, int k, int iv) {
  if (iv == 0) {   // positive definite constraint
    if (a4(1, k) <= 0.) { a4(2, k) = a4(1, k); a4(3, k) = a4(1, k); a4(4, k) = 0.; }
    else if (std::fabs(a4(3, k) - a4(2, k)) < -a4(4, k)) {
      if ((a4(1, k) + 0.25 * ((a4(3, k) - a4(2, k)) * (a4(3, k) - a4(2, k))) / a4(4, k) + a4(4, k) * r12) < 0.) {
        if (a4(1, k) < a4(3, k) && a4(1, k) < a4(2, k)) { a4(3, k) = a4(1, k); a4(2, k) = a4(1, k); a4(4, k) = 0.; }
        else if (a4(3, k) > a4(2, k)) { a4(4, k) = 3. * (a4(2, k) - a4(1, k)); a4(3, k) = a4(2, k) - a4(4, k); }
        else { a4(4, k) = 3. * (a4(3, k) - a4(1, k)); a4(2, k) = a4(3, k) - a4(4, k); }
      }
    }
  } else if (iv == 1) {
    if ((a4(1, k) - a4(2, k)) * (a4(1, k) - a4(3, k)) >= 0.) { a4(2, k) = a4(1, k); a4(3, k) = a4(1, k); a4(4, k) = 0.; }
    else {
      const double da1 = a4(3, k) - a4(2, k), da2 = da1 * da1, a6da = a4(4, k) * da1;
      if (a6da < -da2) { a4(4, k) = 3. * (a4(2, k) - a4(1, k)); a4(3, k) = a4(2, k) - a4(4, k); }
      else if (a6da > da2) { a4(4, k) = 3. * (a4(3, k) - a4(1, k)); a4(2, k) = a4(3, k) - a4(4, k); }
    }
  } else {   // standard PPM constraint
    if (extm) { a4(2, k) = a4(1, k); a4(3, k) = a4(1, k); a4(4, k) = 0.; }
    else {
      const double da1 = a4(3, k) - a4(2, k), da2 = da1 * da1, a6da = a4(4, k) * da1;
      if (a6da < -da2) { a4(4, k) = 3. * (a4(2, k) - a4(1, k)); a4(3, k) = a4(2, k) - a4(4, k); }
      else if (a6da > da2) { a4(4, k) = 3. * (a4(3, k) - a4(1, k)); a4(2, k) = a4(3, k) - a4(4, k); }
    }
  }
}

// scalar_profile (scalar = true: fv_operators.F90:546-916, with the q_min tests) / cs_profile (:919-1300) of one column.
// delp(1:km), a4(1,:) = the layer means on entry.  Returns -2 for a scheme outside the restated set.
int profile(double qs, A4& a4, const std::vector<double>& delp, int km, int iv, int kord, double qmin, bool scalar) {
  const int ak = std::abs(kord);
  if (ak < 8 || ak > 15 || iv == -3) return -2;
  std::vector<double> gam(km + 3, 0.), q(km + 3, 0.);
  std::vector<char> extm(km + 2, 0), ext5(km + 2, 0), ext6(km + 2, 0);
  if (iv == -2) {   // lower boundary condition q(km+1) = qs (:570-592 / :941-963)
    gam[2] = 0.5;
    q[1] = 1.5 * a4(1, 1);
    for (int k = 2; k <= km - 1; k++) {
      const double grat = delp[k - 1] / delp[k];
      const double bet = 2. + grat + grat - gam[k];
      q[k] = (3. * (a4(1, k - 1) + a4(1, k)) - q[k - 1]) / bet;
      gam[k + 1] = grat / bet;
    }
    const double grat = delp[km - 1] / delp[km];
    q[km] = (3. * (a4(1, km - 1) + a4(1, km)) - grat * qs - q[km - 1]) / (2. + grat + grat - gam[km]);
    q[km + 1] = qs;
    for (int k = km - 1; k >= 1; k--) q[k] = q[k] - gam[k + 1] * q[k + 1];
  } else {          // (:593-622 / :965-1013)
    double d4 = 0.;
    {
      const double grat = delp[2] / delp[1];
      const double bet = grat * (grat + 0.5);
      q[1] = ((grat + grat) * (grat + 1.) * a4(1, 1) + a4(1, 2)) / bet;
      gam[1] = (1. + grat * (grat + 1.5)) / bet;
    }
    for (int k = 2; k <= km; k++) {
      d4 = delp[k - 1] / delp[k];
      const double bet = 2. + d4 + d4 - gam[k - 1];
      q[k] = (3. * (a4(1, k - 1) + d4 * a4(1, k)) - q[k - 1]) / bet;
      gam[k] = d4 / bet;
    }
    const double a_bot = 1. + d4 * (d4 + 1.5);
    q[km + 1] = (2. * d4 * (d4 + 1.) * a4(1, km) + a4(1, km - 1) - a_bot * q[km]) / (d4 * (d4 + 0.5) - a_bot * gam[km]);
    for (int k = km; k >= 1; k--) q[k] = q[k] - gam[k] * q[k + 1];
  }
  // large-scale constraints on the interface values (:639-682 / :1034-1073)
  q[2] = std::min(q[2], std::max(a4(1, 1), a4(1, 2)));
  q[2] = std::max(q[2], std::min(a4(1, 1), a4(1, 2)));
  for (int k = 2; k <= km; k++) gam[k] = a4(1, k) - a4(1, k - 1);
  for (int k = 3; k <= km - 1; k++) {
    if (ak >= 14 || gam[k - 1] * gam[k + 1] > 0.) {   // all interfaces for the strictly monotone schemes, else away from extrema
      q[k] = std::min(q[k], std::max(a4(1, k - 1), a4(1, k)));
      q[k] = std::max(q[k], std::min(a4(1, k - 1), a4(1, k)));
    } else if (gam[k - 1] > 0.) {
      q[k] = std::max(q[k], std::min(a4(1, k - 1), a4(1, k)));
    } else {
      q[k] = std::min(q[k], std::max(a4(1, k - 1), a4(1, k)));
      if (iv == 0) q[k] = std::max(0., q[k]);
    }
  }
  q[km] = std::min(q[km], std::max(a4(1, km - 1), a4(1, km)));
  q[km] = std::max(q[km], std::min(a4(1, km - 1), a4(1, km)));
  for (int k = 1; k <= km; k++) { a4(2, k) = q[k]; a4(3, k) = q[k + 1]; }
  // extremum flags (:695-715 / :1082-1102)
  for (int k = 1; k <= km; k++) {
    if (k == 1 || k == km) extm[k] = (a4(2, k) - a4(1, k)) * (a4(3, k) - a4(1, k)) > 0.;
    else extm[k] = gam[k] * gam[k + 1] < 0.;
    if (ak > 9) {
      const double x0 = 2. * a4(1, k) - (a4(2, k) + a4(3, k)), x1 = std::fabs(a4(2, k) - a4(3, k));
      a4(4, k) = 3. * x0;
      ext5[k] = std::fabs(x0) > x1;
      ext6[k] = std::fabs(a4(4, k)) > x1;
    }
  }
  // top two layers (:721-754 / :1109-1140)
  if (iv == 0) a4(2, 1) = std::max(0., a4(2, 1));
  else if (iv == -1) { if (a4(2, 1) * a4(1, 1) <= 0.) a4(2, 1) = 0.; }
  else if (iv == 2) { a4(2, 1) = a4(1, 1); a4(3, 1) = a4(1, 1); a4(4, 1) = 0.; }
  if (iv != 2) {
    a4(4, 1) = 3. * (2. * a4(1, 1) - (a4(2, 1) + a4(3, 1)));
    cs_limiters(extm[1], a4, 1, 1);
  }
  a4(4, 2) = 3. * (2. * a4(1, 2) - (a4(2, 2) + a4(3, 2)));
  cs_limiters(extm[2], a4, 2, 2);
  // Huynh's second constraint in the interior (:759-893 / :1142-1276)
  auto huynh = [&](int k) {
    const double pmp_1 = a4(1, k) - 2. * gam[k + 1], lac_1 = pmp_1 + 1.5 * gam[k + 2];
    a4(2, k) = std::min(std::max(a4(2, k), min3(a4(1, k), pmp_1, lac_1)), max3(a4(1, k), pmp_1, lac_1));
    const double pmp_2 = a4(1, k) + 2. * gam[k], lac_2 = pmp_2 - 1.5 * gam[k - 1];
    a4(3, k) = std::min(std::max(a4(3, k), min3(a4(1, k), pmp_2, lac_2)), max3(a4(1, k), pmp_2, lac_2));
  };
  auto flat = [&](int k) { a4(2, k) = a4(1, k); a4(3, k) = a4(1, k); a4(4, k) = 0.; };
  auto a6_a = [&](int k) { return 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k))); };
  auto a6_b = [&](int k) { return 6. * a4(1, k) - 3. * (a4(2, k) + a4(3, k)); };
  for (int k = 3; k <= km - 2; k++) {
    const bool small = scalar && a4(1, k) < qmin;
    switch (ak) {
      case 8:
        huynh(k);
        a4(4, k) = a6_a(k);
        break;
      case 9:
        if (extm[k] && extm[k - 1]) flat(k);
        else if (extm[k] && extm[k + 1]) flat(k);
        else if (scalar && extm[k] && a4(1, k) < qmin) flat(k);
        else {
          a4(4, k) = scalar ? a6_a(k) : a6_b(k);
          if (std::fabs(a4(4, k)) > std::fabs(a4(2, k) - a4(3, k))) {
            huynh(k);
            a4(4, k) = scalar ? a6_a(k) : a6_b(k);
          }
        }
        break;
      case 10:
        if (extm[k]) {
          if (small || extm[k - 1] || extm[k + 1]) flat(k);
          else a4(4, k) = a6_b(k);
        } else {
          a4(4, k) = a6_b(k);
          if (std::fabs(a4(4, k)) > std::fabs(a4(2, k) - a4(3, k))) {
            huynh(k);
            a4(4, k) = a6_b(k);
          }
        }
        break;
      case 11:
        if (ext5[k] && (ext5[k - 1] || ext5[k + 1] || small)) flat(k);
        else a4(4, k) = a6_a(k);
        break;
      case 12:
        if (ext5[k]) {
          if (ext5[k - 1] || ext5[k + 1]) { a4(2, k) = a4(1, k); a4(3, k) = a4(1, k); }
          else if (ext6[k - 1] || ext6[k + 1]) huynh(k);
        } else if (ext6[k]) {
          if (ext5[k - 1] || ext5[k + 1]) huynh(k);
        }
        a4(4, k) = a6_a(k);
        break;
      case 13:
        a4(4, k) = a6_a(k);
        break;
      case 14:   // strict monotonicity constraint (a4(4) as the flag loop left it: 3 * x0)
        cs_limiters(extm[k], a4, k, 2);
        break;
      case 15:
        cs_limiters(extm[k], a4, k, 1);
        break;
    }
    if (iv == 0 && ak <= 13) cs_limiters(extm[k], a4, k, 0);
  }
  // bottom two layers (:898-914 / :1281-1298)
  if (iv == 0) a4(3, km) = std::max(0., a4(3, km));
  else if (iv == -1) { if (a4(3, km) * a4(1, km) <= 0.) a4(3, km) = 0.; }
  for (int k = km - 1; k <= km; k++) {
    a4(4, k) = 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k)));
    if (k == km - 1) cs_limiters(extm[k], a4, k, 2);
    if (k == km) cs_limiters(extm[k], a4, k, 1);
  }
  return 0;
}

// fv_operators.F90:1642-1723, one element (dm = the limited slope dc(k))
void ppm_limiters(double dm, A4& a4, int k, int lmt) {
  if (lmt == 3) return;
  if (lmt == 0) {          // standard PPM constraint
    if (dm == 0.) { a4(2, k) = a4(1, k); a4(3, k) = a4(1, k); a4(4, k) = 0.; }
    else {
      const double da1 = a4(3, k) - a4(2, k), da2 = da1 * da1, a6da = a4(4, k) * da1;
      if (a6da < -da2) { a4(4, k) = 3. * (a4(2, k) - a4(1, k)); a4(3, k) = a4(2, k) - a4(4, k); }
      else if (a6da > da2) { a4(4, k) = 3. * (a4(3, k) - a4(1, k)); a4(2, k) = a4(3, k) - a4(4, k); }
    }
  } else if (lmt == 1) {   // improved full monotonicity constraint (no first guess of a4(4) needed)
    const double qmp = 2. * dm;
    a4(2, k) = a4(1, k) - fsign(std::min(std::fabs(qmp), std::fabs(a4(2, k) - a4(1, k))), qmp);
    a4(3, k) = a4(1, k) + fsign(std::min(std::fabs(qmp), std::fabs(a4(3, k) - a4(1, k))), qmp);
    a4(4, k) = 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k)));
  } else if (lmt == 2) {   // positive definite constraint
    if (std::fabs(a4(3, k) - a4(2, k)) < -a4(4, k)) {
      const double fmin = a4(1, k) + 0.25 * ((a4(3, k) - a4(2, k)) * (a4(3, k) - a4(2, k))) / a4(4, k) + a4(4, k) * r12;
      if (fmin < 0.) {
        if (a4(1, k) < a4(3, k) && a4(1, k) < a4(2, k)) { a4(3, k) = a4(1, k); a4(2, k) = a4(1, k); a4(4, k) = 0.; }
        else if (a4(3, k) > a4(2, k)) { a4(4, k) = 3. * (a4(2, k) - a4(1, k)); a4(3, k) = a4(2, k) - a4(4, k); }
        else { a4(4, k) = 3. * (a4(3, k) - a4(1, k)); a4(2, k) = a4(3, k) - a4(4, k); }
      }
    }
  }
}

// ppm_profile (fv_operators.F90:1382-1639; BOT_MONO not defined; steepz is commented out in the reference) of one column, the
// piecewise parabolic reconstruction map_scalar / map1_ppm / map1_q2 use for kord <= 7.  kord in 1..7, km >= 5.
int ppm_profile(A4& a4, const std::vector<double>& delp, int km, int iv, int kord) {
  if (kord < 1 || kord > 7 || km < 5 || iv == -3) return -2;
  const int km1 = km - 1;
  std::vector<double> dc(km + 2, 0.), h2(km + 2, 0.), delq(km + 2, 0.), df2(km + 2, 0.), d4(km + 2, 0.);
  for (int k = 2; k <= km; k++) { delq[k - 1] = a4(1, k) - a4(1, k - 1); d4[k] = delp[k - 1] + delp[k]; }
  for (int k = 2; k <= km1; k++) {
    const double c1 = (delp[k - 1] + 0.5 * delp[k]) / d4[k + 1];
    const double c2 = (delp[k + 1] + 0.5 * delp[k]) / d4[k];
    df2[k] = delp[k] * (c1 * delq[k] + c2 * delq[k - 1]) / (d4[k] + delp[k + 1]);
    dc[k] = fsign(std::min(std::min(std::fabs(df2[k]), max3(a4(1, k - 1), a4(1, k), a4(1, k + 1)) - a4(1, k)),
                           a4(1, k) - min3(a4(1, k - 1), a4(1, k), a4(1, k + 1))), df2[k]);
  }
  // 4th order interpolation of the provisional cell edge value (:1445-1454)
  for (int k = 3; k <= km1; k++) {
    const double c1 = delq[k - 1] * delp[k - 1] / d4[k];
    const double a1 = d4[k - 1] / (d4[k] + delp[k - 1]);
    const double a2 = d4[k + 1] / (d4[k] + delp[k]);
    a4(2, k) = a4(1, k - 1) + c1 + 2. / (d4[k - 1] + d4[k + 1]) * (delp[k] * (c1 * (a1 - a2) + a2 * dc[k - 1]) - delp[k - 1] * a1 * dc[k]);
  }
  {   // top: area preserving cubic with zero second derivative at the boundary (:1460-1478)
    const double d1 = delp[1], d2 = delp[2];
    const double qm = (d2 * a4(1, 1) + d1 * a4(1, 2)) / (d1 + d2);
    const double dq = 2. * (a4(1, 2) - a4(1, 1)) / (d1 + d2);
    const double c1 = 4. * (a4(2, 3) - qm - d2 * dq) / (d2 * (2. * d2 * d2 + d1 * (d2 + 3. * d1)));
    const double c3 = dq - 0.5 * c1 * (d2 * (5. * d1 + d2) - 3. * d1 * d1);
    a4(2, 2) = qm - 0.25 * c1 * d1 * d2 * (d2 + 3. * d1);
    a4(2, 1) = d1 * (2. * c1 * (d1 * d1) - c3) + a4(2, 2);
    a4(2, 2) = std::max(a4(2, 2), std::min(a4(1, 1), a4(1, 2)));
    a4(2, 2) = std::min(a4(2, 2), std::max(a4(1, 1), a4(1, 2)));
    dc[1] = 0.5 * (a4(2, 2) - a4(1, 1));
  }
  if (iv == 0) { a4(2, 1) = std::max(0., a4(2, 1)); a4(2, 2) = std::max(0., a4(2, 2)); }   // :1482-1496
  else if (iv == -1) { if (a4(2, 1) * a4(1, 1) <= 0.) a4(2, 1) = 0.; }
  else if (std::abs(iv) == 2) { a4(2, 1) = a4(1, 1); a4(3, 1) = a4(1, 1); }
  {   // bottom (:1500-1518)
    const double d1 = delp[km], d2 = delp[km1];
    const double qm = (d2 * a4(1, km) + d1 * a4(1, km1)) / (d1 + d2);
    const double dq = 2. * (a4(1, km1) - a4(1, km)) / (d1 + d2);
    const double c1 = (a4(2, km1) - qm - d2 * dq) / (d2 * (2. * d2 * d2 + d1 * (d2 + 3. * d1)));
    const double c3 = dq - 2.0 * c1 * (d2 * (5. * d1 + d2) - 3. * d1 * d1);
    a4(2, km) = qm - c1 * d1 * d2 * (d2 + 3. * d1);
    a4(3, km) = d1 * (8. * c1 * (d1 * d1) - c3) + a4(2, km);
    a4(2, km) = std::max(a4(2, km), std::min(a4(1, km), a4(1, km1)));
    a4(2, km) = std::min(a4(2, km), std::max(a4(1, km), a4(1, km1)));
    dc[km] = 0.5 * (a4(1, km) - a4(2, km));
  }
  if (iv == 0) { a4(2, km) = std::max(0., a4(2, km)); a4(3, km) = std::max(0., a4(3, km)); }   // :1539-1548
  else if (iv < 0) { if (a4(1, km) * a4(3, km) <= 0.) a4(3, km) = 0.; }
  for (int k = 1; k <= km1; k++) a4(3, k) = a4(2, k + 1);
  // top 2 and bottom 2 layers always use the monotonic mapping
  for (int k = 1; k <= 2; k++) {
    a4(4, k) = 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k)));
    ppm_limiters(dc[k], a4, k, 0);
  }
  if (kord >= 7) {   // Huynh's 2nd constraint (:1568-1614)
    for (int k = 2; k <= km1; k++)
      h2[k] = 2. * (dc[k + 1] / delp[k + 1] - dc[k - 1] / delp[k - 1]) / (delp[k] + 0.5 * (delp[k - 1] + delp[k + 1])) * (delp[k] * delp[k]);
    const double fac = 1.5;
    for (int k = 3; k <= km - 2; k++) {
      const double pmp = 2. * dc[k];
      double qmp = a4(1, k) + pmp;
      double lac = a4(1, k) + fac * h2[k - 1] + dc[k];
      a4(3, k) = std::min(std::max(a4(3, k), min3(a4(1, k), qmp, lac)), max3(a4(1, k), qmp, lac));
      qmp = a4(1, k) - pmp;
      lac = a4(1, k) + fac * h2[k + 1] - dc[k];
      a4(2, k) = std::min(std::max(a4(2, k), min3(a4(1, k), qmp, lac)), max3(a4(1, k), qmp, lac));
      a4(4, k) = 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k)));
      if (iv == 0 && kord >= 6) ppm_limiters(dc[k], a4, k, 2);
    }
  } else {           // (:1616-1630)
    int lmt = std::max(0, kord - 3);
    if (iv == 0) lmt = std::min(2, lmt);
    for (int k = 3; k <= km - 2; k++) {
      if (kord != 4) a4(4, k) = 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k)));
      if (kord != 6) ppm_limiters(dc[k], a4, k, lmt);
    }
  }
  for (int k = km1; k <= km; k++) {
    a4(4, k) = 3. * (2. * a4(1, k) - (a4(2, k) + a4(3, k)));
    ppm_limiters(dc[k], a4, k, 0);
  }
  return 0;
}

// the conservative mapping loop shared by map_scalar / map1_ppm / map1_q2 (fv_operators.F90:88-132, 183-227, 399-441) for one
// column: pe1(1:km+1) -> pe2(1:kn+1).  dp2 != nullptr: divide by dp2(k) (map1_q2) instead of pe2(k+1) - pe2(k).
// mapn: the operation order of mapn_tracer (:276-336: the geometric factors fac1, fac2 are formed first), which fv_mapz uses for nq > 5
void map_column(int km, const std::vector<double>& pe1, A4& q4, const std::vector<double>& dp1, int kn, const std::vector<double>& pe2,
                std::vector<double>& q2, const double* dp2, bool mapn = false) {
  int k0 = 1;
  for (int k = 1; k <= kn; k++) {
    double qsum = 0.;
    bool done = false, have_sum = false;
    for (int l = k0; l <= km && !done && !have_sum; l++) {
      if (pe2[k] >= pe1[l] && pe2[k] <= pe1[l + 1]) {
        const double pl = (pe2[k] - pe1[l]) / dp1[l];
        if (pe2[k + 1] <= pe1[l + 1]) {   // the new layer lies within one old layer
          const double pr = (pe2[k + 1] - pe1[l]) / dp1[l];
          if (mapn) {
            double fac1 = pr + pl;
            const double fac2 = r3 * (pr * fac1 + pl * pl);
            fac1 = 0.5 * fac1;
            q2[k] = q4(2, l) + (q4(4, l) + q4(3, l) - q4(2, l)) * fac1 - q4(4, l) * fac2;
          } else
          q2[k] = q4(2, l) + 0.5 * (q4(4, l) + q4(3, l) - q4(2, l)) * (pr + pl) - q4(4, l) * r3 * (pr * (pr + pl) + pl * pl);
          k0 = l;
          done = true;
        } else {                          // fractional area of layer l, whole layers, fraction of the last one
          if (mapn) {
            const double dp = pe1[l + 1] - pe2[k];
            double fac1 = 1. + pl;
            const double fac2 = r3 * (1. + pl * fac1);
            fac1 = 0.5 * fac1;
            qsum = dp * (q4(2, l) + (q4(4, l) + q4(3, l) - q4(2, l)) * fac1 - q4(4, l) * fac2);
          } else
          qsum = (pe1[l + 1] - pe2[k]) * (q4(2, l) + 0.5 * (q4(4, l) + q4(3, l) - q4(2, l)) * (1. + pl) - q4(4, l) * (r3 * (1. + pl * (1. + pl))));
          for (int m = l + 1; m <= km; m++) {
            if (pe2[k + 1] > pe1[m + 1]) qsum = qsum + dp1[m] * q4(1, m);
            else {
              const double dp = pe2[k + 1] - pe1[m], esl = dp / dp1[m];
              if (mapn) { const double fac1 = 0.5 * esl, fac2 = 1. - r23 * esl; qsum = qsum + dp * (q4(2, m) + fac1 * (q4(3, m) - q4(2, m) + q4(4, m) * fac2)); }
              else
              qsum = qsum + dp * (q4(2, m) + 0.5 * esl * (q4(3, m) - q4(2, m) + q4(4, m) * (1. - r23 * esl)));
              k0 = m;
              break;
            }
          }
          have_sum = true;
        }
      }
    }
    if (!done) q2[k] = qsum / (dp2 ? dp2[k] : (pe2[k + 1] - pe2[k]));
  }
}

// map_scalar (scalar = true) / map1_ppm (false) / map1_q2 (scalar = true, dp2 given) of one column, in place on q(1:km)
int remap_field(int km, const std::vector<double>& pe1, const std::vector<double>& pe2, std::vector<double>& q, double qs, int iv, int kord,
                double qmin, bool scalar, const double* dp2 = nullptr, bool mapn = false) {
  A4 q4(km);
  std::vector<double> dp1(km + 2, 0.);
  for (int k = 1; k <= km; k++) { dp1[k] = pe1[k + 1] - pe1[k]; q4(1, k) = q[k]; }
  // kord > 7: scalar_profile / cs_profile, else ppm_profile (:86-90, 181-185, 394-398)
  // (the callers pass abs(kord); mapn_tracer has no ppm_profile branch, :262-273: it calls scalar_profile whatever kord is, whose
  //  interior `select case (abs(kord))` treats 0..8 alike, :756, as do the tests abs(kord) > 9 / >= 14 / <= 13)
  int ak = std::abs(kord);
  if (mapn && ak <= 7) ak = 8;
  const int rc = ak > 7 ? profile(qs, q4, dp1, km, iv, ak, qmin, scalar) : ppm_profile(q4, dp1, km, iv, ak);
  if (rc) return rc;
  map_column(km, pe1, q4, dp1, km, pe2, q, dp2, mapn);
  return 0;
}

}  // namespace

// fillz (fv_fill.F90:34-139; the default branch, DEV_GFS_PHYS not defined) for one column of one tracer: negative mixing ratios
// borrow mass from the layer above, then below; columns that needed it get the non-local rescaling of :113-135.  q, dp: 1..km
void fillz_column(int km, std::vector<double>& q, const std::vector<double>& dp) {
  auto pos = [](double x) { return x > 0. ? x : 0.; };   // max(0., x)
  if (q[1] < 0.) { q[2] = q[2] + q[1] * dp[1] / dp[2]; q[1] = 0.; }
  bool zfix = false;
  for (int k = 2; k <= km - 1; k++) {
    if (q[k] < 0.) {
      zfix = true;
      if (q[k - 1] > 0.) {   // borrow from above
        const double dq = std::min(q[k - 1] * dp[k - 1], -q[k] * dp[k]);
        q[k - 1] = q[k - 1] - dq / dp[k - 1];
        q[k] = q[k] + dq / dp[k];
      }
      if (q[k] < 0. && q[k + 1] > 0.) {   // borrow from below
        const double dq = std::min(q[k + 1] * dp[k + 1], -q[k] * dp[k]);
        q[k + 1] = q[k + 1] - dq / dp[k + 1];
        q[k] = q[k] + dq / dp[k];
      }
    }
  }
  if (q[km] < 0. && q[km - 1] > 0.) {
    zfix = true;
    const double qup = q[km - 1] * dp[km - 1], qly = -q[km] * dp[km];
    const double dup = std::min(qly, qup);
    q[km - 1] = q[km - 1] - dup / dp[km - 1];
    q[km] = q[km] + dup / dp[km];
  }
  if (zfix) {
    std::vector<double> dm(km + 1, 0.);
    double sum0 = 0.;
    for (int k = 2; k <= km; k++) { dm[k] = q[k] * dp[k]; sum0 = sum0 + dm[k]; }
    if (sum0 > 0.) {
      double sum1 = 0.;
      for (int k = 2; k <= km; k++) sum1 = sum1 + pos(dm[k]);
      const double fac = sum0 / sum1;
      for (int k = 2; k <= km; k++) q[k] = pos(fac * dm[k] / dp[k]);
    }
  }
}

// stand-alone column operator on FV3_WORK_Q (compute domain): pe1 = FV3_PE of the context, pe2 = ak + bk * pe1(km+1)
// mode 0: map_scalar (iv, kord, qmin), 1: map1_ppm (iv, kord; qs = FV3_WS when iv = -2), 2: map1_q2 (iv = 0, kord)
int remap_work_q(V3 q, V2 ws, double* pe, const std::vector<double>& ak, const std::vector<double>& bk, const fv3_flags_t& f, const Bd& bd,
                 int mode, int iv, int kord, double qmin) {
  const int km = bd.npz, is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
  const size_t nip = ie - is + 3;
  auto PE = [&](int i, int k, int j) -> double& { return pe[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (km + 1)]; };
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(max : bad)
  for (int j = js; j <= je; j++)
    for (int i = is; i <= ie; i++) {
      std::vector<double> pe1(km + 2), pe2(km + 2), col(km + 2), dp2(km + 2);
      for (int k = 1; k <= km + 1; k++) { pe1[k] = PE(i, k, j); pe2[k] = ak[k - 1] + bk[k - 1] * PE(i, km + 1, j); }
      pe2[1] = f.ptop; pe2[km + 1] = pe1[km + 1];
      for (int k = 1; k <= km; k++) { col[k] = q(i, j, k); dp2[k] = pe2[k + 1] - pe2[k]; }
      const int rc = remap_field(km, pe1, pe2, col, iv == -2 ? ws(i, j) : 0., iv, kord, qmin, mode != 1, mode == 2 ? dp2.data() : nullptr);
      if (rc) bad = 1;
      for (int k = 1; k <= km; k++) q(i, j, k) = col[k];
    }
  return bad ? -2 : 0;
}

// fillz of q with the thicknesses dp on the compute domain (fv3o_fillz)
void fillz(V3 q, V3 dp, const Bd& bd) {
  const int km = bd.npz;
#pragma omp parallel for schedule(static)
  for (int j = bd.js; j <= bd.je; j++)
    for (int i = bd.is; i <= bd.ie; i++) {
      std::vector<double> col(km + 2), d(km + 2);
      for (int k = 1; k <= km; k++) { col[k] = q(i, j, k); d[k] = dp(i, j, k); }
      fillz_column(km, col, d);
      for (int k = 1; k <= km; k++) q(i, j, k) = col[k];
    }
}

// fv_mapz.F90:56-845 on one face; fill: flagstruct%fill (fillz after each tracer, :391 / fv_operators.F90:337); use_tracer: the number of tracers (F.qtr) remapped with kord_tr (map1_q2 for nq <= 5, :398-408; mapn_tracer for nq > 5, :390-393).
int lagrangian_to_eulerian(const L2EFields& F, const std::vector<double>& ak, const std::vector<double>& bk, const fv3_flags_t& f, const Bd& bd,
                           int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum, double r_vir, int fill) {
  if (f.moist_kappa || kord_wz < 0) return -2;
  if (sphum >= use_tracer) return -1;
  const int km = bd.npz, is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
  const bool hydrostatic = f.hydrostatic != 0;
  const double akap = f.kappa, cv_air = f.cp_air - f.rdgas, k1k = f.rdgas / cv_air, rrg = -f.rdgas / f.grav, ptop = f.ptop;
  const double t_min = 184.;   // fv_mapz.F90:43
  const size_t nip = ie - is + 3, nie = ie - is + 1;
  double *pe = F.pe, *peln = F.peln;
  auto PE = [&](int i, int k, int j) -> double& { return pe[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (km + 1)]; };
  auto PELN = [&](int i, int k, int j) -> double& { return peln[(i - is) + (size_t)(k - 1) * nie + (size_t)(j - js) * nie * (km + 1)]; };
  V3 pt = F.pt, delp = F.delp, delz = F.delz, w = F.w, u = F.u, v = F.v, pk = F.pk, pkz = F.pkz, omga = F.omga;
  if ((int)F.qtr.size() < use_tracer) return -1;
  V2 ws = F.ws;
  std::vector<double> pe4((size_t)nie * (je - js + 1) * (km + 1), 0.);   // the new interface pressures, stored until every row has used the old ones
  auto PE4 = [&](int i, int j, int k) -> double& { return pe4[(i - is) + (size_t)(j - js) * nie + (size_t)(k - 1) * nie * (je - js + 1)]; };
  int bad = 0;
  // the rows are independent (pe of rows j - 1, j is read, never written, inside the loop; u of row j is written by row j only)
#pragma omp parallel for schedule(static) reduction(max : bad)
  for (int j = js; j <= je + 1; j++) {
    std::vector<double> pe1(km + 2), pe2(km + 2), pn1(km + 2), pn2(km + 2), pk2(km + 2), dp2(km + 2), col(km + 2), pe0(km + 2), pe3(km + 2);
    if (j != je + 1) {
      for (int i = is; i <= ie; i++) {
        for (int k = 1; k <= km + 1; k++) pe1[k] = PE(i, k, j);
        pe2[1] = ptop; pe2[km + 1] = PE(i, km + 1, j);
        // 0) temperature (:200-230): theta_v -> T_v when kord_tm < 0
        if (kord_tm < 0) {
          if (hydrostatic) for (int k = 1; k <= km; k++) pt(i, j, k) = pt(i, j, k) * (pk(i, j, k + 1) - pk(i, j, k)) / (akap * (PELN(i, k + 1, j) - PELN(i, k, j)));
          else for (int k = 1; k <= km; k++) pt(i, j, k) = pt(i, j, k) * std::exp(k1k * std::log(rrg * delp(i, j, k) / delz(i, j, k) * pt(i, j, k)));
        }
        if (!hydrostatic) for (int k = 1; k <= km; k++) delz(i, j, k) = -delz(i, j, k) / delp(i, j, k);   // :297-303
        for (int k = 2; k <= km; k++) pe2[k] = ak[k - 1] + bk[k - 1] * PE(i, km + 1, j);                   // :313-317
        for (int k = 1; k <= km; k++) { dp2[k] = pe2[k + 1] - pe2[k]; delp(i, j, k) = dp2[k]; }           // :318-331
        for (int k = 1; k <= km + 1; k++) pn1[k] = PELN(i, k, j);
        pn2[1] = pn1[1]; pn2[km + 1] = pn1[km + 1]; pk2[1] = pk(i, j, 1); pk2[km + 1] = pk(i, j, km + 1);
        for (int k = 2; k <= km; k++) { pn2[k] = std::log(pe2[k]); pk2[k] = std::exp(akap * pn2[k]); }    // :350-355
        // 1) T_v in log(p) (kord_tm < 0) or theta_v in p (:373-386)
        for (int k = 1; k <= km; k++) col[k] = pt(i, j, k);
        int rc = kord_tm < 0 ? remap_field(km, pn1, pn2, col, 0., 1, std::abs(kord_tm), t_min, true)
                             : remap_field(km, pe1, pe2, col, 0., 1, std::abs(kord_tm), 0., false);
        if (rc) bad = 1;
        for (int k = 1; k <= km; k++) pt(i, j, k) = col[k];
        // 2) the tracer (:395-408, map1_q2, no fillz)
        for (int iq = 0; iq < use_tracer; iq++) {
          V3 qtr = F.qtr[iq];
          for (int k = 1; k <= km; k++) col[k] = qtr(i, j, k);
          rc = remap_field(km, pe1, pe2, col, 0., 0, kord_tr, 0., true, dp2.data(), use_tracer > 5);   // nq > 5: mapn_tracer (:390-393)
          if (rc) bad = 1;
          if (fill && !rc) fillz_column(km, col, dp2);
          for (int k = 1; k <= km; k++) qtr(i, j, k) = col[k];
        }
        // 3) w with the lower boundary condition ws, then delz (:411-433)
        if (!hydrostatic) {
          for (int k = 1; k <= km; k++) col[k] = w(i, j, k);
          rc = remap_field(km, pe1, pe2, col, ws(i, j), -2, std::abs(kord_wz), 0., false);
          if (rc) bad = 1;
          for (int k = 1; k <= km; k++) w(i, j, k) = col[k];
          for (int k = 1; k <= km; k++) col[k] = delz(i, j, k);
          rc = remap_field(km, pe1, pe2, col, 0., 1, std::abs(kord_tm), 0., false);
          if (rc) bad = 1;
          for (int k = 1; k <= km; k++) delz(i, j, k) = -col[k] * dp2[k];
        }
        for (int k = 1; k <= km + 1; k++) pk(i, j, k) = pk2[k];                                            // :436-440
        if (last_step) { pe3[1] = 0.; for (int k = 2; k <= km + 1; k++) pe3[k] = omga(i, j, k - 1); }      // :442-453
        for (int k = 1; k <= km + 1; k++) { pe0[k] = PELN(i, k, j); PELN(i, k, j) = pn2[k]; }              // :455-460
        // 3.2) pkz (:463-506)
        if (hydrostatic) for (int k = 1; k <= km; k++) pkz(i, j, k) = (pk2[k + 1] - pk2[k]) / (akap * (PELN(i, k + 1, j) - PELN(i, k, j)));
        else if (kord_tm < 0) for (int k = 1; k <= km; k++) pkz(i, j, k) = std::exp(akap * std::log(rrg * delp(i, j, k) / delz(i, j, k) * pt(i, j, k)));
        else for (int k = 1; k <= km; k++) pkz(i, j, k) = std::exp(k1k * std::log(rrg * delp(i, j, k) / delz(i, j, k) * pt(i, j, k)));
        if (kord_tm > 0) for (int k = 1; k <= km; k++) pt(i, j, k) = pt(i, j, k) * pkz(i, j, k);
        // 3.3) omega to the new layer centres (:509-526)
        if (last_step) {
          for (int k = 1; k <= km; k++) dp2[k] = 0.5 * (PELN(i, k, j) + PELN(i, k + 1, j));
          int k_next = 1;
          for (int n = 1; n <= km; n++) {
            const int kp = k_next;
            for (int k = kp; k <= km; k++) {
              if (dp2[n] <= pe0[k + 1] && dp2[n] >= pe0[k]) {
                omga(i, j, n) = pe3[k] + (pe3[k + 1] - pe3[k]) * (dp2[n] - pe0[k]) / (pe0[k + 1] - pe0[k]);
                k_next = k;
                break;
              }
            }
          }
        }
        for (int k = 1; k <= km; k++) PE4(i, j, k) = pe2[k + 1];   // :652-656
      }
    }
    // 4.1) u on the south faces of row j (:535-552): pressures averaged with row j - 1
    for (int i = is; i <= ie; i++) {
      pe0[1] = PE(i, 1, j);
      for (int k = 2; k <= km + 1; k++) pe0[k] = 0.5 * (PE(i, k, j - 1) + PE(i, k, j));
      for (int k = 1; k <= km + 1; k++) { const double bkh = 0.5 * bk[k - 1]; pe3[k] = ak[k - 1] + bkh * (PE(i, km + 1, j - 1) + PE(i, km + 1, j)); }
      for (int k = 1; k <= km; k++) col[k] = u(i, j, k);
      if (remap_field(km, pe0, pe3, col, 0., -1, kord_mt, 0., false)) bad = 1;
      for (int k = 1; k <= km; k++) u(i, j, k) = col[k];
    }
    // 4.2) v on the west faces (:556-571)
    if (j < je + 1) {
      for (int i = is; i <= ie + 1; i++) {
        pe0[1] = PE(i, 1, j);
        pe3[1] = ak[0];
        for (int k = 2; k <= km + 1; k++) {
          const double bkh = 0.5 * bk[k - 1];
          pe0[k] = 0.5 * (PE(i - 1, k, j) + PE(i, k, j));
          pe3[k] = ak[k - 1] + bkh * (PE(i - 1, km + 1, j) + PE(i, km + 1, j));
        }
        for (int k = 1; k <= km; k++) col[k] = v(i, j, k);
        if (remap_field(km, pe0, pe3, col, 0., -1, kord_mt, 0., false)) bad = 1;
        for (int k = 1; k <= km; k++) v(i, j, k) = col[k];
      }
    }
  }
  if (bad) return -2;
  // 6) the new interface pressures (:661-668)
  for (int k = 2; k <= km; k++)
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++) PE(i, k, j) = PE4(i, j, k - 1);
  // 9) last step: T_v -> T (:792-822, dtmp = 0; no condensates: `.not. use_cond`, `.not. adiabatic` when a specific humidity is
  //    named, the identity otherwise); not the last step: back to theta_v for dyn_core (:833-843)
  if (last_step && sphum >= 0) {
    V3 qv = F.qtr[sphum];
    for (int k = 1; k <= km; k++)
      for (int j = js; j <= je; j++)
        for (int i = is; i <= ie; i++) pt(i, j, k) = (pt(i, j, k) + 0. / (hydrostatic ? f.cp_air : cv_air) * pkz(i, j, k)) / (1. + r_vir * qv(i, j, k));
  }
  if (!last_step)
    for (int k = 1; k <= km; k++)
      for (int j = js; j <= je; j++)
        for (int i = is; i <= ie; i++) pt(i, j, k) = pt(i, j, k) / pkz(i, j, k);
  return 0;
}

}  // namespace fv3o
