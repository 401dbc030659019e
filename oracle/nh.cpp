// TEST INFRASTRUCTURE ONLY -- CPU oracle (see fv3_oracle.hpp header).
// Restates the non-hydrostatic column solvers and pressure-gradient stencils of the reference:
//   model/nh_utils.F90: update_dz_c (:59-201), update_dz_d (:204-321), Riem_Solver_c (:323-480),
//                       SIM1_solver (:1277-1394), SIM_solver (:1396-1537), edge_profile (:1590-1696)
//   model/nh_core.F90:  Riem_Solver3 (:47-241)
//   model/dyn_core.F90: p_grad_c (:1635-1694), nh_p_grad (:1697-1792), pk3_halo (:1395-1447),
//                       pe_halo (:1498-1526)
// Rayleigh w damping (fast_tau_w_sec) is not restated (off in both contract flag sets).
#include "fv3_oracle.hpp"

namespace fv3o {

static const double dz_min = 2.;   // nh_utils.F90:46-50 (no DZ_MIN_6)
static const double r3 = 1. / 3.;

// nh_utils.F90:59-201
void update_dz_c(int is, int ie, int js, int je, int km, int ng, double dt, const double* dp0, V2 zs, V2 area,
                 V3 ut, V3 vt, V3 gz, V2 ws, const Bd& bd) {
  const double rdt = 1. / dt;
  const double top_ratio = dp0[0] / (dp0[0] + dp0[1]);
  const double bot_ratio = dp0[km - 1] / (dp0[km - 2] + dp0[km - 1]);
  const int is1 = is - 1, js1 = js - 1, ie1 = ie + 1, je1 = je + 1, ie2 = ie + 2, je2 = je + 2;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= km + 1; k++) {
    L2 gz2(is - ng, ie + ng, js - ng, je + ng);
    L2 xfx(is - 1, ie + 2, js - 1, je + 1), fx(is - 1, ie + 2, js - 1, je + 1);
    L2 yfx(is - 1, ie + 1, js - 1, je + 2), fy(is - 1, ie + 1, js - 1, je + 2);
    if (k == 1) {
      for (int j = js1; j <= je1; j++) for (int i = is1; i <= ie2; i++) xfx(i, j) = ut(i, j, 1) + (ut(i, j, 1) - ut(i, j, 2)) * top_ratio;
      for (int j = js1; j <= je2; j++) for (int i = is1; i <= ie1; i++) yfx(i, j) = vt(i, j, 1) + (vt(i, j, 1) - vt(i, j, 2)) * top_ratio;
    } else if (k == km + 1) {
      for (int j = js1; j <= je1; j++) for (int i = is1; i <= ie2; i++) xfx(i, j) = ut(i, j, km) + (ut(i, j, km) - ut(i, j, km - 1)) * bot_ratio;
      for (int j = js1; j <= je2; j++) for (int i = is1; i <= ie1; i++) yfx(i, j) = vt(i, j, km) + (vt(i, j, km) - vt(i, j, km - 1)) * bot_ratio;
    } else {
      const double int_ratio = 1. / (dp0[k - 2] + dp0[k - 1]);
      for (int j = js1; j <= je1; j++) for (int i = is1; i <= ie2; i++) xfx(i, j) = (dp0[k - 1] * ut(i, j, k - 1) + dp0[k - 2] * ut(i, j, k)) * int_ratio;
      for (int j = js1; j <= je2; j++) for (int i = is1; i <= ie1; i++) yfx(i, j) = (dp0[k - 1] * vt(i, j, k - 1) + dp0[k - 2] * vt(i, j, k)) * int_ratio;
    }
    for (int j = js - ng; j <= je + ng; j++) for (int i = is - ng; i <= ie + ng; i++) gz2(i, j) = gz(i, j, k);
    if (bd.grid_type < 3) fill_4corners(gz2, 1, bd);
    for (int j = js1; j <= je1; j++)
      for (int i = is1; i <= ie2; i++) {
        fx(i, j) = (xfx(i, j) > 0.) ? gz2(i - 1, j) : gz2(i, j);
        fx(i, j) = xfx(i, j) * fx(i, j);
      }
    if (bd.grid_type < 3) fill_4corners(gz2, 2, bd);
    for (int j = js1; j <= je2; j++)
      for (int i = is1; i <= ie1; i++) {
        fy(i, j) = (yfx(i, j) > 0.) ? gz2(i, j - 1) : gz2(i, j);
        fy(i, j) = yfx(i, j) * fy(i, j);
      }
    for (int j = js1; j <= je1; j++)
      for (int i = is1; i <= ie1; i++)
        gz(i, j, k) = (gz2(i, j) * area(i, j) + fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) /
                      (area(i, j) + xfx(i, j) - xfx(i + 1, j) + yfx(i, j) - yfx(i, j + 1));
  }
#pragma omp parallel for schedule(static)
  for (int j = js1; j <= je1; j++) {
    for (int i = is1; i <= ie1; i++) ws(i, j) = (zs(i, j) - gz(i, j, km + 1)) * rdt;
    for (int k = km; k >= 1; k--)
      for (int i = is1; i <= ie1; i++) gz(i, j, k) = std::max(gz(i, j, k), gz(i, j, k + 1) + dz_min);
  }
}

// nh_utils.F90:1590-1696 (uniform_grid = .false., limiter = 0 as called from update_dz_d)
static void edge_profile(V3 q1, V3 q2, V3 q1e, V3 q2e, int i1, int i2, int j, int km, const double* dp0) {
  L2 qe1(i1, i2, 1, km + 1), qe2(i1, i2, 1, km + 1), gam(i1, i2, 1, km + 1);
  double g0 = dp0[1] / dp0[0];
  double xt1 = 2. * g0 * (g0 + 1.);
  double bet = g0 * (g0 + 0.5);
  for (int i = i1; i <= i2; i++) {
    qe1(i, 1) = (xt1 * q1(i, j, 1) + q1(i, j, 2)) / bet;
    qe2(i, 1) = (xt1 * q2(i, j, 1) + q2(i, j, 2)) / bet;
    gam(i, 1) = (1. + g0 * (g0 + 1.5)) / bet;
  }
  double gk = 0.;
  for (int k = 2; k <= km; k++) {
    gk = dp0[k - 2] / dp0[k - 1];
    for (int i = i1; i <= i2; i++) {
      bet = 2. + 2. * gk - gam(i, k - 1);
      qe1(i, k) = (3. * (q1(i, j, k - 1) + gk * q1(i, j, k)) - qe1(i, k - 1)) / bet;
      qe2(i, k) = (3. * (q2(i, j, k - 1) + gk * q2(i, j, k)) - qe2(i, k - 1)) / bet;
      gam(i, k) = gk / bet;
    }
  }
  double a_bot = 1. + gk * (gk + 1.5);
  xt1 = 2. * gk * (gk + 1.);
  for (int i = i1; i <= i2; i++) {
    double xt2 = gk * (gk + 0.5) - a_bot * gam(i, km);
    qe1(i, km + 1) = (xt1 * q1(i, j, km) + q1(i, j, km - 1) - a_bot * qe1(i, km)) / xt2;
    qe2(i, km + 1) = (xt1 * q2(i, j, km) + q2(i, j, km - 1) - a_bot * qe2(i, km)) / xt2;
  }
  for (int k = km; k >= 1; k--)
    for (int i = i1; i <= i2; i++) {
      qe1(i, k) = qe1(i, k) - gam(i, k) * qe1(i, k + 1);
      qe2(i, k) = qe2(i, k) - gam(i, k) * qe2(i, k + 1);
    }
  for (int k = 1; k <= km + 1; k++)
    for (int i = i1; i <= i2; i++) { q1e(i, j, k) = qe1(i, k); q2e(i, j, k) = qe2(i, k); }
}

// nh_utils.F90:204-321
void update_dz_d(int* ndif, double* damp, int hord, int is, int ie, int js, int je, int km, int ng, int npx,
                 int npy, const double* dp0, V2 zs, V3 zh, V3 crx, V3 cry, V3 xfx, V3 yfx, V2 ws, double rdt,
                 const Grid& g, const Bd& bd, double lim_fac) {
  damp[km] = damp[km - 1];   // damp(km+1) = damp(km)
  ndif[km] = ndif[km - 1];
  const int isd = is - ng, ied = ie + ng, jsd = js - ng, jed = je + ng;
  const size_t nb1 = (size_t)(ie + 1 - is + 1) * (jed - jsd + 1) * (km + 1), nb3 = (size_t)(ied - isd + 1) * (je + 1 - js + 1) * (km + 1);
  LRaw b1(nb1), b2(nb1), b3(nb3), b4(nb3);
  V3 crx_adv(b1.data(), is, ie + 1, jsd, jed), xfx_adv(b2.data(), is, ie + 1, jsd, jed);
  V3 cry_adv(b3.data(), isd, ied, js, je + 1), yfx_adv(b4.data(), isd, ied, js, je + 1);
#pragma omp parallel for schedule(static)
  for (int j = jsd; j <= jed; j++) {
    edge_profile(crx, xfx, crx_adv, xfx_adv, is, ie + 1, j, km, dp0);
    if (j <= je + 1 && j >= js) edge_profile(cry, yfx, cry_adv, yfx_adv, isd, ied, j, km, dp0);
  }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= km + 1; k++) {
    L2 fx(is, ie + 1, js, je), fy(is, ie, js, je + 1);
    L2 fx2(isd, ied + 1, jsd, jed), fy2(isd, ied, jsd, jed + 1), wk2(isd, ied, jsd, jed), z2(isd, ied, jsd, jed);
    L2 ra_x(is, ie, jsd, jed), ra_y(isd, ied, js, je);
    for (int j = jsd; j <= jed; j++) for (int i = is; i <= ie; i++) ra_x(i, j) = g.area(i, j) + xfx_adv(i, j, k) - xfx_adv(i + 1, j, k);
    for (int j = js; j <= je; j++) for (int i = isd; i <= ied; i++) ra_y(i, j) = g.area(i, j) + yfx_adv(i, j, k) - yfx_adv(i, j + 1, k);
    if (damp[k - 1] > 1.E-5) {
      for (int j = jsd; j <= jed; j++) for (int i = isd; i <= ied; i++) z2(i, j) = zh(i, j, k);
      fv_tp_2d(z2, crx_adv.k(k), cry_adv.k(k), npx, npy, hord, fx, fy, xfx_adv.k(k), yfx_adv.k(k), g, bd, ra_x, ra_y,
               lim_fac, nullptr, nullptr, nullptr, false, 0, 0.);
      del6_vt_flux(ndif[k - 1], npx, npy, damp[k - 1], z2, wk2, fx2, fy2, g, bd);
      for (int j = js; j <= je; j++)
        for (int i = is; i <= ie; i++)
          zh(i, j, k) = (z2(i, j) * g.area(i, j) + fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) /
                            (ra_x(i, j) + ra_y(i, j) - g.area(i, j)) +
                        (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * g.rarea(i, j);
    } else {
      fv_tp_2d(zh.k(k), crx_adv.k(k), cry_adv.k(k), npx, npy, hord, fx, fy, xfx_adv.k(k), yfx_adv.k(k), g, bd, ra_x,
               ra_y, lim_fac, nullptr, nullptr, nullptr, false, 0, 0.);
      for (int j = js; j <= je; j++)
        for (int i = is; i <= ie; i++)
          zh(i, j, k) = (zh(i, j, k) * g.area(i, j) + fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) /
                        (ra_x(i, j) + ra_y(i, j) - g.area(i, j));
    }
  }
#pragma omp parallel for schedule(static)
  for (int j = js; j <= je; j++) {
    for (int i = is; i <= ie; i++) ws(i, j) = (zs(i, j) - zh(i, j, km + 1)) * rdt;
    for (int k = km; k >= 1; k--)
      for (int i = is; i <= ie; i++) zh(i, j, k) = std::max(zh(i, j, k), zh(i, j, k + 1) + dz_min);
  }
}

// (i,k) slab helper, 1-based k
struct S2 {
  std::vector<double> b; int i0, ni;
  S2(int ilo, int ihi, int nk) : b((size_t)(ihi - ilo + 1) * nk), i0(ilo), ni(ihi - ilo + 1) {}
  inline double& operator()(int i, int k) { return b[(i - i0) + (size_t)(k - 1) * ni]; }
};

// nh_utils.F90:1277-1394
static void sim1_solver(double dt, int is, int ie, int km, double rgas, S2& gm2, S2& cp2, S2& pe, S2& dm2, S2& pm2,
                        S2& pem, S2& w2, S2& dz2, S2& pt2, const double* ws /*ws[i-is]*/, double p_fac,
                        const double* rff = nullptr, int k_rf = 0) {
  S2 aa(is, ie, km), bb(is, ie, km), dd(is, ie, km), w1(is, ie, km), g_rat(is, ie, km), gam(is, ie, km), pp(is, ie, km + 1);
  std::vector<double> p1v(ie - is + 1), betv(ie - is + 1);
  auto p1 = [&](int i) -> double& { return p1v[i - is]; };
  auto bet = [&](int i) -> double& { return betv[i - is]; };
  const double t1g = 2. * dt * dt, rdt = 1. / dt;
  for (int k = 1; k <= km; k++)
    for (int i = is; i <= ie; i++) {
      pe(i, k) = std::exp(gm2(i, k) * std::log(-dm2(i, k) / dz2(i, k) * rgas * pt2(i, k))) - pm2(i, k);
      w1(i, k) = w2(i, k);
    }
  for (int k = 1; k <= km - 1; k++)
    for (int i = is; i <= ie; i++) {
      g_rat(i, k) = dm2(i, k) / dm2(i, k + 1);
      bb(i, k) = 2. * (1. + g_rat(i, k));
      dd(i, k) = 3. * (pe(i, k) + g_rat(i, k) * pe(i, k + 1));
    }
  for (int i = is; i <= ie; i++) {
    bet(i) = bb(i, 1);
    pp(i, 1) = 0.;
    pp(i, 2) = dd(i, 1) / bet(i);
    bb(i, km) = 2.;
    dd(i, km) = 3. * pe(i, km);
  }
  for (int k = 2; k <= km; k++)
    for (int i = is; i <= ie; i++) {
      gam(i, k) = g_rat(i, k - 1) / bet(i);
      bet(i) = bb(i, k) - gam(i, k);
      pp(i, k + 1) = (dd(i, k) - pp(i, k)) / bet(i);
    }
  for (int k = km; k >= 2; k--)
    for (int i = is; i <= ie; i++) pp(i, k) = pp(i, k) - gam(i, k) * pp(i, k + 1);
  // w-solver
  for (int k = 2; k <= km; k++)
    for (int i = is; i <= ie; i++)
      aa(i, k) = t1g * 0.5 * (gm2(i, k - 1) + gm2(i, k)) / (dz2(i, k - 1) + dz2(i, k)) * (pem(i, k));
  for (int i = is; i <= ie; i++) {
    bet(i) = dm2(i, 1) - aa(i, 2);
    w2(i, 1) = (dm2(i, 1) * w1(i, 1) + dt * pp(i, 2)) / bet(i);
  }
  for (int k = 2; k <= km - 1; k++)
    for (int i = is; i <= ie; i++) {
      gam(i, k) = aa(i, k) / bet(i);
      bet(i) = dm2(i, k) - (aa(i, k) + aa(i, k + 1) + aa(i, k) * gam(i, k));
      w2(i, k) = (dm2(i, k) * w1(i, k) + dt * (pp(i, k + 1) - pp(i, k)) - aa(i, k) * w2(i, k - 1)) / bet(i);
    }
  for (int i = is; i <= ie; i++) {
    p1(i) = t1g * gm2(i, km) / dz2(i, km) * (pem(i, km + 1));
    gam(i, km) = aa(i, km) / bet(i);
    bet(i) = dm2(i, km) - (aa(i, km) + p1(i) + aa(i, km) * gam(i, km));
    w2(i, km) = (dm2(i, km) * w1(i, km) + dt * (pp(i, km + 1) - pp(i, km)) - p1(i) * ws[i - is] - aa(i, km) * w2(i, km - 1)) / bet(i);
  }
  for (int k = km - 1; k >= 1; k--)
    for (int i = is; i <= ie; i++) w2(i, k) = w2(i, k) - gam(i, k + 1) * w2(i, k + 1);
  if (rff)   // Rayleigh damping of w, nh_utils.F90:1363-1371
    for (int k = 1; k <= k_rf; k++) for (int i = is; i <= ie; i++) w2(i, k) = w2(i, k) * rff[k - 1];
  for (int i = is; i <= ie; i++) pe(i, 1) = 0.;
  for (int k = 1; k <= km; k++)
    for (int i = is; i <= ie; i++) pe(i, k + 1) = pe(i, k) + dm2(i, k) * (w2(i, k) - w1(i, k)) * rdt;
  for (int i = is; i <= ie; i++) {
    p1(i) = (pe(i, km) + 2. * pe(i, km + 1)) * r3;
    dz2(i, km) = -dm2(i, km) * rgas * pt2(i, km) *
                 std::exp((cp2(i, km) - 1.) * std::log(std::max(p_fac * pm2(i, km), p1(i) + pm2(i, km))));
  }
  for (int k = km - 1; k >= 1; k--)
    for (int i = is; i <= ie; i++) {
      p1(i) = (pe(i, k) + bb(i, k) * pe(i, k + 1) + g_rat(i, k) * pe(i, k + 2)) * r3 - g_rat(i, k) * p1(i);
      dz2(i, k) = -dm2(i, k) * rgas * pt2(i, k) *
                  std::exp((cp2(i, k) - 1.) * std::log(std::max(p_fac * pm2(i, k), p1(i) + pm2(i, k))));
    }
}

// nh_utils.F90:1396-1537 (scale_m = 0)
static void sim_solver(double dt, int is, int ie, int km, double rgas, S2& gm2, S2& cp2, S2& pe2, S2& dm2, S2& pm2,
                       S2& pem, S2& w2, S2& dz2, S2& pt2, const double* ws, double alpha, double p_fac, double scale_m,
                       const double* rff = nullptr, int k_rf = 0) {
  S2 aa(is, ie, km), bb(is, ie, km), dd(is, ie, km), w1(is, ie, km), wk(is, ie, km), g_rat(is, ie, km), gam(is, ie, km), pp(is, ie, km + 1);
  std::vector<double> p1v(ie - is + 1), wk1v(ie - is + 1), betv(ie - is + 1);
  auto p1 = [&](int i) -> double& { return p1v[i - is]; };
  auto wk1 = [&](int i) -> double& { return wk1v[i - is]; };
  auto bet = [&](int i) -> double& { return betv[i - is]; };
  const double beta = 1. - alpha, ra = 1. / alpha, t2 = beta / alpha, t1g = 2. * (alpha * dt) * (alpha * dt), rdt = 1. / dt;
  for (int k = 1; k <= km; k++)
    for (int i = is; i <= ie; i++) {
      w1(i, k) = w2(i, k);
      pe2(i, k) = std::exp(gm2(i, k) * std::log(-dm2(i, k) / dz2(i, k) * rgas * pt2(i, k))) - pm2(i, k);
    }
  for (int k = 1; k <= km - 1; k++)
    for (int i = is; i <= ie; i++) {
      g_rat(i, k) = dm2(i, k) / dm2(i, k + 1);
      bb(i, k) = 2. * (1. + g_rat(i, k));
      dd(i, k) = 3. * (pe2(i, k) + g_rat(i, k) * pe2(i, k + 1));
    }
  for (int i = is; i <= ie; i++) {
    bet(i) = bb(i, 1);
    pp(i, 1) = 0.;
    pp(i, 2) = dd(i, 1) / bet(i);
    bb(i, km) = 2.;
    dd(i, km) = 3. * pe2(i, km);
  }
  for (int k = 2; k <= km; k++)
    for (int i = is; i <= ie; i++) {
      gam(i, k) = g_rat(i, k - 1) / bet(i);
      bet(i) = bb(i, k) - gam(i, k);
      pp(i, k + 1) = (dd(i, k) - pp(i, k)) / bet(i);
    }
  for (int k = km; k >= 2; k--)
    for (int i = is; i <= ie; i++) pp(i, k) = pp(i, k) - gam(i, k) * pp(i, k + 1);
  for (int k = 1; k <= km + 1; k++) for (int i = is; i <= ie; i++) pe2(i, k) = pem(i, k);
  for (int k = 2; k <= km; k++)
    for (int i = is; i <= ie; i++) {
      aa(i, k) = t1g * 0.5 * (gm2(i, k - 1) + gm2(i, k)) / (dz2(i, k - 1) + dz2(i, k)) * pe2(i, k);
      wk(i, k) = t2 * aa(i, k) * (w1(i, k - 1) - w1(i, k));
      aa(i, k) = aa(i, k) - scale_m * dm2(i, 1);
    }
  for (int i = is; i <= ie; i++) {
    bet(i) = dm2(i, 1) - aa(i, 2);
    w2(i, 1) = (dm2(i, 1) * w1(i, 1) + dt * pp(i, 2) + wk(i, 2)) / bet(i);
  }
  for (int k = 2; k <= km - 1; k++)
    for (int i = is; i <= ie; i++) {
      gam(i, k) = aa(i, k) / bet(i);
      bet(i) = dm2(i, k) - (aa(i, k) + aa(i, k + 1) + aa(i, k) * gam(i, k));
      w2(i, k) = (dm2(i, k) * w1(i, k) + dt * (pp(i, k + 1) - pp(i, k)) + wk(i, k + 1) - wk(i, k) - aa(i, k) * w2(i, k - 1)) / bet(i);
    }
  for (int i = is; i <= ie; i++) {
    wk1(i) = t1g * gm2(i, km) / dz2(i, km) * pe2(i, km + 1);
    gam(i, km) = aa(i, km) / bet(i);
    bet(i) = dm2(i, km) - (aa(i, km) + wk1(i) + aa(i, km) * gam(i, km));
    w2(i, km) = (dm2(i, km) * w1(i, km) + dt * (pp(i, km + 1) - pp(i, km)) - wk(i, km) +
                 wk1(i) * (t2 * w1(i, km) - ra * ws[i - is]) - aa(i, km) * w2(i, km - 1)) / bet(i);
  }
  for (int k = km - 1; k >= 1; k--)
    for (int i = is; i <= ie; i++) w2(i, k) = w2(i, k) - gam(i, k + 1) * w2(i, k + 1);
  if (rff)   // nh_utils.F90:1498-1506
    for (int k = 1; k <= k_rf; k++) for (int i = is; i <= ie; i++) w2(i, k) = w2(i, k) * rff[k - 1];
  for (int i = is; i <= ie; i++) pe2(i, 1) = 0.;
  for (int k = 1; k <= km; k++)
    for (int i = is; i <= ie; i++)
      pe2(i, k + 1) = pe2(i, k) + (dm2(i, k) * (w2(i, k) - w1(i, k)) * rdt - beta * (pp(i, k + 1) - pp(i, k))) * ra;
  for (int i = is; i <= ie; i++) {
    p1(i) = (pe2(i, km) + 2. * pe2(i, km + 1)) * r3;
    dz2(i, km) = -dm2(i, km) * rgas * pt2(i, km) *
                 std::exp((cp2(i, km) - 1.) * std::log(std::max(p_fac * pm2(i, km), p1(i) + pm2(i, km))));
  }
  for (int k = km - 1; k >= 1; k--)
    for (int i = is; i <= ie; i++) {
      p1(i) = (pe2(i, k) + bb(i, k) * pe2(i, k + 1) + g_rat(i, k) * pe2(i, k + 2)) * r3 - g_rat(i, k) * p1(i);
      dz2(i, k) = -dm2(i, k) * rgas * pt2(i, k) *
                  std::exp((cp2(i, k) - 1.) * std::log(std::max(p_fac * pm2(i, k), p1(i) + pm2(i, k))));
    }
  for (int k = 1; k <= km + 1; k++)
    for (int i = is; i <= ie; i++) pe2(i, k) = pe2(i, k) + beta * (pp(i, k) - pe2(i, k));
}

// nh_utils.F90:323-480
void riem_solver_c(int ms, double dt, int is, int ie, int js, int je, int km, int ng, double akap, V3 cappa,
                   double cp, double ptop, V2 hs, V3 w3, V3 pt, V3 q_con, V3 delp, V3 gz, V3 pef, V2 ws,
                   double p_fac, double a_imp, bool use_cond, bool moist_kappa, const Consts& c) {
  (void)ms; (void)ng; (void)cp;
  const double rgrav = 1. / c.grav;
  const int is1 = is - 1, ie1 = ie + 1;
#pragma omp parallel for schedule(static)
  for (int j = js - 1; j <= je + 1; j++) {
    S2 dm(is1, ie1, km), dz2(is1, ie1, km), w2(is1, ie1, km), pm2(is1, ie1, km), gm2(is1, ie1, km), cp2(is1, ie1, km), pt2(is1, ie1, km);
    S2 pem(is1, ie1, km + 1), pe2(is1, ie1, km + 1), peg(is1, ie1, km + 1);
    for (int k = 1; k <= km; k++) for (int i = is1; i <= ie1; i++) dm(i, k) = delp(i, j, k);
    for (int i = is1; i <= ie1; i++) { pef(i, j, 1) = ptop; pem(i, 1) = ptop; if (use_cond) peg(i, 1) = ptop; }
    for (int k = 2; k <= km + 1; k++)
      for (int i = is1; i <= ie1; i++) {
        pem(i, k) = pem(i, k - 1) + dm(i, k - 1);
        if (use_cond) peg(i, k) = peg(i, k - 1) + dm(i, k - 1) * (1. - q_con(i, j, k - 1));
      }
    for (int k = 1; k <= km; k++)
      for (int i = is1; i <= ie1; i++) {
        dz2(i, k) = gz(i, j, k + 1) - gz(i, j, k);
        if (use_cond) pm2(i, k) = (peg(i, k + 1) - peg(i, k)) / std::log(peg(i, k + 1) / peg(i, k));
        else pm2(i, k) = dm(i, k) / std::log(pem(i, k + 1) / pem(i, k));
        cp2(i, k) = (use_cond && moist_kappa) ? cappa(i, j, k) : akap;
        gm2(i, k) = 1. / (1. - cp2(i, k));
        dm(i, k) = dm(i, k) * rgrav;
        w2(i, k) = w3(i, j, k);
        pt2(i, k) = pt(i, j, k);
      }
    std::vector<double> wsr(ie1 - is1 + 1);
    for (int i = is1; i <= ie1; i++) wsr[i - is1] = ws(i, j);
    // a_imp > 0.5 -> SIM1_solver (nh_utils.F90:456-458); other solvers out of contract
    sim1_solver(dt, is1, ie1, km, c.rdgas, gm2, cp2, pe2, dm, pm2, pem, w2, dz2, pt2, wsr.data(), p_fac, c.rff, c.k_rf);
    (void)a_imp;
    for (int k = 2; k <= km + 1; k++) for (int i = is1; i <= ie1; i++) pef(i, j, k) = pe2(i, k) + pem(i, k);
    for (int i = is1; i <= ie1; i++) gz(i, j, km + 1) = hs(i, j);
    for (int k = km; k >= 1; k--) for (int i = is1; i <= ie1; i++) gz(i, j, k) = gz(i, j, k + 1) - dz2(i, k) * c.grav;
  }
}

// nh_core.F90:47-241
void riem_solver3(int ms, double dt, int is, int ie, int js, int je, int km, int ng, int isd, int ied, int jsd,
                  int jed, double akap, V3 cappa, double cp, double ptop, V2 zs, V3 q_con, V3 w, V3 delz, V3 pt,
                  V3 delp, V3 zh, double* pe, V3 ppe, V3 pk3, V3 pk, double* peln, V2 ws, double p_fac, double a_imp,
                  bool use_logp, bool use_cond, bool moist_kappa, bool last_call, bool fp_out, const Consts& c) {
  (void)ms; (void)ng; (void)isd; (void)ied; (void)jsd; (void)jed; (void)cp;
  const double rgrav = 1. / c.grav;
  const double peln1 = std::log(ptop);
  const double ptk = std::exp(akap * peln1);
  const int nie = ie - is + 1;         // peln (is:ie, km+1, js:je)
  const int nip = ie + 1 - (is - 1) + 1;  // pe (is-1:ie+1, km+1, js-1:je+1)
#pragma omp parallel for schedule(static)
  for (int j = js; j <= je; j++) {
    S2 dm(is, ie, km), dz2(is, ie, km), pm2(is, ie, km), w2(is, ie, km), gm2(is, ie, km), cp2(is, ie, km), pt2(is, ie, km);
    S2 pem(is, ie, km + 1), pe2(is, ie, km + 1), peln2(is, ie, km + 1), peg(is, ie, km + 1), pelng(is, ie, km + 1);
    for (int k = 1; k <= km; k++)
      for (int i = is; i <= ie; i++) { dm(i, k) = delp(i, j, k); cp2(i, k) = moist_kappa ? cappa(i, j, k) : akap; }
    for (int i = is; i <= ie; i++) {
      pem(i, 1) = ptop; peln2(i, 1) = peln1; pk3(i, j, 1) = ptk;
      if (use_cond) { peg(i, 1) = ptop; pelng(i, 1) = peln1; }
    }
    for (int k = 2; k <= km + 1; k++)
      for (int i = is; i <= ie; i++) {
        pem(i, k) = pem(i, k - 1) + dm(i, k - 1);
        peln2(i, k) = std::log(pem(i, k));
        if (use_cond) {
          peg(i, k) = peg(i, k - 1) + dm(i, k - 1) * (1. - q_con(i, j, k - 1));
          pelng(i, k) = std::log(peg(i, k));
        }
        pk3(i, j, k) = std::exp(akap * peln2(i, k));
      }
    for (int k = 1; k <= km; k++)
      for (int i = is; i <= ie; i++) {
        if (use_cond) pm2(i, k) = (peg(i, k + 1) - peg(i, k)) / (pelng(i, k + 1) - pelng(i, k));
        else pm2(i, k) = dm(i, k) / (peln2(i, k + 1) - peln2(i, k));
        gm2(i, k) = 1. / (1. - cp2(i, k));
        dm(i, k) = dm(i, k) * rgrav;
        dz2(i, k) = zh(i, j, k + 1) - zh(i, j, k);
        w2(i, k) = w(i, j, k);
        pt2(i, k) = pt(i, j, k);
      }
    std::vector<double> wsr(ie - is + 1);
    for (int i = is; i <= ie; i++) wsr[i - is] = ws(i, j);
    if (a_imp > 0.999)
      sim1_solver(dt, is, ie, km, c.rdgas, gm2, cp2, pe2, dm, pm2, pem, w2, dz2, pt2, wsr.data(), p_fac, c.rff, c.k_rf);
    else
      sim_solver(dt, is, ie, km, c.rdgas, gm2, cp2, pe2, dm, pm2, pem, w2, dz2, pt2, wsr.data(), a_imp, p_fac, 0.0, c.rff, c.k_rf);
    for (int k = 1; k <= km; k++)
      for (int i = is; i <= ie; i++) { w(i, j, k) = w2(i, k); delz(i, j, k) = dz2(i, k); }
    if (last_call) {
      for (int k = 1; k <= km + 1; k++)
        for (int i = is; i <= ie; i++) {
          peln[(i - is) + (size_t)(k - 1) * nie + (size_t)(j - js) * nie * (km + 1)] = peln2(i, k);
          pk(i, j, k) = pk3(i, j, k);
          pe[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (km + 1)] = pem(i, k);
        }
    }
    if (fp_out) {
      for (int k = 1; k <= km + 1; k++) for (int i = is; i <= ie; i++) ppe(i, j, k) = pe2(i, k) + pem(i, k);
    } else {
      for (int k = 1; k <= km + 1; k++) for (int i = is; i <= ie; i++) ppe(i, j, k) = pe2(i, k);
    }
    if (use_logp)
      for (int k = 2; k <= km + 1; k++) for (int i = is; i <= ie; i++) pk3(i, j, k) = peln2(i, k);
    for (int i = is; i <= ie; i++) zh(i, j, km + 1) = zs(i, j);
    for (int k = km; k >= 1; k--) for (int i = is; i <= ie; i++) zh(i, j, k) = zh(i, j, k + 1) - dz2(i, k);
  }
}

// dyn_core.F90:1635-1694
void p_grad_c(double dt2, int npz, V3 delpc, V3 pkc, V3 gz, V3 uc, V3 vc, const Bd& bd, V2 rdxc, V2 rdyc,
              bool hydrostatic) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz; k++) {
    L2 wk(is - 1, ie + 1, js - 1, je + 1);
    if (hydrostatic) {
      for (int j = js - 1; j <= je + 1; j++) for (int i = is - 1; i <= ie + 1; i++) wk(i, j) = pkc(i, j, k + 1) - pkc(i, j, k);
    } else {
      for (int j = js - 1; j <= je + 1; j++) for (int i = is - 1; i <= ie + 1; i++) wk(i, j) = delpc(i, j, k);
    }
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++)
        uc(i, j, k) = uc(i, j, k) + dt2 * rdxc(i, j) / (wk(i - 1, j) + wk(i, j)) *
                      ((gz(i - 1, j, k + 1) - gz(i, j, k)) * (pkc(i, j, k + 1) - pkc(i - 1, j, k)) +
                       (gz(i - 1, j, k) - gz(i, j, k + 1)) * (pkc(i - 1, j, k + 1) - pkc(i, j, k)));
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++)
        vc(i, j, k) = vc(i, j, k) + dt2 * rdyc(i, j) / (wk(i, j - 1) + wk(i, j)) *
                      ((gz(i, j - 1, k + 1) - gz(i, j, k)) * (pkc(i, j, k + 1) - pkc(i, j - 1, k)) +
                       (gz(i, j - 1, k) - gz(i, j, k + 1)) * (pkc(i, j - 1, k + 1) - pkc(i, j, k)));
  }
}

// dyn_core.F90:1697-1792
void nh_p_grad(V3 u, V3 v, V3 pp, V3 gz, V3 delp, V3 pk, double dt, int ng, const Grid& g, const Bd& bd, int npx,
               int npy, int npz, bool use_logp, double ptop, double akap) {
  (void)ng; (void)npx; (void)npy;
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const double top_value = use_logp ? std::log(ptop) : std::pow(ptop, akap);  // peln1 / ptk, dyn_core.F90:220-222
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz + 1; k++) {
    L2 wk1(isd, ied, jsd, jed);
    if (k == 1) {
      for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) { pp(i, j, 1) = 0.; pk(i, j, 1) = top_value; }
    } else {
      a2b_ord4(pp.k(k), wk1, g, bd, true);
      a2b_ord4(pk.k(k), wk1, g, bd, true);
    }
    a2b_ord4(gz.k(k), wk1, g, bd, true);
  }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz; k++) {
    L2 wk1(isd, ied, jsd, jed), wk(is, ie + 1, js, je + 1);
    a2b_ord4(delp.k(k), wk1, g, bd, false);
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) wk(i, j) = pk(i, j, k + 1) - pk(i, j, k);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) {
        double du1 = dt / (wk(i, j) + wk(i + 1, j)) *
                     ((gz(i, j, k + 1) - gz(i + 1, j, k)) * (pk(i + 1, j, k + 1) - pk(i, j, k)) +
                      (gz(i, j, k) - gz(i + 1, j, k + 1)) * (pk(i, j, k + 1) - pk(i + 1, j, k)));
        u(i, j, k) = (u(i, j, k) + du1 + dt / (wk1(i, j) + wk1(i + 1, j)) *
                      ((gz(i, j, k + 1) - gz(i + 1, j, k)) * (pp(i + 1, j, k + 1) - pp(i, j, k)) +
                       (gz(i, j, k) - gz(i + 1, j, k + 1)) * (pp(i, j, k + 1) - pp(i + 1, j, k)))) * g.rdx(i, j);
      }
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) {
        double dv1 = dt / (wk(i, j) + wk(i, j + 1)) *
                     ((gz(i, j, k + 1) - gz(i, j + 1, k)) * (pk(i, j + 1, k + 1) - pk(i, j, k)) +
                      (gz(i, j, k) - gz(i, j + 1, k + 1)) * (pk(i, j, k + 1) - pk(i, j + 1, k)));
        v(i, j, k) = (v(i, j, k) + dv1 + dt / (wk1(i, j) + wk1(i, j + 1)) *
                      ((gz(i, j, k + 1) - gz(i, j + 1, k)) * (pp(i, j + 1, k + 1) - pp(i, j, k)) +
                       (gz(i, j, k) - gz(i, j + 1, k + 1)) * (pp(i, j, k + 1) - pp(i, j + 1, k)))) * g.rdy(i, j);
      }
  }
}

// dyn_core.F90:2202-2356 geopk (not SW_DYNAMICS, not bounded_domain).  pe is (is-1:ie+1, km+1, js-1:je+1), peln (is:ie, km+1, js:je)
void geopk(double ptop, double* pe, double* peln, V3 delp, V3 pk, V3 gz, V2 hs, V3 pt, V3 q_con, V3 pkz, int km, double akap,
           double cp_air, bool CG, bool use_cond, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
  const double ptk = std::pow(ptop, akap), peln1 = std::log(ptop);   // dyn_core.F90:220-222
  int ifirst, ilast, jfirst, jlast;
  if (!CG) { ifirst = is - 2; ilast = ie + 2; jfirst = js - 2; jlast = je + 2; }
  else { ifirst = is - 1; ilast = ie + 1; jfirst = js - 1; jlast = je + 1; }
  const size_t nip = ie - is + 3, nie = ie - is + 1;
  auto PE = [&](int i, int k, int j) -> double& { return pe[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (km + 1)]; };
  auto PELN = [&](int i, int k, int j) -> double& { return peln[(i - is) + (size_t)(k - 1) * nie + (size_t)(j - js) * nie * (km + 1)]; };
#pragma omp parallel for schedule(static)
  for (int j = jfirst; j <= jlast; j++) {
    L2 peg(ifirst, ilast, 1, km + 1), pkg(ifirst, ilast, 1, km + 1);
    L1 p1d(ifirst, ilast), logp(ifirst, ilast);
    for (int i = ifirst; i <= ilast; i++) {
      p1d(i) = ptop; pk(i, j, 1) = ptk; gz(i, j, km + 1) = hs(i, j);
      if (use_cond) { peg(i, 1) = ptop; pkg(i, 1) = ptk; }
    }
    if (j >= js && j <= je) for (int i = is; i <= ie; i++) PELN(i, 1, j) = peln1;
    if (j > (js - 2) && j < (je + 2)) for (int i = std::max(ifirst, is - 1); i <= std::min(ilast, ie + 1); i++) PE(i, 1, j) = ptop;
    for (int k = 2; k <= km + 1; k++) {   // top down
      for (int i = ifirst; i <= ilast; i++) {
        p1d(i) = p1d(i) + delp(i, j, k - 1);
        logp(i) = std::log(p1d(i));
        pk(i, j, k) = std::exp(akap * logp(i));
        if (use_cond) {
          peg(i, k) = peg(i, k - 1) + delp(i, j, k - 1) * (1. - q_con(i, j, k - 1));
          pkg(i, k) = std::exp(akap * std::log(peg(i, k)));
        }
      }
      if (j > (js - 2) && j < (je + 2)) {
        for (int i = std::max(ifirst, is - 1); i <= std::min(ilast, ie + 1); i++) PE(i, k, j) = p1d(i);
        if (j >= js && j <= je) for (int i = is; i <= ie; i++) PELN(i, k, j) = logp(i);
      }
    }
    for (int k = km; k >= 1; k--)   // bottom up
      for (int i = ifirst; i <= ilast; i++)
        gz(i, j, k) = use_cond ? gz(i, j, k + 1) + cp_air * pt(i, j, k) * (pkg(i, k + 1) - pkg(i, k))
                               : gz(i, j, k + 1) + cp_air * pt(i, j, k) * (pk(i, j, k + 1) - pk(i, j, k));
    if (!CG && j >= js && j <= je)
      for (int k = 1; k <= km; k++)
        for (int i = is; i <= ie; i++) pkz(i, j, k) = (pk(i, j, k + 1) - pk(i, j, k)) / (akap * (PELN(i, k + 1, j) - PELN(i, k, j)));
  }
}

// dyn_core.F90:1795-1905 split_p_grad (beta > 0, non-hydrostatic): du, dv = the hydrostatic increment of the previous substep
void split_p_grad(V3 u, V3 v, V3 pp, V3 gz, V3 delp, V3 pk, V3 du, V3 dv, double beta, double dt, const Grid& g, const Bd& bd, int npz,
                  bool use_logp, double ptop, double akap) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const double top_value = use_logp ? std::log(ptop) : std::pow(ptop, akap);
  const double alpha = 1. - beta;
  for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) { pp(i, j, 1) = 0.; pk(i, j, 1) = top_value; }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz + 1; k++) {
    L2 wk1(isd, ied, jsd, jed);
    if (k != 1) { a2b_ord4(pp.k(k), wk1, g, bd, true); a2b_ord4(pk.k(k), wk1, g, bd, true); }
    a2b_ord4(gz.k(k), wk1, g, bd, true);
  }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz; k++) {
    L2 wk1(isd, ied, jsd, jed), wk(is, ie + 1, js, je + 1);
    a2b_ord4(delp.k(k), wk1, g, bd, false);
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) wk(i, j) = pk(i, j, k + 1) - pk(i, j, k);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) {
        u(i, j, k) = u(i, j, k) + beta * du(i, j, k);
        du(i, j, k) = dt / (wk(i, j) + wk(i + 1, j)) *
                      ((gz(i, j, k + 1) - gz(i + 1, j, k)) * (pk(i + 1, j, k + 1) - pk(i, j, k)) +
                       (gz(i, j, k) - gz(i + 1, j, k + 1)) * (pk(i, j, k + 1) - pk(i + 1, j, k)));
        u(i, j, k) = (u(i, j, k) + alpha * du(i, j, k) + dt / (wk1(i, j) + wk1(i + 1, j)) *
                      ((gz(i, j, k + 1) - gz(i + 1, j, k)) * (pp(i + 1, j, k + 1) - pp(i, j, k)) +
                       (gz(i, j, k) - gz(i + 1, j, k + 1)) * (pp(i, j, k + 1) - pp(i + 1, j, k)))) * g.rdx(i, j);
      }
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) {
        v(i, j, k) = v(i, j, k) + beta * dv(i, j, k);
        dv(i, j, k) = dt / (wk(i, j) + wk(i, j + 1)) *
                      ((gz(i, j, k + 1) - gz(i, j + 1, k)) * (pk(i, j + 1, k + 1) - pk(i, j, k)) +
                       (gz(i, j, k) - gz(i, j + 1, k + 1)) * (pk(i, j, k + 1) - pk(i, j + 1, k)));
        v(i, j, k) = (v(i, j, k) + alpha * dv(i, j, k) + dt / (wk1(i, j) + wk1(i, j + 1)) *
                      ((gz(i, j, k + 1) - gz(i, j + 1, k)) * (pp(i, j + 1, k + 1) - pp(i, j, k)) +
                       (gz(i, j, k) - gz(i, j + 1, k + 1)) * (pp(i, j, k + 1) - pp(i, j + 1, k)))) * g.rdy(i, j);
      }
  }
}

// dyn_core.F90:2033-2116 grad1_p_update (beta > 0, hydrostatic), d_ext = 0 (divg2 = 0, dyn_core.F90:745-747)
void grad1_p_update(V3 u, V3 v, V3 pk, V3 gz, V3 du, V3 dv, double dt, const Grid& g, const Bd& bd, int npz, double ptop, double akap, double beta,
                    const double* divg2) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  // divg2(is:ie+1, js:je+1): the external-mode damping term of dyn_core.F90:828-847 (nullptr: d_ext = 0, divg2 = 0)
  const int nd = ie - is + 2;
  auto D2 = [&](int i, int j) { return divg2 ? divg2[(i - is) + (size_t)(j - js) * nd] : 0.; };
  const double alpha = 1. - beta, top_value = std::pow(ptop, akap);
  for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) pk(i, j, 1) = top_value;
#pragma omp parallel for schedule(static)
  for (int k = 2; k <= npz + 1; k++) { L2 wk(isd, ied, jsd, jed); a2b_ord4(pk.k(k), wk, g, bd, true); }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz + 1; k++) { L2 wk(isd, ied, jsd, jed); a2b_ord4(gz.k(k), wk, g, bd, true); }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz; k++) {
    L2 wk(isd, ied, jsd, jed);
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) wk(i, j) = pk(i, j, k + 1) - pk(i, j, k);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) {
        u(i, j, k) = u(i, j, k) + beta * du(i, j, k);
        du(i, j, k) = dt / (wk(i, j) + wk(i + 1, j)) *
                      ((gz(i, j, k + 1) - gz(i + 1, j, k)) * (pk(i + 1, j, k + 1) - pk(i, j, k)) +
                       (gz(i, j, k) - gz(i + 1, j, k + 1)) * (pk(i, j, k + 1) - pk(i + 1, j, k)));
        u(i, j, k) = (u(i, j, k) + D2(i, j) - D2(i + 1, j) + alpha * du(i, j, k)) * g.rdx(i, j);
      }
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) {
        v(i, j, k) = v(i, j, k) + beta * dv(i, j, k);
        dv(i, j, k) = dt / (wk(i, j) + wk(i, j + 1)) *
                      ((gz(i, j, k + 1) - gz(i, j + 1, k)) * (pk(i, j + 1, k + 1) - pk(i, j, k)) +
                       (gz(i, j, k) - gz(i, j + 1, k + 1)) * (pk(i, j, k + 1) - pk(i, j + 1, k)));
        v(i, j, k) = (v(i, j, k) + D2(i, j) - D2(i, j + 1) + alpha * dv(i, j, k)) * g.rdy(i, j);
      }
  }
}

// dyn_core.F90:1909-2030 one_grad_p; divg2 (nullable: d_ext = 0, wk1 = wk2 = 0) is the external-mode damping term (:1969-1984)
void one_grad_p(V3 u, V3 v, V3 pk, V3 gz, V3 delp, double dt, const Grid& g, const Bd& bd, int npz, double ptop, double akap,
                bool hydrostatic, const double* divg2) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const int nd = ie - is + 2;
  auto D2 = [&](int i, int j) { return divg2 ? divg2[(i - is) + (size_t)(j - js) * nd] : 0.; };
  const double top_value = hydrostatic ? std::pow(ptop, akap) : ptop;
  for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) pk(i, j, 1) = top_value;
#pragma omp parallel for schedule(static)
  for (int k = 2; k <= npz + 1; k++) { L2 wk(isd, ied, jsd, jed); a2b_ord4(pk.k(k), wk, g, bd, true); }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz + 1; k++) { L2 wk(isd, ied, jsd, jed); a2b_ord4(gz.k(k), wk, g, bd, true); }
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= npz; k++) {
    L2 wk(isd, ied, jsd, jed);
    if (hydrostatic) { for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) wk(i, j) = pk(i, j, k + 1) - pk(i, j, k); }
    else a2b_ord4(delp.k(k), wk, g, bd, false);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++)
        u(i, j, k) = g.rdx(i, j) * ((D2(i, j) - D2(i + 1, j)) + u(i, j, k) + dt / (wk(i, j) + wk(i + 1, j)) *
                                    ((gz(i, j, k + 1) - gz(i + 1, j, k)) * (pk(i + 1, j, k + 1) - pk(i, j, k)) +
                                     (gz(i, j, k) - gz(i + 1, j, k + 1)) * (pk(i, j, k + 1) - pk(i + 1, j, k))));
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++)
        v(i, j, k) = g.rdy(i, j) * ((D2(i, j) - D2(i, j + 1)) + v(i, j, k) + dt / (wk(i, j) + wk(i, j + 1)) *
                                    ((gz(i, j, k + 1) - gz(i, j + 1, k)) * (pk(i, j + 1, k + 1) - pk(i, j, k)) +
                                     (gz(i, j, k) - gz(i, j + 1, k + 1)) * (pk(i, j, k + 1) - pk(i, j + 1, k))));
  }
}

// dyn_core.F90:1395-1447
void pk3_halo(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, int npz, double ptop,
              double akap, V3 pk3, V3 delp) {
  (void)isd; (void)ied; (void)jsd; (void)jed;
  for (int j = js; j <= je; j++) {
    const int ii[4] = {is - 2, is - 1, ie + 1, ie + 2};
    for (int n = 0; n < 4; n++) {
      double pei = ptop;
      for (int k = 1; k <= npz; k++) { pei = pei + delp(ii[n], j, k); pk3(ii[n], j, k + 1) = std::exp(akap * std::log(pei)); }
    }
  }
  for (int i = is - 2; i <= ie + 2; i++) {
    const int jj[4] = {js - 2, js - 1, je + 1, je + 2};
    for (int n = 0; n < 4; n++) {
      double pej = ptop;
      for (int k = 1; k <= npz; k++) { pej = pej + delp(i, jj[n], k); pk3(i, jj[n], k + 1) = std::exp(akap * std::log(pej)); }
    }
  }
}

// dyn_core.F90:1449-1496 (use_logp: pk3 holds log(pe) instead of pe**akap)
void pln_halo(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, int npz, double ptop, V3 pk3, V3 delp) {
  (void)isd; (void)ied; (void)jsd; (void)jed;
  for (int j = js; j <= je; j++) {
    const int ii[4] = {is - 2, is - 1, ie + 1, ie + 2};
    for (int n = 0; n < 4; n++) {
      double pet = ptop;
      for (int k = 1; k <= npz; k++) { pet = pet + delp(ii[n], j, k); pk3(ii[n], j, k + 1) = std::log(pet); }
    }
  }
  for (int i = is - 2; i <= ie + 2; i++) {
    const int jj[4] = {js - 2, js - 1, je + 1, je + 2};
    for (int n = 0; n < 4; n++) {
      double pet = ptop;
      for (int k = 1; k <= npz; k++) { pet = pet + delp(i, jj[n], k); pk3(i, jj[n], k + 1) = std::log(pet); }
    }
  }
}

// dyn_core.F90:1498-1526
void pe_halo(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, int npz, double ptop,
             double* pe, V3 delp) {
  (void)isd; (void)ied; (void)jsd; (void)jed;
  const int nip = ie + 1 - (is - 1) + 1;
  auto PE = [&](int i, int k, int j) -> double& {
    return pe[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (npz + 1)];
  };
  for (int j = js; j <= je; j++) {
    PE(is - 1, 1, j) = ptop; PE(ie + 1, 1, j) = ptop;
    for (int k = 1; k <= npz; k++) {
      PE(is - 1, k + 1, j) = PE(is - 1, k, j) + delp(is - 1, j, k);
      PE(ie + 1, k + 1, j) = PE(ie + 1, k, j) + delp(ie + 1, j, k);
    }
  }
  for (int i = is - 1; i <= ie + 1; i++) {
    PE(i, 1, js - 1) = ptop; PE(i, 1, je + 1) = ptop;
    for (int k = 1; k <= npz; k++) {
      PE(i, k + 1, js - 1) = PE(i, k, js - 1) + delp(i, js - 1, k);
      PE(i, k + 1, je + 1) = PE(i, k, je + 1) + delp(i, je + 1, k);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// dyn_core.F90:2356-2465 del2_cubed: nmax (<= 3) passes of a del-2 filter on an A-grid scalar whose halo has been updated by
// the caller (the mpp_update_domains at :2401 is the harness's / dyn_core's exchange).  Non-USE_SG build (del6_u / del6_v).
void del2_cubed(V3 q, double cd, const Grid& g, const Bd& bd, int km, int nmax) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const int npx = bd.npx, npy = bd.npy;
  const double r3 = 1. / 3.;
  const int ntimes = std::min(3, nmax);
  for (int n = 1; n <= ntimes; n++) {
    const int nt = ntimes - n;
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= km; k++) {
      L2 fx(isd, ied + 1, jsd, jed), fy(isd, ied, jsd, jed + 1);
      V2 qk = q.k(k);
      if (bd.sw_corner) { qk(1, 1) = (qk(1, 1) + qk(0, 1) + qk(1, 0)) * r3; qk(0, 1) = qk(1, 1); qk(1, 0) = qk(1, 1); }
      if (bd.se_corner) { qk(ie, 1) = (qk(ie, 1) + qk(npx, 1) + qk(ie, 0)) * r3; qk(npx, 1) = qk(ie, 1); qk(ie, 0) = qk(ie, 1); }
      if (bd.ne_corner) { qk(ie, je) = (qk(ie, je) + qk(npx, je) + qk(ie, npy)) * r3; qk(npx, je) = qk(ie, je); qk(ie, npy) = qk(ie, je); }
      if (bd.nw_corner) { qk(1, je) = (qk(1, je) + qk(0, je) + qk(1, npy)) * r3; qk(0, je) = qk(1, je); qk(1, npy) = qk(1, je); }
      if (nt > 0 && !bd.bounded_domain) copy_corners(qk, npx, npy, 1, bd);
      for (int j = js - nt; j <= je + nt; j++)
        for (int i = is - nt; i <= ie + 1 + nt; i++) fx(i, j) = g.del6_v(i, j) * (qk(i - 1, j) - qk(i, j));
      if (nt > 0 && !bd.bounded_domain) copy_corners(qk, npx, npy, 2, bd);
      for (int j = js - nt; j <= je + 1 + nt; j++)
        for (int i = is - nt; i <= ie + nt; i++) fy(i, j) = g.del6_u(i, j) * (qk(i, j - 1) - qk(i, j));
      for (int j = js - nt; j <= je + nt; j++)
        for (int i = is - nt; i <= ie + nt; i++)
          qk(i, j) = qk(i, j) + cd * g.rarea(i, j) * (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1));
    }
  }
}

// dyn_core.F90:296-307: number of levels that receive the dissipative heating
int n_con_levels(const fv3_flags_t& f, int npz) {
  if (f.convert_ke || (f.do_vort_damp && f.vtdm4 > 1.E-4)) return npz;
  if (f.d2_bg_k1 < 1.E-3) return 0;
  return (f.d2_bg_k2 < 1.E-3) ? 1 : 2;
}

// dyn_core.F90:1305-1356: the filtered heat_source becomes a temperature tendency, limited by delt_max, added to pt
// (pt = cp*(virtual temperature / pkz) scaling of the acoustic loop).  moist_kappa = F branch for the non-hydrostatic case.
void dcon_heating(V3 pt, V3 heat_source, V3 delp, V3 delz, V3 pkz, int n_con, double bdt, const fv3_flags_t& f, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
  const double rdg = -f.rdgas / f.grav, cv_air = f.cp_air - f.rdgas, k1k = f.kappa / (1. - f.kappa);
  if (f.hydrostatic) {
    for (int j = js; j <= je; j++)
      for (int k = 1; k <= n_con; k++) {
        if (k < 3) {
          for (int i = is; i <= ie; i++) pt(i, j, k) = pt(i, j, k) + heat_source(i, j, k) / (f.cp_air * delp(i, j, k) * pkz(i, j, k));
        } else {
          for (int i = is; i <= ie; i++) {
            const double dtmp = heat_source(i, j, k) / (f.cp_air * delp(i, j, k));
            pt(i, j, k) = pt(i, j, k) + fsign(std::min(std::fabs(bdt) * f.delt_max, std::fabs(dtmp)), dtmp) / pkz(i, j, k);
            heat_source(i, j, k) = dtmp;
          }
        }
      }
  } else {
    for (int k = 1; k <= n_con; k++) {
      double delt = std::fabs(bdt * f.delt_max);
      if (k == 1) delt = 0.1 * delt;
      if (k == 2) delt = 0.5 * delt;
      for (int j = js; j <= je; j++)
        for (int i = is; i <= ie; i++) {
          pkz(i, j, k) = std::exp(k1k * std::log(rdg * delp(i, j, k) / delz(i, j, k) * pt(i, j, k)));
          const double dtmp = heat_source(i, j, k) / (cv_air * delp(i, j, k));
          pt(i, j, k) = pt(i, j, k) + fsign(std::min(delt, std::fabs(dtmp)), dtmp) / pkz(i, j, k);
          heat_source(i, j, k) = dtmp;
        }
    }
  }
}

// fv_dynamics.F90:303-328, :377-398 (not SW_DYNAMICS; moist_kappa = F): dp1 = zvir q_v, pkz from the gas law (non-hydrostatic;
// the hydrostatic pkz comes from p_var / the previous remap), pt -> virtual potential (density) temperature pt (1+dp1)[(1-q_con)]/pkz
void pt_to_theta(V3 pt, V3 delp, V3 delz, V3 qv, V3 q_con, V3 dp1, V3 pkz, double zvir, const fv3_flags_t& f, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je;
  const double rdg = -f.rdgas / f.grav;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= bd.npz; k++)
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++) {
        dp1(i, j, k) = zvir * qv(i, j, k);
        if (!f.hydrostatic) pkz(i, j, k) = std::exp(f.kappa * std::log(rdg * delp(i, j, k) * pt(i, j, k) * (1. + dp1(i, j, k)) / delz(i, j, k)));
        if (f.use_cond) pt(i, j, k) = pt(i, j, k) * (1. + dp1(i, j, k)) * (1. - q_con(i, j, k)) / pkz(i, j, k);
        else pt(i, j, k) = pt(i, j, k) * (1. + dp1(i, j, k)) / pkz(i, j, k);
      }
}

// dyn_core.F90:409-422: pem(is-1:ie+1, npz+1, js-1:je+1), the interface pressures before the last substep
void pem_from_delp(double* pem, V3 delp, double ptop, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, km = bd.npz;
  const size_t nip = ie - is + 3;
  auto PEM = [&](int i, int k, int j) -> double& { return pem[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (km + 1)]; };
#pragma omp parallel for schedule(static)
  for (int j = js - 1; j <= je + 1; j++) {
    for (int i = is - 1; i <= ie + 1; i++) PEM(i, 1, j) = ptop;
    for (int k = 1; k <= km; k++)
      for (int i = is - 1; i <= ie + 1; i++) PEM(i, k + 1, j) = PEM(i, k, j) + delp(i, j, k);
  }
}

// dyn_core.F90:1182-1195 (use_old_omega = T): omga = (pe - pem) * rdt, then adv_pe (:1529-1630): the advective term
// 0.5 * rarea * V3 . grad(pe), grad by Green's theorem around the cell from the corner values of pem (a2b_ord2)
void omega_old(V3 omga, const double* pe, const double* pem, V3 ua, V3 va, double rdt, const Grid& g, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, km = bd.npz, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const size_t nip = ie - is + 3;
  auto P = [&](const double* a, int i, int k, int j) { return a[(i - (is - 1)) + (size_t)(k - 1) * nip + (size_t)(j - (js - 1)) * nip * (km + 1)]; };
  const int nia = ied - isd + 1, nic = ie - is + 1;
  auto EC = [&](const double* e, int n, int i, int j) { return e[(n - 1) + 3 * ((i - isd) + (size_t)(j - jsd) * nia)]; };
  auto EN1 = [&](int n, int i, int j) { return g.en1_p[(n - 1) + 3 * ((i - is) + (size_t)(j - js) * nic)]; };          // (3, is:ie, js:je+1)
  auto EN2 = [&](int n, int i, int j) { return g.en2_p[(n - 1) + 3 * ((i - is) + (size_t)(j - js) * (nic + 1))]; };    // (3, is:ie+1, js:je)
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= km; k++)
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++) omga(i, j, k) = (P(pe, i, k + 1, j) - P(pem, i, k + 1, j)) * rdt;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= km; k++) {
    L2 up(is, ie, js, je), vp(is, ie, js, je), pin(isd, ied, jsd, jed), pb(isd, ied, jsd, jed);
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++) {
        if (k == km) { up(i, j) = ua(i, j, km); vp(i, j) = va(i, j, km); }
        else { up(i, j) = 0.5 * (ua(i, j, k) + ua(i, j, k + 1)); vp(i, j) = 0.5 * (va(i, j, k) + va(i, j, k + 1)); }
      }
    for (int j = js - 1; j <= je + 1; j++) for (int i = is - 1; i <= ie + 1; i++) pin(i, j) = P(pem, i, k + 1, j);
    a2b_ord2(pin, pb, g, bd);
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie; i++) {
        double acc = 0.;
        for (int n = 1; n <= 3; n++) {
          const double v3 = up(i, j) * EC(g.ec1_p, n, i, j) + vp(i, j) * EC(g.ec2_p, n, i, j);
          const double pdx0 = (pb(i, j) + pb(i + 1, j)) * g.dx(i, j) * EN1(n, i, j);
          const double pdx1 = (pb(i, j + 1) + pb(i + 1, j + 1)) * g.dx(i, j + 1) * EN1(n, i, j + 1);
          const double pdy0 = (pb(i, j) + pb(i, j + 1)) * g.dy(i, j) * EN2(n, i, j);
          const double pdy1 = (pb(i + 1, j) + pb(i + 1, j + 1)) * g.dy(i + 1, j) * EN2(n, i + 1, j);
          const double grad = pdx1 - pdx0 - pdy0 + pdy1;
          if (n == 1) acc = v3 * grad; else acc = acc + v3 * grad;
        }
        omga(i, j, k) = omga(i, j, k) + 0.5 * g.rarea(i, j) * acc;
      }
  }
}

}  // namespace fv3o
