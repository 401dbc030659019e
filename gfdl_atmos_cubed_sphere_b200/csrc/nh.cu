// Non-hydrostatic column solvers and height advection (sm_100a).
//
// Reference semantics: model/nh_utils.F90 update_dz_c (:59-201), update_dz_d (:204-321),
// Riem_Solver_c (:323-480), SIM1_solver (:1277-1394), SIM_solver (:1396-1537),
// edge_profile (:1590-1696); model/nh_core.F90 Riem_Solver3 (:47-241);
// model/dyn_core.F90 pk3_halo (:1395-1447), pe_halo (:1498-1526), gz bookkeeping (:370-385,
// :491-521, :982-989).
//
// Design: the reference solves (i,k) slabs per j under OpenMP; here ONE THREAD OWNS ONE
// COLUMN (consecutive threads on consecutive i, so every level access is a coalesced row
// segment) and marches the Thomas recurrences in k.  The few per-level temporaries that
// must survive between the forward and backward sweeps live in [k][j][i] scratch planes
// (gam, pp, w2, pm2); everything else (pem, dm, g_rat, bb, dd, aa) is recomputed from the
// inputs in registers in exactly the reference's operation order.
#include "tp2d.cuh"
#include "ppm.cuh"
#include <cmath>

#define CB 128
#define COL_SETUP(I0, I1, J0, J1)                                   \
  const int ni_ = (I1) - (I0) + 1, nj_ = (J1) - (J0) + 1;            \
  const int t_ = blockIdx.x * blockDim.x + threadIdx.x;              \
  if (t_ >= ni_ * nj_) return;                                       \
  const int i = (I0) + t_ % ni_, j = (J0) + t_ / ni_;                \
  const long long o = LIDX(L, i, j);                                 \
  const long long P = L.plane;
static inline int col_blocks(int ni, int nj) { return (ni * nj + CB - 1) / CB; }

namespace {
constexpr double r3 = 1. / 3.;
constexpr double dz_min = 2.;   // nh_utils.F90:46-50

struct SolverIn {
  const double *delp, *pt, *hgt /*interface heights (m): dz2 = hgt(k+1)-hgt(k)*/, *w, *ws, *q_con, *cappa;
  double *pe /*(km+1) out: perturbation pressure*/, *pm2, *gam, *pp, *w2;
  double dt, rgrav, rdgas, akap, ptop, p_fac, a_imp;
  int km, use_cond, moist_kappa, d_grid;
};

// One column of SIM1_solver (a_imp > 0.999) or SIM_solver.  On exit: S.w2 holds the new w,
// S.pe the perturbation pressure at interfaces, and dz2(k) is returned through out_dz(k) in
// the order k = km..1 (the caller rebuilds heights / geopotential while it is produced).
template <class DzSink>
__device__ __forceinline__ void solve_column(const SolverIn& S, long long o, long long P, DzSink out_dz) {
  const int km = S.km;
  const bool sim1 = S.d_grid ? (S.a_imp > 0.999) : true;   // nh_utils.F90:450-459, nh_core.F90:169-185
  const double alpha = sim1 ? 1.0 : S.a_imp;
  const double beta = 1. - alpha, ra = 1. / alpha, t2 = beta / alpha;
  const double t1g = sim1 ? 2. * S.dt * S.dt : 2. * (alpha * S.dt) * (alpha * S.dt);
  const double rdt = 1. / S.dt, dt = S.dt;
  auto DM = [&](int k) { return __ldg(S.delp + o + (long long)(k - 1) * P) * S.rgrav; };
  auto DZ = [&](int k) { return __ldg(S.hgt + o + (long long)k * P) - __ldg(S.hgt + o + (long long)(k - 1) * P); };
  auto CP2 = [&](int k) { return S.moist_kappa ? __ldg(S.cappa + o + (long long)(k - 1) * P) : S.akap; };
  auto GM2 = [&](int k) { return 1. / (1. - CP2(k)); };
  auto W1 = [&](int k) { return __ldg(S.w + o + (long long)(k - 1) * P); };
  auto PT = [&](int k) { return __ldg(S.pt + o + (long long)(k - 1) * P); };
  // ---- pass 1: pm2, pe(k) = p_gas' (nh_utils.F90:377-447, :1297-1302)
  {
    double pem = S.ptop, peg = S.ptop, lpem = 0., lpeg = 0.;
    if (S.d_grid) { lpem = log(S.ptop); lpeg = lpem; }
    for (int k = 1; k <= km; k++) {
      const double dmr = __ldg(S.delp + o + (long long)(k - 1) * P);
      const double pem1 = pem + dmr;
      double pm2;
      if (S.d_grid) {   // nh_core.F90:120-165 (logs of interface pressures)
        const double lpem1 = log(pem1);
        if (S.use_cond) {
          const double peg1 = peg + dmr * (1. - __ldg(S.q_con + o + (long long)(k - 1) * P));
          const double lpeg1 = log(peg1);
          pm2 = (peg1 - peg) / (lpeg1 - lpeg);
          peg = peg1; lpeg = lpeg1;
        } else pm2 = dmr / (lpem1 - lpem);
        lpem = lpem1;
      } else {          // nh_utils.F90:412-447
        if (S.use_cond) {
          const double peg1 = peg + dmr * (1. - __ldg(S.q_con + o + (long long)(k - 1) * P));
          pm2 = (peg1 - peg) / log(peg1 / peg);
          peg = peg1;
        } else pm2 = dmr / log(pem1 / pem);
      }
      pem = pem1;
      const double dm = dmr * S.rgrav;
      S.pm2[o + (long long)(k - 1) * P] = pm2;
      S.pe[o + (long long)(k - 1) * P] = exp(GM2(k) * log(-dm / DZ(k) * S.rdgas * PT(k))) - pm2;
    }
  }
  auto PE = [&](int k) -> double& { return S.pe[o + (long long)(k - 1) * P]; };
  auto PP = [&](int k) -> double& { return S.pp[o + (long long)(k - 1) * P]; };
  auto GAM = [&](int k) -> double& { return S.gam[o + (long long)(k - 1) * P]; };
  auto W2 = [&](int k) -> double& { return S.w2[o + (long long)(k - 1) * P]; };
  // ---- cubic-spline edge pressures pp (nh_utils.F90:1304-1332)
  {
    double g_rat = DM(1) / DM(2);
    double bet = 2. * (1. + g_rat);
    PP(1) = 0.;
    double ppk = 3. * (PE(1) + g_rat * PE(2)) / bet;   // pp(2)
    PP(2) = ppk;
    for (int k = 2; k <= km; k++) {
      const double gam = g_rat / bet;   // g_rat(k-1)/bet
      double bb, dd;
      if (k < km) { g_rat = DM(k) / DM(k + 1); bb = 2. * (1. + g_rat); dd = 3. * (PE(k) + g_rat * PE(k + 1)); }
      else { bb = 2.; dd = 3. * PE(km); }
      GAM(k) = gam;
      bet = bb - gam;
      ppk = (dd - ppk) / bet;
      PP(k + 1) = ppk;
    }
    double nxt = PP(km + 1);
    for (int k = km; k >= 2; k--) { nxt = PP(k) - GAM(k) * nxt; PP(k) = nxt; }
  }
  // ---- w solver (nh_utils.F90:1335-1361 / :1463-1496)
  {
    double pem = S.ptop + __ldg(S.delp + o);   // pem(2)
    auto AA = [&](int k, double pemk) { return t1g * 0.5 * (GM2(k - 1) + GM2(k)) / (DZ(k - 1) + DZ(k)) * pemk; };
    double aa_k1 = AA(2, pem);                // aa(2)
    double wk_k1 = sim1 ? 0. : t2 * aa_k1 * (W1(1) - W1(2));   // wk(2)
    double bet = DM(1) - aa_k1;
    double w2p = sim1 ? (DM(1) * W1(1) + dt * PP(2)) / bet : (DM(1) * W1(1) + dt * PP(2) + wk_k1) / bet;
    W2(1) = w2p;
    double aa_k = aa_k1, wk_k = wk_k1;
    for (int k = 2; k <= km - 1; k++) {
      pem = pem + __ldg(S.delp + o + (long long)(k - 1) * P);   // pem(k+1)
      const double aa_n = AA(k + 1, pem);
      const double wk_n = sim1 ? 0. : t2 * aa_n * (W1(k) - W1(k + 1));
      const double gam = aa_k / bet;
      GAM(k) = gam;
      bet = DM(k) - (aa_k + aa_n + aa_k * gam);
      if (sim1) w2p = (DM(k) * W1(k) + dt * (PP(k + 1) - PP(k)) - aa_k * w2p) / bet;
      else w2p = (DM(k) * W1(k) + dt * (PP(k + 1) - PP(k)) + wk_n - wk_k - aa_k * w2p) / bet;
      W2(k) = w2p;
      aa_k = aa_n; wk_k = wk_n;
    }
    pem = pem + __ldg(S.delp + o + (long long)(km - 1) * P);     // pem(km+1)
    const double p1 = t1g * GM2(km) / DZ(km) * pem;
    const double gam = aa_k / bet;
    GAM(km) = gam;
    bet = DM(km) - (aa_k + p1 + aa_k * gam);
    const double wsv = __ldg(S.ws + o);
    if (sim1) w2p = (DM(km) * W1(km) + dt * (PP(km + 1) - PP(km)) - p1 * wsv - aa_k * w2p) / bet;
    else w2p = (DM(km) * W1(km) + dt * (PP(km + 1) - PP(km)) - wk_k + p1 * (t2 * W1(km) - ra * wsv) - aa_k * w2p) / bet;
    W2(km) = w2p;
    for (int k = km - 1; k >= 1; k--) { w2p = W2(k) - GAM(k + 1) * w2p; W2(k) = w2p; }
  }
  // ---- perturbation pressure (nh_utils.F90:1373-1380 / :1508-1516)
  {
    double pe = 0.;
    double pe_prev_store = 0.;
    (void)pe_prev_store;
    // pe(1) = 0 ; pe(k+1) = pe(k) + ...   (PE(k) for k<=km is consumed: overwrite in order)
    double carry = 0.;   // pe(k)
    for (int k = 1; k <= km; k++) {
      double nxt;
      if (sim1) nxt = carry + DM(k) * (W2(k) - W1(k)) * rdt;
      else nxt = carry + (DM(k) * (W2(k) - W1(k)) * rdt - beta * (PP(k + 1) - PP(k))) * ra;
      PE(k) = carry;
      carry = nxt;
    }
    PE(km + 1) = carry;
    (void)pe;
  }
  // ---- dz2 from the spline of pe (nh_utils.F90:1382-1392), k = km..1
  {
    double p1 = (PE(km) + 2. * PE(km + 1)) * r3;
    auto DZ2 = [&](int k, double p1v) {
      const double pm2 = S.pm2[o + (long long)(k - 1) * P];
      return -DM(k) * S.rdgas * PT(k) * exp((CP2(k) - 1.) * log(fmax(S.p_fac * pm2, p1v + pm2)));
    };
    out_dz(km, DZ2(km, p1));
    for (int k = km - 1; k >= 1; k--) {
      const double g_rat = DM(k) / DM(k + 1), bb = 2. * (1. + g_rat);
      p1 = (PE(k) + bb * PE(k + 1) + g_rat * PE(k + 2)) * r3 - g_rat * p1;
      out_dz(k, DZ2(k, p1));
    }
  }
  if (!sim1) {   // nh_utils.F90:1531-1535
    for (int k = 1; k <= km + 1; k++) PE(k) = PE(k) + beta * (PP(k) - PE(k));
  }
}
}  // namespace

// ---- Riem_Solver_c (nh_utils.F90:323-480) on columns [is-1, ie+1]^2 -------------------------
// 8 CTAs x 128 threads per SM: all columns of a C384 face are resident in ONE wave (72-88 registers gave 1.56 waves)
__global__ void __launch_bounds__(CB, 8) k_riem_c(Lay L, SolverIn S, const double* __restrict__ hs, double* __restrict__ gz,
                                               double* __restrict__ pef, double grav) {
  COL_SETUP(L.is - 1, L.ie + 1, L.js - 1, L.je + 1)
  const int km = S.km;
  double gzk = __ldg(hs + o);   // gz(km+1) = hs
  // heights are read from gz (input, m) inside solve_column; gz is overwritten bottom-up only
  // after the w-solver has consumed dz2, in the same backward order as nh_utils.F90:468-476
  solve_column(S, o, P, [&](int k, double dz2) {
    if (k == km) gz[o + (long long)km * P] = gzk;
    gzk = gzk - dz2 * grav;
    gz[o + (long long)(k - 1) * P] = gzk;
  });
  // pef = pe2 + pem (nh_utils.F90:461-465), top = ptop
  double pem = S.ptop;
  pef[o] = S.ptop;
  for (int k = 2; k <= km + 1; k++) {
    pem = pem + __ldg(S.delp + o + (long long)(k - 2) * P);
    pef[o + (long long)(k - 1) * P] = S.pe[o + (long long)(k - 1) * P] + pem;
  }
}

// ---- Riem_Solver3 (nh_core.F90:47-241) on columns [is, ie]x[js, je] -------------------------
__global__ void __launch_bounds__(CB, 8) k_riem3(Lay L, SolverIn S, const double* __restrict__ zs, double* __restrict__ zh,
                                              double* __restrict__ w, double* __restrict__ delz, double* __restrict__ ppe,
                                              double* __restrict__ pk3, double* __restrict__ pk, double* __restrict__ pe,
                                              double* __restrict__ peln, int last_call, int fp_out, int use_logp) {
  COL_SETUP(L.is, L.ie, L.js, L.je)
  const int km = S.km;
  double zk = __ldg(zs + o);
  solve_column(S, o, P, [&](int k, double dz2) {
    if (k == km) zh[o + (long long)km * P] = zk;
    delz[o + (long long)(k - 1) * P] = dz2;
    zk = zk - dz2;
    zh[o + (long long)(k - 1) * P] = zk;
  });
  // w, pk3, ppe (+ pe, pk, peln on the last call)
  const double peln1 = log(S.ptop);
  const double ptk = exp(S.akap * peln1);
  double pem = S.ptop;
  for (int k = 1; k <= km + 1; k++) {
    const long long ok = o + (long long)(k - 1) * P;
    double pl, pkv;
    if (k == 1) { pl = peln1; pkv = ptk; }
    else {
      pem = pem + __ldg(S.delp + o + (long long)(k - 2) * P);
      pl = log(pem);
      pkv = exp(S.akap * pl);
    }
    if (last_call) { peln[ok] = pl; pk[ok] = pkv; pe[ok] = pem; }
    const double pe2 = S.pe[ok];
    ppe[ok] = fp_out ? pe2 + pem : pe2;
    pk3[ok] = (use_logp && k >= 2) ? pl : pkv;
    if (k <= km) w[ok] = S.w2[ok];
  }
}

// ---- update_dz_c (nh_utils.F90:59-201) ------------------------------------------------------
#define TI 32
#define TJ 8
__global__ void __launch_bounds__(TI* TJ) k_dzc_adv(Lay L, DevGrid G, const double* __restrict__ ut, const double* __restrict__ vt,
                                                   const double* __restrict__ gz, double* __restrict__ gzn, const double* __restrict__ dp0) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  const int k = blockIdx.z + 1;   // 1..km+1
  if (i < L.is - 1 || i > L.ie + 1 || j < L.js - 1 || j > L.je + 1) return;
  const int km = L.npz;
  const long long P = L.plane;
  const double* gzk = gz + (long long)(k - 1) * P;
  auto U = [&](int ii, int jj, int kk) { return __ldg(ut + LIDX(L, ii, jj) + (long long)(kk - 1) * P); };
  auto V = [&](int ii, int jj, int kk) { return __ldg(vt + LIDX(L, ii, jj) + (long long)(kk - 1) * P); };
  double r0 = 0., r1 = 0.;
  int mode;
  if (k == 1) { mode = 0; r0 = dp0[0] / (dp0[0] + dp0[1]); }
  else if (k == km + 1) { mode = 1; r0 = dp0[km - 1] / (dp0[km - 2] + dp0[km - 1]); }
  else { mode = 2; r0 = 1. / (dp0[k - 2] + dp0[k - 1]); r1 = 0.; }
  (void)r1;
  auto XF = [&](int ii, int jj) {
    if (mode == 0) return U(ii, jj, 1) + (U(ii, jj, 1) - U(ii, jj, 2)) * r0;
    if (mode == 1) return U(ii, jj, km) + (U(ii, jj, km) - U(ii, jj, km - 1)) * r0;
    return (dp0[k - 1] * U(ii, jj, k - 1) + dp0[k - 2] * U(ii, jj, k)) * r0;
  };
  auto YF = [&](int ii, int jj) {
    if (mode == 0) return V(ii, jj, 1) + (V(ii, jj, 1) - V(ii, jj, 2)) * r0;
    if (mode == 1) return V(ii, jj, km) + (V(ii, jj, km) - V(ii, jj, km - 1)) * r0;
    return (dp0[k - 1] * V(ii, jj, k - 1) + dp0[k - 2] * V(ii, jj, k)) * r0;
  };
  // fill_4corners views of gz2 (sw_core.F90:3496-3555)
  auto GX = [&](int ii, int jj) {
    if (L.cube) {
      if (jj == 0) {
        if (ii == -1) { ii = 0; jj = 2; } else if (ii == 0) { jj = 1; }
        else if (ii == L.npx + 1) { ii = L.npx; jj = 2; } else if (ii == L.npx) { jj = 1; }
      } else if (jj == L.npy) {
        if (ii == 0) { jj = L.npy - 1; } else if (ii == -1) { ii = 0; jj = L.npy - 2; }
        else if (ii == L.npx) { jj = L.npy - 1; } else if (ii == L.npx + 1) { ii = L.npx; jj = L.npy - 2; }
      }
    }
    return __ldg(gzk + LIDX(L, ii, jj));
  };
  auto GY = [&](int ii, int jj) {
    if (L.cube) {
      if (ii == 0) {
        if (jj == 0) { ii = 1; } else if (jj == -1) { ii = 2; jj = 0; }
        else if (jj == L.npy) { ii = 1; } else if (jj == L.npy + 1) { ii = 2; jj = L.npy; }
      } else if (ii == L.npx) {
        if (jj == 0) { ii = L.npx - 1; } else if (jj == -1) { ii = L.npx - 2; jj = 0; }
        else if (jj == L.npy) { ii = L.npx - 1; } else if (jj == L.npy + 1) { ii = L.npx - 2; jj = L.npy; }
      }
    }
    return __ldg(gzk + LIDX(L, ii, jj));
  };
  const double x0 = XF(i, j), x1 = XF(i + 1, j), y0 = YF(i, j), y1 = YF(i, j + 1);
  const double fx0 = x0 * ((x0 > 0.) ? GX(i - 1, j) : GX(i, j));
  const double fx1 = x1 * ((x1 > 0.) ? GX(i, j) : GX(i + 1, j));
  const double fy0 = y0 * ((y0 > 0.) ? GY(i, j - 1) : GY(i, j));
  const double fy1 = y1 * ((y1 > 0.) ? GY(i, j) : GY(i, j + 1));
  const double ar = __ldg(G.area + LIDX(L, i, j));
  // gz2(i,j) at the centre is the array after BOTH fills (dir=2 last), nh_utils.F90:163,177
  gzn[LIDX(L, i, j) + (long long)(k - 1) * P] = (GY(i, j) * ar + fx0 - fx1 + fy0 - fy1) / (ar + x0 - x1 + y0 - y1);
}
// ws and the monotonic-height clamp (nh_utils.F90:183-199 / :303-319); writes h in place
__global__ void __launch_bounds__(CB) k_dz_clamp(Lay L, const double* __restrict__ hn, double* __restrict__ h, const double* __restrict__ phis,
                                                double* __restrict__ ws, double rgrav, double rdt, int halo) {
  COL_SETUP(L.is - halo, L.ie + halo, L.js - halo, L.je + halo)
  const int km = L.npz;
  double below = __ldg(hn + o + (long long)km * P);
  ws[o] = (__ldg(phis + o) * rgrav - below) * rdt;
  h[o + (long long)km * P] = below;
  for (int k = km; k >= 1; k--) {
    const double v = fmax(__ldg(hn + o + (long long)(k - 1) * P), below + dz_min);
    h[o + (long long)(k - 1) * P] = v;
    below = v;
  }
}

// ---- update_dz_d (nh_utils.F90:204-321) -----------------------------------------------------
// edge_profile (nh_utils.F90:1638-1672, non-uniform branch, limiter = 0) for a pair of fields
__global__ void __launch_bounds__(CB, 8) k_edge_profile(Lay L, const double* __restrict__ q1, const double* __restrict__ q2,
                                                    double* __restrict__ q1e, double* __restrict__ q2e, double* __restrict__ gam,
                                                    const double* __restrict__ dp0, int i0, int i1, int j0, int j1) {
  COL_SETUP(i0, i1, j0, j1)
  const int km = L.npz;
  auto Q1 = [&](int k) { return __ldg(q1 + o + (long long)(k - 1) * P); };
  auto Q2 = [&](int k) { return __ldg(q2 + o + (long long)(k - 1) * P); };
  const double g0 = dp0[1] / dp0[0];
  double xt1 = 2. * g0 * (g0 + 1.);
  double bet = g0 * (g0 + 0.5);
  double e1 = (xt1 * Q1(1) + Q1(2)) / bet, e2 = (xt1 * Q2(1) + Q2(2)) / bet;
  double gm = (1. + g0 * (g0 + 1.5)) / bet;
  q1e[o] = e1; q2e[o] = e2; gam[o] = gm;
  double gk = 0.;
  for (int k = 2; k <= km; k++) {
    gk = dp0[k - 2] / dp0[k - 1];
    bet = 2. + 2. * gk - gm;
    e1 = (3. * (Q1(k - 1) + gk * Q1(k)) - e1) / bet;
    e2 = (3. * (Q2(k - 1) + gk * Q2(k)) - e2) / bet;
    gm = gk / bet;
    const long long ok = o + (long long)(k - 1) * P;
    q1e[ok] = e1; q2e[ok] = e2; gam[ok] = gm;
  }
  const double a_bot = 1. + gk * (gk + 1.5);
  xt1 = 2. * gk * (gk + 1.);
  const double xt2 = gk * (gk + 0.5) - a_bot * gm;
  e1 = (xt1 * Q1(km) + Q1(km - 1) - a_bot * e1) / xt2;
  e2 = (xt1 * Q2(km) + Q2(km - 1) - a_bot * e2) / xt2;
  q1e[o + (long long)km * P] = e1; q2e[o + (long long)km * P] = e2;
  for (int k = km; k >= 1; k--) {
    const long long ok = o + (long long)(k - 1) * P;
    const double g = gam[ok];
    e1 = q1e[ok] - g * e1; e2 = q2e[ok] - g * e2;
    q1e[ok] = e1; q2e[ok] = e2;
  }
}
// zh update from the transport fluxes (nh_utils.F90:282-299); del6 term only where damp(k) > 1e-5
__global__ void __launch_bounds__(TI* TJ) k_dzd_upd(Lay L, DevGrid G, const double* __restrict__ zh, const double* __restrict__ fx,
                                                   const double* __restrict__ fy, const double* __restrict__ xfa, const double* __restrict__ yfa,
                                                   const double* __restrict__ dfx, const double* __restrict__ dfy, const double* kdbl,
                                                   double* __restrict__ zn) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  const int k = blockIdx.z;   // 0..km
  if (i < L.is || i > L.ie || j < L.js || j > L.je) return;
  const long long o = LIDX(L, i, j) + (long long)k * L.plane;
  const double ar = __ldg(G.area + LIDX(L, i, j));
  const double rax = ar + xfa[o] - xfa[o + 1], ray = ar + yfa[o] - yfa[o + L.NI];
  double z = (__ldg(zh + o) * ar + fx[o] - fx[o + 1] + fy[o] - fy[o + L.NI]) / (rax + ray - ar);
  if (kdbl[KD_DZ * (L.npz + 1) + k] != 0.) z = z + (dfx[o] - dfx[o + 1] + dfy[o] - dfy[o + L.NI]) * __ldg(G.rarea + LIDX(L, i, j));
  zn[o] = z;
}

// ---- small column / pointwise helpers -------------------------------------------------------
__global__ void __launch_bounds__(CB) k_pk3_halo(Lay L, const double* __restrict__ delp, double* __restrict__ pk3, double ptop, double akap) {
  // ring cells: 2-wide frame around the compute domain excluding ... (dyn_core.F90:1405-1445)
  const int n = L.ie - L.is + 1;
  const int nring = 4 * n + 4 * (n + 4);   // 2 cols x n rows x 2 sides + 2 rows x (n+4) x 2 sides
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nring) return;
  int i, j;
  if (t < 4 * n) {          // west/east columns, j in [js, je]
    const int c = t / n; j = L.js + t % n;
    i = (c == 0) ? L.is - 2 : (c == 1) ? L.is - 1 : (c == 2) ? L.ie + 1 : L.ie + 2;
  } else {                  // south/north rows, i in [is-2, ie+2]
    const int t2 = t - 4 * n; const int r = t2 / (n + 4); i = L.is - 2 + t2 % (n + 4);
    j = (r == 0) ? L.js - 2 : (r == 1) ? L.js - 1 : (r == 2) ? L.je + 1 : L.je + 2;
  }
  const long long o = LIDX(L, i, j);
  double pe = ptop;
  for (int k = 1; k <= L.npz; k++) {
    pe = pe + __ldg(delp + o + (long long)(k - 1) * L.plane);
    pk3[o + (long long)k * L.plane] = exp(akap * log(pe));
  }
}
__global__ void __launch_bounds__(CB) k_pe_halo(Lay L, const double* __restrict__ delp, double* __restrict__ pe, double ptop) {
  const int n = L.ie - L.is + 1;
  const int nring = 2 * n + 2 * (n + 2);   // dyn_core.F90:1507-1524
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nring) return;
  int i, j;
  if (t < 2 * n) { j = L.js + t % n; i = (t / n == 0) ? L.is - 1 : L.ie + 1; }
  else { const int t2 = t - 2 * n; i = L.is - 1 + t2 % (n + 2); j = (t2 / (n + 2) == 0) ? L.js - 1 : L.je + 1; }
  const long long o = LIDX(L, i, j);
  double p = ptop;
  pe[o] = ptop;
  for (int k = 1; k <= L.npz; k++) {
    p = p + __ldg(delp + o + (long long)(k - 1) * L.plane);
    pe[o + (long long)k * L.plane] = p;
  }
}
__global__ void __launch_bounds__(TI* TJ) k_gz_from_zh(Lay L, const double* __restrict__ zh, double* __restrict__ gz, double grav) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  if (i < L.is - 2 || i > L.ie + 2 || j < L.js - 2 || j > L.je + 2) return;
  const long long o = LIDX(L, i, j) + (long long)blockIdx.z * L.plane;
  gz[o] = __ldg(zh + o) * grav;
}
__global__ void __launch_bounds__(CB) k_gz_init(Lay L, const double* __restrict__ phis, const double* __restrict__ delz, double* __restrict__ gz,
                                               double rgrav) {
  COL_SETUP(L.is, L.ie, L.js, L.je)
  double g = __ldg(phis + o) * rgrav;
  gz[o + (long long)L.npz * P] = g;
  for (int k = L.npz; k >= 1; k--) { g = g - __ldg(delz + o + (long long)(k - 1) * P); gz[o + (long long)(k - 1) * P] = g; }
}

// =============================================================================================
__global__ void k_zs(Lay L, const double* __restrict__ phis, double* __restrict__ zs, double rgrav);
static SolverIn make_solver(fv3_ctx* c, double dt, int d_grid) {
  SolverIn S;
  const fv3_flags_t& f = c->f;
  S.dt = dt; S.rgrav = 1.0 / f.grav; S.rdgas = f.rdgas; S.akap = f.kappa; S.ptop = f.ptop; S.p_fac = f.p_fac; S.a_imp = f.a_imp;
  S.km = c->L.npz; S.use_cond = f.use_cond; S.moist_kappa = f.moist_kappa && (d_grid || f.use_cond); S.d_grid = d_grid;
  S.q_con = c->fld[FV3_QCON]; S.cappa = c->fld[FV3_CAPPA];
  S.pm2 = c->scr[0]; S.gam = c->scr[1]; S.pp = c->scr[2]; S.w2 = c->scr[3];
  return S;
}

int stage_update_dz_c(fv3_ctx* c, double dt2) {
  StageScope ts(c, "UPDATE_DZ_C");
  const Lay& L = c->L;
  dim3 blk(TI, TJ), grd((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, L.npz + 1);
  double* gzn = c->scr[0];
  k_dzc_adv<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_UT], c->fld[FV3_VT], c->fld[FV3_GZ], gzn, c->d_dp_ref);
  const int n = L.ie - L.is + 3;
  k_dz_clamp<<<col_blocks(n, n), CB, 0, c->stream>>>(L, gzn, c->fld[FV3_GZ], c->fld[FV3_PHIS], c->fld[FV3_WS3], 1.0 / c->f.grav, 1.0 / dt2, 1);
  c->launches += 2;
  return 0;
}

int stage_riem_solver_c(fv3_ctx* c, double dt2) {
  StageScope ts(c, "Riem_Solver_C");
  const Lay& L = c->L;
  if (c->f.a_imp <= 0.5) return fv3_fail(c, -2, "Riem_Solver_c: a_imp <= 0.5 (RIM_2D / SIM3p0) not supported");
  if (c->f.fast_tau_w_sec > 1.e-5) return fv3_fail(c, -2, "Riem_Solver_c: fast_tau_w_sec not supported");
  SolverIn S = make_solver(c, dt2, 0);
  S.delp = c->fld[FV3_DELPC]; S.pt = c->fld[FV3_PTC]; S.hgt = c->fld[FV3_GZ]; S.w = c->fld[FV3_OMGA]; S.ws = c->fld[FV3_WS3];
  S.pe = c->fld[FV3_PKC];   // perturbation pressure accumulates in pef's storage, pem added last
  const int n = L.ie - L.is + 3;
  k_riem_c<<<col_blocks(n, n), CB, 0, c->stream>>>(L, S, c->fld[FV3_PHIS], c->fld[FV3_GZ], c->fld[FV3_PKC], c->f.grav);
  c->launches++;
  return 0;
}

int stage_riem_solver3(fv3_ctx* c, double dt, int last_call) {
  StageScope ts(c, "Riem_Solver3");
  const Lay& L = c->L;
  if (c->f.a_imp <= 0.5) return fv3_fail(c, -2, "Riem_Solver3: a_imp <= 0.5 (RIM_2D / SIM3) not supported");
  if (c->f.fast_tau_w_sec > 1.e-5 || c->f.d2bg_zq > 0.0001) return fv3_fail(c, -2, "Riem_Solver3: fast_tau_w_sec / d2bg_zq not supported");
  SolverIn S = make_solver(c, dt, 1);
  S.delp = c->fld[FV3_DELP]; S.pt = c->fld[FV3_PT]; S.hgt = c->fld[FV3_ZH]; S.w = c->fld[FV3_W]; S.ws = c->fld[FV3_WS];
  S.pe = c->scr[4];
  const int n = L.ie - L.is + 1;
  // zs = phis*rgrav (dyn_core.F90:247-251) into scr[5]
  {
    dim3 blk2(TI, TJ), grd2((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, 1);
    k_zs<<<grd2, blk2, 0, c->stream>>>(L, c->fld[FV3_PHIS], c->scr[5], 1.0 / c->f.grav);
    c->launches++;
  }
  k_riem3<<<col_blocks(n, n), CB, 0, c->stream>>>(L, S, c->scr[5], c->fld[FV3_ZH], c->fld[FV3_W], c->fld[FV3_DELZ], c->fld[FV3_PKC],
                                                 c->fld[FV3_PK3], c->fld[FV3_PK], c->fld[FV3_PE], c->fld[FV3_PELN], last_call,
                                                 c->f.beta < -0.1 ? 1 : 0, c->f.use_logp);
  c->launches++;
  return 0;
}

// zs plane helper (dyn_core.F90:247-251)
__global__ void __launch_bounds__(TI* TJ) k_zs(Lay L, const double* __restrict__ phis, double* __restrict__ zs, double rgrav) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  if (i < L.isd || i > L.ied || j > L.jed) return;
  zs[LIDX(L, i, j)] = __ldg(phis + LIDX(L, i, j)) * rgrav;
}


int stage_update_dz_d(fv3_ctx* c, double dt) {
  StageScope ts(c, "UPDATE_DZ");
  const Lay& L = c->L;
  const int km = L.npz, n1 = km + 1;
  if (!ppm::hord_supported(c->f.hord_tm)) return fv3_fail(c, -2, "update_dz_d: unsupported hord_tm");
  // damp(km+1) = damp(km), ndif(km+1) = ndif(km)  (nh_utils.F90:240-241); tables set by the d_sw prologue
  c->damp_vt[km] = c->damp_vt[km - 1]; c->nord_v[km] = c->nord_v[km - 1];
  std::vector<int> ki(n1); std::vector<double> kd(n1);
  bool any = false;
  for (int k = 0; k < n1; k++) { ki[k] = c->nord_v[k]; kd[k] = c->damp_vt[k] > 1.E-5 ? c->damp_vt[k] : 0.; any |= kd[k] != 0.; }
  FV3_CUDA(c, cudaMemcpyAsync(c->d_kint + KI_NORD_V * n1, ki.data(), n1 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  FV3_CUDA(c, cudaMemcpyAsync(c->d_kdbl + KD_DZ * n1, kd.data(), n1 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double *crxa = c->scr[0], *xfxa = c->scr[1], *crya = c->scr[2], *yfxa = c->scr[3], *gam = c->scr[4];
  double *fx = c->scr[5], *fy = c->scr[6], *fx2 = c->scr[7], *fy2 = c->scr[8], *q_i = c->scr[9], *q_j = c->scr[10];
  double *zn = c->scr[11], *dfx = c->scr[12], *dfy = c->scr[13], *d2 = c->scr[14];
  const int nix = L.ie + 1 - L.is + 1, njx = L.jed - L.jsd + 1, niy = L.ied - L.isd + 1, njy = L.je + 1 - L.js + 1;
  k_edge_profile<<<col_blocks(nix, njx), CB, 0, c->stream>>>(L, c->fld[FV3_CRX], c->fld[FV3_XFX], crxa, xfxa, gam, c->d_dp_ref, L.is, L.ie + 1, L.jsd, L.jed);
  k_edge_profile<<<col_blocks(niy, njy), CB, 0, c->stream>>>(L, c->fld[FV3_CRY], c->fld[FV3_YFX], crya, yfxa, gam, c->d_dp_ref, L.isd, L.ied, L.js, L.je + 1);
  c->launches += 2;
  Tp2d tp;
  tp.q = c->fld[FV3_ZH]; tp.crx = crxa; tp.cry = crya; tp.xfx = xfxa; tp.yfx = yfxa; tp.ra_x = nullptr; tp.ra_y = nullptr;
  tp.fx = fx; tp.fy = fy; tp.mfx = nullptr; tp.mfy = nullptr; tp.hord = c->f.hord_tm; tp.nk = n1;
  tp.fx2 = fx2; tp.fy2 = fy2; tp.q_i = q_i; tp.q_j = q_j;
  int rc = launch_tp2d(c, tp); if (rc) return rc;
  if (any) {
    Deln dl;
    dl.q = c->fld[FV3_ZH]; dl.fx2 = dfx; dl.fy2 = dfy; dl.d2 = d2; dl.slot_nord = KI_NORD_V; dl.slot_damp = KD_DZ; dl.thresh = 0;
    dl.premul = 1; dl.nk = n1; dl.nord_const = 0; dl.damp_const = 0;
    launch_deln(c, dl);
  }
  dim3 blk(TI, TJ), grd((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, n1);
  k_dzd_upd<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_ZH], fx, fy, xfxa, yfxa, dfx, dfy, c->d_kdbl, zn);
  const int n = L.ie - L.is + 1;
  k_dz_clamp<<<col_blocks(n, n), CB, 0, c->stream>>>(L, zn, c->fld[FV3_ZH], c->fld[FV3_PHIS], c->fld[FV3_WS], 1.0 / c->f.grav, 1.0 / dt, 0);
  c->launches += 2;
  return 0;
}

int stage_pk3_halo(fv3_ctx* c) {
  const Lay& L = c->L;
  if (c->f.use_logp) return fv3_fail(c, -2, "pln_halo (use_logp) not supported");
  const int n = L.ie - L.is + 1, nring = 4 * n + 4 * (n + 4);
  k_pk3_halo<<<(nring + CB - 1) / CB, CB, 0, c->stream>>>(L, c->fld[FV3_DELP], c->fld[FV3_PK3], c->f.ptop, c->f.kappa);
  c->launches++;
  return 0;
}
int stage_pe_halo(fv3_ctx* c) {
  const Lay& L = c->L;
  const int n = L.ie - L.is + 1, nring = 2 * n + 2 * (n + 2);
  k_pe_halo<<<(nring + CB - 1) / CB, CB, 0, c->stream>>>(L, c->fld[FV3_DELP], c->fld[FV3_PE], c->f.ptop);
  c->launches++;
  return 0;
}
int stage_gz_from_zh(fv3_ctx* c) {
  const Lay& L = c->L;
  dim3 blk(TI, TJ), grd((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, L.npz + 1);
  k_gz_from_zh<<<grd, blk, 0, c->stream>>>(L, c->fld[FV3_ZH], c->fld[FV3_GZ], c->f.grav);
  c->launches++;
  return 0;
}
int stage_gz_init(fv3_ctx* c) {
  const Lay& L = c->L;
  const int n = L.ie - L.is + 1;
  k_gz_init<<<col_blocks(n, n), CB, 0, c->stream>>>(L, c->fld[FV3_PHIS], c->fld[FV3_DELZ], c->fld[FV3_GZ], 1.0 / c->f.grav);
  c->launches++;
  return 0;
}
int stage_copy_field(fv3_ctx* c, int dst, int src) {
  if (dst < 0 || src < 0 || dst >= FV3_NUM_FIELDS || src >= FV3_NUM_FIELDS || c->dim[dst].nk != c->dim[src].nk) return fv3_fail(c, -1, "copy_field: bad ids");
  FV3_CUDA(c, cudaMemcpyAsync(c->fld[dst], c->fld[src], (size_t)c->L.plane * c->dim[dst].nk * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int stage_zero_field(fv3_ctx* c, int f) {
  if (f < 0 || f >= FV3_NUM_FIELDS) return fv3_fail(c, -1, "zero_field: bad id");
  FV3_CUDA(c, cudaMemsetAsync(c->fld[f], 0, (size_t)c->L.plane * c->dim[f].nk * sizeof(double), c->stream));
  return 0;
}
