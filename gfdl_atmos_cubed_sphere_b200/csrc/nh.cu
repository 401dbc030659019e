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
  double *pm2, *gam, *pp, *w2;   // scratch planes [km+1][NJ][NI]
  double dt, rgrav, rdgas, akap, ptop, p_fac, a_imp;
  int km, use_cond, moist_kappa, d_grid;
  // appended (keeps the constant-bank offsets of everything above): Rayleigh damping of w, rff(1:k_rf) (nh_utils.F90:356-368)
  const double* rff; int k_rf;
};

// One column of SIM1_solver (a_imp > 0.999) or SIM_solver, in SIX sweeps over k with the loads of NB levels issued
// together (the first version made ten sweeps with one dependent load chain per level: ncu 78 % long-scoreboard stalls,
// 3.4 GB of traffic per call at C384L79):
//   A  down: pm2, p_gas' (registers only) and the forward elimination of the spline system for pp, one level behind
//            (nh_utils.F90:377-447, :1297-1332)
//   B  up:   back-substitution of pp                                                   (:1328-1332)
//   C  down: forward elimination of the w system, one level behind the loads          (:1335-1361 / :1463-1496)
//   D  up:   back-substitution of w                                                   (:1357-1361)
//   E  down: perturbation pressure pe (:1373-1380 / :1508-1516, :1531-1535); emit(k, pe_final, pem, w2) hands every
//            interface to the caller (pef / ppe, pk3, pe, pk, peln, w); the raw pe goes to the storage of pp
//   F  up:   dz2 from the spline of pe (:1382-1392); out_dz(k, dz2) for k = km..1
// The arithmetic of every statement is the reference's, in the reference's order.
#ifndef RIEM_NB
#define RIEM_NB 2
#endif
#ifndef RIEM_MINB
#define RIEM_MINB 8
#endif
constexpr int NB = RIEM_NB;
// RF: fast_tau_w_sec > 0 -- w2(k <= k_rf) *= rff(k) after the back-substitution (nh_utils.F90:1363-1371, :1498-1506), applied where
// sweep E loads the final w2.  A template flag so that the default instantiations keep their exact code.
template <bool MK, bool UC, bool RF, class IfaceSink, class DzSink>
__device__ __forceinline__ void solve_column(const SolverIn& S, long long o, long long P, IfaceSink emit, DzSink out_dz) {
  const int km = S.km;
  const bool sim1 = S.d_grid ? (S.a_imp > 0.999) : true;   // nh_utils.F90:450-459, nh_core.F90:169-185
  const double alpha = sim1 ? 1.0 : S.a_imp;
  const double beta = 1. - alpha, ra = 1. / alpha, t2 = beta / alpha;
  const double t1g = sim1 ? 2. * S.dt * S.dt : 2. * (alpha * S.dt) * (alpha * S.dt);
  const double rdt = 1. / S.dt, dt = S.dt;
  // array bases stay in the constant bank (kernel parameter S); only the column offset o lives in registers
  constexpr bool mk = MK, uc = UC;
#define LV(k) (o + (long long)((k)-1) * P)
#define delp S.delp
#define hgt S.hgt
#define pt S.pt
#define w1 S.w
#define qcon S.q_con
#define cappa S.cappa
#define pm2a S.pm2
#define gama S.gam
#define ppa S.pp
#define w2a S.w2
  // ---- A
  double ppk;   // pp(km+1) after the sweep
  {
    double pem = S.ptop, peg = S.ptop, lpem = 0., lpeg = 0.;
    if (S.d_grid) { lpem = log(S.ptop); lpeg = lpem; }
    double h_prev = __ldg(hgt + o);
    double dm_prev = 0., pe_prev = 0., g_rat = 0., bet = 1.;
    ppk = 0.;
    for (int k0 = 1; k0 <= km; k0 += NB) {
      double d[NB], h[NB], t[NB], qc[NB], cp[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int kk = min(k0 + u, km);
        d[u] = __ldg(delp + LV(kk)); h[u] = __ldg(hgt + o + (long long)kk * P); t[u] = __ldg(pt + LV(kk));
        qc[u] = uc ? __ldg(qcon + LV(kk)) : 0.;
        cp[u] = mk ? __ldg(cappa + LV(kk)) : S.akap;
      }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int k = k0 + u;
        if (k > km) break;
        const double dmr = d[u];
        const double pem1 = pem + dmr;
        double pm2;
        if (S.d_grid) {   // nh_core.F90:120-165 (logs of interface pressures)
          const double lpem1 = log(pem1);
          if (uc) {
            const double peg1 = peg + dmr * (1. - qc[u]);
            const double lpeg1 = log(peg1);
            pm2 = (peg1 - peg) / (lpeg1 - lpeg);
            peg = peg1; lpeg = lpeg1;
          } else pm2 = dmr / (lpem1 - lpem);
          lpem = lpem1;
        } else {          // nh_utils.F90:412-447
          if (uc) {
            const double peg1 = peg + dmr * (1. - qc[u]);
            pm2 = (peg1 - peg) / log(peg1 / peg);
            peg = peg1;
          } else pm2 = dmr / log(pem1 / pem);
        }
        pem = pem1;
        const double dm = dmr * S.rgrav;
        pm2a[LV(k)] = pm2;
        const double dz = h[u] - h_prev; h_prev = h[u];
        const double pe = exp((1. / (1. - cp[u])) * log(-dm / dz * S.rdgas * t[u])) - pm2;
        // spline system, step k-1 (needs levels k-1 and k)
        if (k == 2) {
          g_rat = dm_prev / dm;
          bet = 2. * (1. + g_rat);
          ppa[o] = 0.;
          ppk = 3. * (pe_prev + g_rat * pe) / bet;   // pp(2)
          ppa[o + P] = ppk;
        } else if (k >= 3) {
          const double gam = g_rat / bet;   // g_rat(k-2)/bet
          g_rat = dm_prev / dm;
          const double bb = 2. * (1. + g_rat), dd = 3. * (pe_prev + g_rat * pe);
          gama[LV(k - 1)] = gam;
          bet = bb - gam;
          ppk = (dd - ppk) / bet;
          ppa[LV(k)] = ppk;
        }
        dm_prev = dm; pe_prev = pe;
      }
    }
    const double gam = g_rat / bet;   // step km (bb = 2, dd = 3 pe(km))
    gama[LV(km)] = gam;
    bet = 2. - gam;
    ppk = (3. * pe_prev - ppk) / bet;
    ppa[LV(km + 1)] = ppk;
  }
  // ---- B
  {
    double nxt = ppk;
    for (int k0 = km; k0 >= 2; k0 -= NB) {
      double p[NB], g[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) { const int kk = max(k0 - u, 2); p[u] = ppa[LV(kk)]; g[u] = gama[LV(kk)]; }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int k = k0 - u;
        if (k < 2) break;
        nxt = p[u] - g[u] * nxt; ppa[LV(k)] = nxt;
      }
    }
  }
  // ---- C
  double w2p;   // w2(km) after the sweep
  {
    double pem = S.ptop;
    double h_prev = __ldg(hgt + o);
    double d_p = 0., w_p = 0., g_p = 0., z_p = 0., pp_p = 0.;   // level m-1
    double aa_k = 0., wk_k = 0., bet = 1.;
    w2p = 0.;
    for (int k0 = 1; k0 <= km; k0 += NB) {
      double d[NB], h[NB], wv[NB], pv[NB], cp[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int kk = min(k0 + u, km);
        d[u] = __ldg(delp + LV(kk)); h[u] = __ldg(hgt + o + (long long)kk * P); wv[u] = __ldg(w1 + LV(kk)); pv[u] = ppa[LV(kk)];
        cp[u] = mk ? __ldg(cappa + LV(kk)) : S.akap;
      }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int m = k0 + u;
        if (m > km) break;
        const double dz = h[u] - h_prev; h_prev = h[u];
        const double gm2 = 1. / (1. - cp[u]);
        if (m >= 2) {
          pem = pem + d_p;                            // pem(m) = ptop + sum_{l<m} delp(l)
          const double aa_n = t1g * 0.5 * (g_p + gm2) / (z_p + dz) * pem;   // aa(m)
          const double wk_n = sim1 ? 0. : t2 * aa_n * (w_p - wv[u]);        // wk(m)
          const double dmk = d_p * S.rgrav;
          if (m == 2) {                               // step k = 1
            bet = dmk - aa_n;
            w2p = sim1 ? (dmk * w_p + dt * pv[u]) / bet : (dmk * w_p + dt * pv[u] + wk_n) / bet;
          } else {                                    // step k = m-1
            const double gam = aa_k / bet;
            gama[LV(m - 1)] = gam;
            bet = dmk - (aa_k + aa_n + aa_k * gam);
            if (sim1) w2p = (dmk * w_p + dt * (pv[u] - pp_p) - aa_k * w2p) / bet;
            else w2p = (dmk * w_p + dt * (pv[u] - pp_p) + wk_n - wk_k - aa_k * w2p) / bet;
          }
          w2a[LV(m - 1)] = w2p;
          aa_k = aa_n; wk_k = wk_n;
        }
        d_p = d[u]; w_p = wv[u]; g_p = gm2; z_p = dz; pp_p = pv[u];
      }
    }
    // step k = km
    pem = pem + d_p;                                  // pem(km+1)
    const double dmk = d_p * S.rgrav;
    const double p1 = t1g * g_p / z_p * pem;
    const double gam = aa_k / bet;
    gama[LV(km)] = gam;
    bet = dmk - (aa_k + p1 + aa_k * gam);
    const double wsv = __ldg(S.ws + o);
    if (sim1) w2p = (dmk * w_p + dt * (ppk - pp_p) - p1 * wsv - aa_k * w2p) / bet;
    else w2p = (dmk * w_p + dt * (ppk - pp_p) - wk_k + p1 * (t2 * w_p - ra * wsv) - aa_k * w2p) / bet;
    w2a[LV(km)] = w2p;
  }
  // ---- D
  {
    double nxt = w2p;
    for (int k0 = km - 1; k0 >= 1; k0 -= NB) {
      double wv[NB], g[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) { const int kk = max(k0 - u, 1); wv[u] = w2a[LV(kk)]; g[u] = gama[LV(kk + 1)]; }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int k = k0 - u;
        if (k < 1) break;
        nxt = wv[u] - g[u] * nxt; w2a[LV(k)] = nxt;
      }
    }
  }
  // ---- E
  double pe_k1, pe_k2;   // raw pe(km), pe(km+1) for sweep F
  {
    double carry = 0., prevc = 0.;   // pe(k), pe(k-1)
    double pem = S.ptop;             // hydrostatic interface pressure pem(k)
    double pp_k = 0.;                // pp(k): pp(1) = 0
    for (int k0 = 1; k0 <= km; k0 += NB) {
      double d[NB], w2v[NB], wv[NB], pn[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int kk = min(k0 + u, km);
        d[u] = __ldg(delp + LV(kk)); w2v[u] = w2a[LV(kk)]; wv[u] = __ldg(w1 + LV(kk));
        if (RF && kk <= S.k_rf) w2v[u] = w2v[u] * __ldg(S.rff + kk - 1);
        pn[u] = sim1 ? 0. : ppa[LV(kk + 1)];
      }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int k = k0 + u;
        if (k > km) break;
        const double dm = d[u] * S.rgrav;
        double nxt;
        if (sim1) nxt = carry + dm * (w2v[u] - wv[u]) * rdt;
        else nxt = carry + (dm * (w2v[u] - wv[u]) * rdt - beta * (pn[u] - pp_k)) * ra;
        ppa[LV(k)] = carry;                                           // raw pe(k)
        emit(k, sim1 ? carry : carry + beta * (pp_k - carry), pem, w2v[u]);
        pem = pem + d[u];
        prevc = carry; carry = nxt; pp_k = pn[u];
      }
    }
    ppa[LV(km + 1)] = carry;
    emit(km + 1, sim1 ? carry : carry + beta * (pp_k - carry), pem, 0.);
    pe_k1 = prevc; pe_k2 = carry;
  }
  // ---- F
  {
    double p1 = (pe_k1 + 2. * pe_k2) * r3;
    double dm_n = 0.;          // dm(k+1)
    double pe_a = pe_k1, pe_b = pe_k2;   // pe(k+1), pe(k+2) when level k is processed (k < km)
    for (int k0 = km; k0 >= 1; k0 -= NB) {
      double d[NB], t[NB], pm[NB], pe[NB], cp[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int kk = max(k0 - u, 1);
        d[u] = __ldg(delp + LV(kk)); t[u] = __ldg(pt + LV(kk)); pm[u] = pm2a[LV(kk)]; pe[u] = ppa[LV(kk)];
        cp[u] = mk ? __ldg(cappa + LV(kk)) : S.akap;
      }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const int k = k0 - u;
        if (k < 1) break;
        const double dm = d[u] * S.rgrav;
        if (k < km) {
          const double g_rat = dm / dm_n, bb = 2. * (1. + g_rat);
          p1 = (pe[u] + bb * pe_a + g_rat * pe_b) * r3 - g_rat * p1;
          pe_b = pe_a; pe_a = pe[u];
        }
        out_dz(k, -dm * S.rdgas * t[u] * exp((cp[u] - 1.) * log(fmax(S.p_fac * pm[u], p1 + pm[u]))));
        dm_n = dm;
      }
    }
  }
#undef LV
#undef delp
#undef hgt
#undef pt
#undef w1
#undef qcon
#undef cappa
#undef pm2a
#undef gama
#undef ppa
#undef w2a
}
}  // namespace

// ---- Riem_Solver_c (nh_utils.F90:323-480) on columns [is-1, ie+1]^2 -------------------------
// 8 CTAs x 128 threads per SM: all columns of a C384 face are resident in ONE wave (72-88 registers gave 1.56 waves)
template <bool MK, bool UC, bool RF>
__global__ void __launch_bounds__(CB, RIEM_MINB) k_riem_c(Lay L, const __grid_constant__ SolverIn S, const double* __restrict__ hs, double* __restrict__ gz,
                                               double* __restrict__ pef, double grav) {
  COL_SETUP(L.is - 1, L.ie + 1, L.js - 1, L.je + 1)
  const int km = S.km;
  double gzk = __ldg(hs + o);   // gz(km+1) = hs
  // heights are read from gz (input, m) by the sweeps A and C; gz is overwritten bottom-up by the last sweep only,
  // in the same backward order as nh_utils.F90:468-476
  solve_column<MK, UC, RF>(S, o, P,
               [&](int k, double pe2, double pem, double) {   // pef = pe2 + pem (nh_utils.F90:461-465), top = ptop
                 pef[o + (long long)(k - 1) * P] = (k == 1) ? S.ptop : pe2 + pem;
               },
               [&](int k, double dz2) {
                 if (k == km) gz[o + (long long)km * P] = gzk;
                 gzk = gzk - dz2 * grav;
                 gz[o + (long long)(k - 1) * P] = gzk;
               });
}

// ---- Riem_Solver3 (nh_core.F90:47-241) on columns [is, ie]x[js, je] -------------------------
template <bool MK, bool UC, bool RF>
__global__ void __launch_bounds__(CB, RIEM_MINB) k_riem3(Lay L, const __grid_constant__ SolverIn S, const double* __restrict__ zs, double* __restrict__ zh,
                                              double* __restrict__ w, double* __restrict__ delz, double* __restrict__ ppe,
                                              double* __restrict__ pk3, double* __restrict__ pk, double* __restrict__ pe,
                                              double* __restrict__ peln, int last_call, int fp_out, int use_logp) {
  COL_SETUP(L.is, L.ie, L.js, L.je)
  const int km = S.km;
  double zk = __ldg(zs + o);
  const double peln1 = log(S.ptop);
  const double ptk = exp(S.akap * peln1);
  solve_column<MK, UC, RF>(S, o, P,
               [&](int k, double pe2, double pem, double w2) {   // w, pk3, ppe (+ pe, pk, peln on the last call)
                 const long long ok = o + (long long)(k - 1) * P;
                 double pl, pkv;
                 if (k == 1) { pl = peln1; pkv = ptk; }
                 else { pl = log(pem); pkv = exp(S.akap * pl); }
                 if (last_call) { peln[ok] = pl; pk[ok] = pkv; pe[ok] = pem; }
                 ppe[ok] = fp_out ? pe2 + pem : pe2;
                 pk3[ok] = (use_logp && k >= 2) ? pl : pkv;
                 if (k <= km) w[ok] = w2;
               },
               [&](int k, double dz2) {
                 if (k == km) zh[o + (long long)km * P] = zk;
                 delz[o + (long long)(k - 1) * P] = dz2;
                 zk = zk - dz2;
                 zh[o + (long long)(k - 1) * P] = zk;
               });
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
  if (L.cube && i >= 2 && i <= L.npx - 2 && j >= 2 && j <= L.npy - 2) {
    // points whose 5-point cross touches no face-corner remap (all but a 2-wide ring): plain offsets from one index.  The general
    // path below evaluates the fill_4corners remap tests on every access and made the kernel instruction bound (358 warp
    // instructions per thread, 80 % issue-active: profiles/r2_dsw_summary.md)
    const int o2 = LIDX(L, i, j), NI = L.NI;
    const double* g = gzk + o2;
    double x0, x1, y0, y1;
    if (k == 1 || k == km + 1) {
      const int ka = (k == 1) ? 1 : km, kb = (k == 1) ? 2 : km - 1;
      const double r0 = (k == 1) ? dp0[0] / (dp0[0] + dp0[1]) : dp0[km - 1] / (dp0[km - 2] + dp0[km - 1]);
      const double* ua = ut + o2 + (long long)(ka - 1) * P; const double* ub = ut + o2 + (long long)(kb - 1) * P;
      const double* va = vt + o2 + (long long)(ka - 1) * P; const double* vb = vt + o2 + (long long)(kb - 1) * P;
      const double a0 = __ldg(ua), a1 = __ldg(ua + 1), c0 = __ldg(va), c1 = __ldg(va + NI);
      x0 = a0 + (a0 - __ldg(ub)) * r0; x1 = a1 + (a1 - __ldg(ub + 1)) * r0;
      y0 = c0 + (c0 - __ldg(vb)) * r0; y1 = c1 + (c1 - __ldg(vb + NI)) * r0;
    } else {
      const double r0 = 1. / (dp0[k - 2] + dp0[k - 1]), da = dp0[k - 1], db = dp0[k - 2];
      const double* ua = ut + o2 + (long long)(k - 2) * P; const double* ub = ua + P;
      const double* va = vt + o2 + (long long)(k - 2) * P; const double* vb = va + P;
      x0 = (da * __ldg(ua) + db * __ldg(ub)) * r0; x1 = (da * __ldg(ua + 1) + db * __ldg(ub + 1)) * r0;
      y0 = (da * __ldg(va) + db * __ldg(vb)) * r0; y1 = (da * __ldg(va + NI) + db * __ldg(vb + NI)) * r0;
    }
    const double gW = __ldg(g - 1), gC = __ldg(g), gE = __ldg(g + 1), gS = __ldg(g - NI), gN = __ldg(g + NI);
    const double ar = __ldg(G.area + o2);
    const double fx0 = x0 * ((x0 > 0.) ? gW : gC), fx1 = x1 * ((x1 > 0.) ? gC : gE);
    const double fy0 = y0 * ((y0 > 0.) ? gS : gC), fy1 = y1 * ((y1 > 0.) ? gC : gN);
    gzn[o2 + (long long)(k - 1) * P] = (gC * ar + fx0 - fx1 + fy0 - fy1) / (ar + x0 - x1 + y0 - y1);
    return;
  }
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
  // both upwind candidates of every face are loaded before the winds are known: one memory round trip instead of two
  const double gW = GX(i - 1, j), gCx = GX(i, j), gE = GX(i + 1, j), gS = GY(i, j - 1), gC = GY(i, j), gN = GY(i, j + 1);
  const double ar = __ldg(G.area + LIDX(L, i, j));
  const double fx0 = x0 * ((x0 > 0.) ? gW : gCx);
  const double fx1 = x1 * ((x1 > 0.) ? gCx : gE);
  const double fy0 = y0 * ((y0 > 0.) ? gS : gC);
  const double fy1 = y1 * ((y1 > 0.) ? gC : gN);
  // gz2(i,j) at the centre is the array after BOTH fills (dir=2 last), nh_utils.F90:163,177
  gzn[LIDX(L, i, j) + (long long)(k - 1) * P] = (gC * ar + fx0 - fx1 + fy0 - fy1) / (ar + x0 - x1 + y0 - y1);
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
// edge_profile (nh_utils.F90:1638-1672, non-uniform branch, limiter = 0) for a pair of fields.
// The tridiagonal coefficients depend on dp0 only, i.e. they are the same for every column: a one-thread kernel builds
// the per-level tables once per context (gk, bet, gam + the four boundary scalars) instead of every column recomputing
// them and storing gam as a full 3-D plane.  The two sweeps load NB levels at a time.
// table layout: T[0..km) = gk(k), T[km..2km) = bet(k), T[2km..3km) = gam(k) (k = 1..km at index k-1), then
// T[3km+0] = xt1 (top), +1 = a_bot, +2 = xt1 (bottom), +3 = xt2
__global__ void k_edge_tables(const double* __restrict__ dp0, double* __restrict__ T, int km) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double g0 = dp0[1] / dp0[0];
  double bet = g0 * (g0 + 0.5);
  double gm = (1. + g0 * (g0 + 1.5)) / bet;
  T[3 * km + 0] = 2. * g0 * (g0 + 1.);
  T[0] = g0; T[km] = bet; T[2 * km] = gm;
  double gk = 0.;
  for (int k = 2; k <= km; k++) {
    gk = dp0[k - 2] / dp0[k - 1];
    bet = 2. + 2. * gk - gm;
    gm = gk / bet;
    T[k - 1] = gk; T[km + k - 1] = bet; T[2 * km + k - 1] = gm;
  }
  const double a_bot = 1. + gk * (gk + 1.5);
  T[3 * km + 1] = a_bot;
  T[3 * km + 2] = 2. * gk * (gk + 1.);
  T[3 * km + 3] = gk * (gk + 0.5) - a_bot * gm;
}
__global__ void __launch_bounds__(CB, 8) k_edge_profile(Lay L, const double* __restrict__ q1, const double* __restrict__ q2,
                                                    double* __restrict__ q1e, double* __restrict__ q2e, const double* __restrict__ T,
                                                    int i0, int i1, int j0, int j1) {
  COL_SETUP(i0, i1, j0, j1)
  const int km = L.npz;
  constexpr int NBE = 4;
#define LV(k) (o + (long long)((k)-1) * P)
  const double* __restrict__ GK = T; const double* __restrict__ BET = T + km; const double* __restrict__ GM = T + 2 * km;
  double e1, e2, a_prev, b_prev;   // a_prev, b_prev = Q1(k-1), Q2(k-1)
  {
    const double a1 = __ldg(q1 + LV(1)), a2 = __ldg(q1 + LV(2)), b1 = __ldg(q2 + LV(1)), b2 = __ldg(q2 + LV(2));
    const double xt1 = __ldg(T + 3 * km), bet = __ldg(BET);
    e1 = (xt1 * a1 + a2) / bet; e2 = (xt1 * b1 + b2) / bet;
    q1e[o] = e1; q2e[o] = e2;
    a_prev = a1; b_prev = b1;
  }
  for (int k0 = 2; k0 <= km; k0 += NBE) {
    double a[NBE], b[NBE];
#pragma unroll
    for (int u = 0; u < NBE; u++) { const int kk = min(k0 + u, km); a[u] = __ldg(q1 + LV(kk)); b[u] = __ldg(q2 + LV(kk)); }
#pragma unroll
    for (int u = 0; u < NBE; u++) {
      const int k = k0 + u;
      if (k > km) break;
      const double gk = __ldg(GK + k - 1), bet = __ldg(BET + k - 1);
      e1 = (3. * (a_prev + gk * a[u]) - e1) / bet;
      e2 = (3. * (b_prev + gk * b[u]) - e2) / bet;
      q1e[LV(k)] = e1; q2e[LV(k)] = e2;
      if (k < km) { a_prev = a[u]; b_prev = b[u]; }   // keep Q(km-1) for the bottom edge
    }
  }
  {
    const double a_bot = __ldg(T + 3 * km + 1), xt1 = __ldg(T + 3 * km + 2), xt2 = __ldg(T + 3 * km + 3);
    const double akm = __ldg(q1 + LV(km)), bkm = __ldg(q2 + LV(km));
    e1 = (xt1 * akm + a_prev - a_bot * e1) / xt2;
    e2 = (xt1 * bkm + b_prev - a_bot * e2) / xt2;
    q1e[LV(km + 1)] = e1; q2e[LV(km + 1)] = e2;
  }
  for (int k0 = km; k0 >= 1; k0 -= NBE) {
    double a[NBE], b[NBE];
#pragma unroll
    for (int u = 0; u < NBE; u++) { const int kk = max(k0 - u, 1); a[u] = q1e[LV(kk)]; b[u] = q2e[LV(kk)]; }
#pragma unroll
    for (int u = 0; u < NBE; u++) {
      const int k = k0 - u;
      if (k < 1) break;
      const double g = __ldg(GM + k - 1);
      e1 = a[u] - g * e1; e2 = b[u] - g * e2;
      q1e[LV(k)] = e1; q2e[LV(k)] = e2;
    }
  }
#undef LV
}
// ---- small column / pointwise helpers -------------------------------------------------------
// LOGP: pln_halo (dyn_core.F90:1449-1496, use_logp) instead of pk3_halo (:1395-1447)
template <bool LOGP>
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
    pk3[o + (long long)k * L.plane] = LOGP ? log(pe) : exp(akap * log(pe));
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
  S.rff = nullptr; S.k_rf = 0;
  return S;
}

// nh_utils.F90:356-368: the Rayleigh table is set up ONCE, by the first solver call of the context, with that call's dt (the
// reference keeps it in SAVEd module variables); pfull as fv_dynamics.F90:276-280 with p_ref = 1.e5 (fv_arrays.F90 default)
static int rayleigh_setup(fv3_ctx* c, double dt, SolverIn& S) {
  const fv3_flags_t& f = c->f;
  if (!(f.fast_tau_w_sec > 1.e-5)) return 0;
  const int km = c->L.npz;
  if (!c->d_rff) {
    std::vector<double> rff(km, 1.0);
    c->k_rf = 0;
    for (int k = 1; k <= km; k++) {
      const double ph1 = c->ak[k - 1] + c->bk[k - 1] * 1.e5, ph2 = c->ak[k] + c->bk[k] * 1.e5;
      const double pfull = (ph2 - ph1) / log(ph2 / ph1);
      if (pfull > f.rf_cutoff) break;
      c->k_rf = k;
      const double sn = sin(0.5 * f.pi * log(f.rf_cutoff / pfull) / log(f.rf_cutoff / f.ptop));
      rff[k - 1] = 1.0 / (1.0 + dt / f.fast_tau_w_sec * (sn * sn));
    }
    FV3_CUDA(c, cudaMalloc(&c->d_rff, sizeof(double) * km));
    FV3_CUDA(c, cudaMemcpy(c->d_rff, rff.data(), sizeof(double) * km, cudaMemcpyHostToDevice));
  }
  S.rff = c->d_rff; S.k_rf = c->k_rf;
  return 0;
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
  SolverIn S = make_solver(c, dt2, 0);
  { int rc = rayleigh_setup(c, dt2, S); if (rc) return rc; }
  S.delp = c->fld[FV3_DELPC]; S.pt = c->fld[FV3_PTC]; S.hgt = c->fld[FV3_GZ]; S.w = c->fld[FV3_OMGA]; S.ws = c->fld[FV3_WS3];
  const int n = L.ie - L.is + 3;
#define RIEM_C(MK, UC)                                                                                                                    \
  do {                                                                                                                                    \
    if (S.rff) k_riem_c<MK, UC, true><<<col_blocks(n, n), CB, 0, c->stream>>>(L, S, c->fld[FV3_PHIS], c->fld[FV3_GZ], c->fld[FV3_PKC], c->f.grav); \
    else k_riem_c<MK, UC, false><<<col_blocks(n, n), CB, 0, c->stream>>>(L, S, c->fld[FV3_PHIS], c->fld[FV3_GZ], c->fld[FV3_PKC], c->f.grav);    \
  } while (0)
  if (S.moist_kappa) { if (S.use_cond) RIEM_C(true, true); else RIEM_C(true, false); }
  else { if (S.use_cond) RIEM_C(false, true); else RIEM_C(false, false); }
#undef RIEM_C
  c->launches++;
  return 0;
}

int stage_riem_solver3(fv3_ctx* c, double dt, int last_call) {
  StageScope ts(c, "Riem_Solver3");
  const Lay& L = c->L;
  if (c->f.a_imp <= 0.5) return fv3_fail(c, -2, "Riem_Solver3: a_imp <= 0.5 (RIM_2D / SIM3) not supported");
  if (c->f.d2bg_zq > 0.0001) return fv3_fail(c, -2, "Riem_Solver3: d2bg_zq (imp_diff_w) not supported");
  SolverIn S = make_solver(c, dt, 1);
  { int rc = rayleigh_setup(c, dt, S); if (rc) return rc; }
  S.delp = c->fld[FV3_DELP]; S.pt = c->fld[FV3_PT]; S.hgt = c->fld[FV3_ZH]; S.w = c->fld[FV3_W]; S.ws = c->fld[FV3_WS];
  const int n = L.ie - L.is + 1;
  // zs = phis*rgrav (dyn_core.F90:247-251) into scr[5]
  {
    dim3 blk2(TI, TJ), grd2((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, 1);
    k_zs<<<grd2, blk2, 0, c->stream>>>(L, c->fld[FV3_PHIS], c->scr[5], 1.0 / c->f.grav);
    c->launches++;
  }
#define RIEM_3A(MK, UC, RF)                                                                                                         \
  k_riem3<MK, UC, RF><<<col_blocks(n, n), CB, 0, c->stream>>>(L, S, c->scr[5], c->fld[FV3_ZH], c->fld[FV3_W], c->fld[FV3_DELZ], c->fld[FV3_PKC], \
                                                             c->fld[FV3_PK3], c->fld[FV3_PK], c->fld[FV3_PE], c->fld[FV3_PELN], last_call,      \
                                                             c->f.beta < -0.1 ? 1 : 0, c->f.use_logp)
#define RIEM_3(MK, UC) do { if (S.rff) RIEM_3A(MK, UC, true); else RIEM_3A(MK, UC, false); } while (0)
  if (S.moist_kappa) { if (S.use_cond) RIEM_3(true, true); else RIEM_3(true, false); }
  else { if (S.use_cond) RIEM_3(false, true); else RIEM_3(false, false); }
#undef RIEM_3
#undef RIEM_3A
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
  if (!ppm::hord_supported(c->f.hord_tm, c->f.lim_fac)) return fv3_fail(c, -2, "update_dz_d: unsupported hord_tm");
  // damp(km+1) = damp(km), ndif(km+1) = ndif(km)  (nh_utils.F90:240-241); tables set by the d_sw prologue
  c->damp_vt[km] = c->damp_vt[km - 1]; c->nord_v[km] = c->nord_v[km - 1];
  std::vector<int> ki(n1); std::vector<double> kd(n1);
  bool any = false;
  for (int k = 0; k < n1; k++) { ki[k] = c->nord_v[k]; kd[k] = c->damp_vt[k] > 1.E-5 ? c->damp_vt[k] : 0.; any |= kd[k] != 0.; }
  if (!c->capturing) {   // (see stage_d_sw)
    FV3_CUDA(c, cudaMemcpyAsync(c->d_kint + KI_NORD_V * n1, ki.data(), n1 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    FV3_CUDA(c, cudaMemcpyAsync(c->d_kdbl + KD_DZ * n1, kd.data(), n1 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  double *crxa = c->scr[0], *xfxa = c->scr[1], *crya = c->scr[2], *yfxa = c->scr[3];
  double *fx = c->scr[5], *fy = c->scr[6];
  double *zn = c->scr[11], *dfx = c->scr[12], *dfy = c->scr[13], *d2 = c->scr[14];
  const int nix = L.ie + 1 - L.is + 1, njx = L.jed - L.jsd + 1, niy = L.ied - L.isd + 1, njy = L.je + 1 - L.js + 1;
  if (!c->d_edge_tab) {
    FV3_CUDA(c, cudaMalloc(&c->d_edge_tab, sizeof(double) * (3 * km + 4)));
    k_edge_tables<<<1, 32, 0, c->stream>>>(c->d_dp_ref, c->d_edge_tab, km);
    c->launches++;
  }
  k_edge_profile<<<col_blocks(nix, njx), CB, 0, c->stream>>>(L, c->fld[FV3_CRX], c->fld[FV3_XFX], crxa, xfxa, c->d_edge_tab, L.is, L.ie + 1, L.jsd, L.jed);
  k_edge_profile<<<col_blocks(niy, njy), CB, 0, c->stream>>>(L, c->fld[FV3_CRY], c->fld[FV3_YFX], crya, yfxa, c->d_edge_tab, L.isd, L.ied, L.js, L.je + 1);
  c->launches += 2;
  if (any) {   // del-n damping fluxes of zh first: the transport epilogue consumes them
    Deln dl;
    dl.q = c->fld[FV3_ZH]; dl.fx2 = dfx; dl.fy2 = dfy; dl.d2 = d2; dl.slot_nord = KI_NORD_V; dl.slot_damp = KD_DZ; dl.thresh = 0;
    dl.premul = 1; dl.nk = n1; dl.nord_const = 0; dl.damp_const = 0;
    dl.k_lo = n1; dl.k_hi = -1; dl.nord_max = 0;
    for (int k = 0; k < n1; k++)
      if (kd[k] != 0.) { dl.k_lo = std::min(dl.k_lo, k); dl.k_hi = std::max(dl.k_hi, k); dl.nord_max = std::max(dl.nord_max, ki[k]); }
    launch_deln(c, dl);
  }
  Tp2d tp;
  tp.q = c->fld[FV3_ZH]; tp.crx = crxa; tp.cry = crya; tp.xfx = xfxa; tp.yfx = yfxa; tp.ra_x = nullptr; tp.ra_y = nullptr;
  tp.fx = fx; tp.fy = fy; tp.mfx = nullptr; tp.mfy = nullptr; tp.hord = c->f.hord_tm; tp.nk = n1;
  tp.zn = zn; tp.zn_dfx = dfx; tp.zn_dfy = dfy; tp.zn_slot = KD_DZ;   // height update fused into the transport epilogue
  int rc = launch_tp2d(c, tp); if (rc) return rc;
  const int n = L.ie - L.is + 1;
  k_dz_clamp<<<col_blocks(n, n), CB, 0, c->stream>>>(L, zn, c->fld[FV3_ZH], c->fld[FV3_PHIS], c->fld[FV3_WS], 1.0 / c->f.grav, 1.0 / dt, 0);
  c->launches += 1;
  return 0;
}

// ---- geopk (dyn_core.F90:2202-2356): hydrostatic pressure / geopotential integration, one thread per column --------
// C-grid call (cg): columns (is-1:ie+1)^2 from delpc, ptc; D-grid call: (is-2:ie+2)^2 from delp, pt, plus pkz.
// pe is stored on (is-1:ie+1)^2, peln on (is:ie)^2, as the reference's extents (:2211-2212).
template <bool UC>
__global__ void __launch_bounds__(CB) k_geopk(Lay L, const double* __restrict__ delp, const double* __restrict__ pt,
                                             const double* __restrict__ q_con, const double* __restrict__ hs, double* __restrict__ pk,
                                             double* __restrict__ gz, double* __restrict__ pe, double* __restrict__ peln,
                                             double* __restrict__ pkz, double ptop, double akap, double cp_air, int cg, int halo) {
  COL_SETUP(L.is - halo, L.ie + halo, L.js - halo, L.je + halo)
  const int km = L.npz;
  const bool in_pe = i >= L.is - 1 && i <= L.ie + 1 && j >= L.js - 1 && j <= L.je + 1;
  const bool in_c = i >= L.is && i <= L.ie && j >= L.js && j <= L.je;
  const double ptk = pow(ptop, akap), peln1 = log(ptop);
  double p1d = ptop, peg = ptop;
  pk[o] = ptk;
  if (in_c) peln[o] = peln1;
  if (in_pe) pe[o] = ptop;
  // top down (:2291-2318)
  for (int k = 2; k <= km + 1; k++) {
    const long long ok = o + (long long)(k - 1) * P;
    p1d = p1d + __ldg(delp + ok - P);
    const double lp = log(p1d);
    pk[ok] = exp(akap * lp);
    if (in_pe) pe[ok] = p1d;
    if (in_c) peln[ok] = lp;
  }
  // bottom up.  With condensate the integration uses pkg = peg^kappa of the condensate-free pressure: peg is re-accumulated
  // top down first (same additions in the same order as :2297-2298), stored in gz's own column as scratch, then consumed.
  double g = __ldg(hs + o);
  if (UC) {
    gz[o] = ptk;   // pkg(1)
    for (int k = 2; k <= km + 1; k++) {
      const long long ok = o + (long long)(k - 1) * P;
      peg = peg + __ldg(delp + ok - P) * (1. - __ldg(q_con + ok - P));
      gz[ok] = exp(akap * log(peg));
    }
    double pkg_hi = gz[o + (long long)km * P];
    gz[o + (long long)km * P] = g;
    for (int k = km; k >= 1; k--) {
      const long long ok = o + (long long)(k - 1) * P;
      const double pkg_lo = gz[ok];
      g = g + cp_air * __ldg(pt + ok) * (pkg_hi - pkg_lo);
      gz[ok] = g;
      pkg_hi = pkg_lo;
    }
  } else {
    gz[o + (long long)km * P] = g;
    double pk_hi = pk[o + (long long)km * P];
    for (int k = km; k >= 1; k--) {
      const long long ok = o + (long long)(k - 1) * P;
      const double pk_lo = pk[ok];
      g = g + cp_air * __ldg(pt + ok) * (pk_hi - pk_lo);
      gz[ok] = g;
      pk_hi = pk_lo;
    }
  }
  if (!cg && in_c) {
    double pk_lo = pk[o], ln_lo = peln[o];
    for (int k = 1; k <= km; k++) {
      const long long ok = o + (long long)(k - 1) * P;
      const double pk_hi = pk[ok + P], ln_hi = peln[ok + P];
      pkz[ok] = (pk_hi - pk_lo) / (akap * (ln_hi - ln_lo));
      pk_lo = pk_hi; ln_lo = ln_hi;
    }
  }
}
int stage_geopk(fv3_ctx* c, int cg) {
  StageScope ts(c, cg ? "GEOPK_C" : "GEOPK_D");
  const Lay& L = c->L;
  const int halo = cg ? 1 : 2, n = L.ie - L.is + 1 + 2 * halo;
  const double* delp = c->fld[cg ? FV3_DELPC : FV3_DELP];
  const double* pt = c->fld[cg ? FV3_PTC : FV3_PT];
  if (c->f.use_cond)
    k_geopk<true><<<col_blocks(n, n), CB, 0, c->stream>>>(L, delp, pt, c->fld[FV3_QCON], c->fld[FV3_PHIS], c->fld[FV3_PKC], c->fld[FV3_GZ],
                                                          c->fld[FV3_PE], c->fld[FV3_PELN], c->fld[FV3_PKZ], c->f.ptop, c->f.kappa,
                                                          c->f.cp_air, cg, halo);
  else
    k_geopk<false><<<col_blocks(n, n), CB, 0, c->stream>>>(L, delp, pt, c->fld[FV3_QCON], c->fld[FV3_PHIS], c->fld[FV3_PKC], c->fld[FV3_GZ],
                                                           c->fld[FV3_PE], c->fld[FV3_PELN], c->fld[FV3_PKZ], c->f.ptop, c->f.kappa,
                                                           c->f.cp_air, cg, halo);
  c->launches++;
  return 0;
}

int stage_pk3_halo(fv3_ctx* c) {
  const Lay& L = c->L;
  const int n = L.ie - L.is + 1, nring = 4 * n + 4 * (n + 4);
  if (c->f.use_logp) k_pk3_halo<true><<<(nring + CB - 1) / CB, CB, 0, c->stream>>>(L, c->fld[FV3_DELP], c->fld[FV3_PK3], c->f.ptop, c->f.kappa);
  else k_pk3_halo<false><<<(nring + CB - 1) / CB, CB, 0, c->stream>>>(L, c->fld[FV3_DELP], c->fld[FV3_PK3], c->f.ptop, c->f.kappa);
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
