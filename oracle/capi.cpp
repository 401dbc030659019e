// TEST INFRASTRUCTURE ONLY -- CPU oracle (see fv3_oracle.hpp header).
// extern "C" surface of the oracle, shaped like include/fv3_dyncore.h (prefix fv3o_) so the
// parity tests drive oracle and product through the same harness.  Fields are held in the
// reference's native Fortran layouts.  The per-stage drivers restate the call sites in
// model/dyn_core.F90 (:436-447 c_sw, :525 update_dz_c, :531 Riem_Solver_C, :562 p_grad_c,
// :666-812 d_sw with its per-k damping prologue, :911 update_dz_d, :932 Riem_Solver3,
// :953-989 halo pressure fills + gz, :1032 nh_p_grad).
#include "fv3_oracle.hpp"
#include <string>
#include <cstdio>
#include <map>
#include <omp.h>

using namespace fv3o;

struct FieldDim { int ilo, ni, jlo, nj, nk, kmid; };

struct fv3o_ctx {
  fv3_bounds_t b; fv3_grid_t g; fv3_flags_t f;
  std::vector<double> ak, bk;
  std::vector<std::vector<double>> fld;
  FieldDim dim[FV3_NUM_FIELDS];
  std::string err;
  // per-k damping arrays shared between d_sw and update_dz_d (dyn_core.F90:186-187)
  std::vector<double> damp_vt; std::vector<int> nord_v;
  // Rayleigh damping table of nh_utils (SAVEd rff, k_rf, RFw_initialized, nh_utils.F90:53-55)
  std::vector<double> rff; int k_rf = 0; bool rf_init = false;
  // external-mode damping term divg2(is:ie+1, js:je+1) of the current substep (dyn_core.F90:828-847; empty: d_ext = 0)
  // tracers: store[i] holds tracer i except for i == tracer_sel, which lives in fld[FV3_WORK_Q] (fv3o_select_tracer swaps)
  std::vector<std::vector<double>> store; int tracer_sel = 0;
  int tracer_fill = 0;   // flagstruct%fill: fillz after each remapped tracer (fv3o_set_tracer_fill)
  std::vector<double> pem;   // interface pressures before the last substep (omega diagnostic, dyn_core.F90:409-422)
  std::vector<double> divg2, ext_dpc;   // ext_dpc: delp at the cell corners (is:ie+1, js:je+1, npz), taken before d_sw (:745-747)
  explicit fv3o_ctx(const fv3_bounds_t& b_, const fv3_grid_t& g_, const fv3_flags_t& f_) : b(b_), g(g_), f(f_) {}
};

static void set_dims(fv3o_ctx* c) {
  const fv3_bounds_t& b = c->b;
  const int nia = b.ied - b.isd + 1, nja = b.jed - b.jsd + 1, nic = b.ie - b.is + 1, njc = b.je - b.js + 1, kz = b.npz;
  auto A = [&](int nk) { return FieldDim{b.isd, nia, b.jsd, nja, nk, 0}; };
  FieldDim* d = c->dim;
  d[FV3_U] = {b.isd, nia, b.jsd, nja + 1, kz, 0};
  d[FV3_V] = {b.isd, nia + 1, b.jsd, nja, kz, 0};
  d[FV3_W] = A(kz); d[FV3_PT] = A(kz); d[FV3_DELP] = A(kz); d[FV3_QCON] = A(kz); d[FV3_CAPPA] = A(kz);
  d[FV3_DELZ] = {b.is, nic, b.js, njc, kz, 0};
  d[FV3_PHIS] = A(1); d[FV3_OMGA] = A(kz); d[FV3_UA] = A(kz); d[FV3_VA] = A(kz);
  d[FV3_UC] = {b.isd, nia + 1, b.jsd, nja, kz, 0};
  d[FV3_VC] = {b.isd, nia, b.jsd, nja + 1, kz, 0};
  d[FV3_MFX] = {b.is, nic + 1, b.js, njc, kz, 0};
  d[FV3_MFY] = {b.is, nic, b.js, njc + 1, kz, 0};
  d[FV3_CX] = {b.is, nic + 1, b.jsd, nja, kz, 0};
  d[FV3_CY] = {b.isd, nia, b.js, njc + 1, kz, 0};
  d[FV3_DELPC] = A(kz); d[FV3_PTC] = A(kz); d[FV3_UT] = A(kz); d[FV3_VT] = A(kz);
  d[FV3_DIVGD] = {b.isd, nia + 1, b.jsd, nja + 1, kz, 0};
  d[FV3_CRX] = d[FV3_CX]; d[FV3_XFX] = d[FV3_CX]; d[FV3_CRY] = d[FV3_CY]; d[FV3_YFX] = d[FV3_CY];
  d[FV3_GZ] = A(kz + 1); d[FV3_ZH] = A(kz + 1); d[FV3_PKC] = A(kz + 1); d[FV3_PK3] = A(kz + 1);
  d[FV3_WS3] = A(1);
  d[FV3_WS] = {b.is, nic, b.js, njc, 1, 0};
  d[FV3_PE] = {b.is - 1, nic + 2, b.js - 1, njc + 2, kz + 1, 1};
  d[FV3_PELN] = {b.is, nic, b.js, njc, kz + 1, 1};
  d[FV3_PK] = {b.is, nic, b.js, njc, kz + 1, 0};
  d[FV3_PKZ] = {b.is, nic, b.js, njc, kz, 0};
  d[FV3_HEAT] = A(kz); d[FV3_DISS] = A(kz);
  d[FV3_WORK_Q] = A(kz);
  d[FV3_WORK_FX] = {b.is, nic + 1, b.js, njc, kz, 0};
  d[FV3_WORK_FY] = {b.is, nic, b.js, njc + 1, kz, 0};
  d[FV3_WORK_RAX] = {b.is, nic, b.jsd, nja, kz, 0};
  d[FV3_WORK_RAY] = {b.isd, nia, b.js, njc, kz, 0};
  d[FV3_DP1] = A(kz);
  d[FV3_DU] = d[FV3_U]; d[FV3_DV] = d[FV3_V];
}

static V3 F3(fv3o_ctx* c, int id) {
  const FieldDim& d = c->dim[id];
  return V3(c->fld[id].data(), d.ilo, d.ilo + d.ni - 1, d.jlo, d.jlo + d.nj - 1);
}
static V2 F2(fv3o_ctx* c, int id, int k = 1) { return F3(c, id).k(k); }

extern "C" {

int fv3o_create(const fv3_bounds_t* bd, const fv3_grid_t* grid, const fv3_flags_t* flags, int device, fv3o_ctx** out) {
  (void)device;
  fv3o_ctx* c = new fv3o_ctx(*bd, *grid, *flags);
  c->ak.assign(flags->ak, flags->ak + bd->npz + 1);
  c->bk.assign(flags->bk, flags->bk + bd->npz + 1);
  c->f.ak = c->ak.data(); c->f.bk = c->bk.data();
  set_dims(c);
  c->fld.resize(FV3_NUM_FIELDS);
  for (int i = 0; i < FV3_NUM_FIELDS; i++) {
    const FieldDim& d = c->dim[i];
    c->fld[i].assign((size_t)d.ni * d.nj * d.nk, 0.0);
  }
  c->damp_vt.assign(bd->npz + 1, 0.); c->nord_v.assign(bd->npz + 1, 0);
  *out = c;
  return 0;
}
void fv3o_destroy(fv3o_ctx* c) { delete c; }
const char* fv3o_last_error(const fv3o_ctx* c) { return c->err.c_str(); }
int fv3o_field_dims(const fv3o_ctx* c, int field, int dims[6]) {
  if (field < 0 || field >= FV3_NUM_FIELDS) return -1;
  const FieldDim& d = c->dim[field];
  dims[0] = d.ilo; dims[1] = d.ni; dims[2] = d.jlo; dims[3] = d.nj; dims[4] = d.nk; dims[5] = d.kmid;
  return 0;
}
int fv3o_put_field(fv3o_ctx* c, int field, const double* host) {
  if (field < 0 || field >= FV3_NUM_FIELDS) return -1;
  std::memcpy(c->fld[field].data(), host, c->fld[field].size() * sizeof(double));
  return 0;
}
int fv3o_get_field(fv3o_ctx* c, int field, double* host) {
  if (field < 0 || field >= FV3_NUM_FIELDS) return -1;
  std::memcpy(host, c->fld[field].data(), c->fld[field].size() * sizeof(double));
  return 0;
}
int fv3o_sync(fv3o_ctx*) { return 0; }
int fv3o_set_threads(int n) { omp_set_num_threads(n); return omp_get_max_threads(); }
int fv3o_max_threads(void) { return omp_get_max_threads(); }
// all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
int fv3o_use_all_cores(void) { omp_set_num_threads(omp_get_num_procs()); return omp_get_max_threads(); }

// tp_core.F90:85, batched over nk levels (same field conventions as fv3_fv_tp_2d)
int fv3o_fv_tp_2d(fv3o_ctx* c, int nk, int hord, int use_mfx, int use_mass, int nord, double damp_c) {
  Bd bd(c->b); Grid g(c->g, bd);
  V3 q = F3(c, FV3_WORK_Q), crx = F3(c, FV3_CRX), cry = F3(c, FV3_CRY), xfx = F3(c, FV3_XFX), yfx = F3(c, FV3_YFX);
  V3 rax = F3(c, FV3_WORK_RAX), ray = F3(c, FV3_WORK_RAY), fx = F3(c, FV3_WORK_FX), fy = F3(c, FV3_WORK_FY);
  V3 mfx = F3(c, FV3_MFX), mfy = F3(c, FV3_MFY), mass = F3(c, FV3_DELP);
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= nk; k++) {
    V2 mx = mfx.k(k), my = mfy.k(k), ms = mass.k(k);
    fv_tp_2d(q.k(k), crx.k(k), cry.k(k), bd.npx, bd.npy, hord, fx.k(k), fy.k(k), xfx.k(k), yfx.k(k), g, bd, rax.k(k),
             ray.k(k), c->f.lim_fac, use_mfx ? &mx : nullptr, use_mfx ? &my : nullptr, use_mass ? &ms : nullptr,
             nord >= 0, nord, damp_c);
  }
  return 0;
}

// dyn_core.F90:436-447
int fv3o_c_sw(fv3o_ctx* c, double dt2) {
  Bd bd(c->b); Grid g(c->g, bd);
  const bool hydro = c->f.hydrostatic != 0;
#pragma omp parallel for schedule(dynamic)
  for (int k = 1; k <= bd.npz; k++) {
    c_sw(F2(c, FV3_DELPC, k), F2(c, FV3_DELP, k), F2(c, FV3_PTC, k), F2(c, FV3_PT, k), F2(c, FV3_U, k), F2(c, FV3_V, k),
         F2(c, FV3_W, k), F2(c, FV3_UC, k), F2(c, FV3_VC, k), F2(c, FV3_UA, k), F2(c, FV3_VA, k), F2(c, FV3_OMGA, k),
         F2(c, FV3_UT, k), F2(c, FV3_VT, k), F2(c, FV3_DIVGD, k), c->f.nord, dt2, hydro, true, bd, g);
  }
  return 0;
}

static std::vector<double> dp_ref(fv3o_ctx* c) {  // dyn_core.F90:242-244
  std::vector<double> d(c->b.npz);
  for (int k = 0; k < c->b.npz; k++) d[k] = c->ak[k + 1] - c->ak[k] + (c->bk[k + 1] - c->bk[k]) * 1.E5;
  return d;
}
static std::vector<double> zs_of(fv3o_ctx* c) {  // dyn_core.F90:247-251
  const double rgrav = 1.0 / c->f.grav;
  std::vector<double> zs(c->fld[FV3_PHIS].size());
  for (size_t i = 0; i < zs.size(); i++) zs[i] = c->fld[FV3_PHIS][i] * rgrav;
  return zs;
}

// dyn_core.F90:525-527
int fv3o_update_dz_c(fv3o_ctx* c, double dt2) {
  Bd bd(c->b); Grid g(c->g, bd);
  std::vector<double> dp0 = dp_ref(c), zs = zs_of(c);
  update_dz_c(bd.is, bd.ie, bd.js, bd.je, bd.npz, bd.ng, dt2, dp0.data(), V2(zs.data(), bd.isd, bd.jsd, bd.ied - bd.isd + 1),
              g.area, F3(c, FV3_UT), F3(c, FV3_VT), F3(c, FV3_GZ), F2(c, FV3_WS3), bd);
  return 0;
}
// dyn_core.F90:531-536
// nh_utils.F90:356-368: set up once, by the first solver call, with that call's dt.  pfull as fv_dynamics.F90:276-280 with
// p_ref = 1.e5 (fv_arrays.F90 default).
static void rayleigh_init(fv3o_ctx* c, double dt) {
  if (c->rf_init || !(c->f.fast_tau_w_sec > 1.e-5)) return;
  c->rf_init = true;
  const int km = c->b.npz;
  c->rff.assign(km, 1.0); c->k_rf = 0;
  for (int k = 1; k <= km; k++) {
    const double ph1 = c->ak[k - 1] + c->bk[k - 1] * 1.e5, ph2 = c->ak[k] + c->bk[k] * 1.e5;
    const double pfull = (ph2 - ph1) / std::log(ph2 / ph1);
    if (pfull > c->f.rf_cutoff) break;
    c->k_rf = k;
    const double sn = std::sin(0.5 * c->f.pi * std::log(c->f.rf_cutoff / pfull) / std::log(c->f.rf_cutoff / c->f.ptop));
    c->rff[k - 1] = 1.0 / (1.0 + dt / c->f.fast_tau_w_sec * (sn * sn));
  }
}
int fv3o_riem_solver_c(fv3o_ctx* c, double dt2) {
  Bd bd(c->b);
  Consts k{c->f.rdgas, c->f.cp_air, c->f.grav, c->f.kappa, c->f.radius, c->f.omega, c->f.pi};
  const int ms = std::max(1, c->f.m_split / 2);
  rayleigh_init(c, dt2);
  if (c->rf_init) { k.rff = c->rff.data(); k.k_rf = c->k_rf; }
  riem_solver_c(ms, dt2, bd.is, bd.ie, bd.js, bd.je, bd.npz, bd.ng, c->f.kappa, F3(c, FV3_CAPPA), c->f.cp_air, c->f.ptop,
                F2(c, FV3_PHIS), F3(c, FV3_OMGA), F3(c, FV3_PTC), F3(c, FV3_QCON), F3(c, FV3_DELPC), F3(c, FV3_GZ),
                F3(c, FV3_PKC), F2(c, FV3_WS3), c->f.p_fac, c->f.a_imp, c->f.use_cond != 0, c->f.moist_kappa != 0, k);
  return 0;
}
// dyn_core.F90:562
int fv3o_p_grad_c(fv3o_ctx* c, double dt2) {
  Bd bd(c->b); Grid g(c->g, bd);
  p_grad_c(dt2, bd.npz, F3(c, FV3_DELPC), F3(c, FV3_PKC), F3(c, FV3_GZ), F3(c, FV3_UC), F3(c, FV3_VC), bd, g.rdxc, g.rdyc,
           c->f.hydrostatic != 0);
  return 0;
}

// dyn_core.F90:666-812 (use_old_omega=T, d_ext=0, do_f3d=F paths)
int fv3o_d_sw(fv3o_ctx* c, double dt) {
  Bd bd(c->b); Grid g(c->g, bd);
  const fv3_flags_t& f = c->f;
  const int npz = bd.npz;
  std::vector<int> nord_k_(npz + 1), nord_w_(npz + 1), nord_t_(npz + 1);
  std::vector<double> d2_divg_(npz + 1), damp_w_(npz + 1), damp_t_(npz + 1), d_con_k_(npz + 1);
  for (int k = 1; k <= npz; k++) {
    int nord_k = f.nord;
    c->nord_v[k - 1] = std::min(2, f.nord);
    double d2_divg = std::min(0.20, f.d2_bg);
    c->damp_vt[k - 1] = f.do_vort_damp ? f.vtdm4 : 0.;
    int nord_w = c->nord_v[k - 1], nord_t = c->nord_v[k - 1];
    double damp_w = c->damp_vt[k - 1], damp_t = c->damp_vt[k - 1], d_con_k = f.d_con;
    if (npz == 1 || f.n_sponge < 0) {
      d2_divg = f.d2_bg;
    } else {
      if (k == 1) {
        nord_k = 0;
        if (f.is_ideal_case) d2_divg = std::max(f.d2_bg, f.d2_bg_k1);
        else d2_divg = max3(0.01, f.d2_bg, f.d2_bg_k1);
        nord_w = 0; damp_w = d2_divg;
        if (f.do_vort_damp) { c->nord_v[k - 1] = 0; c->damp_vt[k - 1] = 0.5 * d2_divg; }
        d_con_k = 0.;
      } else if (k == 2 && f.d2_bg_k2 > 0.01) {
        nord_k = 0; d2_divg = std::max(f.d2_bg, f.d2_bg_k2);
        nord_w = 0; damp_w = d2_divg;
        if (f.do_vort_damp) { c->nord_v[k - 1] = 0; c->damp_vt[k - 1] = 0.5 * d2_divg; }
        d_con_k = 0.;
      } else if (k == 3 && f.d2_bg_k2 > 0.05) {
        nord_k = 0; d2_divg = std::max(f.d2_bg, 0.2 * f.d2_bg_k2);
        nord_w = 0; damp_w = d2_divg;
        d_con_k = 0.;
      }
    }
    nord_k_[k] = nord_k; nord_w_[k] = nord_w; nord_t_[k] = nord_t;
    d2_divg_[k] = d2_divg; damp_w_[k] = damp_w; damp_t_[k] = damp_t; d_con_k_[k] = d_con_k;
  }
  const int nic = bd.ie - bd.is + 1, njc = bd.je - bd.js + 1;
#pragma omp parallel for schedule(dynamic)
  for (int k = 1; k <= npz; k++) {
    DswArgs a;
    a.dt = dt; a.hord_tr = f.hord_tr; a.hord_mt = f.hord_mt; a.hord_vt = f.hord_vt; a.hord_tm = f.hord_tm; a.hord_dp = f.hord_dp;
    a.nord = nord_k_[k]; a.nord_v = c->nord_v[k - 1]; a.nord_w = nord_w_[k]; a.nord_t = nord_t_[k];
    a.dddmp = f.dddmp; a.d2_bg = d2_divg_[k]; a.d4_bg = f.d4_bg; a.damp_v = c->damp_vt[k - 1]; a.damp_w = damp_w_[k];
    a.damp_t = damp_t_[k]; a.d_con = d_con_k_[k]; a.kgb = f.ke_bg; a.hydrostatic = f.hydrostatic != 0;
    a.use_cond = f.use_cond != 0; a.do_f3d = false; a.prevent_diss_cooling = f.prevent_diss_cooling != 0;
    a.do_diss_est = f.do_diss_est != 0; a.lim_fac = f.lim_fac; a.sw_test_case = f.sw_test_case;
    L2 heat_s(bd.is, bd.ie, bd.js, bd.je), diss_e(bd.is, bd.ie, bd.js, bd.je), z_rat(bd.isd, bd.ied, bd.jsd, bd.jed, 1.0);
    // note the aliasing at dyn_core.F90:762: d_sw's delpc dummy is dyn_core's vt work array
    d_sw(F2(c, FV3_VT, k), F2(c, FV3_DELP, k), F2(c, FV3_PTC, k), F2(c, FV3_PT, k), F2(c, FV3_U, k), F2(c, FV3_V, k),
         F2(c, FV3_W, k), F2(c, FV3_UC, k), F2(c, FV3_VC, k), F2(c, FV3_UA, k), F2(c, FV3_VA, k), F2(c, FV3_DIVGD, k),
         F2(c, FV3_MFX, k), F2(c, FV3_MFY, k), F2(c, FV3_CX, k), F2(c, FV3_CY, k), F2(c, FV3_CRX, k), F2(c, FV3_CRY, k),
         F2(c, FV3_XFX, k), F2(c, FV3_YFX, k), F2(c, FV3_QCON, k), z_rat, heat_s, diss_e, a, g, bd);
    if (f.d_con > 1.0E-5) {
      V2 hs = F2(c, FV3_HEAT, k);
      for (int j = bd.js; j <= bd.je; j++) for (int i = bd.is; i <= bd.ie; i++) hs(i, j) = hs(i, j) + heat_s(i, j);
    }
    if (f.do_diss_est) {
      V2 ds = F2(c, FV3_DISS, k);
      for (int j = bd.js; j <= bd.je; j++) for (int i = bd.is; i <= bd.ie; i++) ds(i, j) = ds(i, j) + diss_e(i, j);
    }
  }
  (void)nic; (void)njc;
  return 0;
}

// dyn_core.F90:911-912 (nord_v, damp_vt as left by the d_sw prologue)
int fv3o_update_dz_d(fv3o_ctx* c, double dt) {
  Bd bd(c->b); Grid g(c->g, bd);
  std::vector<double> dp0 = dp_ref(c), zs = zs_of(c);
  update_dz_d(c->nord_v.data(), c->damp_vt.data(), c->f.hord_tm, bd.is, bd.ie, bd.js, bd.je, bd.npz, bd.ng, bd.npx, bd.npy,
              dp0.data(), V2(zs.data(), bd.isd, bd.jsd, bd.ied - bd.isd + 1), F3(c, FV3_ZH), F3(c, FV3_CRX), F3(c, FV3_CRY),
              F3(c, FV3_XFX), F3(c, FV3_YFX), F2(c, FV3_WS), 1.0 / dt, g, bd, c->f.lim_fac);
  return 0;
}
// dyn_core.F90:932-940
int fv3o_riem_solver3(fv3o_ctx* c, double dt, int last_call) {
  Bd bd(c->b);
  Consts k{c->f.rdgas, c->f.cp_air, c->f.grav, c->f.kappa, c->f.radius, c->f.omega, c->f.pi};
  std::vector<double> zs = zs_of(c);
  rayleigh_init(c, dt);
  if (c->rf_init) { k.rff = c->rff.data(); k.k_rf = c->k_rf; }
  riem_solver3(c->f.m_split, dt, bd.is, bd.ie, bd.js, bd.je, bd.npz, bd.ng, bd.isd, bd.ied, bd.jsd, bd.jed, c->f.kappa,
               F3(c, FV3_CAPPA), c->f.cp_air, c->f.ptop, V2(zs.data(), bd.isd, bd.jsd, bd.ied - bd.isd + 1), F3(c, FV3_QCON),
               F3(c, FV3_W), F3(c, FV3_DELZ), F3(c, FV3_PT), F3(c, FV3_DELP), F3(c, FV3_ZH), c->fld[FV3_PE].data(),
               F3(c, FV3_PKC), F3(c, FV3_PK3), F3(c, FV3_PK), c->fld[FV3_PELN].data(), F2(c, FV3_WS), c->f.p_fac, c->f.a_imp,
               c->f.use_logp != 0, c->f.use_cond != 0, c->f.moist_kappa != 0, last_call != 0, c->f.beta < -0.1, k);
  return 0;
}
int fv3o_pk3_halo(fv3o_ctx* c) {
  Bd bd(c->b);
  if (c->f.use_logp) pln_halo(bd.is, bd.ie, bd.js, bd.je, bd.isd, bd.ied, bd.jsd, bd.jed, bd.npz, c->f.ptop, F3(c, FV3_PK3), F3(c, FV3_DELP));   // dyn_core.F90:955-959
  else pk3_halo(bd.is, bd.ie, bd.js, bd.je, bd.isd, bd.ied, bd.jsd, bd.jed, bd.npz, c->f.ptop, c->f.kappa, F3(c, FV3_PK3), F3(c, FV3_DELP));
  return 0;
}
int fv3o_pe_halo(fv3o_ctx* c) {
  Bd bd(c->b);
  pe_halo(bd.is, bd.ie, bd.js, bd.je, bd.isd, bd.ied, bd.jsd, bd.jed, bd.npz, c->f.ptop, c->fld[FV3_PE].data(), F3(c, FV3_DELP));
  return 0;
}
// dyn_core.F90:982-989
int fv3o_gz_from_zh(fv3o_ctx* c) {
  Bd bd(c->b);
  V3 gz = F3(c, FV3_GZ), zh = F3(c, FV3_ZH);
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= bd.npz + 1; k++)
    for (int j = bd.js - 2; j <= bd.je + 2; j++)
      for (int i = bd.is - 2; i <= bd.ie + 2; i++) gz(i, j, k) = zh(i, j, k) * c->f.grav;
  return 0;
}
// dyn_core.F90:1028 split_p_grad / :1019 grad1_p_update (beta > 0): beta_d = 0 on the first substep of a call (:404-406)
// d_ext > 0 (external-mode divergence damping, hydrostatic branch): before d_sw the corner values of delp (a2b_ord2, :745-747;
// the reference parks them in ptc after d_sw, :791-797 -- d_sw uses ptc as scratch, so they wait in a private array here) ...
int fv3o_ext_mode_prepare(fv3o_ctx* c) {
  Bd bd(c->b); Grid g(c->g, bd);
  V3 delp = F3(c, FV3_DELP);
  const int nd = bd.ie - bd.is + 2, md = bd.je - bd.js + 2;
  c->ext_dpc.assign((size_t)nd * md * bd.npz, 0.);
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= bd.npz; k++) {
    L2 wk(bd.isd, bd.ied, bd.jsd, bd.jed);
    a2b_ord2(delp.k(k), wk, g, bd);
    for (int j = bd.js; j <= bd.je + 1; j++)
      for (int i = bd.is; i <= bd.ie + 1; i++) c->ext_dpc[(i - bd.is) + (size_t)(j - bd.js) * nd + (size_t)(k - 1) * nd * md] = wk(i, j);
  }
  return 0;
}
// ... and after d_sw the mass-weighted vertical mean of the divergence d_sw left in vt (:828-847)
int fv3o_ext_mode_divg2(fv3o_ctx* c) {
  Bd bd(c->b);
  V3 vt = F3(c, FV3_VT);
  const int nd = bd.ie - bd.is + 2, md = bd.je - bd.js + 2;
  if (c->ext_dpc.size() != (size_t)nd * md * bd.npz) return -1;
  auto ptc = [&](int i, int j, int k) { return c->ext_dpc[(i - bd.is) + (size_t)(j - bd.js) * nd + (size_t)(k - 1) * nd * md]; };
  c->divg2.assign((size_t)nd * md, 0.);
  const double d2_divg = c->f.d_ext * c->g.da_min_c;
  for (int j = bd.js; j <= bd.je + 1; j++)
    for (int i = bd.is; i <= bd.ie + 1; i++) {
      double wk = ptc(i, j, 1), d = wk * vt(i, j, 1);
      for (int k = 2; k <= bd.npz; k++) { wk = wk + ptc(i, j, k); d = d + ptc(i, j, k) * vt(i, j, k); }
      c->divg2[(i - bd.is) + (size_t)(j - bd.js) * nd] = d2_divg * d / wk;
    }
  return 0;
}
int fv3o_split_p_grad(fv3o_ctx* c, double dt, double beta_d) {
  Bd bd(c->b); Grid g(c->g, bd);
  const double* d2 = (c->f.d_ext > 0. && !c->divg2.empty()) ? c->divg2.data() : nullptr;
  if (c->f.hydrostatic)
    grad1_p_update(F3(c, FV3_U), F3(c, FV3_V), F3(c, FV3_PKC), F3(c, FV3_GZ), F3(c, FV3_DU), F3(c, FV3_DV), dt, g, bd, bd.npz, c->f.ptop, c->f.kappa, beta_d, d2);
  else
    split_p_grad(F3(c, FV3_U), F3(c, FV3_V), F3(c, FV3_PKC), F3(c, FV3_GZ), F3(c, FV3_DELP), F3(c, FV3_PK3), F3(c, FV3_DU), F3(c, FV3_DV), beta_d, dt,
                 g, bd, bd.npz, c->f.use_logp != 0, c->f.ptop, c->f.kappa);
  return 0;
}
// dyn_core.F90:1032
int fv3o_nh_p_grad(fv3o_ctx* c, double dt) {
  Bd bd(c->b); Grid g(c->g, bd);
  nh_p_grad(F3(c, FV3_U), F3(c, FV3_V), F3(c, FV3_PKC), F3(c, FV3_GZ), F3(c, FV3_DELP), F3(c, FV3_PK3), dt, bd.ng, g, bd,
            bd.npx, bd.npy, bd.npz, c->f.use_logp != 0, c->f.ptop, c->f.kappa);
  return 0;
}
// dyn_core.F90:2202 geopk: cg != 0 -> C-grid call (:478-480, delpc/ptc), else D-grid call (:905-907, delp/pt)
int fv3o_geopk(fv3o_ctx* c, int cg) {
  Bd bd(c->b);
  geopk(c->f.ptop, c->fld[FV3_PE].data(), c->fld[FV3_PELN].data(), F3(c, cg ? FV3_DELPC : FV3_DELP), F3(c, FV3_PKC), F3(c, FV3_GZ),
        F2(c, FV3_PHIS), F3(c, cg ? FV3_PTC : FV3_PT), F3(c, FV3_QCON), F3(c, FV3_PKZ), bd.npz, c->f.kappa, c->f.cp_air, cg != 0,
        c->f.use_cond != 0, bd);
  return 0;
}
// dyn_core.F90:1001-1010 (remap_step .and. hydrostatic): pk = pkc on the compute domain
int fv3o_pk_from_pkc(fv3o_ctx* c) {
  Bd bd(c->b);
  V3 pk = F3(c, FV3_PK), pkc = F3(c, FV3_PKC);
  for (int k = 1; k <= bd.npz + 1; k++)
    for (int j = bd.js; j <= bd.je; j++)
      for (int i = bd.is; i <= bd.ie; i++) pk(i, j, k) = pkc(i, j, k);
  return 0;
}
// dyn_core.F90:1909 one_grad_p (hydrostatic call :1019-1021, d_ext = 0)
int fv3o_one_grad_p(fv3o_ctx* c, double dt) {
  Bd bd(c->b); Grid g(c->g, bd);
  const double* d2 = (c->f.d_ext > 0. && !c->divg2.empty()) ? c->divg2.data() : nullptr;
  one_grad_p(F3(c, FV3_U), F3(c, FV3_V), F3(c, FV3_PKC), F3(c, FV3_GZ), F3(c, FV3_DELP), dt, g, bd, bd.npz, c->f.ptop, c->f.kappa,
             c->f.hydrostatic != 0, d2);
  return 0;
}
// dyn_core.F90:2356 del2_cubed on one field (FV3_HEAT: :1303 with cd = 0.20 da_min, nmax = min(3, nord+1); FV3_OMGA:
// fv_dynamics.F90:640 with cd = 0.18 da_min, nmax = nf_omega).  The caller has exchanged the halo of the field.
int fv3o_del2_cubed(fv3o_ctx* c, int field, double cd, int nmax) {
  if (field != FV3_HEAT && field != FV3_OMGA) return -1;
  Bd bd(c->b); Grid g(c->g, bd);
  del2_cubed(F3(c, field), cd, g, bd, bd.npz, nmax);
  return 0;
}
// fv_dynamics.F90:303-328, :377-398: specific humidity in FV3_WORK_Q (read only when zvir != 0 is meaningful; it is multiplied anyway)
// omega diagnostic of the last substep of the last dyn_core call (end_step): before the substep ...
int fv3o_omega_new(fv3o_ctx* c, int phase, double dt);
int fv3o_omega_begin(fv3o_ctx* c) {
  if (!c->f.use_old_omega) return 0;
  Bd bd(c->b);
  c->pem.assign(c->fld[FV3_PE].size(), 0.);
  pem_from_delp(c->pem.data(), F3(c, FV3_DELP), c->f.ptop, bd);
  return 0;
}
// ... and after it (use_old_omega = T, dyn_core.F90:1182-1195)
int fv3o_omega_end(fv3o_ctx* c, double dt) {
  if (!c->f.use_old_omega) return fv3o_omega_new(c, 2, dt);
  if (c->pem.size() != c->fld[FV3_PE].size() || !c->g.ec1 || !c->g.ec2 || !c->g.en1 || !c->g.en2) return -1;
  Bd bd(c->b); Grid g(c->g, bd);
  omega_old(F3(c, FV3_OMGA), c->fld[FV3_PE].data(), c->pem.data(), F3(c, FV3_UA), F3(c, FV3_VA), 1. / dt, g, bd);
  return 0;
}
// use_old_omega = F (dyn_core.F90:735-742, 774-781, 1196-1214): phase 0 before d_sw: omga = delp; phase 1 after d_sw: times the
// convergence of the area fluxes / dt; phase 2 at the end of the substep: running sum over k
int fv3o_omega_new(fv3o_ctx* c, int phase, double dt) {
  Bd bd(c->b); Grid g(c->g, bd);
  V3 omga = F3(c, FV3_OMGA), delp = F3(c, FV3_DELP), xfx = F3(c, FV3_XFX), yfx = F3(c, FV3_YFX);
  const double rdt = 1. / dt;
  if (phase == 0) {
    for (int k = 1; k <= bd.npz; k++) for (int j = bd.js; j <= bd.je; j++) for (int i = bd.is; i <= bd.ie; i++) omga(i, j, k) = delp(i, j, k);
  } else if (phase == 1) {
    for (int k = 1; k <= bd.npz; k++)
      for (int j = bd.js; j <= bd.je; j++)
        for (int i = bd.is; i <= bd.ie; i++)
          omga(i, j, k) = omga(i, j, k) * (xfx(i, j, k) - xfx(i + 1, j, k) + yfx(i, j, k) - yfx(i, j + 1, k)) * g.rarea(i, j) * rdt;
  } else {
    for (int j = bd.js; j <= bd.je; j++)
      for (int i = bd.is; i <= bd.ie; i++) {
        double om = omga(i, j, 1);
        for (int k = 2; k <= bd.npz; k++) { om = om + omga(i, j, k); omga(i, j, k) = om; }
      }
  }
  return 0;
}
// fv_operators.F90 map_scalar (mode 0) / map1_ppm (1) / map1_q2 (2) of FV3_WORK_Q on the compute domain: from the layers of FV3_PE to
// the hybrid levels ak + bk * pe(km+1) (remap.cpp)
int fv3o_remap_work_q(fv3o_ctx* c, int mode, int iv, int kord, double qmin) {
  Bd bd(c->b);
  return remap_work_q(F3(c, FV3_WORK_Q), F2(c, FV3_WS), c->fld[FV3_PE].data(), c->ak, c->bk, c->f, bd, mode, iv, kord, qmin);
}
// fv_mapz.F90:56-845 Lagrangian_to_Eulerian on one face (remap.cpp states the supported subset)
int fv3o_set_num_tracers(fv3o_ctx* c, int nq) {
  if (nq < 1 || nq > 64) return -1;
  if (c->store.empty()) c->store.resize(1);
  if (c->tracer_sel >= nq) return -1;
  c->store.resize(nq);
  for (int i = 0; i < nq; i++) if (i != c->tracer_sel && c->store[i].size() != c->fld[FV3_WORK_Q].size()) c->store[i].assign(c->fld[FV3_WORK_Q].size(), 0.);
  return 0;
}
int fv3o_num_tracers(const fv3o_ctx* c) { return c->store.empty() ? 1 : (int)c->store.size(); }
int fv3o_select_tracer(fv3o_ctx* c, int iq) {
  if (c->store.empty()) c->store.resize(1);
  if (iq < 0 || iq >= (int)c->store.size()) return -1;
  if (iq == c->tracer_sel) return 0;
  c->store[c->tracer_sel].swap(c->fld[FV3_WORK_Q]);   // park the selected tracer
  c->fld[FV3_WORK_Q].swap(c->store[iq]);
  c->tracer_sel = iq;
  return 0;
}
// use_tracer: the number of tracers to remap (the first use_tracer of the context's table)
int fv3o_lagrangian_to_eulerian_qv(fv3o_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum, double r_vir);
int fv3o_lagrangian_to_eulerian(fv3o_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr) {
  return fv3o_lagrangian_to_eulerian_qv(c, last_step, kord_mt, kord_wz, kord_tm, use_tracer, kord_tr, -1, 0.);
}
// sphum: index of the specific-humidity tracer (< use_tracer) or -1; r_vir = rvgas / rdgas - 1
int fv3o_lagrangian_to_eulerian_qv(fv3o_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum, double r_vir) {
  Bd bd(c->b);
  if (use_tracer < 0 || use_tracer > fv3o_num_tracers(c)) return -1;
  L2EFields F{F3(c, FV3_PT), F3(c, FV3_DELP), F3(c, FV3_DELZ), F3(c, FV3_W), F3(c, FV3_U), F3(c, FV3_V), F3(c, FV3_PK), F3(c, FV3_PKZ),
              F3(c, FV3_OMGA), {}, F2(c, FV3_WS), c->fld[FV3_PE].data(), c->fld[FV3_PELN].data()};
  const FieldDim& d = c->dim[FV3_WORK_Q];
  for (int iq = 0; iq < use_tracer; iq++) {
    std::vector<double>& buf = (iq == c->tracer_sel || c->store.empty()) ? c->fld[FV3_WORK_Q] : c->store[iq];
    F.qtr.push_back(V3(buf.data(), d.ilo, d.ilo + d.ni - 1, d.jlo, d.jlo + d.nj - 1));
  }
  return lagrangian_to_eulerian(F, c->ak, c->bk, c->f, bd, last_step, kord_mt, kord_wz, kord_tm, use_tracer, kord_tr, sphum, r_vir, c->tracer_fill);
}
// flagstruct%fill; fillz (fv_fill.F90:34) of FV3_WORK_Q with the thicknesses FV3_DELP on the compute domain
int fv3o_set_tracer_fill(fv3o_ctx* c, int on) { c->tracer_fill = on != 0; return 0; }
int fv3o_fillz(fv3o_ctx* c) { Bd bd(c->b); fillz(F3(c, FV3_WORK_Q), F3(c, FV3_DELP), bd); return 0; }
int fv3o_pt_to_theta(fv3o_ctx* c, double zvir) {
  if (c->f.moist_kappa) return -2;
  Bd bd(c->b);
  pt_to_theta(F3(c, FV3_PT), F3(c, FV3_DELP), F3(c, FV3_DELZ), F3(c, FV3_WORK_Q), F3(c, FV3_QCON), F3(c, FV3_DP1), F3(c, FV3_PKZ), zvir, c->f, bd);
  return 0;
}
// dyn_core.F90:1300-1356 without the del2_cubed call
int fv3o_dcon_heating(fv3o_ctx* c, double bdt) {
  Bd bd(c->b);
  const int n_con = n_con_levels(c->f, bd.npz);
  if (n_con == 0 || !(c->f.d_con > 1.e-5)) return 0;
  dcon_heating(F3(c, FV3_PT), F3(c, FV3_HEAT), F3(c, FV3_DELP), F3(c, FV3_DELZ), F3(c, FV3_PKZ), n_con, bdt, c->f, bd);
  return 0;
}
// dyn_core.F90:370-385 (it==1): gz from zs and delz on the compute domain
int fv3o_gz_init(fv3o_ctx* c) {
  Bd bd(c->b);
  std::vector<double> zs = zs_of(c);
  V2 zsv(zs.data(), bd.isd, bd.jsd, bd.ied - bd.isd + 1);
  V3 gz = F3(c, FV3_GZ), delz = F3(c, FV3_DELZ);
  for (int j = bd.js; j <= bd.je; j++) {
    for (int i = bd.is; i <= bd.ie; i++) gz(i, j, bd.npz + 1) = zsv(i, j);
    for (int k = bd.npz; k >= 1; k--) for (int i = bd.is; i <= bd.ie; i++) gz(i, j, k) = gz(i, j, k + 1) - delz(i, j, k);
  }
  return 0;
}
// dyn_core.F90:491-521: zh = gz (it==1) or gz = zh
int fv3o_copy_field(fv3o_ctx* c, int dst, int src) {
  if (c->fld[dst].size() != c->fld[src].size()) return -1;
  double* d = c->fld[dst].data(); const double* s = c->fld[src].data();
  const ptrdiff_t n = (ptrdiff_t)c->fld[dst].size();
#pragma omp parallel for schedule(static)
  for (ptrdiff_t i = 0; i < n; i++) d[i] = s[i];
  return 0;
}
int fv3o_zero_field(fv3o_ctx* c, int f) {
  double* d = c->fld[f].data(); const ptrdiff_t n = (ptrdiff_t)c->fld[f].size();
#pragma omp parallel for schedule(static)
  for (ptrdiff_t i = 0; i < n; i++) d[i] = 0.0;
  return 0;
}

// ---- 6-tile halo exchange inside the library (the FMS group updates of dyn_core.F90:350-1169; FMS itself is not in the
// reference checkout).  The index / sign tables are built by gfdl_atmos_cubed_sphere_b200/cubed_sphere.py (the tables the
// NumPy exchange of the harness uses, tests/test_grid_and_halo.py validates them) and registered once per table id.
struct HaloTab { std::vector<long long> dst, src; std::vector<int> src_tile, src_comp; std::vector<double> sign; };
static std::map<int, std::vector<HaloTab>>& halo_registry() { static std::map<int, std::vector<HaloTab>> r; return r; }

int fv3o_halo_set_table(int table_id, int tile, int comp, long long n, const long long* dst, const int* src_tile, const int* src_comp,
                        const long long* src, const double* sign) {
  if (tile < 1 || tile > 6 || comp < 0 || comp > 1) return -1;
  auto& v = halo_registry()[table_id];
  if (v.size() != 12) v.assign(12, HaloTab());
  HaloTab& t = v[(tile - 1) * 2 + comp];
  t.dst.assign(dst, dst + n); t.src.assign(src, src + n); t.src_tile.assign(src_tile, src_tile + n);
  t.src_comp.assign(src_comp, src_comp + n); t.sign.assign(sign, sign + n);
  return 0;
}
// ctxs[0..5] = tiles 1..6.  field_y < 0: scalar exchange of field_x; else the (x, y) pair with the table's component swaps / signs.
// Sources are compute-domain points, destinations halo (or, for the boundary-only tables, north / east edge) points: never the
// same point, so the gather runs in place.
int fv3o_halo_exchange(fv3o_ctx** ctxs, int field_x, int field_y, int table_id) {
  auto it = halo_registry().find(table_id);
  if (it == halo_registry().end()) return -1;
  const std::vector<HaloTab>& tabs = it->second;
  const int ncomp = field_y < 0 ? 1 : 2;
  const int fields[2] = {field_x, field_y};
  const int nk = ctxs[0]->dim[field_x].nk;
#pragma omp parallel for collapse(2) schedule(static)
  for (int job = 0; job < 6 * ncomp; job++) {
    for (int k = 0; k < nk; k++) {
      const int t = job / ncomp, ci = job % ncomp;
      const HaloTab& tb = tabs[t * 2 + ci];
      const FieldDim& dd = ctxs[t]->dim[fields[ci]];
      double* dp = ctxs[t]->fld[fields[ci]].data() + (size_t)k * dd.ni * dd.nj;
      const size_t n = tb.dst.size();
      for (size_t e = 0; e < n; e++) {
        const int sf = fields[tb.src_comp[e]];
        const fv3o_ctx* sc = ctxs[tb.src_tile[e] - 1];
        const FieldDim& sd = sc->dim[sf];
        dp[tb.dst[e]] = sc->fld[sf][(size_t)k * sd.ni * sd.nj + tb.src[e]] * tb.sign[e];
      }
    }
  }
  return 0;
}

// 1-D periodic xppm / yppm (interior formulas only, grid_type = 4): used to pin the oracle against the
// NumPy restatement in the reference's docs/examples/tp_core.ipynb (tests/golden/ppm_notebook.npz)
int fv3o_ppm_periodic(int n, const double* q, const double* cn, int iord, int ydir, double* flux) {
  const int ng = 3, isd = 1 - ng, ied = n + ng;
  if (!ydir) {
    std::vector<double> qh((size_t)(ied - isd + 1)), ch(n + 1), fh(n + 1), dxa((size_t)(ied - isd + 1), 1.0);
    for (int i = isd; i <= ied; i++) qh[i - isd] = q[((i - 1) % n + n) % n];
    for (int i = 0; i <= n; i++) ch[i] = cn[i];
    xppm(V2(fh.data(), 1, 1, n + 1), V2(qh.data(), isd, 1, ied - isd + 1), V2(ch.data(), 1, 1, n + 1), iord, 1, n, isd, ied, 1, 1, 1, 1,
         n + 1, n + 1, V2(dxa.data(), isd, 1, ied - isd + 1), false, 4, 1.0);
    for (int i = 0; i <= n; i++) flux[i] = fh[i];
  } else {
    // one column (ifirst = ilast = 1), j is the sweep index
    std::vector<double> qh((size_t)(ied - isd + 1)), ch(n + 1), fh(n + 1), dya((size_t)(ied - isd + 1), 1.0);
    for (int j = isd; j <= ied; j++) qh[j - isd] = q[((j - 1) % n + n) % n];
    for (int j = 0; j <= n; j++) ch[j] = cn[j];
    yppm(V2(fh.data(), 1, 1, 1), V2(qh.data(), 1, isd, 1), V2(ch.data(), 1, 1, 1), iord, 1, 1, 1, 1, 1, n, isd, ied, n + 1, n + 1,
         V2(dya.data(), 1, isd, 1), false, 4, 1.0);
    for (int j = 0; j <= n; j++) flux[j] = fh[j];
  }
  return 0;
}

// One full cube-face line of xppm / yppm (is = 1, ie = n, cube-edge formulas active at both ends): q and dxa carry the 3-cell
// halo (index -2..n+3), cn and flux the faces 1..n+1.  Used by tests/test_host_device_math.py to pin the cube-edge operator.
int fv3o_ppm_cube_line(int n, const double* q, const double* cn, const double* dxa, int iord, int ydir, double* flux) {
  const int isd = -2, ied = n + 3, npx = n + 1;
  std::vector<double> qh(q, q + n + 6), dh(dxa, dxa + n + 6), ch(cn, cn + n + 1), fh(n + 1);
  if (!ydir) {
    xppm(V2(fh.data(), 1, 1, n + 1), V2(qh.data(), isd, 1, n + 6), V2(ch.data(), 1, 1, n + 1), iord, 1, n, isd, ied, 1, 1, 1, 1,
         npx, npx, V2(dh.data(), isd, 1, n + 6), false, 0, 1.0);
  } else {   // one column (ifirst = ilast = 1), j is the sweep index
    yppm(V2(fh.data(), 1, 1, 1), V2(qh.data(), 1, isd, 1), V2(ch.data(), 1, 1, 1), iord, 1, 1, 1, 1, 1, n, isd, ied, npx, npx,
         V2(dh.data(), 1, isd, 1), false, 0, 1.0);
  }
  for (int i = 0; i <= n; i++) flux[i] = fh[i];
  return 0;
}
// One full cube-face line of xtp_u (row j of a face with npy = npx): u, dx, rdx carry the halo (-2..n+3), c and flux the faces
// 1..n+1.  j = 1 or npy is a face-edge line (bl = br = 0 at the corner cells, sw_core.F90:2206-2210).
int fv3o_xtp_u_line(int n, int j, const double* u, const double* cn, const double* dx, const double* rdx, int iord, double* flux) {
  const int isd = -2, ied = n + 3, npx = n + 1;
  std::vector<double> uh(u, u + n + 6), dh(dx, dx + n + 6), rh(rdx, rdx + n + 6), ch(cn, cn + n + 1), fh(n + 1);
  // js = j, je = j - 1: the routine sweeps rows js..je+1 = the single row j
  xtp_u(1, n, j, j - 1, isd, ied, j, j, V2(ch.data(), 1, j, n + 1), V2(uh.data(), isd, j, n + 6), V2(nullptr, 0, 0, 0),
        V2(fh.data(), 1, j, n + 1), iord, V2(dh.data(), isd, j, n + 6), V2(rh.data(), isd, j, n + 6), npx, npx, 0, false, 1.0);
  for (int i = 0; i <= n; i++) flux[i] = fh[i];
  return 0;
}

// One full cube-face line of ytp_v (column i of a face): v, dy, rdy carry the halo (-2..n+3), c and flux the faces 1..n+1.
// i = 1 or npx is a face-edge line.
int fv3o_ytp_v_line(int n, int i, const double* v, const double* cn, const double* dy, const double* rdy, int jord, double* flux) {
  const int jsd = -2, jed = n + 3, npy = n + 1;
  std::vector<double> vh(v, v + n + 6), dh(dy, dy + n + 6), rh(rdy, rdy + n + 6), ch(cn, cn + n + 1), fh(n + 1);
  // is = i, ie = i - 1: the routine sweeps columns is..ie+1 = the single column i
  ytp_v(i, i - 1, 1, n, i, i, jsd, jed, V2(ch.data(), i, 1, 1), V2(nullptr, 0, 0, 0), V2(vh.data(), i, jsd, 1), V2(fh.data(), i, 1, 1), jord,
        V2(dh.data(), i, jsd, 1), V2(rh.data(), i, jsd, 1), npy, npy, 0, false, 1.0);
  for (int j = 0; j <= n; j++) flux[j] = fh[j];
  return 0;
}

// stand-alone operators for unit parity
int fv3o_a2b_ord4(fv3o_ctx* c, int field, int k, double* qout /*(isd:ied,jsd:jed)*/, int replace) {
  Bd bd(c->b); Grid g(c->g, bd);
  a2b_ord4(F2(c, field, k), V2(qout, bd.isd, bd.jsd, bd.ied - bd.isd + 1), g, bd, replace != 0);
  return 0;
}

}  // extern "C"
