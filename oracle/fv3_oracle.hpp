// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FV3 acoustic-dynamics hot path.
//
// A line-faithful C++ restatement of the reference Fortran (same loop bounds,
// same floating-point operation order; build with -ffp-contract=off for the
// parity build).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product
// (gfdl_atmos_cubed_sphere_b200/csrc) never links or calls it.
//
// PARITY UNPINNED: the reference ships no golden vectors, fixtures or tests
// for this path and cannot be compiled here (no Fortran compiler, FMS not
// vendored) -- see DESIGN.md.  What pins this oracle instead: operator
// invariants (tests/test_oracle_invariants.py) and the NumPy interior-PPM
// restatement in docs/examples/tp_core.ipynb (tests/golden/).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../include/fv3_dyncore.h"

namespace fv3o {

// Fortran-style 2-D view: element (i,j) with explicit lower bounds, i fastest.
struct V2 {
  double* p; int i0, j0, ni;
  V2() : p(nullptr), i0(0), j0(0), ni(0) {}
  V2(double* p_, int i0_, int j0_, int ni_) : p(p_), i0(i0_), j0(j0_), ni(ni_) {}
  inline double& operator()(int i, int j) const {
    return p[(i - i0) + (ptrdiff_t)(j - j0) * ni];
  }
};
// Owning local array (Fortran automatic array).
#ifndef FV3O_ARENA
struct L2 : V2 {
  std::vector<double> buf;
  L2(int ilo, int ihi, int jlo, int jhi, double fill = 0.0)
      : buf((size_t)(ihi - ilo + 1) * (size_t)(jhi - jlo + 1), fill) {
    p = buf.data(); i0 = ilo; j0 = jlo; ni = ihi - ilo + 1;
  }
  L2(const L2&) = delete;
  void fill(double v) { std::fill(buf.begin(), buf.end(), v); }
};
#else
// Timing build (-DFV3O_ARENA): like a Fortran automatic array the storage comes from a per-thread stack (LIFO bump
// allocator, no malloc / munmap / page faults per level) and is NOT initialised unless a fill value is given.  With
// FV3O_POISON=1 in the environment every such array is filled with signalling garbage (NaN) first, which is how
// tests/test_oracle_invariants.py proves that no routine reads an element it did not write.
struct Arena {
  char* base = nullptr; size_t cap = 0, top = 0;
  static Arena& get() { thread_local Arena a; return a; }
  static bool poison() { static const bool p = [] { const char* e = std::getenv("FV3O_POISON"); return e && e[0] == '1'; }(); return p; }
  double* push(size_t n) {
    const size_t bytes = (n * sizeof(double) + 63) & ~(size_t)63;
    if (!base) { cap = (size_t)1 << 30; base = (char*)std::malloc(cap); }   // 1 GiB of address space per thread, touched lazily
    if (top + bytes > cap) return nullptr;
    double* r = (double*)(base + top); top += bytes;
    return r;
  }
  void pop(size_t n) { top -= (n * sizeof(double) + 63) & ~(size_t)63; }
};
struct L2 : V2 {
  size_t n_; bool heap_;
  void init(int ilo, int ihi, int jlo, int jhi) {
    n_ = (size_t)(ihi - ilo + 1) * (size_t)(jhi - jlo + 1);
    p = Arena::get().push(n_); heap_ = (p == nullptr);
    if (heap_) p = (double*)std::malloc(n_ * sizeof(double));
    i0 = ilo; j0 = jlo; ni = ihi - ilo + 1;
  }
  L2(int ilo, int ihi, int jlo, int jhi) {
    init(ilo, ihi, jlo, jhi);
    if (Arena::poison()) { const double nan = std::nan(""); for (size_t i = 0; i < n_; i++) p[i] = nan; }
  }
  L2(int ilo, int ihi, int jlo, int jhi, double fillv) { init(ilo, ihi, jlo, jhi); fill(fillv); }
  ~L2() { if (heap_) std::free(p); else Arena::get().pop(n_); }
  L2(const L2&) = delete;
  void fill(double v) { for (size_t i = 0; i < n_; i++) p[i] = v; }
};
#endif
// raw automatic 3-D work array (update_dz_d's crx_adv ... yfx_adv): zero-filled vector in the parity build, stack storage in the
// timing build
#ifndef FV3O_ARENA
struct LRaw { std::vector<double> v; explicit LRaw(size_t n) : v(n) {} double* data() { return v.data(); } };
#else
struct LRaw {
  double* p; size_t n_; bool heap_;
  explicit LRaw(size_t n) : n_(n) {
    p = Arena::get().push(n); heap_ = (p == nullptr);
    if (heap_) p = (double*)std::malloc(n * sizeof(double));
    if (Arena::poison()) { const double nan = std::nan(""); for (size_t i = 0; i < n; i++) p[i] = nan; }
  }
  ~LRaw() { if (heap_) std::free(p); else Arena::get().pop(n_); }
  LRaw(const LRaw&) = delete;
  double* data() { return p; }
};
#endif
struct L1 {
  std::vector<double> buf; int i0;
  L1(int ilo, int ihi, double fill = 0.0) : buf((size_t)(ihi - ilo + 1), fill), i0(ilo) {}
  inline double& operator()(int i) { return buf[i - i0]; }
};
struct LB1 {  // logical
  std::vector<char> buf; int i0;
  LB1(int ilo, int ihi) : buf((size_t)(ihi - ilo + 1), 0), i0(ilo) {}
  inline char& operator()(int i) { return buf[i - i0]; }
};
struct LB2 {
  std::vector<char> buf; int i0, j0, ni;
  LB2(int ilo, int ihi, int jlo, int jhi)
      : buf((size_t)(ihi - ilo + 1) * (size_t)(jhi - jlo + 1), 0), i0(ilo), j0(jlo), ni(ihi - ilo + 1) {}
  inline char& operator()(int i, int j) { return buf[(i - i0) + (size_t)(j - j0) * ni]; }
};
// 3-D view (i,j,k), k is 1-based.
struct V3 {
  double* p; int i0, j0, ni, nj;
  V3() : p(nullptr), i0(0), j0(0), ni(0), nj(0) {}
  V3(double* p_, int i0_, int i1, int j0_, int j1)
      : p(p_), i0(i0_), j0(j0_), ni(i1 - i0_ + 1), nj(j1 - j0_ + 1) {}
  inline double& operator()(int i, int j, int k) const {
    return p[(i - i0) + (ptrdiff_t)(j - j0) * ni + (ptrdiff_t)(k - 1) * ni * nj];
  }
  inline V2 k(int kk) const { return V2(p ? p + (ptrdiff_t)(kk - 1) * ni * nj : nullptr, i0, j0, ni); }
  inline bool ok() const { return p != nullptr; }
};

inline double fsign(double a, double b) { return (b >= 0.0 && !std::signbit(b)) ? std::fabs(a) : -std::fabs(a); }
inline double min3(double a, double b, double c) { return std::min(std::min(a, b), c); }
inline double max3(double a, double b, double c) { return std::max(std::max(a, b), c); }
inline double min4(double a, double b, double c, double d) { return std::min(std::min(a, b), std::min(c, d)); }
inline double max4(double a, double b, double c, double d) { return std::max(std::max(a, b), std::max(c, d)); }

struct Bd {
  int npx, npy, npz, ng, is, ie, js, je, isd, ied, jsd, jed, grid_type;
  bool bounded_domain, sw_corner, se_corner, ne_corner, nw_corner, stretched_grid;
  explicit Bd(const fv3_bounds_t& b)
      : npx(b.npx), npy(b.npy), npz(b.npz), ng(b.ng), is(b.is), ie(b.ie), js(b.js), je(b.je),
        isd(b.isd), ied(b.ied), jsd(b.jsd), jed(b.jed), grid_type(b.grid_type),
        bounded_domain(b.bounded_domain != 0), sw_corner(b.sw_corner != 0), se_corner(b.se_corner != 0),
        ne_corner(b.ne_corner != 0), nw_corner(b.nw_corner != 0), stretched_grid(b.stretched_grid != 0) {}
};

// gridstruct view (fv_arrays.F90:75-205)
struct Grid {
  V2 area, rarea, dxa, dya, rdxa, rdya, cosa_s, rsin2, f0;
  V2 dy, rdy, dxc, rdxc, cosa_u, sina_u, rsin_u, divg_v, del6_v;
  V2 dx, rdx, dyc, rdyc, cosa_v, sina_v, rsin_v, divg_u, del6_u;
  V2 area_c, rarea_c, fC, cosa, sina, rsina;
  const double *sin_sg_p, *cos_sg_p;
  const double *edge_w, *edge_e, *edge_s, *edge_n;  // 1-based: edge_w[j-1]
  const double *grid_p, *agrid_p;
  const double *ec1_p, *ec2_p, *en1_p, *en2_p;   // (3, ...) component fastest
  double da_min, da_min_c;
  int isd, jsd, nia, nja;
  Grid(const fv3_grid_t& g, const Bd& b) {
    isd = b.isd; jsd = b.jsd; nia = b.ied - b.isd + 1; nja = b.jed - b.jsd + 1;
    auto A = [&](const double* p) { return V2(const_cast<double*>(p), b.isd, b.jsd, nia); };
    auto B = [&](const double* p) { return V2(const_cast<double*>(p), b.isd, b.jsd, nia + 1); };
    area = A(g.area); rarea = A(g.rarea); dxa = A(g.dxa); dya = A(g.dya); rdxa = A(g.rdxa);
    rdya = A(g.rdya); cosa_s = A(g.cosa_s); rsin2 = A(g.rsin2); f0 = A(g.f0);
    dy = B(g.dy); rdy = B(g.rdy); dxc = B(g.dxc); rdxc = B(g.rdxc); cosa_u = B(g.cosa_u);
    sina_u = B(g.sina_u); rsin_u = B(g.rsin_u); divg_v = B(g.divg_v); del6_v = B(g.del6_v);
    dx = A(g.dx); rdx = A(g.rdx); dyc = A(g.dyc); rdyc = A(g.rdyc); cosa_v = A(g.cosa_v);
    sina_v = A(g.sina_v); rsin_v = A(g.rsin_v); divg_u = A(g.divg_u); del6_u = A(g.del6_u);
    area_c = B(g.area_c); rarea_c = B(g.rarea_c); fC = B(g.fC); cosa = B(g.cosa); sina = B(g.sina);
    rsina = V2(const_cast<double*>(g.rsina), b.is, b.js, b.ie + 1 - b.is + 1);
    sin_sg_p = g.sin_sg; cos_sg_p = g.cos_sg;
    edge_w = g.edge_w; edge_e = g.edge_e; edge_s = g.edge_s; edge_n = g.edge_n;
    grid_p = g.grid; agrid_p = g.agrid; da_min = g.da_min; da_min_c = g.da_min_c;
    ec1_p = g.ec1; ec2_p = g.ec2; en1_p = g.en1; en2_p = g.en2;
  }
  inline double sin_sg(int i, int j, int n) const {
    return sin_sg_p[(i - isd) + (ptrdiff_t)(j - jsd) * nia + (ptrdiff_t)(n - 1) * nia * nja];
  }
  inline double cos_sg(int i, int j, int n) const {
    return cos_sg_p[(i - isd) + (ptrdiff_t)(j - jsd) * nia + (ptrdiff_t)(n - 1) * nia * nja];
  }
  inline double grid(int i, int j, int n) const {  // (isd:ied+1, jsd:jed+1, 2)
    return grid_p[(i - isd) + (ptrdiff_t)(j - jsd) * (nia + 1) + (ptrdiff_t)(n - 1) * (nia + 1) * (nja + 1)];
  }
  inline double agrid(int i, int j, int n) const {
    return agrid_p[(i - isd) + (ptrdiff_t)(j - jsd) * nia + (ptrdiff_t)(n - 1) * nia * nja];
  }
};

// ---- tp_core.F90 ----
void copy_corners(V2 q, int npx, int npy, int dir, const Bd& bd);
void xppm(V2 flux, V2 q, V2 c, int iord, int is, int ie, int isd, int ied, int jfirst, int jlast,
          int jsd, int jed, int npx, int npy, V2 dxa, bool bounded_domain, int grid_type, double lim_fac);
void yppm(V2 flux, V2 q, V2 c, int jord, int ifirst, int ilast, int isd, int ied, int js, int je,
          int jsd, int jed, int npx, int npy, V2 dya, bool bounded_domain, int grid_type, double lim_fac);
void pert_ppm(int im, const double* a0, double* al, double* ar, int iv);
void deln_flux(int nord, int is, int ie, int js, int je, int npx, int npy, double damp, V2 q, V2 fx, V2 fy,
               const Grid& g, const Bd& bd, const V2* mass);
// optional: mfx,mfy (both or none), mass, nord/damp_c (use_damp)
void fv_tp_2d(V2 q, V2 crx, V2 cry, int npx, int npy, int hord, V2 fx, V2 fy, V2 xfx, V2 yfx,
              const Grid& g, const Bd& bd, V2 ra_x, V2 ra_y, double lim_fac, const V2* mfx, const V2* mfy,
              const V2* mass, bool use_damp, int nord, double damp_c);

// ---- a2b_edge.F90 ----
double great_circle_dist(const double q1[2], const double q2[2], double radius);
void a2b_ord4(V2 qin, V2 qout, const Grid& g, const Bd& bd, bool replace);
void a2b_ord2(V2 qin, V2 qout, const Grid& g, const Bd& bd);
void a2b_ord2(V2 qin, V2 qout, const Grid& g, const Bd& bd, bool replace);

// ---- fv_mp_mod.F90 fill_corners ----
void fill_corners_bgrid(V2 q, int npx, int npy, int ng, int fill_dir /*1 X,2 Y*/);
void fill_corners_dgrid_vec(V2 x, V2 y, int npx, int npy, int ng, double mysign);

// ---- sw_core.F90 ----
struct SwFlags {
  int npx, npy, grid_type; bool hydrostatic, do_f3d, prevent_diss_cooling, do_diss_est, inline_q;
  double lim_fac;
};
void fill_4corners(V2 q, int dir, const Bd& bd);
void fill2_4corners(V2 q1, V2 q2, int dir, const Bd& bd);
void d2a2c_vect(V2 u, V2 v, V2 ua, V2 va, V2 uc, V2 vc, V2 ut, V2 vt, bool dord4, const Grid& g, const Bd& bd);
void divergence_corner(V2 u, V2 v, V2 ua, V2 va, V2 divg_d, const Grid& g, const Bd& bd);
void del6_vt_flux(int nord, int npx, int npy, double damp, V2 q, V2 d2, V2 fx2, V2 fy2, const Grid& g, const Bd& bd);
void xtp_u(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, V2 c, V2 u, V2 v, V2 flux,
           int iord, V2 dx, V2 rdx, int npx, int npy, int grid_type, bool bounded_domain, double lim_fac);
void ytp_v(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, V2 c, V2 u, V2 v, V2 flux,
           int jord, V2 dy, V2 rdy, int npx, int npy, int grid_type, bool bounded_domain, double lim_fac);
void c_sw(V2 delpc, V2 delp, V2 ptc, V2 pt, V2 u, V2 v, V2 w, V2 uc, V2 vc, V2 ua, V2 va, V2 wc, V2 ut,
          V2 vt, V2 divg_d, int nord, double dt2, bool hydrostatic, bool dord4, const Bd& bd, const Grid& g);
struct DswArgs {
  double dt; int hord_tr, hord_mt, hord_vt, hord_tm, hord_dp; int nord, nord_v, nord_w, nord_t;
  double dddmp, d2_bg, d4_bg, damp_v, damp_w, damp_t, d_con, kgb; bool hydrostatic, use_cond;
  bool do_f3d, prevent_diss_cooling, do_diss_est; double lim_fac;
  int sw_test_case = 0;   // 1: SW_DYNAMICS build, test_case = 1 (pure advection by the prescribed uc, vc)
};
void d_sw(V2 delpc, V2 delp, V2 ptc, V2 pt, V2 u, V2 v, V2 w, V2 uc, V2 vc, V2 ua, V2 va, V2 divg_d,
          V2 xflux, V2 yflux, V2 cx, V2 cy, V2 crx_adv, V2 cry_adv, V2 xfx_adv, V2 yfx_adv, V2 q_con,
          V2 z_rat, V2 heat_source, V2 diss_est, const DswArgs& a, const Grid& g, const Bd& bd);

// ---- nh_utils.F90 / nh_core.F90 / dyn_core.F90 ----
struct Consts {
  double rdgas, cp_air, grav, kappa, radius, omega, pi;
  // Rayleigh damping of w (fast_tau_w_sec > 0): nh_utils.F90 keeps rff(1:k_rf) as SAVEd module state, set up by the FIRST solver call
  // with that call's dt (:356-368) and reused by every later one; the caller owns it here
  const double* rff = nullptr; int k_rf = 0;
};
void update_dz_c(int is, int ie, int js, int je, int km, int ng, double dt, const double* dp0, V2 zs, V2 area,
                 V3 ut, V3 vt, V3 gz, V2 ws, const Bd& bd);
void update_dz_d(int* ndif, double* damp, int hord, int is, int ie, int js, int je, int km, int ng, int npx,
                 int npy, const double* dp0, V2 zs, V3 zh, V3 crx, V3 cry, V3 xfx, V3 yfx, V2 ws, double rdt,
                 const Grid& g, const Bd& bd, double lim_fac);
void riem_solver_c(int ms, double dt, int is, int ie, int js, int je, int km, int ng, double akap, V3 cappa,
                   double cp, double ptop, V2 hs, V3 w3, V3 pt, V3 q_con, V3 delp, V3 gz, V3 pef, V2 ws,
                   double p_fac, double a_imp, bool use_cond, bool moist_kappa, const Consts& c);
void riem_solver3(int ms, double dt, int is, int ie, int js, int je, int km, int ng, int isd, int ied, int jsd,
                  int jed, double akap, V3 cappa, double cp, double ptop, V2 zs, V3 q_con, V3 w, V3 delz, V3 pt,
                  V3 delp, V3 zh, double* pe /*(is-1:ie+1,km+1,js-1:je+1)*/, V3 ppe, V3 pk3, V3 pk,
                  double* peln /*(is:ie,km+1,js:je)*/, V2 ws, double p_fac, double a_imp, bool use_logp,
                  bool use_cond, bool moist_kappa, bool last_call, bool fp_out, const Consts& c);
void pt_to_theta(V3 pt, V3 delp, V3 delz, V3 qv, V3 q_con, V3 dp1, V3 pkz, double zvir, const fv3_flags_t& f, const Bd& bd);
void del2_cubed(V3 q, double cd, const Grid& g, const Bd& bd, int km, int nmax);
int n_con_levels(const fv3_flags_t& f, int npz);
void dcon_heating(V3 pt, V3 heat_source, V3 delp, V3 delz, V3 pkz, int n_con, double bdt, const fv3_flags_t& f, const Bd& bd);
void p_grad_c(double dt2, int npz, V3 delpc, V3 pkc, V3 gz, V3 uc, V3 vc, const Bd& bd, V2 rdxc, V2 rdyc,
              bool hydrostatic);
void nh_p_grad(V3 u, V3 v, V3 pp, V3 gz, V3 delp, V3 pk, double dt, int ng, const Grid& g, const Bd& bd, int npx,
               int npy, int npz, bool use_logp, double ptop, double akap);
void geopk(double ptop, double* pe, double* peln, V3 delp, V3 pk, V3 gz, V2 hs, V3 pt, V3 q_con, V3 pkz, int km, double akap,
           double cp_air, bool CG, bool use_cond, const Bd& bd);
void split_p_grad(V3 u, V3 v, V3 pp, V3 gz, V3 delp, V3 pk, V3 du, V3 dv, double beta, double dt, const Grid& g, const Bd& bd, int npz,
                  bool use_logp, double ptop, double akap);
void grad1_p_update(V3 u, V3 v, V3 pk, V3 gz, V3 du, V3 dv, double dt, const Grid& g, const Bd& bd, int npz, double ptop, double akap, double beta,
                    const double* divg2 = nullptr);
void one_grad_p(V3 u, V3 v, V3 pk, V3 gz, V3 delp, double dt, const Grid& g, const Bd& bd, int npz, double ptop, double akap,
                bool hydrostatic, const double* divg2 = nullptr);
void pln_halo(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, int npz, double ptop, V3 pk3, V3 delp);
void pk3_halo(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, int npz, double ptop,
              double akap, V3 pk3, V3 delp);
void pe_halo(int is, int ie, int js, int je, int isd, int ied, int jsd, int jed, int npz, double ptop,
             double* pe, V3 delp);

// omega diagnostic of the last substep (dyn_core.F90:409-422, 1182-1195, adv_pe :1529-1630), nh.cpp
void pem_from_delp(double* pem, V3 delp, double ptop, const Bd& bd);
void omega_old(V3 omga, const double* pe, const double* pem, V3 ua, V3 va, double rdt, const Grid& g, const Bd& bd);

// remap.cpp (fv_mapz.F90 Lagrangian_to_Eulerian, fv_operators.F90 map_scalar / map1_ppm / map1_q2 and their profiles)
struct L2EFields { V3 pt, delp, delz, w, u, v, pk, pkz, omga; std::vector<V3> qtr; V2 ws; double *pe, *peln; };
int remap_work_q(V3 q, V2 ws, double* pe, const std::vector<double>& ak, const std::vector<double>& bk, const fv3_flags_t& f, const Bd& bd,
                 int mode, int iv, int kord, double qmin);
int lagrangian_to_eulerian(const L2EFields& F, const std::vector<double>& ak, const std::vector<double>& bk, const fv3_flags_t& f, const Bd& bd,
                           int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum = -1, double r_vir = 0., int fill = 0);
void fillz(V3 q, V3 dp, const Bd& bd);

}  // namespace fv3o
