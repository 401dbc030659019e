// Vertical remapping on device-resident state (the step between the k_split iterations of fv_dynamics).
//
// Reference semantics: model/fv_mapz.F90:56-845 (Lagrangian_to_Eulerian) and the column operators of model/fv_operators.F90:
// map_scalar (:40-134), map1_ppm (:137-229), map1_q2 (:352-443), scalar_profile (:546-916), cs_profile (:919-1300),
// cs_limiters (:1303-1378), ppm_profile / ppm_limiters (:1382-1723).  Built: remap_te = F, moist_kappa = F, consv = 0 (no energy
// fixer), dry air or water vapour without condensates, abs(kord) in 8..15 (cs / scalar profiles) and 1..7 (ppm_profile: separate
// instantiations of the kernels, so the kernels of the default schemes do not carry it), kord_wz > 0 (iv = -2), the tracers of the
// context's table, fillz after each of them when fv3_set_tracer_fill is on (flagstruct%fill); anything else is an error (-2), never a silent fall-back.
//
// Design: column-parallel like the vertical solvers -- one thread per column, consecutive threads on consecutive i, so every
// level access is a coalesced row segment of the [k][NJ][NI] arrays.  The reconstruction (a4(1:4), the interface values, the
// differences and the extremum flags) lives in seven scratch planes of the context instead of per-thread local arrays
// (km = 79..127 levels x 7 arrays would be 4-7 KB of local memory per thread); each sweep keeps its loop-carried and neighbouring
// values in registers.  Measured inside an fv_dynamics step at C384L79 (profiles/prof_fv_dynamics.py, ncu launch list
// profiles/r2/r2_remap_launches.csv): 20 ms per face for six fields (pt, tracer, w, delz, u, v) against 58 ms for the eight
// acoustic substeps it follows (26 ms before the register carries, the four-level load batches of remap_col.cuh, the per-field
// launches and the launch bounds below; one fv_dynamics step 506 -> 476 ms).  The kernels are bound by occupancy x memory-level
// parallelism, not by bytes (1.5 TB/s): every sweep is a chain of L2 / DRAM round trips.
#include "fv3_ctx.hpp"
#include <cmath>
#include <string>
#include <type_traits>
#include <vector>

#include "remap_col.cuh"

namespace {

using namespace rmp;
constexpr int CB = 128;

struct Scr { double *a1, *a2, *a3, *a4, *q, *gam, *fl, *pe4; };
__device__ __forceinline__ Col col_of(const Scr& S, long long o, long long plane) {
  return Col{S.a1 + o, S.a2 + o, S.a3 + o, S.a4 + o, S.q + o, S.gam + o, reinterpret_cast<int*>(S.fl) + o, plane};
}

struct L2E {
  double *pt, *delp, *delz, *w, *u, *v, *pk, *pkz, *omga, *qtr, *pe, *peln;
  const double* qv;      // the specific-humidity tracer (last-step conversion), or nullptr
  const double *ws, *ak, *bk;   // ak, bk: device tables (km + 1)
  double akap, k1k, rrg, ptop, t_min, r_vir;
  int hydrostatic, last_step, kord_mt, kord_wz, kord_tm, use_tracer, kord_tr;
};

// steps 0 - 3.3 of fv_mapz.F90 (:188-526) for the cell columns, in four launches (PART 0: temperature, 1: tracers, 2: w and delz,
// 3: pressure variables, pkz, omega) -- the columns are independent, so the launch boundaries change nothing, and each part
// keeps the register allocation of its batched sweeps (remap_col.cuh) to itself: one kernel for everything needed 128 registers
// and still spilled
template <int PART, bool PPM>
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
  if constexpr (PART == 0) {
  if (a.kord_tm < 0) {   // theta_v -> T_v (:200-230)
    if (a.hydrostatic) for (int k = 1; k <= km; k++) LV(pt, k) = LV(pt, k) * (LV(pk, k + 1) - LV(pk, k)) / (a.akap * (LV(peln, k + 1) - LV(peln, k)));
    else for (int k = 1; k <= km; k++) { const double p = LV(pt, k); LV(pt, k) = p * exp(a.k1k * log(a.rrg * LV(delp, k) / LV(delz, k) * p)); }
  }
  if (!a.hydrostatic) for (int k = 1; k <= km; k++) LV(delz, k) = -LV(delz, k) / LV(delp, k);   // :297-303
  for (int k = 1; k <= km; k++) LV(delp, k) = pe2(k + 1) - pe2(k);                              // :318-331
  // pn2 = log(pe2) is recomputed where it is needed (the temperature map, then the peln / pk update): no column array is kept
  auto pn1 = [&](int k) { return LV(peln, k); };
  auto pn2 = [&](int k) { return (k == 1 || k == km + 1) ? LV(peln, k) : log(pe2(k)); };
  if (a.kord_tm < 0) remap_field<PPM>(C, km, pn1, pn2, pt, 0., 1, a.kord_tm, a.t_min, true);          // :373-386
  else remap_field<PPM>(C, km, pe1, pe2, pt, 0., 1, a.kord_tm, 0., false);
  }
  if constexpr (PART == 1)   // one launch per tracer (a.qtr): a loop over the tracers around the inlined sweeps made the compiler spill 480 bytes
    remap_field<PPM>(C, km, pe1, pe2, a.qtr + o, 0., 0, a.kord_tr, 0., true, a.use_tracer > 5);   // :390-408 (map1_q2; mapn_tracer order for nq > 5)
  if constexpr (PART == 2)
  if (!a.hydrostatic) {                                                                          // :411-433
    remap_field<PPM>(C, km, pe1, pe2, a.w + o, a.ws[o], -2, a.kord_wz, 0., false);
    remap_field<PPM>(C, km, pe1, pe2, delz, 0., 1, a.kord_tm, 0., false);
    for (int k = 1; k <= km; k++) LV(delz, k) = -LV(delz, k) * (pe2(k + 1) - pe2(k));
  }
  if constexpr (PART == 3) {
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
}

// 4.1 / 4.2 (:535-571): u on the south faces (DIR = 0: i in is..ie, j in js..je+1), v on the west faces (DIR = 1)
template <int DIR, bool PPM>
__global__ void __launch_bounds__(CB, 8) k_remap_wind(Lay L, L2E a, Scr S) {
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
  remap_field<PPM>(C, km, pe0, pe3, (DIR == 0 ? a.u : a.v) + o, 0., -1, a.kord_mt, 0., false);
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
  else if (a.qv) {   // T_v -> T (:792-822 with dtmp = 0, no condensates)
    const double* qv = a.qv + o;
    for (int k = 1; k <= km; k++) LV(a.pt + o, k) = (LV(a.pt + o, k) + 0. * LV(a.pkz + o, k)) / (1. + a.r_vir * LV(qv, k));
  }
}

// stand-alone column operator on FV3_WORK_Q (parity of the profiles for every scheme / iv): pe1 = FV3_PE, pe2 = the hybrid levels
template <bool PPM>
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
  remap_field<PPM>(C, km, pe1, pe2, a.qtr + o, iv == -2 ? a.ws[o] : 0., iv, kord, qmin, mode != 1);
}

// fillz (flagstruct%fill: fv_mapz.F90:391, fv_operators.F90:337) of one tracer with the layer thicknesses dp, one thread per column
__global__ void __launch_bounds__(CB) k_fillz(Lay L, double* q, const double* dp) {
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const int t = blockIdx.x * CB + threadIdx.x;
  if (t >= nx * ny) return;
  const long long o = LIDX(L, L.is + t % nx, L.js + t / nx);
  fillz_column(L.npz, q + o, dp + o, L.plane);
}

// abs(kord) in 8..15: cs_profile / scalar_profile; 1..7: ppm_profile (the PPM instantiations of the kernels)
int check_kord(fv3_ctx* c, int kord, const char* what, bool* ppm) {
  const int ak = kord < 0 ? -kord : kord;
  if (ak < 1 || ak > 15) return fv3_fail(c, -2, std::string("remap: ") + what + " outside 1..15");
  if (ak <= 7) *ppm = true;
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

// fillz of FV3_WORK_Q with the thicknesses FV3_DELP (compute domain)
int stage_fillz(fv3_ctx* c) {
  StageScope ts(c, "REMAP_OP");
  if (c->L.npz < 3) return fv3_fail(c, -2, "fillz: npz >= 3");
  const int n = (c->L.ie - c->L.is + 1) * (c->L.je - c->L.js + 1);
  k_fillz<<<(n + CB - 1) / CB, CB, 0, c->stream>>>(c->L, c->fld[FV3_WORK_Q], c->fld[FV3_DELP]);
  c->launches++;
  FV3_CUDA(c, cudaGetLastError());
  return 0;
}

int stage_remap_work_q(fv3_ctx* c, int mode, int iv, int kord, double qmin) {
  StageScope ts(c, "REMAP_OP");
  if (mode < 0 || mode > 2 || iv < -2 || iv > 2) return fv3_fail(c, -1, "remap_work_q: mode in 0..2, iv in -2..2");
  bool ppm = false;
  int rc = check_kord(c, kord, "kord", &ppm); if (rc) return rc;
  if (ppm && c->L.npz < 5) return fv3_fail(c, -2, "remap: ppm_profile (kord <= 7) needs npz >= 5");
  L2E a{}; Scr S{};
  rc = fill(c, a, S); if (rc) return rc;
  const int n = (c->L.ie - c->L.is + 1) * (c->L.je - c->L.js + 1);
  if (ppm) k_remap_work_q<true><<<(n + CB - 1) / CB, CB, 0, c->stream>>>(c->L, a, S, mode, iv, kord, qmin);
  else k_remap_work_q<false><<<(n + CB - 1) / CB, CB, 0, c->stream>>>(c->L, a, S, mode, iv, kord, qmin);
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
  bool ppm = false;   // any field with abs(kord) <= 7: the launches below use the instantiations that hold ppm_profile
  int rc = check_kord(c, kord_mt, "kord_mt", &ppm); if (rc) return rc;
  rc = check_kord(c, kord_tm, "kord_tm", &ppm); if (rc) return rc;
  if (!f.hydrostatic) { rc = check_kord(c, kord_wz, "kord_wz", &ppm); if (rc) return rc; }
  if (use_tracer < 0 || use_tracer > 64) return fv3_fail(c, -1, "remap: use_tracer (the number of tracers to remap) in 0..64");
  if (use_tracer) {
    if (kord_tr < 0) return fv3_fail(c, -2, "remap: kord_tr must be positive");
    // more than 5 tracers: mapn_tracer (fv_operators.F90:262-273) has no ppm_profile branch -- it calls scalar_profile whatever
    // kord is, and scalar_profile treats abs(kord) 0..8 alike (`case (0:8)`, :756; abs(kord) > 9 / >= 14 false, <= 13 true)
    if (use_tracer > 5 && kord_tr >= 1 && kord_tr <= 7) kord_tr = 8;
    rc = check_kord(c, kord_tr, "kord_tr", &ppm); if (rc) return rc;
  }
  if (ppm && c->L.npz < 5) return fv3_fail(c, -2, "remap: ppm_profile (kord <= 7) needs npz >= 5");
  L2E a{}; Scr S{};
  rc = fill(c, a, S); if (rc) return rc;
  if (use_tracer) {   // the first use_tracer tracers of the context's table (fv3_set_num_tracers; one tracer: FV3_WORK_Q's own array)
    if (c->tracers.empty()) c->tracers.push_back(c->fld[FV3_WORK_Q]);
    if (use_tracer > (int)c->tracers.size()) return fv3_fail(c, -1, "remap: use_tracer exceeds the number of tracers of the context");
  }
  if (sphum >= use_tracer) return fv3_fail(c, -1, "remap: sphum must be one of the remapped tracers (or -1)");
  if (sphum >= 0 && c->f.use_cond) return fv3_fail(c, -2, "remap: specific humidity with use_cond (condensates, moist_cv) not supported");
  a.qv = sphum < 0 ? nullptr : c->tracers[sphum]; a.r_vir = r_vir;
  a.last_step = last_step; a.kord_mt = kord_mt; a.kord_wz = kord_wz; a.kord_tm = kord_tm; a.use_tracer = use_tracer; a.kord_tr = kord_tr;
  const Lay& L = c->L;
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const int gc = (nx * ny + CB - 1) / CB;
  auto launch = [&](auto P) {   // P: std::bool_constant, the PPM instantiation or not
    constexpr bool PPM = decltype(P)::value;
    k_remap_cells<0, PPM><<<gc, CB, 0, c->stream>>>(L, a, S);
    for (int iq = 0; iq < use_tracer; iq++) {
      a.qtr = c->tracers[iq];
      k_remap_cells<1, PPM><<<gc, CB, 0, c->stream>>>(L, a, S);
      c->launches++;
      if (c->tracer_fill) { k_fillz<<<gc, CB, 0, c->stream>>>(L, a.qtr, a.delp); c->launches++; }   // delp holds dp2 since PART 0
    }
    if (!f.hydrostatic) { k_remap_cells<2, PPM><<<gc, CB, 0, c->stream>>>(L, a, S); c->launches++; }
    k_remap_cells<3, false><<<gc, CB, 0, c->stream>>>(L, a, S);   // (pressure variables only: no profile)
    c->launches++;
    k_remap_wind<0, PPM><<<(nx * (ny + 1) + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
    k_remap_wind<1, PPM><<<((nx + 1) * ny + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
  };
  if (ppm) launch(std::true_type{}); else launch(std::false_type{});
  k_remap_finish<<<(nx * ny + CB - 1) / CB, CB, 0, c->stream>>>(L, a, S);
  c->launches += 4;
  FV3_CUDA(c, cudaGetLastError());
  return 0;
}
