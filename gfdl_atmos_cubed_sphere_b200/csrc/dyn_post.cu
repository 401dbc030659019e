// Post-loop part of dyn_core: the del-2 filter `del2_cubed` and the dissipative-heating update of pt.
//
// Reference semantics: model/dyn_core.F90:2356-2465 (del2_cubed; non-USE_SG build: del6_u / del6_v), :296-307 (n_con),
// :1300-1356 (heat_source -> pt).  del2_cubed is also the omega filter of fv_dynamics (fv_dynamics.F90:637-642).
//
// Design: the reference updates q in place, pass by pass, after (a) averaging the three cells around each cube corner and
// (b) copy_corners(dir=1) before the x fluxes / copy_corners(dir=2) before the y fluxes.  Here one pass is ONE kernel
// reading q_in and writing q_out (ping-pong with a scratch field): (a) and (b) are read-side remaps (ppm::QAccX / QAccY
// index maps composed with the corner average), the four fluxes of a cell are evaluated on the fly, so no flux plane and no
// corner-filled copy ever reaches HBM.  The 3x3 corner blocks of the halo are never read through anything but the remaps,
// so they are copied through unchanged.
#include "fv3_ctx.hpp"
#include "ppm.cuh"
#include <algorithm>
#include <cmath>

#define TI 32
#define TJ 8
#define G2(p, i, j) __ldg((G.p) + LIDX(L, (i), (j)))

namespace {

struct Del2Read {
  const double* q; Lay L;
  // value of q(i, j) after the corner averaging of dyn_core.F90:2410-2429 ((i, j) is not inside a corner block)
  __device__ __forceinline__ double avg(int i, int j) const {
    const double r3 = 1. / 3.;
    if (L.cube) {
      const int npx = L.npx, npy = L.npy, ie = L.ie, je = L.je;
      const bool lo_i = (i == 0 || i == 1), hi_i = (i == ie || i == npx), lo_j = (j == 0 || j == 1), hi_j = (j == je || j == npy);
      if ((lo_i || hi_i) && (lo_j || hi_j)) {
        const int ic = lo_i ? 1 : ie, io = lo_i ? 0 : npx, jc = lo_j ? 1 : je, jo = lo_j ? 0 : npy;
        // the three averaged cells of this corner: (ic, jc), (io, jc), (ic, jo); (io, jo) is a corner-block cell
        if (!(i == io && j == jo))
          return (__ldg(q + LIDX(L, ic, jc)) + __ldg(q + LIDX(L, io, jc)) + __ldg(q + LIDX(L, ic, jo))) * r3;
      }
    }
    return __ldg(q + LIDX(L, i, j));
  }
  // copy_corners(dir = 1) view (tp_core.F90:257-286), then the corner average
  __device__ __forceinline__ double x(int i, int j) const {
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
    return avg(ii, jj);
  }
  // copy_corners(dir = 2) view (tp_core.F90:288-318), then the corner average
  __device__ __forceinline__ double y(int i, int j) const {
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
    return avg(ii, jj);
  }
};

// one pass of del2_cubed with shrink parameter nt (= ntimes - n): updates (is-nt:ie+nt, js-nt:je+nt), copies the rest
__global__ void __launch_bounds__(TI* TJ) k_del2_iter(Lay L, DevGrid G, const double* __restrict__ qin, double* __restrict__ qout, double cd, int nt) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  const long long ko = (long long)blockIdx.z * L.plane;
  if (i < L.isd || i > L.ied || j > L.jed) return;
  const long long o = ko + LIDX(L, i, j);
  const bool corner_block = L.cube && (i < 1 || i >= L.npx) && (j < 1 || j >= L.npy);
  if (corner_block) { qout[o] = __ldg(qin + o); return; }
  const Del2Read R{qin + ko, L};
  const double q0 = R.avg(i, j);
  if (i < L.is - nt || i > L.ie + nt || j < L.js - nt || j > L.je + nt) { qout[o] = q0; return; }
  // fx(i,j) = del6_v(i,j) (q(i-1,j) - q(i,j)), fy(i,j) = del6_u(i,j) (q(i,j-1) - q(i,j))   (:2434-2451)
  const double fx0 = G2(del6_v, i, j) * (R.x(i - 1, j) - R.x(i, j));
  const double fx1 = G2(del6_v, i + 1, j) * (R.x(i, j) - R.x(i + 1, j));
  const double fy0 = G2(del6_u, i, j) * (R.y(i, j - 1) - R.y(i, j));
  const double fy1 = G2(del6_u, i, j + 1) * (R.y(i, j) - R.y(i, j + 1));
  qout[o] = q0 + cd * G2(rarea, i, j) * (fx0 - fx1 + fy0 - fy1);   // :2456-2460
}

// dyn_core.F90:1305-1356
__global__ void __launch_bounds__(TI* TJ) k_dcon_heat(Lay L, double* __restrict__ pt, double* __restrict__ hs, const double* __restrict__ delp,
                                                    const double* __restrict__ delz, double* __restrict__ pkz, double bdt, double delt_max,
                                                    double cp_air, double cv_air, double rdg, double k1k, int hydrostatic) {
  const int i = L.is + blockIdx.x * TI + threadIdx.x;
  const int j = L.js + blockIdx.y * TJ + threadIdx.y;
  const int k = blockIdx.z;   // 0-based level; reference k = k + 1
  if (i > L.ie || j > L.je) return;
  const long long o = (long long)k * L.plane + LIDX(L, i, j);
  const double h = hs[o], dp = __ldg(delp + o);
  if (hydrostatic) {
    if (k + 1 < 3) { pt[o] = pt[o] + h / (cp_air * dp * __ldg(pkz + o)); return; }
    const double dtmp = h / (cp_air * dp);
    pt[o] = pt[o] + copysign(fmin(fabs(bdt) * delt_max, fabs(dtmp)), dtmp) / __ldg(pkz + o);
    hs[o] = dtmp;
    return;
  }
  double delt = fabs(bdt * delt_max);
  if (k == 0) delt = 0.1 * delt;
  if (k == 1) delt = 0.5 * delt;
  const double p = pt[o];
  const double pz = exp(k1k * log(rdg * dp / __ldg(delz + o) * p));
  pkz[o] = pz;
  const double dtmp = h / (cv_air * dp);
  pt[o] = p + copysign(fmin(delt, fabs(dtmp)), dtmp) / pz;
  hs[o] = dtmp;
}

// fv_dynamics.F90:303-328, :377-398
__global__ void __launch_bounds__(TI* TJ) k_pt_to_theta(Lay L, double* __restrict__ pt, const double* __restrict__ delp, const double* __restrict__ delz,
                                                      const double* __restrict__ qv, const double* __restrict__ qcon, double* __restrict__ dp1,
                                                      double* __restrict__ pkz, double zvir, double rdg, double kappa, int hydrostatic, int use_cond) {
  const int i = L.is + blockIdx.x * TI + threadIdx.x;
  const int j = L.js + blockIdx.y * TJ + threadIdx.y;
  if (i > L.ie || j > L.je) return;
  const long long o = (long long)blockIdx.z * L.plane + LIDX(L, i, j);
  const double d1 = zvir * __ldg(qv + o), p = pt[o];
  dp1[o] = d1;
  double pz;
  if (!hydrostatic) { pz = exp(kappa * log(rdg * __ldg(delp + o) * p * (1. + d1) / __ldg(delz + o))); pkz[o] = pz; }
  else pz = pkz[o];
  pt[o] = use_cond ? p * (1. + d1) * (1. - __ldg(qcon + o)) / pz : p * (1. + d1) / pz;
}

}  // namespace

// dyn_core.F90:296-307
int fv3_n_con(const fv3_flags_t& f, int npz) {
  if (f.convert_ke || (f.do_vort_damp && f.vtdm4 > 1.E-4)) return npz;
  if (f.d2_bg_k1 < 1.E-3) return 0;
  return (f.d2_bg_k2 < 1.E-3) ? 1 : 2;
}

// del2_cubed on FV3_HEAT or FV3_OMGA; the caller has updated the halo of the field (dyn_core.F90:2401)
int stage_del2_cubed(fv3_ctx* c, int field, double cd, int nmax) {
  StageScope ts(c, "DEL2_CUBED");
  if (field != FV3_HEAT && field != FV3_OMGA) return fv3_fail(c, -1, "del2_cubed: field must be FV3_HEAT or FV3_OMGA");
  const Lay& L = c->L;
  const int nk = L.npz, ntimes = std::min(3, nmax);
  if (ntimes < 1) return 0;
  double* a = c->fld[field];
  double* b = c->scr[0];
  const dim3 blk(TI, TJ), grd((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, nk);
  for (int n = 1; n <= ntimes; n++) {
    k_del2_iter<<<grd, blk, 0, c->stream>>>(L, c->G, a, b, cd, ntimes - n);
    c->launches++;
    std::swap(a, b);
  }
  if (a != c->fld[field]) FV3_CUDA(c, cudaMemcpyAsync(c->fld[field], a, (size_t)L.plane * nk * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

// dyn_core.F90:1305-1356 (after the del2_cubed call)
int stage_dcon_heating(fv3_ctx* c, double bdt) {
  StageScope ts(c, "DCON_HEAT");
  const fv3_flags_t& f = c->f;
  const int n_con = fv3_n_con(f, c->L.npz);
  if (n_con == 0 || !(f.d_con > 1.e-5)) return 0;
  if (!f.hydrostatic && f.moist_kappa) return fv3_fail(c, -2, "dcon_heating: moist_kappa branch (dyn_core.F90:1336-1344) not supported");
  const Lay& L = c->L;
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const dim3 blk(TI, TJ), grd((nx + TI - 1) / TI, (ny + TJ - 1) / TJ, n_con);
  k_dcon_heat<<<grd, blk, 0, c->stream>>>(L, c->fld[FV3_PT], c->fld[FV3_HEAT], c->fld[FV3_DELP], c->fld[FV3_DELZ], c->fld[FV3_PKZ], bdt,
                                           f.delt_max, f.cp_air, f.cp_air - f.rdgas, -f.rdgas / f.grav, f.kappa / (1. - f.kappa),
                                           f.hydrostatic ? 1 : 0);
  c->launches++;
  return 0;
}

// fv_dynamics.F90:303-328 + :377-398: the entry conversion of fv_dynamics (temperature -> virtual potential temperature,
// pkz from the gas law); specific humidity is read from FV3_WORK_Q, dp1 = zvir q_v is left in FV3_DP1
int stage_pt_to_theta(fv3_ctx* c, double zvir) {
  StageScope ts(c, "PT_TO_THETA");
  const fv3_flags_t& f = c->f;
  if (f.moist_kappa) return fv3_fail(c, -2, "pt_to_theta: moist_kappa (moist_cv) not supported");
  const Lay& L = c->L;
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  const dim3 blk(TI, TJ), grd((nx + TI - 1) / TI, (ny + TJ - 1) / TJ, L.npz);
  k_pt_to_theta<<<grd, blk, 0, c->stream>>>(L, c->fld[FV3_PT], c->fld[FV3_DELP], c->fld[FV3_DELZ], c->fld[FV3_WORK_Q], c->fld[FV3_QCON],
                                             c->fld[FV3_DP1], c->fld[FV3_PKZ], zvir, -f.rdgas / f.grav, f.kappa, f.hydrostatic ? 1 : 0,
                                             f.use_cond ? 1 : 0);
  c->launches++;
  return 0;
}
