// d_sw: full-step D-grid shallow-water sweep for all levels, batched over k.
//
// Reference semantics: model/sw_core.F90:494-1606 d_sw (non-SW_DYNAMICS, AM4 variant of the
// pt damping :1014-1016, inline_q = F), :1608-1737 del6_vt_flux, :2154-2998 xtp_u / ytp_v;
// per-level damping prologue and call site model/dyn_core.F90:666-812; fill_corners
// tools/fv_mp_mod.F90:1031-1062 (B-grid) and :1249-1281 (D-grid vector).
//
// Design notes (B200-first, not a translation):
//  * k is batched in gridDim.z; the per-level damping orders/coefficients that dyn_core
//    computes inside its OpenMP k-loop are small device tables read by blockIdx.z.
//  * contravariant winds incl. the 4 edge strips and the 4 2x2 corner systems are a
//    closed-form per-point evaluation (no ordered passes over edges/corners).
//  * every "fill corners then sweep" of the reference is a read-side index remap.
//  * prognostics are ping-ponged (fld <-> alt) so no kernel reads what a neighbour thread
//    writes; the writers copy the halo through so a frozen halo stays frozen.
//  * uc, vc, divg_d are left intact (the reference trashes them as scratch, sw_core.F90:
//    1392-1424; the next c_sw overwrites them anyway).
#include "tp2d.cuh"
#include "ppm.cuh"
#include "tp_tile.cuh"
#include "tp_line.cuh"
#include <cmath>

using namespace ppm;

// FV3_TP_LINES: 2 (default) line-per-warp kernels on the interior tiles and, for the multi-field transport, on the frame tiles;
// 3: on the frame tiles of the single-field transports too; 1: interior tiles only; 0: nowhere (the first-generation tile kernel
// of tp_tile.cuh everywhere) -- A/B timing and bisecting
static int use_line_kernels() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FV3_TP_LINES"); on = e ? atoi(e) : 2; }
  return on;
}

#define TI 32
#define TJ 8
#ifndef PLANE_MINB
#define PLANE_MINB 6
#endif
#define PLANE_IJK                                              \
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x; \
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;          \
  const int k = blockIdx.z;                                     \
  const long long ko = (long long)k * L.plane;
#define AT(p, i, j) __ldg((p) + ko + LIDX(L, (i), (j)))
#define G2(p, i, j) __ldg((G.p) + LIDX(L, (i), (j)))
#define SG(n, i, j) __ldg(G.sin_sg + (long long)((n)-1) * L.plane + LIDX(L, (i), (j)))

static inline dim3 plane_grid(const Lay& L, int nk) { return dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, nk); }
void launch_deln_add(fv3_ctx* c, double* fx, double* fy, const double* fx2, const double* fy2, const double* mass,
                     int slot_damp, double damp_const, int nk);
int launch_a2b_ord4(fv3_ctx* c, const double* qin, double* qout, int nk, int replace_into_qin);

// ---------------------------------------------------------------------------------------------
// contravariant winds, Courant numbers, area fluxes  (sw_core.F90:653-917)
// ---------------------------------------------------------------------------------------------
struct WindCtx {
  const double *uc, *vc; Lay L; DevGrid G; long long ko; double dt;
  __device__ __forceinline__ double UC(int i, int j) const { return __ldg(uc + ko + LIDX(L, i, j)); }
  __device__ __forceinline__ double VC(int i, int j) const { return __ldg(vc + ko + LIDX(L, i, j)); }
  __device__ __forceinline__ double CU(int i, int j) const { return __ldg(G.cosa_u + LIDX(L, i, j)); }
  __device__ __forceinline__ double CV(int i, int j) const { return __ldg(G.cosa_v + LIDX(L, i, j)); }
  __device__ __forceinline__ double S(int n, int i, int j) const { return __ldg(G.sin_sg + (long long)(n - 1) * L.plane + LIDX(L, i, j)); }
  // generic interior formulas (:670-687)
  __device__ __forceinline__ double UTg(int i, int j) const {
    return (UC(i, j) - 0.25 * CU(i, j) * (VC(i - 1, j) + VC(i, j) + VC(i - 1, j + 1) + VC(i, j + 1))) * __ldg(G.rsin_u + LIDX(L, i, j));
  }
  __device__ __forceinline__ double VTg(int i, int j) const {
    return (VC(i, j) - 0.25 * CV(i, j) * (UC(i, j - 1) + UC(i + 1, j - 1) + UC(i, j) + UC(i + 1, j))) * __ldg(G.rsin_v + LIDX(L, i, j));
  }
  // face-edge values (:692-699, :709-716, :727-735, :746-753)
  __device__ __forceinline__ double UTe(int i, int j) const {
    const double u = UC(i, j);
    return (u * dt > 0.) ? u / S(3, i - 1, j) : u / S(1, i, j);
  }
  __device__ __forceinline__ double VTe(int i, int j) const {
    const double v = VC(i, j);
    return (v * dt > 0.) ? v / S(4, i, j - 1) : v / S(2, i, j);
  }
  __device__ __forceinline__ double UTb(int i, int j) const { return (i == 1 || i == L.npx) ? UTe(i, j) : UTg(i, j); }
  __device__ __forceinline__ double VTb(int i, int j) const { return (j == 1 || j == L.npy) ? VTe(i, j) : VTg(i, j); }

  __device__ double ut_final(int i, int j) const {
    if (!L.cube) return UC(i, j);
    const int npx = L.npx, npy = L.npy;
    if (i == 1 || i == npx) return UTe(i, j);
    const bool erow = (j == 0 || j == 1 || j == npy - 1 || j == npy);
    if (!erow) return UTg(i, j);
    if (i >= 3 && i <= npx - 2)  // :737-742, :754-759
      return UC(i, j) - 0.25 * CU(i, j) * (VTb(i - 1, j) + VTb(i, j) + VTb(i - 1, j + 1) + VTb(i, j + 1));
    // 2x2 corner systems (:772-844)
    if (i == 2 && j == 0) {
      const double damp = 1. / (1. - 0.0625 * CU(2, 0) * CV(1, 0));
      return (UC(2, 0) - 0.25 * CU(2, 0) * (VTb(1, 1) + VTb(2, 1) + VTb(2, 0) + VC(1, 0) - 0.25 * CV(1, 0) * (UTb(1, 0) + UTb(1, -1) + UTb(2, -1)))) * damp;
    }
    if (i == 2 && j == 1) {
      const double damp = 1. / (1. - 0.0625 * CU(2, 1) * CV(1, 2));
      return (UC(2, 1) - 0.25 * CU(2, 1) * (VTb(1, 1) + VTb(2, 1) + VTb(2, 2) + VC(1, 2) - 0.25 * CV(1, 2) * (UTb(1, 1) + UTb(1, 2) + UTb(2, 2)))) * damp;
    }
    if (i == npx - 1 && j == 0) {
      const double damp = 1. / (1. - 0.0625 * CU(npx - 1, 0) * CV(npx - 1, 0));
      return (UC(npx - 1, 0) - 0.25 * CU(npx - 1, 0) * (VTb(npx - 1, 1) + VTb(npx - 2, 1) + VTb(npx - 2, 0) + VC(npx - 1, 0) -
                                                        0.25 * CV(npx - 1, 0) * (UTb(npx, 0) + UTb(npx, -1) + UTb(npx - 1, -1)))) * damp;
    }
    if (i == npx - 1 && j == 1) {
      const double damp = 1. / (1. - 0.0625 * CU(npx - 1, 1) * CV(npx - 1, 2));
      return (UC(npx - 1, 1) - 0.25 * CU(npx - 1, 1) * (VTb(npx - 1, 1) + VTb(npx - 2, 1) + VTb(npx - 2, 2) + VC(npx - 1, 2) -
                                                        0.25 * CV(npx - 1, 2) * (UTb(npx, 1) + UTb(npx, 2) + UTb(npx - 1, 2)))) * damp;
    }
    if (i == npx - 1 && j == npy) {
      const double damp = 1. / (1. - 0.0625 * CU(npx - 1, npy) * CV(npx - 1, npy + 1));
      return (UC(npx - 1, npy) - 0.25 * CU(npx - 1, npy) * (VTb(npx - 1, npy) + VTb(npx - 2, npy) + VTb(npx - 2, npy + 1) + VC(npx - 1, npy + 1) -
                                                            0.25 * CV(npx - 1, npy + 1) * (UTb(npx, npy) + UTb(npx, npy + 1) + UTb(npx - 1, npy + 1)))) * damp;
    }
    if (i == npx - 1 && j == npy - 1) {
      const double damp = 1. / (1. - 0.0625 * CU(npx - 1, npy - 1) * CV(npx - 1, npy - 1));
      return (UC(npx - 1, npy - 1) - 0.25 * CU(npx - 1, npy - 1) * (VTb(npx - 1, npy) + VTb(npx - 2, npy) + VTb(npx - 2, npy - 1) + VC(npx - 1, npy - 1) -
                                                                    0.25 * CV(npx - 1, npy - 1) * (UTb(npx, npy - 1) + UTb(npx, npy - 2) + UTb(npx - 1, npy - 2)))) * damp;
    }
    if (i == 2 && j == npy) {
      const double damp = 1. / (1. - 0.0625 * CU(2, npy) * CV(1, npy + 1));
      return (UC(2, npy) - 0.25 * CU(2, npy) * (VTb(1, npy) + VTb(2, npy) + VTb(2, npy + 1) + VC(1, npy + 1) -
                                                0.25 * CV(1, npy + 1) * (UTb(1, npy) + UTb(1, npy + 1) + UTb(2, npy + 1)))) * damp;
    }
    if (i == 2 && j == npy - 1) {
      const double damp = 1. / (1. - 0.0625 * CU(2, npy - 1) * CV(1, npy - 1));
      return (UC(2, npy - 1) - 0.25 * CU(2, npy - 1) * (VTb(1, npy) + VTb(2, npy) + VTb(2, npy - 1) + VC(1, npy - 1) -
                                                        0.25 * CV(1, npy - 1) * (UTb(1, npy - 1) + UTb(1, npy - 2) + UTb(2, npy - 2)))) * damp;
    }
    return UTg(i, j);  // never set by the reference at these rows (i = 0, npx+1)
  }

  __device__ double vt_final(int i, int j) const {
    if (!L.cube) return VC(i, j);
    const int npx = L.npx, npy = L.npy;
    if (j == 1 || j == npy) return VTe(i, j);
    const bool ecol = (i == 0 || i == 1 || i == npx - 1 || i == npx);
    if (!ecol) return VTg(i, j);
    if (j >= 3 && j <= npy - 2)  // :700-705, :718-723
      return VC(i, j) - 0.25 * CV(i, j) * (UTb(i, j - 1) + UTb(i + 1, j - 1) + UTb(i, j) + UTb(i + 1, j));
    if (i == 0 && j == 2) {
      const double damp = 1. / (1. - 0.0625 * CU(0, 1) * CV(0, 2));
      return (VC(0, 2) - 0.25 * CV(0, 2) * (UTb(1, 1) + UTb(1, 2) + UTb(0, 2) + UC(0, 1) - 0.25 * CU(0, 1) * (VTb(0, 1) + VTb(-1, 1) + VTb(-1, 2)))) * damp;
    }
    if (i == 1 && j == 2) {
      const double damp = 1. / (1. - 0.0625 * CU(2, 1) * CV(1, 2));
      return (VC(1, 2) - 0.25 * CV(1, 2) * (UTb(1, 1) + UTb(1, 2) + UTb(2, 2) + UC(2, 1) - 0.25 * CU(2, 1) * (VTb(1, 1) + VTb(2, 1) + VTb(2, 2)))) * damp;
    }
    if (i == npx && j == 2) {
      const double damp = 1. / (1. - 0.0625 * CU(npx + 1, 1) * CV(npx, 2));
      return (VC(npx, 2) - 0.25 * CV(npx, 2) * (UTb(npx, 1) + UTb(npx, 2) + UTb(npx + 1, 2) + UC(npx + 1, 1) -
                                                0.25 * CU(npx + 1, 1) * (VTb(npx, 1) + VTb(npx + 1, 1) + VTb(npx + 1, 2)))) * damp;
    }
    if (i == npx - 1 && j == 2) {
      const double damp = 1. / (1. - 0.0625 * CU(npx - 1, 1) * CV(npx - 1, 2));
      return (VC(npx - 1, 2) - 0.25 * CV(npx - 1, 2) * (UTb(npx, 1) + UTb(npx, 2) + UTb(npx - 1, 2) + UC(npx - 1, 1) -
                                                        0.25 * CU(npx - 1, 1) * (VTb(npx - 1, 1) + VTb(npx - 2, 1) + VTb(npx - 2, 2)))) * damp;
    }
    if (i == npx && j == npy - 1) {
      const double damp = 1. / (1. - 0.0625 * CU(npx + 1, npy - 1) * CV(npx, npy - 1));
      return (VC(npx, npy - 1) - 0.25 * CV(npx, npy - 1) * (UTb(npx, npy - 1) + UTb(npx, npy - 2) + UTb(npx + 1, npy - 2) + UC(npx + 1, npy - 1) -
                                                            0.25 * CU(npx + 1, npy - 1) * (VTb(npx, npy) + VTb(npx + 1, npy) + VTb(npx + 1, npy - 1)))) * damp;
    }
    if (i == npx - 1 && j == npy - 1) {
      const double damp = 1. / (1. - 0.0625 * CU(npx - 1, npy - 1) * CV(npx - 1, npy - 1));
      return (VC(npx - 1, npy - 1) - 0.25 * CV(npx - 1, npy - 1) * (UTb(npx, npy - 1) + UTb(npx, npy - 2) + UTb(npx - 1, npy - 2) + UC(npx - 1, npy - 1) -
                                                                    0.25 * CU(npx - 1, npy - 1) * (VTb(npx - 1, npy) + VTb(npx - 2, npy) + VTb(npx - 2, npy - 1)))) * damp;
    }
    if (i == 0 && j == npy - 1) {
      const double damp = 1. / (1. - 0.0625 * CU(0, npy - 1) * CV(0, npy - 1));
      return (VC(0, npy - 1) - 0.25 * CV(0, npy - 1) * (UTb(1, npy - 1) + UTb(1, npy - 2) + UTb(0, npy - 2) + UC(0, npy - 1) -
                                                        0.25 * CU(0, npy - 1) * (VTb(0, npy) + VTb(-1, npy) + VTb(-1, npy - 1)))) * damp;
    }
    if (i == 1 && j == npy - 1) {
      const double damp = 1. / (1. - 0.0625 * CU(2, npy - 1) * CV(1, npy - 1));
      return (VC(1, npy - 1) - 0.25 * CV(1, npy - 1) * (UTb(1, npy - 1) + UTb(1, npy - 2) + UTb(2, npy - 2) + UC(2, npy - 1) -
                                                        0.25 * CU(2, npy - 1) * (VTb(1, npy) + VTb(2, npy) + VTb(2, npy - 1)))) * damp;
    }
    return VTg(i, j);
  }
};

__global__ void __launch_bounds__(TI* TJ, PLANE_MINB) k_dsw_wind(Lay L, DevGrid G, const double* __restrict__ uc, const double* __restrict__ vc,
                                                    double* __restrict__ uts, double* __restrict__ vts, double* __restrict__ crx,
                                                    double* __restrict__ cry, double* __restrict__ xfx, double* __restrict__ yfx,
                                                    double* __restrict__ cx, double* __restrict__ cy, double dt) {
  PLANE_IJK
  if (i < L.isd || i > L.ied + 1 || j > L.jed + 1) return;
  WindCtx W{uc, vc, L, G, ko, dt};
  const long long o = ko + LIDX(L, i, j);
  // away from the face edges the closed-form edge / corner cases cannot apply: plain interior formulas, no call
  const bool inner = L.cube && i >= 3 && i <= L.npx - 2 && j >= 3 && j <= L.npy - 2;
  // ut, vt themselves are read again only by k_dsw_ke, and there only in the edge / corner formulas (sw_core.F90:1078-1228:
  // i or j within two cells of a face edge); everywhere else they need not reach HBM
  const bool keep = !L.cube || i <= 4 || i >= L.npx - 3 || j <= 4 || j >= L.npy - 3;
  // ut on (is-1:ie+2, jsd:jed)
  if (j <= L.jed && i >= L.is - 1 && i <= L.ie + 2) {
    const double ut = inner ? W.UTg(i, j) : W.ut_final(i, j);
    if (keep) uts[o] = ut;
    if (i >= L.is && i <= L.ie + 1) {  // :863-890, :923-927
      // both upwind candidates of the metric terms are loaded before the sign of ut is known (one memory round trip, not two)
      const double rm = G2(rdxa, i - 1, j), r0 = G2(rdxa, i, j), s3 = SG(3, i - 1, j), s1 = SG(1, i, j), dyv = G2(dy, i, j), cx0 = cx[o];
      double xf = dt * ut, cr;
      if (xf > 0.) { cr = xf * rm; xf = dyv * xf * s3; }
      else { cr = xf * r0; xf = dyv * xf * s1; }
      crx[o] = cr; xfx[o] = xf; cx[o] = cx0 + cr;
    }
  }
  // vt on (isd:ied, js-1:je+2)
  if (i <= L.ied && j >= L.js - 1 && j <= L.je + 2) {
    const double vt = inner ? W.VTg(i, j) : W.vt_final(i, j);
    if (keep) vts[o] = vt;
    if (j >= L.js && j <= L.je + 1) {  // :869-902, :933-936
      const double rm = G2(rdya, i, j - 1), r0 = G2(rdya, i, j), s4 = SG(4, i, j - 1), s2 = SG(2, i, j), dxv = G2(dx, i, j), cy0 = cy[o];
      double yf = dt * vt, cr;
      if (yf > 0.) { cr = yf * rm; yf = dxv * yf * s4; }
      else { cr = yf * r0; yf = dxv * yf * s2; }
      cry[o] = cr; yfx[o] = yf; cy[o] = cy0 + cr;
    }
  }
}

// SW_DYNAMICS build, test_case == 1 (sw_core.F90:626-651): Courant numbers / area fluxes from the PRESCRIBED C-grid winds
__global__ void __launch_bounds__(TI* TJ) k_dsw_wind_sw1(Lay L, DevGrid G, const double* __restrict__ uc, const double* __restrict__ vc,
                                                       double* __restrict__ crx, double* __restrict__ cry, double* __restrict__ xfx,
                                                       double* __restrict__ yfx, double* __restrict__ cx, double* __restrict__ cy, double dt) {
  PLANE_IJK
  if (i < L.isd || i > L.ied + 1 || j > L.jed + 1) return;
  const long long o = ko + LIDX(L, i, j);
  if (j <= L.jed && i >= L.is && i <= L.ie + 1) {
    const double sa = G2(sina_u, i, j);
    double xf = dt * __ldg(uc + o) / sa;
    const double cr = xf > 0. ? xf * G2(rdxa, i - 1, j) : xf * G2(rdxa, i, j);
    xf = G2(dy, i, j) * xf * sa;
    crx[o] = cr; xfx[o] = xf; cx[o] = cx[o] + cr;
  }
  if (i <= L.ied && j >= L.js && j <= L.je + 1) {
    const double sa = G2(sina_v, i, j);
    double yf = dt * __ldg(vc + o) / sa;
    const double cr = yf > 0. ? yf * G2(rdya, i, j - 1) : yf * G2(rdya, i, j);
    yf = G2(dx, i, j) * yf * sa;
    cry[o] = cr; yfx[o] = yf; cy[o] = cy[o] + cr;
  }
}

// ---------------------------------------------------------------------------------------------
// scalar updates
// ---------------------------------------------------------------------------------------------
// w damping increment and heating (sw_core.F90:951-982); hs = heat_s work plane, ds = diss_e
__global__ void __launch_bounds__(TI* TJ) k_dsw_dw(Lay L, DevGrid G, const double* __restrict__ w, const double* __restrict__ fx2,
                                                  const double* __restrict__ fy2, double* __restrict__ dw, double* __restrict__ hs,
                                                  double* __restrict__ ds, const double* kdbl, double kgb, double dt, int prevent, int do_diss,
                                                  int need_hs, int kofs) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  const int k = blockIdx.z + kofs;
  const long long ko = (long long)k * L.plane;
  if (i < L.is || i > L.ie || j < L.js || j > L.je) return;
  const long long o = ko + LIDX(L, i, j);
  const double coef = kdbl[KD_DAMP4_W * (L.npz + 1) + k];
  // levels without w damping: dw is never read (the transport tests the same coefficient); hs, ds only by k_dsw_heat
  if (coef == 0.) { if (need_hs) { hs[o] = 0.; ds[o] = 0.; } return; }
  const double dd8 = kgb * fabs(dt);
  const double d = (fx2[o] - fx2[o + 1] + fy2[o] - fy2[o + L.NI]) * G2(rarea, i, j);
  dw[o] = d;
  const double wv = __ldg(w + o);
  const double tmp = d * (wv + 0.5 * d);
  if (prevent) { hs[o] = dd8 - fmin(0., tmp); ds[o] = do_diss ? dd8 - tmp : 0.; }
  else { hs[o] = dd8 - tmp; ds[o] = do_diss ? dd8 - tmp : 0.; }
}
// ---------------------------------------------------------------------------------------------
// kinetic energy at cell corners (sw_core.F90:1078-1228)
// ---------------------------------------------------------------------------------------------
// RARE = true: the instantiation that also carries hord_mt 1..4, 7, 9, 11 (out-of-line); the common 5 / 6 / 8 / 10 keep their code
template <bool RARE>
__global__ void __launch_bounds__(TI* TJ, PLANE_MINB) k_dsw_ke(Lay L, DevGrid G, const double* __restrict__ u, const double* __restrict__ v,
                                                  const double* __restrict__ uc, const double* __restrict__ vc, const double* __restrict__ uts,
                                                  const double* __restrict__ vts, double* __restrict__ ke, double dt, int hord_mt) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const int npx = L.npx, npy = L.npy;
  const bool cube = L.cube;
  const double dt5 = 0.5 * dt, dt4 = 0.25 * dt;
  auto UT = [&](int ii, int jj) { return AT(uts, ii, jj); };
  auto VT = [&](int ii, int jj) { return AT(vts, ii, jj); };
  double vb, ub;
  if (!cube) {
    vb = dt5 * (AT(vc, i - 1, j) + AT(vc, i, j));
    ub = dt5 * (AT(uc, i, j - 1) + AT(uc, i, j));
  } else {
    if (j == 1 || j == npy) vb = dt5 * (VT(i - 1, j) + VT(i, j));
    else if (i == 1 || i == npx) vb = dt4 * (-VT(i - 2, j) + 3. * (VT(i - 1, j) + VT(i, j)) - VT(i + 1, j));
    else vb = dt5 * (AT(vc, i - 1, j) + AT(vc, i, j) - (AT(uc, i, j - 1) + AT(uc, i, j)) * G2(cosa, i, j)) * G2(rsina, i, j);
    if (i == 1 || i == npx) ub = dt5 * (UT(i, j - 1) + UT(i, j));
    else if (j == 1 || j == npy) ub = dt4 * (-UT(i, j - 2) + 3. * (UT(i, j - 1) + UT(i, j)) - UT(i, j + 1));
    else ub = dt5 * (AT(uc, i, j - 1) + AT(uc, i, j) - (AT(vc, i - 1, j) + AT(vc, i, j)) * G2(cosa, i, j)) * G2(rsina, i, j);
  }
  // ytp_v: advect v along y with vb (:1134)
  double k1;
  {
    double f;
    if (cube && j >= 4 && j <= npy - 3) {
      const long long o = ko + LIDX(L, i, j);
      if constexpr (RARE) f = flux_wind_fast_g(v + o, L.NI, vb, G2(rdy, i, j - 1), G2(rdy, i, j), hord_mt);
      else f = flux_wind_fast(v + o, L.NI, vb, G2(rdy, i, j - 1), G2(rdy, i, j), hord_mt);
    } else {
      Acc va{v + ko, LIDX(L, i, 0), L.NI};
      Acc dya{G.dy, LIDX(L, i, 0), L.NI}, rdy{G.rdy, LIDX(L, i, 0), L.NI};
      f = flux_wind<RARE>(va, dya, rdy, j, vb, hord_mt, npy, cube, cube && (i == 1 || i == npx));
    }
    k1 = vb * f;
  }
  // xtp_u: advect u along x with ub (:1191)
  double k2;
  {
    double f;
    if (cube && i >= 4 && i <= npx - 3) {
      const long long o = ko + LIDX(L, i, j);
      if constexpr (RARE) f = flux_wind_fast_g(u + o, 1, ub, G2(rdx, i - 1, j), G2(rdx, i, j), hord_mt);
      else f = flux_wind_fast(u + o, 1, ub, G2(rdx, i - 1, j), G2(rdx, i, j), hord_mt);
    } else {
      Acc ua{u + ko, LIDX(L, 0, j), 1};
      Acc dxa{G.dx, LIDX(L, 0, j), 1}, rdx{G.rdx, LIDX(L, 0, j), 1};
      f = flux_wind<RARE>(ua, dxa, rdx, i, ub, hord_mt, npx, cube, cube && (j == 1 || j == npy));
    }
    k2 = ub * f;
  }
  double kev = 0.5 * (k1 + k2);
  if (cube) {  // :1203-1228
    const double dt6 = dt / 6.;
    if (i == 1 && j == 1)
      kev = dt6 * ((UT(1, 1) + UT(1, 0)) * AT(u, 1, 1) + (VT(1, 1) + VT(0, 1)) * AT(v, 1, 1) + (UT(1, 1) + VT(1, 1)) * AT(u, 0, 1));
    else if (i == npx && j == 1)
      kev = dt6 * ((UT(i, 1) + UT(i, 0)) * AT(u, i - 1, 1) + (VT(i, 1) + VT(i - 1, 1)) * AT(v, i, 1) + (UT(i, 1) - VT(i - 1, 1)) * AT(u, i, 1));
    else if (i == npx && j == npy)
      kev = dt6 * ((UT(i, j) + UT(i, j - 1)) * AT(u, i - 1, j) + (VT(i, j) + VT(i - 1, j)) * AT(v, i, j - 1) + (UT(i, j - 1) + VT(i - 1, j)) * AT(u, i, j));
    else if (i == 1 && j == npy)
      kev = dt6 * ((UT(1, j) + UT(1, j - 1)) * AT(u, 1, j) + (VT(1, j) + VT(0, j)) * AT(v, 1, j - 1) + (UT(1, j - 1) - VT(1, j)) * AT(u, 0, j));
  }
  ke[ko + LIDX(L, i, j)] = kev;
}

// relative vorticity wk and absolute vorticity wk + f0 on the data domain (:1231-1247, :1476-1496)
__global__ void __launch_bounds__(TI* TJ) k_dsw_vort(Lay L, DevGrid G, const double* __restrict__ u, const double* __restrict__ v,
                                                    double* __restrict__ wk, double* __restrict__ vq) {
  PLANE_IJK
  if (i < L.isd || i > L.ied || j > L.jed) return;
  const long long o = ko + LIDX(L, i, j);
  const double vt0 = AT(u, i, j) * G2(dx, i, j), vt1 = AT(u, i, j + 1) * G2(dx, i, j + 1);
  const double ut0 = AT(v, i, j) * G2(dy, i, j), ut1 = AT(v, i + 1, j) * G2(dy, i + 1, j);
  const double w = G2(rarea, i, j) * (vt0 - vt1 - ut0 + ut1);
  if (wk) wk[o] = w;   // the relative vorticity itself is read only by the Smagorinsky and vorticity-damping branches
  vq[o] = w + G2(f0, i, j);
}

// ---------------------------------------------------------------------------------------------
// divergence damping
// ---------------------------------------------------------------------------------------------
// B-grid scalar with fill_corners(BGRID) views (fv_mp_mod.F90:1031-1062)
struct BFillX {
  const double* q; Lay L; bool on;
  __device__ __forceinline__ double operator()(int ii, int jj) const {
    if (on) {
      const int npx = L.npx, npy = L.npy;
      if (ii < 1 && jj < 1) { const int a = jj, b = 2 - ii; ii = a; jj = b; }
      else if (ii < 1 && jj > npy) { const int a = npy + 1 - jj, b = npy - 1 + ii; ii = a; jj = b; }
      else if (ii > npx && jj < 1) { const int a = npx + 1 - jj, b = ii - npx + 1; ii = a; jj = b; }
      else if (ii > npx && jj > npy) { const int a = npx + jj - npy, b = npy - ii + npx; ii = a; jj = b; }
    }
    return __ldg(q + LIDX(L, ii, jj));
  }
};
struct BFillY {
  const double* q; Lay L; bool on;
  __device__ __forceinline__ double operator()(int ii, int jj) const {
    if (on) {
      const int npx = L.npx, npy = L.npy;
      if (ii < 1 && jj < 1) { const int a = 2 - jj, b = ii; ii = a; jj = b; }
      else if (ii < 1 && jj > npy) { const int a = jj - npy + 1, b = npy + 1 - ii; ii = a; jj = b; }
      else if (ii > npx && jj < 1) { const int a = npx - 1 + jj, b = 1 - ii + npx; ii = a; jj = b; }
      else if (ii > npx && jj > npy) { const int a = npx - jj + npy, b = npy + ii - npx; ii = a; jj = b; }
    }
    return __ldg(q + LIDX(L, ii, jj));
  }
};

// n-th pass of the del-2N divergence damping (sw_core.F90:1387-1424) in ONE kernel:
//   vc = d/dx divg * divg_u, uc = d/dy divg * divg_v   (with fill_corners(divg, BGRID) when nt != 0, :1387-1403)
//   divg = div(uc, vc) * rarea_c                        (with fill_corners(vc, uc, VECTOR, DGRID), :1405-1424)
// uc, vc are evaluated on the fly from divg (5-point footprint, L1-resident) instead of making a round trip through two
// scratch planes: 2 array passes per iteration instead of 6.  dgi -> dgo ping-pong (a thread reads its neighbours).
__global__ void __launch_bounds__(TI* TJ) k_dsw_dd_iter(Lay L, DevGrid G, const double* __restrict__ dgi, double* __restrict__ dgo,
                                                       const int* kint, int n, int stretched) {
  PLANE_IJK
  const int nord = kint[KI_NORD * (L.npz + 1) + k];
  if (n > nord) return;
  const int nt = nord - n;
  if (i < L.is - nt || i > L.ie + 1 + nt || j < L.js - nt || j > L.je + 1 + nt) return;
  const int npx = L.npx, npy = L.npy;
  if (i >= 4 && i <= npx - 4 && j >= 4 && j <= npy - 4) {
    // away from the face edges (96 % of the points at C384): the 5-point cross of divg with plain offsets from one index.  The
    // general path below carries the fill_corners remap tests and the face-corner terms through every access and was
    // instruction bound (262 warp instructions per thread at 75 % issue-active, profiles/r2_dsw_summary.md).  Same operations.
    const int o2 = LIDX(L, i, j), NI = L.NI;
    const double* d = dgi + ko + o2;
    const double dC = __ldg(d), dW = __ldg(d - 1), dE = __ldg(d + 1), dS = __ldg(d - NI), dN = __ldg(d + NI);
    const double ucS = (dC - dS) * __ldg(G.divg_v + o2 - NI), ucC = (dN - dC) * __ldg(G.divg_v + o2);
    const double vcW = (dC - dW) * __ldg(G.divg_u + o2 - 1), vcC = (dE - dC) * __ldg(G.divg_u + o2);
    double dv = ucS - ucC + vcW - vcC;
    if (!stretched) dv = dv * __ldg(G.rarea_c + o2);
    dgo[ko + o2] = dv;
    return;
  }
  const bool fill_c = (nt != 0) && L.cube && (i < 4 || i > npx - 4) && (j < 4 || j > npy - 4);   // remaps exist in the corner regions only
  const double* d = dgi + ko;
  BFillX dx_{d, L, fill_c};
  BFillY dy_{d, L, fill_c};
  auto vcs = [&](int ii, int jj) { return (dx_(ii + 1, jj) - dx_(ii, jj)) * G2(divg_u, ii, jj); };   // "vc"
  auto ucs = [&](int ii, int jj) { return (dy_(ii, jj + 1) - dy_(ii, jj)) * G2(divg_v, ii, jj); };   // "uc"
  const double s = -1.0;
  // x = vc (u-like), y = uc (v-like): fv_mp_mod.F90:1262-1277
  auto VCv = [&](int ii, int jj) -> double {
    if (fill_c) {
      if (ii <= 0 && jj <= 0) return s * ucs(jj, 1 - ii);
      if (ii <= 0 && jj >= npy + 1) return ucs(npy + 1 - jj, npy - 1 + ii);
      if (ii >= npx && jj <= 0) return ucs(npx + 1 - jj, ii - npx + 1);
      if (ii >= npx && jj >= npy + 1) return s * ucs(npx + jj - npy, npy - ii + npx - 1);
    }
    return vcs(ii, jj);
  };
  auto UCv = [&](int ii, int jj) -> double {
    if (fill_c) {
      if (ii <= 0 && jj <= 0) return s * vcs(1 - jj, ii);
      if (ii <= 0 && jj >= npy) return vcs(jj - npy + 1, npy + 1 - ii);
      if (ii >= npx + 1 && jj <= 0) return vcs(npx - 1 + jj, 1 - ii + npx);
      if (ii >= npx + 1 && jj >= npy) return s * vcs(npx - jj + npy - 1, npy + ii - npx);
    }
    return ucs(ii, jj);
  };
  double dv = UCv(i, j - 1) - UCv(i, j) + VCv(i - 1, j) - VCv(i, j);
  if (L.cube) {
    if (i == 1 && j == 1) dv = dv - UCv(1, 0);
    if (i == npx && j == 1) dv = dv - UCv(npx, 0);
    if (i == npx && j == npy) dv = dv + UCv(npx, npy);
    if (i == 1 && j == npy) dv = dv + UCv(1, npy);
  }
  if (!stretched) dv = dv * G2(rarea_c, i, j);
  dgo[ko + LIDX(L, i, j)] = dv;
}

// final damping term (both branches) and ke += term.  dterm = the reference's B-grid "vort".
// nord==0 levels: del-2 from u, v, ua, va (:1290-1371); nord>0: :1376-1458
__global__ void __launch_bounds__(TI* TJ) k_dsw_damp(Lay L, DevGrid G, const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ ua, const double* __restrict__ va, const double* __restrict__ uc,
                                                    const double* __restrict__ vc, const double* __restrict__ divg_in,
                                                    const double* __restrict__ dg_even, const double* __restrict__ dg_odd,
                                                    const double* __restrict__ vortb, double* __restrict__ ke,
                                                    double* __restrict__ dterm, const int* kint, const double* kdbl, double dt, double dddmp,
                                                    double d4_bg, int stretched, double* __restrict__ delpc_out) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const int npx = L.npx, npy = L.npy;
  const bool cube = L.cube;
  const int nord = kint[KI_NORD * (L.npz + 1) + k];
  const double d2_bg = kdbl[KD_D2BG * (L.npz + 1) + k];
  const long long o = ko + LIDX(L, i, j);
  double term;
  if (nord == 0) {
    auto PTC = [&](int ii, int jj) -> double {
      if (cube && (jj == 1 || jj == npy)) {
        return (AT(vc, ii, jj) > 0) ? AT(u, ii, jj) * G2(dyc, ii, jj) * SG(4, ii, jj - 1) : AT(u, ii, jj) * G2(dyc, ii, jj) * SG(2, ii, jj);
      }
      return (AT(u, ii, jj) - 0.5 * (AT(va, ii, jj - 1) + AT(va, ii, jj)) * G2(cosa_v, ii, jj)) * G2(dyc, ii, jj) * G2(sina_v, ii, jj);
    };
    auto VRT = [&](int ii, int jj) -> double {
      if (cube && (ii == 1 || ii == npx)) {
        return (AT(uc, ii, jj) > 0) ? AT(v, ii, jj) * G2(dxc, ii, jj) * SG(3, ii - 1, jj) : AT(v, ii, jj) * G2(dxc, ii, jj) * SG(1, ii, jj);
      }
      return (AT(v, ii, jj) - 0.5 * (AT(ua, ii - 1, jj) + AT(ua, ii, jj)) * G2(cosa_u, ii, jj)) * G2(dxc, ii, jj) * G2(sina_u, ii, jj);
    };
    double dpc = VRT(i, j - 1) - VRT(i, j) + PTC(i - 1, j) - PTC(i, j);
    if (cube) {
      if (i == 1 && j == 1) dpc = dpc - VRT(1, 0);
      if (i == npx && j == 1) dpc = dpc - VRT(npx, 0);
      if (i == npx && j == npy) dpc = dpc + VRT(npx, npy);
      if (i == 1 && j == npy) dpc = dpc + VRT(1, npy);
    }
    dpc = G2(rarea_c, i, j) * dpc;
    const double damp = G.da_min_c * fmax(d2_bg, fmin(0.20, dddmp * fabs(dpc * dt)));
    term = damp * dpc;
    if (delpc_out) delpc_out[o] = dpc;   // d_sw's delpc output (sw_core.F90:1366), read by the external-mode damping (d_ext > 0)
  } else {
    const double dpc = __ldg(divg_in + o);   // delpc = divg_d saved before the loop (:1376-1381)
    if (delpc_out) delpc_out[o] = dpc;
    double vo = 0.;
    if (dddmp >= 1.E-5) {
      const double vb = __ldg(vortb + o);
      vo = fabs(dt) * sqrt(dpc * dpc + vb * vb);
    }
    const double dd8 = kdbl[KD_DD8 * (L.npz + 1) + k];   // (da_min_c*d4_bg)^(nord+1), tabulated per level by the host (a pow per thread cost 0.1 ms)
    const double damp2 = G.da_min_c * fmax(d2_bg, fmin(0.20, dddmp * vo));
    term = damp2 * dpc + dd8 * __ldg(((nord & 1) ? dg_odd : dg_even) + o);   // result plane of the last ping-pong pass
  }
  if (dterm) dterm[o] = term;
  ke[o] = ke[o] + term;
}

// dissipative heating / dissipation estimate (:1462-1473, :1523-1586) and the vorticity-damping
// momentum increments (:1589-1600).  ut_d, vt_d = del6_vt_flux outputs (x-flux "ut", y-flux "vt").
__global__ void __launch_bounds__(TI* TJ) k_dsw_heat(Lay L, DevGrid G, const double* __restrict__ un, const double* __restrict__ vn,
                                                    const double* __restrict__ uold, const double* __restrict__ vold,
                                                    const double* __restrict__ dterm, const double* __restrict__ ut_d, const double* __restrict__ vt_d,
                                                    const double* __restrict__ delp_new, const double* __restrict__ hs, const double* __restrict__ ds,
                                                    double* __restrict__ heat, double* __restrict__ diss, const double* kdbl, int prevent,
                                                    int do_diss, double d_con_flag) {
  PLANE_IJK
  if (i < L.is || i > L.ie || j < L.js || j > L.je) return;
  const double d_con = kdbl[KD_DCON * (L.npz + 1) + k];
  const bool have_v = kdbl[KD_DAMP4_V * (L.npz + 1) + k] != 0.;
  const long long o = ko + LIDX(L, i, j);
  double hsv = __ldg(hs + o), dsv = __ldg(ds + o);
  if (d_con > 1.e-5 || do_diss) {
    auto UB = [&](int ii, int jj) {  // (ub + vt)*rdx on (is:ie, js:je+1)
      const double ub = AT(dterm, ii, jj) - AT(dterm, ii + 1, jj);
      // without vorticity damping and without do_diss_est the reference's vt still holds u*dx (sw_core.F90:1233,1516)
      const double vt = have_v ? AT(vt_d, ii, jj) : (do_diss ? 0. : AT(uold, ii, jj) * G2(dx, ii, jj));
      return (ub + vt) * G2(rdx, ii, jj);
    };
    auto VB = [&](int ii, int jj) {  // (vb - ut)*rdy on (is:ie+1, js:je)
      const double vb = AT(dterm, ii, jj) - AT(dterm, ii, jj + 1);
      const double ut = have_v ? AT(ut_d, ii, jj) : (do_diss ? 0. : AT(vold, ii, jj) * G2(dy, ii, jj));
      return (vb - ut) * G2(rdy, ii, jj);
    };
    const double ub0 = UB(i, j), ub1 = UB(i, j + 1), vb0 = VB(i, j), vb1 = VB(i + 1, j);
    const double fy0 = AT(un, i, j) * G2(rdx, i, j), fy1 = AT(un, i, j + 1) * G2(rdx, i, j + 1);
    const double fx0 = AT(vn, i, j) * G2(rdy, i, j), fx1 = AT(vn, i + 1, j) * G2(rdy, i + 1, j);
    const double gy0 = fy0 * ub0, gy1 = fy1 * ub1, gx0 = fx0 * vb0, gx1 = fx1 * vb1;
    const double u2 = fy0 + fy1, du2 = ub0 + ub1, v2 = fx0 + fx1, dv2 = vb0 + vb1;
    const double inner = ((ub0 * ub0 + ub1 * ub1 + vb0 * vb0 + vb1 * vb1) + 2. * (gy0 + gy1 + gx0 + gx1) -
                          G2(cosa_s, i, j) * (u2 * dv2 + v2 * du2 + du2 * dv2));
    const double damp = 0.25 * d_con;
    if (prevent) {
      const double tmp = G2(rsin2, i, j) * inner;
      if (d_con > 1.e-5) hsv = __ldg(delp_new + o) * (hsv - damp * fmin(0., tmp));
      if (do_diss) dsv = dsv - tmp;
    } else {
      hsv = __ldg(delp_new + o) * (hsv - damp * G2(rsin2, i, j) * inner);
      if (do_diss) dsv = dsv - G2(rsin2, i, j) * inner;
    }
  }
  // dyn_core.F90:798-811 accumulation into the 3-D arrays
  if (d_con_flag > 1.0E-5) heat[o] = heat[o] + hsv;
  if (do_diss) diss[o] = diss[o] + dsv;
}
__global__ void __launch_bounds__(TI* TJ) k_dsw_vdamp(Lay L, double* __restrict__ un, double* __restrict__ vn, const double* __restrict__ ut_d,
                                                     const double* __restrict__ vt_d, const double* kdbl) {
  PLANE_IJK
  if (kdbl[KD_DAMP4_V * (L.npz + 1) + k] == 0.) return;
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const long long o = ko + LIDX(L, i, j);
  if (i <= L.ie) un[o] = un[o] + vt_d[o];
  if (j <= L.je) vn[o] = vn[o] - ut_d[o];
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static void dsw_tables(fv3_ctx* c, std::vector<int>& ki, std::vector<double>& kd) {
  // model/dyn_core.F90:666-733
  const fv3_flags_t& f = c->f;
  const int npz = c->L.npz, n1 = npz + 1;
  ki.assign(FV3_KSLOTS * n1, 0); kd.assign(FV3_KSLOTS * n1, 0.);
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
        d2_divg = f.is_ideal_case ? std::max(f.d2_bg, f.d2_bg_k1) : std::max(0.01, std::max(f.d2_bg, f.d2_bg_k1));
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
    const int kk = k - 1;
    const int nord_v = c->nord_v[kk]; const double damp_v = c->damp_vt[kk];
    ki[KI_NORD * n1 + kk] = nord_k; ki[KI_NORD_V * n1 + kk] = nord_v; ki[KI_NORD_W * n1 + kk] = nord_w; ki[KI_NORD_T * n1 + kk] = nord_t;
    kd[KD_D2BG * n1 + kk] = d2_divg; kd[KD_DAMP_V * n1 + kk] = damp_v; kd[KD_DAMP_W * n1 + kk] = damp_w; kd[KD_DAMP_T * n1 + kk] = damp_t;
    kd[KD_DCON * n1 + kk] = d_con_k;
    kd[KD_DAMP4_W * n1 + kk] = (!f.hydrostatic && damp_w > 1.E-5) ? pow(damp_w * c->G.da_min_c, (double)(nord_w + 1)) : 0.;   // sw_core.F90:951-953
    kd[KD_DAMP4_V * n1 + kk] = (damp_v > 1.E-5) ? pow(damp_v * c->G.da_min_c, (double)(nord_v + 1)) : 0.;   // :1513-1514
    kd[KD_DELN * n1 + kk] = (damp_v > 1.e-4) ? pow(damp_v * c->G.da_min, (double)(nord_v + 1)) : 0.;       // tp_core.F90:202-203, delp (nord_v, damp_v)
    kd[KD_DELN_T * n1 + kk] = (damp_t > 1.e-4) ? pow(damp_t * c->G.da_min, (double)(nord_t + 1)) : 0.;     // pt, q_con (nord_t, damp_t)
    kd[KD_DD8 * n1 + kk] = c->b.stretched_grid ? c->G.da_min * pow(f.d4_bg, (double)(nord_k + 1))
                                              : pow(c->G.da_min_c * f.d4_bg, (double)(nord_k + 1));            // sw_core.F90:1427-1431
  }
}

// ---------------------------------------------------------------------------------------------
// fused transport of delp, w, q_con, pt (sw_core.F90:919-1066, :1262-1283): ONE kernel per d_sw call.
// A CTA owns a 26x26 tile of one level (tp_tile.cuh).  The Courant numbers / area fluxes are staged once and
// shared by the fields; the mass fluxes of delp stay in shared memory and weight the fluxes of the other fields;
// the flux divergences are applied in the epilogue (thread = one cell), so fx, fy, gx, gy never touch HBM.
// The del-n damping fluxes (wide stencil, separate kernels) are read from global and added on the fly.
// ---------------------------------------------------------------------------------------------
struct DswTr {
  const double *delp, *pt, *w, *qcon;                  // inputs (qcon/w nullable)
  const double *crx, *cry, *xfx, *yfx;
  const double *dpx, *dpy;                             // del-n flux of delp  (nullable)
  const double *ptx, *pty;                             // del-n flux of pt    (nullable)
  const double *qcx, *qcy;                             // del-n flux of q_con (nullable)
  const double* dw;                                    // w damping increment (nullable)
  double *delp_o, *pt_o, *w_o, *qcon_o;
  double *mfx, *mfy;                                   // flux capacitors (accumulated)
  const double* kdbl;
  int hord_dp, hord_vt, hord_tm;
};
struct DswSmem {
  tpt::Smem t;
  double mfx[tpt::QH][tpt::QW];   // mass fluxes incl. del-n damping, tile coordinates (west / south face of [r][c])
  double mfy[tpt::QH][tpt::QW];
};
constexpr int CELL_ITERS = (tpt::TY + tpt::NW - 1) / tpt::NW;

// weight the unweighted fluxes of the current field by the mass fluxes (+ mass-weighted del-n flux, tp_core.F90:1390-1445)
__device__ __forceinline__ void weight_by_mass(const Lay& L, DswSmem& S, const tpt::Tile& T, const double* __restrict__ delp,
                                               const double* __restrict__ dx_, const double* __restrict__ dy_, double coef) {
  using namespace tpt;
  const int c = T.lane, i = T.i0 - 3 + c;
#pragma unroll
  for (int r = 3 + T.wid; r <= TY + 3; r += NW) {
    const int j = T.j0 - 3 + r;
    if (c < 3 || j > L.je + 1) continue;
    const int o = T.idx(i, j);
    if (r < TY + 3 && j <= L.je && c <= TX + 3 && i <= L.ie + 1) {
      double g = S.t.fx2[r][c] * S.mfx[r][c];
      if (dx_) g = g + (0.5 * coef) * (__ldg(delp + T.ko + o - 1) + __ldg(delp + T.ko + o)) * __ldg(dx_ + T.ko + o);
      S.t.fx2[r][c] = g;
    }
    if (c <= TX + 2 && i <= L.ie) {
      double g = S.t.fy2[r][c] * S.mfy[r][c];
      if (dy_) g = g + (0.5 * coef) * (__ldg(delp + T.ko + o - T.NI) + __ldg(delp + T.ko + o)) * __ldg(dy_ + T.ko + o);
      S.t.fy2[r][c] = g;
    }
  }
  __syncthreads();
}

// FAM: 0 all fields use an unlimited scheme, 1 all monotone, 2 mixed (per-field choice at run time).
// The fields are a run-time loop around ONE copy of the tile routine (instruction-cache footprint).
template <int FAM, bool EDGE>
__global__ void __launch_bounds__(tpt::NT, 2) k_dsw_transport(Lay L, DevGrid G, tpt::TileMap M, DswTr a) {
  using namespace tpt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DswSmem& S = *reinterpret_cast<DswSmem*>(smem_raw);
  const Tile T = make_tile(L, M);
  const int k = blockIdx.z, n1 = L.npz + 1;
  const int c = T.lane, i = T.i0 - 3 + c;   // tile column / global i of this lane
  const bool xcell = c >= 3 && c <= TX + 2 && i <= L.ie;
  const bool lastx = T.i0 + TX > L.ie, lasty = T.j0 + TY > L.je;
  const double c_dp = a.kdbl[KD_DELN * n1 + k], c_t = a.kdbl[KD_DELN_T * n1 + k];
  const bool dwk = a.dw && a.kdbl[KD_DAMP4_W * n1 + k] != 0.;
  stage_inputs<EDGE>(L, G, S.t, T, a.crx, a.cry, a.xfx, a.yfx);
  double dp[CELL_ITERS], dpn[CELL_ITERS], ra[CELL_ITERS];
#pragma unroll 1
  for (int f = 0; f < 4; f++) {   // 0 delp (:919-920), 1 w (:984-990), 2 q_con (:992-1000), 3 pt (:1014-1016)
    const double* qf = f == 0 ? a.delp : f == 1 ? a.w : f == 2 ? a.qcon : a.pt;
    if (!qf) continue;
    const int hord = (f == 1) ? a.hord_vt : (f == 3) ? a.hord_tm : a.hord_dp;
    stage_q<EDGE>(L, S.t, T, qf);
    tp_compute<FAM, EDGE>(L, G, S.t, T, nullptr, nullptr, (hord == 10) ? 8 : hord, hord);
    if (f == 0) {
      // mass fluxes (+ del-n damping flux) into shared memory and the flux capacitors (:928-940)
#pragma unroll
      for (int n = 0; n < CELL_ITERS; n++) {   // this thread's cells: rows 3 + wid + n*NW
        const int r = 3 + T.wid + n * NW, j = T.j0 - 3 + r;
        const bool ok = xcell && r < TY + 3 && j <= L.je;
        dp[n] = ok ? S.t.q[r][c] : 1.;
        ra[n] = ok ? __ldg(G.rarea + T.idx(i, j)) : 0.;
        dpn[n] = 1.;
      }
      const bool damp = a.dpx && c_dp != 0.;
#pragma unroll 1
      for (int r = 3 + T.wid; r <= TY + 3; r += NW) {
        const int j = T.j0 - 3 + r;
        if (c < 3 || j > L.je + 1) continue;
        const int o = T.idx(i, j);
        if (r < TY + 3 && j <= L.je && c <= TX + 3 && i <= L.ie + 1) {
          double m = S.t.fx2[r][c] * S.t.xfx[r][c];
          if (damp) m = m + __ldg(a.dpx + T.ko + o);
          S.mfx[r][c] = m;
          if (c < TX + 3 || lastx) a.mfx[T.ko + o] = a.mfx[T.ko + o] + m;
        }
        if (c <= TX + 2 && i <= L.ie) {
          double m = S.t.fy2[r][c] * S.t.yfx[r][c];
          if (damp) m = m + __ldg(a.dpy + T.ko + o);
          S.mfy[r][c] = m;
          if (r < TY + 3 || lasty) a.mfy[T.ko + o] = a.mfy[T.ko + o] + m;
        }
      }
      __syncthreads();   // mass fluxes visible; every thread has read its delp and the unweighted fluxes
#pragma unroll
      for (int n = 0; n < CELL_ITERS; n++) {
        const int r = 3 + T.wid + n * NW;
        if (xcell && r < TY + 3) dpn[n] = dp[n] + (S.mfx[r][c] - S.mfx[r][c + 1] + S.mfy[r][c] - S.mfy[r + 1][c]) * ra[n];
        // delp is the only transported field (SW_DYNAMICS test_case 1, sw_core.F90:1055-1066): no pt epilogue will store it
        if (!a.pt && xcell && r < TY + 3 && T.j0 - 3 + r <= L.je) a.delp_o[T.ko + T.idx(i, T.j0 - 3 + r)] = dpn[n];
      }
      continue;
    }
    // mass-weighted fluxes (+ mass-weighted del-n flux of q_con / pt), then (q*delp + div) / delp_new  (:1053-1066, :1262-1283)
    const bool damp = (f == 2) ? (a.qcx && c_t != 0.) : (f == 3) ? (a.ptx && c_t != 0.) : false;
    weight_by_mass(L, S, T, a.delp, damp ? (f == 2 ? a.qcx : a.ptx) : nullptr, damp ? (f == 2 ? a.qcy : a.pty) : nullptr, c_t);
    double* out = (f == 1) ? a.w_o : (f == 2) ? a.qcon_o : a.pt_o;
#pragma unroll
    for (int n = 0; n < CELL_ITERS; n++) {
      const int r = 3 + T.wid + n * NW, j = T.j0 - 3 + r;
      if (!(xcell && r < TY + 3 && j <= L.je)) continue;
      const long long o = T.ko + T.idx(i, j);
      const double div = (S.t.fx2[r][c] - S.t.fx2[r][c + 1] + S.t.fy2[r][c] - S.t.fy2[r + 1][c]) * ra[n];
      // same operand order as the reference: w, q_con: delp*q + div (:986,:998); pt: pt*delp + div (:1061)
      const double num = (f == 3) ? S.t.q[r][c] * dp[n] + div : dp[n] * S.t.q[r][c] + div;
      double v = num / dpn[n];
      if (f == 1 && dwk) v = v + __ldg(a.dw + o);
      out[o] = v;
      if (f == 3) a.delp_o[o] = dpn[n];
    }
    __syncthreads();   // epilogue reads done before the next field overwrites q / fluxes
  }
}

// vorticity transport + momentum update (sw_core.F90:1476-1509): u += ke(i)-ke(i+1) + fy, v += ke(j)-ke(j+1) - fx
template <int FAM, bool EDGE>
__global__ void __launch_bounds__(tpt::NT, 2) k_dsw_vort_uv(Lay L, DevGrid G, tpt::TileMap M, const double* __restrict__ vq, const double* __restrict__ crx,
                                                          const double* __restrict__ cry, const double* __restrict__ xfx,
                                                          const double* __restrict__ yfx, const double* __restrict__ u,
                                                          const double* __restrict__ v, const double* __restrict__ ke,
                                                          double* __restrict__ uo, double* __restrict__ vo, int hord_vt) {
  using namespace tpt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const Tile T = make_tile(L, M);
  stage_inputs<EDGE>(L, G, S, T, crx, cry, xfx, yfx);
  stage_q<EDGE>(L, S, T, vq);
  tp_compute<FAM, EDGE>(L, G, S, T, nullptr, nullptr, (hord_vt == 10) ? 8 : hord_vt, hord_vt);
  const bool lastx = T.i0 + TX > L.ie, lasty = T.j0 + TY > L.je;
  u += T.ko; v += T.ko; ke += T.ko; uo += T.ko; vo += T.ko;
  const int c = T.lane - 3, i = T.i0 + c;
#pragma unroll
  for (int r = T.wid; r <= TY; r += NW) {
    const int j = T.j0 + r;
    if (c < 0 || j > L.je + 1) continue;
    const int o = T.idx(i, j);
    if (r < TY && j <= L.je && i <= L.ie + 1 && (c < TX || (c == TX && lastx))) {   // v (is:ie+1, js:je)
      const double fx = FX(S, r, c) * S.xfx[r + 3][c + 3];
      vo[o] = __ldg(v + o) * __ldg(G.dy + o) + __ldg(ke + o) - __ldg(ke + o + T.NI) - fx;
    }
    if (c < TX && i <= L.ie && (r < TY || lasty)) {                                   // u (is:ie, js:je+1)
      const double fy = FY(S, r, c) * S.yfx[r + 3][c + 3];
      uo[o] = __ldg(u + o) * __ldg(G.dx + o) + __ldg(ke + o) - __ldg(ke + o + 1) + fy;
    }
  }
}

// ---- line-per-warp form (tp_line.cuh): delp, [w,] pt transported phase-major by a CTA that is persistent over a chunk of levels.
// Same arithmetic as k_dsw_transport (the mass fluxes of delp weight the other fields' fluxes in the outer sweep, the flux
// divergences are applied in the epilogue); launched when no del-n flux and no q_con is in play.
// NF = 4: the absolute vorticity rides along as a fourth field (area-flux weighted, sw_core.F90:1476-1509) and the epilogue also
// does the momentum update of k_dsw_vort_uv2 -- all scalar transports of d_sw share the Courant numbers and area fluxes, so the
// one-field kernel's staging (5 of its 6 input arrays), barriers and per-level bookkeeping are saved.
struct DswVU { const double *vq, *u, *v, *ke; double *uo, *vo; };
template <int FAM, int NF, int HORD, bool EDGE, typename R = double>
__global__ void __launch_bounds__(tp2::NT, 1) k_dsw_transport2(Lay L, DevGrid G, tpt::TileMap M, DswTr a, DswVU vu, int nk, int kch) {
  static_assert(NF >= 2 && NF <= 4, "fields: delp, [w,] pt [, vorticity]");
  constexpr int NS = NF == 4 ? 3 : NF;                       // mass-weighted scalars
  constexpr int NEP = (NF == 4 && EDGE) ? 0 : 2;             // four fields + the cube-edge tables leave no room for the prefetch arrays
  constexpr int WM = NF == 4 ? tp2::W_MASS_V : tp2::W_MASS;
  const double* src[4 + NF];
  src[0] = a.crx; src[1] = a.cry; src[2] = a.xfx; src[3] = a.yfx; src[4] = a.delp;
  if (NS == 3) src[5] = a.w;
  src[3 + NS] = a.pt;
  if (NF == 4) src[7] = vu.vq;
  int ord_in[NF], ord_ou[NF];
  ord_ou[0] = a.hord_dp; if (NS == 3) ord_ou[1] = a.hord_vt; ord_ou[NS - 1] = a.hord_tm;
  if (NF == 4) ord_ou[3] = a.hord_vt;
#pragma unroll
  for (int f = 0; f < NF; f++) ord_in[f] = (ord_ou[f] == 10) ? 8 : ord_ou[f];   // tp_core.F90:136-141
  const int n1 = L.npz + 1;
  // this thread's epilogue cell is the same on every level: its 1/area stays in a register
  double ra = 0.;
  bool lastx, lasty;
  {
    const tp2::Geo T = tp2::make_geo(L, M);
    const int i = T.i0 - 3 + T.lane, j = T.j0 + T.wid;
    if (T.wid < tp2::TY && (!EDGE || (i <= L.ied && j <= L.jed))) ra = __ldg(G.rarea + tp2::gidx(T, i, j));
    lastx = T.i0 + tp2::TX > L.ie; lasty = T.j0 + tp2::TY > L.je;
  }
  tp2::run_tile<FAM, NF, NEP, WM, HORD, 32, EDGE, R>(L, G, M, src, nk, kch, ord_in, ord_ou,
    [&](tp2::Smem<NF, NEP, EDGE, R>& S, const tp2::Geo& T, int k, long long ko, int r) {   // the flux capacitors' old values (read-modify-write, :928-940)
      if constexpr (NEP == 2) {
        const int c = T.lane, i = T.i0 - 3 + c, j = T.j0 - 3 + r;
        if (c < 3 || c > tp2::TX + 3 || (EDGE && (i > L.ie + 1 || j > L.je + 1))) return;
        const long long g = ko + tp2::gidx(T, i, j);
        tpt::cp_async8(&S.ep[0][r * tp2::P + c], a.mfx + g);
        tpt::cp_async8(&S.ep[1][r * tp2::P + c], a.mfy + g);
      }
    },
    [&](tp2::Smem<NF, NEP, EDGE, R>& S, int b, const tp2::Geo& T, int k, long long ko, int r) {
      const int c = T.lane, i = T.i0 - 3 + c, j = T.j0 - 3 + r;
      if (c < 3 || c > tp2::TX + 3 || (EDGE && (i > L.ie + 1 || j > L.je + 1))) return;
      const int o = r * tp2::P + c;
      const int gi = tp2::gidx(T, i, j);
      const long long g = ko + gi;
      const double mx0 = (double)S.qi[0][o], my0 = (double)S.qj[0][o];
      // flux capacitors: the tile's own west / south faces, plus the face's last column / row of faces
      const bool xface = r <= tp2::TY + 2 && (!EDGE || j <= L.je) && (c <= tp2::TX + 2 || (EDGE && lastx));
      const bool yface = c <= tp2::TX + 2 && (!EDGE || i <= L.ie) && (r <= tp2::TY + 2 || (EDGE && lasty));
      if (xface) a.mfx[g] = (NEP == 2 ? S.ep[0][NEP == 2 ? o : 0] : a.mfx[g]) + mx0;
      if (yface) a.mfy[g] = (NEP == 2 ? S.ep[NEP == 2 ? 1 : 0][NEP == 2 ? o : 0] : a.mfy[g]) + my0;
      if constexpr (NF == 4) {   // momentum update (k_dsw_vort_uv2): v on the west faces, u on the south faces
        const double kev = (xface || yface) ? __ldg(vu.ke + g) : 0.;
        if (xface) vu.vo[g] = __ldg(vu.v + g) * __ldg(G.dy + gi) + kev - __ldg(vu.ke + g + T.NI) - (double)S.qi[3][o];
        if (yface) vu.uo[g] = __ldg(vu.u + g) * __ldg(G.dx + gi) + kev - __ldg(vu.ke + g + 1) + (double)S.qj[3][o];
      }
      if (r > tp2::TY + 2 || c > tp2::TX + 2 || (EDGE && (i > L.ie || j > L.je))) return;   // not a cell of the tile
      const double mx1 = (double)S.qi[0][o + 1], my1 = (double)S.qj[0][o + tp2::P];
      const double dp = S.q64(b, 0, o);
      const double dpn = dp + (mx0 - mx1 + my0 - my1) * ra;
      const double rdpn = 1. / dpn;   // one division for w and pt (<= 1 ulp from the two divisions of sw_core.F90:986, 1061)
#pragma unroll
      for (int f = 1; f < NS; f++) {
        const double div = ((double)S.qi[f][o] - (double)S.qi[f][o + 1] + (double)S.qj[f][o] - (double)S.qj[f][o + tp2::P]) * ra;
        const double q = S.q64(b, f, o);
        double v = (q * dp + div) * rdpn;
        if (NS == 3 && f == 1) {
          if (a.dw && a.kdbl[KD_DAMP4_W * n1 + k] != 0.) v = v + __ldg(a.dw + g);
          a.w_o[g] = v;
        } else a.pt_o[g] = v;
      }
      a.delp_o[g] = dpn;
    });
}

// one transported field; TP2_NF1_NWC warps per CTA on the interior tiles (16: two CTAs per SM), 32 on the frame tiles (their
// cube-edge tables do not fit twice)
#ifndef TP2_NF1_NWC
#define TP2_NF1_NWC 32   // measured at C384L79: 401 us (one CTA of 32 warps per SM) vs 420 us (two CTAs of 16)
#endif
template <int FAM, int HORD, bool EDGE, typename R = double>
__global__ void __launch_bounds__(EDGE ? 1024 : TP2_NF1_NWC * 32, EDGE ? 1 : 32 / TP2_NF1_NWC) k_dsw_vort_uv2(Lay L, DevGrid G, tpt::TileMap M, const double* __restrict__ vq, const double* __restrict__ crx,
                                                       const double* __restrict__ cry, const double* __restrict__ xfx,
                                                       const double* __restrict__ yfx, const double* __restrict__ u,
                                                       const double* __restrict__ v, const double* __restrict__ ke,
                                                       double* __restrict__ uo, double* __restrict__ vo, int hord_vt, int nk, int kch) {
  const double* src[5] = {crx, cry, xfx, yfx, vq};
  const int ord_ou[1] = {hord_vt}, ord_in[1] = {(hord_vt == 10) ? 8 : hord_vt};
  tp2::run_tile<FAM, 1, 0, tp2::W_AREA, HORD, EDGE ? 32 : TP2_NF1_NWC, EDGE, R>(L, G, M, src, nk, kch, ord_in, ord_ou,
    [&](tp2::Smem<1, 0, EDGE, R>&, const tp2::Geo&, int, long long, int) {},
    [&](tp2::Smem<1, 0, EDGE, R>& S, int b, const tp2::Geo& T, int k, long long ko, int r) {
      const int c = T.lane, i = T.i0 - 3 + c, j = T.j0 - 3 + r;
      if (c < 3 || c > tp2::TX + 3 || (EDGE && (i > L.ie + 1 || j > L.je + 1))) return;
      const bool lastx = T.i0 + tp2::TX > L.ie, lasty = T.j0 + tp2::TY > L.je;
      const int o = r * tp2::P + c;
      const int gi = tp2::gidx(T, i, j);
      const long long g = ko + gi;
      const double kev = __ldg(ke + g);
      // v on (is:ie+1, js:je): west faces; u on (is:ie, js:je+1): south faces
      if (r <= tp2::TY + 2 && (!EDGE || j <= L.je) && (c <= tp2::TX + 2 || (EDGE && lastx)))
        vo[g] = __ldg(v + g) * __ldg(G.dy + gi) + kev - __ldg(ke + g + T.NI) - (double)S.qi[0][o];
      if (c <= tp2::TX + 2 && (!EDGE || i <= L.ie) && (r <= tp2::TY + 2 || (EDGE && lasty)))
        uo[g] = __ldg(u + g) * __ldg(G.dx + gi) + kev - __ldg(ke + g + 1) + (double)S.qj[0][o];
    });
}

// the fused kernels write the computational domain only: copy the halo frame so a frozen halo stays frozen
struct FrameJob { const double* src; double* dst; int i1, j1; };   // computed box is (is:i1, js:j1), array box (isd:i1+ng', ...)
struct FrameJobs { FrameJob j[6]; int n; };
__global__ void __launch_bounds__(TI* TJ) k_copy_frame(Lay L, FrameGrid FG, FrameJobs jobs) {
  int bx_, by_;
  FG.map(blockIdx.x, bx_, by_);
  const int i = L.isd - FV3_IOFF + bx_ * TI + threadIdx.x;
  const int j = L.jsd + by_ * TJ + threadIdx.y;
  const long long ko = (long long)blockIdx.z * L.plane;
  if (i < L.isd || i > L.ied + 1 || j > L.jed + 1) return;
  const long long o = ko + LIDX(L, i, j);
  for (int n = 0; n < jobs.n; n++) {
    const FrameJob& f = jobs.j[n];
    const int ihi = f.i1 + (L.ied - L.ie), jhi = f.j1 + (L.jed - L.je);   // array extents incl. halo
    if (i > ihi || j > jhi) continue;
    if (i >= L.is && i <= f.i1 && j >= L.js && j <= f.j1) continue;
    f.dst[o] = __ldg(f.src + o);
  }
}

// fuse_vort() : the vorticity transport + momentum update ride in the scalar transport kernel (k_dsw_transport2<.., NF = 4>)
static int fuse_vort() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FV3_TP_FUSE4"); v = e ? atoi(e) : 1; }
  return v;
}
static bool can_fuse_vort(const fv3_ctx* c, const DswTr& a) {
  const bool mono = a.hord_dp >= 8 && a.hord_tm >= 8 && a.hord_vt >= 8, none = a.hord_dp < 8 && a.hord_tm < 8 && a.hord_vt < 8;
  const bool rare = hord_is_rare(a.hord_dp) || hord_is_rare(a.hord_tm) || hord_is_rare(a.hord_vt);
  return fuse_vort() && use_line_kernels() > 1 && !c->tp_fp32 && a.w && a.pt && !a.qcon && !a.dpx && !a.ptx && !a.qcx && !rare && (mono || none);
}
template <int FAM>
static int launch_transport_t(fv3_ctx* c, const DswTr& a, int nk, const DswVU* vu = nullptr) {
  static bool attr_set = false;
  if (!attr_set) {
    FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_transport<FAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DswSmem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_transport<FAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DswSmem)));
    attr_set = true;
  }
  tpt::TileMap Min, Mfr; int n_in, n_fr;
  tpt::tile_maps(c->L, Min, Mfr, n_in, n_fr);
  // the line-per-warp kernels (tp_line.cuh) when only delp, [w,] pt are transported with the common schemes and no del-n flux
  // is added; interior tiles and frame tiles are separate instantiations (the frame one carries the cube-edge cells)
  const bool lines = FAM != 2 && use_line_kernels() && a.pt && !a.qcon && !a.dpx && !a.ptx && !a.qcx;
  const bool fp32 = lines && c->tp_fp32;
  const bool lines_fr = (lines && use_line_kernels() > 1) || fp32;
  if (lines) {
    constexpr int F2 = FAM == 2 ? 0 : FAM;
    const int kch = tp2::level_chunk(nk), nch = (nk + kch - 1) / kch, kch_fr = (kch + 1) / 2, nch_fr = (nk + kch_fr - 1) / kch_fr;
    // the scheme is a compile-time constant of the kernel when every field uses the same common one (hord 10 / 8 / 5 / 6)
    const bool same = a.hord_dp == a.hord_tm && (!a.w || a.hord_dp == a.hord_vt);
    const int hs = same ? a.hord_dp : tp2::ORD_RT;
    const DswVU vu0{};
    if (vu) {   // four fields (can_fuse_vort): interior and frame tiles both on the line kernels
#define TR4_LAUNCH1(H_, E_, MAP_, N_)                                                                                              \
    do {                                                                                                                           \
      constexpr int NEP4 = E_ ? 0 : 2;                                                                                             \
      static_assert(sizeof(tp2::Smem<4, NEP4, E_>) <= 232448, "four-field tile exceeds the 227 KB of shared memory");              \
      FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_transport2<F2, 4, H_, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<4, NEP4, E_>))); \
      k_dsw_transport2<F2, 4, H_, E_><<<dim3(N_, E_ ? nch_fr : nch), tp2::NT, sizeof(tp2::Smem<4, NEP4, E_>), c->stream>>>(c->L, c->G, MAP_, a, *vu, nk, E_ ? kch_fr : kch); \
    } while (0)
#define TR4_LAUNCH(H_)                                          \
    do {                                                        \
      if (n_in) TR4_LAUNCH1(H_, false, Min, n_in);              \
      if (n_fr) TR4_LAUNCH1(H_, true, Mfr, n_fr);               \
    } while (0)
      if constexpr (F2 == 1) {
        if (hs == 10) TR4_LAUNCH(10);
        else if (hs == 8) TR4_LAUNCH(8);
        else TR4_LAUNCH(tp2::ORD_RT);
      } else {
        if (hs == 5) TR4_LAUNCH(5);
        else if (hs == 6) TR4_LAUNCH(6);
        else TR4_LAUNCH(tp2::ORD_RT);
      }
#undef TR4_LAUNCH
#undef TR4_LAUNCH1
      c->launches += (n_in ? 1 : 0) + (n_fr ? 1 : 0);
      return 0;
    }
#define TR2_LAUNCH1(NF_, H_, E_, MAP_, N_)                                                                                         \
    do {                                                                                                                           \
      FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_transport2<F2, NF_, H_, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<NF_, 2, E_>))); \
      k_dsw_transport2<F2, NF_, H_, E_><<<dim3(N_, E_ ? nch_fr : nch), tp2::NT, sizeof(tp2::Smem<NF_, 2, E_>), c->stream>>>(c->L, c->G, MAP_, a, vu0, nk, E_ ? kch_fr : kch); \
    } while (0)
#define TR2_LAUNCH32(NF_, H_)                                                                                                      \
    do {                                                                                                                           \
      /* fp32 sweeps (strict float: tp_line.cuh): the interior and the frame instantiation give a shared face the same flux bit for bit */ \
      FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_transport2<F2, NF_, H_, false, tp2::sf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<NF_, 2, false, tp2::sf>))); \
      FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_transport2<F2, NF_, H_, true, tp2::sf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<NF_, 2, true, tp2::sf>))); \
      if (n_in) k_dsw_transport2<F2, NF_, H_, false, tp2::sf><<<dim3(n_in, nch), tp2::NT, sizeof(tp2::Smem<NF_, 2, false, tp2::sf>), c->stream>>>(c->L, c->G, Min, a, vu0, nk, kch); \
      if (n_fr) k_dsw_transport2<F2, NF_, H_, true, tp2::sf><<<dim3(n_fr, nch_fr), tp2::NT, sizeof(tp2::Smem<NF_, 2, true, tp2::sf>), c->stream>>>(c->L, c->G, Mfr, a, vu0, nk, kch_fr); \
    } while (0)
#define TR2_LAUNCH(NF_, H_)                                                        \
    do {                                                                           \
      if (fp32) { TR2_LAUNCH32(NF_, H_); break; }                                  \
      if (n_in) TR2_LAUNCH1(NF_, H_, false, Min, n_in);                            \
      if (n_fr && lines_fr) TR2_LAUNCH1(NF_, H_, true, Mfr, n_fr);                 \
    } while (0)
    if (a.w) {
      if constexpr (F2 == 1) {
        if (hs == 10) TR2_LAUNCH(3, 10);
        else if (hs == 8) TR2_LAUNCH(3, 8);
        else TR2_LAUNCH(3, tp2::ORD_RT);
      } else {
        if (hs == 5) TR2_LAUNCH(3, 5);
        else if (hs == 6) TR2_LAUNCH(3, 6);
        else TR2_LAUNCH(3, tp2::ORD_RT);
      }
    } else TR2_LAUNCH(2, tp2::ORD_RT);
#undef TR2_LAUNCH
#undef TR2_LAUNCH1
#undef TR2_LAUNCH32
  } else if (n_in) k_dsw_transport<FAM, false><<<dim3(n_in, 1, nk), tpt::NT, sizeof(DswSmem), c->stream>>>(c->L, c->G, Min, a);
  if (n_fr && !lines_fr) k_dsw_transport<FAM, true><<<dim3(n_fr, 1, nk), tpt::NT, sizeof(DswSmem), c->stream>>>(c->L, c->G, Mfr, a);
  c->launches += (n_in ? 1 : 0) + (n_fr ? 1 : 0);
  return 0;
}
static int launch_transport(fv3_ctx* c, const DswTr& a, int nk, const DswVU* vu = nullptr) {
  const bool m_dp = a.hord_dp >= 8, m_vt = a.hord_vt >= 8, m_tm = a.hord_tm >= 8;
  const bool vt_used = a.w != nullptr, tm_used = a.pt != nullptr;
  const bool all_mono = m_dp && (m_tm || !tm_used) && (m_vt || !vt_used), none_mono = !m_dp && (!m_tm || !tm_used) && (!m_vt || !vt_used);
  const bool rare = hord_is_rare(a.hord_dp) || (tm_used && hord_is_rare(a.hord_tm)) || (vt_used && hord_is_rare(a.hord_vt));
  if (rare) return launch_transport_t<2>(c, a, nk);   // the general instantiation carries the less common schemes
  if (all_mono) return launch_transport_t<1>(c, a, nk, vu);
  if (none_mono) return launch_transport_t<0>(c, a, nk, vu);
  return launch_transport_t<2>(c, a, nk);
}
template <int FM>
static int launch_vort_uv_t(fv3_ctx* c, const double* vq, const double* u, const double* v, const double* ke, double* uo, double* vo, int nk) {
  static bool attr_set = false;
  if (!attr_set) {
    FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_vort_uv<FM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_vort_uv<FM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    attr_set = true;
  }
  tpt::TileMap Min, Mfr; int n_in, n_fr;
  tpt::tile_maps(c->L, Min, Mfr, n_in, n_fr);
  // frame tiles of a single transported field: the first-generation kernel (two CTAs per SM) is faster than the line kernel with its
  // cube-edge tables (one CTA per SM): measured 217 vs 281 us at C384L79 (profiles/r2_dsw_kernels.md); FV3_TP_LINES=3 forces lines
  const bool lines = FM != 2 && use_line_kernels(), fp32 = lines && c->tp_fp32, lines_fr = lines && use_line_kernels() > 2;
  if (lines) {
    constexpr int F2 = FM == 2 ? 0 : FM;
    const int kch = tp2::level_chunk(nk), nch = (nk + kch - 1) / kch, kch_fr = (kch + 1) / 2, nch_fr = (nk + kch_fr - 1) / kch_fr;
    const int h = c->f.hord_vt;
#define VU2_LAUNCH1(H_, E_, MAP_, N_)                                                                                             \
    do {                                                                                                                          \
      FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_vort_uv2<F2, H_, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<1, 0, E_>))); \
      k_dsw_vort_uv2<F2, H_, E_><<<dim3(N_, E_ ? nch_fr : nch), E_ ? 1024 : TP2_NF1_NWC * 32, sizeof(tp2::Smem<1, 0, E_>), c->stream>>>(                                           \
          c->L, c->G, MAP_, vq, c->fld[FV3_CRX], c->fld[FV3_CRY], c->fld[FV3_XFX], c->fld[FV3_YFX], u, v, ke, uo, vo, h, nk, E_ ? kch_fr : kch); \
    } while (0)
#define VU2_LAUNCH32(H_)                                                                                                          \
    do {                                                                                                                          \
      /* every u, v point is updated by exactly one tile (no flux is applied from two sides), so the frame tiles may keep the fp64 kernel */ \
      FV3_CUDA(c, cudaFuncSetAttribute(k_dsw_vort_uv2<F2, H_, false, tp2::sf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<1, 0, false, tp2::sf>))); \
      k_dsw_vort_uv2<F2, H_, false, tp2::sf><<<dim3(n_in, nch), TP2_NF1_NWC * 32, sizeof(tp2::Smem<1, 0, false, tp2::sf>), c->stream>>>(                                     \
          c->L, c->G, Min, vq, c->fld[FV3_CRX], c->fld[FV3_CRY], c->fld[FV3_XFX], c->fld[FV3_YFX], u, v, ke, uo, vo, h, nk, kch); \
    } while (0)
#define VU2_LAUNCH(H_)                                                 \
    do {                                                               \
      if (n_in && fp32) VU2_LAUNCH32(H_);                              \
      else if (n_in) VU2_LAUNCH1(H_, false, Min, n_in);                \
      if (n_fr && lines_fr) VU2_LAUNCH1(H_, true, Mfr, n_fr);          \
    } while (0)
    if constexpr (F2 == 1) {
      if (h == 10) VU2_LAUNCH(10);
      else if (h == 8) VU2_LAUNCH(8);
      else VU2_LAUNCH(tp2::ORD_RT);
    } else {
      if (h == 5) VU2_LAUNCH(5);
      else if (h == 6) VU2_LAUNCH(6);
      else VU2_LAUNCH(tp2::ORD_RT);
    }
#undef VU2_LAUNCH
#undef VU2_LAUNCH1
#undef VU2_LAUNCH32
  } else if (n_in) k_dsw_vort_uv<FM, false><<<dim3(n_in, 1, nk), tpt::NT, sizeof(tpt::Smem), c->stream>>>(
      c->L, c->G, Min, vq, c->fld[FV3_CRX], c->fld[FV3_CRY], c->fld[FV3_XFX], c->fld[FV3_YFX], u, v, ke, uo, vo, c->f.hord_vt);
  if (n_fr && !lines_fr) k_dsw_vort_uv<FM, true><<<dim3(n_fr, 1, nk), tpt::NT, sizeof(tpt::Smem), c->stream>>>(
      c->L, c->G, Mfr, vq, c->fld[FV3_CRX], c->fld[FV3_CRY], c->fld[FV3_XFX], c->fld[FV3_YFX], u, v, ke, uo, vo, c->f.hord_vt);
  c->launches += (n_in ? 1 : 0) + (n_fr ? 1 : 0);
  return 0;
}

int stage_d_sw(fv3_ctx* c, double dt) {
  StageScope ts(c, "D_SW");
  const Lay& L = c->L;
  const fv3_flags_t& f = c->f;
  const int nk = L.npz;
  if (!hord_supported(f.hord_dp, f.lim_fac) || !hord_supported(f.hord_tm, f.lim_fac) || !hord_supported(f.hord_vt, f.lim_fac) || !hord_wind_supported(f.hord_mt, f.lim_fac))
    return fv3_fail(c, -2, "d_sw: unsupported hord (supported: -5, 1..13 [1 only with lim_fac = 1]; hord_mt: 1..11, 1 only with lim_fac = 1)");
  if (f.inline_q) return fv3_fail(c, -2, "d_sw: inline_q not supported");
  if (f.do_f3d) return fv3_fail(c, -2, "d_sw: do_f3d not supported");
  if (f.nord > 3) return fv3_fail(c, -2, "d_sw: nord > 3 not supported");
  std::vector<int> ki; std::vector<double> kd;
  dsw_tables(c, ki, kd);
  if (!c->capturing) {   // graph capture (dyn_core.cu): the tables depend on the flags only and were uploaded by an earlier direct call
    FV3_CUDA(c, cudaMemcpyAsync(c->d_kint, ki.data(), ki.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    FV3_CUDA(c, cudaMemcpyAsync(c->d_kdbl, kd.data(), kd.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // the tables are consumed asynchronously; ki/kd are pageable -> copy is staged by the driver before return
  }
  const int n1 = nk + 1;
  bool any_w = false, any_v = false, any_deln = false, any_deln_t = false, any_n0 = false, any_nn = false;
  int nord_max = 0;
  for (int k = 0; k < nk; k++) {
    any_w |= kd[KD_DAMP4_W * n1 + k] != 0.; any_v |= kd[KD_DAMP4_V * n1 + k] != 0.; any_deln |= kd[KD_DELN * n1 + k] != 0.; any_deln_t |= kd[KD_DELN_T * n1 + k] != 0.;
    any_n0 |= ki[KI_NORD * n1 + k] == 0; any_nn |= ki[KI_NORD * n1 + k] > 0;
    nord_max = std::max(nord_max, ki[KI_NORD * n1 + k]);
  }
  (void)any_n0; (void)any_nn;
  dim3 blk(TI, TJ), grd = plane_grid(L, nk);
  cudaStream_t st = c->stream;
  double *uts = c->scr[0], *vts = c->scr[1], *fx = c->scr[2], *fy = c->scr[3];
  double *fx2 = c->scr[4], *fy2 = c->scr[5], *q_i = c->scr[6], *q_j = c->scr[7];
  double *gx = c->scr[8], *gy = c->scr[9], *dfx = c->scr[10], *dfy = c->scr[11], *d2 = c->scr[12];
  double *dw = c->scr[13], *hs = c->scr[14], *ds = c->scr[15];
  double *delp = c->fld[FV3_DELP], *pt = c->fld[FV3_PT], *w = c->fld[FV3_W], *u = c->fld[FV3_U], *v = c->fld[FV3_V];
  double *uc = c->fld[FV3_UC], *vc = c->fld[FV3_VC];
  double *crx = c->fld[FV3_CRX], *cry = c->fld[FV3_CRY], *xfx = c->fld[FV3_XFX], *yfx = c->fld[FV3_YFX];

  if (f.sw_test_case == 1) {
    // SW_DYNAMICS build with test_case = 1 (BASELINE config 1a): delp is advected by the prescribed uc, vc; the momentum part,
    // pt, w, the damping of the winds are all skipped (sw_core.F90:626-651, :1055-1066, :1069, :1602)
    k_dsw_wind_sw1<<<grd, blk, 0, st>>>(L, c->G, uc, vc, crx, cry, xfx, yfx, c->fld[FV3_CX], c->fld[FV3_CY], dt);
    c->launches++;
    DswTr tr{};
    tr.delp = delp;
    tr.crx = crx; tr.cry = cry; tr.xfx = xfx; tr.yfx = yfx;
    tr.delp_o = c->alt_delp;
    tr.mfx = c->fld[FV3_MFX]; tr.mfy = c->fld[FV3_MFY]; tr.kdbl = c->d_kdbl;
    tr.hord_dp = f.hord_dp; tr.hord_vt = f.hord_dp; tr.hord_tm = f.hord_dp;
    if (any_deln) {   // :919-920 (nord = nord_v, damp_c = damp_v apply in the SW build too)
      Deln dl;
      dl.d2 = d2; dl.nk = nk; dl.thresh = 0; dl.nord_const = 0; dl.damp_const = 0;
      dl.q = delp; dl.fx2 = dfx; dl.fy2 = dfy; dl.slot_nord = KI_NORD_V; dl.slot_damp = KD_DELN; dl.premul = 1;
      dl.k_lo = nk; dl.k_hi = -1; dl.nord_max = 0;
      for (int k = 0; k < nk; k++)
        if (kd[KD_DELN * n1 + k] != 0.) { dl.k_lo = std::min(dl.k_lo, k); dl.k_hi = std::max(dl.k_hi, k); dl.nord_max = std::max(dl.nord_max, ki[KI_NORD_V * n1 + k]); }
      launch_deln(c, dl);
      tr.dpx = dfx; tr.dpy = dfy;
    }
    int rc1 = launch_transport(c, tr, nk); if (rc1) return rc1;
    FrameJobs fj{};
    fj.j[0] = FrameJob{delp, c->alt_delp, L.ie, L.je}; fj.n = 1;
    { const FrameGrid FG = frame_grid(L, L.is, L.ie, L.js, L.je); k_copy_frame<<<dim3(FG.count(), 1, nk), blk, 0, st>>>(L, FG, fj); }
    c->launches++;
    std::swap(c->fld[FV3_DELP], c->alt_delp);
    return 0;
  }
  k_dsw_wind<<<grd, blk, 0, st>>>(L, c->G, uc, vc, uts, vts, crx, cry, xfx, yfx, c->fld[FV3_CX], c->fld[FV3_CY], dt);
  c->launches++;

  // --- del-n damping fluxes of delp, w, q_con, pt (wide stencils: separate launches), then ONE fused transport
  Deln dl;
  dl.d2 = d2; dl.nk = nk; dl.thresh = 0; dl.nord_const = 0; dl.damp_const = 0;
  auto level_range = [&](Deln& d, int slot_nord, int slot_damp) {   // levels with a nonzero coefficient, and their largest order
    d.k_lo = nk; d.k_hi = -1; d.nord_max = 0;
    for (int k = 0; k < nk; k++)
      if (kd[slot_damp * n1 + k] != 0.) { d.k_lo = std::min(d.k_lo, k); d.k_hi = std::max(d.k_hi, k); d.nord_max = std::max(d.nord_max, ki[slot_nord * n1 + k]); }
  };
  const bool nonhydro = !f.hydrostatic;
  DswTr tr{};
  tr.delp = delp; tr.pt = pt; tr.w = nonhydro ? w : nullptr; tr.qcon = f.use_cond ? c->fld[FV3_QCON] : nullptr;
  tr.crx = crx; tr.cry = cry; tr.xfx = xfx; tr.yfx = yfx;
  tr.delp_o = c->alt_delp; tr.pt_o = c->alt_pt; tr.w_o = c->alt_w; tr.qcon_o = c->alt_qcon;
  tr.mfx = c->fld[FV3_MFX]; tr.mfy = c->fld[FV3_MFY]; tr.kdbl = c->d_kdbl;
  tr.hord_dp = f.hord_dp; tr.hord_vt = f.hord_vt; tr.hord_tm = f.hord_tm;
  if (any_deln) {   // delp (:919-920)
    dl.q = delp; dl.fx2 = dfx; dl.fy2 = dfy; dl.slot_nord = KI_NORD_V; dl.slot_damp = KD_DELN; level_range(dl, KI_NORD_V, KD_DELN); dl.premul = 1;
    launch_deln(c, dl);
    tr.dpx = dfx; tr.dpy = dfy;
  }
  if (nonhydro) {   // w (:950-990)
    if (any_w) {
      dl.q = w; dl.fx2 = fx2; dl.fy2 = fy2; dl.slot_nord = KI_NORD_W; dl.slot_damp = KD_DAMP4_W; level_range(dl, KI_NORD_W, KD_DAMP4_W); dl.premul = 1;
      launch_deln(c, dl);
      tr.dw = dw;
    }
    {
      const int need_hs = (f.d_con > 1.e-5 || f.do_diss_est) ? 1 : 0;
      Deln r; level_range(r, KI_NORD_W, KD_DAMP4_W);
      const int k_lo = need_hs ? 0 : r.k_lo, k_hi = need_hs ? nk - 1 : r.k_hi;
      if (k_hi >= k_lo) {
        k_dsw_dw<<<plane_grid(L, k_hi - k_lo + 1), blk, 0, st>>>(L, c->G, w, fx2, fy2, dw, hs, ds, c->d_kdbl, f.ke_bg, dt, f.prevent_diss_cooling,
                                                              f.do_diss_est, need_hs, k_lo);
        c->launches++;
      }
    }
  } else {
    FV3_CUDA(c, cudaMemsetAsync(hs, 0, (size_t)L.plane * nk * sizeof(double), st));
    FV3_CUDA(c, cudaMemsetAsync(ds, 0, (size_t)L.plane * nk * sizeof(double), st));
  }
  if (f.use_cond && any_deln_t) {   // q_con (:992-1000)
    dl.q = c->fld[FV3_QCON]; dl.fx2 = q_i; dl.fy2 = q_j; dl.slot_nord = KI_NORD_T; dl.slot_damp = KD_DELN_T; level_range(dl, KI_NORD_T, KD_DELN_T); dl.premul = 0;
    launch_deln(c, dl);
    tr.qcx = q_i; tr.qcy = q_j;
  }
  if (any_deln_t) {   // pt (:1014-1016)
    dl.q = pt; dl.fx2 = fx; dl.fy2 = fy; dl.slot_nord = KI_NORD_T; dl.slot_damp = KD_DELN_T; level_range(dl, KI_NORD_T, KD_DELN_T); dl.premul = 0;
    launch_deln(c, dl);
    tr.ptx = fx; tr.pty = fy;
  }
  // the scalar transport, the halo copy-through of its outputs and the ping-pong swap.  With vu (can_fuse_vort) the vorticity
  // transport and the momentum update ride along: the call then moves behind the kinetic energy / damping kernels that feed it
  // (nothing in between reads the transported scalars)
  auto do_transport = [&](const DswVU* vu) -> int {
    int rc_ = launch_transport(c, tr, nk, vu); if (rc_) return rc_;
    FrameJobs fj{}; int n = 0;
    fj.j[n++] = FrameJob{delp, c->alt_delp, L.ie, L.je};
    fj.j[n++] = FrameJob{pt, c->alt_pt, L.ie, L.je};
    if (nonhydro) fj.j[n++] = FrameJob{w, c->alt_w, L.ie, L.je};
    if (f.use_cond) fj.j[n++] = FrameJob{c->fld[FV3_QCON], c->alt_qcon, L.ie, L.je};
    if (vu) { fj.j[n++] = FrameJob{u, c->alt_u, L.ie, L.je + 1}; fj.j[n++] = FrameJob{v, c->alt_v, L.ie + 1, L.je}; }
    fj.n = n;
    { const FrameGrid FG = frame_grid(L, L.is, L.ie, L.js, L.je); k_copy_frame<<<dim3(FG.count(), 1, nk), blk, 0, st>>>(L, FG, fj); }
    c->launches++;
    std::swap(c->fld[FV3_DELP], c->alt_delp); std::swap(c->fld[FV3_PT], c->alt_pt);
    if (nonhydro) std::swap(c->fld[FV3_W], c->alt_w);
    if (f.use_cond) std::swap(c->fld[FV3_QCON], c->alt_qcon);
    return 0;
  };
  const bool fuse4 = nonhydro && can_fuse_vort(c, tr);
  int rc = 0;
  if (!fuse4) { rc = do_transport(nullptr); if (rc) return rc; }
  double* delp_new = fuse4 ? c->alt_delp : c->fld[FV3_DELP];
  // --- KE (:1078-1228); ke lives in fx2's plane from here (tp scratch is rewritten later, so use gx)
  double* ke = gx;      // B-grid (is:ie+1, js:je+1)
  if (hord_wind_is_rare(f.hord_mt)) k_dsw_ke<true><<<grd, blk, 0, st>>>(L, c->G, u, v, uc, vc, uts, vts, ke, dt, f.hord_mt);
  else k_dsw_ke<false><<<grd, blk, 0, st>>>(L, c->G, u, v, uc, vc, uts, vts, ke, dt, f.hord_mt);
  double *wk = uts, *vq = vts;   // contravariant winds are dead after KE
  k_dsw_vort<<<grd, blk, 0, st>>>(L, c->G, u, v, (f.dddmp >= 1.E-5 || any_v) ? wk : nullptr, vq);
  c->launches += 2;
  // --- divergence damping (:1290-1460)
  double *dg_even = gy, *dg_odd = dfx, *vortb = d2, *dterm = q_i;   // ping-pong planes of the damping passes
  if (nord_max > 0) {
    // pass n reads plane (n-1)&1 and writes plane n&1; plane "0" of pass 1 is divg_d itself (read-only)
    for (int n = 1; n <= nord_max; n++) {
      const double* src = (n == 1) ? c->fld[FV3_DIVGD] : ((n - 1) & 1) ? dg_odd : dg_even;
      k_dsw_dd_iter<<<grd, blk, 0, st>>>(L, c->G, src, (n & 1) ? dg_odd : dg_even, c->d_kint, n, c->b.stretched_grid);
      c->launches++;
    }
    if (f.dddmp >= 1.E-5) {
      if (!L.cube) return fv3_fail(c, -2, "d_sw: smag_corner (dddmp>0 on a doubly-periodic grid) not supported");
      rc = launch_a2b_ord4(c, wk, vortb, nk, 0); if (rc) return rc;
    }
  }
  k_dsw_damp<<<grd, blk, 0, st>>>(L, c->G, u, v, c->fld[FV3_UA], c->fld[FV3_VA], uc, vc, c->fld[FV3_DIVGD], dg_even, dg_odd, vortb, ke,
                                  (f.d_con > 1.e-5 || f.do_diss_est) ? dterm : nullptr,
                                  c->d_kint, c->d_kdbl, dt, f.dddmp, f.d4_bg, c->b.stretched_grid, f.d_ext > 0. ? c->fld[FV3_VT] : nullptr);
  c->launches++;
  // --- vorticity transport and momentum update (:1476-1509), fused
  if (fuse4) {
    const DswVU vu{vq, u, v, ke, c->alt_u, c->alt_v};
    rc = do_transport(&vu);
  } else
  rc = hord_is_rare(f.hord_vt) ? launch_vort_uv_t<2>(c, vq, u, v, ke, c->alt_u, c->alt_v, nk)
       : (f.hord_vt >= 8)      ? launch_vort_uv_t<1>(c, vq, u, v, ke, c->alt_u, c->alt_v, nk)
                               : launch_vort_uv_t<0>(c, vq, u, v, ke, c->alt_u, c->alt_v, nk);
  if (rc) return rc;
  if (!fuse4) {
    FrameJobs fj{};
    fj.j[0] = FrameJob{u, c->alt_u, L.ie, L.je + 1};
    fj.j[1] = FrameJob{v, c->alt_v, L.ie + 1, L.je};
    fj.n = 2;
    { const FrameGrid FG = frame_grid(L, L.is, L.ie, L.js, L.je); k_copy_frame<<<dim3(FG.count(), 1, nk), blk, 0, st>>>(L, FG, fj); }
    c->launches++;
  }
  // --- vorticity damping + dissipative heating (:1513-1600)
  if (any_v) {
    dl.q = wk; dl.fx2 = dfx; dl.fy2 = dfy; dl.slot_nord = KI_NORD_V; dl.slot_damp = KD_DAMP4_V; level_range(dl, KI_NORD_V, KD_DAMP4_V); dl.premul = 1;
    launch_deln(c, dl);   // dfx = "ut", dfy = "vt"
  }
  if (f.d_con > 1.e-5 || f.do_diss_est) {
    k_dsw_heat<<<grd, blk, 0, st>>>(L, c->G, c->alt_u, c->alt_v, u, v, dterm, dfx, dfy, delp_new, hs, ds, c->fld[FV3_HEAT], c->fld[FV3_DISS],
                                    c->d_kdbl, f.prevent_diss_cooling, f.do_diss_est, f.d_con);
    c->launches++;
  }
  if (any_v) {
    k_dsw_vdamp<<<grd, blk, 0, st>>>(L, c->alt_u, c->alt_v, dfx, dfy, c->d_kdbl);
    c->launches++;
  }
  std::swap(c->fld[FV3_U], c->alt_u); std::swap(c->fld[FV3_V], c->alt_v);
  return 0;
}
