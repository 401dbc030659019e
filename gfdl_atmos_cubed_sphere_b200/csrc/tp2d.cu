// fv_tp_2d (Lin-Rood 2-D flux-form transport with PPM operators) and the del-n damping
// fluxes, as CUDA kernels for sm_100a.
//
// Reference semantics: model/tp_core.F90:85-241 fv_tp_2d, :245-322 copy_corners,
// :1267-1447 deln_flux; model/sw_core.F90:1608-1737 del6_vt_flux.
// Design (not a translation): ONE launch pair (interior / frame tiles) per transport; a CTA owns a 26x26 tile of one level
// and keeps the inner fluxes, q_i and q_j in shared memory (tp_tile.cuh).
// The cube-corner "copy_corners" transposes are NOT written into q: corner tiles load q
// through a remapping accessor (ppm::QAccX / QAccY), so q stays read-only.
#include "tp2d.cuh"
#include "ppm.cuh"
#include "tp_tile.cuh"
#include "tp_line.cuh"

using namespace ppm;

#define TI 32
#define TJ 8

static inline dim3 plane_grid(const Lay& L, int nk) { return dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, nk); }

#define PLANE_IJK_OFS(kofs)                                    \
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x; \
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;          \
  const int k = blockIdx.z + (kofs);                            \
  const long long ko = (long long)k * L.plane;
#define PLANE_IJK PLANE_IJK_OFS(0)

// One CTA per TX x TY tile and level: see tp_tile.cuh.  Epilogue: weight by the area flux (or the
// mass flux, tp_core.F90:213-226) and store the tile's own faces (+ the face's last column / row).
template <int FAM, bool EDGE>
__global__ void __launch_bounds__(tpt::NT, 2) k_tp_fused(Lay L, DevGrid G, tpt::TileMap M, const double* __restrict__ q,
                                                       const double* __restrict__ crx, const double* __restrict__ cry,
                                                       const double* __restrict__ xfx, const double* __restrict__ yfx,
                                                       const double* __restrict__ ra_x, const double* __restrict__ ra_y,
                                                       const double* __restrict__ mfx, const double* __restrict__ mfy,
                                                       double* __restrict__ fx, double* __restrict__ fy, int ord_in, int ord_ou,
                                                       tpt::ZnEpi Z) {
  using namespace tpt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const Tile T = make_tile(L, M);
  stage_inputs<EDGE>(L, G, S, T, crx, cry, xfx, yfx);
  stage_q<EDGE>(L, S, T, q);
  tp_compute<FAM, EDGE>(L, G, S, T, ra_x, ra_y, ord_in, ord_ou);
  // epilogue: lane = column; the tile stores its own west/south faces, plus the face's last column / row
  const bool lastx = T.i0 + TX > L.ie, lasty = T.j0 + TY > L.je;
  const int c = T.lane - 3, i = T.i0 + c;
  if (Z.zn) {
    // update_dz_d (nh_utils.F90:282-299): the advected height straight from the tile, fluxes never stored:
    //   zn = (zh*area + fx(i)-fx(i+1) + fy(j)-fy(j+1)) / (ra_x + ra_y - area)  [+ del-n damping flux divergence]
    const double coef = Z.kdbl ? Z.kdbl[Z.slot * (L.npz + 1) + blockIdx.z] : 0.;
#pragma unroll
    for (int r = T.wid; r < TY; r += NW) {
      const int j = T.j0 + r;
      if (c < 0 || c >= TX || i > L.ie || j > L.je) continue;
      const int o = T.idx(i, j);
      const double ar = S.area[r + 3][c + 3];
      const double x0 = S.xfx[r + 3][c + 3], x1 = S.xfx[r + 3][c + 4], y0 = S.yfx[r + 3][c + 3], y1 = S.yfx[r + 4][c + 3];
      const double rax = ar + x0 - x1, ray = ar + y0 - y1;
      double z = (S.q[r + 3][c + 3] * ar + FX(S, r, c) * x0 - FX(S, r, c + 1) * x1 + FY(S, r, c) * y0 - FY(S, r + 1, c) * y1) / (rax + ray - ar);
      if (coef != 0.) {
        const long long g = T.ko + o;
        z = z + (Z.dfx[g] - Z.dfx[g + 1] + Z.dfy[g] - Z.dfy[g + T.NI]) * __ldg(G.rarea + o);
      }
      Z.zn[T.ko + o] = z;
    }
    return;
  }
  fx += T.ko; fy += T.ko;
#pragma unroll
  for (int r = T.wid; r <= TY; r += NW) {
    const int j = T.j0 + r;
    if (c < 0 || j > L.je + 1) continue;
    const int o = T.idx(i, j);
    if (r < TY && j <= L.je && i <= L.ie + 1 && (c < TX || (c == TX && lastx)))
      fx[o] = FX(S, r, c) * (mfx ? __ldg(mfx + T.ko + o) : S.xfx[r + 3][c + 3]);
    if (c < TX && i <= L.ie && (r < TY || lasty))
      fy[o] = FY(S, r, c) * (mfy ? __ldg(mfy + T.ko + o) : S.yfx[r + 3][c + 3]);
  }
}

// interior tiles of the fused height update (update_dz_d, nh_utils.F90:282-299) in the line-per-warp form (tp_line.cuh)
template <int FAM, int HORD, bool EDGE>
__global__ void __launch_bounds__(1024, 1) k_tp_zn2(Lay L, DevGrid G, tpt::TileMap M, const double* __restrict__ q, const double* __restrict__ crx,
                                                 const double* __restrict__ cry, const double* __restrict__ xfx, const double* __restrict__ yfx,
                                                 int ord_in_, int ord_ou_, tpt::ZnEpi Z, int nk, int kch) {
  const double* src[5] = {crx, cry, xfx, yfx, q};
  const int ord_in[1] = {ord_in_}, ord_ou[1] = {ord_ou_};
  tp2::run_tile<FAM, 1, 0, tp2::W_AREA, HORD, 32, EDGE>(L, G, M, src, nk, kch, ord_in, ord_ou,
    [&](tp2::Smem<1, 0, EDGE>&, const tp2::Geo&, int, long long, int) {},
    [&](tp2::Smem<1, 0, EDGE>& S, int b, const tp2::Geo& T, int k, long long ko, int r) {
      const int c = T.lane, i = T.i0 - 3 + c, j = T.j0 - 3 + r;
      if (c < 3 || c > tp2::TX + 2 || r > tp2::TY + 2 || (EDGE && (i > L.ie || j > L.je))) return;
      const int o = r * tp2::P + c;
      const int gi = tp2::gidx(T, i, j);
      const double ar = S.area[o];
      const double x0 = S.in[b][tp2::A_XFX][o], x1 = S.in[b][tp2::A_XFX][o + 1], y0 = S.in[b][tp2::A_YFX][o], y1 = S.in[b][tp2::A_YFX][o + tp2::P];
      const double rax = ar + x0 - x1, ray = ar + y0 - y1;
      double z = (S.in[b][tp2::A_Q][o] * ar + S.qi[0][o] - S.qi[0][o + 1] + S.qj[0][o] - S.qj[0][o + tp2::P]) / (rax + ray - ar);
      const double coef = Z.kdbl ? Z.kdbl[Z.slot * (L.npz + 1) + k] : 0.;
      if (coef != 0.) {
        const long long g = ko + gi;
        z = z + (Z.dfx[g] - Z.dfx[g + 1] + Z.dfy[g] - Z.dfy[g + T.NI]) * __ldg(G.rarea + gi);
      }
      Z.zn[ko + gi] = z;
    });
}

int launch_tp2d(fv3_ctx* c, const Tp2d& a) {
  if (!hord_supported(a.hord, c->f.lim_fac)) return fv3_fail(c, -2, "fv_tp_2d: hord " + std::to_string(a.hord) + " not supported on the GPU path (supported: -5, 1..13; 1 only with lim_fac = 1)");
  const Lay& L = c->L;
  const int ord_in = (a.hord == 10) ? 8 : a.hord;   // tp_core.F90:136-141
  tpt::TileMap Min, Mfr; int n_in, n_fr;
  tpt::tile_maps(L, Min, Mfr, n_in, n_fr);
  static bool attr_set = false;
  if (!attr_set) {
    FV3_CUDA(c, cudaFuncSetAttribute(k_tp_fused<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_tp_fused<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_tp_fused<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_tp_fused<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_tp_fused<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c, cudaFuncSetAttribute(k_tp_fused<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    attr_set = true;
  }
  const tpt::ZnEpi Z{a.zn, a.zn_dfx, a.zn_dfy, a.zn ? c->d_kdbl : nullptr, a.zn_slot};
  if (a.zn && (a.ra_x || a.ra_y || a.mfx)) return fv3_fail(c, -1, "fv_tp_2d: the fused height update takes no ra_x / ra_y / mfx");
  // the fused height update with the common schemes: line-per-warp kernel (tp_line.cuh) on the interior tiles; the frame tiles keep
  // the first-generation kernel (faster for a single field, see launch_vort_uv_t in d_sw.cu) unless FV3_TP_LINES=3
  static int lines_on = -1;
  if (lines_on < 0) { const char* e = getenv("FV3_TP_LINES"); lines_on = e ? atoi(e) : 2; }
  if (a.zn && lines_on && !hord_is_rare(a.hord)) {
    const int kch = tp2::level_chunk(a.nk), nch = (a.nk + kch - 1) / kch, kch_fr = (kch + 1) / 2, nch_fr = (a.nk + kch_fr - 1) / kch_fr;
#define ZN2_LAUNCH1(F_, H_, E_, MAP_, N_)                                                                                          \
    do {                                                                                                                           \
      FV3_CUDA(c, cudaFuncSetAttribute(k_tp_zn2<F_, H_, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tp2::Smem<1, 0, E_>))); \
      k_tp_zn2<F_, H_, E_><<<dim3(N_, E_ ? nch_fr : nch), 1024, sizeof(tp2::Smem<1, 0, E_>), c->stream>>>(L, c->G, MAP_, a.q, a.crx, a.cry, a.xfx, a.yfx,    \
                                                                                  ord_in, a.hord, Z, a.nk, E_ ? kch_fr : kch);     \
      c->launches++;                                                                                                               \
    } while (0)
#define ZN2_LAUNCH(F_, H_)                                                          \
    do {                                                                            \
      if (n_in) { ZN2_LAUNCH1(F_, H_, false, Min, n_in); n_in = 0; }                \
      if (n_fr && lines_on > 2) { ZN2_LAUNCH1(F_, H_, true, Mfr, n_fr); n_fr = 0; } \
    } while (0)
    if (a.hord == 10) ZN2_LAUNCH(1, 10);
    else if (a.hord == 8) ZN2_LAUNCH(1, 8);
    else if (a.hord == 5) ZN2_LAUNCH(0, 5);
    else if (a.hord == 6) ZN2_LAUNCH(0, 6);
    else ZN2_LAUNCH(0, tp2::ORD_RT);   // -5
#undef ZN2_LAUNCH
#undef ZN2_LAUNCH1
  }
#define TP_LAUNCH(FAM, EDGE, M, N)                                                                                           \
  k_tp_fused<FAM, EDGE><<<dim3(N, 1, a.nk), tpt::NT, sizeof(tpt::Smem), c->stream>>>(L, c->G, M, a.q, a.crx, a.cry, a.xfx, a.yfx, \
                                                                                      a.ra_x, a.ra_y, a.mfx, a.mfy, a.fx, a.fy, ord_in, a.hord, Z)
  if (hord_is_rare(a.hord)) { if (n_in) TP_LAUNCH(2, false, Min, n_in); if (n_fr) TP_LAUNCH(2, true, Mfr, n_fr); }   // general instantiation
  else if (a.hord >= 8) { if (n_in) TP_LAUNCH(1, false, Min, n_in); if (n_fr) TP_LAUNCH(1, true, Mfr, n_fr); }
  else { if (n_in) TP_LAUNCH(0, false, Min, n_in); if (n_fr) TP_LAUNCH(0, true, Mfr, n_fr); }
#undef TP_LAUNCH
  c->launches += (n_in ? 1 : 0) + (n_fr ? 1 : 0);
  return 0;
}

// ------------------------------------------------------------------ del-n fluxes
__device__ __forceinline__ void kparams(const int* kint, const double* kdbl, int npz1, int slot_nord, int slot_damp, int k,
                                        int nord_const, double damp_const, int& nord, double& coef) {
  nord = slot_nord >= 0 ? kint[slot_nord * npz1 + k] : nord_const;
  coef = slot_damp >= 0 ? kdbl[slot_damp * npz1 + k] : damp_const;
}

__global__ void __launch_bounds__(TI* TJ) k_deln_first(Lay L, DevGrid G, const double* __restrict__ q, double* __restrict__ fx2,
                                                      double* __restrict__ fy2, const int* kint, const double* kdbl, int slot_nord,
                                                      int slot_damp, int nord_const, double damp_const, int premul, int kofs) {
  PLANE_IJK_OFS(kofs)
  int nord; double coef;
  kparams(kint, kdbl, L.npz + 1, slot_nord, slot_damp, k, nord_const, damp_const, nord, coef);
  if (coef == 0.) return;
  const double m = premul ? coef : 1.0;
  if (i >= L.is - nord && i <= L.ie + nord + 1 && j >= L.js - nord && j <= L.je + nord) {
    QAccX qa{q + ko, L, j};
    fx2[ko + LIDX(L, i, j)] = __ldg(G.del6_v + LIDX(L, i, j)) * (m * qa(i - 1) - m * qa(i));
  }
  if (i >= L.is - nord && i <= L.ie + nord && j >= L.js - nord && j <= L.je + nord + 1) {
    QAccY qa{q + ko, L, i};
    fy2[ko + LIDX(L, i, j)] = __ldg(G.del6_u + LIDX(L, i, j)) * (m * qa(j - 1) - m * qa(j));
  }
}

__global__ void __launch_bounds__(TI* TJ) k_deln_d2(Lay L, DevGrid G, const double* __restrict__ fx2, const double* __restrict__ fy2,
                                                   double* __restrict__ d2, const int* kint, const double* kdbl, int slot_nord,
                                                   int slot_damp, int nord_const, double damp_const, int n, int kofs) {
  PLANE_IJK_OFS(kofs)
  int nord; double coef;
  kparams(kint, kdbl, L.npz + 1, slot_nord, slot_damp, k, nord_const, damp_const, nord, coef);
  if (coef == 0. || n > nord) return;
  const int nt = nord - n;
  if (i < L.is - nt - 1 || i > L.ie + nt + 1 || j < L.js - nt - 1 || j > L.je + nt + 1) return;
  const long long o = ko + LIDX(L, i, j);
  d2[o] = (fx2[o] - fx2[o + 1] + fy2[o] - fy2[o + L.NI]) * __ldg(G.rarea + LIDX(L, i, j));
}

__global__ void __launch_bounds__(TI* TJ) k_deln_flux(Lay L, DevGrid G, const double* __restrict__ d2, double* __restrict__ fx2,
                                                     double* __restrict__ fy2, const int* kint, const double* kdbl, int slot_nord,
                                                     int slot_damp, int nord_const, double damp_const, int n, int kofs) {
  PLANE_IJK_OFS(kofs)
  int nord; double coef;
  kparams(kint, kdbl, L.npz + 1, slot_nord, slot_damp, k, nord_const, damp_const, nord, coef);
  if (coef == 0. || n > nord) return;
  const int nt = nord - n;
  if (i >= L.is - nt && i <= L.ie + nt + 1 && j >= L.js - nt && j <= L.je + nt) {
    QAccX da{d2 + ko, L, j};
    fx2[ko + LIDX(L, i, j)] = __ldg(G.del6_v + LIDX(L, i, j)) * (da(i) - da(i - 1));
  }
  if (i >= L.is - nt && i <= L.ie + nt && j >= L.js - nt && j <= L.je + nt + 1) {
    QAccY da{d2 + ko, L, i};
    fy2[ko + LIDX(L, i, j)] = __ldg(G.del6_u + LIDX(L, i, j)) * (da(j) - da(j - 1));
  }
}

int launch_deln(fv3_ctx* c, const Deln& a) {
  const Lay& L = c->L;
  const int k_lo = a.k_hi >= a.k_lo ? a.k_lo : 0, k_hi = a.k_hi >= a.k_lo ? a.k_hi : a.nk - 1;
  dim3 blk(TI, TJ), grd = plane_grid(L, k_hi - k_lo + 1);
  const int nmax = a.nord_max >= 0 ? a.nord_max : (a.slot_nord >= 0 ? 2 : a.nord_const);
  k_deln_first<<<grd, blk, 0, c->stream>>>(L, c->G, a.q, a.fx2, a.fy2, c->d_kint, c->d_kdbl, a.slot_nord, a.slot_damp,
                                           a.nord_const, a.damp_const, a.premul, k_lo);
  c->launches++;
  for (int n = 1; n <= nmax; n++) {
    k_deln_d2<<<grd, blk, 0, c->stream>>>(L, c->G, a.fx2, a.fy2, a.d2, c->d_kint, c->d_kdbl, a.slot_nord, a.slot_damp,
                                          a.nord_const, a.damp_const, n, k_lo);
    k_deln_flux<<<grd, blk, 0, c->stream>>>(L, c->G, a.d2, a.fx2, a.fy2, c->d_kint, c->d_kdbl, a.slot_nord, a.slot_damp,
                                            a.nord_const, a.damp_const, n, k_lo);
    c->launches += 2;
  }
  return 0;
}

// add the del-n fluxes into fx, fy (tp_core.F90:1390-1445).  With mass: fx += 0.5*damp*(m(i-1)+m(i))*fx2
__global__ void __launch_bounds__(TI* TJ) k_deln_add(Lay L, double* __restrict__ fx, double* __restrict__ fy,
                                                    const double* __restrict__ fx2, const double* __restrict__ fy2,
                                                    const double* __restrict__ mass, const double* kdbl, int slot_damp,
                                                    double damp_const) {
  PLANE_IJK
  const double coef = slot_damp >= 0 ? kdbl[slot_damp * (L.npz + 1) + k] : damp_const;
  if (coef == 0.) return;
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const long long o = ko + LIDX(L, i, j);
  if (j <= L.je) {
    if (mass) fx[o] = fx[o] + (0.5 * coef) * (__ldg(mass + o - 1) + __ldg(mass + o)) * fx2[o];
    else fx[o] = fx[o] + fx2[o];
  }
  if (i <= L.ie) {
    if (mass) fy[o] = fy[o] + (0.5 * coef) * (__ldg(mass + o - L.NI) + __ldg(mass + o)) * fy2[o];
    else fy[o] = fy[o] + fy2[o];
  }
}

void launch_deln_add(fv3_ctx* c, double* fx, double* fy, const double* fx2, const double* fy2, const double* mass,
                     int slot_damp, double damp_const, int nk) {
  dim3 blk(TI, TJ), grd = plane_grid(c->L, nk);
  k_deln_add<<<grd, blk, 0, c->stream>>>(c->L, fx, fy, fx2, fy2, mass, c->d_kdbl, slot_damp, damp_const);
  c->launches++;
}

// ------------------------------------------------------------------ stand-alone stage (C ABI fv3_fv_tp_2d)
int stage_fv_tp_2d(fv3_ctx* c, int nk, int hord, int use_mfx, int use_mass, int nord, double damp_c) {
  if (nk < 1 || nk > c->L.npz) return fv3_fail(c, -1, "fv_tp_2d: bad nk");
  Tp2d a;
  a.q = c->fld[FV3_WORK_Q]; a.crx = c->fld[FV3_CRX]; a.cry = c->fld[FV3_CRY]; a.xfx = c->fld[FV3_XFX]; a.yfx = c->fld[FV3_YFX];
  a.ra_x = c->fld[FV3_WORK_RAX]; a.ra_y = c->fld[FV3_WORK_RAY];
  a.fx = c->fld[FV3_WORK_FX]; a.fy = c->fld[FV3_WORK_FY];
  a.mfx = use_mfx ? c->fld[FV3_MFX] : nullptr; a.mfy = use_mfx ? c->fld[FV3_MFY] : nullptr;
  a.hord = hord; a.nk = nk;
  if (nk > c->L.npz && (use_mfx || use_mass)) return fv3_fail(c, -1, "fv_tp_2d: mfx/mass only for nk <= npz");
  int rc = launch_tp2d(c, a);
  if (rc) return rc;
  if (nord >= 0 && damp_c > 1.e-4) {   // tp_core.F90:201-206 / :227-232
    if (nord > 2) return fv3_fail(c, -2, "deln_flux: nord > 2 not supported");
    if (use_mfx && !use_mass) return 0;   // needs mass when mfx present (tp_core.F90:201)
    const double damp = pow(damp_c * c->G.da_min, (double)(nord + 1));
    Deln d;
    d.q = a.q; d.fx2 = c->scr[4]; d.fy2 = c->scr[5]; d.d2 = c->scr[6];
    d.slot_nord = -1; d.slot_damp = -1; d.thresh = 0; d.premul = use_mass ? 0 : 1; d.nk = nk; d.nord_const = nord; d.damp_const = damp;
    launch_deln(c, d);
    launch_deln_add(c, a.fx, a.fy, d.fx2, d.fy2, use_mass ? c->fld[FV3_DELP] : nullptr, -1, damp, nk);
  }
  return 0;
}
