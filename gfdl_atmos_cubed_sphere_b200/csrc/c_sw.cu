// c_sw: half-step C-grid shallow-water sweep for all levels in one batched launch sequence.
//
// Reference semantics: model/sw_core.F90:79-488 c_sw, :3006-3345 d2a2c_vect,
// :1740-1845 divergence_corner, :3348-3359 edge_interpolate4, :3434-3555 fill*_4corners;
// call site model/dyn_core.F90:436-447 (one call per k under OpenMP -> here ONE call,
// k = blockIdx.z).
// Design: 4 kernels, one thread per output point, k batched in gridDim.z.
//   k_a: 4th/2nd-order D->A (utmp, vtmp) + contravariant ua, va
//   k_c: A->C (uc, ut, vc, vt incl. the 6 one-sided edge columns/rows), Courant scaling of
//        ut, vt, and divergence_corner
//   k_t: upwind transport delpc/ptc/wc + KE + C-grid absolute vorticity
//   k_u: uc, vc update
// The reference's in-place corner fills (fill2_4corners / the utmp,vtmp,ua,va corner
// rotations) are read-side index remaps (FillX/FillY, UtmpX, VtmpY, UaX, VaY): no array is
// modified outside its own output region, so no inter-block ordering is needed.
#include "fv3_ctx.hpp"

#define TI 32
#define TJ 8
// resident CTAs per SM the register allocation of the load-latency-bound kernels is sized for (A/B: profiles/r2_dsw_summary.md)
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
#define CG(n, i, j) __ldg(G.cos_sg + (long long)((n)-1) * L.plane + LIDX(L, (i), (j)))

static inline dim3 plane_grid(const Lay& L, int nk) { return dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, nk); }

namespace {
// sw_core.F90:53-59
constexpr double a1 = 0.5625, a2 = -0.0625;
constexpr double c1 = -2. / 14., c2 = 11. / 14., c3 = 5. / 14.;
constexpr double big_number = 1.E30;

// fill_4corners views (sw_core.F90:3496-3555): dir=1 for x-fluxes, dir=2 for y-fluxes
template <bool E>
struct FillX {
  const double* q; Lay L;
  __device__ __forceinline__ double operator()(int i, int j) const {
    if (E && L.cube) {
      if (j == 0) {
        if (i == -1) { i = 0; j = 2; } else if (i == 0) { j = 1; }
        else if (i == L.npx + 1) { i = L.npx; j = 2; } else if (i == L.npx) { j = 1; }
      } else if (j == L.npy) {
        if (i == 0) { j = L.npy - 1; } else if (i == -1) { i = 0; j = L.npy - 2; }
        else if (i == L.npx) { j = L.npy - 1; } else if (i == L.npx + 1) { i = L.npx; j = L.npy - 2; }
      }
    }
    return __ldg(q + LIDX(L, i, j));
  }
};
template <bool E>
struct FillY {
  const double* q; Lay L;
  __device__ __forceinline__ double operator()(int i, int j) const {
    if (E && L.cube) {
      if (i == 0) {
        if (j == 0) { i = 1; } else if (j == -1) { i = 2; j = 0; }
        else if (j == L.npy) { i = 1; } else if (j == L.npy + 1) { i = 2; j = L.npy; }
      } else if (i == L.npx) {
        if (j == 0) { i = L.npx - 1; } else if (j == -1) { i = L.npx - 2; j = 0; }
        else if (j == L.npy) { i = L.npx - 1; } else if (j == L.npy + 1) { i = L.npx - 2; j = L.npy; }
      }
    }
    return __ldg(q + LIDX(L, i, j));
  }
};

__device__ __forceinline__ double edge_interpolate4(double u1, double u2, double u3, double u4, double d1, double d2,
                                                    double d3, double d4) {  // sw_core.F90:3348
  const double t1 = d1 + d2, t2 = d3 + d4;
  return 0.5 * (((t1 + d2) * u2 - d2 * u1) / t1 + ((t2 + d3) * u3 - d3 * u4) / t2);
}
}  // namespace

// ---- k_a: utmp, vtmp, ua, va  (sw_core.F90:3060-3157) -------------------------------------
__global__ void __launch_bounds__(TI* TJ) k_csw_a(Lay L, DevGrid G, const double* __restrict__ u, const double* __restrict__ v,
                                                 double* __restrict__ utmp, double* __restrict__ vtmp, double* __restrict__ ua,
                                                 double* __restrict__ va) {
  PLANE_IJK
  if (i < L.isd || i > L.ied || j > L.jed) return;
  const long long o = ko + LIDX(L, i, j);
  double ut = big_number, vt = big_number;
  if (L.cube) {
    const int npt = 4;
    if (i >= npt && i <= L.npx - npt && j >= npt && j <= L.npy - npt) {
      ut = a2 * (AT(u, i, j - 1) + AT(u, i, j + 2)) + a1 * (AT(u, i, j) + AT(u, i, j + 1));
      vt = a2 * (AT(v, i - 1, j) + AT(v, i + 2, j)) + a1 * (AT(v, i, j) + AT(v, i + 1, j));
    } else {
      ut = 0.5 * (AT(u, i, j) + AT(u, i, j + 1));
      vt = 0.5 * (AT(v, i, j) + AT(v, i + 1, j));
    }
  } else {
    if (j >= L.js - 1 && j <= L.je + 1) ut = a2 * (AT(u, i, j - 1) + AT(u, i, j + 2)) + a1 * (AT(u, i, j) + AT(u, i, j + 1));
    if (i >= L.is - 1 && i <= L.ie + 1) vt = a2 * (AT(v, i - 1, j) + AT(v, i + 2, j)) + a1 * (AT(v, i, j) + AT(v, i + 1, j));
  }
  utmp[o] = ut; vtmp[o] = vt;
  if (i >= L.is - 2 && i <= L.ie + 2 && j >= L.js - 2 && j <= L.je + 2) {
    const double cs = G2(cosa_s, i, j), rs = G2(rsin2, i, j);
    ua[o] = (ut - vt * cs) * rs;
    va[o] = (vt - ut * cs) * rs;
  }
}

namespace {
// utmp with the A->C x-direction corner rotation (sw_core.F90:3166-3185)
template <bool E>
struct UtmpX {
  const double *ut, *vt; Lay L;
  __device__ __forceinline__ double operator()(int i, int j) const {
    if (E && L.cube) {
      if (j == 0) {
        if (i <= 0) return -__ldg(vt + LIDX(L, 0, 1 - i));
        if (i >= L.npx) return __ldg(vt + LIDX(L, L.npx, i - L.npx + 1));
      } else if (j == L.npy) {
        if (i >= L.npx) return -__ldg(vt + LIDX(L, L.npx, L.je - (i - L.npx)));
        if (i <= 0) return __ldg(vt + LIDX(L, 0, L.je + i));
      }
    }
    return __ldg(ut + LIDX(L, i, j));
  }
};
// vtmp with the y-direction corner rotation (sw_core.F90:3259-3278)
template <bool E>
struct VtmpY {
  const double *ut, *vt; Lay L;
  __device__ __forceinline__ double operator()(int i, int j) const {
    if (E && L.cube) {
      if (i == 0) {
        if (j <= 0) return -__ldg(ut + LIDX(L, 1 - j, 0));
        if (j >= L.npy) return __ldg(ut + LIDX(L, j - L.npy + 1, L.npy));
      } else if (i == L.npx) {
        if (j <= 0) return __ldg(ut + LIDX(L, L.ie + j, 0));
        if (j >= L.npy) return -__ldg(ut + LIDX(L, L.ie - (j - L.npy), L.npy));
      }
    }
    return __ldg(vt + LIDX(L, i, j));
  }
};
// ua with the corner rotation used by edge_interpolate4 along x (sw_core.F90:3206-3221)
template <bool E>
struct UaX {
  const double *ua, *va; Lay L;
  __device__ __forceinline__ double operator()(int i, int j) const {
    if (E && L.cube) {
      if (j == 0) {
        if (i == -1) return -__ldg(va + LIDX(L, 0, 2));
        if (i == 0) return -__ldg(va + LIDX(L, 0, 1));
        if (i == L.npx) return __ldg(va + LIDX(L, L.npx, 1));
        if (i == L.npx + 1) return __ldg(va + LIDX(L, L.npx, 2));
      } else if (j == L.npy) {
        if (i == L.npx) return -__ldg(va + LIDX(L, L.npx, L.npy - 1));
        if (i == L.npx + 1) return -__ldg(va + LIDX(L, L.npx, L.npy - 2));
        if (i == -1) return __ldg(va + LIDX(L, 0, L.npy - 2));
        if (i == 0) return __ldg(va + LIDX(L, 0, L.npy - 1));
      }
    }
    return __ldg(ua + LIDX(L, i, j));
  }
};
// va with the corner rotation along y (sw_core.F90:3279-3294); reads the ORIGINAL ua
template <bool E>
struct VaY {
  const double *ua, *va; Lay L;
  __device__ __forceinline__ double operator()(int i, int j) const {
    if (E && L.cube) {
      if (i == 0) {
        if (j == -1) return -__ldg(ua + LIDX(L, 2, 0));
        if (j == 0) return -__ldg(ua + LIDX(L, 1, 0));
        if (j == L.npy) return __ldg(ua + LIDX(L, 1, L.npy));
        if (j == L.npy + 1) return __ldg(ua + LIDX(L, 2, L.npy));
      } else if (i == L.npx) {
        if (j == 0) return __ldg(ua + LIDX(L, L.npx - 1, 0));
        if (j == -1) return __ldg(ua + LIDX(L, L.npx - 2, 0));
        if (j == L.npy) return -__ldg(ua + LIDX(L, L.npx - 1, L.npy));
        if (j == L.npy + 1) return -__ldg(ua + LIDX(L, L.npx - 2, L.npy));
      }
    }
    return __ldg(va + LIDX(L, i, j));
  }
};
}  // namespace

// ---- k_c: uc, ut, vc, vt (+ Courant scaling, sw_core.F90:159-176) and divergence_corner ----
// E = false: the point is far enough from the face edges (3 <= i <= npx-2, same in j) that no one-sided formula and no
// corner remap can apply -- the branches and the per-access index tests are compiled out (they were most of the
// instructions of this kernel: ncu, profiles/r1_stages_ncu_v5.txt)
template <bool E>
__device__ __forceinline__ void csw_c_point(const Lay& L, const DevGrid& G, const double* __restrict__ u, const double* __restrict__ v,
                                            const double* __restrict__ utmp_, const double* __restrict__ vtmp_, double* ua_, double* va_,
                                            double* __restrict__ uc, double* __restrict__ vc, double* __restrict__ ut,
                                            double* __restrict__ vt, double* __restrict__ divg_d, int nord, double dt2, int i, int j,
                                            long long ko) {
  const int npx = L.npx, npy = L.npy;
  const bool cube = E && L.cube;
  const bool ortho = !L.cube;   // doubly periodic orthogonal grid: vt = vc (sw_core.F90:3296-3303)
  const long long o = ko + LIDX(L, i, j);
  UtmpX<E> utx{utmp_ + ko, vtmp_ + ko, L};
  VtmpY<E> vty{utmp_ + ko, vtmp_ + ko, L};
  UaX<E> uax{ua_ + ko, va_ + ko, L};
  VaY<E> vay{ua_ + ko, va_ + ko, L};
  // x direction: j in [js-1, je+1], i in [is-1, ie+2]
  if (j >= L.js - 1 && j <= L.je + 1 && i >= L.is - 1 && i <= L.ie + 2) {
    double ucv, utv;
    if (cube && i == 1) {
      utv = edge_interpolate4(uax(-1, j), uax(0, j), uax(1, j), uax(2, j), G2(dxa, -1, j), G2(dxa, 0, j), G2(dxa, 1, j), G2(dxa, 2, j));
      ucv = (utv > 0.) ? utv * SG(3, 0, j) : utv * SG(1, 1, j);
    } else if (cube && i == npx) {
      utv = edge_interpolate4(uax(npx - 2, j), uax(npx - 1, j), uax(npx, j), uax(npx + 1, j), G2(dxa, npx - 2, j), G2(dxa, npx - 1, j),
                              G2(dxa, npx, j), G2(dxa, npx + 1, j));
      ucv = (utv > 0.) ? utv * SG(3, npx - 1, j) : utv * SG(1, npx, j);
    } else {
      if (cube && i == 0) ucv = c1 * utx(-2, j) + c2 * utx(-1, j) + c3 * utx(0, j);
      else if (cube && i == 2) ucv = c1 * utx(3, j) + c2 * utx(2, j) + c3 * utx(1, j);
      else if (cube && i == npx - 1) ucv = c1 * utx(npx - 3, j) + c2 * utx(npx - 2, j) + c3 * utx(npx - 1, j);
      else if (cube && i == npx + 1) ucv = c3 * utx(npx, j) + c2 * utx(npx + 1, j) + c1 * utx(npx + 2, j);
      else ucv = a2 * (utx(i - 2, j) + utx(i + 1, j)) + a1 * (utx(i - 1, j) + utx(i, j));
      utv = (ucv - AT(v, i, j) * G2(cosa_u, i, j)) * G2(rsin_u, i, j);
    }
    uc[o] = ucv;
    // sw_core.F90:159-167
    { const double s3 = SG(3, i - 1, j), s1 = SG(1, i, j), dyv = G2(dy, i, j);   // both candidates loaded before utv is known
      ut[o] = (utv > 0.) ? dt2 * utv * dyv * s3 : dt2 * utv * dyv * s1; }
  }
  // y direction: j in [js-1, je+2], i in [is-1, ie+1]
  if (j >= L.js - 1 && j <= L.je + 2 && i >= L.is - 1 && i <= L.ie + 1) {
    double vcv, vtv;
    if (ortho) {
      vcv = a2 * (vty(i, j - 2) + vty(i, j + 1)) + a1 * (vty(i, j - 1) + vty(i, j));
      vtv = vcv;
    } else if (cube && (j == 1 || j == npy)) {
      vtv = edge_interpolate4(vay(i, j - 2), vay(i, j - 1), vay(i, j), vay(i, j + 1), G2(dya, i, j - 2), G2(dya, i, j - 1), G2(dya, i, j),
                              G2(dya, i, j + 1));
      vcv = (vtv > 0.) ? vtv * SG(4, i, j - 1) : vtv * SG(2, i, j);
    } else {
      if (cube && (j == 0 || j == npy - 1)) vcv = c1 * vty(i, j - 2) + c2 * vty(i, j - 1) + c3 * vty(i, j);
      else if (cube && (j == 2 || j == npy + 1)) vcv = c1 * vty(i, j + 1) + c2 * vty(i, j) + c3 * vty(i, j - 1);
      else vcv = a2 * (vty(i, j - 2) + vty(i, j + 1)) + a1 * (vty(i, j - 1) + vty(i, j));
      vtv = (vcv - AT(u, i, j) * G2(cosa_v, i, j)) * G2(rsin_v, i, j);
    }
    vc[o] = vcv;
    // sw_core.F90:168-176
    { const double s4 = SG(4, i, j - 1), s2 = SG(2, i, j), dxv = G2(dx, i, j);
      vt[o] = (vtv > 0.) ? dt2 * vtv * dxv * s4 : dt2 * vtv * dxv * s2; }
  }
  // divergence_corner (sw_core.F90:1797-1843), corners [is, ie+1]^2; uses the un-rotated ua, va
  if (nord > 0 && i >= L.is && i <= L.ie + 1 && j >= L.js && j <= L.je + 1) {
    const double* uaq = ua_ + ko; const double* vaq = va_ + ko;
    auto UA = [&](int ii, int jj) { return __ldg(uaq + LIDX(L, ii, jj)); };
    auto VA = [&](int ii, int jj) { return __ldg(vaq + LIDX(L, ii, jj)); };
    if (L.grid_type > 3) {
      auto uf = [&](int ii, int jj) { return AT(u, ii, jj) * G2(dyc, ii, jj); };
      auto vf = [&](int ii, int jj) { return AT(v, ii, jj) * G2(dxc, ii, jj); };
      divg_d[o] = G2(rarea_c, i, j) * (vf(i, j - 1) - vf(i, j) + uf(i - 1, j) - uf(i, j));
    } else {
      auto uf = [&](int ii, int jj) {
        const double s = 0.5 * (SG(4, ii, jj - 1) + SG(2, ii, jj));
        if (E && (jj == 1 || jj == npy)) return AT(u, ii, jj) * G2(dyc, ii, jj) * s;
        return (AT(u, ii, jj) - 0.25 * (VA(ii, jj - 1) + VA(ii, jj)) * (CG(4, ii, jj - 1) + CG(2, ii, jj))) * G2(dyc, ii, jj) * s;
      };
      auto vf = [&](int ii, int jj) {
        const double s = 0.5 * (SG(3, ii - 1, jj) + SG(1, ii, jj));
        if (E && (ii == 1 || ii == npx)) return AT(v, ii, jj) * G2(dxc, ii, jj) * s;
        return (AT(v, ii, jj) - 0.25 * (UA(ii - 1, jj) + UA(ii, jj)) * (CG(3, ii - 1, jj) + CG(1, ii, jj))) * G2(dxc, ii, jj) * s;
      };
      double d = vf(i, j - 1) - vf(i, j) + uf(i - 1, j) - uf(i, j);
      if (E) {
        if (i == 1 && j == 1) d = d - vf(1, 0);
        if (i == npx && j == 1) d = d - vf(npx, 0);
        if (i == npx && j == npy) d = d + vf(npx, npy);
        if (i == 1 && j == npy) d = d + vf(1, npy);
      }
      divg_d[o] = G2(rarea_c, i, j) * d;
    }
  }
  // persist the rotated corner values of ua, va exactly as the reference leaves them
  // (sw_core.F90:3206-3221, :3279-3294): 16 cells per level.  No thread of this kernel reads
  // these cells directly (only through UaX/VaY, which remap them), so the in-place store is safe.
  if (cube) {
    const bool jx = (j == 0 || j == npy), ix = (i == -1 || i == 0 || i == npx || i == npx + 1);
    if (jx && ix) ua_[o] = uax(i, j);
    const bool iy = (i == 0 || i == npx), jy = (j == -1 || j == 0 || j == npy || j == npy + 1);
    if (iy && jy) va_[o] = vay(i, j);
  }
}
// The same for a point with 3 <= i <= npx-2, 3 <= j <= npy-2 of a cubed-sphere face (grid_type <= 3), written with plain offsets
// from ONE in-plane index: the generic routine computes an index (and, through its accessors, the remap tests the E = false
// instantiation folds away only partly) for each of its ~55 loads -- 586 warp instructions per thread (profiles/r2_dsw_summary.md).
// Same operations in the same order.
__device__ __forceinline__ void csw_c_interior(const Lay& L, const DevGrid& G, const double* __restrict__ u, const double* __restrict__ v,
                                               const double* __restrict__ utmp, const double* __restrict__ vtmp, const double* __restrict__ ua,
                                               const double* __restrict__ va, double* __restrict__ uc, double* __restrict__ vc,
                                               double* __restrict__ ut, double* __restrict__ vt, double* __restrict__ divg_d, int nord,
                                               double dt2, int o2, long long ko) {
  const int NI = L.NI;
  const long long P = L.plane, o = ko + o2;
  const double* sg = G.sin_sg + o2; const double* cg = G.cos_sg + o2;
  const double* up = u + o; const double* vp = v + o;
  const double u00 = __ldg(up), v00 = __ldg(vp);
  {
    const double* t = utmp + o;
    const double ucv = a2 * (__ldg(t - 2) + __ldg(t + 1)) + a1 * (__ldg(t - 1) + __ldg(t));
    const double utv = (ucv - v00 * __ldg(G.cosa_u + o2)) * __ldg(G.rsin_u + o2);
    const double s3 = __ldg(sg + 2 * P - 1), s1 = __ldg(sg), dyv = __ldg(G.dy + o2);
    uc[o] = ucv;
    ut[o] = (utv > 0.) ? dt2 * utv * dyv * s3 : dt2 * utv * dyv * s1;
  }
  {
    const double* t = vtmp + o;
    const double vcv = a2 * (__ldg(t - 2 * NI) + __ldg(t + NI)) + a1 * (__ldg(t - NI) + __ldg(t));
    const double vtv = (vcv - u00 * __ldg(G.cosa_v + o2)) * __ldg(G.rsin_v + o2);
    const double s4 = __ldg(sg + 3 * P - NI), s2 = __ldg(sg + P), dxv = __ldg(G.dx + o2);
    vc[o] = vcv;
    vt[o] = (vtv > 0.) ? dt2 * vtv * dxv * s4 : dt2 * vtv * dxv * s2;
  }
  if (nord > 0) {   // divergence_corner (sw_core.F90:1797-1843) at corner (i, j): uf at (i-1, j), (i, j); vf at (i, j-1), (i, j)
    const double* uap = ua + o; const double* vap = va + o;
    // uf(ii,jj) = (u - 0.25*(va(jj-1)+va(jj))*(cos_sg4(jj-1)+cos_sg2(jj))) * dyc * (0.5*(sin_sg4(jj-1)+sin_sg2(jj)))
    auto uf = [&](int d) {   // d = 0: (i, j), d = -1: (i-1, j)
      const double s = 0.5 * (__ldg(sg + 3 * P + d - NI) + __ldg(sg + P + d));
      return (__ldg(up + d) - 0.25 * (__ldg(vap + d - NI) + __ldg(vap + d)) * (__ldg(cg + 3 * P + d - NI) + __ldg(cg + P + d))) * __ldg(G.dyc + o2 + d) * s;
    };
    auto vf = [&](int d) {   // d = 0: (i, j), d = -NI: (i, j-1)
      const double s = 0.5 * (__ldg(sg + 2 * P + d - 1) + __ldg(sg + d));
      return (__ldg(vp + d) - 0.25 * (__ldg(uap + d - 1) + __ldg(uap + d)) * (__ldg(cg + 2 * P + d - 1) + __ldg(cg + d))) * __ldg(G.dxc + o2 + d) * s;
    };
    const double dv = vf(-NI) - vf(0) + uf(-1) - uf(0);
    divg_d[o] = __ldg(G.rarea_c + o2) * dv;
  }
}
#ifndef CSWC_MINB
#define CSWC_MINB 4   // 64 registers: at 6 CTAs per SM (40 registers) this kernel spills 140 bytes (measured 440 / 457 / 491 us at 4 / 6 / 3)
#endif
__global__ void __launch_bounds__(TI* TJ, CSWC_MINB) k_csw_c(Lay L, DevGrid G, const double* __restrict__ u, const double* __restrict__ v,
                                                 const double* __restrict__ utmp_, const double* __restrict__ vtmp_,
                                                 double* ua_, double* va_,
                                                 double* __restrict__ uc, double* __restrict__ vc, double* __restrict__ ut,
                                                 double* __restrict__ vt, double* __restrict__ divg_d, int nord, double dt2) {
  PLANE_IJK
  if (i < L.isd || i > L.ied + 1 || j > L.jed + 1) return;
  if (L.cube && L.grid_type <= 3 && i >= 3 && i <= L.npx - 2 && j >= 3 && j <= L.npy - 2)
    csw_c_interior(L, G, u, v, utmp_, vtmp_, ua_, va_, uc, vc, ut, vt, divg_d, nord, dt2, LIDX(L, i, j), ko);
  else
    csw_c_point<true>(L, G, u, v, utmp_, vtmp_, ua_, va_, uc, vc, ut, vt, divg_d, nord, dt2, i, j, ko);
}

// ---- k_t: delpc, ptc, wc, ke, vort --------------------------------------------------------
template <bool E>
__device__ __forceinline__ void csw_t_point(const Lay& L, const DevGrid& G, const double* __restrict__ delp, const double* __restrict__ pt,
                                            const double* __restrict__ w, const double* __restrict__ u, const double* __restrict__ v,
                                            const double* __restrict__ uc, const double* __restrict__ vc, const double* __restrict__ ua,
                                            const double* __restrict__ va, const double* __restrict__ ut, const double* __restrict__ vt,
                                            double* __restrict__ delpc, double* __restrict__ ptc, double* __restrict__ wc,
                                            double* __restrict__ ke, double* __restrict__ vort, int hydrostatic, double dt2, int i, int j,
                                            long long ko) {
  const int npx = L.npx, npy = L.npy;
  const bool cube = E && L.cube;
  const bool ortho = !L.cube;
  const long long o = ko + LIDX(L, i, j);
  FillX<E> dx_{delp + ko, L}, px_{pt + ko, L}, wx_{w + ko, L};
  FillY<E> dy_{delp + ko, L}, py_{pt + ko, L}, wy_{w + ko, L};
  // upwind fluxes through the 4 faces (sw_core.F90:214-276)
  double fx1[2], fx[2], fx2[2], fy1[2], fy[2], fy2[2];
  if (!E) {
    // interior points: BOTH upwind candidates of every face are loaded before the winds are known (the 5-point cross of delp,
    // pt, w: 15 + 4 independent loads) instead of 12 loads whose addresses depend on the sign of ut / vt -- the dependent form
    // costs two memory round trips per thread and the kernel was load-latency bound (ncu long_scoreboard 12.9 cycles per issue
    // at 40 % of the DRAM bandwidth, profiles/r2/r2_substep_limiters.txt)
    const double* dp_ = delp + o; const double* pt_ = pt + o; const double* w_ = w + o;
    const int NI = L.NI;
    const double u0 = __ldg(ut + o), u1 = __ldg(ut + o + 1), v0 = __ldg(vt + o), v1 = __ldg(vt + o + NI);
    const double dW = __ldg(dp_ - 1), dC = __ldg(dp_), dE = __ldg(dp_ + 1), dS = __ldg(dp_ - NI), dN = __ldg(dp_ + NI);
    const double pW = __ldg(pt_ - 1), pC = __ldg(pt_), pE = __ldg(pt_ + 1), pS = __ldg(pt_ - NI), pN = __ldg(pt_ + NI);
    double wW = 0., wC = 0., wE = 0., wS = 0., wN = 0.;
    if (!hydrostatic) { wW = __ldg(w_ - 1); wC = __ldg(w_); wE = __ldg(w_ + 1); wS = __ldg(w_ - NI); wN = __ldg(w_ + NI); }
    fx1[0] = u0 * (u0 > 0. ? dW : dC); fx[0] = fx1[0] * (u0 > 0. ? pW : pC); fx2[0] = hydrostatic ? 0. : fx1[0] * (u0 > 0. ? wW : wC);
    fx1[1] = u1 * (u1 > 0. ? dC : dE); fx[1] = fx1[1] * (u1 > 0. ? pC : pE); fx2[1] = hydrostatic ? 0. : fx1[1] * (u1 > 0. ? wC : wE);
    fy1[0] = v0 * (v0 > 0. ? dS : dC); fy[0] = fy1[0] * (v0 > 0. ? pS : pC); fy2[0] = hydrostatic ? 0. : fy1[0] * (v0 > 0. ? wS : wC);
    fy1[1] = v1 * (v1 > 0. ? dC : dN); fy[1] = fy1[1] * (v1 > 0. ? pC : pN); fy2[1] = hydrostatic ? 0. : fy1[1] * (v1 > 0. ? wC : wN);
  } else
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int ii = i + s;
    const double utv = AT(ut, ii, j);
    const int iu = (utv > 0.) ? ii - 1 : ii;
    fx1[s] = utv * dx_(iu, j);
    fx[s] = fx1[s] * px_(iu, j);
    fx2[s] = hydrostatic ? 0. : fx1[s] * wx_(iu, j);
    const int jj = j + s;
    const double vtv = AT(vt, i, jj);
    const int ju = (vtv > 0.) ? jj - 1 : jj;
    fy1[s] = vtv * dy_(i, ju);
    fy[s] = fy1[s] * py_(i, ju);
    fy2[s] = hydrostatic ? 0. : fy1[s] * wy_(i, ju);
  }
  const double ra = G2(rarea, i, j);
  const double dp0 = dy_(i, j), pt0 = py_(i, j);
  const double dpc = dp0 + (fx1[0] - fx1[1] + fy1[0] - fy1[1]) * ra;
  delpc[o] = dpc;
  const double rdpc = 1. / dpc;   // one division for ptc and wc (<= 1 ulp from the two of sw_core.F90:279-283)
  ptc[o] = (pt0 * dp0 + (fx[0] - fx[1] + fy[0] - fy[1]) * ra) * rdpc;
  if (!hydrostatic) wc[o] = (wy_(i, j) * dp0 + (fx2[0] - fx2[1] + fy2[0] - fy2[1]) * ra) * rdpc;

  // KE (sw_core.F90:297-366)
  const double uav = AT(ua, i, j), vav = AT(va, i, j);
  double kx, ky;
  (void)ortho;
  if (!cube) {   // no face edge nearby (or not a cubed sphere): sw_core.F90:352-366; both candidates loaded, then selected
    const double uc0 = AT(uc, i, j), uc1 = AT(uc, i + 1, j), vc0 = AT(vc, i, j), vc1 = AT(vc, i, j + 1);
    kx = (uav > 0.) ? uc0 : uc1;
    ky = (vav > 0.) ? vc0 : vc1;
  } else {
    if (uav > 0.) {
      if (i == 1) kx = AT(uc, 1, j) * SG(1, 1, j) + AT(v, 1, j) * CG(1, 1, j);
      else if (i == npx) kx = AT(uc, npx, j) * SG(1, npx, j) + AT(v, npx, j) * CG(1, npx, j);
      else kx = AT(uc, i, j);
    } else {
      if (i == 0) kx = AT(uc, 1, j) * SG(3, 0, j) + AT(v, 1, j) * CG(3, 0, j);
      else if (i == npx - 1) kx = AT(uc, npx, j) * SG(3, npx - 1, j) + AT(v, npx, j) * CG(3, npx - 1, j);
      else kx = AT(uc, i + 1, j);
    }
    if (vav > 0.) {
      if (j == 1) ky = AT(vc, i, 1) * SG(2, i, 1) + AT(u, i, 1) * CG(2, i, 1);
      else if (j == npy) ky = AT(vc, i, npy) * SG(2, i, npy) + AT(u, i, npy) * CG(2, i, npy);
      else ky = AT(vc, i, j);
    } else {
      if (j == 0) ky = AT(vc, i, 1) * SG(4, i, 0) + AT(u, i, 1) * CG(4, i, 0);
      else if (j == npy - 1) ky = AT(vc, i, npy) * SG(4, i, npy - 1) + AT(u, i, npy) * CG(4, i, npy - 1);
      else ky = AT(vc, i, j + 1);
    }
  }
  const double dt4 = 0.5 * dt2;
  ke[o] = dt4 * (uav * kx + vav * ky);

  // absolute vorticity at corners [is, ie+1]x[js, je+1] (sw_core.F90:372-403)
  if (i >= L.is && j >= L.js) {
    auto FX = [&](int ii, int jj) { return AT(uc, ii, jj) * G2(dxc, ii, jj); };
    auto FY = [&](int ii, int jj) { return AT(vc, ii, jj) * G2(dyc, ii, jj); };
    double vo = FX(i, j - 1) - FX(i, j) - FY(i - 1, j) + FY(i, j);
    if (cube) {
      if (i == 1 && j == 1) vo = vo + FY(0, 1);
      if (i == npx && j == 1) vo = vo - FY(npx, 1);
      if (i == npx && j == npy) vo = vo - FY(npx, npy);
      if (i == 1 && j == npy) vo = vo + FY(0, npy);
    }
    vort[o] = G2(fC, i, j) + G2(rarea_c, i, j) * vo;
  }
}
__global__ void __launch_bounds__(TI* TJ, PLANE_MINB) k_csw_t(Lay L, DevGrid G, const double* __restrict__ delp, const double* __restrict__ pt,
                                                 const double* __restrict__ w, const double* __restrict__ u, const double* __restrict__ v,
                                                 const double* __restrict__ uc, const double* __restrict__ vc, const double* __restrict__ ua,
                                                 const double* __restrict__ va, const double* __restrict__ ut, const double* __restrict__ vt,
                                                 double* __restrict__ delpc, double* __restrict__ ptc, double* __restrict__ wc,
                                                 double* __restrict__ ke, double* __restrict__ vort, int hydrostatic, double dt2) {
  PLANE_IJK
  if (i < L.is - 1 || i > L.ie + 1 || j < L.js - 1 || j > L.je + 1) return;
  if (L.cube && i >= 3 && i <= L.npx - 2 && j >= 3 && j <= L.npy - 2)
    csw_t_point<false>(L, G, delp, pt, w, u, v, uc, vc, ua, va, ut, vt, delpc, ptc, wc, ke, vort, hydrostatic, dt2, i, j, ko);
  else
    csw_t_point<true>(L, G, delp, pt, w, u, v, uc, vc, ua, va, ut, vt, delpc, ptc, wc, ke, vort, hydrostatic, dt2, i, j, ko);
}

// ---- k_u: time-centred uc, vc (sw_core.F90:414-486) ---------------------------------------
__global__ void __launch_bounds__(TI* TJ) k_csw_u(Lay L, DevGrid G, const double* __restrict__ u, const double* __restrict__ v,
                                                 const double* __restrict__ ke, const double* __restrict__ vort,
                                                 double* __restrict__ uc, double* __restrict__ vc, double dt2) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const int npx = L.npx, npy = L.npy;
  const bool cube = L.cube;
  const long long o = ko + LIDX(L, i, j);
  if (j <= L.je) {
    double fy1;
    if (cube && (i == 1 || i == npx)) fy1 = dt2 * AT(v, i, j);
    else fy1 = dt2 * (AT(v, i, j) - uc[o] * G2(cosa_u, i, j)) / G2(sina_u, i, j);
    const double vo0 = AT(vort, i, j), vo1 = AT(vort, i, j + 1);
    const double fyv = (fy1 > 0.) ? vo0 : vo1;
    uc[o] = uc[o] + fy1 * fyv + G2(rdxc, i, j) * (AT(ke, i - 1, j) - AT(ke, i, j));
  }
  if (i <= L.ie) {
    double fx1;
    if (cube && (j == 1 || j == npy)) fx1 = dt2 * AT(u, i, j);
    else fx1 = dt2 * (AT(u, i, j) - vc[o] * G2(cosa_v, i, j)) / G2(sina_v, i, j);
    const double vo0 = AT(vort, i, j), vo1 = AT(vort, i + 1, j);
    const double fxv = (fx1 > 0.) ? vo0 : vo1;
    vc[o] = vc[o] - fx1 * fxv + G2(rdyc, i, j) * (AT(ke, i, j - 1) - AT(ke, i, j));
  }
}

int stage_c_sw(fv3_ctx* c, double dt2) {
  StageScope ts(c, "C_SW");
  const Lay& L = c->L;
  const int nk = L.npz;
  dim3 blk(TI, TJ), grd = plane_grid(L, nk);
  double *utmp = c->scr[0], *vtmp = c->scr[1], *ke = c->scr[2], *vort = c->scr[3];
  k_csw_a<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_U], c->fld[FV3_V], utmp, vtmp, c->fld[FV3_UA], c->fld[FV3_VA]);
  k_csw_c<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_U], c->fld[FV3_V], utmp, vtmp, c->fld[FV3_UA], c->fld[FV3_VA],
                                      c->fld[FV3_UC], c->fld[FV3_VC], c->fld[FV3_UT], c->fld[FV3_VT], c->fld[FV3_DIVGD],
                                      c->f.nord, dt2);
  k_csw_t<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_DELP], c->fld[FV3_PT], c->fld[FV3_W], c->fld[FV3_U], c->fld[FV3_V],
                                      c->fld[FV3_UC], c->fld[FV3_VC], c->fld[FV3_UA], c->fld[FV3_VA], c->fld[FV3_UT],
                                      c->fld[FV3_VT], c->fld[FV3_DELPC], c->fld[FV3_PTC], c->fld[FV3_OMGA], ke, vort,
                                      c->f.hydrostatic, dt2);
  k_csw_u<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_U], c->fld[FV3_V], ke, vort, c->fld[FV3_UC], c->fld[FV3_VC], dt2);
  c->launches += 4;
  return 0;
}
