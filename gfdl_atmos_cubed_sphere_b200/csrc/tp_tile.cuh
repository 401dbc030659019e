// fv_tp_2d on a shared-memory tile: inner sweeps -> q_i / q_j -> outer sweeps in ONE kernel.
//
// Reference semantics: model/tp_core.F90:85-241 (fv_tp_2d), :245-322 (copy_corners).
// Why: in the three-launch form (first version) the intermediates fx2, fy2, q_i, q_j made a round trip through
// HBM/L2 and every 6-point window was re-fetched through L1 by six different threads (ncu,
// profiles/r1_dsw_ncu_summary.md).  Here a CTA of 512 threads owns a TX x TY block of cells of one level:
//   I. stage the Courant numbers, area fluxes and cell areas of the tile with cp.async      (once per tile)
//   A. stage the (TY+6) x (TX+6) halo tile of q (cube-corner tiles: both copy_corners views) (per field)
//   B0. per-point limiter input (dm or al) of both sweeps
//   B. inner fluxes fx2 on (TY+6) x (TX+1), fy2 on (TY+1) x (TX+6)          [ord_in]
//   C. q_i on TY x (TX+6), q_j on (TY+6) x TX                                 [one division each]
//   C2. per-point limiter input of q_i (x lines) and q_j (y lines)
//   D. outer fluxes on TY x (TX+1) and (TY+1) x TX, averaged with the inner   [ord_ou]
// All global reads happen in I/A with every load of the tile in flight at once (the first fused version read crx,
// xfx, area inside B, C, D and spent 45 % of its issue slots stalled on those loads: ncu long_scoreboard).
// The recomputed halo fluxes cost ~1.3x arithmetic.  The caller's epilogue weights the fluxes (area or mass
// flux) and, in d_sw, applies the flux divergence in place, so the fluxes of w, pt, q_con never leave the SM
// and the staged Courant numbers / area fluxes are shared by all transported fields.
#pragma once
#include "ppm.cuh"

namespace tpt {

constexpr int TX = 32, TY = 16, NT = 512;
constexpr int QW = TX + 6, QH = TY + 6;

struct Smem {
  // per tile
  double crx[QH][TX + 1];   // crx(i0+c, j0-3+r)
  double xfx[QH][TX + 1];
  double cry[TY + 1][QW];   // cry(i0-3+c, j0+r)
  double yfx[TY + 1][QW];
  double area[QH][QW];      // area(i0-3+c, j0-3+r)
  // per field
  double qx[QH][QW];        // q, copy_corners(dir=1) view
  double ax[QH][QW];        // x-sweep limiter input (dm or al, ppm::aux_point) of q, later of q_i
  double ay[QH][QW];        // y-sweep limiter input of q, later of q_j
  double fx2[QH][TX + 1];   // inner x flux; rows 3..TY+2 hold the averaged outer flux after phase D
  double fy2[TY + 1][QW];   // inner y flux; columns 3..TX+2 hold the averaged outer flux after phase D
  double w[TY * QW + QH * TX];  // phases A-B: q in the copy_corners(dir=2) view (cube-corner tiles); C-D: q_i, q_j
};
__device__ __forceinline__ double& QI(Smem& S, int r, int c) { return S.w[r * QW + c]; }             // [TY][QW]
__device__ __forceinline__ double& QJ(Smem& S, int r, int c) { return S.w[TY * QW + r * TX + c]; }   // [QH][TX]
__device__ __forceinline__ double& FX(Smem& S, int r, int c) { return S.fx2[r + 3][c]; }   // r < TY, c <= TX
__device__ __forceinline__ double& FY(Smem& S, int r, int c) { return S.fy2[r][c + 3]; }   // r <= TY, c < TX

// strided view of a shared-memory line in sweep coordinates: value at sweep index s
struct SAcc {
  const double* p; int stride; int org;
  __device__ __forceinline__ double operator()(int s) const { return p[(s - org) * stride]; }
};

// loop over the points (r, c) of an N = rows x W array with NT threads: compile-time trip count (fully unrolled),
// no integer division in the body
#define TPT_LOOP(N, W, r, c)                                                                        \
  _Pragma("unroll") for (int it_ = 0, t_ = threadIdx.x, r = t_ / (W), c = t_ - r * (W);            \
                         it_ < ((N) + NT - 1) / NT;                                                 \
                         ++it_, t_ += NT, r += NT / (W), c += NT % (W), r += (c >= (W)) ? 1 : 0, c -= (c >= (W)) ? (W) : 0) \
    if (t_ < (N))

// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, all copies of a tile in flight together
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// cube-edge faces (the one-sided formulas of tp_core.F90:643-681): rare, kept out of line so the interior
// path stays small (instruction cache) and its register allocation is not driven by this code
static __device__ __noinline__ double edge_flux(const double* ql, int sq, int org, const double* dl, long long dbase,
                                                long long dstride, int i, double c, int iord, int n) {
  SAcc qa{ql, sq, org};
  ppm::Acc da{dl, dbase, dstride};
  return ppm::flux_scalar(qa, da, i, c, iord, n, true);
}

// interior flux through face f of a line (face f lies between line elements f+2 and f+3)
template <bool MONO>
__device__ __forceinline__ double line_flux(const double* ql, int sq, const double* al, int sa, int f, double c, int iord) {
  if (MONO) {
    const int u = (c > 0.) ? f + 2 : f + 3;
    const double* p = ql + u * sq;
    const double* d = al + u * sa;
    return ppm::flux_mono_aux(p[-2 * sq], p[-sq], p[0], p[sq], p[2 * sq], d[-sa], d[0], d[sa], c, iord);
  }
  const double* p = ql + (f + 2) * sq;
  const double* d = al + (f + 2) * sa;
  return ppm::flux_unlim_aux(p[0], p[sq], d[0], d[sa], d[2 * sa], c, iord);
}

struct Tile {
  int i0, j0;          // first cell of the tile
  long long ko;        // level offset (elements)
  int NI, ib, jb;      // in-plane index = (i + ib) + (j - jb) * NI
  bool corner;         // tile halo overlaps a cube-corner region: q needs the two copy_corners views
  __device__ __forceinline__ int idx(int i, int j) const { return (i + ib) + (j - jb) * NI; }
};
__device__ __forceinline__ Tile make_tile(const Lay& L) {
  Tile T;
  T.i0 = L.is + blockIdx.x * TX; T.j0 = L.js + blockIdx.y * TY;
  T.ko = (long long)blockIdx.z * L.plane;
  T.NI = L.NI; T.ib = FV3_IOFF - L.isd; T.jb = L.jsd;
  T.corner = L.cube && (T.i0 - 3 <= 0 || T.i0 + TX + 2 >= L.npx) && (T.j0 - 3 <= 0 || T.j0 + TY + 2 >= L.npy);
  return T;
}

// I: issue the per-tile inputs (no wait).  Points beyond the face are clamped duplicates.
__device__ __forceinline__ void stage_inputs(const Lay& L, const DevGrid& G, Smem& S, const Tile& T,
                                             const double* __restrict__ crx, const double* __restrict__ cry,
                                             const double* __restrict__ xfx, const double* __restrict__ yfx) {
  crx += T.ko; cry += T.ko; xfx += T.ko; yfx += T.ko;
  TPT_LOOP(QH * (TX + 1), TX + 1, r, c) {
    const int o = T.idx(min(T.i0 + c, L.ie + 1), min(T.j0 - 3 + r, L.jed));
    cp_async8(&S.crx[r][c], crx + o);
    cp_async8(&S.xfx[r][c], xfx + o);
  }
  TPT_LOOP((TY + 1) * QW, QW, r, c) {
    const int o = T.idx(min(T.i0 - 3 + c, L.ied), min(T.j0 + r, L.je + 1));
    cp_async8(&S.cry[r][c], cry + o);
    cp_async8(&S.yfx[r][c], yfx + o);
  }
  TPT_LOOP(QH * QW, QW, r, c) {
    const int o = T.idx(min(T.i0 - 3 + c, L.ied), min(T.j0 - 3 + r, L.jed));
    cp_async8(&S.area[r][c], G.area + o);
  }
}

// A: issue the halo tile of q (no wait)
__device__ __forceinline__ void stage_q(const Lay& L, Smem& S, const Tile& T, const double* __restrict__ q) {
  q += T.ko;
  double(*qyw)[QW] = reinterpret_cast<double(*)[QW]>(S.w);
  TPT_LOOP(QH * QW, QW, r, c) {
    const int i = min(T.i0 - 3 + c, L.ied), j = min(T.j0 - 3 + r, L.jed);
    if (T.corner) {
      S.qx[r][c] = ppm::QAccX{q, L, j}(i);
      qyw[r][c] = ppm::QAccY{q, L, i}(j);
    } else {
      cp_async8(&S.qx[r][c], q + T.idx(i, j));
    }
  }
}

// B0..D for the staged tile.  The caller has issued stage_inputs/stage_q; this waits for them.
// On return (after a __syncthreads) FX(S,r,c) = unweighted Lin-Rood flux 0.5*(outer + inner) through the west face of
// cell (i0+c, j0+r), FY(S,r,c) = south face.  Tile points beyond the face are garbage; the caller must not store them.
// MONO: the scheme family (8, 10 vs 5, 6, -5) is a compile-time choice of the launcher (both sweeps of one hord are in
// the same family: hord 10 runs ord 8 inside, tp_core.F90:136-141).
template <bool MONO>
__device__ __forceinline__ void tp_compute(const Lay& L, const DevGrid& G, Smem& S, const Tile& T,
                                           const double* __restrict__ ra_x, const double* __restrict__ ra_y,
                                           int ord_in, int ord_ou) {
  using namespace ppm;
  const bool cube = L.cube;
  const int npx = L.npx, npy = L.npy, i0 = T.i0, j0 = T.j0;
  cp_async_wait_all();
  __syncthreads();
  const double(*qy)[QW] = T.corner ? reinterpret_cast<const double(*)[QW]>(S.w) : S.qx;
  // ---- B0: per-point limiter inputs of both sweeps (border points are clamped garbage that no face reads)
  TPT_LOOP(QH * QW, QW, r, c) {
    const int c2 = max(c - 2, 0), c1 = max(c - 1, 0), c3 = min(c + 1, QW - 1);
    S.ax[r][c] = aux_point<MONO>(ord_in, S.qx[r][c2], S.qx[r][c1], S.qx[r][c], S.qx[r][c3]);
    const int r2 = max(r - 2, 0), r1 = max(r - 1, 0), r3 = min(r + 1, QH - 1);
    S.ay[r][c] = aux_point<MONO>(ord_in, qy[r2][c], qy[r1][c], qy[r][c], qy[r3][c]);
  }
  __syncthreads();
  // ---- B: inner sweeps
  TPT_LOOP(QH * (TX + 1), TX + 1, r, c) {   // fx2(i0+c, j0-3+r)   tp_core.F90:164-169
    const int i = i0 + c;
    const double cr = S.crx[r][c];
    double f;
    if (!cube || (i >= 4 && i <= npx - 3)) f = line_flux<MONO>(&S.qx[r][0], 1, &S.ax[r][0], 1, c, cr, ord_in);
    else f = edge_flux(&S.qx[r][0], 1, i0 - 3, G.dxa, LIDX(L, 0, min(j0 - 3 + r, L.jed)), 1, min(i, L.ie + 1), cr, ord_in, npx);
    S.fx2[r][c] = f;
  }
  TPT_LOOP((TY + 1) * QW, QW, r, c) {       // fy2(i0-3+c, j0+r)   tp_core.F90:143-148
    const int j = j0 + r;
    const double cr = S.cry[r][c];
    double f;
    if (!cube || (j >= 4 && j <= npy - 3)) f = line_flux<MONO>(&qy[0][c], QW, &S.ay[0][c], QW, r, cr, ord_in);
    else f = edge_flux(&qy[0][c], QW, j0 - 3, G.dya, LIDX(L, min(i0 - 3 + c, L.ied), 0), L.NI, min(j, L.je + 1), cr, ord_in, npy);
    S.fy2[r][c] = f;
  }
  __syncthreads();
  // ---- C: intermediate advected fields
  TPT_LOOP(TY * QW, QW, r, c) {             // q_i(i0-3+c, j0+r)   tp_core.F90:150-159
    const double ar = S.area[r + 3][c];
    const double y0 = S.yfx[r][c], y1 = S.yfx[r + 1][c];
    const double f0 = y0 * S.fy2[r][c], f1 = y1 * S.fy2[r + 1][c];
    const double ray = ra_y ? __ldg(ra_y + T.ko + T.idx(min(i0 - 3 + c, L.ied), min(j0 + r, L.je))) : (ar + y0 - y1);
    QI(S, r, c) = (S.qx[r + 3][c] * ar + f0 - f1) / ray;
  }
  TPT_LOOP(QH * TX, TX, r, c) {             // q_j(i0+c, j0-3+r)   tp_core.F90:171-178
    const double ar = S.area[r][c + 3];
    const double x0 = S.xfx[r][c], x1 = S.xfx[r][c + 1];
    const double f0 = x0 * S.fx2[r][c], f1 = x1 * S.fx2[r][c + 1];
    const double rax = ra_x ? __ldg(ra_x + T.ko + T.idx(min(i0 + c, L.ie), min(j0 - 3 + r, L.jed))) : (ar + x0 - x1);
    QJ(S, r, c) = (S.qx[r][c + 3] * ar + f0 - f1) / rax;
  }
  __syncthreads();
  // ---- C2: limiter inputs of q_i (x lines) and q_j (y lines)
  TPT_LOOP(TY * QW, QW, r, c) {
    const int c2 = max(c - 2, 0), c1 = max(c - 1, 0), c3 = min(c + 1, QW - 1);
    S.ax[r][c] = aux_point<MONO>(ord_ou, QI(S, r, c2), QI(S, r, c1), QI(S, r, c), QI(S, r, c3));
  }
  TPT_LOOP(QH * TX, TX, r, c) {
    const int r2 = max(r - 2, 0), r1 = max(r - 1, 0), r3 = min(r + 1, QH - 1);
    S.ay[r][c] = aux_point<MONO>(ord_ou, QJ(S, r2, c), QJ(S, r1, c), QJ(S, r, c), QJ(S, r3, c));
  }
  __syncthreads();
  // ---- D: outer sweeps, averaged with the inner fluxes in place (tp_core.F90:161,180,193-198)
  TPT_LOOP(TY * (TX + 1), TX + 1, r, c) {
    const int i = i0 + c;
    const double cr = S.crx[r + 3][c];
    double f;
    if (!cube || (i >= 4 && i <= npx - 3)) f = line_flux<MONO>(&QI(S, r, 0), 1, &S.ax[r][0], 1, c, cr, ord_ou);
    else f = edge_flux(&QI(S, r, 0), 1, i0 - 3, G.dxa, LIDX(L, 0, min(j0 + r, L.je)), 1, min(i, L.ie + 1), cr, ord_ou, npx);
    FX(S, r, c) = 0.5 * (f + FX(S, r, c));
  }
  TPT_LOOP((TY + 1) * TX, TX, r, c) {
    const int j = j0 + r;
    const double cr = S.cry[r][c + 3];
    double f;
    if (!cube || (j >= 4 && j <= npy - 3)) f = line_flux<MONO>(&QJ(S, 0, c), TX, &S.ay[0][c], QW, r, cr, ord_ou);
    else f = edge_flux(&QJ(S, 0, c), TX, j0 - 3, G.dya, LIDX(L, min(i0 + c, L.ie), 0), L.NI, min(j, L.je + 1), cr, ord_ou, npy);
    FY(S, r, c) = 0.5 * (f + FY(S, r, c));
  }
  __syncthreads();
}

static inline dim3 tile_grid(const Lay& L, int nk) {
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  return dim3((nx + TX - 1) / TX, (ny + TY - 1) / TY, nk);
}

}  // namespace tpt
