// fv_tp_2d on a shared-memory tile: inner sweeps -> q_i / q_j -> outer sweeps in ONE kernel.
//
// Reference semantics: model/tp_core.F90:85-241 (fv_tp_2d), :245-322 (copy_corners).
// Why: in the three-launch form (first version) the intermediates fx2, fy2, q_i, q_j made a round trip through
// HBM/L2 and every 6-point window was re-fetched through L1 by six different threads (ncu,
// profiles/r1_dsw_ncu_summary.md).  Here a CTA of 16 warps owns a TX x TY = 26 x 26 block of cells of one level.
// Every tile array is [QH][32]: the 32 lanes of a warp are the 26 cells + 3 + 3 halo columns of a row, a warp takes
// whole rows (row index is warp-uniform, column = lane, no index arithmetic and no integer division in the body):
//   I.  stage the Courant numbers, area fluxes and cell areas of the tile with cp.async       (once per tile)
//   A.  stage the halo tile of q (cube-corner tiles: both copy_corners views)                 (per field)
//   B0. per-point limiter input (dm or al) of both sweeps                     [-DTPT_AUX only]
//   B.  inner fluxes fx2 (all rows), fy2 (rows 3..TY+3)                       [ord_in]
//   C.  q_i (rows 3..TY+2, all columns), q_j (all rows, columns 3..TX+2)      [one division each]
//   C2. per-point limiter input of q_i (x lines) and q_j (y lines)            [-DTPT_AUX only]
//   D.  outer fluxes, averaged with the inner ones in place                   [ord_ou]
// All global reads happen in I/A with every load of the tile in flight at once (the first fused version read crx,
// xfx, area inside B, C, D and spent 45 % of its issue slots stalled on those loads: ncu long_scoreboard).
// The recomputed halo fluxes cost ~1.3x arithmetic.  The caller's epilogue weights the fluxes (area or mass
// flux) and, in d_sw, applies the flux divergence in place, so the fluxes of w, pt, q_con never leave the SM
// and the staged Courant numbers / area fluxes are shared by all transported fields.
//
// Tile coordinates: array element [r][c] is the point (i0-3+c, j0-3+r); a flux stored at [r][c] is the flux through
// the WEST (x sweeps) or SOUTH (y sweeps) face of that cell, i.e. interface index i = i0-3+c (j = j0-3+r).
#pragma once
#include "ppm.cuh"

namespace tpt {

// Two forms of the limiter inputs, both parity-green (profiles/r1_dsw_ncu_summary.md, v10):
//   default    : recomputed in registers from the q window the flux loads anyway (TPT_NOAUX) -- 12 tile arrays, which lets
//                the tile be 26 rows high at 2 CTAs per SM (row tasks per phase 59 / 58 / 53 on 16 warps)
//   -DTPT_AUX  : one auxiliary value per point in shared memory (two extra passes + 2 barriers per field); 14 arrays, so
//                at most 24 rows at 2 CTAs per SM.  Measured at C384L79: d_sw 3.76 ms (AUX, 24) / 3.77 (NOAUX, 24) /
//                3.61 (NOAUX, 26) / 3.65 (NOAUX, 28)
#ifndef TPT_AUX
#define TPT_NOAUX
#endif
#ifndef TPT_TY
#ifdef TPT_NOAUX
#define TPT_TY 26
#else
#define TPT_TY 24
#endif
#endif
#ifndef TPT_NW
#define TPT_NW 16
#endif
constexpr int TX = 26, TY = TPT_TY, NW = TPT_NW, NT = NW * 32;
constexpr int QW = 32, QH = TY + 6;

struct Smem {
  // per tile
  double crx[QH][QW], xfx[QH][QW], cry[QH][QW], yfx[QH][QW], area[QH][QW];
  // per field
  double q[QH][QW];     // q, copy_corners(dir=1) view
#ifndef TPT_NOAUX
  double ax[QH][QW];    // x-sweep limiter input (dm or al, ppm::aux_point) of q, later of q_i
  double ay[QH][QW];    // y-sweep limiter input of q, later of q_j
#endif
  double fx2[QH][QW];   // inner x flux; rows 3..TY+2 hold the averaged outer flux after phase D
  double fy2[QH][QW];   // inner y flux; rows 3..TY+3 hold the averaged outer flux after phase D
  double qi[QH][QW];    // phases A-B: q in the copy_corners(dir=2) view (cube-corner tiles); C-D: q_i
  double qj[QH][QW];
};
// averaged, unweighted fluxes through the west / south face of cell (i0+c, j0+r) after tp_compute
__device__ __forceinline__ double& FX(Smem& S, int r, int c) { return S.fx2[r + 3][c + 3]; }   // r < TY, c <= TX
__device__ __forceinline__ double& FY(Smem& S, int r, int c) { return S.fy2[r + 3][c + 3]; }   // r <= TY, c < TX

// optional fused epilogue of the stand-alone transport: the height update of update_dz_d (tp2d.cu)
struct ZnEpi { double* zn; const double *dfx, *dfy; const double* kdbl; int slot; };

// strided view of a shared-memory line in sweep coordinates: value at sweep index s
struct SAcc {
  const double* p; int stride; int org;
  PPM_HD __forceinline__ double operator()(int s) const { return p[(s - org) * stride]; }
};

// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, all copies of a tile in flight together
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// cube-edge faces (the one-sided formulas of tp_core.F90:643-681): rare, kept out of line so the interior
// path stays small (instruction cache) and its register allocation is not driven by this code
template <bool RARE>
static __device__ __noinline__ double edge_flux(const double* ql, int sq, int org, const double* dl, long long dbase,
                                                long long dstride, int i, double c, int iord, int n) {
  SAcc qa{ql, sq, org};
  ppm::Acc da{dl, dbase, dstride};
  return ppm::flux_scalar<RARE>(qa, da, i, c, iord, n, true);
}

// interior flux through the low-side face of line element m (between elements m-1 and m); q, a = element m of the line
template <bool RARE>
__device__ __forceinline__ double line_flux(bool mono, const double* q, int sq, const double* a, int sa, double c, int iord) {
  if (RARE && mono && iord == 7) return ppm::flux_pd7_line(q[-3 * sq], q[-2 * sq], q[-sq], q[0], q[sq], q[2 * sq], c);
  if (mono) {
    const int u = (c > 0.) ? -1 : 0;
    const double* p = q + u * sq;
    const double* d = a + u * sa;
    return ppm::flux_mono_aux<RARE>(p[-2 * sq], p[-sq], p[0], p[sq], p[2 * sq], d[-sa], d[0], d[sa], c, iord);
  }
  return ppm::flux_unlim_aux<RARE>(q[-sq], q[0], a[-sa], a[0], a[sa], c, iord);
}

// TPT_NOAUX form: no limiter-input arrays.  The 5 (monotone) / 6 (unlimited) values of q the flux reads anyway contain every
// operand of the three dm / al values it needs, so they are recomputed in registers: +2 limiter evaluations per flux, but
// 3 shared-memory loads per flux, the two aux passes (6 loads + 2 stores per point) and 2 of the 5 barriers per field go away
// (ncu, profiles/r1_dsw_ncu_summary.md: the shared-memory data pipe is at 60 % of peak next to 58 % issue-active).
// Same operations on the same operands as the aux form => bit-identical results.
template <bool RARE>
PPM_HD __forceinline__ double line_flux_na(bool mono, const double* q, int sq, double c, int iord) {   // host too: tests/host_ppm_test.cu
  if (RARE && mono && iord == 7) return ppm::flux_pd7_line(q[-3 * sq], q[-2 * sq], q[-sq], q[0], q[sq], q[2 * sq], c);
  if (mono) {
    const int u = (c > 0.) ? -1 : 0;
    const double* p = q + u * sq;
    const double a = p[-2 * sq], b = p[-sq], m = p[0], d = p[sq], e = p[2 * sq];
    return ppm::flux_mono_aux<RARE>(a, b, m, d, e, ppm::dm2(a, b, m), ppm::dm2(b, m, d), ppm::dm2(m, d, e), c, iord);
  }
  const double a0 = q[-3 * sq], a1 = q[-2 * sq], a2 = q[-sq], a3 = q[0], a4 = q[sq], a5 = q[2 * sq];
  return ppm::flux_unlim_aux<RARE>(a2, a3, ppm::aux_point(false, iord, a0, a1, a2, a3), ppm::aux_point(false, iord, a1, a2, a3, a4),
                             ppm::aux_point(false, iord, a2, a3, a4, a5), c, iord);
}

struct Tile {
  int i0, j0;          // first cell of the tile
  long long ko;        // level offset (elements)
  int NI, ib, jb;      // in-plane index = (i + ib) + (j - jb) * NI
  int lane, wid;
  bool corner;         // tile halo overlaps a cube-corner region: q needs the two copy_corners views
  __device__ __forceinline__ int idx(int i, int j) const { return (i + ib) + (j - jb) * NI; }
};
// Tile coordinates come from a FrameGrid: interior launches (EDGE = false) enumerate the rectangle of tiles whose every
// flux is an ordinary interior one, frame launches (EDGE = true) the rest (fv3_ctx.hpp).
struct TileMap { FrameGrid fg; int interior; };
__device__ __forceinline__ void tile_xy(const TileMap& M, int& bx, int& by) {
  if (M.interior) { const int nin = M.fg.nbx - M.fg.cl - M.fg.cr; by = blockIdx.x / nin; bx = blockIdx.x - by * nin; bx += M.fg.cl; by += M.fg.a; }
  else M.fg.map(blockIdx.x, bx, by);
}
__device__ __forceinline__ Tile make_tile(const Lay& L, const TileMap& M) {
  Tile T;
  int bx, by;
  tile_xy(M, bx, by);
  T.i0 = L.is + bx * TX; T.j0 = L.js + by * TY;
  T.ko = (long long)blockIdx.z * L.plane;
  T.NI = L.NI; T.ib = FV3_IOFF - L.isd; T.jb = L.jsd;
  T.lane = threadIdx.x & 31; T.wid = threadIdx.x >> 5;
  T.corner = L.cube && (T.i0 - 3 <= 0 || T.i0 + TX + 2 >= L.npx) && (T.j0 - 3 <= 0 || T.j0 + TY + 2 >= L.npy);
  return T;
}

// I: issue the per-tile inputs (no wait).  One index per point serves all five arrays; points beyond the padded plane
// are clamped duplicates (never used by a stored result).
template <bool EDGE>
__device__ __forceinline__ void stage_inputs(const Lay& L, const DevGrid& G, Smem& S, const Tile& T,
                                             const double* __restrict__ crx, const double* __restrict__ cry,
                                             const double* __restrict__ xfx, const double* __restrict__ yfx) {
  const int i = EDGE ? min(T.i0 - 3 + T.lane, L.ied + 1) : T.i0 - 3 + T.lane;
#pragma unroll
  for (int r = T.wid; r < QH; r += NW) {
    const int o = T.idx(i, EDGE ? min(T.j0 - 3 + r, L.jed + 1) : T.j0 - 3 + r);
    const long long g = T.ko + o;
    cp_async8(&S.crx[r][T.lane], crx + g);
    cp_async8(&S.xfx[r][T.lane], xfx + g);
    cp_async8(&S.cry[r][T.lane], cry + g);
    cp_async8(&S.yfx[r][T.lane], yfx + g);
    cp_async8(&S.area[r][T.lane], G.area + o);
  }
}

// A: issue the halo tile of q (no wait)
template <bool EDGE>
__device__ __forceinline__ void stage_q(const Lay& L, Smem& S, const Tile& T, const double* __restrict__ q) {
  q += T.ko;
  const int i = EDGE ? min(T.i0 - 3 + T.lane, L.ied) : T.i0 - 3 + T.lane;
#pragma unroll
  for (int r = T.wid; r < QH; r += NW) {
    const int j = EDGE ? min(T.j0 - 3 + r, L.jed) : T.j0 - 3 + r;
    if (EDGE && T.corner) {
      S.q[r][T.lane] = ppm::QAccX{q, L, j}(i);
      S.qi[r][T.lane] = ppm::QAccY{q, L, i}(j);
    } else {
      cp_async8(&S.q[r][T.lane], q + T.idx(i, j));
    }
  }
}

// B0..D for the staged tile.  The caller has issued stage_inputs/stage_q; this waits for them.
// On return (after a __syncthreads) FX(S,r,c) = unweighted Lin-Rood flux 0.5*(outer + inner) through the west face of
// cell (i0+c, j0+r), FY(S,r,c) = south face.  Points beyond the face are not computed; the caller must not use them.
// FAM: the scheme family is a compile-time choice of the launcher: 0 = unlimited (5, 6, -5), 1 = monotone (8, 10),
// 2 = decided per call from ord_ou (kernels that transport fields of both families).  Both sweeps of one hord are in
// the same family (hord 10 runs ord 8 inside, tp_core.F90:136-141).
// The row loops are deliberately NOT unrolled: the unrolled form of this routine, inlined once per transported field,
// overflowed the instruction cache (ncu: 23 % of the stall samples were no_instruction).
// EDGE = false (interior tiles): every face of the tile, halo faces included, is an ordinary interior face and the tile
// lies inside the face -- the cube-edge operator, the copy_corners views and all bound tests are compiled out.
template <int FAM, bool EDGE>
__device__ __forceinline__ void tp_compute(const Lay& L, const DevGrid& G, Smem& S, const Tile& T,
                                           const double* __restrict__ ra_x, const double* __restrict__ ra_y,
                                           int ord_in, int ord_ou) {
  using namespace ppm;
  const bool mono = (FAM == 2) ? (ord_ou >= 7) : (FAM == 1);   // dm family: iord >= 7 (tp_core.F90:364, 563)
  constexpr bool RARE = (FAM == 2);   // only the general instantiation carries the schemes beyond 5, 6, -5, 8, 10
  const bool cube = EDGE && L.cube;
  const int npx = L.npx, npy = L.npy, i0 = T.i0, j0 = T.j0, c = T.lane, wid = T.wid;
  const int i = i0 - 3 + c;                                     // this lane's column (cell / west-face index)
  const bool xface = c >= 3 && c <= TX + 3 && (!EDGE || i <= L.ie + 1);    // lane owns a west face that is needed
  const bool xcell = c >= 3 && c <= TX + 2 && (!EDGE || i <= L.ie);        // lane owns a cell of the tile
  const bool xfast = !cube || (i >= 4 && i <= npx - 3);
  const int cm2 = max(c - 2, 0), cm1 = max(c - 1, 0), cp1 = min(c + 1, QW - 1);
  // the tile's x faces that need the cube-edge operator: columns 3..2+nwc (faces i <= 3) and ce0..ce0+nec-nwc-1
  // (faces npx-2 <= i <= npx).  They are evaluated in a separate dense pass: in the row-per-warp loops they would keep
  // 3 lanes of every row busy with the long out-of-line operator while 29 wait (the frame tiles ran 2x slower per
  // tile than the interior ones: profiles/r1_dsw_ncu_summary.md, v6).
  int nwc = 0, ce0 = 0, nec = 0;
  if (EDGE && cube) {
    nwc = max(0, min(TX + 3, 6 - i0) - 2);
    ce0 = max(3, npx + 1 - i0);
    nec = nwc + max(0, min(TX + 3, npx + 3 - i0) - ce0 + 1);
  }
  cp_async_wait_all();
  __syncthreads();
  const double(*qy)[QW] = (EDGE && T.corner) ? S.qi : S.q;
#ifndef TPT_NOAUX
  // ---- B0: per-point limiter inputs of both sweeps (border points are clamped garbage that no face reads)
#pragma unroll
  for (int r = wid; r < QH; r += NW) {
    S.ax[r][c] = aux_point(mono, ord_in, S.q[r][cm2], S.q[r][cm1], S.q[r][c], S.q[r][cp1]);
    const int r2 = max(r - 2, 0), r1 = max(r - 1, 0), r3 = min(r + 1, QH - 1);
    S.ay[r][c] = aux_point(mono, ord_in, qy[r2][c], qy[r1][c], qy[r][c], qy[r3][c]);
  }
  __syncthreads();
#define TPT_LF(qp, sq, ap, sa, cr, ord) line_flux<RARE>(mono, qp, sq, ap, sa, cr, ord)
#else
  (void)cm2; (void)cm1; (void)cp1;
#define TPT_LF(qp, sq, ap, sa, cr, ord) line_flux_na<RARE>(mono, qp, sq, cr, ord)
#endif
  // ---- B: inner sweeps; task t < QH: fx2 row t (tp_core.F90:164-169), else fy2 row t-QH+3 (:143-148)
#pragma unroll
  for (int t = wid; t < QH + TY + 1; t += NW) {
    if (t < QH) {
      const int r = t, j = j0 - 3 + r;
      if (xface && xfast && (!EDGE || j <= L.jed)) S.fx2[r][c] = TPT_LF(&S.q[r][c], 1, &S.ax[r][c], 1, S.crx[r][c], ord_in);
    } else {
      const int r = t - QH + 3, j = j0 - 3 + r;
      if (!EDGE || (j <= L.je + 1 && i <= L.ied)) {
        const double cr = S.cry[r][c];
        S.fy2[r][c] = (!cube || (j >= 4 && j <= npy - 3)) ? TPT_LF(&qy[r][c], QW, &S.ay[r][c], QW, cr, ord_in)
                                                          : edge_flux<RARE>(&qy[0][c], QW, j0 - 3, G.dya, LIDX(L, i, 0), L.NI, j, cr, ord_in, npy);
      }
    }
  }
  if (EDGE && nec > 0) {   // cube-edge x faces of all rows, densely packed over the CTA
    for (int e = threadIdx.x; e < QH * nec; e += NT) {
      const int r = e / nec, ci = e - r * nec, cc = ci < nwc ? 3 + ci : ce0 + (ci - nwc), j = j0 - 3 + r;
      if (j <= L.jed) S.fx2[r][cc] = edge_flux<RARE>(&S.q[r][0], 1, i0 - 3, G.dxa, LIDX(L, 0, j), 1, i0 - 3 + cc, S.crx[r][cc], ord_in, npx);
    }
  }
  __syncthreads();
  // ---- C: intermediate advected fields; task t < TY: q_i row t+3 (tp_core.F90:150-159), else q_j row t-TY (:171-178)
#pragma unroll
  for (int t = wid; t < TY + QH; t += NW) {
    if (t < TY) {
      const int r = t + 3, j = j0 - 3 + r;
      if (!EDGE || (j <= L.je && i <= L.ied)) {
        const double ar = S.area[r][c];
        const double y0 = S.yfx[r][c], y1 = S.yfx[r + 1][c];
        const double f0 = y0 * S.fy2[r][c], f1 = y1 * S.fy2[r + 1][c];
        const double ray = ra_y ? __ldg(ra_y + T.ko + T.idx(i, j)) : (ar + y0 - y1);
        S.qi[r][c] = (S.q[r][c] * ar + f0 - f1) / ray;
      }
    } else {
      const int r = t - TY, j = j0 - 3 + r;
      if (xcell && (!EDGE || j <= L.jed)) {
        const double ar = S.area[r][c];
        const double x0 = S.xfx[r][c], x1 = S.xfx[r][c + 1];
        const double f0 = x0 * S.fx2[r][c], f1 = x1 * S.fx2[r][c + 1];
        const double rax = ra_x ? __ldg(ra_x + T.ko + T.idx(i, j)) : (ar + x0 - x1);
        S.qj[r][c] = (S.q[r][c] * ar + f0 - f1) / rax;
      }
    }
  }
  __syncthreads();
#ifndef TPT_NOAUX
  // ---- C2: limiter inputs of q_i (x lines, rows 3..TY+2) and q_j (y lines, all rows)
#pragma unroll
  for (int t = wid; t < TY + QH; t += NW) {
    if (t < TY) {
      const int r = t + 3;
      S.ax[r][c] = aux_point(mono, ord_ou, S.qi[r][cm2], S.qi[r][cm1], S.qi[r][c], S.qi[r][cp1]);
    } else {
      const int r = t - TY;
      const int r2 = max(r - 2, 0), r1 = max(r - 1, 0), r3 = min(r + 1, QH - 1);
      S.ay[r][c] = aux_point(mono, ord_ou, S.qj[r2][c], S.qj[r1][c], S.qj[r][c], S.qj[r3][c]);
    }
  }
  __syncthreads();
#endif
  // ---- D: outer sweeps, averaged with the inner fluxes in place (tp_core.F90:161,180,193-198)
#pragma unroll
  for (int t = wid; t < TY + TY + 1; t += NW) {
    if (t < TY) {
      const int r = t + 3, j = j0 - 3 + r;
      if (xface && xfast && (!EDGE || j <= L.je))
        S.fx2[r][c] = 0.5 * (TPT_LF(&S.qi[r][c], 1, &S.ax[r][c], 1, S.crx[r][c], ord_ou) + S.fx2[r][c]);
    } else {
      const int r = t - TY + 3, j = j0 - 3 + r;
      if (xcell && (!EDGE || j <= L.je + 1)) {
        const double cr = S.cry[r][c];
        const double f = (!cube || (j >= 4 && j <= npy - 3)) ? TPT_LF(&S.qj[r][c], QW, &S.ay[r][c], QW, cr, ord_ou)
                                                             : edge_flux<RARE>(&S.qj[0][c], QW, j0 - 3, G.dya, LIDX(L, i, 0), L.NI, j, cr, ord_ou, npy);
        S.fy2[r][c] = 0.5 * (f + S.fy2[r][c]);
      }
    }
  }
  if (EDGE && nec > 0) {
    for (int e = threadIdx.x; e < TY * nec; e += NT) {
      const int r = 3 + e / nec, ci = e - (r - 3) * nec, cc = ci < nwc ? 3 + ci : ce0 + (ci - nwc), j = j0 - 3 + r;
      if (j <= L.je)
        S.fx2[r][cc] = 0.5 * (edge_flux<RARE>(&S.qi[r][0], 1, i0 - 3, G.dxa, LIDX(L, 0, j), 1, i0 - 3 + cc, S.crx[r][cc], ord_ou, npx) + S.fx2[r][cc]);
    }
  }
  __syncthreads();
#undef TPT_LF
}

// tile maps of a face: the interior rectangle (tiles with 4 <= every face index <= npx-3 and the whole halo inside the
// data domain) and the frame around it
static inline void tile_maps(const Lay& L, TileMap& in, TileMap& fr, int& n_in, int& n_fr) {
  const int nx = L.ie - L.is + 1, ny = L.je - L.js + 1;
  FrameGrid f;
  f.nbx = (nx + TX - 1) / TX; f.nby = (ny + TY - 1) / TY;
  int cl = 0, cr = 0, a = 0, bt = 0;
  if (L.cube) {
    while (cl < f.nbx && L.is + cl * TX < 4) cl++;
    while (cr < f.nbx - cl && L.is + (f.nbx - cr) * TX > L.npx - 3) cr++;
    while (a < f.nby && L.js + a * TY < 4) a++;
    while (bt < f.nby - a && L.js + (f.nby - bt) * TY > L.npy - 3) bt++;
  } else {   // doubly periodic: only the tiles that overhang the face need the bound tests
    while (cr < f.nbx && L.is + (f.nbx - cr) * TX - 1 > L.ie) cr++;
    while (bt < f.nby && L.js + (f.nby - bt) * TY - 1 > L.je) bt++;
  }
  f.cl = cl; f.cr = cr; f.a = a; f.b = f.nby - bt;
  const int nin_x = f.nbx - cl - cr, nin_y = f.b - f.a;
  n_in = (nin_x > 0 && nin_y > 0) ? nin_x * nin_y : 0;
  if (n_in == 0) { f.cl = f.nbx; f.cr = 0; f.a = f.nby; f.b = f.nby; }   // everything is frame
  in.fg = f; in.interior = 1; fr.fg = f; fr.interior = 0;
  n_fr = f.count();
}

}  // namespace tpt
