// fv_tp_2d on a shared-memory tile, LINE-PER-WARP form (round 2) -- the interior-tile instance of the Lin-Rood transport.
//
// Reference semantics: model/tp_core.F90:85-241 (fv_tp_2d), :324-712 (xppm), :715-1152 (yppm) away from the cube edges.
//
// Why a second form (profiles/r1_dsw_ncu_v10.txt): the first tile kernel (tp_tile.cuh) evaluates every flux from scratch --
// a thread re-derives the three dm / al values of the upwind cell from a 5-6 point window, so every limiter input is computed
// three times and every q value is read from shared memory five times; the kernel sat on issue slots (66 %), the fp64 pipe
// (42 %) and the shared-memory crossbar (46 %) at once.  Here a warp owns a whole 32-cell LINE of the 32 x 32 halo tile, one cell
// per lane:
//   * the PPM edge values are computed ONCE per cell (dm, al, then bl / br) and handed to the neighbour lane with warp
//     shuffles (dm <- lane-1, al <- lane+1); each lane forms the two candidate fluxes its cell can feed (through its low face
//     if the wind blows from the high side, through its high face otherwise) and the face picks one (one more shuffle);
//   * x lines are tile rows (lane = column); y lines are tile COLUMNS (lane = row): every tile array has a pitch of 33 doubles,
//     which makes both access patterns bank-conflict free, so the same routine serves both sweeps;
//   * the inner flux of a face stays in a register of the lane that owns the face until the outer sweep averages it (the same
//     warp runs the inner and the outer sweep of a line); q_i / q_j are formed by the inner line task itself (the flux of the
//     next face arrives by shuffle), so a level needs three barriers in all -- whatever the number of fields;
//   * several fields are transported PHASE-major (all inner sweeps, barrier, all outer sweeps): the Courant numbers, area
//     fluxes and 1/ra of a line are loaded once for all fields;
//   * a CTA (32 warps, one per SM) is persistent over a chunk of levels and double-buffers its inputs: level k+1 streams in with
//     cp.async while level k computes; cell areas are staged once per tile.
// Only tiles whose every flux (halo fluxes included) is an ordinary interior flux run here; the frame tiles keep the general
// kernel of tp_tile.cuh (cube-edge operators, copy_corners views, bound tests).  Same tiling (26 x 26) as tp_tile.cuh.
//
// ra_x / ra_y: formed as area + xfx(i) - xfx(i+1) (tp_core.F90 callers in d_sw / update_dz_d, sw_core.F90:905-917); q_i / q_j
// multiply by the reciprocal (one division per line and level instead of one per cell and field): <= 1 ulp from the divide.
#pragma once
#include "ppm.cuh"
#include "tp_tile.cuh"
#include <cstdlib>
#include <cstdio>

namespace tp2 {

constexpr int TX = 26, TY = 26, QW = 32, QH = 32, P = 33;
constexpr int NW = 32, NT = NW * 32;
constexpr int ASZ = QH * P;      // doubles per tile array
constexpr int GUARD = 72;        // >= 2*P + 2: the neighbour reads of lanes 0, 1, 30, 31 fall into the guards (values never used)
static_assert(tpt::TX == TX && tpt::TY == TY, "tp_line.cuh and tp_tile.cuh must tile a face identically");

enum { A_CRX = 0, A_CRY = 1, A_XFX = 2, A_YFX = 3, A_Q = 4 };
// what the outer sweep leaves in qi (x faces) / qj (y faces):
//   W_AREA: flux * area flux (xfx / yfx) for every field         (tp_core.F90:193-198 without mfx)
//   W_MASS: field 0: flux * area flux = the mass flux; field f > 0: flux * mass flux   (sw_core.F90:928-940, tp_core.F90:213-226)
//   W_RAW : the unweighted Lin-Rood flux 0.5 * (outer + inner)
//   W_MASS_V: W_MASS for fields 0 .. NF-2, W_AREA for the last one (d_sw: delp, w, pt and the absolute vorticity in one kernel)
enum { W_AREA = 0, W_MASS = 1, W_RAW = 2, W_MASS_V = 3 };

// T = the type the sweeps compute in.  double: the fp64 product path.  sf ("transport_fp32", BASELINE config 5: fp32 transport on
// fp64 storage): the level's inputs land as doubles (cp.async cannot convert), each thread converts the elements it fetched itself
// into the fp32 working set; the fluxes leave the sweeps as fp32 values and the epilogue applies them in fp64, so the update stays
// conservative: both cells of a face use the same rounded flux.  For a face shared by two TILES that holds only if both tiles
// compute bit-identical fluxes although one may run the interior and the other the frame instantiation of the kernel.  Hence `sf`
// (strict float): every operation is a separately rounded IEEE operation (__fadd_rn, __fmul_rn, ...: never contracted into an
// FMA), so the same expression tree on the same inputs gives the same bits in every instantiation.  (In fp64 the two
// instantiations may differ by an FMA contraction, i.e. 1e-16 -- inside the round-off the conservation statement allows.)
struct sf {
  float v;
  sf() = default;
  __host__ __device__ constexpr sf(float x) : v(x) {}
  __host__ __device__ constexpr sf(double x) : v((float)x) {}
  __host__ __device__ constexpr operator double() const { return (double)v; }
};
__device__ __forceinline__ sf operator+(sf a, sf b) { return sf(__fadd_rn(a.v, b.v)); }
__device__ __forceinline__ sf operator-(sf a, sf b) { return sf(__fsub_rn(a.v, b.v)); }
__device__ __forceinline__ sf operator*(sf a, sf b) { return sf(__fmul_rn(a.v, b.v)); }
__device__ __forceinline__ sf operator/(sf a, sf b) { return sf(__fdiv_rn(a.v, b.v)); }
__device__ __forceinline__ sf operator-(sf a) { return sf(-a.v); }
__device__ __forceinline__ bool operator<(sf a, sf b) { return a.v < b.v; }
__device__ __forceinline__ bool operator>(sf a, sf b) { return a.v > b.v; }
template <int NF, int NEP = 0, bool EDGE = false, typename T = double>
struct Smem {
  double g0[GUARD];
  // frame tiles: the fluxes through the cube-edge faces of every line (sweep index 1..3 and n-2..n: six slots), evaluated densely by
  // dense_edge_fluxes before the line tasks run, [field][direction][line][slot]
  double ef[EDGE ? NF * 2 * 32 * 6 : 1];
  // frame tiles: dxa (x lines) / dya (y lines) at the four cells around the low (e = 1) and the high (e = n) cube edge of every
  // line, [direction][line][edge][e-2 .. e+1] -- k-invariant, filled once per tile (the two-sided edge value, :376-377, 647-648)
  double et[EDGE ? 2 * 32 * 2 * 4 : 1];
  // fp32 sweeps: where cp.async lands the next level.  The fields' landing zone is double-buffered: the epilogue applies the fluxes to
  // the fp64 values of the level (q64), which must survive the arrival of the next one.
  double landc[sizeof(T) == 4 ? 4 : 1][sizeof(T) == 4 ? ASZ : 1];
  double landq[sizeof(T) == 4 ? 2 : 1][sizeof(T) == 4 ? NF : 1][sizeof(T) == 4 ? ASZ : 1];
  T in[2][4 + NF][ASZ];        // per level, double-buffered: crx, cry, xfx, yfx, q_0 .. q_{NF-1}
  T area[ASZ];
  T qi[NF][ASZ];               // q_i (tp_core.F90:150-159); after the outer sweep: the x fluxes
  T qj[NF][ASZ];               // q_j (:171-178);            after the outer sweep: the y fluxes
  double ep[NEP > 0 ? NEP : 1][NEP > 0 ? ASZ : 1];   // operands of the epilogue fetched with cp.async while the sweeps run
  double g1[GUARD];
  // field f of the level in buffer b at element o, in fp64 (what the epilogues update)
  __device__ __forceinline__ double q64(int b, int f, int o) const {
    if constexpr (sizeof(T) == 4) return landq[b][f][o];
    else return in[b][A_Q + f][o];
  }
};

template <typename T> __device__ __forceinline__ T up1(T x) { return __shfl_up_sync(0xffffffffu, x, 1); }
template <typename T> __device__ __forceinline__ T dn1(T x) { return __shfl_down_sync(0xffffffffu, x, 1); }
// the limiter arithmetic for both types: fp64 = the ppm.cuh routines (compare-select min / max: no fp64 min / max instruction),
// fp32 = single FMNMX / |x| / copysign instructions
using ppm::mn; using ppm::mx; using ppm::fsign; using ppm::min3; using ppm::max3; using ppm::dm2;
template <> __device__ __forceinline__ sf up1<sf>(sf x) { return sf(__shfl_up_sync(0xffffffffu, x.v, 1)); }
template <> __device__ __forceinline__ sf dn1<sf>(sf x) { return sf(__shfl_down_sync(0xffffffffu, x.v, 1)); }
__device__ __forceinline__ sf mn(sf a, sf b) { return sf(fminf(a.v, b.v)); }
__device__ __forceinline__ sf mx(sf a, sf b) { return sf(fmaxf(a.v, b.v)); }
__device__ __forceinline__ sf min3(sf a, sf b, sf c) { return sf(fminf(fminf(a.v, b.v), c.v)); }
__device__ __forceinline__ sf max3(sf a, sf b, sf c) { return sf(fmaxf(fmaxf(a.v, b.v), c.v)); }
__device__ __forceinline__ sf fsign(sf a, sf b) { return sf(copysignf(a.v, b.v)); }
__device__ __forceinline__ double ab(double x) { return fabs(x); }
__device__ __forceinline__ sf ab(sf x) { return sf(fabsf(x.v)); }
__device__ __forceinline__ sf dm2(sf qm, sf q0, sf qp) {   // ppm::dm2 in fp32
  const sf a = q0 - qm, b = qp - q0, xt = sf(0.25f) * (qp - qm);
  const bool same = (__float_as_int(a.v) ^ __float_as_int(b.v)) >= 0;
  const float m = fminf(fminf(fabsf(xt.v), fabsf(a.v)), fabsf(b.v));
  return sf(same ? copysignf(m, xt.v) : 0.f);
}

// Unweighted upwind PPM fluxes through the LOW-side face of this lane's cell for NF fields of one 32-cell line held one cell per
// lane.  o = this lane's element in every tile array, sa = element stride between lanes; cl / cr = Courant number of this
// lane's low face / of the next lane's low face.  Valid for lanes 3..29 (cells 2..29 are reconstructed).  Same operations on
// the same operands as ppm::flux_mono_aux / flux_unlim_aux.  FAM 1: iord 8, 10 (tp_core.F90:563-637, 701-707); FAM 0: 5, 6, -5
// (:369-373, 491-558).  ORD: the scheme as a compile-time constant when every field uses the same one, else ORD_RT (ord[f]).
// The routine is written stage by stage over the fields (every stage a fully unrolled loop over f) and without divergent
// branches, so that the NF dependency chains -- each with four shuffle round trips -- interleave: with one CTA of 32 warps per SM
// (64 registers per thread) instruction-level parallelism is what hides the fp64 / shuffle latencies (ncu, first version of
// this file: 55 % issue-active with the fields evaluated one after the other).
constexpr int ORD_RT = 99;
// Frame tiles (EDGE = true): the faces next to a cube edge (sweep index 1..3 and n-2..n) take the reference's one-sided operator
// (tp_core.F90:374-392, 536-545, 643-681).  Their fluxes are NOT formed by the lanes that own them -- three lanes of a line would
// run the long operator while 29 wait, once per line, field and sweep -- but densely, all lanes busy, by dense_edge_fluxes into
// Smem::ef; the line task only picks them up.  ic = base + lane is the sweep index of a lane's cell / low face, n = npx (npy),
// ef -> the six slots of this line for field 0 (field stride 2*32*6).
struct EdgeLine { bool on; int base, n; const double* ef; };
constexpr int EF_FIELD = 2 * 32 * 6;
__device__ __forceinline__ int edge_slot(int ic, int n) { return (ic >= 1 && ic <= 3) ? ic - 1 : (ic >= n - 2 && ic <= n) ? ic - (n - 2) + 3 : -1; }
template <int FAM, int NF, int ORD, bool EDGE, typename T>
__device__ __forceinline__ void line_fluxes(const T (*__restrict__ q)[ASZ], int o, int sa, int lane, T cl, T cr,
                                            const int (&ord)[NF], T (&q0)[NF], T (&flux)[NF], const EdgeLine& E) {
  constexpr T r3 = T(ppm::r3), r12 = T(ppm::r12), p1 = T(ppm::p1), p2 = T(ppm::p2), near_zero_tp = T(ppm::near_zero_tp);
  constexpr T Z = T(0.), H = T(0.5), ONE = T(1.), TWO = T(2.), THREE = T(3.), Q3 = T(0.75), Q1 = T(0.25);
  const int eslot = (EDGE && E.on) ? edge_slot(E.base + lane, E.n) : -1;   // >= 0: this lane's low face is a cube-edge face
  if (FAM == 1) {
    T qm[NF], qp[NF], dm0[NF], dmm[NF], dmp[NF], al0[NF], al1[NF], bl[NF], br[NF], FR[NF], FL[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) {
      qm[f] = q[f][o - sa]; q0[f] = q[f][o]; qp[f] = q[f][o + sa];
      dm0[f] = dm2(qm[f], q0[f], qp[f]);
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
      dmm[f] = up1(dm0[f]);
      if (ORD != 8) dmp[f] = dn1(dm0[f]);
    }
#pragma unroll
    for (int f = 0; f < NF; f++) al0[f] = H * (qm[f] + q0[f]) + r3 * (dmm[f] - dm0[f]);
#pragma unroll
    for (int f = 0; f < NF; f++) al1[f] = dn1(al0[f]);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int iord = (ORD == ORD_RT) ? ord[f] : ORD;
      // iord 8 (tp_core.F90:591-597)
      const T xt = TWO * dm0[f];
      const T bl8 = -fsign(mn(ab(xt), ab(al0[f] - q0[f])), xt);
      const T br8 = fsign(mn(ab(xt), ab(al1[f] - q0[f])), xt);
      if (ORD == 8) { bl[f] = bl8; br[f] = br8; continue; }
      // iord 10 (:605-627 with the pmp / lac constraint), branch-free: the constraint applies where the parabola overshoots
      T b_l = al0[f] - q0[f], b_r = al1[f] - q0[f];
      const bool flat = ab(dmm[f]) + ab(dm0[f]) + ab(dmp[f]) < near_zero_tp;
      const bool over = ab(THREE * (b_l + b_r)) > ab(b_l - b_r);
      const T qm2 = q[f][o - 2 * sa], qp2 = q[f][o + 2 * sa];
      const T dqm2 = TWO * (qm[f] - qm2), dqm1 = TWO * (q0[f] - qm[f]), dq0 = TWO * (qp[f] - q0[f]), dqp1 = TWO * (qp2 - qp[f]);
      const T pmp_2 = dqm1, lac_2 = pmp_2 - Q3 * dqm2;
      const T brl = mn(max3(Z, pmp_2, lac_2), mx(b_r, min3(Z, pmp_2, lac_2)));
      const T pmp_1 = -dq0, lac_1 = pmp_1 + Q3 * dqp1;
      const T bll = mn(max3(Z, pmp_1, lac_1), mx(b_l, min3(Z, pmp_1, lac_1)));
      if (over) { b_l = bll; b_r = brl; }
      if (flat) { b_l = Z; b_r = Z; }
      if (ORD == ORD_RT && iord == 8) { b_l = bl8; b_r = br8; }
      bl[f] = b_l; br[f] = b_r;
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const T b0 = bl[f] + br[f];
      FR[f] = q0[f] + (ONE - cr) * (br[f] - cr * b0);   // through the high face, wind from this cell (Courant number > 0 there)
      FL[f] = q0[f] + (ONE + cl) * (bl[f] + cl * b0);   // through the low face, wind from this cell (Courant number <= 0)
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const T FRm = up1(FR[f]);
      flux[f] = cl > Z ? FRm : FL[f];
      if (EDGE && eslot >= 0) flux[f] = T(E.ef[f * EF_FIELD + eslot]);
    }
  } else {
    T qm[NF], al0[NF], al1[NF], F1R[NF], F1L[NF];
    bool smt[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int iord = (ORD == ORD_RT) ? ord[f] : ORD;
      const T qm2 = q[f][o - 2 * sa], qp = q[f][o + sa];
      qm[f] = q[f][o - sa]; q0[f] = q[f][o];
      T a = p1 * (qm[f] + q0[f]) + p2 * (qm2 + qp);
      if (iord < 0) a = mx(Z, a);
      al0[f] = a;
    }
#pragma unroll
    for (int f = 0; f < NF; f++) al1[f] = dn1(al0[f]);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int iord = (ORD == ORD_RT) ? ord[f] : ORD;
      T bl = al0[f] - q0[f], br = al1[f] - q0[f], b0 = bl + br;
      bool sm_;
      if (iord == 5) sm_ = bl * br < Z;
      else if (iord == -5) {
        sm_ = bl * br < Z;
        const T da1 = br - bl, a4 = -THREE * b0;
        if (ab(da1) < -a4) {
          if (q0[f] + Q1 / a4 * (da1 * da1) + a4 * r12 < Z) {
            if (!sm_) { br = Z; bl = Z; b0 = Z; }
            else if (da1 > Z) { br = -TWO * bl; b0 = -bl; }
            else { bl = -TWO * br; b0 = -br; }
          }
        }
      } else sm_ = THREE * ab(b0) < ab(bl - br);
      smt[f] = sm_;
      F1R[f] = (ONE - cr) * (br - cr * b0);
      F1L[f] = (ONE + cl) * (bl + cl * b0);
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const T F1Rm = up1(F1R[f]);
      const unsigned sm = __ballot_sync(0xffffffffu, smt[f]);
      const bool smtA = ((sm << 1) >> lane) & 1u;          // smt of the cell on the low side of the face
      T fl = cl > Z ? qm[f] : q0[f];
      if (smtA || smt[f]) fl = fl + (cl > Z ? F1Rm : F1L[f]);
      flux[f] = fl;
      if (EDGE && eslot >= 0) flux[f] = T(E.ef[f * EF_FIELD + eslot]);
    }
  }
}

// inner sweep of one line for all fields (tp_core.F90:143-148 / 164-169) + the intermediate field it feeds (:150-159 / 171-178).
// o = this lane's element, sa = lane stride; cr_ / xf_ = Courant numbers / area fluxes of the sweep direction.
template <int FAM, int NF, int ORD, bool EDGE, typename T>
__device__ __forceinline__ void inner_line(const T* __restrict__ cr_, const T* __restrict__ xf_, const T* __restrict__ area,
                                           const T (*q)[ASZ], T (*qout)[ASZ], int o, int sa, int lane,
                                           const int (&ord)[NF], T (&fin)[NF], const EdgeLine& E) {
  const T cl = cr_[o], cr = cr_[o + sa], xl = xf_[o], xr = xf_[o + sa], ar = area[o];
  const T rra = T(1.) / (ar + xl - xr);
  T q0[NF], g[NF];
  if constexpr (NF > 3 && sizeof(T) == 8) {   // four fp64 fields staged together do not fit 64 registers: three + one
    const int o3[3] = {ord[0], ord[1], ord[2]}, o1[1] = {ord[3]};
    T q3[3], f3[3], q1[1], f1[1];
    line_fluxes<FAM, 3, ORD, EDGE, T>(q, o, sa, lane, cl, cr, o3, q3, f3, E);
    const EdgeLine E1{E.on, E.base, E.n, E.ef + 3 * EF_FIELD};
    line_fluxes<FAM, 1, ORD, EDGE, T>(q + 3, o, sa, lane, cl, cr, o1, q1, f1, E1);
#pragma unroll
    for (int f = 0; f < 3; f++) { q0[f] = q3[f]; fin[f] = f3[f]; }
    q0[3] = q1[0]; fin[3] = f1[0];
  } else line_fluxes<FAM, NF, ORD, EDGE, T>(q, o, sa, lane, cl, cr, ord, q0, fin, E);
#pragma unroll
  for (int f = 0; f < NF; f++) g[f] = fin[f] * xl;
#pragma unroll
  for (int f = 0; f < NF; f++) {
    const T g1 = dn1(g[f]);
    qout[f][o] = (q0[f] * ar + g[f] - g1) * rra;
  }
}

// outer sweep of one line for all fields, averaged with the inner flux and weighted (tp_core.F90:161, 180, 193-226); in place
template <int FAM, int NF, int WMODE, int ORD, bool EDGE, typename T>
__device__ __forceinline__ void outer_line(const T* __restrict__ cr_, const T* __restrict__ xf_, T (*__restrict__ q)[ASZ], int o,
                                           int sa, int lane, const int (&ord)[NF], const T (&fin)[NF], const EdgeLine& E) {
  const T cl = cr_[o], cr = cr_[o + sa];
  const T xl = (WMODE == W_RAW) ? T(1.) : xf_[o];
  T q0[NF], fo[NF];
  // the fields interleaved (one staged evaluation) when their working set fits the 64 registers of a 1024-thread CTA: the
  // branch-free iord = 10 constraint keeps ~12 doubles live per field, so there the fields go one after the other
  constexpr bool SEQ = ((FAM == 1 && ORD != 8 && NF > 1) || NF > 3) && sizeof(T) == 8;
  if (SEQ) {
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int o1[1] = {ord[f]};
      T q1[1], f1[1];
      const EdgeLine Ef{E.on, E.base, E.n, E.ef + f * EF_FIELD};
      line_fluxes<FAM, 1, ORD, EDGE, T>(q + f, o, sa, lane, cl, cr, o1, q1, f1, Ef);
      fo[f] = f1[0];
    }
  } else line_fluxes<FAM, NF, ORD, EDGE, T>(q, o, sa, lane, cl, cr, ord, q0, fo, E);
  T m = T(0.);
#pragma unroll
  for (int f = 0; f < NF; f++) {
    const T F = T(0.5) * (fo[f] + fin[f]);
    T out;
    if (WMODE == W_RAW) out = F;
    else if (WMODE == W_AREA || f == 0) { out = F * xl; m = out; }
    else if (WMODE == W_MASS_V && f == NF - 1) out = F * xl;
    else out = F * m;
    q[f][o] = out;
  }
}

struct Geo { int i0, j0, NI, ib, jb, lane, wid; bool corner; };
__device__ __forceinline__ int gidx(const Geo& T, int i, int j) { return (i + T.ib) + (j - T.jb) * T.NI; }

__device__ __forceinline__ Geo make_geo(const Lay& L, const tpt::TileMap& M) {
  Geo T;
  int bx, by;
  tpt::tile_xy(M, bx, by);
  T.i0 = L.is + bx * TX; T.j0 = L.js + by * TY;
  T.NI = L.NI; T.ib = FV3_IOFF - L.isd; T.jb = L.jsd;
  T.lane = threadIdx.x & 31; T.wid = threadIdx.x >> 5;
  // the tile's halo overlaps a cube-corner region: q is needed in both copy_corners views (tp_core.F90:245-322)
  T.corner = L.cube && (T.i0 - 3 <= 0 || T.i0 + TX + 2 >= L.npx) && (T.j0 - 3 <= 0 || T.j0 + TY + 2 >= L.npy);
  return T;
}

// issue the inputs of one level into buffer b: 32 / NWC elements of every array per thread (row = warp (+ NWC), column = lane).
// g = level offset + in-plane index of the thread's first element (frame tiles: index clamped to the padded plane), dj = in-plane
// distance to its next element (NWC rows; frame tiles whose second row falls off the plane repeat the last row).
// skip_q: cube-corner tiles stage their q views themselves (stage_corner_q)
template <int NF, int NEP, int NWC, bool EDGE, typename T>
__device__ __forceinline__ void stage_level(Smem<NF, NEP, EDGE, T>& S, int b, const double* const (&src)[4 + NF], long long g, int so, int dj, bool skip_q) {
#pragma unroll
  for (int i = 0; i < 32 / NWC; i++)
#pragma unroll
    for (int a = 0; a < 4 + NF; a++)
      if (a < 4 || !skip_q) {
        if constexpr (sizeof(T) == 8) tpt::cp_async8(&S.in[b][a][so + i * NWC * P], src[a] + g + (long long)i * dj);
        else tpt::cp_async8(a < 4 ? &S.landc[a][so + i * NWC * P] : &S.landq[b][a - 4][so + i * NWC * P], src[a] + g + (long long)i * dj);
      }
}
// fp32 sweeps: the elements this thread fetched (its cp.async group has completed) -> the float working set of buffer b
template <int NF, int NEP, int NWC, bool EDGE, typename T>
__device__ __forceinline__ void convert_level(Smem<NF, NEP, EDGE, T>& S, int b, int so) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < 32 / NWC; i++)
#pragma unroll
      for (int a = 0; a < 4 + NF; a++) S.in[b][a][so + i * NWC * P] = (T)(a < 4 ? S.landc[a][so + i * NWC * P] : S.landq[b][a - 4][so + i * NWC * P]);
  }
}
// cube-corner tiles: the copy_corners(dir = 1) view of every field into the level buffer (read by the x lines) and the dir = 2 view
// into qi (read, then overwritten with q_i, by the y lines -- a column of qi is only ever touched by the warp that owns it)
template <int NF, int NEP, int NWC, bool EDGE, typename R>
__device__ __forceinline__ void stage_corner_q(const Lay& L, Smem<NF, NEP, EDGE, R>& S, int b, const Geo& T, const double* const (&src)[4 + NF], long long ko) {
  const int i = min(T.i0 - 3 + T.lane, L.ied);
#pragma unroll
  for (int n = 0; n < 32 / NWC; n++) {
    const int r = T.wid + n * NWC, j = min(T.j0 - 3 + r, L.jed);
#pragma unroll
    for (int f = 0; f < NF; f++) {
      S.in[b][A_Q + f][r * P + T.lane] = (R)ppm::QAccX{src[A_Q + f] + ko, L, j}(i);
      S.qi[f][r * P + T.lane] = (R)ppm::QAccY{src[A_Q + f] + ko, L, i}(j);
    }
  }
}

// dxa / dya of a line at the cells around its two cube edges, from the per-tile table: sweep indices -1..2 and n-2..n+1
struct EdgeMetric {
  const double* t; int n;
  __device__ __forceinline__ double operator()(int i) const { return i <= 2 ? t[i + 1] : t[4 + i - (n - 2)]; }
};
// Frame tiles: the fluxes through the cube-edge faces of lines l0..l1 of both directions for all fields, one face per thread
// (dense enumeration: every lane of the participating warps has a face), by the general per-face operator of the first-generation
// kernel (tpt::edge_flux -> ppm::flux_scalar: upwind cell first, one-sided (bl, br) / al of that cell from the line in shared
// memory, dxa / dya from the per-tile table).  qx / qy: the field as the x / y sweeps see it.
template <typename R> struct SAccT {   // tpt::SAcc over the working type (the one-sided operators themselves always run in fp64)
  const R* p; int stride; int org;
  __device__ __forceinline__ double operator()(int s) const { return (double)p[(s - org) * stride]; }
};
template <int NF, int NEP, int NWC, typename R>
__device__ __forceinline__ void dense_edge_fluxes(const Lay& L, const DevGrid& G, Smem<NF, NEP, true, R>& S, int b, const Geo& T,
                                                  const R (*qx)[ASZ], const R (*qy)[ASZ], const int (&ord)[NF], int l0, int l1) {
  // slots present in this tile, per direction: face index 1..3 (slots 0..2) and n-2..n (slots 3..5) that fall on columns 3..29
  int sx[6], sy[6], nsx = 0, nsy = 0;
#pragma unroll
  for (int s_ = 0; s_ < 6; s_++) {
    const int ix = (s_ < 3 ? 1 + s_ : L.npx - 5 + s_) - (T.i0 - 3), iy = (s_ < 3 ? 1 + s_ : L.npy - 5 + s_) - (T.j0 - 3);
    if (ix >= 3 && ix <= TX + 3) sx[nsx++] = s_;
    if (iy >= 3 && iy <= TY + 3) sy[nsy++] = s_;
  }
  const int nl = l1 - l0 + 1, ntx = nsx * nl, nty = nsy * nl, per_f = ntx + nty, total = per_f * NF;
  for (int e = threadIdx.x; e < total; e += NWC * 32) {
    const int f = e / per_f, r_ = e - f * per_f;
    const bool xd = r_ < ntx;
    const int q_ = xd ? r_ : r_ - ntx;
    const int si = q_ / nl, line = l0 + q_ - si * nl;   // slot-major: the lanes of a warp share the face, i.e. the one-sided branch
    int slot = 0;
#pragma unroll
    for (int m = 0; m < 6; m++) if (m == si) slot = xd ? sx[m] : sy[m];
    const int n = xd ? L.npx : L.npy;
    const int face = slot < 3 ? 1 + slot : n - 5 + slot;          // sweep index of the face
    int io = 0; const int* ordp = ord;
#pragma unroll
    for (int m = 0; m < NF; m++) if (m == f) io = ordp[m];
    if (xd) {
      const int r = line, c = face - (T.i0 - 3);
      const SAccT<R> qa{qx[f] + r * P, 1, T.i0 - 3};
      const EdgeMetric da{S.et + r * 8, n};
      S.ef[f * EF_FIELD + r * 6 + slot] = ppm::flux_scalar<false>(qa, da, face, (double)S.in[b][A_CRX][r * P + c], io, n, true);
    } else {
      const int c = line, r = face - (T.j0 - 3);
      const SAccT<R> qa{qy[f] + c, P, T.j0 - 3};
      const EdgeMetric da{S.et + 256 + c * 8, n};
      S.ef[f * EF_FIELD + 192 + c * 6 + slot] = ppm::flux_scalar<false>(qa, da, face, (double)S.in[b][A_CRY][r * P + c], io, n, true);
    }
  }
}

#ifdef FV3_TP2_PROF
#define TP2_CLK(i) do { const long long t_ = clock64(); prof[i] += t_ - tprev; tprev = t_; } while (0)
#else
#define TP2_CLK(i)
#endif
// The two sweeps of one staged level by a CTA of NWC warps (32: one line per warp and direction; 16: two).  Warp w owns the
// x lines (tile rows) w, w + NWC and the y lines (tile columns) (w + NWC/2) mod NWC, + NWC -- the offset spreads the lines that
// have no outer task (rows / columns 0..2, 29..31) over different warps.  On return (after a barrier) qi[f] / qj[f] hold the
// fluxes through the west / south face of element [r][c] = cell (i0-3+c, j0-3+r): x faces valid for rows 3..28, columns 3..29;
// y faces for rows 3..29, columns 3..28.
template <int FAM, int NF, int NEP, int WMODE, int HORD, int NWC, bool EDGE, typename R>
__device__ __forceinline__ void compute_level(const Lay& L, const DevGrid& G, Smem<NF, NEP, EDGE, R>& S, int b, const Geo& T, const int (&ord_in)[NF],
                                              const int (&ord_ou)[NF]
#ifdef FV3_TP2_PROF
                                              , long long (&prof)[8], long long& tprev
#endif
) {
  constexpr int OI = (HORD == ORD_RT) ? ORD_RT : (HORD == 10 ? 8 : HORD), OO = HORD;   // tp_core.F90:136-141
  constexpr int LPW = 32 / NWC;
  R finx[LPW][NF], finy[LPW][NF];
  const int yc0 = (T.wid + NWC / 2) & (NWC - 1);
  const bool cube = EDGE && L.cube;
  const R (*qy)[ASZ] = (EDGE && T.corner) ? S.qi : S.in[b] + A_Q;   // the field as the y sweeps see it
  if constexpr (EDGE) {
    if (cube) {   // (CTA-uniform) inner fluxes through the cube-edge faces of all 32 + 32 lines
      dense_edge_fluxes<NF, NEP, NWC, R>(L, G, S, b, T, S.in[b] + A_Q, qy, ord_in, 0, QH - 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < LPW; i++) {
    const int r = T.wid + i * NWC;
    const EdgeLine E{cube, T.i0 - 3, L.npx, S.ef + r * 6};
    inner_line<FAM, NF, OI, EDGE, R>(S.in[b][A_CRX], S.in[b][A_XFX], S.area, S.in[b] + A_Q, S.qj, r * P + T.lane, 1, T.lane, ord_in, finx[i], E);
  }
  TP2_CLK(1);
#pragma unroll
  for (int i = 0; i < LPW; i++) {
    const int c = yc0 + i * NWC;
    const EdgeLine E{cube, T.j0 - 3, L.npy, S.ef + 192 + c * 6};
    inner_line<FAM, NF, OI, EDGE, R>(S.in[b][A_CRY], S.in[b][A_YFX], S.area, qy, S.qi, T.lane * P + c, P, T.lane, ord_in, finy[i], E);
  }
  TP2_CLK(2);
  __syncthreads();
  TP2_CLK(3);
  if constexpr (EDGE) {
    if (cube) {   // outer fluxes through the cube-edge faces of the 26 + 26 lines of the tile proper
      dense_edge_fluxes<NF, NEP, NWC, R>(L, G, S, b, T, S.qi, S.qj, ord_ou, 3, TY + 2);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < LPW; i++) {
    const int r = T.wid + i * NWC;
    const EdgeLine E{cube, T.i0 - 3, L.npx, S.ef + r * 6};
    if (r >= 3 && r <= TY + 2) outer_line<FAM, NF, WMODE, OO, EDGE, R>(S.in[b][A_CRX], S.in[b][A_XFX], S.qi, r * P + T.lane, 1, T.lane, ord_ou, finx[i], E);
  }
#pragma unroll
  for (int i = 0; i < LPW; i++) {
    const int c = yc0 + i * NWC;
    const EdgeLine E{cube, T.j0 - 3, L.npy, S.ef + 192 + c * 6};
    if (c >= 3 && c <= TX + 2) outer_line<FAM, NF, WMODE, OO, EDGE, R>(S.in[b][A_CRY], S.in[b][A_YFX], S.qj, T.lane * P + c, P, T.lane, ord_ou, finy[i], E);
  }
  TP2_CLK(4);
  __syncthreads();
  TP2_CLK(5);
}

// Persistent-over-k driver: the CTA (NWC warps) owns tile blockIdx.x of the map (interior rectangle: EDGE = false, or the frame
// around it: EDGE = true) and the levels [blockIdx.y*kch, +kch) of nk.
//   src: the 4 + NF source arrays (level-0 based; the level offset is added here)
//   pre(S, T, k, ko, r): called for every tile row r = 3..28 this warp finishes, right after level k is staged: may issue cp.async
//        into S.ep[.][r*P + lane] for the operands its epilogue needs (they arrive while the sweeps run; no global-load latency
//        is left in the epilogue, which nothing would overlap in a one-CTA-per-SM kernel)
//   epi(S, b, T, k, ko, r): the per-level epilogue of row r (reads S.qi / S.qj / S.in[b][A_Q + f] / S.ep); frame tiles are also
//        called for r = 29 (the south faces of the face's last row)
//   HORD: the transport scheme of every field as a compile-time constant, or ORD_RT (per-field ord_in / ord_ou at run time)
template <int FAM, int NF, int NEP, int WMODE, int HORD, int NWC, bool EDGE, typename R = double, class Pre, class Epi>
__device__ __forceinline__ void run_tile(const Lay& L, const DevGrid& G, const tpt::TileMap& M, const double* const (&src)[4 + NF], int nk, int kch,
                                         const int (&ord_in)[NF], const int (&ord_ou)[NF], Pre&& pre, Epi&& epi) {
  extern __shared__ __align__(16) unsigned char smem_raw2[];
  Smem<NF, NEP, EDGE, R>& S = *reinterpret_cast<Smem<NF, NEP, EDGE, R>*>(smem_raw2);
  constexpr bool F32 = sizeof(R) == 4;
  const Geo T = make_geo(L, M);
  const int k0 = blockIdx.y * kch, k1 = min(nk, k0 + kch);
  if (k0 >= k1) return;
  const int so = T.wid * P + T.lane;
  // frame tiles may overhang the face: indices are clamped to the padded plane (duplicates are never used by a stored result)
  const int ig = EDGE ? min(T.i0 - 3 + T.lane, L.ied + 1) : T.i0 - 3 + T.lane;
  const int jg = EDGE ? min(T.j0 - 3 + T.wid, L.jed + 1) : T.j0 - 3 + T.wid;
  const int g2 = gidx(T, ig, jg);
  const int dj = (EDGE ? min(T.j0 - 3 + T.wid + NWC, L.jed + 1) - jg : NWC) * T.NI;
  const bool cq = EDGE && T.corner;
  constexpr int REPI = EDGE ? TY + 3 : TY + 2;   // last tile row with an epilogue
  if (EDGE && L.cube && threadIdx.x < 128) {      // the lines' cube-edge metric table (visible after the first level-start barrier)
    const int dir = threadIdx.x >> 6, line = (threadIdx.x >> 1) & 31, e = (threadIdx.x & 1) ? (dir ? L.npy : L.npx) : 1;
    double* t = S.et + dir * 256 + line * 8 + (threadIdx.x & 1) * 4;
    if (dir == 0) { const int j = min(T.j0 - 3 + line, L.jed); for (int m = 0; m < 4; m++) t[m] = __ldg(G.dxa + gidx(T, e - 2 + m, j)); }
    else { const int i = min(T.i0 - 3 + line, L.ied); for (int m = 0; m < 4; m++) t[m] = __ldg(G.dya + gidx(T, i, e - 2 + m)); }
  }
#pragma unroll
  for (int i = 0; i < 32 / NWC; i++) {
    if constexpr (F32) S.area[so + i * NWC * P] = R(__ldg(G.area + g2 + i * dj));
    else tpt::cp_async8(&S.area[so + i * NWC * P], G.area + g2 + i * dj);
  }
  stage_level<NF, NEP, NWC, EDGE, R>(S, 0, src, (long long)k0 * L.plane + g2, so, dj, cq && !F32);   // fp32: q64 needs the plain field
#ifdef FV3_TP2_PROF
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
  for (int k = k0; k < k1; k++) {
    const int b = (k - k0) & 1;
    const long long ko = (long long)k * L.plane;
    tpt::cp_async_wait_all();
    if constexpr (F32) convert_level<NF, NEP, NWC, EDGE, R>(S, b, so);   // own elements only: buffer b was last read two levels ago
    __syncthreads();            // level k is in buffer b; every warp is done with the previous level (buffer b^1, qi, qj, ep)
    if (k + 1 < k1) stage_level<NF, NEP, NWC, EDGE, R>(S, b ^ 1, src, (long long)(k + 1) * L.plane + g2, so, dj, cq && !F32);
    if (NEP > 0) {
#pragma unroll
      for (int i = 0; i < 32 / NWC; i++) { const int r = 3 + T.wid + i * NWC; if (r <= REPI) pre(S, T, k, ko, r); }
    }
    if constexpr (EDGE) {
      if (cq) {                 // (warp-uniform, CTA-uniform) the four corner tiles of a face: q through the remapping accessors
        stage_corner_q<NF, NEP, NWC, EDGE, R>(L, S, b, T, src, ko);
        __syncthreads();
      }
    }
    TP2_CLK(0);
#ifdef FV3_TP2_PROF
    compute_level<FAM, NF, NEP, WMODE, HORD, NWC, EDGE, R>(L, G, S, b, T, ord_in, ord_ou, prof, tprev);
#else
    compute_level<FAM, NF, NEP, WMODE, HORD, NWC, EDGE, R>(L, G, S, b, T, ord_in, ord_ou);
#endif
    if (NEP > 0) tpt::cp_async_wait_all();   // this thread's own epilogue operands (read back by the thread that fetched them)
#pragma unroll
    for (int i = 0; i < 32 / NWC; i++) { const int r = 3 + T.wid + i * NWC; if (r <= REPI) epi(S, b, T, k, ko, r); }
    TP2_CLK(6);
  }
#ifdef FV3_TP2_PROF
  if (NF == 3 && blockIdx.x == 40 && blockIdx.y == 2 && T.lane == 0 && (T.wid % 5 == 0 || T.wid == 31))
    printf("tp2prof warp %2d levels %d: wait+bar %lld | Bx %lld By %lld bar %lld | D %lld bar %lld | epi %lld  (clk per level)\n", T.wid, k1 - k0,
           prof[0] / (k1 - k0), prof[1] / (k1 - k0), prof[2] / (k1 - k0), prof[3] / (k1 - k0), prof[4] / (k1 - k0), prof[5] / (k1 - k0), prof[6] / (k1 - k0));
#endif
}

// interior tiles are launched as a (tiles, level chunks) grid of one-CTA-per-SM kernels
static inline int level_chunk(int nk) {
  static int kch = 0;
  if (!kch) { const char* e = getenv("FV3_TP2_KCH"); kch = e ? atoi(e) : 8; if (kch < 1) kch = 1; }
  return kch < nk ? kch : nk;
}

}  // namespace tp2
