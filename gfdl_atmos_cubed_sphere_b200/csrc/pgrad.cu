// Pressure-gradient stencils and the A->B (cell-corner) 4th-order interpolation (sm_100a).
//
// Reference semantics: model/a2b_edge.F90 a2b_ord4 (:47-327), extrap_corner (:452-462);
// model/dyn_core.F90 p_grad_c (:1635-1694), nh_p_grad (:1697-1792).
// Design: a2b_ord4 is two launches (1-D cubic sweeps qx/qy incl. the closed-form edge columns,
// then the corner-point combination); the three-way corner extrapolation uses 12 weights
// precomputed at fv3_create instead of calling great_circle_dist on every call.  nh_p_grad
// keeps the interpolated (B-grid) pp, pk, gz, delp in scratch planes rather than writing
// them back in place (the reference's replace=.true.), so the A-grid arrays stay readable by
// neighbouring threads; nothing downstream reads the replaced arrays before they are rebuilt.
#include "fv3_ctx.hpp"
#include <algorithm>
#include <cmath>

#define TI 32
#define TJ 8
#define PLANE_IJK                                              \
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x; \
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;          \
  const int k = blockIdx.z;                                     \
  const long long ko = (long long)k * L.plane;
#define AT(p, i, j) __ldg((p) + ko + LIDX(L, (i), (j)))
#define G2(p, i, j) __ldg((G.p) + LIDX(L, (i), (j)))
static inline dim3 plane_grid(const Lay& L, int nk) { return dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, nk); }

#define FRAME_IJK                                              \
  int bx_, by_;                                                \
  FG.map(blockIdx.x, bx_, by_);                                \
  const int i = L.isd - FV3_IOFF + bx_ * TI + threadIdx.x;     \
  const int j = L.jsd + by_ * TJ + threadIdx.y;                \
  const int k = blockIdx.z;                                    \
  const long long ko = (long long)k * L.plane;

namespace {
constexpr double r3 = 1. / 3.;
constexpr double a1 = 0.5625, a2 = -0.0625;         // a2b_edge.F90:34-35
constexpr double b1 = 7. / 12., b2 = -1. / 12.;     // :39-40
constexpr double c1 = 2. / 3., c2 = -1. / 6.;       // :54-55
}
// Up to four independent a2b_ord4 calls in one launch triple (nh_p_grad interpolates pp, pk3, gz, delp back to back,
// dyn_core.F90:1743-1746 + nh_p_grad): the cube-edge passes are latency-bound launches of a few thousand points per level,
// so four of them in one grid cost about what one does.  Field = blockIdx.y (edge passes) or blockIdx.z / nkmax (interior).
struct A2bBatch { const double* qin[4]; double* qout[4]; double* qx[4]; double* qy[4]; int nk[4]; int nkmax; };

// pass 1: qx (is:ie+1, 1:npy-1) and qy (1:npx-1, js:je+1)   (a2b_edge.F90:132-233)
// frame != 0: only the points the frame outputs of pass 2 read (the interior is done by k_a2b_fused)
__global__ void __launch_bounds__(128) k_a2b_1(Lay L, DevGrid G, FramePts FP, A2bBatch B) {
  int i, j;
  if (!FP.map(blockIdx.x * 128 + threadIdx.x, i, j)) return;
  const int k = blockIdx.z, fld = blockIdx.y;
  if (k >= B.nk[fld]) return;
  const double* __restrict__ qin = B.qin[fld];
  double* __restrict__ qx = B.qx[fld];
  double* __restrict__ qy = B.qy[fld];
  const long long ko = (long long)k * L.plane;
  const int npx = L.npx, npy = L.npy;
  const bool cube = L.cube;
  auto Q = [&](int ii, int jj) { return AT(qin, ii, jj); };
  if (i >= L.is && i <= L.ie + 1 && (!cube || (j >= 1 && j <= npy - 1))) {
    double v;
    auto gen = [&](int ii) { return b2 * (Q(ii - 2, j) + Q(ii + 1, j)) + b1 * (Q(ii - 1, j) + Q(ii, j)); };
    if (!cube || (i >= 3 && i <= npx - 2)) v = gen(i);
    else if (i <= 2) {
      const double g_in = G2(dxa, 2, j) / G2(dxa, 1, j), g_ou = G2(dxa, -1, j) / G2(dxa, 0, j);
      const double q1 = 0.5 * (((2. + g_in) * Q(1, j) - Q(2, j)) / (1. + g_in) + ((2. + g_ou) * Q(0, j) - Q(-1, j)) / (1. + g_ou));
      v = (i == 1) ? q1 : (3. * (g_in * Q(1, j) + Q(2, j)) - (g_in * q1 + gen(3))) / (2. + 2. * g_in);
    } else {
      const double g_in = G2(dxa, npx - 2, j) / G2(dxa, npx - 1, j), g_ou = G2(dxa, npx + 1, j) / G2(dxa, npx, j);
      const double qn = 0.5 * (((2. + g_in) * Q(npx - 1, j) - Q(npx - 2, j)) / (1. + g_in) + ((2. + g_ou) * Q(npx, j) - Q(npx + 1, j)) / (1. + g_ou));
      v = (i == npx) ? qn : (3. * (Q(npx - 2, j) + g_in * Q(npx - 1, j)) - (g_in * qn + gen(npx - 2))) / (2. + 2. * g_in);
    }
    qx[ko + LIDX(L, i, j)] = v;
  }
  if (j >= L.js && j <= L.je + 1 && (!cube || (i >= 1 && i <= npx - 1))) {
    double v;
    auto gen = [&](int jj) { return b2 * (Q(i, jj - 2) + Q(i, jj + 1)) + b1 * (Q(i, jj - 1) + Q(i, jj)); };
    if (!cube || (j >= 3 && j <= npy - 2)) v = gen(j);
    else if (j <= 2) {
      const double g_in = G2(dya, i, 2) / G2(dya, i, 1), g_ou = G2(dya, i, -1) / G2(dya, i, 0);
      const double q1 = 0.5 * (((2. + g_in) * Q(i, 1) - Q(i, 2)) / (1. + g_in) + ((2. + g_ou) * Q(i, 0) - Q(i, -1)) / (1. + g_ou));
      v = (j == 1) ? q1 : (3. * (g_in * Q(i, 1) + Q(i, 2)) - (g_in * q1 + gen(3))) / (2. + 2. * g_in);
    } else {
      const double g_in = G2(dya, i, npy - 2) / G2(dya, i, npy - 1), g_ou = G2(dya, i, npy + 1) / G2(dya, i, npy);
      const double qn = 0.5 * (((2. + g_in) * Q(i, npy - 1) - Q(i, npy - 2)) / (1. + g_in) + ((2. + g_ou) * Q(i, npy) - Q(i, npy + 1)) / (1. + g_ou));
      v = (j == npy) ? qn : (3. * (Q(i, npy - 2) + g_in * Q(i, npy - 1)) - (g_in * qn + gen(npy - 2))) / (2. + 2. * g_in);
    }
    qy[ko + LIDX(L, i, j)] = v;
  }
}

// pass 2: qout on (is:ie+1, js:je+1)   (a2b_edge.F90:104-130, :141-166, :199-224, :258-288)
__global__ void __launch_bounds__(128) k_a2b_2(Lay L, DevGrid G, FramePts FP, A2bBatch B) {
  int i, j;
  if (!FP.map(blockIdx.x * 128 + threadIdx.x, i, j)) return;
  const int k = blockIdx.z, fld = blockIdx.y;
  if (k >= B.nk[fld]) return;
  const double* __restrict__ qin = B.qin[fld];
  const double* __restrict__ qx = B.qx[fld];
  const double* __restrict__ qy = B.qy[fld];
  double* __restrict__ qout = B.qout[fld];
  const long long ko = (long long)k * L.plane;
  const int npx = L.npx, npy = L.npy;
  auto Q = [&](int ii, int jj) { return AT(qin, ii, jj); };
  auto QX = [&](int ii, int jj) { return AT(qx, ii, jj); };
  auto QY = [&](int ii, int jj) { return AT(qy, ii, jj); };
  double out;
  if (!L.cube) {
    out = 0.5 * (a1 * (QX(i, j - 1) + QX(i, j) + QY(i - 1, j) + QY(i, j)) + a2 * (QX(i, j - 2) + QX(i, j + 1) + QY(i - 2, j) + QY(i + 1, j)));
    qout[ko + LIDX(L, i, j)] = out;
    return;
  }
  auto ex = [&](int c, int m, double q1, double q2) { return q1 + G.a2b_w[c][m] * (q1 - q2); };
  // edge values (closed form, also needed by the first interior row/column)
  auto west = [&](int jj) {
    auto q2 = [&](int j2) { return (Q(0, j2) * G2(dxa, 1, j2) + Q(1, j2) * G2(dxa, 0, j2)) / (G2(dxa, 0, j2) + G2(dxa, 1, j2)); };
    const double ew = __ldg(G.edge_w + jj - 1);
    return ew * q2(jj - 1) + (1. - ew) * q2(jj);
  };
  auto east = [&](int jj) {
    auto q2 = [&](int j2) { return (Q(npx - 1, j2) * G2(dxa, npx, j2) + Q(npx, j2) * G2(dxa, npx - 1, j2)) / (G2(dxa, npx - 1, j2) + G2(dxa, npx, j2)); };
    const double ee = __ldg(G.edge_e + jj - 1);
    return ee * q2(jj - 1) + (1. - ee) * q2(jj);
  };
  auto south = [&](int ii) {
    auto q1 = [&](int i2) { return (Q(i2, 0) * G2(dya, i2, 1) + Q(i2, 1) * G2(dya, i2, 0)) / (G2(dya, i2, 0) + G2(dya, i2, 1)); };
    const double es = __ldg(G.edge_s + ii - 1);
    return es * q1(ii - 1) + (1. - es) * q1(ii);
  };
  auto north = [&](int ii) {
    auto q1 = [&](int i2) { return (Q(i2, npy - 1) * G2(dya, i2, npy) + Q(i2, npy) * G2(dya, i2, npy - 1)) / (G2(dya, i2, npy - 1) + G2(dya, i2, npy)); };
    const double en = __ldg(G.edge_n + ii - 1);
    return en * q1(ii - 1) + (1. - en) * q1(ii);
  };
  const bool ic = (i == 1 || i == npx), jc = (j == 1 || j == npy);
  if (ic && jc) {
    if (i == 1 && j == 1)
      out = (ex(0, 0, Q(1, 1), Q(2, 2)) + ex(0, 1, Q(0, 1), Q(-1, 2)) + ex(0, 2, Q(1, 0), Q(2, -1))) * r3;
    else if (i == npx && j == 1)
      out = (ex(1, 0, Q(npx - 1, 1), Q(npx - 2, 2)) + ex(1, 1, Q(npx - 1, 0), Q(npx - 2, -1)) + ex(1, 2, Q(npx, 1), Q(npx + 1, 2))) * r3;
    else if (i == npx && j == npy)
      out = (ex(2, 0, Q(npx - 1, npy - 1), Q(npx - 2, npy - 2)) + ex(2, 1, Q(npx, npy - 1), Q(npx + 1, npy - 2)) +
             ex(2, 2, Q(npx - 1, npy), Q(npx - 2, npy + 1))) * r3;
    else
      out = (ex(3, 0, Q(1, npy - 1), Q(2, npy - 2)) + ex(3, 1, Q(0, npy - 1), Q(-1, npy - 2)) + ex(3, 2, Q(1, npy), Q(2, npy + 1))) * r3;
  } else if (i == 1) out = west(j);
  else if (i == npx) out = east(j);
  else if (j == 1) out = south(i);
  else if (j == npy) out = north(i);
  else {
    auto qxx_gen = [&](int jj) { return a2 * (QX(i, jj - 2) + QX(i, jj + 1)) + a1 * (QX(i, jj - 1) + QX(i, jj)); };
    auto qyy_gen = [&](int ii) { return a2 * (QY(ii - 2, j) + QY(ii + 1, j)) + a1 * (QY(ii - 1, j) + QY(ii, j)); };
    double qxx, qyy;
    if (j == 2) qxx = c1 * (QX(i, 1) + QX(i, 2)) + c2 * (south(i) + qxx_gen(3));
    else if (j == npy - 1) qxx = c1 * (QX(i, npy - 2) + QX(i, npy - 1)) + c2 * (north(i) + qxx_gen(npy - 2));
    else qxx = qxx_gen(j);
    if (i == 2) qyy = c1 * (QY(1, j) + QY(2, j)) + c2 * (west(j) + qyy_gen(3));
    else if (i == npx - 1) qyy = c1 * (QY(npx - 2, j) + QY(npx - 1, j)) + c2 * (east(j) + qyy_gen(npx - 2));
    else qyy = qyy_gen(i);
    out = 0.5 * (qxx + qyy);
  }
  qout[ko + LIDX(L, i, j)] = out;
}

// Interior of a2b_ord4 in ONE kernel: outputs (i, j) in [3, npx-2]^2 use only the generic 4th-order formulas
// (a2b_edge.F90:168-170, :226-228, :258-288 interior branch), a 4x4 stencil of qin that never leaves the face.  A CTA
// stages a (32+3) x (16+3) tile of qin, forms qx and qy in shared memory (same expressions as the two-pass kernels) and
// combines them; the intermediates never touch HBM (the two-pass form moved ~5x the field).
#define AF_TX 32
#define AF_TY 16
__global__ void __launch_bounds__(256) k_a2b_fused(Lay L, A2bBatch B) {
  __shared__ double q[AF_TY + 3][AF_TX + 4];
  __shared__ double qx[AF_TY + 3][AF_TX];
  __shared__ double qy[AF_TY][AF_TX + 4];
  const int fld = blockIdx.z / B.nkmax, kk = blockIdx.z - fld * B.nkmax;
  if (kk >= B.nk[fld]) return;
  const double* __restrict__ qin = B.qin[fld];
  double* __restrict__ qout = B.qout[fld];
  const int i0 = 3 + blockIdx.x * AF_TX, j0 = 3 + blockIdx.y * AF_TY;   // first output point of the tile
  const long long ko = (long long)kk * L.plane;
  const int tid = threadIdx.x, hi = L.npx - 2;                          // last output index in both directions
  for (int e = tid; e < (AF_TY + 3) * (AF_TX + 3); e += 256) {
    const int r = e / (AF_TX + 3), c = e - r * (AF_TX + 3);
    q[r][c] = __ldg(qin + ko + LIDX(L, min(i0 - 2 + c, L.ie), min(j0 - 2 + r, L.je)));
  }
  __syncthreads();
  for (int e = tid; e < (AF_TY + 3) * AF_TX; e += 256) {   // qx(i0+c, j0-2+r)
    const int r = e / AF_TX, c = e - r * AF_TX;
    qx[r][c] = b2 * (q[r][c] + q[r][c + 3]) + b1 * (q[r][c + 1] + q[r][c + 2]);
  }
  for (int e = tid; e < AF_TY * (AF_TX + 3); e += 256) {   // qy(i0-2+c, j0+r)
    const int r = e / (AF_TX + 3), c = e - r * (AF_TX + 3);
    qy[r][c] = b2 * (q[r][c] + q[r + 3][c]) + b1 * (q[r + 1][c] + q[r + 2][c]);
  }
  __syncthreads();
  const int c = tid & 31;
  for (int r = tid >> 5; r < AF_TY; r += 8) {
    const int i = i0 + c, j = j0 + r;
    if (i > hi || j > hi) continue;
    const double qxx = a2 * (qx[r][c] + qx[r + 3][c]) + a1 * (qx[r + 1][c] + qx[r + 2][c]);
    const double qyy = a2 * (qy[r][c] + qy[r][c + 3]) + a1 * (qy[r][c + 1] + qy[r][c + 2]);
    qout[ko + LIDX(L, i, j)] = 0.5 * (qxx + qyy);
  }
}

// n <= 4 fields; qx, qy scratch of field f = c->scr[scr0 + 2 f], c->scr[scr0 + 2 f + 1]
int launch_a2b_ord4_batch(fv3_ctx* c, int n, const double* const* qin, double* const* qout, const int* nk, int scr0) {
  const Lay& L = c->L;
  if (n < 1 || n > 4 || scr0 + 2 * n > fv3_ctx::NSCR) return fv3_fail(c, -1, "a2b_ord4 batch: bad field count / scratch range");
  A2bBatch B{};
  B.nkmax = 0;
  for (int f = 0; f < n; f++) {
    B.qin[f] = qin[f]; B.qout[f] = qout[f]; B.qx[f] = c->scr[scr0 + 2 * f]; B.qy[f] = c->scr[scr0 + 2 * f + 1]; B.nk[f] = nk[f];
    B.nkmax = std::max(B.nkmax, nk[f]);
  }
  const int ni = L.npx - 4;   // outputs 3 .. npx-2
  const int fused = L.cube && ni >= 1 && L.npx == L.npy;
  if (fused) {
    dim3 g((ni + AF_TX - 1) / AF_TX, (ni + AF_TY - 1) / AF_TY, B.nkmax * n);
    k_a2b_fused<<<g, 256, 0, c->stream>>>(L, B);
    c->launches++;
  }
  // edge passes, one thread per point: pass 1 on (is-2:ie+2)^2 where i <= 5 || i >= npx-4 (same in j) -- what the
  // frame outputs read --, pass 2 on (is:ie+1)^2 where i <= 2 || i >= npx-1 (same in j); everything when not fused
  const FramePts P1 = fused ? frame_pts(L.is - 2, L.ie + 2, L.js - 2, L.je + 2, 6, L.npx - 5, 6, L.npy - 5)
                            : frame_pts(L.is - 2, L.ie + 2, L.js - 2, L.je + 2, 1, 0, 1, 0);
  const FramePts P2 = fused ? frame_pts(L.is, L.ie + 1, L.js, L.je + 1, 3, L.npx - 2, 3, L.npy - 2)
                            : frame_pts(L.is, L.ie + 1, L.js, L.je + 1, 1, 0, 1, 0);
  k_a2b_1<<<dim3((P1.count() + 127) / 128, n, B.nkmax), 128, 0, c->stream>>>(L, c->G, P1, B);
  k_a2b_2<<<dim3((P2.count() + 127) / 128, n, B.nkmax), 128, 0, c->stream>>>(L, c->G, P2, B);
  c->launches += 2;
  return 0;
}
// single field; qx, qy scratch = c->scr[4], c->scr[5] (free in both callers: d_sw between transports, one_grad_p)
int launch_a2b_ord4(fv3_ctx* c, const double* qin, double* qout, int nk, int /*replace_into_qin*/) {
  return launch_a2b_ord4_batch(c, 1, &qin, &qout, &nk, 4);
}

// ---- p_grad_c (dyn_core.F90:1635-1694), in place on uc, vc ---------------------------------
__global__ void __launch_bounds__(TI* TJ) k_pgrad_c(Lay L, DevGrid G, const double* __restrict__ delpc, const double* __restrict__ pkc,
                                                   const double* __restrict__ gz, double* __restrict__ uc, double* __restrict__ vc,
                                                   double dt2, int hydrostatic) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const long long P = L.plane;
  const long long o = ko + LIDX(L, i, j);
  auto WK = [&](long long oo) { return hydrostatic ? (__ldg(pkc + oo + P) - __ldg(pkc + oo)) : __ldg(delpc + oo); };
  auto GZ = [&](long long oo) { return __ldg(gz + oo); };
  auto PK = [&](long long oo) { return __ldg(pkc + oo); };
  if (j <= L.je) {
    const long long w = o - 1;
    uc[o] = uc[o] + dt2 * G2(rdxc, i, j) / (WK(w) + WK(o)) *
                        ((GZ(w + P) - GZ(o)) * (PK(o + P) - PK(w)) + (GZ(w) - GZ(o + P)) * (PK(w + P) - PK(o)));
  }
  if (i <= L.ie) {
    const long long s = o - L.NI;
    vc[o] = vc[o] + dt2 * G2(rdyc, i, j) / (WK(s) + WK(o)) *
                        ((GZ(s + P) - GZ(o)) * (PK(o + P) - PK(s)) + (GZ(s) - GZ(o + P)) * (PK(s + P) - PK(o)));
  }
}
int stage_p_grad_c(fv3_ctx* c, double dt2) {
  StageScope ts(c, "PG_C");
  const Lay& L = c->L;
  dim3 blk(TI, TJ), grd = plane_grid(L, L.npz);
  k_pgrad_c<<<grd, blk, 0, c->stream>>>(L, c->G, c->fld[FV3_DELPC], c->fld[FV3_PKC], c->fld[FV3_GZ], c->fld[FV3_UC], c->fld[FV3_VC], dt2,
                                        c->f.hydrostatic);
  c->launches++;
  return 0;
}

// ---- nh_p_grad (dyn_core.F90:1697-1792) -----------------------------------------------------
__global__ void __launch_bounds__(TI* TJ) k_nh_top(Lay L, double* __restrict__ ppb, double* __restrict__ pkb, double top_value) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  ppb[LIDX(L, i, j)] = 0.;
  pkb[LIDX(L, i, j)] = top_value;
}
// SPLIT: split_p_grad (dyn_core.F90:1795-1905, beta > 0): u first takes beta times the hydrostatic increment of the PREVIOUS substep
// (du, dv), the current one is stored and enters with weight alpha = 1 - beta
template <bool SPLIT>
__global__ void __launch_bounds__(TI* TJ) k_nh_pgrad(Lay L, DevGrid G, const double* __restrict__ pp, const double* __restrict__ pk,
                                                    const double* __restrict__ gz, const double* __restrict__ dpb, double* __restrict__ u,
                                                    double* __restrict__ v, double* __restrict__ du, double* __restrict__ dv, double beta,
                                                    double dt) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const long long P = L.plane;
  const long long o = ko + LIDX(L, i, j);
  const double alpha = 1. - beta;
  auto WKo = [&](long long oo) { return __ldg(pk + oo + P) - __ldg(pk + oo); };
  if (i <= L.ie) {
    const long long e = o + 1;
    const double du1 = dt / (WKo(o) + WKo(e)) *
                       ((__ldg(gz + o + P) - __ldg(gz + e)) * (__ldg(pk + e + P) - __ldg(pk + o)) +
                        (__ldg(gz + o) - __ldg(gz + e + P)) * (__ldg(pk + o + P) - __ldg(pk + e)));
    const double nh = dt / (__ldg(dpb + o) + __ldg(dpb + e)) *
                      ((__ldg(gz + o + P) - __ldg(gz + e)) * (__ldg(pp + e + P) - __ldg(pp + o)) +
                       (__ldg(gz + o) - __ldg(gz + e + P)) * (__ldg(pp + o + P) - __ldg(pp + e)));
    if (SPLIT) {
      const double u1 = u[o] + beta * du[o];
      du[o] = du1;
      u[o] = (u1 + alpha * du1 + nh) * G2(rdx, i, j);
    } else u[o] = (u[o] + du1 + nh) * G2(rdx, i, j);
  }
  if (j <= L.je) {
    const long long n = o + L.NI;
    const double dv1 = dt / (WKo(o) + WKo(n)) *
                       ((__ldg(gz + o + P) - __ldg(gz + n)) * (__ldg(pk + n + P) - __ldg(pk + o)) +
                        (__ldg(gz + o) - __ldg(gz + n + P)) * (__ldg(pk + o + P) - __ldg(pk + n)));
    const double nh = dt / (__ldg(dpb + o) + __ldg(dpb + n)) *
                      ((__ldg(gz + o + P) - __ldg(gz + n)) * (__ldg(pp + n + P) - __ldg(pp + o)) +
                       (__ldg(gz + o) - __ldg(gz + n + P)) * (__ldg(pp + o + P) - __ldg(pp + n)));
    if (SPLIT) {
      const double v1 = v[o] + beta * dv[o];
      dv[o] = dv1;
      v[o] = (v1 + alpha * dv1 + nh) * G2(rdy, i, j);
    } else v[o] = (v[o] + dv1 + nh) * G2(rdy, i, j);
  }
}

// beta_d < 0: nh_p_grad (dyn_core.F90:1032); beta_d >= 0: split_p_grad with that beta (:1028; 0 on the first substep, :404-406)
int stage_nh_p_grad(fv3_ctx* c, double dt, double beta_d) {
  StageScope ts(c, "PG_D");
  const Lay& L = c->L;
  const int km = L.npz;
  double *ppb = c->scr[0], *pkb = c->scr[1], *gzb = c->scr[2], *dpb = c->scr[3];
  const long long P = L.plane;
  const double top_value = c->f.use_logp ? log(c->f.ptop) : pow(c->f.ptop, c->f.kappa);   // peln1 / ptk, dyn_core.F90:220-222
  dim3 blk(TI, TJ), g1 = plane_grid(L, 1);
  k_nh_top<<<g1, blk, 0, c->stream>>>(L, ppb, pkb, top_value);
  c->launches++;
  int rc;
  {  // pp, pk (k = 2..km+1), gz (k = 1..km+1), delp -> wk1: four a2b_ord4 calls in one launch triple; scratch scr[4..11]
    const double* qin[4] = {c->fld[FV3_PKC] + P, c->fld[FV3_PK3] + P, c->fld[FV3_GZ], c->fld[FV3_DELP]};
    double* qout[4] = {ppb + P, pkb + P, gzb, dpb};
    const int nks[4] = {km, km, km + 1, km};
    if ((rc = launch_a2b_ord4_batch(c, 4, qin, qout, nks, 4))) return rc;
  }
  if (beta_d >= 0.) k_nh_pgrad<true><<<plane_grid(L, km), blk, 0, c->stream>>>(L, c->G, ppb, pkb, gzb, dpb, c->fld[FV3_U], c->fld[FV3_V], c->fld[FV3_DU], c->fld[FV3_DV], beta_d, dt);
  else k_nh_pgrad<false><<<plane_grid(L, km), blk, 0, c->stream>>>(L, c->G, ppb, pkb, gzb, dpb, c->fld[FV3_U], c->fld[FV3_V], nullptr, nullptr, 0., dt);
  c->launches++;
  return 0;
}

// ---- one_grad_p (dyn_core.F90:1909-2030), hydrostatic call, d_ext = 0 -------------------------------------------------
// pk (= pe^kappa) and gz are interpolated to the cell corners into scratch planes (the reference replaces them in
// place; nothing reads them before the next geopk rebuilds them), wk = pk(k+1) - pk(k) at the corners.
// GRAD1: grad1_p_update (dyn_core.F90:2033-2116, hydrostatic, beta > 0, d_ext = 0 so divg2 = 0): the beta-weighted increment of the
// previous substep first, the current one stored in du, dv and applied with weight 1 - beta
template <bool GRAD1>
__global__ void __launch_bounds__(TI* TJ) k_one_grad_p(Lay L, DevGrid G, const double* __restrict__ pk, const double* __restrict__ gz,
                                                      double* __restrict__ u, double* __restrict__ v, double* __restrict__ du,
                                                      double* __restrict__ dv, double beta, double dt, const double* __restrict__ divg2) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const long long P = L.plane;
  const int o2 = LIDX(L, i, j);
  const long long o = ko + o2;
  const double alpha = 1. - beta;
  // external-mode damping (d_ext > 0, dyn_core.F90:1969-1984 / :2102, 2111): differences of the 2-D corner field divg2
  const double d0 = divg2 ? __ldg(divg2 + o2) : 0.;
  auto WK = [&](long long oo) { return __ldg(pk + oo + P) - __ldg(pk + oo); };
  if (i <= L.ie) {
    const long long e = o + 1;
    const double d1 = dt / (WK(o) + WK(e)) *
                      ((__ldg(gz + o + P) - __ldg(gz + e)) * (__ldg(pk + e + P) - __ldg(pk + o)) +
                       (__ldg(gz + o) - __ldg(gz + e + P)) * (__ldg(pk + o + P) - __ldg(pk + e)));
    if (GRAD1) {
      const double u1 = u[o] + beta * du[o];
      du[o] = d1;
      const double dE = divg2 ? __ldg(divg2 + o2 + 1) : 0.;
      u[o] = (u1 + d0 - dE + alpha * d1) * G2(rdx, i, j);
    } else {
      const double wk2 = divg2 ? d0 - __ldg(divg2 + o2 + 1) : 0.;
      u[o] = G2(rdx, i, j) * (wk2 + u[o] + d1);
    }
  }
  if (j <= L.je) {
    const long long n = o + L.NI;
    const double d1 = dt / (WK(o) + WK(n)) *
                      ((__ldg(gz + o + P) - __ldg(gz + n)) * (__ldg(pk + n + P) - __ldg(pk + o)) +
                       (__ldg(gz + o) - __ldg(gz + n + P)) * (__ldg(pk + o + P) - __ldg(pk + n)));
    if (GRAD1) {
      const double v1 = v[o] + beta * dv[o];
      dv[o] = d1;
      const double dN = divg2 ? __ldg(divg2 + o2 + L.NI) : 0.;
      v[o] = (v1 + d0 - dN + alpha * d1) * G2(rdy, i, j);
    } else {
      const double wk1 = divg2 ? d0 - __ldg(divg2 + o2 + L.NI) : 0.;
      v[o] = G2(rdy, i, j) * (wk1 + v[o] + d1);
    }
  }
}
__global__ void __launch_bounds__(TI* TJ) k_set_top(Lay L, double* __restrict__ pkb, double top_value) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  pkb[LIDX(L, i, j)] = top_value;
}
// beta_d < 0: one_grad_p (dyn_core.F90:1021); beta_d >= 0: grad1_p_update with that beta (:1019)
int stage_one_grad_p(fv3_ctx* c, double dt, double beta_d) {
  StageScope ts(c, "PG_D");
  const Lay& L = c->L;
  const int km = L.npz;
  if (!c->f.hydrostatic) return fv3_fail(c, -2, "one_grad_p: only the hydrostatic call is supported (non-hydrostatic uses nh_p_grad)");
  const double* divg2 = nullptr;
  if (c->f.d_ext > 0.0) {
    if (!c->d_divg2) return fv3_fail(c, -1, "one_grad_p: d_ext > 0 needs fv3_ext_mode_prepare / fv3_ext_mode_divg2 around d_sw of this substep");
    divg2 = c->d_divg2;
  }
  double *pkb = c->scr[1], *gzb = c->scr[2];
  const long long P = L.plane;
  dim3 blk(TI, TJ);
  k_set_top<<<plane_grid(L, 1), blk, 0, c->stream>>>(L, pkb, pow(c->f.ptop, c->f.kappa));   // ptk, dyn_core.F90:222,1944
  c->launches++;
  int rc;
  {  // pk (k = 2..km+1) and gz (k = 1..km+1) in one launch triple; scratch scr[4..7]
    const double* qin[2] = {c->fld[FV3_PKC] + P, c->fld[FV3_GZ]};
    double* qout[2] = {pkb + P, gzb};
    const int nks[2] = {km, km + 1};
    if ((rc = launch_a2b_ord4_batch(c, 2, qin, qout, nks, 4))) return rc;
  }
  if (beta_d >= 0.) k_one_grad_p<true><<<plane_grid(L, km), blk, 0, c->stream>>>(L, c->G, pkb, gzb, c->fld[FV3_U], c->fld[FV3_V], c->fld[FV3_DU], c->fld[FV3_DV], beta_d, dt, divg2);
  else k_one_grad_p<false><<<plane_grid(L, km), blk, 0, c->stream>>>(L, c->G, pkb, gzb, c->fld[FV3_U], c->fld[FV3_V], nullptr, nullptr, 0., dt, divg2);
  c->launches++;
  return 0;
}

// ---- external-mode divergence damping (d_ext > 0; hydrostatic branch of dyn_core) -----------------------------------------
// a2b_edge.F90:329-450 a2b_ord2 (no replace): A-grid -> cell corners (is:ie+1, js:je+1), all levels in one launch
__global__ void __launch_bounds__(TI* TJ) k_a2b_ord2(Lay L, DevGrid G, const double* __restrict__ qin, double* __restrict__ qout) {
  PLANE_IJK
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const int npx = L.npx, npy = L.npy;
  const long long o = ko + LIDX(L, i, j);
  auto Q = [&](int ii, int jj) { return __ldg(qin + ko + LIDX(L, ii, jj)); };
  const double r3 = 1. / 3.;
  double out;
  if (!L.cube) out = 0.25 * (Q(i - 1, j - 1) + Q(i, j - 1) + Q(i - 1, j) + Q(i, j));
  else if (i == 1 && j == 1) out = r3 * (Q(1, 1) + Q(1, 0) + Q(0, 1));
  else if (i == npx && j == 1) out = r3 * (Q(npx - 1, 1) + Q(npx - 1, 0) + Q(npx, 1));
  else if (i == npx && j == npy) out = r3 * (Q(npx - 1, npy - 1) + Q(npx, npy - 1) + Q(npx - 1, npy));
  else if (i == 1 && j == npy) out = r3 * (Q(1, npy - 1) + Q(0, npy - 1) + Q(1, npy));
  else if (i == 1 || i == npx) {
    const int ia = i == 1 ? 0 : npx - 1;
    const double ew = __ldg((i == 1 ? G.edge_w : G.edge_e) + j - 1);
    const double qa = 0.5 * (Q(ia, j - 1) + Q(ia + 1, j - 1)), qb = 0.5 * (Q(ia, j) + Q(ia + 1, j));
    out = ew * qa + (1. - ew) * qb;
  } else if (j == 1 || j == npy) {
    const int ja = j == 1 ? 0 : npy - 1;
    const double es = __ldg((j == 1 ? G.edge_s : G.edge_n) + i - 1);
    const double qa = 0.5 * (Q(i - 1, ja) + Q(i - 1, ja + 1)), qb = 0.5 * (Q(i, ja) + Q(i, ja + 1));
    out = es * qa + (1. - es) * qb;
  } else out = 0.25 * (Q(i - 1, j - 1) + Q(i, j - 1) + Q(i - 1, j) + Q(i, j));
  qout[o] = out;
}
// dyn_core.F90:828-847: divg2 = d_ext * da_min_c * sum_k(dpc * divg) / sum_k(dpc) at the cell corners
__global__ void __launch_bounds__(TI* TJ) k_ext_divg2(Lay L, const double* __restrict__ dpc, const double* __restrict__ divg, double* __restrict__ divg2,
                                                      double d2_divg) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  if (i < L.is || i > L.ie + 1 || j < L.js || j > L.je + 1) return;
  const int o2 = LIDX(L, i, j);
  double wk = __ldg(dpc + o2), d = wk * __ldg(divg + o2);
  for (int k = 1; k < L.npz; k++) {
    const double p = __ldg(dpc + o2 + (long long)k * L.plane);
    wk = wk + p;
    d = d + p * __ldg(divg + o2 + (long long)k * L.plane);
  }
  divg2[o2] = d2_divg * d / wk;
}
// before d_sw (:745-747): delp at the corners -> FV3_PTC (dead between p_grad_c and the next c_sw; the reference parks it there too, :791-797)
int stage_ext_mode_prepare(fv3_ctx* c) {
  StageScope ts(c, "EXT_MODE");
  const Lay& L = c->L;
  k_a2b_ord2<<<plane_grid(L, L.npz), dim3(TI, TJ), 0, c->stream>>>(L, c->G, c->fld[FV3_DELP], c->fld[FV3_PTC]);
  c->launches++;
  return 0;
}
// after d_sw, which left its divergence (delpc, sw_core.F90:1366 / :1379) in FV3_VT
int stage_ext_mode_divg2(fv3_ctx* c) {
  StageScope ts(c, "EXT_MODE");
  const Lay& L = c->L;
  if (!c->d_divg2) {
    FV3_CUDA(c, cudaMalloc(&c->d_divg2, (size_t)L.plane * sizeof(double)));
    FV3_CUDA(c, cudaMemsetAsync(c->d_divg2, 0, (size_t)L.plane * sizeof(double), c->stream));   // on the context stream: a legacy-stream memset could land after the kernel below
  }
  k_ext_divg2<<<dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ), dim3(TI, TJ), 0, c->stream>>>(L, c->fld[FV3_PTC], c->fld[FV3_VT], c->d_divg2,
                                                                                              c->f.d_ext * c->G.da_min_c);
  c->launches++;
  return 0;
}

// ---- omega diagnostic of the last substep of the last dyn_core call (end_step; use_old_omega = T) ------------------------------
// dyn_core.F90:409-422: pem = ptop + partial sums of delp on (is-1:ie+1, js-1:je+1), one thread per column
__global__ void __launch_bounds__(TI* TJ) k_pem(Lay L, const double* __restrict__ delp, double* __restrict__ pem, double ptop) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  if (i < L.is - 1 || i > L.ie + 1 || j < L.js - 1 || j > L.je + 1) return;
  const int o2 = LIDX(L, i, j);
  double p = ptop;
  pem[o2] = p;
  for (int k = 0; k < L.npz; k++) { p = p + __ldg(delp + o2 + (long long)k * L.plane); pem[o2 + (long long)(k + 1) * L.plane] = p; }
}
// dyn_core.F90:1182-1195 + adv_pe (:1529-1630): omga = (pe - pem) * rdt + 0.5 * rarea * V3 . grad(pem); pb = a2b_ord2 of the
// interfaces 2..npz+1 of pem (plane k of pb <-> interface k+2)
__global__ void __launch_bounds__(TI* TJ) k_omega_old(Lay L, DevGrid G, const double* __restrict__ pe, const double* __restrict__ pem,
                                                      const double* __restrict__ pb, const double* __restrict__ ua, const double* __restrict__ va,
                                                      double* __restrict__ omga, double rdt) {
  PLANE_IJK
  if (i < L.is || i > L.ie || j < L.js || j > L.je) return;
  const int o2 = LIDX(L, i, j);
  const long long o = ko + o2, P = L.plane;
  double om = (__ldg(pe + o + P) - __ldg(pem + o + P)) * rdt;
  double up, vp;
  if (k == L.npz - 1) { up = __ldg(ua + o); vp = __ldg(va + o); }
  else { up = 0.5 * (__ldg(ua + o) + __ldg(ua + o + P)); vp = 0.5 * (__ldg(va + o) + __ldg(va + o + P)); }
  const double b00 = __ldg(pb + o), b10 = __ldg(pb + o + 1), b01 = __ldg(pb + o + L.NI), b11 = __ldg(pb + o + L.NI + 1);
  const double dx0 = G2(dx, i, j), dx1 = G2(dx, i, j + 1), dy0 = G2(dy, i, j), dy1 = G2(dy, i + 1, j);
  double acc = 0.;
#pragma unroll
  for (int n = 0; n < 3; n++) {
    const long long no = (long long)n * P;
    const double v3 = up * __ldg(G.ec1 + no + o2) + vp * __ldg(G.ec2 + no + o2);
    const double pdx0 = (b00 + b10) * dx0 * __ldg(G.en1 + no + o2);
    const double pdx1 = (b01 + b11) * dx1 * __ldg(G.en1 + no + o2 + L.NI);
    const double pdy0 = (b00 + b01) * dy0 * __ldg(G.en2 + no + o2);
    const double pdy1 = (b10 + b11) * dy1 * __ldg(G.en2 + no + o2 + 1);
    const double grad = pdx1 - pdx0 - pdy0 + pdy1;
    acc = n == 0 ? v3 * grad : acc + v3 * grad;
  }
  omga[o] = om + 0.5 * G2(rarea, i, j) * acc;
}
// use_old_omega = F (dyn_core.F90:735-742, 774-781, 1196-1214): phase 0 before d_sw: omga = delp; phase 1 after d_sw: times the
// convergence of the area fluxes / dt; phase 2 at the end of the substep: running sum over k
__global__ void __launch_bounds__(TI* TJ) k_omega_new(Lay L, DevGrid G, double* __restrict__ omga, const double* __restrict__ delp,
                                                      const double* __restrict__ xfx, const double* __restrict__ yfx, int phase, double rdt) {
  PLANE_IJK
  if (i < L.is || i > L.ie || j < L.js || j > L.je) return;
  const long long o = ko + LIDX(L, i, j);
  if (phase == 0) { omga[o] = __ldg(delp + o); return; }
  if (phase == 1) { omga[o] = omga[o] * (__ldg(xfx + o) - __ldg(xfx + o + 1) + __ldg(yfx + o) - __ldg(yfx + o + L.NI)) * G2(rarea, i, j) * rdt; return; }
  if (k != 0) return;   // phase 2: one thread per column
  double om = omga[o];
  for (int kk = 1; kk < L.npz; kk++) { om = om + omga[o + (long long)kk * L.plane]; omga[o + (long long)kk * L.plane] = om; }
}
int stage_omega_new(fv3_ctx* c, int phase, double dt) {
  StageScope ts(c, "OMEGA");
  const Lay& L = c->L;
  if (phase < 0 || phase > 2) return fv3_fail(c, -1, "omega_new: phase in 0..2");
  k_omega_new<<<plane_grid(L, phase == 2 ? 1 : L.npz), dim3(TI, TJ), 0, c->stream>>>(L, c->G, c->fld[FV3_OMGA], c->fld[FV3_DELP], c->fld[FV3_XFX],
                                                                                    c->fld[FV3_YFX], phase, 1. / dt);
  c->launches++;
  return 0;
}
int stage_omega_begin(fv3_ctx* c) {
  StageScope ts(c, "OMEGA");
  const Lay& L = c->L;
  if (!c->f.use_old_omega) return 0;
  if (!c->d_pem) FV3_CUDA(c, cudaMalloc(&c->d_pem, (size_t)L.plane * (L.npz + 1) * sizeof(double)));
  k_pem<<<dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ), dim3(TI, TJ), 0, c->stream>>>(L, c->fld[FV3_DELP], c->d_pem, c->f.ptop);
  c->launches++;
  return 0;
}
int stage_omega_end(fv3_ctx* c, double dt) {
  StageScope ts(c, "OMEGA");
  const Lay& L = c->L;
  if (!c->f.use_old_omega) return stage_omega_new(c, 2, dt);
  if (!c->d_pem) return fv3_fail(c, -1, "omega_end without omega_begin");
  if (!c->G.ec1 || !c->G.ec2 || !c->G.en1 || !c->G.en2) return fv3_fail(c, -1, "omega diagnostic: fv3_grid_t.ec1 / ec2 / en1 / en2 were not given to fv3_create");
  double* pb = c->scr[0];
  k_a2b_ord2<<<plane_grid(L, L.npz), dim3(TI, TJ), 0, c->stream>>>(L, c->G, c->d_pem + L.plane, pb);
  k_omega_old<<<plane_grid(L, L.npz), dim3(TI, TJ), 0, c->stream>>>(L, c->G, c->fld[FV3_PE], c->d_pem, pb, c->fld[FV3_UA], c->fld[FV3_VA],
                                                                  c->fld[FV3_OMGA], 1. / dt);
  c->launches += 2;
  return 0;
}
