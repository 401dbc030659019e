// tracer_2d_1L on the device: sub-cycled Lin-Rood transport of a tracer with the mass fluxes and Courant numbers that the
// acoustic loop accumulated ("pure re-use of the fv_tp_2d kernel ... adds the mp_reduce_max CFL all-reduce").
//
// Reference semantics: model/fv_tracer2d.F90:49-295 (tracer_2d_1L), nq = 1, trdm = 0 (no del-n damping), id_divg_mean = 0.
// Design: one prep kernel per face (area fluxes from cx, cy + per-level CFL maximum by block reduction and an atomic max
// on the bit pattern), the maxima of the faces / ranks are combined (host max over the faces of the process,
// ncclAllReduce(max) over ranks), one scale kernel, then per sub-cycle ONE halo exchange of the tracer and ONE launch pair
// (interior / frame tiles) of the shared-memory fv_tp_2d tile kernel whose epilogue applies
//   dp2 = dp1 + div(mfx, mfy)*rarea,   q = (q*dp1 + div(fx, fy)*rarea) / dp2
// so the fluxes never reach HBM.  Levels whose sub-cycle count is exhausted copy the tracer through (the tracer is
// ping-ponged between FV3_WORK_Q and a scratch field because a tile reads its neighbours' halos).
#include "tp2d.cuh"
#include "ppm.cuh"
#include "tp_tile.cuh"
#include <cmath>
#include <vector>

using namespace ppm;

int halo_allreduce_max(fv3_ctx* c, double* vals, int n);
extern "C" int fv3_halo_exchange(fv3_ctx** ctxs, int nctx, int group);

#define TI 32
#define TJ 8
static inline dim3 plane_grid(const Lay& L, int nk) { return dim3((L.NI + TI - 1) / TI, (L.NJ + TJ - 1) / TJ, nk); }
#define G2(p, i, j) __ldg((G.p) + LIDX(L, (i), (j)))
#define SG(n, i, j) __ldg(G.sin_sg + (long long)((n)-1) * L.plane + LIDX(L, (i), (j)))

// fv_tracer2d.F90:110-145: xfx, yfx from the accumulated Courant numbers; cmax(k)
__global__ void __launch_bounds__(TI* TJ) k_tr_prep(Lay L, DevGrid G, const double* __restrict__ cx, const double* __restrict__ cy,
                                                   double* __restrict__ xfx, double* __restrict__ yfx, unsigned long long* __restrict__ cmax) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  const int k = blockIdx.z;
  const long long ko = (long long)k * L.plane;
  double m = 0.;
  if (i >= L.isd && i <= L.ied + 1 && j <= L.jed + 1) {
    const long long o = ko + LIDX(L, i, j);
    if (j <= L.jed && i >= L.is && i <= L.ie + 1) {
      const double c = cx[o];
      xfx[o] = (c > 0.) ? c * G2(dxa, i - 1, j) * G2(dy, i, j) * SG(3, i - 1, j) : c * G2(dxa, i, j) * G2(dy, i, j) * SG(1, i, j);
    }
    if (i <= L.ied && j >= L.js && j <= L.je + 1) {
      const double c = cy[o];
      yfx[o] = (c > 0.) ? c * G2(dya, i, j - 1) * G2(dx, i, j) * SG(4, i, j - 1) : c * G2(dya, i, j) * G2(dx, i, j) * SG(2, i, j);
    }
    if (i >= L.is && i <= L.ie && j >= L.js && j <= L.je) {
      const double a = fmax(fabs(cx[o]), fabs(cy[o]));
      m = ((k + 1) < L.npz / 6) ? a : a + 1. - SG(5, i, j);   // 1-based k < npz/6 (:133-144)
    }
  }
  // block maximum (values are >= 0: the unsigned bit pattern orders like the value)
  __shared__ double red[TI * TJ / 32];
  for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
  const int t = threadIdx.y * TI + threadIdx.x;
  if ((t & 31) == 0) red[t >> 5] = m;
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < TI * TJ / 32; w++) m = fmax(m, red[w]);
    if (m > 0.) atomicMax(cmax + k, (unsigned long long)__double_as_longlong(m));
  }
}

// fv_tracer2d.F90:165-192: cx, xfx, cy, yfx, mfx, mfy *= frac(k) where nsplt(k) > 1
__global__ void __launch_bounds__(TI* TJ) k_tr_scale(Lay L, double* __restrict__ cx, double* __restrict__ xfx, double* __restrict__ cy,
                                                    double* __restrict__ yfx, double* __restrict__ mfx, double* __restrict__ mfy,
                                                    const double* __restrict__ frac) {
  const int i = L.isd - FV3_IOFF + blockIdx.x * TI + threadIdx.x;
  const int j = L.jsd + blockIdx.y * TJ + threadIdx.y;
  const int k = blockIdx.z;
  const double f = frac[k];
  if (f == 1.) return;
  if (i < L.isd || i > L.ied + 1 || j > L.jed + 1) return;
  const long long o = (long long)k * L.plane + LIDX(L, i, j);
  if (j <= L.jed && i >= L.is && i <= L.ie + 1) { cx[o] = cx[o] * f; xfx[o] = xfx[o] * f; }
  if (j >= L.js && j <= L.je && i >= L.is && i <= L.ie + 1) mfx[o] = mfx[o] * f;
  if (i <= L.ied && j >= L.js && j <= L.je + 1) { cy[o] = cy[o] * f; yfx[o] = yfx[o] * f; }
  if (i >= L.is && i <= L.ie && j >= L.js && j <= L.je + 1) mfy[o] = mfy[o] * f;
}

// one sub-cycle `it` (1-based) of all levels: fv_tp_2d with mass-flux weighting + the tracer update (:206-275)
template <int FAM, bool EDGE>
__global__ void __launch_bounds__(tpt::NT, 2) k_tr_step(Lay L, DevGrid G, tpt::TileMap M, const double* __restrict__ q, double* __restrict__ qo,
                                                      double* __restrict__ dp1, const double* __restrict__ cx, const double* __restrict__ cy,
                                                      const double* __restrict__ xfx, const double* __restrict__ yfx,
                                                      const double* __restrict__ mfx, const double* __restrict__ mfy,
                                                      const int* __restrict__ nsplt, int it, int ord_in, int ord_ou, int update_dp1) {
  using namespace tpt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const Tile T = make_tile(L, M);
  const int ns = nsplt[blockIdx.z];
  const int c = T.lane - 3, i = T.i0 + c;
  if (it > ns) {   // this level is done: carry the tracer over to the other buffer
    for (int r = T.wid; r < TY; r += NW) {
      const int j = T.j0 + r;
      if (c < 0 || c >= TX || i > L.ie || j > L.je) continue;
      const long long o = T.ko + T.idx(i, j);
      qo[o] = __ldg(q + o);
    }
    return;
  }
  stage_inputs<EDGE>(L, G, S, T, cx, cy, xfx, yfx);
  stage_q<EDGE>(L, S, T, q);
  tp_compute<FAM, EDGE>(L, G, S, T, nullptr, nullptr, ord_in, ord_ou);
#pragma unroll
  for (int r = T.wid; r < TY; r += NW) {
    const int j = T.j0 + r;
    if (c < 0 || c >= TX || i > L.ie || j > L.je) continue;
    const int oi = T.idx(i, j);
    const long long o = T.ko + oi;
    const double mx0 = __ldg(mfx + o), mx1 = __ldg(mfx + o + 1), my0 = __ldg(mfy + o), my1 = __ldg(mfy + o + T.NI);
    const double ra = __ldg(G.rarea + oi), d1 = dp1[o];
    const double dp2 = d1 + (mx0 - mx1 + my0 - my1) * ra;
    const double qn = (S.q[r + 3][c + 3] * d1 + (FX(S, r, c) * mx0 - FX(S, r, c + 1) * mx1 + FY(S, r, c) * my0 - FY(S, r + 1, c) * my1) * ra) / dp2;
    qo[o] = qn;
    if (update_dp1 && it < ns) dp1[o] = dp2;   // after the LAST tracer of a sub-cycle that is not the last one (:268-274)
  }
}

extern "C" int fv3_tracer_2d(fv3_ctx** ctxs, int nctx, int hord, double* cmax_out) {
  if (!ctxs || nctx < 1) return -1;
  fv3_ctx* c0 = ctxs[0];
  if (!hord_supported(hord, c0->f.lim_fac)) return fv3_fail(c0, -2, "tracer_2d: hord " + std::to_string(hord) + " not supported (supported: -5, 1..13; 1 only with lim_fac = 1)");
  const int npz = c0->L.npz;
  const bool linked = c0->halo != nullptr;
  std::vector<double> cmax(npz, 0.), tmp(npz);
  std::vector<unsigned long long*> d_cmax(nctx, nullptr);
  // All nq tracers of the contexts are advected (fv_tracer2d.F90:206-275: per sub-cycle every tracer is updated with the same
  // dp1 -> dp2, then dp1 = dp2).  cur[a][iq]: where tracer iq of context a lives right now; spare[a]: the buffer the next kernel
  // writes -- it starts as the scratch plane scr[0] and rotates through the tracer buffers (the halo exchange reads
  // fld[FV3_WORK_Q], which is pointed at the tracer being exchanged).
  for (int a = 0; a < nctx; a++) if (ctxs[a]->tracers.empty()) ctxs[a]->tracers.push_back(ctxs[a]->fld[FV3_WORK_Q]);
  const int nq = (int)c0->tracers.size();
  for (int a = 0; a < nctx; a++)
    if ((int)ctxs[a]->tracers.size() != nq) return fv3_fail(ctxs[a], -1, "tracer_2d: the linked contexts disagree on the number of tracers");
  std::vector<std::vector<double*>> cur(nctx);
  std::vector<double*> spare(nctx), scr0(nctx);
  for (int a = 0; a < nctx; a++) { cur[a] = ctxs[a]->tracers; spare[a] = scr0[a] = ctxs[a]->scr[0]; }
  // whatever path leaves this function: the tracer table of every context holds nq distinct tracer-sized buffers or better, scr[0]
  // is the scratch plane again, fld[WORK_Q] aliases the selected tracer, the per-call tables are freed.  finish(copy): when the
  // scratch plane ended up holding a tracer, that tracer moves (copy = true: with its data) into the spare buffer.
  struct Cleanup {
    fv3_ctx** ctxs; int n; std::vector<std::vector<double*>>& cur; std::vector<double*>& spare; std::vector<double*>& scr0;
    std::vector<unsigned long long*>& tab; bool done = false;
    int finish(bool copy) {
      int rc = 0;
      for (int a = 0; a < n; a++) {
        fv3_ctx* c = ctxs[a];
        cudaSetDevice(c->device);
        if (spare[a] != scr0[a]) {
          for (double*& p : cur[a])
            if (p == scr0[a]) {
              if (copy && cudaMemcpyAsync(spare[a], p, (size_t)c->L.plane * c->L.npz * sizeof(double), cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) rc = 1;
              p = spare[a];
            }
          spare[a] = scr0[a];
        }
        c->tracers = cur[a];
        c->fld[FV3_WORK_Q] = c->tracers[c->tracer_sel];
      }
      done = true;
      return rc;
    }
    ~Cleanup() {
      if (!done) finish(false);
      for (int a = 0; a < n; a++)
        if (tab[a]) { cudaSetDevice(ctxs[a]->device); cudaStreamSynchronize(ctxs[a]->stream); cudaFree(tab[a]); }
    }
  } cleanup{ctxs, nctx, cur, spare, scr0, d_cmax};
  // ---- xfx, yfx, cmax
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    cudaSetDevice(c->device);
    // small per-level tables live behind the edge-profile table allocation pattern: allocate on first use
    FV3_CUDA(c, cudaMalloc(&d_cmax[a], sizeof(double) * 2 * npz + sizeof(int) * npz));
    FV3_CUDA(c, cudaMemsetAsync(d_cmax[a], 0, sizeof(double) * npz, c->stream));
    k_tr_prep<<<plane_grid(c->L, npz), dim3(TI, TJ), 0, c->stream>>>(c->L, c->G, c->fld[FV3_CX], c->fld[FV3_CY], c->fld[FV3_XFX],
                                                                       c->fld[FV3_YFX], d_cmax[a]);
    c->launches++;
  }
  for (int a = 0; a < nctx; a++) {   // maximum over the faces of this process
    fv3_ctx* c = ctxs[a];
    cudaSetDevice(c->device);
    FV3_CUDA(c, cudaMemcpyAsync(tmp.data(), d_cmax[a], sizeof(double) * npz, cudaMemcpyDeviceToHost, c->stream));
    FV3_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < npz; k++) cmax[k] = std::max(cmax[k], tmp[k]);
  }
  {   // maximum over the ranks (mp_reduce_max, :161)
    fv3_ctx* c = c0;
    cudaSetDevice(c->device);
    FV3_CUDA(c, cudaMemcpyAsync(d_cmax[0], cmax.data(), sizeof(double) * npz, cudaMemcpyHostToDevice, c->stream));
    int rc = halo_allreduce_max(c, (double*)d_cmax[0], npz);
    if (rc) return rc;
    FV3_CUDA(c, cudaMemcpyAsync(cmax.data(), d_cmax[0], sizeof(double) * npz, cudaMemcpyDeviceToHost, c->stream));
    FV3_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (cmax_out) for (int k = 0; k < npz; k++) cmax_out[k] = cmax[k];
  std::vector<int> nsplt(npz);
  std::vector<double> frac(npz);
  int nmax = 1;
  for (int k = 0; k < npz; k++) {
    nsplt[k] = (int)(1. + cmax[k]);
    frac[k] = nsplt[k] > 1 ? 1. / (double)nsplt[k] : 1.;
    nmax = std::max(nmax, nsplt[k]);
  }
  const int ord_in = (hord == 10) ? 8 : hord;
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    cudaSetDevice(c->device);
    double* d_frac = (double*)d_cmax[a] + npz;
    int* d_ns = (int*)((double*)d_cmax[a] + 2 * npz);
    FV3_CUDA(c, cudaMemcpyAsync(d_frac, frac.data(), sizeof(double) * npz, cudaMemcpyHostToDevice, c->stream));
    FV3_CUDA(c, cudaMemcpyAsync(d_ns, nsplt.data(), sizeof(int) * npz, cudaMemcpyHostToDevice, c->stream));
    k_tr_scale<<<plane_grid(c->L, npz), dim3(TI, TJ), 0, c->stream>>>(c->L, c->fld[FV3_CX], c->fld[FV3_XFX], c->fld[FV3_CY], c->fld[FV3_YFX],
                                                                        c->fld[FV3_MFX], c->fld[FV3_MFY], d_frac);
    c->launches++;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FV3_CUDA(c0, cudaFuncSetAttribute(k_tr_step<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c0, cudaFuncSetAttribute(k_tr_step<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c0, cudaFuncSetAttribute(k_tr_step<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c0, cudaFuncSetAttribute(k_tr_step<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c0, cudaFuncSetAttribute(k_tr_step<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    FV3_CUDA(c0, cudaFuncSetAttribute(k_tr_step<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tpt::Smem)));
    attr_set = true;
  }
  // ---- sub-cycles
  for (int it = 1; it <= nmax; it++) {
    for (int iq = 0; iq < nq; iq++) {
      for (int a = 0; a < nctx; a++) ctxs[a]->fld[FV3_WORK_Q] = cur[a][iq];
      if (linked) { int rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_TRACER); if (rc) return rc; }   // q_pack (:188) / qn2 (:282)
      const int upd = iq == nq - 1 ? 1 : 0;
      for (int a = 0; a < nctx; a++) {
        fv3_ctx* c = ctxs[a];
        cudaSetDevice(c->device);
        const Lay& L = c->L;
        int* d_ns = (int*)((double*)d_cmax[a] + 2 * npz);
        tpt::TileMap Min, Mfr; int n_in, n_fr;
        tpt::tile_maps(L, Min, Mfr, n_in, n_fr);
#define TR_LAUNCH(FAM, EDGE, MAP, N)                                                                                              \
  k_tr_step<FAM, EDGE><<<dim3(N, 1, npz), tpt::NT, sizeof(tpt::Smem), c->stream>>>(                                               \
      L, c->G, MAP, cur[a][iq], spare[a], c->fld[FV3_DP1], c->fld[FV3_CX], c->fld[FV3_CY], c->fld[FV3_XFX], c->fld[FV3_YFX],      \
      c->fld[FV3_MFX], c->fld[FV3_MFY], d_ns, it, ord_in, hord, upd)
        if (hord_is_rare(hord)) { if (n_in) TR_LAUNCH(2, false, Min, n_in); if (n_fr) TR_LAUNCH(2, true, Mfr, n_fr); }
        else if (hord >= 8) { if (n_in) TR_LAUNCH(1, false, Min, n_in); if (n_fr) TR_LAUNCH(1, true, Mfr, n_fr); }
        else { if (n_in) TR_LAUNCH(0, false, Min, n_in); if (n_fr) TR_LAUNCH(0, true, Mfr, n_fr); }
#undef TR_LAUNCH
        c->launches += (n_in ? 1 : 0) + (n_fr ? 1 : 0);
        std::swap(cur[a][iq], spare[a]);   // the tracer now lives in what was the spare buffer
      }
    }
  }
  if (cleanup.finish(true)) return fv3_fail(c0, 1, "tracer_2d: device copy failed");
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    cudaSetDevice(c->device);
    FV3_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fv3_fail(c, (int)e, std::string("tracer_2d: ") + cudaGetErrorString(e));
  }
  return 0;
}
