// The acoustic (n_split) loop of dyn_core on device-resident state.
//
// Reference semantics: model/dyn_core.F90:289-294 (flux-capacitor reset) and :313-1286, the
// non-hydrostatic, non-nested, non-regional global cubed-sphere branch:
//   halo(u,v,w[,delp,pt]) -> c_sw -> [it==1: gz from delz, halo(gz), zh=gz | gz=zh]
//   -> update_dz_c -> Riem_Solver_C -> p_grad_c -> halo(divgd@corner, uc,vc) -> d_sw
//   -> halo(delp,pt[,q_con]) -> update_dz_d -> Riem_Solver3 -> halo(zh,pkc)
//   -> [pe_halo] pk3_halo -> gz = zh*grav -> nh_p_grad -> [last: shared-edge u,v de-dup]
// the SW_DYNAMICS / test_case = 1 branch (BASELINE config 1a, flags.sw_test_case = 1):  d_sw (advection of delp only) -> halo(delp)
// and the hydrostatic branch (BASELINE config 1b):
//   halo(u,v[,delp,pt]) -> c_sw -> geopk(C) -> p_grad_c -> halo(divgd, uc,vc) -> d_sw -> halo(delp,pt) -> geopk(D) -> one_grad_p
// One call replaces the whole it-loop; all faces owned by this process advance in lockstep
// (each on its own stream), exchanges are fv3_halo_exchange (device-local gathers and/or NCCL).
// After the loop (:1300-1356): halo(heat_source) -> del2_cubed -> heating of pt (d_con > 0; csrc/dyn_post.cu).
// FV3_DYN_END_STEP: the omega diagnostic of the last substep (:409-422, :1182-1195, use_old_omega = T).
// Not included (documented in DESIGN.md): use_old_omega = F, Rayleigh friction, fast physics.
#include "fv3_ctx.hpp"
#include <cstdlib>
#include <algorithm>
#include <vector>

extern "C" int fv3_halo_exchange(fv3_ctx** ctxs, int nctx, int group);
extern "C" int fv3_halo_start(fv3_ctx** ctxs, int nctx, int group);
extern "C" int fv3_halo_wait(fv3_ctx** ctxs, int nctx);

#define FORALL(stmt)                                   \
  for (int a_ = 0; a_ < nctx; a_++) {                  \
    fv3_ctx* c = ctxs[a_];                             \
    cudaSetDevice(c->device);                          \
    int rc_ = (stmt);                                  \
    if (rc_) return rc_;                               \
  }

static int dyn_core_direct(fv3_ctx** ctxs, int nctx, double bdt, int n_split, bool end_step);
static int dyn_core_graph(fv3_ctx** ctxs, int nctx, double bdt, int n_split, bool end_step);

extern "C" int fv3_dyn_core(fv3_ctx** ctxs, int nctx, double bdt, int n_split, int flags) {
  if (!ctxs || nctx < 1 || n_split < 1) return -1;
  if (flags & ~(FV3_DYN_GRAPH | FV3_DYN_END_STEP)) return fv3_fail(ctxs[0], -2, "dyn_core: unknown flag bit (defined: FV3_DYN_GRAPH = 1, FV3_DYN_END_STEP = 2)");
  const bool end_step = (flags & FV3_DYN_END_STEP) != 0;
  int rc = (flags & FV3_DYN_GRAPH) ? dyn_core_graph(ctxs, nctx, bdt, n_split, end_step) : dyn_core_direct(ctxs, nctx, bdt, n_split, end_step);
  if (!rc) for (int a = 0; a < nctx; a++) ctxs[a]->dyn_calls++;
  return rc;
}

// ---- FV3_DYN_GRAPH: the whole call (n_split substeps of every face of this process, exchanges included) as ONE CUDA graph.
// Captured once per (context list, bdt, n_split, precision mode, ping-pong state) with stream capture of the very same launch
// sequence the direct path issues -- the face streams are forked from / joined into the first face's stream -- and replayed
// afterwards with a single cudaGraphLaunch; results are bit-identical to the direct path (same kernels, same arguments, same
// order per stream).  What the host must keep in step with a replay: the fld <-> alt ping-pong pointers (a graph is bound to the
// pointer state it was captured in and leaves the state it recorded at the end of the capture) and the launch counter.
// Falls back to the direct path (same results) when a capture would not be valid: first call of a context (it performs the lazy
// allocations / one-time tables), remote faces (the peer-mapped exchange passes sequence numbers as kernel arguments, NCCL calls
// are not captured here), stage timers on, overlapped exchange.
struct DynGraph {
  std::vector<fv3_ctx*> ctxs; double bdt; int n_split; int tp_fp32; bool end_step;
  std::vector<double*> ptr_in, ptr_out;   // fld[] + alt_* of every context at the start / end of the captured call
  long long launches;                     // per context: launches the captured call accounts for (same on every face)
  std::vector<long long> launches_ctx;
  cudaGraphExec_t exec;
};
struct DynGraphs { std::vector<DynGraph> g; };
void fv3_free_graphs(fv3_ctx* c) {
  if (!c->graphs) return;
  for (auto& g : c->graphs->g) cudaGraphExecDestroy(g.exec);
  delete c->graphs; c->graphs = nullptr;
}
static void ptr_state(fv3_ctx** ctxs, int nctx, std::vector<double*>& out) {
  out.clear();
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    for (int i = 0; i < FV3_NUM_FIELDS; i++) out.push_back(c->fld[i]);
    for (double* p : {c->alt_delp, c->alt_pt, c->alt_w, c->alt_u, c->alt_v, c->alt_qcon}) out.push_back(p);
  }
}
static void set_ptr_state(fv3_ctx** ctxs, int nctx, const std::vector<double*>& in) {
  size_t n = 0;
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    for (int i = 0; i < FV3_NUM_FIELDS; i++) c->fld[i] = in[n++];
    c->alt_delp = in[n++]; c->alt_pt = in[n++]; c->alt_w = in[n++]; c->alt_u = in[n++]; c->alt_v = in[n++]; c->alt_qcon = in[n++];
  }
}
static int dyn_core_graph(fv3_ctx** ctxs, int nctx, double bdt, int n_split, bool end_step) {
  fv3_ctx* c0 = ctxs[0];
  bool direct = std::getenv("FV3_HALO_OVERLAP") != nullptr;
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    if (c->dyn_calls == 0 || c->timers_on || c->device != c0->device) direct = true;
    if (fv3_halo_has_remote(c)) direct = true;
  }
  if (end_step) for (int a = 0; a < nctx; a++) if (!ctxs[a]->d_pem) direct = true;   // (its first use allocates)
  if (direct) return dyn_core_direct(ctxs, nctx, bdt, n_split, end_step);
  if (!c0->graphs) c0->graphs = new DynGraphs;
  std::vector<double*> now;
  ptr_state(ctxs, nctx, now);
  DynGraph* G = nullptr;
  for (auto& g : c0->graphs->g)
    if (g.bdt == bdt && g.n_split == n_split && g.tp_fp32 == c0->tp_fp32 && g.end_step == end_step && (int)g.ctxs.size() == nctx &&
        std::equal(g.ctxs.begin(), g.ctxs.end(), ctxs) && g.ptr_in == now) { G = &g; break; }
  cudaSetDevice(c0->device);
  cudaEvent_t ev;
  FV3_CUDA(c0, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (!G) {
    DynGraph g;
    g.ctxs.assign(ctxs, ctxs + nctx); g.bdt = bdt; g.n_split = n_split; g.tp_fp32 = c0->tp_fp32; g.end_step = end_step; g.ptr_in = now;
    std::vector<long long> l0(nctx);
    for (int a = 0; a < nctx; a++) { l0[a] = ctxs[a]->launches; ctxs[a]->capturing = true; }
    // fork: the other faces' streams join the capture of the first face's stream
    cudaError_t e = cudaStreamBeginCapture(c0->stream, cudaStreamCaptureModeRelaxed);
    int rc = 0;
    if (e == cudaSuccess) {
      cudaEventRecord(ev, c0->stream);
      for (int a = 1; a < nctx; a++) if (ctxs[a]->stream != c0->stream) cudaStreamWaitEvent(ctxs[a]->stream, ev, 0);
      rc = dyn_core_direct(ctxs, nctx, bdt, n_split, end_step);
      // join
      for (int a = 1; a < nctx; a++) {
        if (ctxs[a]->stream == c0->stream) continue;
        cudaEvent_t ej; cudaEventCreateWithFlags(&ej, cudaEventDisableTiming);
        cudaEventRecord(ej, ctxs[a]->stream); cudaStreamWaitEvent(c0->stream, ej, 0); cudaEventDestroy(ej);
      }
    }
    cudaGraph_t graph = nullptr;
    if (e == cudaSuccess) e = cudaStreamEndCapture(c0->stream, &graph);
    for (int a = 0; a < nctx; a++) ctxs[a]->capturing = false;
    if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&g.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc) {
      cudaGetLastError();
      set_ptr_state(ctxs, nctx, now);                       // nothing ran: undo the host-side bookkeeping of the failed capture
      for (int a = 0; a < nctx; a++) ctxs[a]->launches = l0[a];
      cudaEventDestroy(ev);
      if (rc) return rc;
      return fv3_fail(c0, (int)e, std::string("dyn_core: CUDA graph capture failed: ") + cudaGetErrorString(e));
    }
    ptr_state(ctxs, nctx, g.ptr_out);
    g.launches_ctx.resize(nctx);
    for (int a = 0; a < nctx; a++) { g.launches_ctx[a] = ctxs[a]->launches - l0[a]; ctxs[a]->launches = l0[a]; }
    set_ptr_state(ctxs, nctx, now);                         // the capture executed nothing; the replay below does
    if (c0->graphs->g.size() >= 8) { cudaGraphExecDestroy(c0->graphs->g.front().exec); c0->graphs->g.erase(c0->graphs->g.begin()); }
    c0->graphs->g.push_back(std::move(g));
    G = &c0->graphs->g.back();
  }
  // replay: after everything already queued on the faces' streams, and before anything queued on them later
  for (int a = 1; a < nctx; a++) {
    if (ctxs[a]->stream == c0->stream) continue;
    cudaEvent_t ej; cudaEventCreateWithFlags(&ej, cudaEventDisableTiming);
    cudaEventRecord(ej, ctxs[a]->stream); cudaStreamWaitEvent(c0->stream, ej, 0); cudaEventDestroy(ej);
  }
  FV3_CUDA(c0, cudaGraphLaunch(G->exec, c0->stream));
  cudaEventRecord(ev, c0->stream);
  for (int a = 1; a < nctx; a++) if (ctxs[a]->stream != c0->stream) cudaStreamWaitEvent(ctxs[a]->stream, ev, 0);
  cudaEventDestroy(ev);
  set_ptr_state(ctxs, nctx, G->ptr_out);
  for (int a = 0; a < nctx; a++) ctxs[a]->launches += G->launches_ctx[a];
  return 0;
}

static int dyn_core_direct(fv3_ctx** ctxs, int nctx, double bdt, int n_split, bool end_step) {
  for (int a = 0; a < nctx; a++) {
    // the loop below branches on the first context's switches: the linked faces must agree on them
    const fv3_flags_t &f0 = ctxs[0]->f, &fa = ctxs[a]->f;
    if (fa.hydrostatic != f0.hydrostatic || fa.sw_test_case != f0.sw_test_case || fa.d_con != f0.d_con || fa.use_cond != f0.use_cond ||
        fa.nord != f0.nord || ctxs[a]->L.npz != ctxs[0]->L.npz || ctxs[a]->L.npx != ctxs[0]->L.npx)
      return fv3_fail(ctxs[a], -1, "dyn_core: the linked contexts disagree on hydrostatic / sw_test_case / d_con / use_cond / nord / npx / npz");
    // beta > 0: split_p_grad / grad1_p_update; beta < -0.1 selects one_grad_p in the non-hydrostatic branch (:1029-1030), not built
    if (fa.beta != f0.beta) return fv3_fail(ctxs[a], -1, "dyn_core: the linked contexts disagree on beta");
    if (fa.beta < 0.0) return fv3_fail(ctxs[a], -2, "dyn_core: beta < 0 (one_grad_p in the non-hydrostatic branch) not supported");
    // d_ext > 0 builds divg2 (dyn_core.F90:745-747, 791-797, 828-845), which only one_grad_p / grad1_p_update read (:1019-1021, :1030):
    // with nh_p_grad (non-hydrostatic, beta = 0) it has no effect on any result and is skipped; the hydrostatic branch builds it
    if (fa.d_ext != f0.d_ext) return fv3_fail(ctxs[a], -1, "dyn_core: the linked contexts disagree on d_ext");
  }
  const double dt = bdt / (double)n_split;   // dyn_core.F90:223
  const double dt2 = 0.5 * dt;
  const bool linked = ctxs[0]->halo != nullptr;
  // dyn_core.F90:289-294
  FORALL(stage_zero_field(c, FV3_MFX)) FORALL(stage_zero_field(c, FV3_MFY)) FORALL(stage_zero_field(c, FV3_CX))
  FORALL(stage_zero_field(c, FV3_CY)) FORALL(stage_zero_field(c, FV3_HEAT))
  int rc;
  const bool hydrostatic = ctxs[0]->f.hydrostatic != 0;
  // beta > 0 (dyn_core.F90:278-283): the hydrostatic pressure-gradient increments du, dv of the previous substep start at zero
  const double beta = ctxs[0]->f.beta;
  if (beta > 0.) { FORALL(stage_zero_field(c, FV3_DU)) FORALL(stage_zero_field(c, FV3_DV)) }
  // Overlapped delp/pt exchange (fv3_halo_start / fv3_halo_wait) is OFF by default: measured at N = 2, C384L79 it LOSES 2.6 %
  // (209.8 vs 204.4 ms per step) -- the NCCL send/recv kernels of the side stream spin on SMs that the concurrent
  // update_dz_d / Riem_Solver3 kernels need.  FV3_HALO_OVERLAP=1 turns it on for experiments.
  static const bool overlap = std::getenv("FV3_HALO_OVERLAP") != nullptr;
  const bool sw_advection = ctxs[0]->f.sw_test_case == 1;
  for (int it = 1; it <= n_split; it++) {
    const bool last_step = (it == n_split);
    if (sw_advection) {
      // SW_DYNAMICS build, test_case = 1 (BASELINE config 1a): every `test_case > 1` block of the loop is skipped
      // (dyn_core.F90:394-395 c_sw, :567-581 uc/vc exchange, :998-1176 pressure gradient); what is left is d_sw
      // (pure advection of delp by the prescribed uc, vc) and the delp halo update (:823-851)
      if (linked && it == 1 && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DELP_PT))) return rc;
      FORALL(stage_d_sw(c, dt))
      if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DELP_PT))) return rc;
      continue;
    }
    const bool omega = last_step && end_step && !sw_advection;   // the omega diagnostic (dyn_core.F90:409-422, 1182-1195)
    const bool omega_new = omega && !ctxs[0]->f.use_old_omega;    // ... in its convergence form (:735-742, 774-781, 1196-1214)
    if (hydrostatic) {   // geopk replaces the vertical solvers, one_grad_p the pressure gradient (dyn_core.F90:478-480, :905-907, :1017-1021)
      if (linked) {
        if (it == 1 && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DELP_PT))) return rc;
        if ((rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_UVW))) return rc;
      }
      if (omega) { FORALL(stage_omega_begin(c)) }
      FORALL(stage_c_sw(c, dt2))
      FORALL(stage_geopk(c, 1))
      FORALL(stage_p_grad_c(c, dt2))
      if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DIVGD_UCVC))) return rc;
      const bool ext_mode = ctxs[0]->f.d_ext > 0.;                                        // external-mode divergence damping
      if (ext_mode) { FORALL(stage_ext_mode_prepare(c)) }                                 // :745-747
      if (omega_new) { FORALL(stage_omega_new(c, 0, dt)) }                                // :735-742
      FORALL(stage_d_sw(c, dt))
      if (omega_new) { FORALL(stage_omega_new(c, 1, dt)) }                                // :774-781
      if (ext_mode) { FORALL(stage_ext_mode_divg2(c)) }                                   // :828-847
      if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DELP_PT))) return rc;
      FORALL(stage_geopk(c, 0))
      if (last_step) { FORALL(stage_copy_field(c, FV3_PK, FV3_PKC)) }   // :1001-1010: remap_step .and. hydrostatic: pk = pkc
      if (beta > 0.) { FORALL(stage_one_grad_p(c, dt, it == 1 ? 0. : beta)) }   // :1018-1019 grad1_p_update, beta_d (:404-406)
      else { FORALL(stage_one_grad_p(c, dt)) }
      if (last_step && linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_UV_EDGE))) return rc;
      if (omega) { FORALL(stage_omega_end(c, dt)) }
      continue;
    }
    if (linked) {
      if (it == 1 && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DELP_PT))) return rc;   // :402 (started in fv_dynamics.F90:467)
      if ((rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_UVW))) return rc;                  // :430-432
    }
    if (omega) { FORALL(stage_omega_begin(c)) }                                           // :409-422
    if (it == 1) {
      FORALL(stage_gz_init(c))                                                            // :370-385
      if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_GZ))) return rc;         // :387, :488
    }
    FORALL(stage_c_sw(c, dt2))                                                            // :436-447
    if (it == 1) { FORALL(stage_copy_field(c, FV3_ZH, FV3_GZ)) }                          // :491-499
    else { FORALL(stage_copy_field(c, FV3_GZ, FV3_ZH)) }                                  // :514-521
    FORALL(stage_update_dz_c(c, dt2))                                                     // :525
    FORALL(stage_riem_solver_c(c, dt2))                                                   // :531
    FORALL(stage_p_grad_c(c, dt2))                                                        // :562
    if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_DIVGD_UCVC))) return rc;   // :451,:565,:577-578
    if (omega_new) { FORALL(stage_omega_new(c, 0, dt)) }                                  // :735-742
    FORALL(stage_d_sw(c, dt))                                                             // :666-812
    if (omega_new) { FORALL(stage_omega_new(c, 1, dt)) }                                  // :774-781
    // delp, pt[, q_con] halos (:823-825 start, :851 complete): update_dz_d and Riem_Solver3 read the compute domain of delp, pt
    // only, so the exchange MAY run on the side stream underneath them (see `overlap` above)
    if (linked && (rc = (overlap ? fv3_halo_start(ctxs, nctx, FV3_HALO_DELP_PT) : fv3_halo_exchange(ctxs, nctx, FV3_HALO_DELP_PT)))) return rc;
    FORALL(stage_update_dz_d(c, dt))                                                      // :911
    FORALL(stage_riem_solver3(c, dt, last_step ? 1 : 0))                                  // :932
    if (linked && (rc = fv3_halo_wait(ctxs, nctx))) return rc;
    if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_ZH_PKC))) return rc;       // :945-949,:980,:992
    if (last_step) { FORALL(stage_pe_halo(c)) }                                           // :952-953
    FORALL(stage_pk3_halo(c))                                                             // :958
    FORALL(stage_gz_from_zh(c))                                                           // :982-989
    if (beta > 0.) { FORALL(stage_nh_p_grad(c, dt, it == 1 ? 0. : beta)) }                // :1027-1028 split_p_grad
    else { FORALL(stage_nh_p_grad(c, dt)) }                                               // :1032
    if (last_step && linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_UV_EDGE))) return rc;   // :1151-1163
    if (omega) { FORALL(stage_omega_end(c, dt)) }                                         // :1182-1195
  }
  // dyn_core.F90:1300-1356: the dissipative heating accumulated over the substeps is filtered and added to pt
  if (!sw_advection && ctxs[0]->f.d_con > 1.e-5 && fv3_n_con(ctxs[0]->f, ctxs[0]->L.npz) != 0) {
    const int nf_ke = std::min(3, ctxs[0]->f.nord + 1);
    if (linked && (rc = fv3_halo_exchange(ctxs, nctx, FV3_HALO_HEAT))) return rc;         // :2401
    FORALL(stage_del2_cubed(c, FV3_HEAT, 0.20 * c->G.da_min, nf_ke))                      // :1303
    FORALL(stage_dcon_heating(c, bdt))                                                    // :1305-1356
  }
  for (int a = 0; a < nctx; a++) {
    cudaSetDevice(ctxs[a]->device);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fv3_fail(ctxs[a], (int)e, std::string("dyn_core: ") + cudaGetErrorString(e));
  }
  return 0;
}

// del2_cubed incl. its halo update on all faces of this process (the omega filter of fv_dynamics.F90:637-642 is
// fv3_del2_cubed_cube(ctxs, n, FV3_OMGA, 0.18 * da_min, nf_omega))
extern "C" int fv3_del2_cubed_cube(fv3_ctx** ctxs, int nctx, int field, double cd, int nmax) {
  if (!ctxs || nctx < 1) return -1;
  if (field != FV3_HEAT && field != FV3_OMGA) return fv3_fail(ctxs[0], -1, "del2_cubed_cube: field must be FV3_HEAT or FV3_OMGA");
  int rc;
  if (ctxs[0]->halo != nullptr && (rc = fv3_halo_exchange(ctxs, nctx, field == FV3_HEAT ? FV3_HALO_HEAT : FV3_HALO_OMGA))) return rc;
  FORALL(stage_del2_cubed(c, field, cd, nmax))
  return 0;
}

// ---- fv_dynamics: the k_split loop on device-resident state ------------------------------------
// Reference semantics: model/fv_dynamics.F90:303-398 (entry: pkz, pt -> theta_v), :445-662 (n_map loop: dp1 = delp -> dyn_core ->
// tracer_2d -> Lagrangian_to_Eulerian -> on the last step the omega filter).  Dry, adiabatic subset: zvir * q_v = 0, no
// moist_kappa / inline physics / energy fixer (consv_te = 0), the context's tracers (fv3_set_num_tracers; advected with hord_tr and remapped with kord_tr
// when hord_tr != 0).  pt is temperature on entry and on exit.
extern "C" int fv3_tracer_2d(fv3_ctx** ctxs, int nctx, int hord, double* cmax_out);
extern "C" int fv3_select_tracer(fv3_ctx* c, int iq);
// sphum >= 0: that tracer is the specific humidity q_v (no condensates): theta_v and T_v carry the factor 1 + zvir q_v
// (fv_dynamics.F90:303-398 on entry, fv_mapz.F90:792-822 on exit); sphum = -1: dry
extern "C" int fv3_fv_dynamics_qv(fv3_ctx** ctxs, int nctx, double bdt, int k_split, int n_split, int kord_mt, int kord_wz, int kord_tm,
                                  int kord_tr, int hord_tr, int nf_omega, int flags, int sphum, double zvir);
extern "C" int fv3_fv_dynamics(fv3_ctx** ctxs, int nctx, double bdt, int k_split, int n_split, int kord_mt, int kord_wz, int kord_tm,
                               int kord_tr, int hord_tr, int nf_omega, int flags) {
  return fv3_fv_dynamics_qv(ctxs, nctx, bdt, k_split, n_split, kord_mt, kord_wz, kord_tm, kord_tr, hord_tr, nf_omega, flags, -1, 0.);
}
extern "C" int fv3_fv_dynamics_qv(fv3_ctx** ctxs, int nctx, double bdt, int k_split, int n_split, int kord_mt, int kord_wz, int kord_tm,
                                  int kord_tr, int hord_tr, int nf_omega, int flags, int sphum, double zvir) {
  if (!ctxs || nctx < 1 || k_split < 1 || n_split < 1) return -1;
  if (sphum >= 0 && hord_tr == 0) return fv3_fail(ctxs[0], -1, "fv_dynamics: a specific-humidity tracer needs tracer transport (hord_tr != 0)");
  if (ctxs[0]->f.sw_test_case) return fv3_fail(ctxs[0], -2, "fv_dynamics: the SW_DYNAMICS build has no k_split loop body beyond dyn_core");
  if (ctxs[0]->L.npz <= 4) return fv3_fail(ctxs[0], -2, "fv_dynamics: npz <= 4 (no vertical remapping, fv_dynamics.F90:567) not supported");
  int rc;
  if (sphum >= 0) {   // the entry conversion reads q_v through FV3_WORK_Q
    for (int a = 0; a < nctx; a++) {
      const int sel = ctxs[a]->tracer_sel;
      if ((rc = fv3_select_tracer(ctxs[a], sphum))) return rc;
      cudaSetDevice(ctxs[a]->device);
      rc = stage_pt_to_theta(ctxs[a], zvir);
      fv3_select_tracer(ctxs[a], sel);
      if (rc) return rc;
    }
  } else {
    FORALL(stage_pt_to_theta(c, 0.))                                                      // fv_dynamics.F90:303-398
  }
  const double mdt = bdt / (double)k_split;                                               // :268
  for (int n_map = 1; n_map <= k_split; n_map++) {
    const int last_step = n_map == k_split;
    FORALL(stage_copy_field(c, FV3_DP1, FV3_DELP))                                        // :473-481 (compute domain + halo)
    if ((rc = fv3_dyn_core(ctxs, nctx, mdt, n_split, (flags & FV3_DYN_GRAPH) | (last_step ? FV3_DYN_END_STEP : 0)))) return rc;   // :495-502
    if (hord_tr != 0 && (rc = fv3_tracer_2d(ctxs, nctx, hord_tr, nullptr))) return rc;    // :512-535
    static const bool remap_serial = std::getenv("FV3_REMAP_SERIAL") != nullptr;          // experiment: one face at a time
    for (int a_ = 0; a_ < nctx; a_++) {                                                   // :578-625
      fv3_ctx* c = ctxs[a_];
      cudaSetDevice(c->device);
      if ((rc = stage_lagrangian_to_eulerian(c, last_step, kord_mt, kord_wz, kord_tm, hord_tr != 0 ? std::max<int>(1, (int)c->tracers.size()) : 0, kord_tr,
                                             sphum, zvir))) return rc;
      if (remap_serial) cudaStreamSynchronize(c->stream);
    }
    if (last_step && nf_omega > 0) {                                                      // :658-662
      const double cd = 0.18 * ctxs[0]->G.da_min;
      if ((rc = fv3_del2_cubed_cube(ctxs, nctx, FV3_OMGA, cd, nf_omega))) return rc;
    }
  }
  for (int a = 0; a < nctx; a++) {
    cudaSetDevice(ctxs[a]->device);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fv3_fail(ctxs[a], (int)e, std::string("fv_dynamics: ") + cudaGetErrorString(e));
  }
  return 0;
}

// ---- whole-state transfer (dyn_core entry/exit) ---------------------------------------------
extern "C" int fv3_upload_state(fv3_ctx* c, const fv3_state_t* s) {
  if (!c || !s) return -1;
  struct { int id; const double* p; } m[] = {{FV3_U, s->u}, {FV3_V, s->v}, {FV3_W, s->w}, {FV3_DELZ, s->delz}, {FV3_PT, s->pt},
                                             {FV3_DELP, s->delp}, {FV3_QCON, s->q_con}, {FV3_CAPPA, s->cappa}, {FV3_PHIS, s->phis},
                                             {FV3_OMGA, s->omga}, {FV3_UA, s->ua}, {FV3_VA, s->va}, {FV3_UC, s->uc}, {FV3_VC, s->vc},
                                             {FV3_PE, s->pe}, {FV3_PELN, s->peln}, {FV3_PK, s->pk}, {FV3_PKZ, s->pkz},
                                             {FV3_DISS, s->diss_est}};
  for (auto& e : m)
    if (e.p) { int rc = fv3_put_field(c, e.id, e.p); if (rc) return rc; }
  return 0;
}
extern "C" int fv3_download_state(fv3_ctx* c, fv3_state_t* s) {
  if (!c || !s) return -1;
  struct { int id; double* p; } m[] = {{FV3_U, s->u}, {FV3_V, s->v}, {FV3_W, s->w}, {FV3_DELZ, s->delz}, {FV3_PT, s->pt},
                                       {FV3_DELP, s->delp}, {FV3_QCON, s->q_con}, {FV3_OMGA, s->omga}, {FV3_UA, s->ua}, {FV3_VA, s->va},
                                       {FV3_UC, s->uc}, {FV3_VC, s->vc}, {FV3_MFX, s->mfx}, {FV3_MFY, s->mfy}, {FV3_CX, s->cx},
                                       {FV3_CY, s->cy}, {FV3_PE, s->pe}, {FV3_PELN, s->peln}, {FV3_PK, s->pk}, {FV3_PKZ, s->pkz},
                                       {FV3_WS, s->ws}, {FV3_HEAT, s->heat_source}, {FV3_DISS, s->diss_est}};
  for (auto& e : m)
    if (e.p) { int rc = fv3_get_field(c, e.id, e.p); if (rc) return rc; }
  return 0;
}

// ---- per-call host-buffer drop-ins ----------------------------------------------------------
#define PUT(id, p) if ((p) && (rc = fv3_put_field(c, id, p))) return rc;
#define GET(id, p) if ((p) && (rc = fv3_get_field(c, id, p))) return rc;
extern "C" int fv3_c_sw_host(fv3_ctx* c, double* delpc, double* delp, double* ptc, double* pt, double* u, double* v, double* w,
                             double* uc, double* vc, double* ua, double* va, double* wc, double* ut, double* vt, double* divg_d,
                             double dt2) {
  if (!c) return -1;
  int rc;
  PUT(FV3_DELP, delp) PUT(FV3_PT, pt) PUT(FV3_U, u) PUT(FV3_V, v) PUT(FV3_W, w)
  if ((rc = fv3_c_sw(c, dt2))) return rc;
  GET(FV3_DELPC, delpc) GET(FV3_PTC, ptc) GET(FV3_UC, uc) GET(FV3_VC, vc) GET(FV3_UA, ua) GET(FV3_VA, va) GET(FV3_OMGA, wc)
  GET(FV3_UT, ut) GET(FV3_VT, vt) GET(FV3_DIVGD, divg_d)
  return 0;
}
extern "C" int fv3_d_sw_host(fv3_ctx* c, double* delp, double* pt, double* u, double* v, double* w, double* uc, double* vc, double* ua,
                             double* va, double* divg_d, double* mfx, double* mfy, double* cx, double* cy, double* crx, double* cry,
                             double* xfx, double* yfx, double* q_con, double* heat_source, double* diss_est, double dt) {
  if (!c) return -1;
  int rc;
  PUT(FV3_DELP, delp) PUT(FV3_PT, pt) PUT(FV3_U, u) PUT(FV3_V, v) PUT(FV3_W, w) PUT(FV3_UC, uc) PUT(FV3_VC, vc) PUT(FV3_UA, ua)
  PUT(FV3_VA, va) PUT(FV3_DIVGD, divg_d) PUT(FV3_MFX, mfx) PUT(FV3_MFY, mfy) PUT(FV3_CX, cx) PUT(FV3_CY, cy) PUT(FV3_QCON, q_con)
  PUT(FV3_HEAT, heat_source) PUT(FV3_DISS, diss_est)
  if ((rc = fv3_d_sw(c, dt))) return rc;
  GET(FV3_DELP, delp) GET(FV3_PT, pt) GET(FV3_U, u) GET(FV3_V, v) GET(FV3_W, w) GET(FV3_MFX, mfx) GET(FV3_MFY, mfy) GET(FV3_CX, cx)
  GET(FV3_CY, cy) GET(FV3_CRX, crx) GET(FV3_CRY, cry) GET(FV3_XFX, xfx) GET(FV3_YFX, yfx) GET(FV3_QCON, q_con)
  GET(FV3_HEAT, heat_source) GET(FV3_DISS, diss_est)
  return 0;
}
extern "C" int fv3_fv_tp_2d_host(fv3_ctx* c, int nk, double* q, const double* crx, const double* cry, const double* xfx,
                                 const double* yfx, const double* ra_x, const double* ra_y, int hord, double* fx, double* fy,
                                 const double* mfx, const double* mfy, const double* mass, int nord, double damp_c) {
  if (!c) return -1;
  if (nk != c->L.npz) return fv3_fail(c, -1, "fv_tp_2d_host: nk must be npz (buffers are full 3-D arrays of npz levels)");
  if (!q || !crx || !cry || !xfx || !yfx || !ra_x || !ra_y || !fx || !fy) return fv3_fail(c, -1, "fv_tp_2d_host: null argument");
  if ((mfx == nullptr) != (mfy == nullptr)) return fv3_fail(c, -1, "fv_tp_2d_host: mfx and mfy must be given together");
  int rc;
  PUT(FV3_WORK_Q, q) PUT(FV3_CRX, crx) PUT(FV3_CRY, cry) PUT(FV3_XFX, xfx) PUT(FV3_YFX, yfx)
  PUT(FV3_WORK_RAX, ra_x) PUT(FV3_WORK_RAY, ra_y) PUT(FV3_MFX, mfx) PUT(FV3_MFY, mfy) PUT(FV3_DELP, mass)
  if ((rc = fv3_fv_tp_2d(c, nk, hord, mfx ? 1 : 0, mass ? 1 : 0, nord, damp_c))) return rc;
  GET(FV3_WORK_FX, fx) GET(FV3_WORK_FY, fy)
  return 0;
}
extern "C" int fv3_riem_solver_c_host(fv3_ctx* c, double dt2, const double* cappa, const double* phis, const double* w3,
                                      const double* ptc, const double* q_con, const double* delpc, double* gz, double* pef,
                                      const double* ws3) {
  if (!c) return -1;
  int rc;
  PUT(FV3_CAPPA, cappa) PUT(FV3_PHIS, phis) PUT(FV3_OMGA, w3) PUT(FV3_PTC, ptc) PUT(FV3_QCON, q_con) PUT(FV3_DELPC, delpc)
  PUT(FV3_GZ, gz) PUT(FV3_WS3, ws3)
  if ((rc = fv3_riem_solver_c(c, dt2))) return rc;
  GET(FV3_GZ, gz) GET(FV3_PKC, pef)
  return 0;
}
