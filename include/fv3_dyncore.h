/*
 * fv3_dyncore.h -- C ABI of the B200-native FV3 acoustic-dynamics hot path.
 *
 * This is the drop-in boundary a Fortran ISO_C_BINDING shim inside
 * model/dyn_core.F90 binds to (INTEGRATION.md shows the shim).  Every entry
 * point cites the reference interface it replaces (paths relative to the
 * reference checkout, NOAA-GFDL/GFDL_atmos_cubed_sphere @ 202411).
 *
 * Conventions
 *  - all reals are fp64 ("real" in the 64bit build, fv_arrays.F90:39);
 *  - arrays are Fortran column-major, i fastest, passed as the address of
 *    element (lower_i, lower_j, 1) with the NATIVE Fortran extents written
 *    beside each argument (fv_arrays.F90:1515-1563 state, :1749-1878 metrics);
 *  - logicals are int (0/1); optional arrays are NULL;
 *  - every call returns 0 on success, <0 for a bad argument / unsupported
 *    flag combination (there is NO CPU fallback), >0 for a CUDA/NCCL error;
 *    fv3_last_error() returns the message (the shim maps non-zero to
 *    mpp_error(FATAL, ...), the reference's only error convention,
 *    fv_control.F90:1100-1112);
 *  - one fv3_ctx per GPU / cube face; not thread-safe: call from the master
 *    thread outside any OpenMP region (dyn_core.F90:436,658 are replaced by
 *    ONE batched-over-k call each).
 */
#ifndef FV3_DYNCORE_H
#define FV3_DYNCORE_H

#ifdef __cplusplus
extern "C" {
#endif

/* fv_grid_bounds_type (fv_arrays.F90:1192-1200) + the few gridstruct scalars
 * the hot path reads (fv_arrays.F90:75-205). One rank owns one whole face
 * (layout 1x1), so is=js=1, ie=npx-1, je=npy-1, isd=is-ng ... */
typedef struct fv3_bounds_t {
  int npx, npy, npz, ng;
  int is, ie, js, je;
  int isd, ied, jsd, jed;
  int grid_type;        /* 0..2 gnomonic cubed sphere, 4 doubly-periodic Cartesian */
  int bounded_domain;   /* nested/regional: must be 0 (out of scope) */
  int sw_corner, se_corner, ne_corner, nw_corner;
  int stretched_grid;
  int tile;             /* 1..6 (informational; halo topology uses it) */
} fv3_bounds_t;

/* fv_grid_type metric terms used on the path (fv_arrays.F90:75-205,
 * allocation extents fv_arrays.F90:1749-1878). Immutable after
 * grid_utils_init (fv_grid_utils.F90:84-790); uploaded once by fv3_create. */
typedef struct fv3_grid_t {
  /* (isd:ied, jsd:jed) */
  const double *area, *rarea, *dxa, *dya, *rdxa, *rdya, *cosa_s, *rsin2, *f0;
  /* (isd:ied, jsd:jed, 9) */
  const double *sin_sg, *cos_sg;
  /* (isd:ied+1, jsd:jed) */
  const double *dy, *rdy, *dxc, *rdxc, *cosa_u, *sina_u, *rsin_u, *divg_v, *del6_v;
  /* (isd:ied, jsd:jed+1) */
  const double *dx, *rdx, *dyc, *rdyc, *cosa_v, *sina_v, *rsin_v, *divg_u, *del6_u;
  /* (isd:ied+1, jsd:jed+1) */
  const double *area_c, *rarea_c, *fC, *cosa, *sina;
  /* (is:ie+1, js:je+1)  -- no halo, fv_arrays.F90:1784 */
  const double *rsina;
  /* edge_w, edge_e (npy); edge_s, edge_n (npx)  fv_arrays.F90:1796-1799 */
  const double *edge_w, *edge_e, *edge_s, *edge_n;
  /* grid (isd:ied+1, jsd:jed+1, 2) lon,lat of corners; agrid (isd:ied, jsd:jed, 2) */
  const double *grid, *agrid;
  /* unit vectors of the omega diagnostic (adv_pe, dyn_core.F90:1529-1630), component first as in fv_arrays.F90:1845-1854:
   * ec1, ec2 (3, isd:ied, jsd:jed); en1 (3, is:ie, js:je+1); en2 (3, is:ie+1, js:je).  May be NULL when fv3_dyn_core is never
   * called with FV3_DYN_END_STEP. */
  const double *ec1, *ec2, *en1, *en2;
  double da_min, da_min_c;
} fv3_grid_t;

/* The subset of fv_flags_type (fv_arrays.F90:207-906) + FMS constants_mod
 * values + vertical coordinate that dyn_core and its callees read. */
typedef struct fv3_flags_t {
  int hord_mt, hord_vt, hord_tm, hord_dp, hord_tr;
  int nord, n_sponge, m_split;
  int hydrostatic, do_vort_damp, use_cond, moist_kappa, inline_q, do_f3d;
  int use_logp, convert_ke, prevent_diss_cooling, do_diss_est, is_ideal_case;
  int use_old_omega, fill_dp;
  int sw_test_case;     /* 0: the full (non-SW_DYNAMICS) path.  1: SW_DYNAMICS build with test_case = 1 (BASELINE config 1a):
                         * d_sw advects delp with Courant numbers built from the PRESCRIBED uc, vc and skips the momentum
                         * part (sw_core.F90:626-651, 1025-1026, 1069, 1602-1604); dyn_core skips c_sw, the vertical
                         * solvers and the pressure gradient (dyn_core.F90:394-395, 567-581, 998-1176). */
  double d4_bg, d2_bg, dddmp, d2_bg_k1, d2_bg_k2, vtdm4, d_con, ke_bg, d_ext;
  double a_imp, p_fac, beta, lim_fac, fast_tau_w_sec, rf_cutoff, d2bg_zq, delt_max;
  /* constants_mod (FMS, not in the reference repo): run-time parameters */
  double rdgas, cp_air, grav, kappa, radius, omega, pi;
  double ptop;
  const double *ak, *bk;      /* (npz+1) */
} fv3_flags_t;

/* Prognostic + auxiliary state of fv_atmos_type that crosses the dyn_core
 * boundary (dyn_core.F90:94-98; extents fv_arrays.F90:1521-1563). */
typedef struct fv3_state_t {
  double *u;      /* (isd:ied,   jsd:jed+1, npz) */
  double *v;      /* (isd:ied+1, jsd:jed,   npz) */
  double *w;      /* (isd:ied,   jsd:jed,   npz) */
  double *delz;   /* (is:ie,     js:je,     npz) */
  double *pt;     /* (isd:ied,   jsd:jed,   npz) */
  double *delp;   /* (isd:ied,   jsd:jed,   npz) */
  double *q_con;  /* (isd:ied,   jsd:jed,   npz) or NULL */
  double *cappa;  /* (isd:ied,   jsd:jed,   npz) or NULL */
  double *phis;   /* (isd:ied,   jsd:jed) */
  double *omga;   /* (isd:ied,   jsd:jed,   npz) */
  double *ua, *va;/* (isd:ied,   jsd:jed,   npz) */
  double *uc;     /* (isd:ied+1, jsd:jed,   npz) */
  double *vc;     /* (isd:ied,   jsd:jed+1, npz) */
  double *mfx;    /* (is:ie+1,   js:je,     npz) */
  double *mfy;    /* (is:ie,     js:je+1,   npz) */
  double *cx;     /* (is:ie+1,   jsd:jed,   npz) */
  double *cy;     /* (isd:ied,   js:je+1,   npz) */
  double *pe;     /* (is-1:ie+1, npz+1, js-1:je+1)  k in the middle */
  double *peln;   /* (is:ie,     npz+1, js:je)      k in the middle */
  double *pk;     /* (is:ie,     js:je, npz+1) */
  double *pkz;    /* (is:ie,     js:je, npz) */
  double *ws;     /* (is:ie,     js:je) */
  double *heat_source; /* (isd:ied, jsd:jed, npz) */
  double *diss_est;    /* (isd:ied, jsd:jed, npz) */
} fv3_state_t;

typedef struct fv3_ctx fv3_ctx;

/* ---- lifetime ---------------------------------------------------------- */
/* Replaces the allocate/deallocate of dyn_core's module work arrays
 * (dyn_core.F90:254-286, :1365-1390). Copies the metric terms to the device. */
int  fv3_create(const fv3_bounds_t *bd, const fv3_grid_t *grid,
                const fv3_flags_t *flags, int device, fv3_ctx **out);
void fv3_destroy(fv3_ctx *ctx);
const char *fv3_last_error(const fv3_ctx *ctx);
/* struct sizes, so a binding can check it was built against this header */
int  fv3_abi_sizeof(int which); /* 0 bounds, 1 grid, 2 flags, 3 state */
int  fv3_abi_version(void);

/* ---- named device fields (parity harness + shim plumbing) --------------- */
/* Field ids name the arrays dyn_core touches; fv3_field_dims gives the
 * native Fortran lower bounds and extents of each (what the host buffer
 * must hold). fv3_put/get copy host<->device through pinned staging. */
enum fv3_field_id {
  FV3_U = 0, FV3_V, FV3_W, FV3_DELZ, FV3_PT, FV3_DELP, FV3_QCON, FV3_CAPPA,
  FV3_PHIS, FV3_OMGA, FV3_UA, FV3_VA, FV3_UC, FV3_VC,
  FV3_MFX, FV3_MFY, FV3_CX, FV3_CY,
  FV3_DELPC, FV3_PTC, FV3_UT, FV3_VT, FV3_DIVGD,
  FV3_CRX, FV3_CRY, FV3_XFX, FV3_YFX,
  FV3_GZ, FV3_ZH, FV3_PKC, FV3_PK3, FV3_WS3, FV3_WS,
  FV3_PE, FV3_PELN, FV3_PK, FV3_PKZ, FV3_HEAT, FV3_DISS,
  FV3_WORK_Q,   /* scalar for stand-alone fv_tp_2d (isd:ied, jsd:jed, nk) */
  FV3_WORK_FX,  /* (is:ie+1, js:je, nk) */
  FV3_WORK_FY,  /* (is:ie, js:je+1, nk) */
  FV3_WORK_RAX, /* (is:ie, jsd:jed, nk) */
  FV3_WORK_RAY, /* (isd:ied, js:je, nk) */
  FV3_DP1,      /* delp before dyn_core (isd:ied, jsd:jed, npz): input of tracer_2d (fv_tracer2d.F90:50,63) */
  FV3_DU,       /* (isd:ied, jsd:jed+1, npz) hydrostatic pressure-gradient increment of u kept between substeps when beta > 0 */
  FV3_DV,       /* (isd:ied+1, jsd:jed, npz)   (module arrays du, dv of dyn_core.F90:278-283, 1866-1870, 2098-2099)          */
  FV3_NUM_FIELDS
};
/* dims = {i_lo, ni, j_lo, nj, nk, k_middle(0/1)} */
int fv3_field_dims(const fv3_ctx *ctx, int field, int dims[6]);
int fv3_put_field(fv3_ctx *ctx, int field, const double *host);
int fv3_get_field(fv3_ctx *ctx, int field, double *host);
int fv3_sync(fv3_ctx *ctx);

/* ---- whole-state transfer at dyn_core entry/exit ------------------------ */
int fv3_upload_state(fv3_ctx *ctx, const fv3_state_t *host);
int fv3_download_state(fv3_ctx *ctx, fv3_state_t *host);

/* ---- operator entry points (one call replaces one OpenMP k-loop) -------- */
/* All operate on the DEVICE-RESIDENT fields named above; the *_host variants
 * below wrap them with put/get for a per-call drop-in. */

/* tp_core.F90:85 fv_tp_2d, batched over nk levels, on FV3_WORK_Q with
 * Courant numbers FV3_CRX/CRY, fluxes FV3_XFX/YFX, FV3_WORK_RAX/RAY;
 * writes FV3_WORK_FX/FY. use_mfx: multiply by FV3_MFX/MFY instead of xfx/yfx
 * (tp_core.F90:187-200); nord/damp_c: deln_flux (tp_core.F90:201-206),
 * mass-weighted with FV3_DELP when use_mass. */
int fv3_fv_tp_2d(fv3_ctx *ctx, int nk, int hord, int use_mfx, int use_mass,
                 int nord, double damp_c);

/* sw_core.F90:79 c_sw for k = 1..npz (dyn_core.F90:436-447). */
int fv3_c_sw(fv3_ctx *ctx, double dt2);
/* nh_utils.F90:59 update_dz_c (dyn_core.F90:525). */
int fv3_update_dz_c(fv3_ctx *ctx, double dt2);
/* nh_utils.F90:323 Riem_Solver_c (dyn_core.F90:531-536). */
int fv3_riem_solver_c(fv3_ctx *ctx, double dt2);
/* dyn_core.F90:1635 p_grad_c (dyn_core.F90:562). */
int fv3_p_grad_c(fv3_ctx *ctx, double dt2);
/* sw_core.F90:494 d_sw for k = 1..npz incl. the per-k damping prologue
 * dyn_core.F90:666-772. */
int fv3_d_sw(fv3_ctx *ctx, double dt);
/* nh_utils.F90:204 update_dz_d (dyn_core.F90:911). */
int fv3_update_dz_d(fv3_ctx *ctx, double dt);
/* nh_core.F90:47 Riem_Solver3 (dyn_core.F90:932-940). */
int fv3_riem_solver3(fv3_ctx *ctx, double dt, int last_call);
/* dyn_core.F90:1395 pk3_halo + :1498 pe_halo + gz = zh*grav (:982-989). */
int fv3_pk3_halo(fv3_ctx *ctx);
int fv3_pe_halo(fv3_ctx *ctx);
int fv3_gz_from_zh(fv3_ctx *ctx);
/* dyn_core.F90:1697 nh_p_grad (dyn_core.F90:1032). */
int fv3_nh_p_grad(fv3_ctx *ctx, double dt);
/* Hydrostatic branch.  dyn_core.F90:2202 geopk: cg != 0 is the C-grid call (dyn_core.F90:478-480: delpc, ptc -> pkc, gz,
 * pe, peln), cg == 0 the D-grid call (:905-907: delp, pt -> pkc, gz, pe, peln, pkz).  dyn_core.F90:1909 one_grad_p
 * (call :1019-1021): a2b_ord4 of pkc, gz and the D-grid pressure-gradient update of u, v (+ the d_ext term, see fv3_ext_mode_*). */
/* del2_cubed (dyn_core.F90:2356-2465) on FV3_HEAT or FV3_OMGA; the halo of the field must be current (fv3_halo_exchange with
 * FV3_HALO_HEAT / FV3_HALO_OMGA, or fv3_del2_cubed_cube which does both).  nmax passes (at most 3). */
int fv3_del2_cubed(fv3_ctx *ctx, int field, double cd, int nmax);
int fv3_del2_cubed_cube(fv3_ctx **ctxs, int nctx, int field, double cd, int nmax);
/* Entry conversion of fv_dynamics (fv_dynamics.F90:303-328, 377-398; moist_kappa = F): dp1 = zvir*q_v (q_v in FV3_WORK_Q, result
 * in FV3_DP1), non-hydrostatic pkz = exp(kappa*log(rdg*delp*pt*(1+dp1)/delz)), pt = pt*(1+dp1)[*(1-q_con)]/pkz. */
int fv3_pt_to_theta(fv3_ctx *ctx, double zvir);
/* Vertical remapping, the step between the k_split iterations of fv_dynamics (fv_dynamics.F90:578-625 call site).
 * fv3_lagrangian_to_eulerian: fv_mapz.F90:56-845 on one face after fv3_dyn_core (which leaves pe incl. its one-cell halo, peln,
 * pk, ws and omga current): delp <- ak/bk hybrid levels, pt (theta_v in; theta_v out, or T_v when last_step), w, delz, u, v,
 * [the first use_tracer tracers of the context], pe, peln, pk, pkz remapped; omga interpolated when last_step.  Built: remap_te = F, moist_kappa = F,
 * consv = 0 (no energy fixer), dry air (the last-step T_v -> T conversion is the identity), abs(kord_*) in 8..15 (cs_profile /
 * scalar_profile) or 1..7 (ppm_profile, fv_operators.F90:1382-1723; npz >= 5; with more than 5 tracers kord_tr 1..7 acts as 8, as
 * in mapn_tracer, which calls scalar_profile whatever kord is),
 * kord_wz > 0; everything else returns -2.  kord_tm < 0: T_v is mapped in log p (map_scalar), > 0: theta_v in p.
 * fv3_remap_work_q: the column operators alone on FV3_WORK_Q, from the layers of FV3_PE to the hybrid levels -- mode 0 map_scalar
 * (fv_operators.F90:40), 1 map1_ppm (:137; iv = -2 takes its lower boundary value from FV3_WS), 2 map1_q2 (:352). */
/* External-mode divergence damping, flags.d_ext > 0 (0.02 by default in non-SW builds), hydrostatic branch (dyn_core.F90:745-747,
 * 791-797, 828-847, 1969-1984): fv3_ext_mode_prepare before fv3_d_sw (delp at the cell corners, a2b_ord2), fv3_ext_mode_divg2 after it
 * (mass-weighted vertical mean of the divergence d_sw leaves in FV3_VT); fv3_one_grad_p / fv3_split_p_grad then add its
 * differences to u, v.  fv3_dyn_core does this itself; in the non-hydrostatic beta = 0 path divg2 is read by nothing. */
/* the two halves of the omega diagnostic (FV3_DYN_END_STEP), for stage-by-stage drivers: before / after the last substep */
int fv3_omega_begin(fv3_ctx *ctx);
int fv3_omega_end(fv3_ctx *ctx, double dt);
/* use_old_omega = F (dyn_core.F90:735-742, 774-781, 1196-1214): phase 0 before d_sw (omga = delp), 1 after it (times the convergence
 * of the area fluxes / dt); fv3_omega_end then forms the running sum over k */
int fv3_omega_new(fv3_ctx *ctx, int phase, double dt);
int fv3_ext_mode_prepare(fv3_ctx *ctx);
int fv3_ext_mode_divg2(fv3_ctx *ctx);
int fv3_lagrangian_to_eulerian(fv3_ctx *ctx, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr);
/* the same with water vapour: tracer `sphum` (< use_tracer; -1: none) is the specific humidity, so the last-step conversion T_v -> T
 * divides by 1 + r_vir q_v (fv_mapz.F90:792-822, dtmp = 0; use_cond / condensates are not supported); r_vir = rvgas / rdgas - 1 */
int fv3_lagrangian_to_eulerian_qv(fv3_ctx *ctx, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum,
                                  double r_vir);
int fv3_remap_work_q(fv3_ctx *ctx, int mode, int iv, int kord, double qmin);
/* flagstruct%fill (fv_arrays.F90 `fill`): when on, fv3_lagrangian_to_eulerian / fv3_fv_dynamics apply fillz (fv_fill.F90:34-139, the
 * default branch -- DEV_GFS_PHYS not defined) to each tracer after its remap (fv_mapz.F90:391, fv_operators.F90:337).  Off by
 * default.  fv3_fillz: the operator alone on FV3_WORK_Q with the layer thicknesses FV3_DELP, compute domain. */
int fv3_set_tracer_fill(fv3_ctx *ctx, int on);
int fv3_fillz(fv3_ctx *ctx);
/* dyn_core.F90:1305-1356: filtered heat_source -> pt (levels 1..n_con, limited by delt_max); part of fv3_dyn_core */
int fv3_dcon_heating(fv3_ctx *ctx, double bdt);
int fv3_geopk(fv3_ctx *ctx, int cg);
int fv3_one_grad_p(fv3_ctx *ctx, double dt);
/* beta > 0 (dyn_core.F90:1018-1019, 1027-1028): split_p_grad (non-hydrostatic, :1795-1905) / grad1_p_update (hydrostatic,
 * :2033-2116) with beta_d = beta (0 on the first substep of a call, :404-406); the previous substep's hydrostatic increments
 * live in FV3_DU, FV3_DV.  fv3_dyn_core takes this branch when flags.beta > 0. */
int fv3_split_p_grad(fv3_ctx *ctx, double dt, double beta_d);

/* dyn_core.F90:370-385 (it==1): gz on the compute domain from zs and delz. */
int fv3_gz_init(fv3_ctx *ctx);
/* dyn_core.F90:491-521: zh = gz / gz = zh; and the init_ijk_mem resets (:289-294). */
int fv3_copy_field(fv3_ctx *ctx, int dst_field, int src_field);
int fv3_zero_field(fv3_ctx *ctx, int field);

/* device-side timing of a region over all faces of this process (CUDA events on the
 * library's own streams); ms = max over faces. */
int fv3_timer_start(fv3_ctx **ctxs, int nctx);
int fv3_timer_stop(fv3_ctx **ctxs, int nctx, double *ms);

/* ---- halo exchange (replaces fv_mp_mod.F90:646-874 group updates) ------- */
/* One process may own several faces (tiles) of the cube; peers are other
 * contexts in the same process (device-local copies) or other ranks (NCCL
 * P2P, attached with fv3_comm_attach). Without any peer the halo is frozen. */
enum fv3_halo_group {
  FV3_HALO_UVW = 0,   /* i_pack(8)+(7): u,v D-grid vector + w   dyn_core.F90:430-432 */
  FV3_HALO_GZ,        /* i_pack(5): gz (it==1)                   dyn_core.F90:387,488 */
  FV3_HALO_DIVGD_UCVC,/* i_pack(3)+(9): divgd@corner, uc,vc      dyn_core.F90:451,565,577-578 */
  FV3_HALO_DELP_PT,   /* i_pack(1)(+11): delp, pt[, q_con]       dyn_core.F90:823-825,851 */
  FV3_HALO_ZH_PKC,    /* i_pack(4)/(5): zh, pkc                  dyn_core.F90:945-949,980,992 */
  FV3_HALO_UV_EDGE,   /* mpp_get_boundary(u,v) last substep      dyn_core.F90:1151-1163 */
  FV3_HALO_TRACER,    /* q_pack / mpp_update_domains(qn2)         fv_tracer2d.F90:188,282 (the tracer lives in FV3_WORK_Q) */
  FV3_HALO_HEAT,      /* mpp_update_domains(heat_source) inside del2_cubed   dyn_core.F90:1303, 2401 */
  FV3_HALO_OMGA,      /* mpp_update_domains(omga) inside del2_cubed          fv_dynamics.F90:640, dyn_core.F90:2401 */
  FV3_NUM_HALO_GROUPS
};
/* Link the six (or fewer) contexts of one process into a cube. tiles[i] is
 * the face number (1..6) of ctxs[i]. */
int fv3_cube_link(fv3_ctx **ctxs, const int *tiles, int nctx);
/* Attach a communicator the CALLER owns (ncclComm_t passed as void*; the library never
 * destroys it) to the linked contexts of this process: `rank` = this process's rank in it,
 * tile_rank = the rank owning each face (6 ints, -1 = face absent -> its halo stays frozen). */
int fv3_comm_attach(fv3_ctx **ctxs, int nctx, void *nccl_comm, int rank, const int tile_rank[6]);
/* Exchange one group for all linked contexts of this process. */
int fv3_halo_exchange(fv3_ctx **ctxs, int nctx, int group);
/* Overlapped form (start_group_halo_update / complete_group_halo_update of the reference, fv_mp_mod.F90:646-874): the
 * exchange runs on a side stream; until fv3_halo_wait the caller may only enqueue work that touches neither the halo cells
 * nor the edge cells of the group's fields.  One overlapped exchange may be pending at a time. */
int fv3_halo_start(fv3_ctx **ctxs, int nctx, int group);
int fv3_halo_wait(fv3_ctx **ctxs, int nctx);
/* The library's own communicator: rank 0 calls fv3_nccl_unique_id (128 bytes), the id is
 * broadcast by the host program (torch.distributed / MPI), then every rank calls
 * fv3_comm_init with the face->rank map (6 ints, -1 = face absent). */
int fv3_nccl_unique_id(char *out128);
int fv3_comm_init(fv3_ctx **ctxs, int nctx, const char *id128, int nranks, int rank, const int *tile_rank);
/* fv3_comm_init / fv3_comm_attach also try to set up the PEER-MAPPED exchange (ranks on one NVLink / NVSwitch node): every rank maps
 * the others' receive arenas with CUDA IPC, pack kernels store straight into the receiver's buffer, one-thread kernels publish /
 * await a sequence number; no NCCL kernel on the data path (FV3_HALO_P2P=0 in the environment, or any failure of the IPC setup on
 * any rank, keeps NCCL send / recv).  Returns 1 when it is active for ctx's process; *err (nullable) = 1 after an arrival timeout. */
int fv3_halo_p2p_status(fv3_ctx *ctx, int *err);
/* Mixed precision (BASELINE config 5: "fp32 transport / fp64 Riemann"; the reference's analogue is its 32-bit build, a compile-time
 * choice, not a namelist flag -- hence a setter, not a member of fv3_flags_t).  on = 1: the PPM sweeps of d_sw's interior tiles
 * (fv_tp_2d of delp, w, pt and of the vorticity) compute in fp32 from fp64 fields; the fluxes are applied to the fp64 prognostics
 * in fp64, so the flux-form update stays conservative.  Cube-edge (frame) tiles, the height transport of update_dz_d, the column
 * solvers and every other stage stay fp64.  Default 0 (everything fp64: the configuration every parity number refers to). */
int fv3_set_transport_fp32(fv3_ctx *ctx, int on);
/* Halo index tables (host logic, no device needed): entries of face `tile` for one array
 * of a scalar (ncomp=1) or pair (ncomp=2) field; positions 0 centre, 1 corner, 2 north-
 * staggered (D-grid u, C-grid vc), 3 east-staggered (D-grid v, C-grid uc). Returns the
 * entry count (indices refer to the padded device plane, see fv3_plane_index). */
int fv3_halo_entries(int npx, int ng, int tile, int ncomp, int posx, int posy, int ci,
                     int vector, int halo, int boundary_only, int cap, int *dst,
                     int *src_tile, int *src_comp, int *src, int *sign);
int fv3_halo_table(fv3_ctx *ctx, int group, int spec, int ci, int cap, int *dst,
                   int *src_tile, int *src_comp, int *src, int *sign);
int fv3_plane_index(const fv3_ctx *ctx, int i, int j);

/* ---- the acoustic loop -------------------------------------------------- */
/* dyn_core.F90:313-1286: n_split substeps on device-resident state for the
 * linked contexts of this process (nctx faces), halo exchanges included.
 * bdt is the large (k_split) time step; dt = bdt/n_split (dyn_core.F90:223).
 * flags: 0, or FV3_DYN_GRAPH: the call is captured once into a CUDA graph (one per context list, bdt, n_split, precision mode
 * and ping-pong state of the fields) and replayed with a single cudaGraphLaunch afterwards; bit-identical to flags = 0.  Calls
 * that cannot be captured validly run directly with the same result: the first call of a context (one-time allocations), faces
 * on other ranks (peer-mapped / NCCL exchange), stage timers on.  FV3_DYN_END_STEP: see below.  Any other bit: -2.
 * Hydrostatic branch: on the last substep pk = pkc on the compute domain (dyn_core.F90:1001-1010), so the
 * pk a caller downloads for the remapping is current. */
#define FV3_DYN_GRAPH 1
/* FV3_DYN_END_STEP: this is the last dyn_core call of the k_split loop (the reference's end_step argument): on its last substep the
 * omega diagnostic is formed (dyn_core.F90:409-422, 1182-1195: omga = (pe - pem) / dt + adv_pe(ua, va, pem), use_old_omega = T;
 * use_old_omega = F: the convergence form, :735-742, 774-781, 1196-1214).  Without the bit omga is left untouched.  Needs ec1, ec2, en1, en2 in fv3_grid_t. */
#define FV3_DYN_END_STEP 2
int fv3_dyn_core(fv3_ctx **ctxs, int nctx, double bdt, int n_split, int flags);

/* fv_dynamics.F90:303-398 + :445-662: one call of fv_dynamics on device-resident state -- entry conversion
 * T -> theta_v, then k_split times { dp1 = delp; dyn_core(bdt / k_split, n_split); tracer_2d of FV3_WORK_Q when hord_tr != 0;
 * Lagrangian_to_Eulerian }, then the omega filter del2_cubed(omga, 0.18 da_min, nf_omega); the last dyn_core call runs with
 * FV3_DYN_END_STEP (omega diagnostic).  Dry adiabatic subset (no q_v, no
 * moist_kappa, no inline physics, no energy fixer); pt is temperature on entry and on exit.  flags: as fv3_dyn_core. */
int fv3_fv_dynamics(fv3_ctx **ctxs, int nctx, double bdt, int k_split, int n_split, int kord_mt, int kord_wz, int kord_tm,
                    int kord_tr, int hord_tr, int nf_omega, int flags);
/* the same with water vapour (no condensates): tracer `sphum` is q_v, zvir = rvgas / rdgas - 1: the entry conversion forms
 * theta_v = T (1 + zvir q_v) / pkz and the last remap returns T = T_v / (1 + zvir q_v) */
int fv3_fv_dynamics_qv(fv3_ctx **ctxs, int nctx, double bdt, int k_split, int n_split, int kord_mt, int kord_wz, int kord_tm,
                       int kord_tr, int hord_tr, int nf_omega, int flags, int sphum, double zvir);

/* Tracers: fv3_set_num_tracers(ctx, nq) gives the context nq tracer arrays (isd:ied, jsd:jed, npz); FV3_WORK_Q names the one chosen
 * with fv3_select_tracer (0 by default), so fv3_put_field / fv3_get_field(FV3_WORK_Q) move the selected tracer.  fv3_tracer_2d
 * advects ALL nq tracers (one CFL / sub-cycle count, dp1 advanced once per sub-cycle: fv_tracer2d.F90:206-275), the vertical
 * remap maps the first `use_tracer` of them. */
int fv3_set_num_tracers(fv3_ctx *ctx, int nq);
int fv3_num_tracers(const fv3_ctx *ctx);
int fv3_select_tracer(fv3_ctx *ctx, int iq);

/* fv_tracer2d.F90:49-295 tracer_2d_1L for ONE tracer (nq = 1, trdm = 0, id_divg_mean = 0), all faces of this process in
 * lockstep, after fv3_dyn_core: the tracer in FV3_WORK_Q (halo included) is advected in place with the accumulated mass
 * fluxes / Courant numbers FV3_MFX, MFY, CX, CY of the acoustic loop and dp1 = FV3_DP1; like the reference the call
 * rescales cx, cy, mfx, mfy by 1/nsplt(k) and leaves dp1 at its last intermediate value.  The per-level CFL maximum
 * (mp_reduce_max, :161) is reduced over the faces of the process and, when fv3_comm_init was called, over all ranks with
 * ncclAllReduce(max).  cmax_out (nullable, npz values) receives the reduced maxima. */
int fv3_tracer_2d(fv3_ctx **ctxs, int nctx, int hord, double *cmax_out);
/* kernels launched by this context since creation (bench.py gpu_launches) */
long long fv3_launch_count(const fv3_ctx *ctx);
/* device time (ms) accumulated in a named stage since last reset:
 * names follow fv_timing blocks (dyn_core.F90:435,524,530,657,910,931,1016):
 * "C_SW","UPDATE_DZ_C","Riem_Solver_C","PG_C","D_SW","UPDATE_DZ","Riem_Solver3","PG_D","HALO" */
int fv3_stage_time_ms(fv3_ctx *ctx, const char *stage, double *ms, long long *calls);
int fv3_stage_timers(fv3_ctx *ctx, int enable);

/* ---- per-call host-buffer drop-ins (Fortran call-site shaped) ----------- */
/* sw_core.F90:79 / dyn_core.F90:439-447, batched over k=1..npz; arrays in
 * native extents; w/wc NULL when hydrostatic. */
int fv3_c_sw_host(fv3_ctx *ctx, double *delpc, double *delp, double *ptc, double *pt,
                  double *u, double *v, double *w, double *uc, double *vc,
                  double *ua, double *va, double *wc, double *ut, double *vt,
                  double *divg_d, double dt2);
/* sw_core.F90:494 / dyn_core.F90:762-772, batched over k=1..npz. delpc/ptc
 * are the scratch arrays dyn_core passes (vt and ptc). */
int fv3_d_sw_host(fv3_ctx *ctx, double *delp, double *pt, double *u, double *v,
                  double *w, double *uc, double *vc, double *ua, double *va,
                  double *divg_d, double *mfx, double *mfy, double *cx, double *cy,
                  double *crx, double *cry, double *xfx, double *yfx,
                  double *q_con, double *heat_source, double *diss_est, double dt);
/* tp_core.F90:85, nk levels, host buffers in native extents; mfx/mfy/mass NULL
 * to select the xfx/yfx weighting. q's corner halos are rewritten as in the
 * reference (copy_corners, tp_core.F90:143,164). */
int fv3_fv_tp_2d_host(fv3_ctx *ctx, int nk, double *q, const double *crx,
                      const double *cry, const double *xfx, const double *yfx,
                      const double *ra_x, const double *ra_y, int hord,
                      double *fx, double *fy, const double *mfx,
                      const double *mfy, const double *mass, int nord,
                      double damp_c);
/* nh_utils.F90:323 / dyn_core.F90:531-536 */
int fv3_riem_solver_c_host(fv3_ctx *ctx, double dt2, const double *cappa,
                           const double *phis, const double *w3, const double *ptc,
                           const double *q_con, const double *delpc, double *gz,
                           double *pef, const double *ws3);

#ifdef __cplusplus
}
#endif
#endif /* FV3_DYNCORE_H */
