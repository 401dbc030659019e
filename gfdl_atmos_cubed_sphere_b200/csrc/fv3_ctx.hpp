// Device context of the B200 FV3 acoustic-dynamics library (internal header).
//
// HBM layout: every horizontal field, whatever its staggering, lives in the same
// padded plane  [NJ][NI]  with
//     idx(i,j) = (i - isd + IOFF) + (j - jsd) * NI,   IOFF = 5,  NI % 8 == 0
// so that i = 1 (first compute cell) starts a 64-byte aligned segment of every row
// and one index function serves A-, C-, D- and B-grid arrays (extents isd:ied+1,
// jsd:jed+1).  3-D fields are [nk][NJ][NI] (k slowest) -- coalesced for the (i,j)
// stencil kernels (threadIdx.x -> i) and for the column solvers (one thread per
// column, consecutive threads on consecutive i, marching in k).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/fv3_dyncore.h"

#define FV3_IOFF 5

struct Lay {
  int npx, npy, npz, ng;
  int is, ie, js, je, isd, ied, jsd, jed;
  int NI, NJ;
  long long plane;   // NI*NJ rounded up to a multiple of 16 doubles
  int cube;          // grid_type < 3 && !bounded_domain : cube-edge / corner logic on
  int grid_type;
};

#ifdef __CUDACC__
// in-plane index, 32-bit on purpose (a plane is < 2^31 elements; 64-bit index arithmetic was a large share of the
// integer instructions in every plane kernel); the level offset k*plane is added as long long by the caller
#define LIDX(L, i, j) (((i) - (L).isd + FV3_IOFF) + ((j) - (L).jsd) * (L).NI)
#endif

#ifdef __CUDACC__
// The 32 x 8 thread-block tiles of the padded plane that are NOT entirely inside an interior box, enumerated compactly
// (bottom + top tile rows full width, then the left / right tile columns of the rows in between), so a kernel that
// only has cube-edge work launches a few hundred CTAs per level instead of ~20 000 that exit at once.
struct FrameGrid {
  int nbx, a, b, cl, cr, nby;   // tile rows [0,a) and [b,nby) are frame rows; tile columns [0,cl) and [nbx-cr,nbx) frame columns
  int count() const { return nbx * (a + nby - b) + (b - a) * (cl + cr); }
  __host__ __device__ __forceinline__ void map(int t, int& bx, int& by) const {
    const int nyf = a + nby - b;
    if (t < nbx * nyf) { by = t / nbx; bx = t - by * nbx; if (by >= a) by = by - a + b; }
    else { t -= nbx * nyf; const int nc = cl + cr, row = t / nc, cc = t - row * nc; by = a + row; bx = cc < cl ? cc : nbx - cr + (cc - cl); }
  }
};
// interior box (ilo..ihi, jlo..jhi): points strictly inside need no frame work
static inline FrameGrid frame_grid(const Lay& L, int ilo, int ihi, int jlo, int jhi) {
  FrameGrid f;
  f.nbx = (L.NI + 31) / 32; f.nby = (L.NJ + 7) / 8;
  const int i0 = L.isd - FV3_IOFF, j0 = L.jsd;
  // a tile column bx covers i0 + bx*TI .. +TI-1; it is interior iff it lies within [ilo, ihi]
  int cl = 0; while (cl < f.nbx && i0 + cl * 32 < ilo) cl++;
  int cr = 0; while (cr < f.nbx - cl && i0 + (f.nbx - cr) * 32 - 1 > ihi) cr++;
  int a = 0; while (a < f.nby && j0 + a * 8 < jlo) a++;
  int bt = 0; while (bt < f.nby - a && j0 + (f.nby - bt) * 8 - 1 > jhi) bt++;
  f.cl = cl; f.cr = cr; f.a = a; f.b = f.nby - bt;
  return f;
}
// The POINTS of a frame (outer box minus inner box) enumerated densely: south strip, north strip (full width), then the
// west and east strips of the rows in between.  Edge-formula kernels launch one thread per frame point: with tile-shaped
// launches most threads of every CTA exit at once and the few long-latency edge threads left per CTA ran at ~1 % of the
// machine (a2b_ord4 frame passes: 50 us each for ~3000 points per level).
struct FramePts {
  int io0, io1, jo0, jo1, ii0, ii1, ji0, ji1;
  __host__ __device__ int W() const { return io1 - io0 + 1; }
  __host__ __device__ int nS() const { return W() * (ji0 - jo0); }
  __host__ __device__ int nN() const { return W() * (jo1 - ji1); }
  __host__ __device__ int nW() const { return (ii0 - io0) * (ji1 - ji0 + 1); }
  __host__ __device__ int nE() const { return (io1 - ii1) * (ji1 - ji0 + 1); }
  __host__ __device__ int count() const { return nS() + nN() + nW() + nE(); }
  __host__ __device__ __forceinline__ bool map(int t, int& i, int& j) const {
    const int w = W();
    if (t < nS()) { j = jo0 + t / w; i = io0 + t % w; return true; }
    t -= nS();
    if (t < nN()) { j = ji1 + 1 + t / w; i = io0 + t % w; return true; }
    t -= nN();
    const int ww = ii0 - io0, we = io1 - ii1;
    if (t < nW()) { j = ji0 + t / ww; i = io0 + t % ww; return true; }
    t -= nW();
    if (t < nE()) { j = ji0 + t / we; i = ii1 + 1 + t % we; return true; }
    return false;
  }
};
// outer box [io0,io1]x[jo0,jo1], inner (excluded) box [ii0,ii1]x[ji0,ji1]; an empty inner box makes everything "south"
static inline FramePts frame_pts(int io0, int io1, int jo0, int jo1, int ii0, int ii1, int ji0, int ji1) {
  FramePts f{io0, io1, jo0, jo1, ii0, ii1, ji0, ji1};
  if (ii1 < ii0 || ji1 < ji0) { f.ii0 = io0; f.ii1 = io0 - 1; f.ji0 = jo1 + 1; f.ji1 = jo1; }   // nS = whole box
  return f;
}
#endif

// pointers to the 2-D metric planes on the device
struct DevGrid {
  const double *area, *rarea, *dxa, *dya, *rdxa, *rdya, *cosa_s, *rsin2, *f0;
  const double *sin_sg, *cos_sg;   // 9 planes each
  const double *dy, *rdy, *dxc, *rdxc, *cosa_u, *sina_u, *rsin_u, *divg_v, *del6_v;
  const double *dx, *rdx, *dyc, *rdyc, *cosa_v, *sina_v, *rsin_v, *divg_u, *del6_u;
  const double *area_c, *rarea_c, *fC, *cosa, *sina, *rsina;
  const double *edge_w, *edge_e, *edge_s, *edge_n;  // 1-based: edge_w[j-1]
  const double *ec1, *ec2, *en1, *en2;   // omega diagnostic: three planes each (component slowest), nullptr when not given
  double a2b_w[4][3];   // a2b_ord4 corner extrapolation weights x1/(x2-x1) (a2b_edge.F90:106-130,452-462)
  double da_min, da_min_c;
};

struct FieldDim { int ilo, ni, jlo, nj, nk, kmid; };

struct StageTimer { cudaEvent_t e0, e1; double ms; long long calls; bool pending; };

struct HaloPlan;  // halo.cu

struct fv3_ctx {
  fv3_bounds_t b;
  fv3_flags_t f;
  std::vector<double> ak, bk;
  Lay L;
  DevGrid G;
  int device;
  cudaStream_t stream;
  std::string err;
  // device fields (padded layout); fld[id] may alias (ping-pong targets swapped by stages)
  double* fld[FV3_NUM_FIELDS];
  FieldDim dim[FV3_NUM_FIELDS];
  // ping-pong partners for the in-place updated prognostics (delp, pt, w, u, v, q_con)
  double *alt_delp, *alt_pt, *alt_w, *alt_u, *alt_v, *alt_qcon;
  // scratch 3-D planes (nk = npz+1)
  static const int NSCR = 16;
  double* scr[NSCR];
  // metrics storage
  std::vector<double*> metric_alloc;
  // staging
  double* h_stage; size_t h_stage_bytes;
  double* d_stage; size_t d_stage_bytes;
  // per-k damping parameters (dyn_core.F90:666-733) on host and device
  std::vector<int> nord_v; std::vector<double> damp_vt;
  int* d_kint; double* d_kdbl;     // device copies of per-k coefficient tables
  double* d_dp_ref;                // dp_ref(npz)  dyn_core.F90:242-244
  double* d_edge_tab;              // edge_profile coefficient tables (nh.cu), built on first use
  // tracers (fv3_set_num_tracers): every tracer array incl. the one FV3_WORK_Q was created with; fld[FV3_WORK_Q] always aliases
  // tracers[tracer_sel] (fv3_select_tracer), so every single-tracer entry point works on the selected one
  std::vector<double*> tracers; int tracer_sel = 0;
  int tracer_fill = 0;             // flagstruct%fill: fillz after each remapped tracer (fv3_set_tracer_fill)
  double* d_pem = nullptr;         // interface pressures before the last substep (omega diagnostic), npz + 1 planes, built on first use
  double* d_divg2 = nullptr;       // external-mode damping term (d_ext > 0), one plane, built on first use
  double* d_akbk = nullptr;        // ak(0:km), bk(0:km) for the vertical remap (remap.cu), built on first use
  double* d_rff = nullptr; int k_rf = 0;   // Rayleigh damping table of the vertical solvers (fast_tau_w_sec > 0), built by the first solver call
  long long launches;
  bool capturing = false;   // a CUDA-graph capture of fv3_dyn_core is in progress: stages skip their (unchanged) table uploads
  long long dyn_calls = 0;  // completed fv3_dyn_core calls in which this context took part (the first one does every lazy allocation)
  struct DynGraphs* graphs = nullptr;   // captured fv3_dyn_core graphs (dyn_core.cu), owned by the first context of the call
  int tp_fp32 = 0;   // fv3_set_transport_fp32: the interior-tile PPM sweeps of d_sw compute in fp32 (fp64 storage and updates)
  bool timers_on;
  std::map<std::string, StageTimer> timers;
  HaloPlan* halo;
  int tile;
};

bool fv3_halo_has_remote(const fv3_ctx* c);   // halo.cu
void fv3_free_graphs(fv3_ctx* c);             // dyn_core.cu

// error helpers
int fv3_fail(fv3_ctx* c, int code, const std::string& msg);
#define FV3_CUDA(c, call)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fv3_fail((c), (int)e__, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct StageScope {
  fv3_ctx* c; StageTimer* t;
  StageScope(fv3_ctx* c_, const char* name);
  ~StageScope();
};

// stage implementations (each enqueues kernels on c->stream)
int stage_omega_begin(fv3_ctx* c);
int stage_omega_new(fv3_ctx* c, int phase, double dt);
int stage_omega_end(fv3_ctx* c, double dt);
int stage_ext_mode_prepare(fv3_ctx* c);
int stage_ext_mode_divg2(fv3_ctx* c);
int stage_remap_work_q(fv3_ctx* c, int mode, int iv, int kord, double qmin);
int stage_fillz(fv3_ctx* c);
int stage_lagrangian_to_eulerian(fv3_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum = -1,
                                 double r_vir = 0.);
int stage_fv_tp_2d(fv3_ctx* c, int nk, int hord, int use_mfx, int use_mass, int nord, double damp_c);
int stage_c_sw(fv3_ctx* c, double dt2);
int stage_d_sw(fv3_ctx* c, double dt);
int stage_update_dz_c(fv3_ctx* c, double dt2);
int stage_riem_solver_c(fv3_ctx* c, double dt2);
int stage_p_grad_c(fv3_ctx* c, double dt2);
int stage_update_dz_d(fv3_ctx* c, double dt);
int stage_riem_solver3(fv3_ctx* c, double dt, int last_call);
int stage_pk3_halo(fv3_ctx* c);
int stage_pe_halo(fv3_ctx* c);
int stage_gz_from_zh(fv3_ctx* c);
int stage_nh_p_grad(fv3_ctx* c, double dt, double beta_d = -1.);   // beta_d >= 0: split_p_grad
int stage_geopk(fv3_ctx* c, int cg);
int stage_del2_cubed(fv3_ctx* c, int field, double cd, int nmax);
int stage_dcon_heating(fv3_ctx* c, double bdt);
int stage_pt_to_theta(fv3_ctx* c, double zvir);
int fv3_n_con(const fv3_flags_t& f, int npz);
int stage_one_grad_p(fv3_ctx* c, double dt, double beta_d = -1.);  // beta_d >= 0: grad1_p_update
int stage_gz_init(fv3_ctx* c);
int stage_copy_field(fv3_ctx* c, int dst, int src);
int stage_zero_field(fv3_ctx* c, int f);
void halo_destroy(fv3_ctx* c);
