// Context lifetime, field registry, host<->device repacking (C ABI plumbing).
// Replaces the allocate/deallocate of dyn_core's module work arrays
// (reference model/dyn_core.F90:254-286, :1365-1390) and mirrors the extents of
// fv_arrays.F90:1521-1563 (state) / :1749-1878 (metrics) on the host side.
#include "fv3_ctx.hpp"
#include <cstdio>
#include <cstring>
#include <cmath>

int fv3_fail(fv3_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

StageScope::StageScope(fv3_ctx* c_, const char* name) : c(c_), t(nullptr) {
  if (!c->timers_on) return;
  auto it = c->timers.find(name);
  if (it == c->timers.end()) {
    StageTimer st; st.ms = 0; st.calls = 0; st.pending = false;
    cudaEventCreate(&st.e0); cudaEventCreate(&st.e1);
    it = c->timers.emplace(name, st).first;
  }
  t = &it->second;
  if (t->pending) {  // fold the previous interval in before reusing the events
    cudaEventSynchronize(t->e1);
    float ms = 0; cudaEventElapsedTime(&ms, t->e0, t->e1); t->ms += ms; t->pending = false;
  }
  cudaEventRecord(t->e0, c->stream);
}
StageScope::~StageScope() {
  if (!t) return;
  cudaEventRecord(t->e1, c->stream);
  t->pending = true; t->calls++;
}

static void set_dims(fv3_ctx* c) {
  const fv3_bounds_t& b = c->b;
  const int nia = b.ied - b.isd + 1, nja = b.jed - b.jsd + 1, nic = b.ie - b.is + 1, njc = b.je - b.js + 1, kz = b.npz;
  auto A = [&](int nk) { return FieldDim{b.isd, nia, b.jsd, nja, nk, 0}; };
  FieldDim* d = c->dim;
  d[FV3_U] = {b.isd, nia, b.jsd, nja + 1, kz, 0};
  d[FV3_V] = {b.isd, nia + 1, b.jsd, nja, kz, 0};
  d[FV3_W] = A(kz); d[FV3_PT] = A(kz); d[FV3_DELP] = A(kz); d[FV3_QCON] = A(kz); d[FV3_CAPPA] = A(kz);
  d[FV3_DELZ] = {b.is, nic, b.js, njc, kz, 0};
  d[FV3_PHIS] = A(1); d[FV3_OMGA] = A(kz); d[FV3_UA] = A(kz); d[FV3_VA] = A(kz);
  d[FV3_UC] = {b.isd, nia + 1, b.jsd, nja, kz, 0};
  d[FV3_VC] = {b.isd, nia, b.jsd, nja + 1, kz, 0};
  d[FV3_MFX] = {b.is, nic + 1, b.js, njc, kz, 0};
  d[FV3_MFY] = {b.is, nic, b.js, njc + 1, kz, 0};
  d[FV3_CX] = {b.is, nic + 1, b.jsd, nja, kz, 0};
  d[FV3_CY] = {b.isd, nia, b.js, njc + 1, kz, 0};
  d[FV3_DELPC] = A(kz); d[FV3_PTC] = A(kz); d[FV3_UT] = A(kz); d[FV3_VT] = A(kz);
  d[FV3_DIVGD] = {b.isd, nia + 1, b.jsd, nja + 1, kz, 0};
  d[FV3_CRX] = d[FV3_CX]; d[FV3_XFX] = d[FV3_CX]; d[FV3_CRY] = d[FV3_CY]; d[FV3_YFX] = d[FV3_CY];
  d[FV3_GZ] = A(kz + 1); d[FV3_ZH] = A(kz + 1); d[FV3_PKC] = A(kz + 1); d[FV3_PK3] = A(kz + 1);
  d[FV3_WS3] = A(1);
  d[FV3_WS] = {b.is, nic, b.js, njc, 1, 0};
  d[FV3_PE] = {b.is - 1, nic + 2, b.js - 1, njc + 2, kz + 1, 1};
  d[FV3_PELN] = {b.is, nic, b.js, njc, kz + 1, 1};
  d[FV3_PK] = {b.is, nic, b.js, njc, kz + 1, 0};
  d[FV3_PKZ] = {b.is, nic, b.js, njc, kz, 0};
  d[FV3_HEAT] = A(kz); d[FV3_DISS] = A(kz);
  d[FV3_WORK_Q] = A(kz);
  d[FV3_WORK_FX] = {b.is, nic + 1, b.js, njc, kz, 0};
  d[FV3_WORK_FY] = {b.is, nic, b.js, njc + 1, kz, 0};
  d[FV3_WORK_RAX] = {b.is, nic, b.jsd, nja, kz, 0};
  d[FV3_WORK_RAY] = {b.isd, nia, b.js, njc, kz, 0};
  d[FV3_DP1] = A(kz);
  d[FV3_DU] = d[FV3_U]; d[FV3_DV] = d[FV3_V];
}

// native (Fortran) <-> padded device plane repacking.  dir=0: native->device, 1: device->native
__global__ void k_repack(double* __restrict__ dev, double* __restrict__ nat, Lay L, int ilo, int ni, int jlo,
                         int nj, int nk, int kmid, int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= ni || j >= nj || k >= nk) return;
  const long long di = LIDX(L, ilo + i, jlo + j) + (long long)k * L.plane;
  const long long ni_ = ni, nj_ = nj, nk_ = nk;
  const long long hi = kmid ? (i + ni_ * (k + nk_ * j)) : (i + ni_ * (j + nj_ * k));
  if (dir == 0) dev[di] = nat[hi]; else nat[hi] = dev[di];
}

static int upload_metric(fv3_ctx* c, const double* host, int ilo, int ni, int jlo, int nj, int nplanes, const double** out) {
  double* d = nullptr;
  const size_t bytes = (size_t)c->L.plane * nplanes * sizeof(double);
  FV3_CUDA(c, cudaMalloc(&d, bytes));
  FV3_CUDA(c, cudaMemsetAsync(d, 0, bytes, c->stream));
  c->metric_alloc.push_back(d);
  const size_t nb = (size_t)ni * nj * nplanes * sizeof(double);
  if (nb > c->d_stage_bytes) return fv3_fail(c, -3, "staging too small");
  FV3_CUDA(c, cudaMemcpyAsync(c->d_stage, host, nb, cudaMemcpyHostToDevice, c->stream));
  dim3 blk(32, 8), grd((ni + 31) / 32, (nj + 7) / 8, nplanes);
  k_repack<<<grd, blk, 0, c->stream>>>(d, c->d_stage, c->L, ilo, ni, jlo, nj, nplanes, 0, 0);
  FV3_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = d;
  return 0;
}

static int upload_vec(fv3_ctx* c, const double* host, int n, const double** out) {
  double* d = nullptr;
  FV3_CUDA(c, cudaMalloc(&d, sizeof(double) * n));
  FV3_CUDA(c, cudaMemcpy(d, host, sizeof(double) * n, cudaMemcpyHostToDevice));
  c->metric_alloc.push_back(d);
  *out = d;
  return 0;
}

static double gcd_h(const double* q1, const double* q2) {  // fv_grid_utils.F90:1974 (angle)
  double s1 = sin((q1[1] - q2[1]) / 2.), s2 = sin((q1[0] - q2[0]) / 2.);
  return 2. * asin(sqrt(s1 * s1 + cos(q1[1]) * cos(q2[1]) * (s2 * s2)));
}

// a2b_ord4 corner extrapolation weights: precompute x1/(x2-x1) for the 3 pairs at each of
// the 4 face corners (reference recomputes great_circle_dist on every call,
// a2b_edge.F90:106-130,452-462).
static void corner_weights(fv3_ctx* c, const fv3_grid_t* g) {
  const fv3_bounds_t& b = c->b;
  const int nia = b.ied - b.isd + 1, nja = b.jed - b.jsd + 1, npx = b.npx, npy = b.npy;
  auto AG = [&](int i, int j, double* p) {
    p[0] = g->agrid[(i - b.isd) + (size_t)(j - b.jsd) * nia];
    p[1] = g->agrid[(i - b.isd) + (size_t)(j - b.jsd) * nia + (size_t)nia * nja];
  };
  auto GR = [&](int i, int j, double* p) {
    p[0] = g->grid[(i - b.isd) + (size_t)(j - b.jsd) * (nia + 1)];
    p[1] = g->grid[(i - b.isd) + (size_t)(j - b.jsd) * (nia + 1) + (size_t)(nia + 1) * (nja + 1)];
  };
  // pairs (i1,j1,i2,j2) in the order of a2b_edge.F90:108-129
  const int P[4][3][4] = {
      {{1, 1, 2, 2}, {0, 1, -1, 2}, {1, 0, 2, -1}},
      {{npx - 1, 1, npx - 2, 2}, {npx - 1, 0, npx - 2, -1}, {npx, 1, npx + 1, 2}},
      {{npx - 1, npy - 1, npx - 2, npy - 2}, {npx, npy - 1, npx + 1, npy - 2}, {npx - 1, npy, npx - 2, npy + 1}},
      {{1, npy - 1, 2, npy - 2}, {0, npy - 1, -1, npy - 2}, {1, npy, 2, npy + 1}}};
  const int C0[4][2] = {{1, 1}, {npx, 1}, {npx, npy}, {1, npy}};
  for (int q = 0; q < 4; q++) {
    double p0[2]; GR(C0[q][0], C0[q][1], p0);
    for (int m = 0; m < 3; m++) {
      double p1[2], p2[2];
      AG(P[q][m][0], P[q][m][1], p1); AG(P[q][m][2], P[q][m][3], p2);
      double x1 = gcd_h(p1, p0), x2 = gcd_h(p2, p0);
      c->G.a2b_w[q][m] = x1 / (x2 - x1);
    }
  }
}

extern "C" {

int fv3_abi_version(void) { return 1; }
int fv3_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(fv3_bounds_t);
    case 1: return (int)sizeof(fv3_grid_t);
    case 2: return (int)sizeof(fv3_flags_t);
    case 3: return (int)sizeof(fv3_state_t);
  }
  return -1;
}

const char* fv3_last_error(const fv3_ctx* c) { return c ? c->err.c_str() : "null context"; }

int fv3_create(const fv3_bounds_t* bd, const fv3_grid_t* grid, const fv3_flags_t* flags, int device, fv3_ctx** out) {
  if (!bd || !grid || !flags || !out) return -1;
  if (bd->bounded_domain) return -2;                       // nested/regional out of scope
  if (bd->ng != 3) return -2;
  if (bd->is != 1 || bd->js != 1 || bd->ie != bd->npx - 1 || bd->je != bd->npy - 1) return -2;  // one face per context
  if (bd->npx != bd->npy) return -2;
  if (flags->sw_test_case != 0 && flags->sw_test_case != 1) return -2;   // only test_case 1 of the SW_DYNAMICS build is built
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return (int)cudaErrorNoDevice;
  fv3_ctx* c = new fv3_ctx();
  c->b = *bd; c->f = *flags; c->device = device; c->halo = nullptr; c->tile = bd->tile;
  c->launches = 0; c->timers_on = false;
  c->ak.assign(flags->ak, flags->ak + bd->npz + 1);
  c->bk.assign(flags->bk, flags->bk + bd->npz + 1);
  c->f.ak = c->ak.data(); c->f.bk = c->bk.data();
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return (int)cudaErrorInvalidDevice; }
  Lay& L = c->L;
  L.npx = bd->npx; L.npy = bd->npy; L.npz = bd->npz; L.ng = bd->ng;
  L.is = bd->is; L.ie = bd->ie; L.js = bd->js; L.je = bd->je;
  L.isd = bd->isd; L.ied = bd->ied; L.jsd = bd->jsd; L.jed = bd->jed;
  L.NI = ((FV3_IOFF + (bd->ied + 1 - bd->isd + 1)) + 7) / 8 * 8;
  L.NJ = bd->jed + 1 - bd->jsd + 1;
  L.plane = ((long long)L.NI * L.NJ + 15) / 16 * 16;
  L.grid_type = bd->grid_type;
  L.cube = (bd->grid_type < 3 && !bd->bounded_domain) ? 1 : 0;
  set_dims(c);
  cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  const int nkmax = bd->npz + 1;
  c->d_stage_bytes = (size_t)(bd->ied - bd->isd + 2) * (bd->jed - bd->jsd + 2) * (size_t)std::max(nkmax, 9) * sizeof(double);
  c->h_stage = nullptr; c->h_stage_bytes = 0;
  if (cudaMalloc(&c->d_stage, c->d_stage_bytes) != cudaSuccess) { delete c; return (int)cudaErrorMemoryAllocation; }
  int rc = 0;
  // fields
  for (int i = 0; i < FV3_NUM_FIELDS; i++) {
    const size_t bytes = (size_t)L.plane * c->dim[i].nk * sizeof(double);
    if (cudaMalloc(&c->fld[i], bytes) != cudaSuccess) { rc = (int)cudaErrorMemoryAllocation; break; }
    cudaMemsetAsync(c->fld[i], 0, bytes, c->stream);
  }
  double** alts[6] = {&c->alt_delp, &c->alt_pt, &c->alt_w, &c->alt_u, &c->alt_v, &c->alt_qcon};
  for (int i = 0; i < 6 && rc == 0; i++) {
    const size_t bytes = (size_t)L.plane * bd->npz * sizeof(double);
    if (cudaMalloc(alts[i], bytes) != cudaSuccess) { rc = (int)cudaErrorMemoryAllocation; break; }
    cudaMemsetAsync(*alts[i], 0, bytes, c->stream);
  }
  for (int i = 0; i < fv3_ctx::NSCR && rc == 0; i++) {
    const size_t bytes = (size_t)L.plane * nkmax * sizeof(double);
    if (cudaMalloc(&c->scr[i], bytes) != cudaSuccess) { rc = (int)cudaErrorMemoryAllocation; break; }
    cudaMemsetAsync(c->scr[i], 0, bytes, c->stream);
  }
  if (rc) { c->err = "cudaMalloc failed"; *out = c; return rc; }
  // metrics
  DevGrid& G = c->G;
  const int isd = bd->isd, jsd = bd->jsd, nia = bd->ied - isd + 1, nja = bd->jed - jsd + 1;
#define UPA(name) if ((rc = upload_metric(c, grid->name, isd, nia, jsd, nja, 1, &G.name))) { *out = c; return rc; }
#define UPE(name) if ((rc = upload_metric(c, grid->name, isd, nia + 1, jsd, nja, 1, &G.name))) { *out = c; return rc; }
#define UPN(name) if ((rc = upload_metric(c, grid->name, isd, nia, jsd, nja + 1, 1, &G.name))) { *out = c; return rc; }
#define UPC(name) if ((rc = upload_metric(c, grid->name, isd, nia + 1, jsd, nja + 1, 1, &G.name))) { *out = c; return rc; }
  UPA(area) UPA(rarea) UPA(dxa) UPA(dya) UPA(rdxa) UPA(rdya) UPA(cosa_s) UPA(rsin2) UPA(f0)
  if ((rc = upload_metric(c, grid->sin_sg, isd, nia, jsd, nja, 9, &G.sin_sg))) { *out = c; return rc; }
  if ((rc = upload_metric(c, grid->cos_sg, isd, nia, jsd, nja, 9, &G.cos_sg))) { *out = c; return rc; }
  UPE(dy) UPE(rdy) UPE(dxc) UPE(rdxc) UPE(cosa_u) UPE(sina_u) UPE(rsin_u) UPE(divg_v) UPE(del6_v)
  UPN(dx) UPN(rdx) UPN(dyc) UPN(rdyc) UPN(cosa_v) UPN(sina_v) UPN(rsin_v) UPN(divg_u) UPN(del6_u)
  UPC(area_c) UPC(rarea_c) UPC(fC) UPC(cosa) UPC(sina)
  if ((rc = upload_metric(c, grid->rsina, bd->is, bd->ie + 1 - bd->is + 1, bd->js, bd->je + 1 - bd->js + 1, 1, &G.rsina))) { *out = c; return rc; }
  if ((rc = upload_vec(c, grid->edge_w, bd->npy, &G.edge_w))) { *out = c; return rc; }
  if ((rc = upload_vec(c, grid->edge_e, bd->npy, &G.edge_e))) { *out = c; return rc; }
  if ((rc = upload_vec(c, grid->edge_s, bd->npx, &G.edge_s))) { *out = c; return rc; }
  if ((rc = upload_vec(c, grid->edge_n, bd->npx, &G.edge_n))) { *out = c; return rc; }
  {  // unit vectors of the omega diagnostic (nullable): (3, i, j) component first on the host -> three planes on the device
    auto up3 = [&](const double* h, int ilo, int ni, int jlo, int nj, const double** out_p) -> int {
      *out_p = nullptr;
      if (!h) return 0;
      std::vector<double> t((size_t)3 * ni * nj);
      for (int n = 0; n < 3; n++)
        for (size_t e = 0; e < (size_t)ni * nj; e++) t[(size_t)n * ni * nj + e] = h[3 * e + n];
      return upload_metric(c, t.data(), ilo, ni, jlo, nj, 3, out_p);
    };
    const int nic = bd->ie - bd->is + 1, njc = bd->je - bd->js + 1;
    if ((rc = up3(grid->ec1, isd, nia, jsd, nja, &G.ec1)) || (rc = up3(grid->ec2, isd, nia, jsd, nja, &G.ec2)) ||
        (rc = up3(grid->en1, bd->is, nic, bd->js, njc + 1, &G.en1)) || (rc = up3(grid->en2, bd->is, nic + 1, bd->js, njc, &G.en2))) { *out = c; return rc; }
  }
  G.da_min = grid->da_min; G.da_min_c = grid->da_min_c;
  if (L.cube) corner_weights(c, grid);
  // per-k tables
  c->nord_v.assign(bd->npz + 1, 0); c->damp_vt.assign(bd->npz + 1, 0.);
  cudaMalloc(&c->d_kint, sizeof(int) * 12 * (bd->npz + 1));
  cudaMalloc(&c->d_kdbl, sizeof(double) * 12 * (bd->npz + 1));
  std::vector<double> dp(bd->npz);
  for (int k = 0; k < bd->npz; k++) dp[k] = c->ak[k + 1] - c->ak[k] + (c->bk[k + 1] - c->bk[k]) * 1.E5;  // dyn_core.F90:242-244
  c->d_edge_tab = nullptr;
  cudaMalloc(&c->d_dp_ref, sizeof(double) * bd->npz);
  cudaMemcpy(c->d_dp_ref, dp.data(), sizeof(double) * bd->npz, cudaMemcpyHostToDevice);
  cudaStreamSynchronize(c->stream);
  *out = c;
  return 0;
}

void fv3_destroy(fv3_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  halo_destroy(c);
  fv3_free_graphs(c);
  if (!c->tracers.empty()) {   // the tracer table owns every tracer array; fld[FV3_WORK_Q] is an alias of one of them
    for (double* t : c->tracers) cudaFree(t);
    c->fld[FV3_WORK_Q] = nullptr;
  }
  for (int i = 0; i < FV3_NUM_FIELDS; i++) cudaFree(c->fld[i]);
  double* alts[6] = {c->alt_delp, c->alt_pt, c->alt_w, c->alt_u, c->alt_v, c->alt_qcon};
  for (auto p : alts) cudaFree(p);
  for (int i = 0; i < fv3_ctx::NSCR; i++) cudaFree(c->scr[i]);
  for (auto p : c->metric_alloc) cudaFree(p);
  cudaFree(c->d_stage); cudaFree(c->d_kint); cudaFree(c->d_kdbl); cudaFree(c->d_dp_ref); cudaFree(c->d_edge_tab); cudaFree(c->d_rff); cudaFree(c->d_akbk); cudaFree(c->d_divg2); cudaFree(c->d_pem);
  for (auto& kv : c->timers) { cudaEventDestroy(kv.second.e0); cudaEventDestroy(kv.second.e1); }
  cudaStreamDestroy(c->stream);
  delete c;
}

int fv3_field_dims(const fv3_ctx* c, int field, int dims[6]) {
  if (!c || field < 0 || field >= FV3_NUM_FIELDS) return -1;
  const FieldDim& d = c->dim[field];
  dims[0] = d.ilo; dims[1] = d.ni; dims[2] = d.jlo; dims[3] = d.nj; dims[4] = d.nk; dims[5] = d.kmid;
  return 0;
}

int fv3_put_field(fv3_ctx* c, int field, const double* host) {
  if (!c || field < 0 || field >= FV3_NUM_FIELDS || !host) return -1;
  FV3_CUDA(c, cudaSetDevice(c->device));
  const FieldDim& d = c->dim[field];
  const size_t nb = (size_t)d.ni * d.nj * d.nk * sizeof(double);
  if (nb > c->d_stage_bytes) return fv3_fail(c, -3, "staging buffer too small");
  FV3_CUDA(c, cudaMemcpyAsync(c->d_stage, host, nb, cudaMemcpyHostToDevice, c->stream));
  dim3 blk(32, 8), grd((d.ni + 31) / 32, (d.nj + 7) / 8, d.nk);
  k_repack<<<grd, blk, 0, c->stream>>>(c->fld[field], c->d_stage, c->L, d.ilo, d.ni, d.jlo, d.nj, d.nk, d.kmid, 0);
  c->launches++;
  FV3_CUDA(c, cudaGetLastError());
  // the staging buffer is reused by the next put/get: order is guaranteed by the stream,
  // but the HOST buffer must stay valid until the copy has been issued from it
  FV3_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int fv3_get_field(fv3_ctx* c, int field, double* host) {
  if (!c || field < 0 || field >= FV3_NUM_FIELDS || !host) return -1;
  FV3_CUDA(c, cudaSetDevice(c->device));
  const FieldDim& d = c->dim[field];
  const size_t nb = (size_t)d.ni * d.nj * d.nk * sizeof(double);
  if (nb > c->d_stage_bytes) return fv3_fail(c, -3, "staging buffer too small");
  dim3 blk(32, 8), grd((d.ni + 31) / 32, (d.nj + 7) / 8, d.nk);
  k_repack<<<grd, blk, 0, c->stream>>>(c->fld[field], c->d_stage, c->L, d.ilo, d.ni, d.jlo, d.nj, d.nk, d.kmid, 1);
  c->launches++;
  FV3_CUDA(c, cudaGetLastError());
  FV3_CUDA(c, cudaMemcpyAsync(host, c->d_stage, nb, cudaMemcpyDeviceToHost, c->stream));
  FV3_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int fv3_sync(fv3_ctx* c) {
  if (!c) return -1;
  FV3_CUDA(c, cudaSetDevice(c->device));
  FV3_CUDA(c, cudaStreamSynchronize(c->stream));
  FV3_CUDA(c, cudaGetLastError());
  return 0;
}

long long fv3_launch_count(const fv3_ctx* c) { return c ? c->launches : -1; }
int fv3_set_transport_fp32(fv3_ctx* c, int on) {
  if (!c) return -1;
  c->tp_fp32 = on ? 1 : 0;
  return 0;
}

int fv3_stage_timers(fv3_ctx* c, int enable) {
  if (!c) return -1;
  c->timers_on = enable != 0;
  if (enable) for (auto& kv : c->timers) { kv.second.ms = 0; kv.second.calls = 0; kv.second.pending = false; }
  return 0;
}
int fv3_stage_time_ms(fv3_ctx* c, const char* stage, double* ms, long long* calls) {
  if (!c) return -1;
  auto it = c->timers.find(stage);
  if (it == c->timers.end()) { *ms = 0; *calls = 0; return 0; }
  StageTimer& t = it->second;
  if (t.pending) {
    cudaEventSynchronize(t.e1);
    float m = 0; cudaEventElapsedTime(&m, t.e0, t.e1); t.ms += m; t.pending = false;
  }
  *ms = t.ms; *calls = t.calls;
  return 0;
}

// Device-side timing of a region spanning all faces of this process: CUDA events recorded on the
// library's own launch streams (torch.cuda.Event would only see torch's current stream).
static cudaEvent_t g_t0[8], g_t1[8];
static bool g_tinit = false;
int fv3_timer_start(fv3_ctx** ctxs, int nctx) {
  if (!ctxs || nctx < 1 || nctx > 8) return -1;
  if (!g_tinit) { for (int a = 0; a < 8; a++) { cudaEventCreate(&g_t0[a]); cudaEventCreate(&g_t1[a]); } g_tinit = true; }
  for (int a = 0; a < nctx; a++) { cudaSetDevice(ctxs[a]->device); cudaEventRecord(g_t0[a], ctxs[a]->stream); }
  return 0;
}
int fv3_timer_stop(fv3_ctx** ctxs, int nctx, double* ms_out) {
  if (!ctxs || nctx < 1 || nctx > 8 || !g_tinit) return -1;
  for (int a = 0; a < nctx; a++) { cudaSetDevice(ctxs[a]->device); cudaEventRecord(g_t1[a], ctxs[a]->stream); }
  double best = 0.;
  for (int a = 0; a < nctx; a++) {
    cudaEventSynchronize(g_t1[a]);
    for (int b = 0; b < nctx; b++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, g_t0[b], g_t1[a]) == cudaSuccess && ms > best) best = ms;
    }
  }
  *ms_out = best;
  return 0;
}

// ---- stage entry points -------------------------------------------------------------------
#define STAGE_PROLOGUE(c) if (!(c)) return -1; FV3_CUDA(c, cudaSetDevice((c)->device));
#define STAGE_EPILOGUE(c) FV3_CUDA(c, cudaGetLastError()); return 0;

int fv3_fv_tp_2d(fv3_ctx* c, int nk, int hord, int use_mfx, int use_mass, int nord, double damp_c) {
  STAGE_PROLOGUE(c) int rc = stage_fv_tp_2d(c, nk, hord, use_mfx, use_mass, nord, damp_c); if (rc) return rc; STAGE_EPILOGUE(c)
}
int fv3_c_sw(fv3_ctx* c, double dt2) { STAGE_PROLOGUE(c) int rc = stage_c_sw(c, dt2); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_d_sw(fv3_ctx* c, double dt) { STAGE_PROLOGUE(c) int rc = stage_d_sw(c, dt); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_update_dz_c(fv3_ctx* c, double dt2) { STAGE_PROLOGUE(c) int rc = stage_update_dz_c(c, dt2); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_riem_solver_c(fv3_ctx* c, double dt2) { STAGE_PROLOGUE(c) int rc = stage_riem_solver_c(c, dt2); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_p_grad_c(fv3_ctx* c, double dt2) { STAGE_PROLOGUE(c) int rc = stage_p_grad_c(c, dt2); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_update_dz_d(fv3_ctx* c, double dt) { STAGE_PROLOGUE(c) int rc = stage_update_dz_d(c, dt); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_riem_solver3(fv3_ctx* c, double dt, int last_call) { STAGE_PROLOGUE(c) int rc = stage_riem_solver3(c, dt, last_call); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_pk3_halo(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_pk3_halo(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_pe_halo(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_pe_halo(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_gz_from_zh(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_gz_from_zh(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_nh_p_grad(fv3_ctx* c, double dt) { STAGE_PROLOGUE(c) int rc = stage_nh_p_grad(c, dt); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_del2_cubed(fv3_ctx* c, int field, double cd, int nmax) { STAGE_PROLOGUE(c) int rc = stage_del2_cubed(c, field, cd, nmax); if (rc) return rc; STAGE_EPILOGUE(c) }
// ---- tracers: nq arrays of the shape of FV3_WORK_Q; the selected one is what FV3_WORK_Q names
int fv3_set_num_tracers(fv3_ctx* c, int nq) {
  if (!c || nq < 1 || nq > 64) return -1;
  FV3_CUDA(c, cudaSetDevice(c->device));
  if (c->tracers.empty()) c->tracers.push_back(c->fld[FV3_WORK_Q]);
  const size_t bytes = (size_t)c->L.plane * c->dim[FV3_WORK_Q].nk * sizeof(double);
  while ((int)c->tracers.size() < nq) {
    double* d = nullptr;
    FV3_CUDA(c, cudaMalloc(&d, bytes));
    FV3_CUDA(c, cudaMemsetAsync(d, 0, bytes, c->stream));
    c->tracers.push_back(d);
  }
  if (c->tracer_sel >= nq) { c->tracer_sel = 0; c->fld[FV3_WORK_Q] = c->tracers[0]; }
  while ((int)c->tracers.size() > nq) { FV3_CUDA(c, cudaStreamSynchronize(c->stream)); cudaFree(c->tracers.back()); c->tracers.pop_back(); }
  return 0;
}
int fv3_num_tracers(const fv3_ctx* c) { return c ? (c->tracers.empty() ? 1 : (int)c->tracers.size()) : -1; }
int fv3_select_tracer(fv3_ctx* c, int iq) {
  if (!c) return -1;
  if (c->tracers.empty()) c->tracers.push_back(c->fld[FV3_WORK_Q]);
  if (iq < 0 || iq >= (int)c->tracers.size()) return fv3_fail(c, -1, "select_tracer: no such tracer (fv3_set_num_tracers first)");
  c->tracer_sel = iq;
  c->fld[FV3_WORK_Q] = c->tracers[iq];
  return 0;
}
int fv3_set_tracer_fill(fv3_ctx* c, int on) { if (!c) return -1; c->tracer_fill = on != 0; return 0; }
int fv3_fillz(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_fillz(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_omega_new(fv3_ctx* c, int phase, double dt) { STAGE_PROLOGUE(c) int rc = stage_omega_new(c, phase, dt); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_omega_begin(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_omega_begin(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_omega_end(fv3_ctx* c, double dt) { STAGE_PROLOGUE(c) int rc = stage_omega_end(c, dt); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_ext_mode_prepare(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_ext_mode_prepare(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_ext_mode_divg2(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_ext_mode_divg2(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_remap_work_q(fv3_ctx* c, int mode, int iv, int kord, double qmin) { STAGE_PROLOGUE(c) int rc = stage_remap_work_q(c, mode, iv, kord, qmin); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_lagrangian_to_eulerian(fv3_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr) {
  STAGE_PROLOGUE(c) int rc = stage_lagrangian_to_eulerian(c, last_step, kord_mt, kord_wz, kord_tm, use_tracer, kord_tr); if (rc) return rc; STAGE_EPILOGUE(c)
}
int fv3_lagrangian_to_eulerian_qv(fv3_ctx* c, int last_step, int kord_mt, int kord_wz, int kord_tm, int use_tracer, int kord_tr, int sphum, double r_vir) {
  STAGE_PROLOGUE(c) int rc = stage_lagrangian_to_eulerian(c, last_step, kord_mt, kord_wz, kord_tm, use_tracer, kord_tr, sphum, r_vir); if (rc) return rc; STAGE_EPILOGUE(c)
}
int fv3_pt_to_theta(fv3_ctx* c, double zvir) { STAGE_PROLOGUE(c) int rc = stage_pt_to_theta(c, zvir); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_dcon_heating(fv3_ctx* c, double bdt) { STAGE_PROLOGUE(c) int rc = stage_dcon_heating(c, bdt); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_geopk(fv3_ctx* c, int cg) { STAGE_PROLOGUE(c) int rc = stage_geopk(c, cg); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_one_grad_p(fv3_ctx* c, double dt) { STAGE_PROLOGUE(c) int rc = stage_one_grad_p(c, dt); if (rc) return rc; STAGE_EPILOGUE(c) }
// split_p_grad (non-hydrostatic) / grad1_p_update (hydrostatic) with beta_d >= 0 (dyn_core.F90:1018-1028); FV3_DU, FV3_DV carry the
// hydrostatic increment from call to call
int fv3_split_p_grad(fv3_ctx* c, double dt, double beta_d) {
  STAGE_PROLOGUE(c)
  if (beta_d < 0.) return fv3_fail(c, -1, "split_p_grad: beta_d must be >= 0");
  int rc = c->f.hydrostatic ? stage_one_grad_p(c, dt, beta_d) : stage_nh_p_grad(c, dt, beta_d);
  if (rc) return rc;
  STAGE_EPILOGUE(c)
}
int fv3_gz_init(fv3_ctx* c) { STAGE_PROLOGUE(c) int rc = stage_gz_init(c); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_copy_field(fv3_ctx* c, int dst, int src) { STAGE_PROLOGUE(c) int rc = stage_copy_field(c, dst, src); if (rc) return rc; STAGE_EPILOGUE(c) }
int fv3_zero_field(fv3_ctx* c, int f) { STAGE_PROLOGUE(c) int rc = stage_zero_field(c, f); if (rc) return rc; STAGE_EPILOGUE(c) }

}  // extern "C"
