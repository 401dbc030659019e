// Cubed-sphere halo exchange: index tables (host), gather kernels (device-local faces) and
// NCCL point-to-point (faces on other ranks) -- replaces the FMS mpp_domains group updates
// the reference drives through tools/fv_mp_mod.F90:646-874.
//
// Topology: the 12 contacts of tools/fv_mp_mod.F90:498-546.  FMS itself is not part of the
// reference checkout, so the index/sign rules are derived geometrically: each face is the
// square [0,n]^2, the neighbour's coordinates are an affine map p' = M p + c (M a signed
// permutation), staggered points map by geometric edge/corner, vector pairs transform with
// M^T.  tests/test_halo_tables.py checks these tables entry-by-entry against the independent
// NumPy builder (gfdl_atmos_cubed_sphere_b200/cubed_sphere.py), which itself is validated
// against analytic vector fields on the sphere.
//
// Data path: one table per (group, field) and per source face; for a face owned by this
// process the halo is a single gather kernel reading the peer context's field directly
// (same GPU); for a face on another rank the source side packs the requested points into
// one contiguous message per peer and phase (all fields of the phase concatenated, like the
// reference's i_pack groups, dyn_core.F90:823-824), ncclSend/ncclRecv inside one
// ncclGroupStart/End on the context's stream, and the destination unpacks with the sign.
#include "fv3_ctx.hpp"
#include <dlfcn.h>
#include <cstring>
#include <cmath>
#include <algorithm>

enum { POS_CENTER = 0, POS_CORNER = 1, POS_NORTH = 2, POS_EAST = 3 };
enum { EW = 0, EE = 1, ES = 2, EN = 3 };

namespace {
struct Contact { int a, ea, b, eb, rev; };
// tools/fv_mp_mod.F90:499-546
const Contact CONTACTS[12] = {{1, EE, 2, EW, 0}, {1, EN, 3, EW, 1}, {1, EW, 5, EN, 1}, {1, ES, 6, EN, 0}, {2, EN, 3, ES, 0}, {2, EE, 4, ES, 1},
                              {2, ES, 6, EE, 1}, {3, EE, 4, EW, 0}, {3, EN, 5, EW, 1}, {4, EN, 5, ES, 0}, {4, EE, 6, ES, 1}, {5, EE, 6, EW, 0}};
struct Nbr { int tile, edge, rev; };
void neighbour(int tile, int edge, Nbr& out) {
  for (const Contact& c : CONTACTS) {
    if (c.a == tile && c.ea == edge) { out = {c.b, c.eb, c.rev}; return; }
    if (c.b == tile && c.eb == edge) { out = {c.a, c.ea, c.rev}; return; }
  }
  out = {0, 0, 0};
}
void edge_frame(int e, int n, int o[2], int t[2], int nn[2]) {
  o[0] = o[1] = 0; t[0] = t[1] = 0; nn[0] = nn[1] = 0;
  if (e == EW) { t[1] = 1; nn[0] = -1; }
  else if (e == EE) { o[0] = n; t[1] = 1; nn[0] = 1; }
  else if (e == ES) { t[0] = 1; nn[1] = -1; }
  else { o[1] = n; t[0] = 1; nn[1] = 1; }
}
// p_B = M p_A + c  (all quantities are integers in units of half cells: coordinates are doubled)
void affine(int ea, int eb, int rev, int n, int M[2][2], int c2[2]) {
  int oa[2], ta[2], na[2], ob[2], tb[2], nb[2];
  edge_frame(ea, n, oa, ta, na); edge_frame(eb, n, ob, tb, nb);
  const int sg = rev ? -1 : 1;
  for (int r = 0; r < 2; r++)
    for (int s = 0; s < 2; s++) M[r][s] = sg * tb[r] * ta[s] - nb[r] * na[s];
  for (int r = 0; r < 2; r++) {
    const int moa = M[r][0] * oa[0] + M[r][1] * oa[1];
    c2[r] = 2 * (ob[r] + (rev ? n * tb[r] : 0) - moa);   // doubled
  }
}
// doubled offsets of the point of entity (i,j): x2 = 2*i + offx, y2 = 2*j + offy
void pos_off(int pos, int& ox, int& oy) {
  if (pos == POS_CENTER) { ox = -1; oy = -1; }
  else if (pos == POS_CORNER) { ox = -2; oy = -2; }
  else if (pos == POS_NORTH) { ox = -1; oy = -2; }
  else { ox = -2; oy = -1; }
}
void pos_ext(int pos, int& ex, int& ey) { ex = (pos == POS_CORNER || pos == POS_EAST); ey = (pos == POS_CORNER || pos == POS_NORTH); }
}  // namespace

struct HaloEntry { int dst; int src_tile; int src_comp; int src; int sign; };

// entries for one destination array of tile t.  pos[2]: position types of the (x, y) pair
// (ncomp = 1 for scalars).  dst/src are PADDED-plane indices (Lay).
static void build_entries(const Lay& L, int tile, int ncomp, const int pos[2], int ci, int vector, int halo, int boundary_only,
                          std::vector<HaloEntry>& out) {
  const int n = L.npx - 1, ng = L.ng;
  int ox, oy, ex, ey;
  pos_off(pos[ci], ox, oy); pos_ext(pos[ci], ex, ey);
  for (int e = 0; e < 4; e++) {
    Nbr nb; neighbour(tile, e, nb);
    int M[2][2], c2[2];
    affine(e, nb.edge, nb.rev, n, M, c2);
    for (int j = 1 - ng; j <= n + ng + ey; j++)
      for (int i = 1 - ng; i <= n + ng + ex; i++) {
        const int x2 = 2 * i + ox, y2 = 2 * j + oy;   // doubled coordinates
        bool take;
        if (boundary_only) {
          if (ci == 0 && pos[ci] == POS_NORTH && e == EN) take = (y2 == 2 * n) && x2 > 0 && x2 < 2 * n;
          else if (ci == 1 && pos[ci] == POS_EAST && e == EE) take = (x2 == 2 * n) && y2 > 0 && y2 < 2 * n;
          else take = false;
        } else {
          const bool inx = x2 >= 0 && x2 <= 2 * n, iny = y2 >= 0 && y2 <= 2 * n;
          if (e == EW) take = x2 < 0 && x2 >= -2 * halo && iny;
          else if (e == EE) take = x2 > 2 * n && x2 <= 2 * (n + halo) && iny;
          else if (e == ES) take = y2 < 0 && y2 >= -2 * halo && inx;
          else take = y2 > 2 * n && y2 <= 2 * (n + halo) && inx;
        }
        if (!take) continue;
        const int qx = M[0][0] * x2 + M[0][1] * y2 + c2[0], qy = M[1][0] * x2 + M[1][1] * y2 + c2[1];
        int cj = 0, sign = 1, posb = pos[ci];
        if (ncomp == 2) {
          // direction carried by array ci (x-array: i-component) expressed in B's axes: M e_ci
          const int v0 = M[0][ci], v1 = M[1][ci];
          cj = (v0 != 0) ? 0 : 1;
          sign = (cj == 0 ? v0 : v1);
          if (!vector) sign = 1;
          posb = pos[cj];
        }
        int bx, by; pos_off(posb, bx, by);
        const int ib = (qx - bx) / 2, jb = (qy - by) / 2;
        HaloEntry h;
        h.dst = (int)LIDX(L, i, j); h.src_tile = nb.tile; h.src_comp = cj; h.src = (int)LIDX(L, ib, jb); h.sign = sign;
        out.push_back(h);
      }
  }
}

// ---------------------------------------------------------------------------------------------
struct FieldSpec { int fx, fy; int posx, posy; int vector; int nk; int halo; int boundary_only; };
struct DevTable { int n; int *dst, *src, *sign; };   // entries from ONE source tile, ONE source component
struct GroupPlan {
  std::vector<FieldSpec> specs;
  // [spec][ci][src_tile-1][src_comp]
  std::vector<std::vector<std::vector<std::vector<DevTable>>>> tab;
  // the mirror: what THIS tile must send to tile t (entries whose src_tile == me in t's tables)
  // [spec][dst_tile-1][ci][src_comp]
  std::vector<std::vector<std::vector<std::vector<DevTable>>>> send;
};

struct NcclId { char internal[128]; };   // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
  void* lib;
  int (*GetUniqueId)(NcclId*);
  int (*CommInitRank)(void**, int, NcclId, int);
  int (*CommDestroy)(void*);
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
};

// Peer-mapped exchange (one process per GPU on one NVLink / NVSwitch node): every rank owns an ARENA of receive buffers and
// arrival flags in its device memory and maps the arenas of all ranks with CUDA IPC.  A face's pack kernel stores straight into
// the receiver's buffer over NVLink, a one-thread kernel publishes the sequence number in the receiver's flag, the receiver's
// one-thread wait kernel holds its stream until the flag arrives, then the unpack kernel runs.  No NCCL kernel (their channel
// CTAs spin on SMs the compute kernels need, which is why overlapping the NCCL exchange LOST time: DESIGN.md 4) and no host
// round trip are on the data path.  Buffers are per (receiving face, sending face, phase, parity of the sequence number): a
// buffer is rewritten two exchanges of the same phase later, by which time the writer has received the other side's message of
// the exchange in between -- which that side sent after it unpacked this buffer (stream order).  NCCL remains the bootstrap
// (all-gather of the IPC handles and layouts) and the fallback.
constexpr int P2P_MAXR = 8;
struct P2PArena {
  bool on = false;
  int nranks = 0, my_rank = 0;
  char* base = nullptr; size_t bytes = 0;
  char* peer[P2P_MAXR] = {nullptr};
  // layout of every rank's arena (byte offsets, -1: none): buffers [rank][recv face][send face][group][parity], flags [..][group]
  std::vector<long long> boff, foff;
  unsigned long long seq[FV3_NUM_HALO_GROUPS] = {0};
  int* d_err = nullptr;
  long long& B(int r, int f, int t, int g, int par) { return boff[((((size_t)r * 6 + f) * 6 + t) * FV3_NUM_HALO_GROUPS + g) * 2 + par]; }
  long long& F(int r, int f, int t, int g) { return foff[(((size_t)r * 6 + f) * 6 + t) * FV3_NUM_HALO_GROUPS + g]; }
};

struct HaloPlan {
  GroupPlan grp[FV3_NUM_HALO_GROUPS];
  P2PArena* p2p = nullptr;   // owned by the first context of the process
  fv3_ctx* peer[6];        // contexts of the faces owned by this process (nullptr otherwise)
  int tile_rank[6];        // rank owning each face (-1: absent -> halo frozen)
  int my_rank;
  void* comm;              // ncclComm_t
  bool comm_owned;         // created by fv3_comm_init (destroyed with the context) / borrowed through fv3_comm_attach
  double *sendbuf[6], *recvbuf[6];
  size_t bufcap[6];
  std::vector<int*> dev_alloc;
  cudaStream_t xstream;    // side stream of overlapped exchanges (fv3_halo_start / fv3_halo_wait)
  cudaEvent_t xdone;
  bool xpending;
};

// a face of another rank takes part in this context's exchanges (dyn_core.cu: such calls are not graph-captured)
bool fv3_halo_has_remote(const fv3_ctx* c) {
  if (!c->halo) return false;
  for (int t = 0; t < 6; t++) {
    const int r = c->halo->tile_rank[t];
    if (r >= 0 && r != c->halo->my_rank && !c->halo->peer[t]) return true;
  }
  return false;
}

static NcclApi g_nccl = {nullptr};
static int p2p_setup(fv3_ctx** ctxs, int nctx, int nranks, int rank);
static int nccl_load(fv3_ctx* c) {
  if (g_nccl.lib) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fv3_fail(c, -4, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
  g_nccl.lib = h;
  *(void**)(&g_nccl.GetUniqueId) = dlsym(h, "ncclGetUniqueId");
  *(void**)(&g_nccl.CommInitRank) = dlsym(h, "ncclCommInitRank");
  *(void**)(&g_nccl.CommDestroy) = dlsym(h, "ncclCommDestroy");
  *(void**)(&g_nccl.Send) = dlsym(h, "ncclSend");
  *(void**)(&g_nccl.Recv) = dlsym(h, "ncclRecv");
  *(void**)(&g_nccl.GroupStart) = dlsym(h, "ncclGroupStart");
  *(void**)(&g_nccl.GroupEnd) = dlsym(h, "ncclGroupEnd");
  *(void**)(&g_nccl.GetErrorString) = dlsym(h, "ncclGetErrorString");
  *(void**)(&g_nccl.AllReduce) = dlsym(h, "ncclAllReduce");
  *(void**)(&g_nccl.AllGather) = dlsym(h, "ncclAllGather");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd)
    return fv3_fail(c, -4, "libnccl is missing required symbols");
  return 0;
}

static int upload_table(fv3_ctx* c, HaloPlan* hp, const std::vector<HaloEntry>& ent, int src_tile, int src_comp, bool as_send, DevTable& t) {
  std::vector<int> d, s, g;
  for (const HaloEntry& h : ent)
    if (h.src_tile == src_tile && h.src_comp == src_comp) { d.push_back(h.dst); s.push_back(h.src); g.push_back(h.sign); }
  (void)as_send;
  t.n = (int)d.size(); t.dst = t.src = t.sign = nullptr;
  if (!t.n) return 0;
  int* buf = nullptr;
  FV3_CUDA(c, cudaMalloc(&buf, sizeof(int) * 3 * t.n));
  FV3_CUDA(c, cudaMemcpy(buf, d.data(), sizeof(int) * t.n, cudaMemcpyHostToDevice));
  FV3_CUDA(c, cudaMemcpy(buf + t.n, s.data(), sizeof(int) * t.n, cudaMemcpyHostToDevice));
  FV3_CUDA(c, cudaMemcpy(buf + 2 * t.n, g.data(), sizeof(int) * t.n, cudaMemcpyHostToDevice));
  t.dst = buf; t.src = buf + t.n; t.sign = buf + 2 * t.n;
  hp->dev_alloc.push_back(buf);
  return 0;
}

static void group_specs(fv3_ctx* c, int g, std::vector<FieldSpec>& s) {
  const int kz = c->L.npz;
  const bool nh = !c->f.hydrostatic;
  switch (g) {
    case FV3_HALO_UVW:
      s.push_back({FV3_U, FV3_V, POS_NORTH, POS_EAST, 1, kz, 3, 0});
      if (nh) s.push_back({FV3_W, -1, POS_CENTER, 0, 0, kz, 3, 0});
      break;
    case FV3_HALO_GZ: s.push_back({FV3_GZ, -1, POS_CENTER, 0, 0, kz + 1, 3, 0}); break;
    case FV3_HALO_DIVGD_UCVC:
      if (c->f.nord > 0) s.push_back({FV3_DIVGD, -1, POS_CORNER, 0, 0, kz, 3, 0});
      s.push_back({FV3_UC, FV3_VC, POS_EAST, POS_NORTH, 1, kz, 3, 0});
      break;
    case FV3_HALO_DELP_PT:
      s.push_back({FV3_DELP, -1, POS_CENTER, 0, 0, kz, 3, 0});
      s.push_back({FV3_PT, -1, POS_CENTER, 0, 0, kz, 3, 0});
      if (c->f.use_cond) s.push_back({FV3_QCON, -1, POS_CENTER, 0, 0, kz, 3, 0});
      break;
    case FV3_HALO_ZH_PKC:
      s.push_back({FV3_ZH, -1, POS_CENTER, 0, 0, kz + 1, 3, 0});
      s.push_back({FV3_PKC, -1, POS_CENTER, 0, 0, kz + 1, 3, 0});
      break;
    case FV3_HALO_UV_EDGE: s.push_back({FV3_U, FV3_V, POS_NORTH, POS_EAST, 1, kz, 3, 1}); break;
    case FV3_HALO_TRACER: s.push_back({FV3_WORK_Q, -1, POS_CENTER, 0, 0, kz, 3, 0}); break;
    case FV3_HALO_HEAT: s.push_back({FV3_HEAT, -1, POS_CENTER, 0, 0, kz, 3, 0}); break;
    case FV3_HALO_OMGA: s.push_back({FV3_OMGA, -1, POS_CENTER, 0, 0, kz, 3, 0}); break;
  }
}

static int halo_build(fv3_ctx* c) {
  if (c->halo) return 0;
  if (!c->L.cube) return fv3_fail(c, -2, "halo exchange needs the cubed-sphere grid (grid_type < 3)");
  HaloPlan* hp = new HaloPlan();
  for (int t = 0; t < 6; t++) { hp->peer[t] = nullptr; hp->tile_rank[t] = -1; hp->sendbuf[t] = hp->recvbuf[t] = nullptr; }
  hp->comm = nullptr; hp->comm_owned = false; hp->my_rank = 0; for (int t = 0; t < 6; t++) hp->bufcap[t] = 0;
  hp->xstream = nullptr; hp->xdone = nullptr; hp->xpending = false;
  c->halo = hp;
  const int me = c->tile;
  for (int g = 0; g < FV3_NUM_HALO_GROUPS; g++) {
    GroupPlan& gp = hp->grp[g];
    group_specs(c, g, gp.specs);
    gp.tab.resize(gp.specs.size()); gp.send.resize(gp.specs.size());
    for (size_t si = 0; si < gp.specs.size(); si++) {
      const FieldSpec& fs = gp.specs[si];
      const int ncomp = fs.fy >= 0 ? 2 : 1;
      const int pos[2] = {fs.posx, fs.posy};
      gp.tab[si].resize(ncomp);
      for (int ci = 0; ci < ncomp; ci++) {
        std::vector<HaloEntry> ent;
        build_entries(c->L, me, ncomp, pos, ci, fs.vector, fs.halo, fs.boundary_only, ent);
        gp.tab[si][ci].resize(6);
        for (int st = 1; st <= 6; st++) {
          gp.tab[si][ci][st - 1].resize(ncomp);
          for (int sc = 0; sc < ncomp; sc++) {
            int rc = upload_table(c, hp, ent, st, sc, false, gp.tab[si][ci][st - 1][sc]);
            if (rc) return rc;
          }
        }
      }
      // mirror tables: entries of every other tile that read from me
      gp.send[si].resize(6);
      for (int dt = 1; dt <= 6; dt++) {
        gp.send[si][dt - 1].resize(ncomp);
        for (int ci = 0; ci < ncomp; ci++) {
          gp.send[si][dt - 1][ci].resize(ncomp);
          std::vector<HaloEntry> ent;
          if (dt != me) build_entries(c->L, dt, ncomp, pos, ci, fs.vector, fs.halo, fs.boundary_only, ent);
          for (int sc = 0; sc < ncomp; sc++) {
            int rc = upload_table(c, hp, ent, me, sc, true, gp.send[si][dt - 1][ci][sc]);
            if (rc) return rc;
          }
        }
      }
    }
  }
  return 0;
}

void halo_destroy(fv3_ctx* c) {
  if (!c->halo) return;
  HaloPlan* hp = c->halo;
  for (int* p : hp->dev_alloc) cudaFree(p);
  for (int t = 0; t < 6; t++) { cudaFree(hp->sendbuf[t]); cudaFree(hp->recvbuf[t]); }
  if (hp->p2p) {
    P2PArena* a = hp->p2p;
    for (int r = 0; r < a->nranks; r++) if (r != a->my_rank && a->peer[r]) cudaIpcCloseMemHandle(a->peer[r]);
    cudaFree(a->base); cudaFree(a->d_err);
    delete a;
  }
  if (hp->comm && hp->comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(hp->comm);
  if (hp->xstream) cudaStreamDestroy(hp->xstream);
  if (hp->xdone) cudaEventDestroy(hp->xdone);
  delete hp;
  c->halo = nullptr;
}

// dst[k][d[e]] = sign[e] * src[k][s[e]] for a batch of (destination field, source field, table) jobs in ONE launch: the faces of a
// process exchange with ~50 small gathers per face and substep, each a few microseconds of work behind a launch -- batched, an
// exchange of all six faces is 1-2 launches (blockIdx.z = job; jobs shorter than the grid's extent leave their surplus blocks at once).
struct GatherJob { double* dst; const double* src; const int *d, *s, *sg; int n, nk; };
constexpr int GATHER_BATCH = 64;   // 64 x 48 B of kernel parameters
struct GatherJobs { GatherJob j[GATHER_BATCH]; };
__global__ void k_halo_gather_batch(GatherJobs jobs, long long plane) {
  const GatherJob& J = jobs.j[blockIdx.z];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= J.n || (int)blockIdx.y >= J.nk) return;
  const long long ko = (long long)blockIdx.y * plane;
  J.dst[ko + J.d[e]] = (double)J.sg[e] * __ldg(J.src + ko + J.s[e]);
}
// publish / await the sequence number of a peer-mapped message (one thread each).  The wait gives up after ~2e9 polls (a protocol
// error must not hang the GPU): it raises *err and lets the stream continue.
// one thread per message of the exchange (<= P2P_MAXMSG per launch)
constexpr int P2P_MAXMSG = 32;
struct FlagSet { volatile unsigned long long* f[P2P_MAXMSG]; };
__global__ void k_p2p_signal(FlagSet flags, unsigned long long seq) {
  __threadfence_system();
  *flags.f[threadIdx.x] = seq;
  __threadfence_system();
}
__global__ void k_p2p_wait(FlagSet flags, unsigned long long seq, int* err) {
  unsigned long long n = 0;
  while (*flags.f[threadIdx.x] < seq) {
    if (++n > 2000000000ull) { atomicExch(err, 1); break; }
    __nanosleep(64);
  }
  __threadfence_system();
}
// batched pack / unpack: all tables of all messages of an exchange in one launch each (blockIdx.z = job), see k_halo_gather_batch
struct PackJob { double* buf; const double* src; const int* s; int n, nk; };
struct UnpackJob { double* dst; const double* buf; const int *d, *sg; int n, nk; };
struct PackJobs { PackJob j[GATHER_BATCH]; };
struct UnpackJobs { UnpackJob j[GATHER_BATCH]; };
__global__ void k_halo_pack_batch(PackJobs jobs, long long plane) {
  const PackJob& J = jobs.j[blockIdx.z];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= J.n || (int)blockIdx.y >= J.nk) return;
  J.buf[(long long)blockIdx.y * J.n + e] = __ldg(J.src + (long long)blockIdx.y * plane + J.s[e]);
}
__global__ void k_halo_unpack_batch(UnpackJobs jobs, long long plane) {
  const UnpackJob& J = jobs.j[blockIdx.z];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= J.n || (int)blockIdx.y >= J.nk) return;
  J.dst[(long long)blockIdx.y * plane + J.d[e]] = (double)J.sg[e] * J.buf[(long long)blockIdx.y * J.n + e];
}

extern "C" {

int fv3_halo_table(fv3_ctx* c, int group, int spec, int ci, int cap, int* dst, int* src_tile, int* src_comp, int* src, int* sign) {
  // host copy of the entries (for the cross-check against the NumPy builder); returns count
  if (!c) return -1;
  std::vector<FieldSpec> specs; group_specs(c, group, specs);
  if (spec < 0 || spec >= (int)specs.size()) return -1;
  const FieldSpec& fs = specs[spec];
  const int ncomp = fs.fy >= 0 ? 2 : 1;
  const int pos[2] = {fs.posx, fs.posy};
  std::vector<HaloEntry> ent;
  build_entries(c->L, c->tile, ncomp, pos, ci, fs.vector, fs.halo, fs.boundary_only, ent);
  if ((int)ent.size() > cap) return (int)ent.size();
  for (size_t e = 0; e < ent.size(); e++) { dst[e] = ent[e].dst; src_tile[e] = ent[e].src_tile; src_comp[e] = ent[e].src_comp; src[e] = ent[e].src; sign[e] = ent[e].sign; }
  return (int)ent.size();
}
int fv3_plane_index(const fv3_ctx* c, int i, int j) { return (int)LIDX(c->L, i, j); }

// device-free variant (host logic only): entries of tile `tile` for a face of npx-1 cells
int fv3_halo_entries(int npx, int ng, int tile, int ncomp, int posx, int posy, int ci, int vector, int halo, int boundary_only,
                     int cap, int* dst, int* src_tile, int* src_comp, int* src, int* sign) {
  if (npx < 4 || ng < 1 || tile < 1 || tile > 6 || ncomp < 1 || ncomp > 2 || ci < 0 || ci >= ncomp) return -1;
  Lay L;
  L.npx = npx; L.npy = npx; L.npz = 1; L.ng = ng; L.is = 1; L.ie = npx - 1; L.js = 1; L.je = npx - 1;
  L.isd = 1 - ng; L.ied = npx - 1 + ng; L.jsd = 1 - ng; L.jed = npx - 1 + ng;
  L.NI = ((FV3_IOFF + (L.ied + 1 - L.isd + 1)) + 7) / 8 * 8; L.NJ = L.jed + 1 - L.jsd + 1; L.plane = (long long)L.NI * L.NJ; L.cube = 1; L.grid_type = 0;
  const int pos[2] = {posx, posy};
  std::vector<HaloEntry> ent;
  build_entries(L, tile, ncomp, pos, ci, vector, halo, boundary_only, ent);
  if ((int)ent.size() > cap) return (int)ent.size();
  for (size_t e = 0; e < ent.size(); e++) { dst[e] = ent[e].dst; src_tile[e] = ent[e].src_tile; src_comp[e] = ent[e].src_comp; src[e] = ent[e].src; sign[e] = ent[e].sign; }
  return (int)ent.size();
}

int fv3_cube_link(fv3_ctx** ctxs, const int* tiles, int nctx) {
  if (!ctxs || nctx < 1 || nctx > 6) return -1;
  for (int a = 0; a < nctx; a++) {
    fv3_ctx* c = ctxs[a];
    if (tiles[a] < 1 || tiles[a] > 6) return fv3_fail(c, -1, "cube_link: tile out of range");
    c->tile = tiles[a];
    cudaSetDevice(c->device);
    halo_destroy(c);
    int rc = halo_build(c);
    if (rc) return rc;
  }
  for (int a = 0; a < nctx; a++)
    for (int b = 0; b < nctx; b++) {
      ctxs[a]->halo->peer[tiles[b] - 1] = ctxs[b];
      ctxs[a]->halo->tile_rank[tiles[b] - 1] = ctxs[a]->halo->my_rank;
    }
  return 0;
}

int fv3_nccl_unique_id(char* out128) {
  fv3_ctx dummy;
  int rc = nccl_load(nullptr);
  (void)dummy;
  if (rc) return rc;
  return g_nccl.GetUniqueId((NcclId*)out128);
}

// create this rank's communicator and record which rank owns which face
int fv3_comm_init(fv3_ctx** ctxs, int nctx, const char* id128, int nranks, int rank, const int* tile_rank) {
  if (!ctxs || nctx < 1) return -1;
  fv3_ctx* c0 = ctxs[0];
  int rc = nccl_load(c0);
  if (rc) return rc;
  if (!c0->halo) return fv3_fail(c0, -1, "comm_init: call fv3_cube_link first");
  NcclId id;
  memcpy(id.internal, id128, 128);
  cudaSetDevice(c0->device);
  void* comm = nullptr;
  int nrc = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (nrc != 0) return fv3_fail(c0, 1000 + nrc, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"));
  for (int a = 0; a < nctx; a++) {
    HaloPlan* hp = ctxs[a]->halo;
    hp->comm = (a == 0) ? comm : nullptr;   // owned by the first context
    hp->my_rank = rank;
    for (int t = 0; t < 6; t++) hp->tile_rank[t] = tile_rank[t];
  }
  // all contexts share the communicator handle (only ctx 0 destroys it)
  for (int a = 1; a < nctx; a++) ctxs[a]->halo->comm = nullptr;
  ctxs[0]->halo->comm = comm;
  ctxs[0]->halo->comm_owned = true;
  return p2p_setup(ctxs, nctx, nranks, rank);
}
// a communicator owned by the caller (halo_destroy leaves it alone)
int fv3_comm_attach(fv3_ctx** ctxs, int nctx, void* nccl_comm, int rank, const int tile_rank[6]) {
  if (!ctxs || nctx < 1 || !nccl_comm || !tile_rank || rank < 0) return -1;
  fv3_ctx* c0 = ctxs[0];
  int rc = nccl_load(c0);
  if (rc) return rc;
  for (int a = 0; a < nctx; a++)
    if (!ctxs[a]->halo) return fv3_fail(ctxs[a], -1, "comm_attach: call fv3_cube_link first");
  for (int a = 0; a < nctx; a++) {
    HaloPlan* hp = ctxs[a]->halo;
    hp->comm = (a == 0) ? nccl_comm : nullptr;   // the exchange reads the handle of the first context (as fv3_comm_init sets it)
    hp->comm_owned = false;
    hp->my_rank = rank;
    for (int t = 0; t < 6; t++) hp->tile_rank[t] = tile_rank[t];
  }
  int nranks = 0;
  for (int t = 0; t < 6; t++) nranks = std::max(nranks, tile_rank[t] + 1);
  return p2p_setup(ctxs, nctx, nranks, rank);
}

}  // extern "C"

// ---- peer-mapped exchange: arena layout, IPC handle exchange (collective over the communicator) --------------------------------
static size_t msg_elems(const GroupPlan& gp, int t /*remote face, 0-based*/, bool recv) {
  size_t n = 0;
  for (size_t si = 0; si < gp.specs.size(); si++) {
    const int ncomp = gp.specs[si].fy >= 0 ? 2 : 1;
    for (int ci = 0; ci < ncomp; ci++)
      for (int sc = 0; sc < ncomp; sc++)
        n += (size_t)(recv ? gp.tab[si][ci][t][sc].n : gp.send[si][t][ci][sc].n) * gp.specs[si].nk;
  }
  return n;
}
static int p2p_setup(fv3_ctx** ctxs, int nctx, int nranks, int rank) {
  fv3_ctx* c0 = ctxs[0];
  const char* e = getenv("FV3_HALO_P2P");
  if (e && e[0] == '0') return 0;                                   // FV3_HALO_P2P=0: NCCL send / recv on the data path
  if (nranks < 2 || nranks > P2P_MAXR || !g_nccl.AllGather) return 0;
  void* comm = c0->halo->comm;
  constexpr int G = FV3_NUM_HALO_GROUPS;
  P2PArena* A = new P2PArena();
  A->nranks = nranks; A->my_rank = rank;
  A->boff.assign((size_t)nranks * 6 * 6 * G * 2, -1); A->foff.assign((size_t)nranks * 6 * 6 * G, -1);
  size_t off = 0;
  for (int pass = 0; pass < 2; pass++) {                            // flags first, then the buffers (256-byte aligned)
    for (int a = 0; a < nctx; a++) {
      HaloPlan* hp = ctxs[a]->halo; const int f = ctxs[a]->tile - 1;
      for (int t = 0; t < 6; t++) {
        const int r = hp->tile_rank[t];
        if (r < 0 || r == rank || hp->peer[t]) continue;
        for (int g = 0; g < G; g++) {
          const size_t nr = msg_elems(hp->grp[g], t, true);
          if (!nr) continue;
          if (pass == 0) { A->F(rank, f, t, g) = (long long)off; off += 128; }
          else for (int par = 0; par < 2; par++) { A->B(rank, f, t, g, par) = (long long)off; off += (nr * sizeof(double) + 255) / 256 * 256; }
        }
      }
    }
    off = (off + 255) / 256 * 256;
  }
  A->bytes = off > 0 ? off : 256;
  cudaSetDevice(c0->device);
  bool ok = cudaMalloc(&A->base, A->bytes) == cudaSuccess && cudaMemset(A->base, 0, A->bytes) == cudaSuccess &&
            cudaMalloc(&A->d_err, sizeof(int)) == cudaSuccess && cudaMemset(A->d_err, 0, sizeof(int)) == cudaSuccess;
  // record = IPC handle + this rank's slice of the layout tables; all-gathered over the communicator
  const size_t nB = (size_t)6 * 6 * G * 2, nF = (size_t)6 * 6 * G, rec = sizeof(cudaIpcMemHandle_t) + (nB + nF) * sizeof(long long);
  std::vector<char> mine(rec, 0), all(rec * nranks, 0);
  cudaIpcMemHandle_t h; memset(&h, 0, sizeof h);
  ok = ok && cudaIpcGetMemHandle(&h, A->base) == cudaSuccess;
  int okflag = ok ? 1 : 0;
  memcpy(mine.data(), &h, sizeof h);
  memcpy(mine.data() + sizeof h, &A->boff[(size_t)rank * nB], nB * sizeof(long long));
  memcpy(mine.data() + sizeof h + nB * sizeof(long long), &A->foff[(size_t)rank * nF], nF * sizeof(long long));
  if (!ok) memset(mine.data(), 0xff, sizeof h);                      // marks "no arena" for the peers
  char *d_mine = nullptr, *d_all = nullptr;
  FV3_CUDA(c0, cudaMalloc(&d_mine, rec)); FV3_CUDA(c0, cudaMalloc(&d_all, rec * nranks));
  FV3_CUDA(c0, cudaMemcpy(d_mine, mine.data(), rec, cudaMemcpyHostToDevice));
  const int nrc = g_nccl.AllGather(d_mine, d_all, rec, /*ncclChar*/ 0, comm, c0->stream);
  if (nrc != 0) { cudaFree(d_mine); cudaFree(d_all); return fv3_fail(c0, 1000 + nrc, "ncclAllGather (peer-mapped halo setup) failed"); }
  FV3_CUDA(c0, cudaStreamSynchronize(c0->stream));
  FV3_CUDA(c0, cudaMemcpy(all.data(), d_all, rec * nranks, cudaMemcpyDeviceToHost));
  cudaFree(d_mine); cudaFree(d_all);
  for (int r = 0; r < nranks && okflag; r++) {
    const char* p = all.data() + rec * r;
    bool none = true;
    for (size_t b = 0; b < sizeof(cudaIpcMemHandle_t); b++) if ((unsigned char)p[b] != 0xff) { none = false; break; }
    if (none) { okflag = 0; break; }
    memcpy(&A->boff[(size_t)r * nB], p + sizeof h, nB * sizeof(long long));
    memcpy(&A->foff[(size_t)r * nF], p + sizeof h + nB * sizeof(long long), nF * sizeof(long long));
    if (r == rank) { A->peer[r] = A->base; continue; }
    cudaIpcMemHandle_t hr; memcpy(&hr, p, sizeof hr);
    void* q = nullptr;
    if (cudaIpcOpenMemHandle(&q, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); okflag = 0; break; }
    A->peer[r] = (char*)q;
  }
  // every rank must take the same path: agree on success with a max-reduce of the failure flag
  {
    double* d_f = nullptr; double fl = okflag ? 0. : 1.;
    FV3_CUDA(c0, cudaMalloc(&d_f, sizeof(double)));
    FV3_CUDA(c0, cudaMemcpy(d_f, &fl, sizeof fl, cudaMemcpyHostToDevice));
    if (g_nccl.AllReduce(d_f, d_f, 1, /*ncclFloat64*/ 8, /*ncclMax*/ 2, comm, c0->stream) != 0) fl = 1.;
    else { cudaStreamSynchronize(c0->stream); cudaMemcpy(&fl, d_f, sizeof fl, cudaMemcpyDeviceToHost); }
    cudaFree(d_f);
    if (fl != 0.) okflag = 0;
  }
  if (!okflag) {   // stay on NCCL send / recv
    for (int r = 0; r < nranks; r++) if (r != rank && A->peer[r]) cudaIpcCloseMemHandle(A->peer[r]);
    cudaFree(A->base); cudaFree(A->d_err); delete A;
    return 0;
  }
  A->on = true;
  c0->halo->p2p = A;
  return 0;
}
// 1 when the peer-mapped exchange is active for this process's faces, else 0; *err (nullable) receives the arrival-timeout flag
extern "C" int fv3_halo_p2p_status(fv3_ctx* c, int* err) {
  if (!c || !c->halo || !c->halo->p2p || !c->halo->p2p->on) { if (err) *err = 0; return 0; }
  if (err) { cudaSetDevice(c->device); cudaMemcpy(err, c->halo->p2p->d_err, sizeof(int), cudaMemcpyDeviceToHost); }
  return 1;
}

// overlapped != 0: the exchange runs on ctx0's side stream after everything enqueued so far on the face streams, and the
// face streams do NOT wait for it (fv3_halo_wait joins them): the caller may enqueue work that touches neither the halo
// cells nor the edge cells of the group's fields in between (SURVEY 8e "Overlap": dyn_core hides the delp/pt exchange
// behind update_dz_d + Riem_Solver3, which read the compute domain of delp, pt only).
static int halo_exchange_impl(fv3_ctx** ctxs, int nctx, int group, int overlapped) {
  if (!ctxs || nctx < 1 || group < 0 || group >= FV3_NUM_HALO_GROUPS) return -1;
  fv3_ctx* c0 = ctxs[0];
  if (!c0->halo) return 0;   // no topology linked: frozen halo
  void* comm = c0->halo->comm;
  const int my_rank = c0->halo->my_rank;
  HaloPlan* hp0 = c0->halo;
  if (hp0->xpending) return fv3_fail(c0, -1, "halo_exchange: an overlapped exchange is still pending (call fv3_halo_wait first)");
  cudaSetDevice(c0->device);
  if (overlapped && !hp0->xstream) {
    FV3_CUDA(c0, cudaStreamCreateWithFlags(&hp0->xstream, cudaStreamNonBlocking));
    FV3_CUDA(c0, cudaEventCreateWithFlags(&hp0->xdone, cudaEventDisableTiming));
  }
  cudaStream_t st = overlapped ? hp0->xstream : c0->stream;
  // all faces of one process share one stream ordering for the exchange: make that stream wait for the face streams
  // (single-GPU multi-face mode enqueues every face on its own stream)
  for (int a = 0; a < nctx; a++) {
    if (ctxs[a]->stream == st) continue;
    cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventRecord(ev, ctxs[a]->stream); cudaStreamWaitEvent(st, ev, 0); cudaEventDestroy(ev);
  }
  // ---- remote: size the message buffers, pack, send/recv, unpack
  bool any_remote = false;
  for (int a = 0; a < nctx; a++)
    for (int t = 0; t < 6; t++) {
      const int r = ctxs[a]->halo->tile_rank[t];
      if (r >= 0 && r != my_rank && !ctxs[a]->halo->peer[t]) any_remote = true;
    }
  if (any_remote && !comm) return fv3_fail(c0, -1, "halo_exchange: remote faces but no communicator");
  struct Msg { fv3_ctx* c; int tile; int rank; size_t nsend, nrecv; };
  std::vector<Msg> msgs;
  P2PArena* A = hp0->p2p;
  const bool p2p = any_remote && A && A->on;
  const unsigned long long seq = p2p ? ++A->seq[group] : 0;
  const int par = (int)(seq & 1);
  if (any_remote) {
    PackJobs pj; int npj = 0, pn = 0, pk = 0;
    FlagSet sig; int nsig = 0;
    auto flush_pack = [&]() {
      if (!npj) return;
      k_halo_pack_batch<<<dim3((pn + 127) / 128, pk, npj), 128, 0, st>>>(pj, c0->L.plane);
      c0->launches++;
      npj = 0; pn = 0; pk = 0;
    };
    for (int a = 0; a < nctx; a++) {
      fv3_ctx* c = ctxs[a]; HaloPlan* hp = c->halo; GroupPlan& gp = hp->grp[group];
      if (c->L.plane != c0->L.plane) return fv3_fail(c, -1, "halo_exchange: the faces of a process must share one plane layout");
      for (int t = 1; t <= 6; t++) {
        const int r = hp->tile_rank[t - 1];
        if (r < 0 || r == my_rank || hp->peer[t - 1]) continue;
        size_t ns = 0, nr = 0;
        for (size_t si = 0; si < gp.specs.size(); si++) {
          const int ncomp = gp.specs[si].fy >= 0 ? 2 : 1;
          for (int ci = 0; ci < ncomp; ci++)
            for (int sc = 0; sc < ncomp; sc++) {
              ns += (size_t)gp.send[si][t - 1][ci][sc].n * gp.specs[si].nk;
              nr += (size_t)gp.tab[si][ci][t - 1][sc].n * gp.specs[si].nk;
            }
        }
        if (ns == 0 && nr == 0) continue;
        const size_t need = std::max(ns, nr);
        // peer-mapped mode: the pack kernels store into the receiver's arena: its buffer for (its face t, my face, phase, parity)
        double* pack_dst = nullptr;
        if (p2p) {
          if (ns) {
            const long long bo = A->B(r, t - 1, c->tile - 1, group, par), fo = A->F(r, t - 1, c->tile - 1, group);
            if (bo < 0 || fo < 0) return fv3_fail(c, -1, "halo_exchange: the receiving rank's arena has no buffer for this message");
            pack_dst = (double*)(A->peer[r] + bo);
          }
        } else if (need > hp->bufcap[t - 1] || !hp->sendbuf[t - 1]) {
          const size_t cap = need + need / 4;
          cudaFree(hp->sendbuf[t - 1]); cudaFree(hp->recvbuf[t - 1]);
          FV3_CUDA(c, cudaMalloc(&hp->sendbuf[t - 1], cap * sizeof(double)));
          FV3_CUDA(c, cudaMalloc(&hp->recvbuf[t - 1], cap * sizeof(double)));
          hp->bufcap[t - 1] = cap;
        }
        // pack: every table of the message is a job of the exchange's one batched launch
        size_t off = 0;
        for (size_t si = 0; si < gp.specs.size(); si++) {
          const FieldSpec& fs = gp.specs[si];
          const int ncomp = fs.fy >= 0 ? 2 : 1;
          for (int ci = 0; ci < ncomp; ci++)
            for (int sc = 0; sc < ncomp; sc++) {
              const DevTable& tb = gp.send[si][t - 1][ci][sc];
              if (!tb.n) continue;
              pj.j[npj++] = PackJob{(p2p ? pack_dst : hp->sendbuf[t - 1]) + off, c->fld[sc == 0 ? fs.fx : fs.fy], tb.src, tb.n, fs.nk};
              pn = std::max(pn, tb.n); pk = std::max(pk, fs.nk);
              if (npj == GATHER_BATCH) flush_pack();
              off += (size_t)tb.n * fs.nk;
            }
        }
        if (p2p && ns) {   // once the message is complete in the receiver's memory its sequence number is published there
          if (nsig == P2P_MAXMSG) return fv3_fail(c, -1, "halo_exchange: more than 32 messages in one exchange");
          sig.f[nsig++] = (volatile unsigned long long*)(A->peer[r] + A->F(r, t - 1, c->tile - 1, group));
        }
        msgs.push_back({c, t, r, ns, nr});
      }
    }
    flush_pack();
    if (nsig) { k_p2p_signal<<<1, nsig, 0, st>>>(sig, seq); c0->launches++; }
    // canonical issue order per peer: (face on the lower rank, face on the higher rank), so the
    // i-th send of rank A to rank B meets the i-th receive B posts for A
    std::sort(msgs.begin(), msgs.end(), [&](const Msg& x, const Msg& y) {
      if (x.rank != y.rank) return x.rank < y.rank;
      const int xl = (my_rank < x.rank) ? x.c->tile : x.tile, xh = (my_rank < x.rank) ? x.tile : x.c->tile;
      const int yl = (my_rank < y.rank) ? y.c->tile : y.tile, yh = (my_rank < y.rank) ? y.tile : y.c->tile;
      return xl != yl ? xl < yl : xh < yh;
    });
    if (!p2p) g_nccl.GroupStart();
    for (const Msg& m : msgs) {
      if (p2p) break;
      HaloPlan* hp = m.c->halo;
      // tag-free ordering: one message per (my face, peer face) pair and phase; NCCL matches
      // sends and receives between a pair of ranks in issue order, and both ranks enumerate
      // (local face, remote face) pairs in the same canonical order (ascending face of the
      // lower rank first is not needed: each pair of ranks exchanges pairs sorted by
      // (min face, max face) because both loops run a,t ascending and faces are unique).
      if (m.nsend) g_nccl.Send(hp->sendbuf[m.tile - 1], m.nsend, /*ncclDouble*/ 8, m.rank, comm, st);
      if (m.nrecv) g_nccl.Recv(hp->recvbuf[m.tile - 1], m.nrecv, 8, m.rank, comm, st);
    }
    const int nrc = p2p ? 0 : g_nccl.GroupEnd();
    if (nrc != 0) return fv3_fail(c0, 1000 + nrc, "ncclGroupEnd failed");
  }
  // ---- local gathers.  Two-phase so that no face reads a halo another gather is writing:
  // sources are compute-domain points, destinations halo points -> disjoint, one phase suffices.
  {
    GatherJobs jobs; int nj = 0, nmax = 0, kmax = 0;
    auto flush = [&]() {
      if (!nj) return;
      k_halo_gather_batch<<<dim3((nmax + 127) / 128, kmax, nj), 128, 0, st>>>(jobs, c0->L.plane);
      c0->launches++;
      nj = 0; nmax = 0; kmax = 0;
    };
    for (int a = 0; a < nctx; a++) {
      fv3_ctx* c = ctxs[a]; HaloPlan* hp = c->halo; GroupPlan& gp = hp->grp[group];
      if (c->L.plane != c0->L.plane) return fv3_fail(c, -1, "halo_exchange: the faces of a process must share one plane layout");
      for (size_t si = 0; si < gp.specs.size(); si++) {
        const FieldSpec& fs = gp.specs[si];
        const int ncomp = fs.fy >= 0 ? 2 : 1;
        for (int ci = 0; ci < ncomp; ci++) {
          double* dst = c->fld[ci == 0 ? fs.fx : fs.fy];
          for (int t = 1; t <= 6; t++) {
            fv3_ctx* p = hp->peer[t - 1];
            if (!p) continue;
            for (int sc = 0; sc < ncomp; sc++) {
              const DevTable& tb = gp.tab[si][ci][t - 1][sc];
              if (!tb.n) continue;
              jobs.j[nj++] = GatherJob{dst, p->fld[sc == 0 ? fs.fx : fs.fy], tb.dst, tb.src, tb.sign, tb.n, fs.nk};
              nmax = std::max(nmax, tb.n); kmax = std::max(kmax, fs.nk);
              if (nj == GATHER_BATCH) flush();
            }
          }
        }
      }
    }
    flush();
  }
  // ---- unpack remote: await every message of the exchange (one thread per message), then one batched unpack
  {
    UnpackJobs uj; int nuj = 0, un = 0, uk = 0;
    FlagSet wt; int nwt = 0;
    std::vector<const double*> rbufs(msgs.size(), nullptr);
    for (size_t mi = 0; mi < msgs.size(); mi++) {
      const Msg& m = msgs[mi];
      fv3_ctx* c = m.c; HaloPlan* hp = c->halo;
      rbufs[mi] = hp->recvbuf[m.tile - 1];
      if (p2p) {
        if (!m.nrecv) continue;
        const long long bo = A->B(A->my_rank, c->tile - 1, m.tile - 1, group, par), fo = A->F(A->my_rank, c->tile - 1, m.tile - 1, group);
        if (bo < 0 || fo < 0) return fv3_fail(c, -1, "halo_exchange: no arena buffer for an expected message");
        rbufs[mi] = (const double*)(A->base + bo);
        if (nwt == P2P_MAXMSG) return fv3_fail(c, -1, "halo_exchange: more than 32 messages in one exchange");
        wt.f[nwt++] = (volatile unsigned long long*)(A->base + fo);
      }
    }
    if (nwt) { k_p2p_wait<<<1, nwt, 0, st>>>(wt, seq, A->d_err); c0->launches++; }   // holds the stream until the senders' flags arrive
    auto flush_unpack = [&]() {
      if (!nuj) return;
      k_halo_unpack_batch<<<dim3((un + 127) / 128, uk, nuj), 128, 0, st>>>(uj, c0->L.plane);
      c0->launches++;
      nuj = 0; un = 0; uk = 0;
    };
    for (size_t mi = 0; mi < msgs.size(); mi++) {
      const Msg& m = msgs[mi];
      if (p2p && !m.nrecv) continue;
      fv3_ctx* c = m.c; GroupPlan& gp = c->halo->grp[group];
      size_t off = 0;
      for (size_t si = 0; si < gp.specs.size(); si++) {
        const FieldSpec& fs = gp.specs[si];
        const int ncomp = fs.fy >= 0 ? 2 : 1;
        for (int ci = 0; ci < ncomp; ci++)
          for (int sc = 0; sc < ncomp; sc++) {
            const DevTable& tb = gp.tab[si][ci][m.tile - 1][sc];
            if (!tb.n) continue;
            uj.j[nuj++] = UnpackJob{c->fld[ci == 0 ? fs.fx : fs.fy], rbufs[mi] + off, tb.dst, tb.sign, tb.n, fs.nk};
            un = std::max(un, tb.n); uk = std::max(uk, fs.nk);
            if (nuj == GATHER_BATCH) flush_unpack();
            off += (size_t)tb.n * fs.nk;
          }
      }
    }
    flush_unpack();
  }
  if (overlapped) {
    cudaEventRecord(hp0->xdone, st);
    hp0->xpending = true;
  } else {
    // the other faces' streams wait for the exchange
    for (int a = 1; a < nctx; a++) {
      cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      cudaEventRecord(ev, st); cudaStreamWaitEvent(ctxs[a]->stream, ev, 0); cudaEventDestroy(ev);
    }
  }
  FV3_CUDA(c0, cudaGetLastError());
  return 0;
}

extern "C" {
int fv3_halo_exchange(fv3_ctx** ctxs, int nctx, int group) { return halo_exchange_impl(ctxs, nctx, group, 0); }
int fv3_halo_start(fv3_ctx** ctxs, int nctx, int group) { return halo_exchange_impl(ctxs, nctx, group, 1); }
int fv3_halo_wait(fv3_ctx** ctxs, int nctx) {
  if (!ctxs || nctx < 1) return -1;
  HaloPlan* hp0 = ctxs[0]->halo;
  if (!hp0 || !hp0->xpending) return 0;
  for (int a = 0; a < nctx; a++) cudaStreamWaitEvent(ctxs[a]->stream, hp0->xdone, 0);
  hp0->xpending = false;
  return 0;
}
}  // extern "C"

// Element-wise maximum of n doubles over all ranks of the library's communicator (mp_reduce_max of the reference,
// fv_mp_mod.F90; used by tracer_2d for the per-level CFL number).  vals: device buffer on c's device, reduced in place on
// c's stream.  Without a communicator (single process) this is a no-op.
int halo_allreduce_max(fv3_ctx* c, double* vals, int n) {
  HaloPlan* hp = c->halo;
  if (!hp || !hp->comm) return 0;
  if (!g_nccl.AllReduce) return fv3_fail(c, -4, "libnccl has no ncclAllReduce");
  const int nrc = g_nccl.AllReduce(vals, vals, (size_t)n, /*ncclFloat64*/ 8, /*ncclMax*/ 2, hp->comm, c->stream);
  if (nrc != 0) return fv3_fail(c, 1000 + nrc, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"));
  return 0;
}
