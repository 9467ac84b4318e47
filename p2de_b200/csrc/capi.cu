// capi.cu — C ABI of libp2de_b200.so (include/p2de_b200.h): handle, setup, launches.
//
// Host language note: the reference's host is Julia (absent from this image); this layer is what
// Julia reaches through `ccall` (julia/P2DEB200.jl) and what the Python mirror (p2de_b200/) and
// the tests reach through ctypes.  No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/p2de_b200.h"
#include "kernels2d.cuh"
#include "stage_fast.cuh"
#include "stage_subcell.cuh"
#include "gauss.cuh"
#include "kernels1d.cuh"

using namespace p2de;

#ifndef P2DE_EPB
#define P2DE_EPB 16   // elements per CTA (x 2*N1D line threads)
#endif

namespace {

thread_local std::string g_create_error;

// NCCL is resolved at run time (dlopen) so that the library also loads on machines without it;
// only p2de_comm_* need it.  Inside a torch process this finds the libnccl torch already loaded.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};

const NcclApi *nccl_api(std::string *why) {
  static NcclApi api;
  static int state = 0;   // 0 untried, 1 ok, -1 failed
  static std::string err;
  if (state == 0) {
    void *lib = nullptr;
    const char *env = getenv("P2DE_NCCL_LIB");
    const char *cands[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
      if (!c) continue;
      lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load NCCL (set P2DE_NCCL_LIB): ") + dlerror(); state = -1; }
    else {
      bool ok = true;
      auto sym = [&](const char *name) { void *f = dlsym(lib, name); if (!f) { ok = false; err = std::string("NCCL symbol missing: ") + name; } return f; };
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
      api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
      api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
      state = ok ? 1 : -1;
    }
  }
  if (state != 1) { if (why) *why = err; return nullptr; }
  return &api;
}

}  // namespace

struct p2de_handle {
  p2de_config cfg{};
  int N1D = 0, Nq = 0, Nfp = 0, Nc = 0, Nd = 0, Ns = 3;
  int dim = 2;
  long long nLloc = 0;    // L_local entries per element: (Nq + N1D) * Nd
  Tables1D<2> u2{}; Tables1D<3> u3{}; Tables1D<4> u4{}; Tables1D<5> u5{};   // 1D path
  double rxJ1 = 1.0;
  long long K = 0;
  int mode = 0;
  bool fast = false;   // default flux configuration -> FAST kernel variant (kernels2d.cuh)
  bool gauss = false;     // 2D GaussCollocation: entropy projection kernel + generic stage kernel
  bool nodewise = false;  // NodewiseScaledExtrapolation
  double *utf = nullptr;          // [K][Nfp][4] entropy-projected face states (halo rows like the state)
  double *theta_local_dev = nullptr, *theta_dev = nullptr;   // [Ns][K][Nfp], [Ns][K]
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  bool have_state = false;

  // tables (one of these is used, by N1D)
  Tables2D<2> t2{};
  Tables2D<3> t3{};
  Tables2D<4> t4{};
  Tables2D<5> t5{};
  MeshTopo topo{};
  double Jq = 0, Jcons = 0;
  std::vector<double> wq;

  // device memory
  double *U[2] = {nullptr, nullptr};  // state ping-pong; U[cur] is Uq, the other one is resW / next
  bool direct = false;                // FAST subcell path: stages 2/3 write the state from the stage kernel
  bool defer = false;                 // ... and stage 2 forms the stage-1 combine on the fly (P2DE_NO_DEFER=1 disables)
  bool rpre_w = false;                // the rpre buffer holds stage 1's W = U + cap rhsU (run_stage: wform), not rhsU
  int cur = 0;
  double *rhsL = nullptr, *dF = nullptr, *lpre = nullptr, *rhsU = nullptr;
  double *rpre = nullptr, *dFend = nullptr;   // FAST subcell scratch
  double *VDM_inv = nullptr;                  // [Np, Nq] (Hennemann indicator)
  unsigned long long *smin_bits = nullptr;    // global min of s_modified at t0 (min-entropy bounds)
  int entropy_bound = 0;                      // 0 none, 1 min entropy, 2 relaxed min entropy
  bool slim = false;                          // generic kernel in its compile-time Gauss configuration: FAST-family scratch
  int tvd = 0;                                // TVD*Bound: needs the low-order rhs of the neighbours (MODE_LOW pre-pass)
  int cell_entropy = 0;                       // 0 none, 1 cell entropy, 2 relaxed cell entropy
  double *rhsLpre = nullptr;                  // [K][Nq][4] (+ halo rows) low-order rhs of the pre-pass
  double *fstar = nullptr;                    // [K][Nfp][2][4] Gauss + cell entropy: normal components of fstar_H, fstar_L
  double *Lz = nullptr;       // [K, Ns]
  double *Llocal = nullptr;   // [Nq+N1D, Nd, K, Ns]
  double *rhsH_diag = nullptr, *rhsL_diag = nullptr;
  unsigned long long *dt_bits = nullptr;
  unsigned long long *dbg = nullptr, *dbg_buf = nullptr;   // p2de_debug_counters: dbg = dbg_buf while counting
  double *partial = nullptr;  // reduction scratch
  double *tab_dev = nullptr;  // device copy of the Tables2D<N1D> struct (coalesced load into shared memory)
  int *mapP32 = nullptr, *bcflag = nullptr;
  double *Ival = nullptr;
  unsigned char *bc_type[4] = {nullptr, nullptr, nullptr, nullptr};
  double *bc_val[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<void *> owned;
  double last_dt = 0;
  // DataHistory snapshots (SSPRK33.jl:45-55): a ring of state copies on the device, filled by p2de_ssp33_run
  int snap_slots = 0;
  int64_t snap_interval = 0, snap_count = 0;
  double *snap_buf = nullptr;
  std::vector<double> snap_t;
  std::vector<int64_t> snap_step;
  double *err_stage = nullptr;   // staging buffer of p2de_calculate_error
  // multi-GPU (y-stripes, one handle per GPU): NCCL communicator and stripe neighbours
  int rank = 0, nranks = 1, rank_lo = 0, rank_hi = 0;
  bool has_lo = false, has_hi = false;
  ncclComm_t comm = nullptr;
  // overlapped halo exchange (run_step_overlapped): NCCL runs on its own high-priority stream
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_rows = nullptr, ev_all = nullptr, ev_halo = nullptr, ev_dt = nullptr;
  int overlap_min_rows = 64;   // P2DE_OVERLAP_MIN_ROWS
  bool overlap = false;        // p2de_comm_init: the direct schedule exchanges each stage's OUTPUT rows behind its interior rows
  bool halo_current = false;   // the halo rows of U[cur] hold the neighbours' current boundary rows
  bool halo_pending = false;   // ev_halo has been recorded and not yet waited for by the compute stream
  // peer-to-peer halo rows (p2p_setup): the neighbouring stripes' buffers mapped through CUDA IPC; boundary rows travel by
  // copy engine over NVLink straight into the neighbour's halo rows, a flag word per direction says "rows of exchange #seq
  // have landed".  Buffer index: 0, 1 = the two state buffers, 2 = rpre.
  bool p2p = false;
  double *peer_lo[3] = {nullptr, nullptr, nullptr}, *peer_hi[3] = {nullptr, nullptr, nullptr};   // first OWNED entry of the neighbour's buffer
  unsigned int *flags = nullptr;                 // [2]: written by the lower / upper neighbour
  unsigned int *peer_lo_flags = nullptr, *peer_hi_flags = nullptr;
  std::vector<void *> ipc_opened;
  unsigned int seq = 0;                          // exchanges issued so far
  unsigned int wait_seq = 0;                     // flag value the compute stream has to see before the next boundary launch (0 = none pending)
  // optional per-kernel timing (p2de_profile): one event pair per launch, on h->stream
  bool profiling = false;
  struct ProfRec { cudaEvent_t a, b; int kid; };
  std::vector<ProfRec> prof;
};

namespace {

int fail(p2de_handle *h, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CU(h, call)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail(h, P2DE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// every entry point runs on the handle's device whatever the caller's current device is, and restores the caller's
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define DEV(h) DeviceGuard dev_guard_((h)->device)

template <class T>
int dev_alloc(p2de_handle *h, T **p, size_t n) {
  void *q = nullptr;
  CU(h, cudaMalloc(&q, n * sizeof(T)));
  h->owned.push_back(q);
  *p = static_cast<T *>(q);
  return 0;
}

void prof_begin(p2de_handle *h, int kid) {
  if (!h->profiling) return;
  p2de_handle::ProfRec r{};
  r.kid = kid;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, h->stream);
  h->prof.push_back(r);
}
void prof_end(p2de_handle *h) {
  if (h->profiling && !h->prof.empty()) cudaEventRecord(h->prof.back().b, h->stream);
}
void prof_clear(p2de_handle *h) {
  for (auto &r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  h->prof.clear();
}

// array of `n` owned doubles plus one halo row of `row` doubles on each side; *p points at the
// first OWNED entry, so element indices -Kx..-1 and K..K+Kx-1 address the halo rows.
int dev_alloc_halo(p2de_handle *h, double **p, size_t n, size_t row) {
  double *base = nullptr;
  if (int rc = dev_alloc(h, &base, n + 2 * row)) return rc;
  CU(h, cudaMemset(base, 0, (n + 2 * row) * sizeof(double)));
  *p = base + row;
  return 0;
}

#define NC(h, call)                                                                                  \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess) return fail(h, P2DE_ERR_NCCL, "%s: %s", #call, nccl_api(nullptr)->GetErrorString(r_)); \
  } while (0)

// halo exchange of one boundary element row (`row` doubles) with the stripes below and above.
// Issue order (sends up, down; receives from below, above) keeps the pairing right when both
// neighbours are the same rank (2 ranks, periodic).
int exchange_rows(p2de_handle *h, double *owned, size_t row, cudaStream_t stream = nullptr, bool on_stream = false) {
  if (!h->comm) return 0;
  if (!on_stream) {
    stream = h->stream;
    // (ordering against the overlapped schedule's communication stream: wait for whatever it still has in flight)
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
  }
  const NcclApi *n = nccl_api(nullptr);
  const size_t nown = (size_t)h->cfg.Ky * row;
  NC(h, n->GroupStart());
  if (h->has_hi) NC(h, n->Send(owned + nown - row, row, ncclDouble, h->rank_hi, h->comm, stream));
  if (h->has_lo) NC(h, n->Send(owned, row, ncclDouble, h->rank_lo, h->comm, stream));
  if (h->has_lo) NC(h, n->Recv(owned - row, row, ncclDouble, h->rank_lo, h->comm, stream));
  if (h->has_hi) NC(h, n->Recv(owned + nown, row, ncclDouble, h->rank_hi, h->comm, stream));
  NC(h, n->GroupEnd());
  return 0;
}

template <int N1D>
Tables2D<N1D> &tables(p2de_handle *h);
template <> Tables2D<2> &tables<2>(p2de_handle *h) { return h->t2; }
template <> Tables2D<3> &tables<3>(p2de_handle *h) { return h->t3; }
template <> Tables2D<4> &tables<4>(p2de_handle *h) { return h->t4; }
template <> Tables2D<5> &tables<5>(p2de_handle *h) { return h->t5; }

#ifndef P2DE_EPB5
#define P2DE_EPB5 12   // N=4 (N1D=5): elements per CTA (120 threads, 3 CTAs/SM at 168 registers: measured best of 6..16)
#endif
template <int N1D> struct Launch { static constexpr int EPB = N1D == 5 ? P2DE_EPB5 : P2DE_EPB; };
#ifndef P2DE_EPB_SUB4
#define P2DE_EPB_SUB4 8   // subcell family at N=3: 8 elements (64 threads) per CTA, 8 CTAs/SM: the CTAs' load phases interleave better (measured +3 %)
#endif
template <int N1D> struct LaunchSub { static constexpr int EPB = N1D == 4 ? P2DE_EPB_SUB4 : Launch<N1D>::EPB; };

// Extract the per-line tables from the caller's operators and verify the structure this
// kernel family relies on (tensor-product LGL collocation on a Cartesian mesh).
template <int N1D>
int build_tables(p2de_handle *h, const p2de_operators *o, const double GJ[4]) {
  constexpr int Nq = N1D * N1D, Nfp = 4 * N1D, Nh = Nq + Nfp;
  Tables2D<N1D> &T = tables<N1D>(h);
  const double tol = 0.0;
  auto node = [](int d, int line, int a) { return d == 0 ? a + line * N1D : line + a * N1D; };
  // fq2q must be the LGL boundary-node map in face order left, right, bottom, top
  for (int f = 0; f < Nfp; ++f) {
    int F = f / N1D, a = f % N1D;
    int expect = F == 0 ? a * N1D : F == 1 ? (N1D - 1) + a * N1D : F == 2 ? a : a + (N1D - 1) * N1D;
    if ((int)o->fq2q[f] - 1 != expect) return fail(h, P2DE_ERR_UNSUPPORTED, "fq2q[%d]=%lld is not the LGL tensor-product face map", f, (long long)o->fq2q[f]);
    T.fq2q[f] = expect;
  }
  // Vf: the 0/1 gather of collocated face nodes (Lobatto), or an extrapolation along the face node's own grid
  // line (Gauss); Vf_low must be the 0/1 gather fq2q in both cases (init.jl:178-185,324-336)
  for (int f = 0; f < Nfp; ++f)
    for (int j = 0; j < Nq; ++j) {
      const int F = f / N1D, a = f % N1D;
      const bool on_line = F < 2 ? (j / N1D == a) : (j % N1D == a);
      double v = o->Vf[f + (size_t)j * Nfp], e = (j == T.fq2q[f]) ? 1.0 : 0.0;
      if (!h->gauss && std::fabs(v - e) > 1e-13) return fail(h, P2DE_ERR_UNSUPPORTED, "Vf is not a 0/1 gather but the basis is not GaussCollocation");
      if (h->gauss && !on_line && v != 0.0) return fail(h, P2DE_ERR_UNSUPPORTED, "Vf couples a face node with nodes off its grid line");
      if (h->gauss && (!o->Vf_low || std::fabs(o->Vf_low[f + (size_t)j * Nfp] - e) > 1e-13)) return fail(h, P2DE_ERR_UNSUPPORTED, "Vf_low is not the nearest-node gather");
    }
  for (int d = 0; d < 2; ++d)
    for (int line = 0; line < N1D; ++line)
      for (int e = 0; e < 2; ++e)
        for (int a = 0; a < N1D; ++a)
          T.VfL[d][line][e][a] = o->Vf[((2 * d + e) * N1D + line) + (size_t)node(d, line, a) * Nfp];
  for (int d = 0; d < 2; ++d) {
    const double *S = o->Srsh_db[d], *S0 = o->Srs0[d];
    const double g = GJ[d == 0 ? 0 : 3];
    if (GJ[1] != 0.0 || GJ[2] != 0.0) return fail(h, P2DE_ERR_UNSUPPORTED, "non-Cartesian geometric factors (sxJ, ryJ != 0)");
    // every volume-volume nonzero of S_d must couple two nodes of one d-line
    for (int i = 0; i < Nq; ++i)
      for (int j = 0; j < Nq; ++j) {
        bool same_line = d == 0 ? (i / N1D == j / N1D) : (i % N1D == j % N1D);
        if (!same_line && (std::fabs(S[i + (size_t)j * Nh]) > tol || std::fabs(S0[i + (size_t)j * Nq]) > tol))
          return fail(h, P2DE_ERR_UNSUPPORTED, "S_%d couples nodes of different grid lines", d);
      }
    for (int line = 0; line < N1D; ++line) {
      for (int a = 0; a < N1D; ++a)
        for (int b = 0; b < N1D; ++b)
          T.SHt[d][a][b][line] = T.SH[d][line][a][b] = g * S[node(d, line, a) + (size_t)node(d, line, b) * Nh];
      for (int a = 0; a < N1D; ++a) {
        T.S0[d][line][a] = (a + 1 < N1D) ? g * S0[node(d, line, a + 1) + (size_t)node(d, line, a) * Nq] : 0.0;
        for (int b = 0; b < N1D; ++b)   // low-order operator must be nearest-neighbour
          if (std::abs(a - b) > 1 && std::fabs(S0[node(d, line, a) + (size_t)node(d, line, b) * Nq]) > tol)
            return fail(h, P2DE_ERR_UNSUPPORTED, "low-order S0 is not tridiagonal along lines");
      }
      for (int e = 0; e < 2; ++e) {
        int f = (2 * d + e) * N1D + line;
        for (int a = 0; a < N1D; ++a) {   // hybridized block -B E (face row, volume column) and its transpose E^T B
          T.SHf[d][line][e][a] = g * S[(Nq + f) + (size_t)node(d, line, a) * Nh];
          if (S[(Nq + f) + (size_t)node(d, line, a) * Nh] != -S[node(d, line, a) + (size_t)(Nq + f) * Nh])
            return fail(h, P2DE_ERR_UNSUPPORTED, "Srsh_db is not skew-symmetric in its face-volume block");
        }
        T.Bf[d][line][e] = g * o->Brs[d][f];
        if (o->Brs[1 - d][f] != 0.0) return fail(h, P2DE_ERR_UNSUPPORTED, "face %d has a tangential boundary weight", f);
      }
    }
    // faces of the other direction must have zero weight in B_d
    for (int f = 0; f < Nfp; ++f)
      if ((f / (2 * N1D)) != d && o->Brs[d][f] != 0.0) return fail(h, P2DE_ERR_UNSUPPORTED, "B_%d nonzero on a face of the other direction", d);
  }
  for (int i = 0; i < Nq; ++i) {
    T.wq[i] = o->wq[i];
    T.rwJ[i] = 1.0 / (h->Jq * o->wq[i]);
    T.rwJl[0][i / N1D][i % N1D] = T.rwJ[i]; T.rwJl[1][i % N1D][i / N1D] = T.rwJ[i];
    T.minv[i] = o->MinvVhT[i + (size_t)i * Nq];
    if (std::fabs(T.minv[i] * o->wq[i] - 1.0) > 1e-13) return fail(h, P2DE_ERR_UNSUPPORTED, "MinvVhT diagonal is not 1/wq");
    for (int j = 0; j < Nq; ++j)
      if (j != i && o->MinvVhT[i + (size_t)j * Nq] != 0.0) return fail(h, P2DE_ERR_UNSUPPORTED, "mass matrix is not diagonal");
  }
  for (int f = 0; f < Nfp; ++f) {
    T.minvf[f] = o->MinvVfT[T.fq2q[f] + (size_t)f * Nq];
    for (int i = 0; i < Nq; ++i)   // M^-1 Vf^T = (1/wq) Vf^T (LGL: the 1/wq-scaled face gather)
      if (std::fabs(o->MinvVfT[i + (size_t)f * Nq] * o->wq[i] - o->Vf[f + (size_t)i * Nfp]) > 1e-12) return fail(h, P2DE_ERR_UNSUPPORTED, "MinvVfT is not (1/wq) Vf^T");
  }
  return 0;
}

void structured_partner(int N1D, int Kx, int Ky, bool px, bool py, long long k, int f, long long *kP, int *fP) {
  int ix = (int)(k % Kx), iy = (int)(k / Kx), F = f / N1D, a = f % N1D;
  int jx = ix + (F == 0 ? -1 : F == 1 ? 1 : 0), jy = iy + (F == 2 ? -1 : F == 3 ? 1 : 0);
  bool out = jx < 0 || jx >= Kx || jy < 0 || jy >= Ky;
  if (out) {
    bool wrap = F < 2 ? px : py;
    if (!wrap) { *kP = k; *fP = f; return; }
    jx = (jx + Kx) % Kx; jy = (jy + Ky) % Ky;
  }
  *kP = jx + (long long)jy * Kx; *fP = (F ^ 1) * N1D + a;
}

int setup_topology(p2de_handle *h, const p2de_bcdata *bc) {
  const int N1D = h->N1D, Nfp = h->Nfp, Kx = h->cfg.Kx, Ky = h->cfg.Ky;
  const long long K = h->K;
  MeshTopo &M = h->topo;
  M.K = K; M.Kx = Kx; M.Ky = Ky;
  bool structured = (long long)Kx * Ky == K;
  int px = bc->periodic_x != 0, py = bc->periodic_y != 0;
  if (bc->mapP && structured) {
    bool found = false;
    for (int c = 0; c < 4 && !found; ++c) {
      bool tx = c & 1, ty = c & 2, ok = true;
      for (long long k = 0; k < K && ok; ++k)
        for (int f = 0; f < Nfp; ++f) {
          long long kP; int fP;
          structured_partner(N1D, Kx, Ky, tx, ty, k, f, &kP, &fP);
          if (bc->mapP[k * Nfp + f] - 1 != kP * Nfp + fP) { ok = false; break; }
        }
      if (ok) { found = true; px = tx; py = ty; }
    }
    structured = found;
  } else if (!bc->mapP && !structured) {
    return fail(h, P2DE_ERR_ARG, "mapP == NULL needs K == Kx*Ky");
  }
  // boundary data must sit on domain-boundary faces for the structured tables
  auto on_boundary = [&](long long idx) {
    long long k = idx / Nfp; int f = (int)(idx % Nfp), F = f / N1D;
    int ix = (int)(k % Kx), iy = (int)(k / Kx);
    return F == 0 ? ix == 0 : F == 1 ? ix == Kx - 1 : F == 2 ? iy == 0 : iy == Ky - 1;
  };
  for (long long i = 0; i < bc->nI; ++i) {
    long long idx = bc->mapI[i] - 1;
    if (idx < 0 || idx >= K * Nfp) return fail(h, P2DE_ERR_ARG, "mapI[%lld] out of range", i);
    if (structured && !on_boundary(idx)) structured = false;
  }
  for (long long i = 0; i < bc->nO; ++i) {
    long long idx = bc->mapO[i] - 1;
    if (idx < 0 || idx >= K * Nfp) return fail(h, P2DE_ERR_ARG, "mapO[%lld] out of range", i);
    if (structured && !on_boundary(idx)) structured = false;
  }
  if (structured) {
    M.periodic_x = px; M.periodic_y = py;
    M.mapP32 = nullptr; M.bcflag = nullptr; M.Ival = nullptr;
    if (bc->nI + bc->nO > 0) {
      std::vector<unsigned char> type[4];
      std::vector<double> val[4];
      for (int F = 0; F < 4; ++F) {
        size_t len = (size_t)(F < 2 ? Ky : Kx) * N1D;
        type[F].assign(len, 0); val[F].assign(len * 4, 0.0);
      }
      auto put = [&](long long idx, int t, const double *v) {
        long long k = idx / Nfp; int f = (int)(idx % Nfp), F = f / N1D, a = f % N1D;
        int pos = F < 2 ? (int)(k / Kx) : (int)(k % Kx);
        type[F][(size_t)pos * N1D + a] = (unsigned char)t;
        if (v) std::memcpy(&val[F][((size_t)pos * N1D + a) * 4], v, 4 * sizeof(double));
      };
      for (long long i = 0; i < bc->nI; ++i) put(bc->mapI[i] - 1, 1, bc->Ival + 4 * i);
      for (long long i = 0; i < bc->nO; ++i) put(bc->mapO[i] - 1, 2, nullptr);
      for (int F = 0; F < 4; ++F) {
        bool any = false;
        for (unsigned char t : type[F]) any |= t != 0;
        if (!any) continue;
        if (int rc = dev_alloc(h, &h->bc_type[F], type[F].size())) return rc;
        if (int rc = dev_alloc(h, &h->bc_val[F], val[F].size())) return rc;
        CU(h, cudaMemcpy(h->bc_type[F], type[F].data(), type[F].size(), cudaMemcpyHostToDevice));
        CU(h, cudaMemcpy(h->bc_val[F], val[F].data(), val[F].size() * sizeof(double), cudaMemcpyHostToDevice));
        M.bc_type[F] = h->bc_type[F]; M.bc_val[F] = h->bc_val[F];
      }
    }
    return 0;
  }
  // generic gather tables
  if (!bc->mapP) return fail(h, P2DE_ERR_ARG, "boundary data off the domain boundary needs an explicit mapP");
  if (K * Nfp >= (1ll << 31)) return fail(h, P2DE_ERR_UNSUPPORTED, "generic mapP limited to 2^31 face nodes");
  std::vector<int> m32((size_t)K * Nfp), fl((size_t)K * Nfp, 0);
  for (size_t i = 0; i < m32.size(); ++i) {
    long long v = bc->mapP[i] - 1;
    if (v < 0 || v >= K * Nfp) return fail(h, P2DE_ERR_ARG, "mapP[%zu] out of range", i);
    m32[i] = (int)v;
  }
  for (long long i = 0; i < bc->nI; ++i) fl[bc->mapI[i] - 1] = (int)(i + 1);
  for (long long i = 0; i < bc->nO; ++i) fl[bc->mapO[i] - 1] = -1;
  if (int rc = dev_alloc(h, &h->mapP32, m32.size())) return rc;
  if (int rc = dev_alloc(h, &h->bcflag, fl.size())) return rc;
  CU(h, cudaMemcpy(h->mapP32, m32.data(), m32.size() * sizeof(int), cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->bcflag, fl.data(), fl.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (bc->nI > 0) {
    if (int rc = dev_alloc(h, &h->Ival, (size_t)bc->nI * 4)) return rc;
    CU(h, cudaMemcpy(h->Ival, bc->Ival, (size_t)bc->nI * 4 * sizeof(double), cudaMemcpyHostToDevice));
  }
  M.mapP32 = h->mapP32; M.bcflag = h->bcflag; M.Ival = h->Ival;
  return 0;
}

// 0 none, 1 *CellEntropyBound, 2 *RelaxedCellEntropyBound (Solver.jl:47-63)
int cell_entropy_of(int bound) {
  if (bound == P2DE_BOUND_POS_CELL_ENTROPY || bound == P2DE_BOUND_TVD_CELL_ENTROPY) return 1;
  if (bound == P2DE_BOUND_POS_RELAXED_CELL_ENTROPY || bound == P2DE_BOUND_TVD_RELAXED_CELL_ENTROPY) return 2;
  return 0;
}

__global__ void set_dt_kernel(unsigned long long *dt_bits, double v, int with_flag) {
  dt_bits[0] = (unsigned long long)__double_as_longlong(v);
  if (with_flag) dt_bits[1] = (unsigned long long)__double_as_longlong(1.0);   // "every candidate was a positive number" (dt_publish)
}

// minimum(s_modified) over all nodes (initialize_s_modified!, subcell.jl:19-35); s_modified > 0
__global__ void smin1d_kernel(const double *U, long long n_nodes, double gamma, unsigned long long *out_bits) {
  double m = INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x)
    m = fmin(m, s_modified1(gamma, load1(U + i * 3)));
  for (int off = 16; off > 0; off >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m < INFINITY) atomicMin(out_bits, (unsigned long long)__double_as_longlong(m));
}
__global__ void smin_kernel(const double *U, long long n_nodes, double gamma, unsigned long long *out_bits) {
  double m = INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x)
    m = fmin(m, s_modified(gamma, load_cons(U + i * 4)));
  for (int off = 16; off > 0; off >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m < INFINITY) atomicMin(out_bits, (unsigned long long)__double_as_longlong(m));
}

__global__ void reduce_kernel(const double *U, const double *wq, int Nq, long long n_nodes, double J, int what, double *partial) {
  __shared__ double sh[256];
  double acc = what == P2DE_REDUCE_CONSERVATION ? 0.0 : INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x) {
    Cons2 u = load_cons(U + i * 4);
    if (what == P2DE_REDUCE_CONSERVATION) acc += J * wq[i % Nq] * (((u.rho + u.m1) + u.m2) + u.E);
    else if (what == P2DE_REDUCE_MIN_RHO) acc = fmin(acc, u.rho);
    else acc = fmin(acc, rhoe2(u));
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = what == P2DE_REDUCE_CONSERVATION ? sh[threadIdx.x] + sh[threadIdx.x + s] : fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// calculate_error (postprocess.jl:1-28): per component sum wJ |ex - U|, sum wJ |ex - U|^2, max |ex - U| and the same of ex.
// One chunk of nodes per launch; per-block partials [block][6][Nc] are added up on the host.
__global__ void error_norm_kernel(const double *U, const double *ex, const double *wq, int Nq, int Nc, long long node0, long long n_nodes,
                                  double J, double *partial) {
  __shared__ double sh[256];
  double acc[6][4];
  for (int q = 0; q < 6; ++q) for (int c = 0; c < 4; ++c) acc[q][c] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x) {
    const double wJ = J * wq[(node0 + i) % Nq];
    for (int c = 0; c < Nc; ++c) {
      const double e = ex[i * Nc + c], d = fabs(e - U[(node0 + i) * Nc + c]);
      acc[0][c] += wJ * d; acc[1][c] += wJ * d * d; acc[2][c] = fmax(acc[2][c], d);
      acc[3][c] += wJ * fabs(e); acc[4][c] += wJ * e * e; acc[5][c] = fmax(acc[5][c], fabs(e));
    }
  }
  for (int q = 0; q < 6; ++q)
    for (int c = 0; c < Nc; ++c) {
      sh[threadIdx.x] = acc[q][c];
      __syncthreads();
      for (int st = blockDim.x / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sh[threadIdx.x] = (q == 2 || q == 5) ? fmax(sh[threadIdx.x], sh[threadIdx.x + st]) : sh[threadIdx.x] + sh[threadIdx.x + st];
        __syncthreads();
      }
      if (threadIdx.x == 0) partial[(blockIdx.x * 6 + q) * 4 + c] = sh[0];
      __syncthreads();
    }
}

template <int N1D, int MODE, bool FAST>
int launch_stage_t(p2de_handle *h, const StageArgs &A) {
  constexpr bool SUBK = FAST && MODE == MODE_SUBCELL;   // the default configuration with the subcell limiter: stage_subcell.cuh
  constexpr int EPB = SUBK ? LaunchSub<N1D>::EPB : Launch<N1D>::EPB, TPE = 2 * N1D;
  constexpr int TBL = FAST ? fast_table_doubles<N1D>() : (int)((sizeof(Tables2D<N1D>) + 15) / 16) * 2;
  const size_t base = sizeof(double) * (TBL + (size_t)EPB * (FAST ? fast_smem_doubles_per_elem<N1D, MODE>() : stage_smem_doubles_per_elem<N1D, MODE>()));
  const bool sub = !FAST && MODE == MODE_SUBCELL;
  size_t smem = base + (sub ? sizeof(double) * EPB * stage_smem_extra_doubles_per_elem<N1D>(A.tvd != 0, A.cell_entropy != 0) : 0);
  // kernel of the subcell family (stage_subcell.cuh): stage 1 in W form when run_stage asked for it, stages 2/3 fused
  const bool nodiag = !A.rhsL_diag && !A.rhsH_diag;
  const bool sub_s1 = SUBK && A.wform && nodiag && A.nstage == 1 && !A.fuse && !A.defer_add;
  const bool sub_s23 = SUBK && (A.defer_add || (nodiag && A.nstage != 1 && A.fuse));
  if (SUBK) smem = sizeof(double) * (fast_table_doubles<N1D>() + (size_t)EPB * subcell_smem_doubles_per_elem<N1D>(sub_s1 || sub_s23));
  static const size_t smem_pad = [] { const char *pad = getenv("P2DE_SMEM_PAD"); return pad ? (size_t)atoi(pad) : (size_t)0; }();
  smem += smem_pad;   // profiling aid: lowers the number of resident CTAs
  void (*kern)(const StageArgs, const MeshTopo, const Tables2D<N1D>);
  if constexpr (SUBK) kern = stage_subcell_rt<N1D, EPB>;
  else if constexpr (FAST) kern = stage_kernel_fast<N1D, MODE, EPB>;
  else kern = stage_kernel<N1D, MODE, EPB, 0>;
  // the shipped-examples configuration on Gauss nodes has its options compiled in (kernels2d.cuh: CFG = 1)
  const bool default_gauss = !FAST && MODE == MODE_SUBCELL && h->slim;
  if constexpr (!FAST && MODE == MODE_SUBCELL) {
    if (default_gauss) kern = stage_kernel<N1D, MODE, EPB, 1>;
  }
  // (the attribute is per device and per function: one slot per device for this instantiation)
  static size_t attr_set_dev[64] = {};
  size_t &attr_set = attr_set_dev[h->device & 63];
  if constexpr (SUBK) {
    const size_t smem_full = sizeof(double) * (fast_table_doubles<N1D>() + (size_t)EPB * subcell_smem_doubles_per_elem<N1D>(false)) + smem_pad;
    if (smem_full > attr_set) {
      CU(h, cudaFuncSetAttribute(stage_subcell_rt<N1D, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
      CU(h, cudaFuncSetAttribute(stage_subcell_s1<N1D, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
      CU(h, cudaFuncSetAttribute(stage_subcell_s2<N1D, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
      CU(h, cudaFuncSetAttribute(stage_subcell_s3<N1D, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full));
      attr_set = smem_full;
    }
    if (A.defer_add) kern = stage_subcell_s2<N1D, EPB>;
    else if (sub_s1) kern = stage_subcell_s1<N1D, EPB>;
    else if (sub_s23) kern = stage_subcell_s3<N1D, EPB>;
  } else if (smem > attr_set) {
    CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if constexpr (!FAST && MODE == MODE_SUBCELL) {
      CU(h, cudaFuncSetAttribute(stage_kernel<N1D, MODE, EPB, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU(h, cudaFuncSetAttribute(stage_kernel<N1D, MODE, EPB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    attr_set = smem;
  }
  dim3 grid((unsigned)((h->K + EPB - 1) / EPB));
  StageArgs A2 = A;
  if (FAST && !h->topo.mapP32 && h->cfg.Kx % EPB == 0 && h->cfg.Ky <= 65535) {
    A2.rowblocks = h->cfg.Kx / EPB;
    if (A.nrows > 0) grid = dim3((unsigned)A2.rowblocks, (unsigned)A.nrows);
    else { A2.row0 = 0; A2.row_stride = 1; grid = dim3((unsigned)A2.rowblocks, (unsigned)h->cfg.Ky); }
  } else if (A.nrows > 0) return fail(h, P2DE_ERR_STATE, "row-range launch needs the 2D grid (Kx a multiple of the batch size)");
  prof_begin(h, 0);
  kern<<<grid, EPB * TPE, smem, h->stream>>>(A2, h->topo, tables<N1D>(h));
  prof_end(h);
  CU(h, cudaGetLastError());
  h->launches++;
  return 0;
}
template <int N1D>
int launch_stage_n(p2de_handle *h, const StageArgs &A, bool low_prepass) {
  if (low_prepass) return launch_stage_t<N1D, MODE_LOW, false>(h, A);
  if (h->fast) {
    switch (h->mode) {
      case MODE_SUBCELL: return launch_stage_t<N1D, MODE_SUBCELL, true>(h, A);
      case MODE_ZHANGSHU: return launch_stage_t<N1D, MODE_ZHANGSHU, true>(h, A);
      case MODE_LOW: return launch_stage_t<N1D, MODE_LOW, true>(h, A);
      default: return launch_stage_t<N1D, MODE_HIGH, true>(h, A);
    }
  }
  switch (h->mode) {
    case MODE_SUBCELL: return launch_stage_t<N1D, MODE_SUBCELL, false>(h, A);
    case MODE_ZHANGSHU: return launch_stage_t<N1D, MODE_ZHANGSHU, false>(h, A);
    case MODE_LOW: return launch_stage_t<N1D, MODE_LOW, false>(h, A);
    default: return launch_stage_t<N1D, MODE_HIGH, false>(h, A);
  }
}
int launch_stage(p2de_handle *h, const StageArgs &A, bool low_prepass = false) {
  switch (h->N1D) {
    case 2: return launch_stage_n<2>(h, A, low_prepass);
    case 3: return launch_stage_n<3>(h, A, low_prepass);
    case 4: return launch_stage_n<4>(h, A, low_prepass);
    case 5: return launch_stage_n<5>(h, A, low_prepass);
  }
  return fail(h, P2DE_ERR_UNSUPPORTED, "N=%d", h->cfg.N);
}

template <int N1D, int MODE>
int launch_update_t(p2de_handle *h, const UpdateArgs &A) {
  constexpr int EPB = Launch<N1D>::EPB, TPE = 2 * N1D;
  unsigned grid = (unsigned)((h->K + EPB - 1) / EPB);
  prof_begin(h, 1);
  update_kernel<N1D, MODE, EPB><<<grid, EPB * TPE, 0, h->stream>>>(A, h->topo, tables<N1D>(h));
  prof_end(h);
  CU(h, cudaGetLastError());
  h->launches++;
  return 0;
}
template <int N1D>
int launch_update_fast(p2de_handle *h, const UpdateArgs &A) {
  constexpr int EPB = 16;
  unsigned grid = (unsigned)((h->K + EPB - 1) / EPB);
  prof_begin(h, 1);
  update_kernel_fast<N1D, EPB><<<grid, EPB * 16, 0, h->stream>>>(A, h->topo, tables<N1D>(h));
  prof_end(h);
  CU(h, cudaGetLastError());
  h->launches++;
  return 0;
}
template <int N1D>
int launch_update_n(p2de_handle *h, const UpdateArgs &A) {
  if (h->mode == MODE_SUBCELL && (h->fast || h->slim)) return launch_update_fast<N1D>(h, A);
  if (h->mode == MODE_SUBCELL) return launch_update_t<N1D, MODE_SUBCELL>(h, A);
  return launch_update_t<N1D, MODE_LOW>(h, A);
}
int launch_update(p2de_handle *h, const UpdateArgs &A) {
  switch (h->N1D) {
    case 2: return launch_update_n<2>(h, A);
    case 3: return launch_update_n<3>(h, A);
    case 4: return launch_update_n<4>(h, A);
    case 5: return launch_update_n<5>(h, A);
  }
  return fail(h, P2DE_ERR_UNSUPPORTED, "N=%d", h->cfg.N);
}

// Uout = a resW + b (Uin + dt r): the SSP stage combine (SSPRK33.jl:31-39) as a flat, fully coalesced pass.
// wcap > 0: r is stage 1's W = Uin + wcap rhsU (stage_subcell.cuh: KIND_S1), so Uin + dt rhsU = Uin + (dt / wcap) (W - Uin)
__global__ void __launch_bounds__(256)
axpy_update_kernel(double2 *__restrict__ Uout, const double2 *__restrict__ resW, const double2 *__restrict__ Uin,
                   const double2 *__restrict__ r, long long n2, double a, double b, const double *dt_dev, double dt_host, int use_dt_dev,
                   double wcap) {
  const double dt = use_dt_dev ? dt_read(dt_dev) : dt_host;
  const bool plain = a == 0.0 && b == 1.0;
  const double theta = wcap > 0.0 ? dt / wcap : 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const double2 u = Uin[i], q = r[i];
    double2 o;
    if (wcap > 0.0) { o.x = fma(theta, q.x - u.x, u.x); o.y = fma(theta, q.y - u.y, u.y); }
    else if (plain) { o.x = u.x + dt * q.x; o.y = u.y + dt * q.y; }
    else { const double2 w = resW[i]; o.x = a * w.x + b * (u.x + dt * q.x); o.y = a * w.y + b * (u.y + dt * q.y); }
    Uout[i] = o;
  }
}
int launch_axpy(p2de_handle *h, double *Uout, const double *resW, const double *Uin, const double *r, double a, double b,
                double dt_host, bool use_dt_dev, double wcap = 0.0) {
  const long long n2 = h->K * h->Nq * 2;
  const unsigned grid = (unsigned)std::min<long long>((n2 + 255) / 256, 148ll * 8 * 4);
  prof_begin(h, 1);
  axpy_update_kernel<<<grid, 256, 0, h->stream>>>(reinterpret_cast<double2 *>(Uout), reinterpret_cast<const double2 *>(resW),
                                                  reinterpret_cast<const double2 *>(Uin), reinterpret_cast<const double2 *>(r), n2, a, b,
                                                  reinterpret_cast<const double *>(h->dt_bits), dt_host, use_dt_dev ? 1 : 0, wcap);
  prof_end(h);
  CU(h, cudaGetLastError());
  h->launches++;
  return 0;
}

StageArgs stage_args(p2de_handle *h, const double *Uq, int nstage, double dt_host, bool use_dt_dev) {
  StageArgs A{};
  A.Uq = Uq;
  A.rhsL = h->rhsL; A.dF = h->dF; A.lpre = h->lpre; A.rhsU = h->rhsU;
  A.rpre = h->rpre; A.dFend = h->dFend;
  A.Lout = h->Lz ? h->Lz + h->K * (nstage - 1) : nullptr;
  A.rhsH_diag = h->rhsH_diag; A.rhsL_diag = h->rhsL_diag;
  A.dt_bits = h->dt_bits;
  A.dt_dev = reinterpret_cast<const double *>(h->dt_bits);
  A.dt_host = dt_host; A.use_dt_dev = use_dt_dev; A.nstage = nstage;
  A.gamma = h->cfg.gamma; A.ZEROTOL = h->cfg.ZEROTOL; A.POSTOL = h->cfg.POSTOL; A.zeta = h->cfg.zeta;
  A.CFL = h->cfg.CFL; A.Jq = h->Jq; A.blend = 1.0;
  A.hennemann = h->cfg.shockcapture == P2DE_SHOCKCAPTURE_HENNEMANN;
  A.entropy_bound = h->entropy_bound; A.N = h->cfg.N;
  A.hen_a = h->cfg.hennemann_a; A.hen_c = h->cfg.hennemann_c;
  A.VDM_inv = h->VDM_inv; A.smin_dev = reinterpret_cast<const double *>(h->smin_bits);
  A.roundtrip = h->cfg.lgl_projection_roundtrip;
  A.half_inv_gm1 = 1.0 / (2.0 * (h->cfg.gamma - 1.0));
  A.tab_dev = h->tab_dev;
  A.gauss = h->gauss ? 1 : 0; A.utf = h->utf;
  A.theta_local = (h->gauss && h->nodewise) ? h->theta_local_dev + (size_t)h->K * h->Nfp * (nstage - 1) : nullptr;
  A.vol_flux = h->cfg.vol_flux; A.surf_low = h->cfg.surf_flux_low; A.surf_high = h->cfg.surf_flux_high;
  A.tvd = h->tvd; A.rhsLpre = h->rhsLpre; A.cell_entropy = h->cell_entropy; A.bound_beta = h->cfg.bound_beta;
  A.fstar = h->fstar;
  A.dbg = h->dbg;
  return A;
}

int ensure_rhsU(p2de_handle *h) {
  if (!h->rhsU) return dev_alloc(h, &h->rhsU, (size_t)h->K * h->Nq * h->Nc);
  return 0;
}
int ensure_Llocal(p2de_handle *h) {
  if (!h->Llocal) {
    size_t n = (size_t)h->nLloc * h->K * h->Ns;
    if (int rc = dev_alloc(h, &h->Llocal, n)) return rc;
    CU(h, cudaMemsetAsync(h->Llocal, 0, n * sizeof(double), h->stream));
  }
  return 0;
}

template <int N1D> Tables1D<N1D> &tables1(p2de_handle *h);
template <> Tables1D<2> &tables1<2>(p2de_handle *h) { return h->u2; }
template <> Tables1D<3> &tables1<3>(p2de_handle *h) { return h->u3; }
template <> Tables1D<4> &tables1<4>(p2de_handle *h) { return h->u4; }
template <> Tables1D<5> &tables1<5>(p2de_handle *h) { return h->u5; }

template <int N1D>
int build_tables1d(p2de_handle *h, const p2de_operators *o) {
  constexpr int Nq = N1D, Nh = N1D + 2;
  Tables1D<N1D> &T = tables1<N1D>(h);
  for (int i = 0; i < Nh; ++i) for (int j = 0; j < Nh; ++j) T.Srh[i][j] = o->Srsh_db[0][i + (size_t)j * Nh];
  for (int i = 0; i < Nq; ++i) for (int j = 0; j < Nq; ++j) T.S0[i][j] = o->Srs0[0][i + (size_t)j * Nq];
  bool gather = true;
  for (int f = 0; f < 2; ++f) {
    T.Br[f] = o->Brs[0][f];
    T.fq2q[f] = (int)o->fq2q[f] - 1;
    if (T.fq2q[f] < 0 || T.fq2q[f] >= Nq) return fail(h, P2DE_ERR_ARG, "fq2q[%d] out of range", f);
    for (int j = 0; j < Nq; ++j) {
      T.Vf[f][j] = o->Vf[f + (size_t)j * 2];
      T.Vf_low[f][j] = o->Vf_low ? o->Vf_low[f + (size_t)j * 2] : (j == T.fq2q[f] ? 1.0 : 0.0);
      if (T.Vf[f][j] != (j == T.fq2q[f] ? 1.0 : 0.0)) gather = false;
    }
  }
  T.vf_is_gather = gather;
  for (int i = 0; i < Nq; ++i) {
    T.wq[i] = o->wq[i];
    for (int hh = 0; hh < Nh; ++hh) T.MinvVhT[i][hh] = o->MinvVhT[i + (size_t)hh * Nq];
    for (int f = 0; f < 2; ++f) T.MinvVfT[i][f] = o->MinvVfT[i + (size_t)f * Nq];
  }
  return 0;
}

template <int N1D>
int run_stage_1d_t(p2de_handle *h, const Args1D &A, const Upd1D &B, bool do_update) {
  unsigned grid = (unsigned)((h->K + 63) / 64);
  prof_begin(h, 0);
  stage1d_kernel<N1D><<<grid, 64, 0, h->stream>>>(A, tables1<N1D>(h));
  prof_end(h);
  CU(h, cudaGetLastError());
  h->launches++;
  if (do_update) {
    prof_begin(h, 1);
    update1d_kernel<N1D><<<grid, 64, 0, h->stream>>>(B, tables1<N1D>(h));
    prof_end(h);
    CU(h, cudaGetLastError());
    h->launches++;
  }
  return 0;
}

int run_stage_1d(p2de_handle *h, const double *Uin, int nstage, double dt_host, bool limiter_dt_dev, bool update_dt_dev,
                 double *Uout, const double *resW, double a, double b, bool want_outputs) {
  Args1D A{};
  A.Uq = Uin; A.rhsL = h->rhsL; A.dF = h->dF; A.lpre = h->lpre; A.rhsU = h->rhsU;
  A.Lout = h->Lz + h->K * (nstage - 1); A.rhsH_diag = h->rhsH_diag; A.rhsL_diag = h->rhsL_diag;
  A.dt_bits = h->dt_bits; A.dt_dev = reinterpret_cast<const double *>(h->dt_bits);
  A.dt_host = dt_host; A.use_dt_dev = limiter_dt_dev; A.nstage = nstage; A.K = h->K;
  A.mapP32 = h->mapP32; A.bcflag = h->bcflag; A.Ival = h->Ival;
  A.gamma = h->cfg.gamma; A.ZEROTOL = h->cfg.ZEROTOL; A.POSTOL = h->cfg.POSTOL; A.zeta = h->cfg.zeta; A.CFL = h->cfg.CFL;
  A.Jq = h->Jq; A.rxJ = h->rxJ1; A.blend = 1.0; A.mode = h->mode;
  A.vol_flux = h->cfg.vol_flux; A.surf_low = h->cfg.surf_flux_low; A.surf_high = h->cfg.surf_flux_high;
  A.roundtrip = h->cfg.lgl_projection_roundtrip;
  A.tvd = h->tvd; A.rhsLpre = h->rhsLpre; A.entropy_bound = h->entropy_bound; A.cell_entropy = h->cell_entropy;
  A.hennemann = h->cfg.shockcapture == P2DE_SHOCKCAPTURE_HENNEMANN; A.N = h->cfg.N;
  A.hen_a = h->cfg.hennemann_a; A.hen_c = h->cfg.hennemann_c; A.bound_beta = h->cfg.bound_beta;
  A.VDM_inv = h->VDM_inv; A.smin_dev = reinterpret_cast<const double *>(h->smin_bits);
  A.nodewise = h->nodewise ? 1 : 0; A.gauss = h->cfg.basis == P2DE_BASIS_GAUSS; A.eta = h->cfg.eta;
  A.theta_local = h->nodewise ? h->theta_local_dev + (size_t)h->K * 2 * (nstage - 1) : nullptr;
  A.theta = h->nodewise ? h->theta_dev + (size_t)h->K * (nstage - 1) : nullptr;
  if (h->tvd) {
    // TVD bounds (subcell.jl:86-110): rho + dt rhsL[1] of the stencil nodes across the element faces needs the neighbours'
    // finished low-order rhs: a MODE_LOW pre-pass of the same kernel writes it for all elements
    Args1D P = A;
    P.mode = MODE_LOW; P.rhsU = h->rhsLpre; P.nstage = 2;   // nstage != 1: no CFL reduction in the pre-pass
    P.rhsH_diag = nullptr; P.rhsL_diag = nullptr; P.tvd = 0; P.entropy_bound = 0; P.cell_entropy = 0; P.hennemann = 0;
    P.theta_local = nullptr; P.theta = nullptr;
    Upd1D none{};
    int rc = 0;
    switch (h->N1D) {
      case 2: rc = run_stage_1d_t<2>(h, P, none, false); break;
      case 3: rc = run_stage_1d_t<3>(h, P, none, false); break;
      case 4: rc = run_stage_1d_t<4>(h, P, none, false); break;
      case 5: rc = run_stage_1d_t<5>(h, P, none, false); break;
    }
    if (rc) return rc;
  }
  Upd1D B{};
  B.rhsL = h->rhsL; B.dF = h->dF; B.lpre = h->lpre; B.rhsU_in = h->rhsU;
  B.Llocal_out = (want_outputs && h->mode == MODE_SUBCELL) ? h->Llocal + (size_t)h->nLloc * h->K * (nstage - 1) : nullptr;
  B.rhsU_out = (want_outputs && h->mode == MODE_SUBCELL) ? h->rhsU : nullptr;
  B.Uq_in = Uin; B.resW = resW; B.Uq_out = Uout; B.a = a; B.b = b;
  B.dt_dev = reinterpret_cast<const double *>(h->dt_bits); B.dt_host = dt_host; B.use_dt_dev = update_dt_dev;
  B.mode = h->mode; B.K = h->K; B.Jq = h->Jq;
  const bool upd = h->mode == MODE_SUBCELL || Uout;
  switch (h->N1D) {
    case 2: return run_stage_1d_t<2>(h, A, B, upd);
    case 3: return run_stage_1d_t<3>(h, A, B, upd);
    case 4: return run_stage_1d_t<4>(h, A, B, upd);
    case 5: return run_stage_1d_t<5>(h, A, B, upd);
  }
  return fail(h, P2DE_ERR_UNSUPPORTED, "N=%d", h->cfg.N);
}

// 1D: device state, gather tables (mapP is always explicit here: K is small), BC flags
int create_1d(p2de_handle *h, const p2de_operators *ops, const p2de_geometry *geom, const p2de_bcdata *bc) {
  const int N1D = h->N1D, Nq = h->Nq;
  const long long K = h->K;
  if (geom->uniform) { h->Jq = geom->J_const; h->Jcons = geom->J_const; h->rxJ1 = geom->GJ_const[0]; }
  else {
    if (!geom->Jq || !geom->GJh[0]) return fail(h, P2DE_ERR_ARG, "geometry arrays missing");
    h->Jq = geom->Jq[0]; h->Jcons = geom->J ? geom->J[0] : geom->Jq[0]; h->rxJ1 = geom->GJh[0][0];
    for (size_t i = 0; i < (size_t)Nq * K; ++i) if (geom->Jq[i] != h->Jq) return fail(h, P2DE_ERR_UNSUPPORTED, "non-uniform Jq");
    for (size_t i = 0; i < (size_t)(Nq + 2) * K; ++i) if (geom->GJh[0][i] != h->rxJ1) return fail(h, P2DE_ERR_UNSUPPORTED, "non-uniform rxJ");
  }
  int rc = 0;
  switch (N1D) {
    case 2: rc = build_tables1d<2>(h, ops); break;
    case 3: rc = build_tables1d<3>(h, ops); break;
    case 4: rc = build_tables1d<4>(h, ops); break;
    case 5: rc = build_tables1d<5>(h, ops); break;
  }
  if (rc) return rc;
  h->wq.assign(ops->wq, ops->wq + Nq);
  if (2 * K >= (1ll << 31)) return fail(h, P2DE_ERR_UNSUPPORTED, "K too large for the 1D path");
  std::vector<int> m32((size_t)2 * K), fl((size_t)2 * K, 0);
  for (long long k = 0; k < K; ++k)
    for (int f = 0; f < 2; ++f) {
      long long v;
      if (bc->mapP) v = bc->mapP[k * 2 + f] - 1;
      else {
        long long kn = f == 0 ? k - 1 : k + 1;
        bool out = kn < 0 || kn >= K;
        if (out && bc->periodic_x) { kn = (kn + K) % K; out = false; }
        v = out ? k * 2 + f : kn * 2 + (1 - f);
      }
      if (v < 0 || v >= 2 * K) return fail(h, P2DE_ERR_ARG, "mapP out of range");
      m32[k * 2 + f] = (int)v;
    }
  for (long long i = 0; i < bc->nI; ++i) { long long idx = bc->mapI[i] - 1; if (idx < 0 || idx >= 2 * K) return fail(h, P2DE_ERR_ARG, "mapI out of range"); fl[idx] = (int)(i + 1); }
  for (long long i = 0; i < bc->nO; ++i) { long long idx = bc->mapO[i] - 1; if (idx < 0 || idx >= 2 * K) return fail(h, P2DE_ERR_ARG, "mapO out of range"); fl[idx] = -1; }
  if ((rc = dev_alloc(h, &h->mapP32, m32.size())) || (rc = dev_alloc(h, &h->bcflag, fl.size()))) return rc;
  CU(h, cudaMemcpy(h->mapP32, m32.data(), m32.size() * sizeof(int), cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->bcflag, fl.data(), fl.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (bc->nI > 0) {
    if ((rc = dev_alloc(h, &h->Ival, (size_t)bc->nI * 3))) return rc;
    CU(h, cudaMemcpy(h->Ival, bc->Ival, (size_t)bc->nI * 3 * sizeof(double), cudaMemcpyHostToDevice));
  }
  h->topo.K = K; h->topo.Kx = (int)K; h->topo.Ky = 1;
  const size_t nU = (size_t)K * Nq * 3;
  if ((rc = dev_alloc(h, &h->U[0], nU)) || (rc = dev_alloc(h, &h->U[1], nU)) || (rc = dev_alloc(h, &h->rhsU, nU)) ||
      (rc = dev_alloc(h, &h->Lz, (size_t)K * h->Ns)) || (rc = dev_alloc(h, &h->dt_bits, 2)) ||
      (rc = dev_alloc(h, &h->partial, 1024 + (size_t)Nq)))
    return rc;
  CU(h, cudaMemset(h->Lz, 0, (size_t)K * h->Ns * sizeof(double)));
  CU(h, cudaMemcpy(h->partial + 1024, h->wq.data(), Nq * sizeof(double), cudaMemcpyHostToDevice));
  if (h->mode == MODE_SUBCELL)
    if ((rc = dev_alloc(h, &h->rhsL, nU)) || (rc = dev_alloc(h, &h->dF, (size_t)K * (Nq + 1) * 3)) || (rc = dev_alloc(h, &h->lpre, (size_t)K * (Nq + 1)))) return rc;
  if (h->tvd && (rc = dev_alloc(h, &h->rhsLpre, nU))) return rc;
  if ((rc = dev_alloc(h, &h->smin_bits, 1))) return rc;
  CU(h, cudaMemset(h->smin_bits, 0, sizeof(unsigned long long)));   // s_modified_min starts at 0.0 (State.jl:180)
  if (ops->VDM_inv) {
    if ((rc = dev_alloc(h, &h->VDM_inv, (size_t)Nq * Nq))) return rc;
    CU(h, cudaMemcpy(h->VDM_inv, ops->VDM_inv, (size_t)Nq * Nq * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (h->cfg.keep_diagnostics) {
    if ((rc = dev_alloc(h, &h->rhsH_diag, nU)) || (rc = dev_alloc(h, &h->rhsL_diag, nU))) return rc;
    CU(h, cudaMemset(h->rhsH_diag, 0, nU * sizeof(double))); CU(h, cudaMemset(h->rhsL_diag, 0, nU * sizeof(double)));
  }
  if (h->nodewise) {   // theta_local [Ns][K][2], theta [Ns][K]: zeros until a stage writes its slot (init.jl:22-23)
    if (h->cfg.basis == P2DE_BASIS_GAUSS && !ops->Vf_low) return fail(h, P2DE_ERR_ARG, "NodewiseScaledExtrapolation needs ops.Vf_low");
    const size_t nth = (size_t)K * 2 * h->Ns, nt = (size_t)K * h->Ns;
    if ((rc = dev_alloc(h, &h->theta_local_dev, nth)) || (rc = dev_alloc(h, &h->theta_dev, nt))) return rc;
    CU(h, cudaMemset(h->theta_local_dev, 0, nth * sizeof(double)));
    CU(h, cudaMemset(h->theta_dev, 0, nt * sizeof(double)));
  }
  h->fast = false;
  return 0;
}

template <int N1D>
int launch_project_n(p2de_handle *h, const double *Uin, int nstage) {
  constexpr int EPB = 8;
  ProjArgs P{};
  P.Uq = Uin; P.utf = h->utf;
  P.theta_local = h->nodewise ? h->theta_local_dev + (size_t)h->K * h->Nfp * (nstage - 1) : nullptr;
  P.theta = h->nodewise ? h->theta_dev + (size_t)h->K * (nstage - 1) : nullptr;
  P.gamma = h->cfg.gamma; P.POSTOL = h->cfg.POSTOL; P.zeta = h->cfg.zeta; P.eta = h->cfg.eta;
  P.nodewise = h->nodewise ? 1 : 0;
  unsigned grid = (unsigned)((h->K + EPB - 1) / EPB);
  prof_begin(h, 2);
  gauss_project_kernel<N1D, EPB><<<grid, EPB * 2 * N1D, 0, h->stream>>>(P, h->topo, tables<N1D>(h));
  prof_end(h);
  CU(h, cudaGetLastError());
  h->launches++;
  return 0;
}
int launch_project(p2de_handle *h, const double *Uin, int nstage) {
  switch (h->N1D) {
    case 2: return launch_project_n<2>(h, Uin, nstage);
    case 3: return launch_project_n<3>(h, Uin, nstage);
    case 4: return launch_project_n<4>(h, Uin, nstage);
    case 5: return launch_project_n<5>(h, Uin, nstage);
  }
  return fail(h, P2DE_ERR_UNSUPPORTED, "N=%d", h->cfg.N);
}

// one stage: stage_kernel + update_kernel.  `Uin` is the stage input; if `Uout` != nullptr the
// SSP combine Uout = a*resW + b*(Uin + dt*rhsU) is fused into the update kernel.
int run_stage(p2de_handle *h, const double *Uin, int nstage, double t, double dt_host, bool limiter_dt_dev,
              bool update_dt_dev, double *Uout, const double *resW, double a, double b, bool want_outputs,
              double *Uadd = nullptr, bool combine = true) {
  if (nstage == 1) {
    double cap = std::fmin(h->cfg.CFL * h->cfg.dt0, h->cfg.T - t);   // low_order_graph_viscosity.jl:230
    set_dt_kernel<<<1, 1, 0, h->stream>>>(h->dt_bits, cap, 1);
    CU(h, cudaGetLastError());
    h->launches++;
  }
  if (h->dim == 1) {
    if (h->entropy_bound && nstage == 1 && t == h->cfg.t0) {   // subcell.jl:32-34: global minimum of the initial condition
      set_dt_kernel<<<1, 1, 0, h->stream>>>(h->smin_bits, INFINITY, 0);
      smin1d_kernel<<<64, 256, 0, h->stream>>>(Uin, h->K * h->Nq, h->cfg.gamma, h->smin_bits);
      CU(h, cudaGetLastError());
      h->launches += 2;
    }
    return run_stage_1d(h, Uin, nstage, dt_host, limiter_dt_dev, update_dt_dev, Uout, resW, a, b, want_outputs);
  }
  if (h->rhsH_diag && h->mode == MODE_SUBCELL)
    CU(h, cudaMemsetAsync(h->rhsH_diag, 0, (size_t)h->K * h->Nq * h->Nc * sizeof(double), h->stream));
  if (h->entropy_bound && nstage == 1 && t == h->cfg.t0) {   // subcell.jl:32-34: global minimum of the initial condition
    set_dt_kernel<<<1, 1, 0, h->stream>>>(h->smin_bits, INFINITY, 0);
    smin_kernel<<<1024, 256, 0, h->stream>>>(Uin, h->K * h->Nq, h->cfg.gamma, h->smin_bits);
    CU(h, cudaGetLastError());
    h->launches += 2;
    if (h->comm) NC(h, nccl_api(nullptr)->AllReduce(h->smin_bits, h->smin_bits, 1, ncclDouble, ncclMin, h->comm, h->stream));
  }
  // E1: face-state halo (the boundary element rows of Uq) from the stripes below / above; with a deferred combine
  // (stage input = Uin + dt Uadd) the halo rows of Uin are still those of the previous stage and Uadd's travel instead
  if (int rc = exchange_rows(h, Uadd ? Uadd : const_cast<double *>(Uin), (size_t)h->cfg.Kx * h->Nq * 4)) return rc;
  if (h->gauss) {   // entropy projection (+ NodewiseScaledExtrapolation) to the face nodes; its halo rows travel like E1
    if (int rc = launch_project(h, Uin, nstage)) return rc;
    if (int rc = exchange_rows(h, h->utf, (size_t)h->cfg.Kx * h->Nfp * 4)) return rc;
  }
  StageArgs A = stage_args(h, Uin, nstage, dt_host, limiter_dt_dev);
  if (h->tvd) {
    // TVD bounds (subcell.jl:119-141) need rho + dt rhsL[1] of the stencil nodes ACROSS element faces, i.e. the
    // neighbours' finished low-order rhs: a low-order pre-pass of the same kernel family writes it for all elements
    StageArgs P = A;
    P.rhsU = h->rhsLpre; P.nstage = 2;   // nstage != 1: no CFL reduction in the pre-pass
    P.rhsH_diag = nullptr; P.rhsL_diag = nullptr; P.Lout = nullptr; P.tvd = 0; P.cell_entropy = 0; P.entropy_bound = 0; P.hennemann = 0;
    if (int rc = launch_stage(h, P, true)) return rc;
    if (int rc = exchange_rows(h, h->rhsLpre, (size_t)h->cfg.Kx * h->Nq * 4)) return rc;
  }
  // stages 2/3 of the FAST subcell path: the limiter's dt is the step's dt, so the stage kernel can
  // already form the SSP combine of the un-corrected rhs and the update kernel only adds corrections
  const bool fuse = h->fast && h->mode == MODE_SUBCELL && Uout && !want_outputs && nstage > 1 &&
                    limiter_dt_dev == update_dt_dev;
  // ... and when the output buffer is not the input buffer it writes the new state itself; the
  // no second kernel is needed then (sym_free below)
  // (Uout == resW is fine: a node's resW is read by the one thread that then writes that node)
  const bool direct = fuse && Uout != Uin;
  if (fuse) { A.fuse = 1; A.fuse_a = a; A.fuse_b = b; A.fuse_resW = resW; }
  if (direct) A.rpre = Uout;
  A.defer_add = Uadd;
  // FAST subcell path: f_bar_H - f_bar_L is exactly zero on interior element faces, so the interface coefficients are 1
  // on both sides and symmetrize_limiting_parameters! (subcell.jl:418-456) is the identity; what is left after the stage
  // kernel is at most the SSP combine (stage 1, where dt is only known once the kernel has finished everywhere)
  const bool sym_free = h->fast && h->mode == MODE_SUBCELL && !want_outputs && (direct || !fuse);
  // ... and the stage kernel's coefficients are final: when L_local is kept (State.jl:21; allocated by keep_diagnostics or
  // by the first p2de_rhs) each stage writes them straight into its own slot L_local[:, :, :, nstage], as SSP33! leaves them
  if (sym_free && h->Llocal) A.lpre = h->Llocal + (size_t)h->nLloc * h->K * (nstage - 1);
  // stage 1 of a step: nobody but the stage-1 combine reads what the kernel leaves in rpre, and dt is not known before the
  // whole grid has finished, so the kernel writes W = U + cap rhsU (cap = dt_host, the dt its limiter uses) and the combine
  // below / the next stage's kernel takes U + (dt / cap) (W - U)  (stage_subcell.cuh: KIND_S1)
  const bool wform = sym_free && nstage == 1 && !fuse && Uout && a == 0.0 && b == 1.0 && dt_host > 0.0;
  A.wform = wform ? 1 : 0;
  A.inv_cap = dt_host > 0.0 ? 1.0 / dt_host : 0.0;
  if (nstage == 1) h->rpre_w = wform;
  if (Uadd && !h->rpre_w) return fail(h, P2DE_ERR_STATE, "deferred stage-1 combine without a W-form stage 1");
  if (int rc = launch_stage(h, A)) return rc;
  if (h->comm) {
    if (nstage == 1 && h->mode != MODE_HIGH)   // global CFL dt (low_order_graph_viscosity.jl:242): min over all stripes
      NC(h, nccl_api(nullptr)->AllReduce(h->dt_bits, h->dt_bits, 2, ncclDouble, ncclMin, h->comm, h->stream));
    // E2: un-symmetrised interface coefficients of the neighbouring stripes' boundary rows (FAST path: the interface
    // coefficients are 1 on both sides of every interior face, so only the L_local output needs them)
    if (h->mode == MODE_SUBCELL && !sym_free)
      if (int rc = exchange_rows(h, h->lpre, (size_t)h->cfg.Kx * 2 * h->N1D * (h->N1D + 1))) return rc;
  }
  UpdateArgs B{};
  B.rhsL = h->rhsL; B.dF = h->dF; B.lpre = h->lpre; B.rhsU_in = h->rhsU;
  B.rpre = h->rpre; B.dFend = h->dFend;
  B.Llocal_out = (want_outputs && h->mode == MODE_SUBCELL) ? h->Llocal + (size_t)h->nLloc * h->K * (nstage - 1) : nullptr;
  B.rhsU_out = (want_outputs && h->mode == MODE_SUBCELL) ? h->rhsU : nullptr;
  B.Uq_in = Uin; B.resW = resW; B.Uq_out = Uout; B.a = a; B.b = b;
  B.dt_dev = reinterpret_cast<const double *>(h->dt_bits); B.dt_host = dt_host; B.use_dt_dev = update_dt_dev;
  B.Jq = h->Jq; B.rotated = h->fast ? 1 : 0; B.pre_updated = fuse ? 1 : 0;
  B.fstar = h->fstar; B.gamma = h->cfg.gamma;
  if (direct) return 0;   // the stage kernel wrote the new state; nothing to symmetrise on the FAST path
  if (sym_free) {
    if (Uout && combine) return launch_axpy(h, Uout, resW, Uin, h->rpre, a, b, dt_host, update_dt_dev, wform ? dt_host : 0.0);   // pure SSP combine
    return 0;             // combine deferred into the next stage's kernel (StageArgs.defer_add)
  }
  if (h->mode == MODE_SUBCELL || Uout) return launch_update(h, B);
  return 0;
}

// ---- peer-to-peer halo rows ---------------------------------------------------------------------------------------------
// "rows of exchange #seq have landed": written into the neighbour's flag word once the copies in front of it are done
__global__ void p2p_signal_kernel(volatile unsigned int *flag_a, volatile unsigned int *flag_b, unsigned int seq) {
  __threadfence_system();
  if (flag_a) *flag_a = seq;
  if (flag_b) *flag_b = seq;
  __threadfence_system();
}
// fallback of cuStreamWaitValue32: one thread spins until both flags have reached seq
__global__ void p2p_wait_kernel(volatile unsigned int *flag_a, volatile unsigned int *flag_b, unsigned int seq) {
  if (flag_a) while ((int)(*flag_a - seq) < 0) __nanosleep(200);
  if (flag_b) while ((int)(*flag_b - seq) < 0) __nanosleep(200);
  __threadfence_system();
}
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);   // cuStreamWaitValue32(CUstream, CUdeviceptr, value, flags)
StreamWaitValue32Fn stream_wait_value32() {
  static StreamWaitValue32Fn fn = [] {
    const char *no = getenv("P2DE_NO_STREAM_MEMOPS");
    if (no && atoi(no)) return (StreamWaitValue32Fn) nullptr;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<StreamWaitValue32Fn>(f);
  }();
  return fn;
}

// An IPC handle names the driver's whole underlying allocation (small cudaMalloc blocks are carved out of larger ones), and
// cudaIpcOpenMemHandle returns that allocation's base: the exporter sends its pointer's offset from the base along.
struct PeerInfo { cudaIpcMemHandle_t buf[3]; cudaIpcMemHandle_t flags; long long off[4]; long long Ky; long long pad; };
typedef int (*MemGetAddressRangeFn)(unsigned long long *, size_t *, unsigned long long);   // cuMemGetAddressRange(CUdeviceptr *, size_t *, CUdeviceptr)
bool alloc_base_offset(const void *p, long long *off) {
  static MemGetAddressRangeFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<MemGetAddressRangeFn>(f);
  }();
  if (!fn) return false;
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return false;
  *off = (long long)((unsigned long long)(uintptr_t)p - base);
  return true;
}

// Map the neighbouring stripes' state buffers and flag words into this process (CUDA IPC; one process per GPU).  Any
// failure leaves h->p2p false and the NCCL send/recv exchange in place.
int p2p_setup(p2de_handle *h) {
  // EXPERIMENTAL, off unless P2DE_P2P=1: bitwise-correct in tests/multigpu_check.py and tests/multigpu_async.py on 2 GPUs,
  // but the first step hung inside bench.py's process on the same box (unresolved at the end of round 2), so the NCCL
  // send/recv exchange stays the default.
  const char *yes = getenv("P2DE_P2P");
  if (!(yes && atoi(yes)) || !h->fast || h->mode != MODE_SUBCELL || !h->rpre) return 0;
  const NcclApi *n = nccl_api(nullptr);
  const size_t rowU = (size_t)h->cfg.Kx * h->Nq * 4;
  // (an allocation of its own underlying block: IPC handles have the granularity of the driver's blocks, small allocations
  //  share one and their handle names the whole block)
  if (int rc = dev_alloc(h, &h->flags, (size_t)(4u << 20) / sizeof(unsigned int))) return rc;
  CU(h, cudaMemset(h->flags, 0, 2 * sizeof(unsigned int)));
  PeerInfo mine{};
  double *bufs[3] = {h->U[0], h->U[1], h->rpre};
  bool ok = true;
  for (int i = 0; i < 3; ++i)   // (the buffer starts one halo row before its first owned entry)
    ok = ok && cudaIpcGetMemHandle(&mine.buf[i], bufs[i] - rowU) == cudaSuccess && alloc_base_offset(bufs[i] - rowU, &mine.off[i]);
  ok = ok && cudaIpcGetMemHandle(&mine.flags, h->flags) == cudaSuccess && alloc_base_offset(h->flags, &mine.off[3]);
  mine.Ky = h->cfg.Ky;
  cudaGetLastError();
  // every rank must take the same decision: gather (ok, info) from all ranks
  static_assert(sizeof(PeerInfo) % 8 == 0, "gathered as doubles");
  const size_t nd = sizeof(PeerInfo) / 8 + 1;
  std::vector<double> sendb(nd, 0.0), recvb(nd * h->nranks, 0.0);
  std::memcpy(sendb.data(), &mine, sizeof(PeerInfo));
  sendb[nd - 1] = ok ? 1.0 : 0.0;
  double *dsend = nullptr, *drecv = nullptr;
  if (int rc = dev_alloc(h, &dsend, nd)) return rc;
  if (int rc = dev_alloc(h, &drecv, nd * h->nranks)) return rc;
  CU(h, cudaMemcpy(dsend, sendb.data(), nd * 8, cudaMemcpyHostToDevice));
  NC(h, n->AllGather(dsend, drecv, nd, ncclDouble, h->comm, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(recvb.data(), drecv, nd * 8 * h->nranks, cudaMemcpyDeviceToHost));
  for (int r = 0; r < h->nranks; ++r) ok = ok && recvb[(size_t)r * nd + nd - 1] == 1.0;
  if (!ok) return 0;
  auto open_peer = [&](int r, double *out[3], unsigned int **fl, long long *Ky) -> bool {
    PeerInfo pi;
    std::memcpy(&pi, recvb.data() + (size_t)r * nd, sizeof(PeerInfo));
    for (int i = 0; i < 3; ++i) {
      void *q = nullptr;
      if (cudaIpcOpenMemHandle(&q, pi.buf[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
      h->ipc_opened.push_back(q);
      out[i] = reinterpret_cast<double *>(static_cast<char *>(q) + pi.off[i]) + rowU;
    }
    void *q = nullptr;
    if (cudaIpcOpenMemHandle(&q, pi.flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
    h->ipc_opened.push_back(q);
    *fl = reinterpret_cast<unsigned int *>(static_cast<char *>(q) + pi.off[3]);
    *Ky = pi.Ky;
    return true;
  };
  long long Ky_lo = 0, Ky_hi = 0;
  bool opened = true;
  if (h->has_lo) opened = opened && open_peer(h->rank_lo, h->peer_lo, &h->peer_lo_flags, &Ky_lo);
  if (opened && h->has_hi) {
    if (h->has_lo && h->rank_hi == h->rank_lo) {   // two ranks, periodic: one neighbour on both sides, one mapping
      for (int i = 0; i < 3; ++i) h->peer_hi[i] = h->peer_lo[i];
      h->peer_hi_flags = h->peer_lo_flags; Ky_hi = Ky_lo;
    } else opened = opened && open_peer(h->rank_hi, h->peer_hi, &h->peer_hi_flags, &Ky_hi);
  }
  cudaGetLastError();
  // the lower neighbour's upper halo row sits behind ITS owned rows
  if (opened && h->has_lo) for (int i = 0; i < 3; ++i) h->peer_lo[i] += (size_t)Ky_lo * rowU;   // now: the neighbour's upper halo row
  if (opened && h->has_hi) for (int i = 0; i < 3; ++i) h->peer_hi[i] -= rowU;                      // now: the neighbour's lower halo row
  // (again a collective decision: a rank that could not map its neighbours keeps everybody on NCCL)
  double flag = opened ? 1.0 : 0.0;
  CU(h, cudaMemcpy(dsend, &flag, 8, cudaMemcpyHostToDevice));
  NC(h, n->AllReduce(dsend, dsend, 1, ncclDouble, ncclMin, h->comm, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(&flag, dsend, 8, cudaMemcpyDeviceToHost));
  h->p2p = flag == 1.0;
  return 0;
}

// boundary rows of `buf` (index bi) into the neighbours' halo rows, then the flags; all on the communication stream
int p2p_push_rows(p2de_handle *h, const double *buf, int bi) {
  const size_t rowU = (size_t)h->cfg.Kx * h->Nq * 4;
  ++h->seq;
  if (h->has_hi) CU(h, cudaMemcpyAsync(h->peer_hi[bi], buf + (size_t)(h->cfg.Ky - 1) * rowU, rowU * sizeof(double), cudaMemcpyDeviceToDevice, h->comm_stream));
  if (h->has_lo) CU(h, cudaMemcpyAsync(h->peer_lo[bi], buf, rowU * sizeof(double), cudaMemcpyDeviceToDevice, h->comm_stream));
  // I am the upper neighbour's LOWER neighbour (its flag word 0) and the lower neighbour's UPPER neighbour (its flag word 1)
  p2p_signal_kernel<<<1, 1, 0, h->comm_stream>>>(h->has_hi ? h->peer_hi_flags + 0 : nullptr, h->has_lo ? h->peer_lo_flags + 1 : nullptr, h->seq);
  CU(h, cudaGetLastError());
  h->launches++;
  h->wait_seq = h->seq;
  return 0;
}
// the compute stream waits until both neighbours' rows of exchange #wait_seq have landed in this stripe's halo rows
int p2p_wait_rows(p2de_handle *h) {
  if (!h->wait_seq) return 0;
  const unsigned int seq = h->wait_seq;
  h->wait_seq = 0;
  if (StreamWaitValue32Fn wv = stream_wait_value32()) {
    bool ok = true;
    if (h->has_lo) ok = ok && wv(h->stream, (unsigned long long)(uintptr_t)(h->flags + 0), seq, 0x1 /* CU_STREAM_WAIT_VALUE_GEQ */) == 0;
    if (ok && h->has_hi) ok = ok && wv(h->stream, (unsigned long long)(uintptr_t)(h->flags + 1), seq, 0x1) == 0;
    if (ok) return 0;
  }
  p2p_wait_kernel<<<1, 1, 0, h->stream>>>(h->has_lo ? h->flags + 0 : nullptr, h->has_hi ? h->flags + 1 : nullptr, seq);
  CU(h, cudaGetLastError());
  h->launches++;
  return 0;
}

// One SSP-RK3 step of the subcell family on a y-stripe with the halo exchange hidden behind the interior rows.
// Per stage: the two boundary element rows run first (a 2-row launch of the same kernel), their OUTPUT rows travel to
// the neighbouring stripes on the communication stream (NCCL send/recv) while the remaining rows run on the compute
// stream; the next stage's boundary launch waits for the exchange.  What travels is what the next stage reads beyond
// the cut: stage 1's W, stage 2's U2, stage 3's U^{n+1} (whose halo stays valid for stages 1 and 2 of the next step).
// The CFL dt all-reduce (low_order_graph_viscosity.jl:242) is issued on the same communication stream once both
// stage-1 launches are done; it is the only exposed collective of the step.
int run_step_overlapped(p2de_handle *h, double t, double cap) {
  double *Ua = h->U[h->cur], *Ub = h->U[1 - h->cur];
  const size_t rowU = (size_t)h->cfg.Kx * h->Nq * 4;
  const NcclApi *n = nccl_api(nullptr);
  auto buf_index = [&](const double *b) { return b == h->U[0] ? 0 : b == h->U[1] ? 1 : 2; };
  // rows of `buf` to the neighbours, on the communication stream, after whatever the compute stream has enqueued so far
  auto send_rows = [&](double *buf, cudaEvent_t after) -> int {
    CU(h, cudaEventRecord(after, h->stream));
    CU(h, cudaStreamWaitEvent(h->comm_stream, after, 0));
    if (h->p2p) { if (int rc = p2p_push_rows(h, buf, buf_index(buf))) return rc; }
    else if (int rc = exchange_rows(h, buf, rowU, h->comm_stream, true)) return rc;
    CU(h, cudaEventRecord(h->ev_halo, h->comm_stream));
    h->halo_pending = true;
    return 0;
  };
  // the compute stream may read the halo rows: NCCL: the exchange (send and receive) is done; P2P: the neighbours' flags
  auto wait_rows = [&]() -> int {
    if (h->p2p) return p2p_wait_rows(h);
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    return 0;
  };
  if (!h->halo_current) {   // first step after set_state: the halo of U^n has not travelled yet
    if (int rc = send_rows(Ua, h->ev_all)) return rc;
    h->halo_current = true;
  }
  set_dt_kernel<<<1, 1, 0, h->stream>>>(h->dt_bits, cap, 1);
  CU(h, cudaGetLastError());
  h->launches++;
  struct StageIO { const double *in; double *out; const double *resW; double a, b; const double *add; };
  const StageIO io[3] = {{Ua, h->rpre, nullptr, 0.0, 1.0, nullptr},
                         {Ua, Ub, Ua, 3.0 / 4.0, 1.0 / 4.0, h->rpre},
                         {Ub, Ua, Ua, 1.0 / 3.0, 2.0 / 3.0, nullptr}};
  for (int st = 1; st <= 3; ++st) {
    const StageIO &s = io[st - 1];
    StageArgs A = stage_args(h, s.in, st, cap, st > 1);
    A.rpre = s.out;
    if (st == 1) { A.wform = 1; }
    else { A.fuse = 1; A.fuse_a = s.a; A.fuse_b = s.b; A.fuse_resW = s.resW; }
    A.inv_cap = 1.0 / cap;
    A.defer_add = s.add;
    if (h->Llocal) A.lpre = h->Llocal + (size_t)h->nLloc * h->K * (st - 1);
    if (st == 1) h->rpre_w = true;
    // boundary rows: need the halo of this stage's input
    if (int rc = wait_rows()) return rc;
    StageArgs B = A;
    B.row0 = 0; B.row_stride = h->cfg.Ky - 1; B.nrows = 2;
    if (int rc = launch_stage(h, B)) return rc;
    if (int rc = send_rows(s.out, h->ev_rows)) return rc;
    // interior rows
    StageArgs I = A;
    I.row0 = 1; I.row_stride = 1; I.nrows = h->cfg.Ky - 2;
    if (int rc = launch_stage(h, I)) return rc;
    if (st == 1) {   // global CFL dt: min over all stripes, after both stage-1 launches
      CU(h, cudaEventRecord(h->ev_all, h->stream));
      CU(h, cudaStreamWaitEvent(h->comm_stream, h->ev_all, 0));
      NC(h, n->AllReduce(h->dt_bits, h->dt_bits, 2, ncclDouble, ncclMin, h->comm, h->comm_stream));
      CU(h, cudaEventRecord(h->ev_dt, h->comm_stream));
      CU(h, cudaStreamWaitEvent(h->stream, h->ev_dt, 0));
    }
  }
  return 0;
}

}  // namespace

extern "C" {

const char *p2de_last_error(const p2de_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int32_t p2de_create(const p2de_config *cfg, const p2de_operators *ops, const p2de_geometry *geom,
                    const p2de_bcdata *bc, p2de_handle **out) {
  if (!cfg || !ops || !geom || !bc || !out) return fail(nullptr, P2DE_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != P2DE_ABI_VERSION) return fail(nullptr, P2DE_ERR_ARG, "abi_version %d != %d", cfg->abi_version, P2DE_ABI_VERSION);
  if (cfg->dim != 1 && cfg->dim != 2) return fail(nullptr, P2DE_ERR_ARG, "dim=%d", cfg->dim);
  if (cfg->N < 1 || cfg->N > 4) return fail(nullptr, P2DE_ERR_UNSUPPORTED, "N=%d outside 1..4", cfg->N);
  const int N1D = cfg->N + 1;
  const bool d1 = cfg->dim == 1;
  if (d1 ? (cfg->Nq != N1D || cfg->Nfp != 2 || cfg->Nh != N1D + 2 || cfg->Np != N1D)
         : (cfg->Nq != N1D * N1D || cfg->Nfp != 4 * N1D || cfg->Nh != cfg->Nq + cfg->Nfp || cfg->Np != cfg->Nq))
    return fail(nullptr, P2DE_ERR_ARG, "sizes inconsistent with a degree-%d %s", cfg->N, d1 ? "line" : "quad");
  if (cfg->K <= 0) return fail(nullptr, P2DE_ERR_ARG, "K must be positive");
  if (cfg->shockcapture != P2DE_SHOCKCAPTURE_NONE && !ops->VDM_inv)
    return fail(nullptr, P2DE_ERR_UNSUPPORTED, "HennemannShockCapture needs ops.VDM_inv");
  int mode;
  if (cfg->rhs_type == P2DE_RHS_LOW_ORDER_POSITIVITY) mode = MODE_LOW;
  else if (cfg->rhs_type == P2DE_RHS_FLUX_DIFF) mode = MODE_HIGH;
  else if (cfg->rhs_type == P2DE_RHS_LIMITED_DG) {
    if (cfg->limiter == P2DE_LIMITER_ZHANGSHU) mode = MODE_ZHANGSHU;
    else if (cfg->limiter == P2DE_LIMITER_SUBCELL) {
      if (cfg->bound < P2DE_BOUND_POSITIVITY || cfg->bound > P2DE_BOUND_TVD_RELAXED_CELL_ENTROPY) return fail(nullptr, P2DE_ERR_ARG, "bound %d", cfg->bound);
      // the smoothness indicator runs for every bound but PositivityBound (shock_capture.jl:4-12) and needs inv(VDM)
      if (cfg->bound != P2DE_BOUND_POSITIVITY && cfg->bound != P2DE_BOUND_TVD && !ops->VDM_inv)
        return fail(nullptr, P2DE_ERR_ARG, "bound %d needs ops.VDM_inv", cfg->bound);
      mode = MODE_SUBCELL;
    } else return fail(nullptr, P2DE_ERR_UNSUPPORTED, "LimitedDG needs ZhangShuLimiter or SubcellLimiter");
  } else return fail(nullptr, P2DE_ERR_ARG, "rhs_type %d", cfg->rhs_type);
  if (cfg->surf_flux_low != P2DE_SURFFLUX_LF_NODAL && cfg->surf_flux_low != P2DE_SURFFLUX_LF_PROJECTED)
    return fail(nullptr, P2DE_ERR_ARG, "surf_flux_low %d", cfg->surf_flux_low);
  if (cfg->surf_flux_high != P2DE_SURFFLUX_LF_PROJECTED && cfg->surf_flux_high != P2DE_SURFFLUX_CHANDRASHEKAR_PROJECTED)
    return fail(nullptr, P2DE_ERR_ARG, "surf_flux_high %d", cfg->surf_flux_high);

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, P2DE_ERR_CUDA, "no CUDA device (%s): libp2de_b200 has no CPU fallback", cudaGetErrorString(e));
  struct RestoreDevice { int prev = -1; RestoreDevice() { cudaGetDevice(&prev); } ~RestoreDevice() { if (prev >= 0) cudaSetDevice(prev); } } restore_device;
  p2de_handle *h = new p2de_handle();
  h->cfg = *cfg;
  {
    bool low_ok = cfg->surf_flux_low == P2DE_SURFFLUX_LF_NODAL;
    bool high_ok = cfg->surf_flux_high == P2DE_SURFFLUX_LF_PROJECTED && cfg->vol_flux == P2DE_VOLFLUX_CHANDRASHEKAR &&
                   !cfg->lgl_projection_roundtrip;
    h->fast = mode == MODE_LOW ? low_ok : mode == MODE_HIGH ? high_ok : (low_ok && high_ok);
  }
  if (mode == MODE_SUBCELL) {
    const int b = cfg->bound;
    h->entropy_bound = (b == P2DE_BOUND_POS_MIN_ENTROPY || b == P2DE_BOUND_TVD_MIN_ENTROPY) ? 1
                     : (b == P2DE_BOUND_POS_RELAXED_MIN_ENTROPY || b == P2DE_BOUND_TVD_RELAXED_MIN_ENTROPY) ? 2 : 0;
    h->tvd = b >= P2DE_BOUND_TVD ? 1 : 0;
    h->cell_entropy = cell_entropy_of(b);
  }
  if (h->entropy_bound || h->tvd || h->cell_entropy || cfg->shockcapture != P2DE_SHOCKCAPTURE_NONE) h->fast = false;   // generic kernel has these features
  h->gauss = !d1 && cfg->basis == P2DE_BASIS_GAUSS;
  h->nodewise = cfg->proj_limiter == P2DE_PROJLIM_NODEWISE;
  if (h->gauss) h->fast = false;
  h->slim = h->gauss && mode == MODE_SUBCELL && !cfg->lgl_projection_roundtrip && cfg->vol_flux == P2DE_VOLFLUX_CHANDRASHEKAR &&
            cfg->surf_flux_low == P2DE_SURFFLUX_LF_PROJECTED && cfg->surf_flux_high == P2DE_SURFFLUX_LF_PROJECTED &&
            cfg->shockcapture == P2DE_SHOCKCAPTURE_NONE && cfg->bound == P2DE_BOUND_POSITIVITY;
  h->N1D = N1D; h->Nq = cfg->Nq; h->Nfp = cfg->Nfp; h->Nc = d1 ? 3 : 4; h->Nd = d1 ? 1 : 2; h->K = cfg->K; h->mode = mode;
  h->dim = cfg->dim;
  h->nLloc = d1 ? 2 * N1D : 2 * N1D * (N1D + 1);   // State.jl:21: zeros(Nq + N1D, Nd, K, Ns); 1D uses the first Nq+1
  auto bail = [&](int rc) { g_create_error = h->err; p2de_destroy(h); return rc; };
  if (cfg->device >= 0) { if (cudaSetDevice(cfg->device) != cudaSuccess) return bail(fail(h, P2DE_ERR_CUDA, "cudaSetDevice(%d) failed", cfg->device)); }
  cudaGetDevice(&h->device);
  if (d1) {
    int rc1 = create_1d(h, ops, geom, bc);
    if (rc1) return bail(rc1);
    *out = h;
    return P2DE_OK;
  }

  // geometry: uniform meshes only (the reference builds nothing else, init.jl:137)
  double GJ[4];
  if (geom->uniform) {
    h->Jq = geom->J_const; h->Jcons = geom->J_const;
    for (int a = 0; a < 4; ++a) GJ[a] = geom->GJ_const[a];
  } else {
    if (!geom->Jq || !geom->GJh[0] || !geom->GJh[1] || !geom->GJh[2] || !geom->GJh[3]) return bail(fail(h, P2DE_ERR_ARG, "geometry arrays missing"));
    h->Jq = geom->Jq[0]; h->Jcons = geom->J ? geom->J[0] : geom->Jq[0];
    for (int a = 0; a < 4; ++a) GJ[a] = geom->GJh[a][0];
    for (size_t i = 0; i < (size_t)cfg->Nq * cfg->K; ++i)
      if (geom->Jq[i] != h->Jq || (geom->J && geom->J[i] != h->Jcons)) return bail(fail(h, P2DE_ERR_UNSUPPORTED, "non-uniform Jq: only uniform meshes are supported"));
    for (int a = 0; a < 4; ++a)
      for (size_t i = 0; i < (size_t)cfg->Nh * cfg->K; ++i)
        if (geom->GJh[a][i] != GJ[a]) return bail(fail(h, P2DE_ERR_UNSUPPORTED, "non-uniform geometric factors"));
  }
  int rc = 0;
  switch (N1D) {
    case 2: rc = build_tables<2>(h, ops, GJ); break;
    case 3: rc = build_tables<3>(h, ops, GJ); break;
    case 4: rc = build_tables<4>(h, ops, GJ); break;
    case 5: rc = build_tables<5>(h, ops, GJ); break;
  }
  if (rc) return bail(rc);
  h->wq.assign(ops->wq, ops->wq + cfg->Nq);
  if ((rc = setup_topology(h, bc))) return bail(rc);

  const size_t nU = (size_t)h->K * h->Nq * 4;
  const size_t rowU = (size_t)cfg->Kx * h->Nq * 4, rowL = (size_t)cfg->Kx * 2 * N1D * (N1D + 1);
  if ((rc = dev_alloc_halo(h, &h->U[0], nU, rowU)) || (rc = dev_alloc_halo(h, &h->U[1], nU, rowU))) return bail(rc);
  if (mode == MODE_SUBCELL) {
    if ((rc = dev_alloc_halo(h, &h->lpre, (size_t)h->K * 2 * N1D * (N1D + 1), rowL))) return bail(rc);
    if (h->fast) {
      // (no dFend buffer: with exact-zero f_bar_H - f_bar_L on interior element faces the interface symmetrisation
      //  is the identity on this path, stage_fast.cuh "End faces")
      if ((rc = dev_alloc_halo(h, &h->rpre, nU, rowU))) return bail(rc);
      const char *nd = getenv("P2DE_NO_DIRECT");   // testing aid: keep the dense update kernel after every stage
      h->direct = !(nd && atoi(nd));
      const char *nf = getenv("P2DE_NO_DEFER");    // testing / A-B aid: materialise U1 with the axpy kernel
      h->defer = h->direct && !(nf && atoi(nf));
    } else if (h->slim) {
      if ((rc = dev_alloc(h, &h->rpre, nU)) || (rc = dev_alloc(h, &h->dFend, (size_t)h->K * h->Nfp * 4))) return bail(rc);
    } else if ((rc = dev_alloc(h, &h->rhsL, nU)) || (rc = dev_alloc(h, &h->dF, (size_t)h->K * 2 * N1D * (N1D + 1) * 4)))
      return bail(rc);
    if (h->tvd && (rc = dev_alloc_halo(h, &h->rhsLpre, nU, rowU))) return bail(rc);
    if (h->cell_entropy && h->gauss && (rc = dev_alloc(h, &h->fstar, (size_t)h->K * h->Nfp * 8))) return bail(rc);
  } else {
    if ((rc = ensure_rhsU(h))) return bail(rc);
  }
  if ((rc = dev_alloc(h, &h->Lz, (size_t)h->K * h->Ns))) return bail(rc);
  if (cudaMemset(h->Lz, 0, (size_t)h->K * h->Ns * sizeof(double)) != cudaSuccess) return bail(fail(h, P2DE_ERR_CUDA, "memset"));
  if (cfg->keep_diagnostics) {
    if (mode == MODE_SUBCELL && (rc = ensure_Llocal(h))) return bail(rc);   // L_local[:, :, :, 1:3] current after every step
    if ((rc = dev_alloc(h, &h->rhsH_diag, nU)) || (rc = dev_alloc(h, &h->rhsL_diag, nU))) return bail(rc);
    cudaMemset(h->rhsH_diag, 0, nU * sizeof(double)); cudaMemset(h->rhsL_diag, 0, nU * sizeof(double));
  }
  if (h->gauss || h->nodewise) {
    const size_t nth = (size_t)h->K * h->Nfp * h->Ns, nt = (size_t)h->K * h->Ns;
    if ((rc = dev_alloc(h, &h->theta_local_dev, nth)) || (rc = dev_alloc(h, &h->theta_dev, nt))) return bail(rc);
    std::vector<double> ones(nth, 1.0);   // Lobatto + Nodewise: theta_local = 1 (filter.jl:18-20), theta stays 0 as in the reference
    if (cudaMemcpy(h->theta_local_dev, ones.data(), nth * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(h, P2DE_ERR_CUDA, "memcpy theta"));
    cudaMemset(h->theta_dev, 0, nt * sizeof(double));
    if (!h->nodewise) cudaMemset(h->theta_local_dev, 0, nth * sizeof(double));   // NoEntropyProjectionLimiter never writes theta_local
  }
  if (h->gauss && (rc = dev_alloc_halo(h, &h->utf, (size_t)h->K * h->Nfp * 4, (size_t)cfg->Kx * h->Nfp * 4))) return bail(rc);
  if ((rc = dev_alloc(h, &h->dt_bits, 2)) || (rc = dev_alloc(h, &h->smin_bits, 1)) || (rc = dev_alloc(h, &h->partial, 1024 + (size_t)h->Nq))) return bail(rc);
  cudaMemset(h->smin_bits, 0, sizeof(unsigned long long));   // s_modified_min starts at 0.0 (State.jl:180)
  if (ops->VDM_inv) {
    if ((rc = dev_alloc(h, &h->VDM_inv, (size_t)h->Nq * h->Nq))) return bail(rc);
    if (cudaMemcpy(h->VDM_inv, ops->VDM_inv, (size_t)h->Nq * h->Nq * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(h, P2DE_ERR_CUDA, "memcpy VDM_inv"));
  }
  if (cudaMemcpy(h->partial + 1024, h->wq.data(), h->Nq * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(h, P2DE_ERR_CUDA, "memcpy wq"));
  {
    const void *src = N1D == 2 ? (const void *)&h->t2 : N1D == 3 ? (const void *)&h->t3 : N1D == 4 ? (const void *)&h->t4 : (const void *)&h->t5;
    const size_t nb = N1D == 2 ? sizeof(h->t2) : N1D == 3 ? sizeof(h->t3) : N1D == 4 ? sizeof(h->t4) : sizeof(h->t5);
    if ((rc = dev_alloc(h, &h->tab_dev, (nb + 7) / 8 + 2))) return bail(rc);
    if (cudaMemcpy(h->tab_dev, src, nb, cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(h, P2DE_ERR_CUDA, "memcpy tables"));
  }
  *out = h;
  return P2DE_OK;
}

int32_t p2de_destroy(p2de_handle *h) {
  if (!h) return P2DE_OK;
  DEV(h);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  if (h->comm) { if (const NcclApi *n = nccl_api(nullptr)) n->CommDestroy(h->comm); }
  for (void *q : h->ipc_opened) cudaIpcCloseMemHandle(q);
  for (cudaEvent_t e : {h->ev_rows, h->ev_all, h->ev_halo, h->ev_dt}) if (e) cudaEventDestroy(e);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  prof_clear(h);
  for (void *p : h->owned) cudaFree(p);
  delete h;
  return P2DE_OK;
}

int32_t p2de_set_stream(p2de_handle *h, void *cuda_stream) {
  if (!h) return P2DE_ERR_ARG;
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  return P2DE_OK;
}

int32_t p2de_synchronize(p2de_handle *h) {
  if (!h) return P2DE_ERR_ARG;
  DEV(h);
  CU(h, cudaStreamSynchronize(h->stream));
  if (h->comm_stream) CU(h, cudaStreamSynchronize(h->comm_stream));
  return P2DE_OK;
}

int32_t p2de_set_state_async(p2de_handle *h, const double *Uq_host) {
  if (!h || !Uq_host) return fail(h, P2DE_ERR_ARG, "null argument");
  DEV(h);
  if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }   // the last exchange still reads the old boundary rows
  CU(h, cudaMemcpyAsync(h->U[h->cur], Uq_host, (size_t)h->K * h->Nq * h->Nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->have_state = true;
  h->halo_current = false;
  return P2DE_OK;
}
int32_t p2de_set_state(p2de_handle *h, const double *Uq_host) {
  if (int rc = p2de_set_state_async(h, Uq_host)) return rc;
  return p2de_synchronize(h);
}
int32_t p2de_get_state_async(p2de_handle *h, double *Uq_host) {
  if (!h || !Uq_host) return fail(h, P2DE_ERR_ARG, "null argument");
  DEV(h);
  if (!h->have_state) return fail(h, P2DE_ERR_STATE, "get_state before set_state");
  CU(h, cudaMemcpyAsync(Uq_host, h->U[h->cur], (size_t)h->K * h->Nq * h->Nc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return P2DE_OK;
}
int32_t p2de_get_state(p2de_handle *h, double *Uq_host) {
  if (int rc = p2de_get_state_async(h, Uq_host)) return rc;
  return p2de_synchronize(h);
}

int32_t p2de_rhs(p2de_handle *h, double t, double dt, int32_t nstage, double *dt_out) {
  if (!h) return P2DE_ERR_ARG;
  DEV(h);
  if (!h->have_state) return fail(h, P2DE_ERR_STATE, "rhs before set_state");
  if (nstage < 1 || nstage > h->Ns) return fail(h, P2DE_ERR_ARG, "nstage %d outside 1..%d", nstage, h->Ns);
  if (int rc = ensure_rhsU(h)) return rc;
  if (h->mode == MODE_SUBCELL) if (int rc = ensure_Llocal(h)) return rc;
  if (int rc = run_stage(h, h->U[h->cur], nstage, t, dt, false, false, nullptr, nullptr, 0, 0, true)) return rc;
  double dtr[2] = {dt, 1.0};
  if (nstage == 1 && h->mode != MODE_HIGH) {
    CU(h, cudaMemcpyAsync(dtr, h->dt_bits, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(h, cudaStreamSynchronize(h->stream));
  if (dt_out) *dt_out = dtr[1] == 1.0 ? dtr[0] : std::numeric_limits<double>::quiet_NaN();   // dt_publish: a NaN candidate was seen
  return P2DE_OK;
}

int32_t p2de_ssp33_step_async(p2de_handle *h, double t) {
  if (!h) return P2DE_ERR_ARG;
  DEV(h);
  if (!h->have_state) return fail(h, P2DE_ERR_STATE, "ssp33_step before set_state");
  // L_local[:, :, :, 1:3] of the step (SSPRK33.jl:31-39 leaves all three stages' coefficients behind): kept current once the
  // buffer exists.  On the FAST path that costs nothing (run_stage: the stage kernel writes its coefficients into the stage's
  // slot, same 4-launch schedule); on the generic path the update kernel that runs anyway writes them.
  const bool fast_direct = h->direct && h->fast && h->mode == MODE_SUBCELL && h->dim == 2;
  const bool outs = h->Llocal != nullptr && !fast_direct;
  double *Ua = h->U[h->cur], *Ub = h->U[1 - h->cur];
  double cap = std::fmin(h->cfg.CFL * h->cfg.dt0, h->cfg.T - t);   // SSPRK33.jl:30
  const bool has_cfl = h->mode != MODE_HIGH;                          // FluxDiffRHS never changes dt (rhs.jl:38)
  // (stripes of at least 64 element rows: below that the two extra launches per stage cost more than the exchange they hide)
  if (fast_direct && h->comm && h->overlap && h->defer && cap > 0.0 && h->cfg.Ky >= (h->p2p ? 3 : h->overlap_min_rows) && h->cfg.Ky <= 65535 &&
      h->cfg.Kx % (h->N1D == 4 ? LaunchSub<4>::EPB : h->N1D == 5 ? LaunchSub<5>::EPB : LaunchSub<2>::EPB) == 0)
    return run_step_overlapped(h, t, cap);
  h->halo_current = false;   // the schedules below exchange in front of every stage and leave the new state's halo stale
  if (fast_direct) {
    // Direct schedule: stages 2 and 3 write their result from the stage kernel (run_stage: `direct`), which
    // needs an output buffer other than the stage input: U1 -> Ub (dense update, dt only known after the
    // stage-1 kernel), U2 -> the rpre buffer (free once the stage-1 update has consumed it), U^{n+1} -> over
    // U^n, which stage 3 reads only as its own resW, node by node, by the thread that then writes the node.
    if (h->defer && cap > 0.0) {
      // ... and the stage-1 combine U1 = U^n + dt rhsU is not materialised at all: stage 1 leaves rhsU in the rpre
      // buffer and the stage-2 kernel forms U1 while loading (same bytes as reading U1 and resW = U^n): 4 launches
      if (int rc = run_stage(h, Ua, 1, t, cap, false, true, Ub, Ua, 0.0, 1.0, false, nullptr, false)) return rc;
      if (int rc = run_stage(h, Ua, 2, t, cap, true, true, Ub, Ua, 3.0 / 4.0, 1.0 / 4.0, false, h->rpre)) return rc;
      if (int rc = run_stage(h, Ub, 3, t, cap, true, true, Ua, Ua, 1.0 / 3.0, 2.0 / 3.0, false)) return rc;
      return P2DE_OK;
    }
    if (int rc = run_stage(h, Ua, 1, t, cap, false, true, Ub, Ua, 0.0, 1.0, false)) return rc;
    if (int rc = run_stage(h, Ub, 2, t, cap, true, true, h->rpre, Ua, 3.0 / 4.0, 1.0 / 4.0, false)) return rc;
    if (int rc = run_stage(h, h->rpre, 3, t, cap, true, true, Ua, Ua, 1.0 / 3.0, 2.0 / 3.0, false)) return rc;
    return P2DE_OK;
  }
  // stage 1: limiter sees the cap (rhs.jl:46,52), the combine sees the CFL-limited dt
  if (int rc = run_stage(h, Ua, 1, t, cap, false, has_cfl, Ub, Ua, 0.0, 1.0, outs)) return rc;
  if (int rc = run_stage(h, Ub, 2, t, cap, has_cfl, has_cfl, Ub, Ua, 3.0 / 4.0, 1.0 / 4.0, outs)) return rc;
  if (int rc = run_stage(h, Ub, 3, t, cap, has_cfl, has_cfl, Ub, Ua, 1.0 / 3.0, 2.0 / 3.0, outs)) return rc;
  h->cur = 1 - h->cur;
  return P2DE_OK;
}

int32_t p2de_last_dt(p2de_handle *h, double *dt_out) {
  if (!h || !dt_out) return P2DE_ERR_ARG;
  DEV(h);
  double v[2];
  CU(h, cudaMemcpyAsync(v, h->dt_bits, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  *dt_out = v[1] == 1.0 ? v[0] : std::numeric_limits<double>::quiet_NaN();   // dt_publish: a NaN candidate was seen
  return P2DE_OK;
}

int32_t p2de_ssp33_step(p2de_handle *h, double t, double *dt_out) {
  if (int rc = p2de_ssp33_step_async(h, t)) return rc;
  double dt = 0;
  if (h->mode == MODE_HIGH) {
    CU(h, cudaStreamSynchronize(h->stream));
    dt = std::fmin(h->cfg.CFL * h->cfg.dt0, h->cfg.T - t);
  } else if (int rc = p2de_last_dt(h, &dt)) return rc;
  h->last_dt = dt;
  if (dt_out) *dt_out = dt;
  return P2DE_OK;
}

int32_t p2de_ssp33_run(p2de_handle *h, double *t_inout, int64_t max_steps, int64_t *steps_out, double *dthist) {
  if (!h || !t_inout) return P2DE_ERR_ARG;
  double t = *t_inout;
  int64_t n = 0;
  while (t < h->cfg.T && n < max_steps) {   // SSPRK33.jl:28
    double dt;
    if (int rc = p2de_ssp33_step(h, t, &dt)) return rc;
    t += dt;
    if (dthist) dthist[n] = dt;
    ++n;
    // DataHistory (SSPRK33.jl:41-55: `i` starts at 1 and is incremented before the test): a copy of Uq every
    // output_interval steps and at the final time, kept in the device-side ring
    if (h->snap_slots > 0 && ((h->snap_interval > 0 && (n + 1) % h->snap_interval == 0) || std::fabs(t - h->cfg.T) < 1e-10)) {
      const size_t nU = (size_t)h->K * h->Nq * h->Nc;
      const int slot = (int)(h->snap_count % h->snap_slots);
      CU(h, cudaMemcpyAsync(h->snap_buf + (size_t)slot * nU, h->U[h->cur], nU * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      h->snap_t[slot] = t; h->snap_step[slot] = n + 1;
      ++h->snap_count;
    }
  }
  *t_inout = t;
  if (steps_out) *steps_out = n;
  return P2DE_OK;
}

int32_t p2de_snapshot_ring(p2de_handle *h, int32_t slots, int64_t output_interval) {
  if (!h || slots < 0) return fail(h, P2DE_ERR_ARG, "slots must be >= 0");
  DEV(h);
  CU(h, cudaStreamSynchronize(h->stream));
  const size_t nU = (size_t)h->K * h->Nq * h->Nc;
  if (slots != h->snap_slots) {
    if (h->snap_buf) { cudaFree(h->snap_buf); h->owned.erase(std::remove(h->owned.begin(), h->owned.end(), (void *)h->snap_buf), h->owned.end()); h->snap_buf = nullptr; }
    if (slots > 0) if (int rc = dev_alloc(h, &h->snap_buf, nU * (size_t)slots)) return rc;
    h->snap_slots = slots;
  }
  h->snap_interval = output_interval; h->snap_count = 0;
  h->snap_t.assign((size_t)slots, 0.0); h->snap_step.assign((size_t)slots, 0);
  return P2DE_OK;
}
int64_t p2de_snapshot_count(const p2de_handle *h) { return h ? h->snap_count : 0; }
int32_t p2de_snapshot_get(p2de_handle *h, int64_t index, double *Uq_host, double *t_out, int64_t *step_out) {
  if (!h) return P2DE_ERR_ARG;
  DEV(h);
  const int64_t first = h->snap_count > h->snap_slots ? h->snap_count - h->snap_slots : 0;
  if (index < first || index >= h->snap_count) return fail(h, P2DE_ERR_ARG, "snapshot %lld is not in the ring (%lld..%lld)", (long long)index, (long long)first, (long long)h->snap_count - 1);
  const int slot = (int)(index % h->snap_slots);
  const size_t nU = (size_t)h->K * h->Nq * h->Nc;
  if (Uq_host) {
    CU(h, cudaMemcpyAsync(Uq_host, h->snap_buf + (size_t)slot * nU, nU * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  if (t_out) *t_out = h->snap_t[slot];
  if (step_out) *step_out = h->snap_step[slot];
  return P2DE_OK;
}

int32_t p2de_calculate_error(p2de_handle *h, const double *exact_host, double *out) {
  if (!h || !exact_host || !out) return fail(h, P2DE_ERR_ARG, "null argument");
  if (!h->have_state) return fail(h, P2DE_ERR_STATE, "calculate_error before set_state");
  DEV(h);
  const long long n_nodes = h->K * h->Nq, chunk = 1ll << 20;
  const int blocks = 256, Nc = h->Nc;
  if (!h->err_stage) if (int rc = dev_alloc(h, &h->err_stage, (size_t)chunk * 4 + (size_t)blocks * 24)) return rc;
  double *partial = h->err_stage + (size_t)chunk * 4;
  std::vector<double> part((size_t)blocks * 24), acc(24, 0.0);
  for (long long n0 = 0; n0 < n_nodes; n0 += chunk) {
    const long long nn = std::min(chunk, n_nodes - n0);
    CU(h, cudaMemcpyAsync(h->err_stage, exact_host + n0 * Nc, (size_t)nn * Nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    error_norm_kernel<<<blocks, 256, 0, h->stream>>>(h->U[h->cur], h->err_stage, h->partial + 1024, h->Nq, Nc, n0, nn, h->Jq, partial);
    CU(h, cudaGetLastError());
    h->launches++;
    CU(h, cudaMemcpyAsync(part.data(), partial, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (int b = 0; b < blocks; ++b)
      for (int q = 0; q < 6; ++q)
        for (int c = 0; c < Nc; ++c) {
          const double v = part[((size_t)b * 6 + q) * 4 + c];
          acc[q * 4 + c] = (q == 2 || q == 5) ? std::fmax(acc[q * 4 + c], v) : acc[q * 4 + c] + v;
        }
  }
  for (int q = 0; q < 6; ++q) for (int c = 0; c < Nc; ++c) out[q * Nc + c] = acc[q * 4 + c];
  return P2DE_OK;
}

int32_t p2de_get_field(p2de_handle *h, int32_t field, double *dst, int64_t n) {
  if (!h || !dst) return fail(h, P2DE_ERR_ARG, "null argument");
  DEV(h);
  const int64_t nU = h->K * h->Nq * h->Nc;
  const double *src = nullptr;
  int64_t cnt = 0;
  switch (field) {
    case P2DE_FIELD_UQ: src = h->U[h->cur]; cnt = nU; break;
    case P2DE_FIELD_RESW: src = h->U[1 - h->cur]; cnt = nU; break;
    case P2DE_FIELD_RHSU: src = h->rhsU; cnt = nU; break;
    case P2DE_FIELD_RHSH: src = h->rhsH_diag; cnt = nU; break;
    case P2DE_FIELD_RHSL: src = h->rhsL_diag; cnt = nU; break;
    case P2DE_FIELD_L: src = h->Lz; cnt = h->K * h->Ns; break;
    case P2DE_FIELD_L_LOCAL: src = h->Llocal; cnt = (int64_t)h->nLloc * h->K * h->Ns; break;
    case P2DE_FIELD_THETA: cnt = h->K * h->Ns; break;          // NoEntropyProjectionLimiter: never written (zeros)
    case P2DE_FIELD_THETA_LOCAL: cnt = (int64_t)h->Nfp * h->K * h->Ns; break;
    default: return fail(h, P2DE_ERR_ARG, "unknown field %d", field);
  }
  if (n < cnt) return fail(h, P2DE_ERR_ARG, "destination too small: %lld < %lld", (long long)n, (long long)cnt);
  if (field == P2DE_FIELD_THETA && h->theta_dev) src = h->theta_dev;
  else if (field == P2DE_FIELD_THETA_LOCAL && h->theta_local_dev) src = h->theta_local_dev;
  else if (field == P2DE_FIELD_THETA || field == P2DE_FIELD_THETA_LOCAL) { std::memset(dst, 0, cnt * sizeof(double)); return P2DE_OK; }
  if (!src) return fail(h, P2DE_ERR_STATE, "field %d is not kept (keep_diagnostics / call p2de_rhs first)", field);
  CU(h, cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return P2DE_OK;
}

int32_t p2de_reduce(p2de_handle *h, int32_t what, double *out) {
  if (!h || !out) return P2DE_ERR_ARG;
  DEV(h);
  if (what < 0 || what > 2) return fail(h, P2DE_ERR_ARG, "unknown reduction %d", what);
  const int blocks = 1024;
  if (h->dim == 1) reduce1d_kernel<<<blocks, 256, 0, h->stream>>>(h->U[h->cur], h->partial + 1024, h->Nq, h->K * h->Nq, h->Jcons, what, h->partial);
  else reduce_kernel<<<blocks, 256, 0, h->stream>>>(h->U[h->cur], h->partial + 1024, h->Nq, h->K * h->Nq, h->Jcons, what, h->partial);
  CU(h, cudaGetLastError());
  h->launches++;
  std::vector<double> part(blocks);
  CU(h, cudaMemcpyAsync(part.data(), h->partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  double r = what == P2DE_REDUCE_CONSERVATION ? 0.0 : std::numeric_limits<double>::infinity();
  for (double v : part) r = what == P2DE_REDUCE_CONSERVATION ? r + v : std::fmin(r, v);
  *out = r;
  return P2DE_OK;
}

int32_t p2de_profile(p2de_handle *h, int32_t enable) {
  if (!h) return P2DE_ERR_ARG;
  DEV(h);
  CU(h, cudaStreamSynchronize(h->stream));
  prof_clear(h);
  h->profiling = enable != 0;
  return P2DE_OK;
}
int32_t p2de_profile_get(p2de_handle *h, int32_t kernel_id, double *total_ms, int64_t *launches) {
  if (!h || !total_ms || !launches) return P2DE_ERR_ARG;
  DEV(h);
  CU(h, cudaStreamSynchronize(h->stream));
  double tot = 0; int64_t n = 0;
  if (kernel_id == 100) {   // device time BETWEEN consecutive profiled launches (exchanges, collectives, launch bubbles)
    for (size_t i = 0; i + 1 < h->prof.size(); ++i) {
      float ms = 0;
      CU(h, cudaEventElapsedTime(&ms, h->prof[i].b, h->prof[i + 1].a));
      tot += ms; ++n;
    }
    *total_ms = tot; *launches = n;
    return P2DE_OK;
  }
  for (auto &r : h->prof) {
    if (r.kid != kernel_id) continue;
    float ms = 0;
    CU(h, cudaEventElapsedTime(&ms, r.a, r.b));
    tot += ms; ++n;
  }
  *total_ms = tot; *launches = n;
  return P2DE_OK;
}

int32_t p2de_debug_counters(p2de_handle *h, int32_t enable, uint64_t out[P2DE_DBG_COUNT]) {
  if (!h) return P2DE_ERR_ARG;
  DEV(h);
  static_assert((int)P2DE_DBG_COUNT == (int)DBG_COUNT && (int)P2DE_DBG_LIMITER_SLOW == (int)DBG_LIMITER_SLOW, "counter ids");
  CU(h, cudaStreamSynchronize(h->stream));
  if (out) {
    if (h->dbg_buf) CU(h, cudaMemcpy(out, h->dbg_buf, DBG_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    else std::memset(out, 0, DBG_COUNT * sizeof(uint64_t));
  }
  if (enable && !h->dbg) {
    if (!h->dbg_buf) if (int rc = dev_alloc(h, &h->dbg_buf, (size_t)DBG_COUNT)) return rc;
    CU(h, cudaMemset(h->dbg_buf, 0, DBG_COUNT * sizeof(uint64_t)));
    h->dbg = h->dbg_buf;
  } else if (!enable) h->dbg = nullptr;
  return P2DE_OK;
}

int64_t p2de_kernel_launch_count(const p2de_handle *h) { return h ? h->launches : 0; }
void *p2de_device_state_ptr(p2de_handle *h) { return h ? h->U[h->cur] : nullptr; }

int32_t p2de_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return fail(nullptr, P2DE_ERR_ARG, "null argument");
  std::string why;
  const NcclApi *n = nccl_api(&why);
  if (!n) return fail(nullptr, P2DE_ERR_NCCL, "%s", why.c_str());
  ncclUniqueId id;
  ncclResult_t r = n->GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, P2DE_ERR_NCCL, "ncclGetUniqueId: %s", n->GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id_out, &id, 128);
  return P2DE_OK;
}

int32_t p2de_comm_init(p2de_handle *h, int32_t rank, int32_t nranks, const uint8_t unique_id[128]) {
  if (!h || !unique_id) return fail(h, P2DE_ERR_ARG, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, P2DE_ERR_ARG, "rank %d of %d", rank, nranks);
  if (h->topo.mapP32 || h->dim == 1) return fail(h, P2DE_ERR_UNSUPPORTED, "multi-GPU needs the structured 2D mesh path (y-stripes)");
  if (h->comm) return fail(h, P2DE_ERR_STATE, "communicator already initialised");
  if (h->fstar && nranks > 1)   // the interface bisection reads the neighbour's fstar_H / fstar_L, which have no halo rows
    return fail(h, P2DE_ERR_UNSUPPORTED, "cell-entropy bounds on Gauss nodes are single-GPU in this build");
  h->rank = rank; h->nranks = nranks;
  if (nranks == 1) return P2DE_OK;
  std::string why;
  const NcclApi *n = nccl_api(&why);
  if (!n) return fail(h, P2DE_ERR_NCCL, "%s", why.c_str());
  ncclUniqueId id;
  std::memcpy(&id, unique_id, 128);
  DEV(h);
  ncclComm_t comm = nullptr;
  ncclResult_t r = n->CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) return fail(h, P2DE_ERR_NCCL, "ncclCommInitRank: %s", n->GetErrorString(r));
  h->comm = comm;
  {
    const char *no = getenv("P2DE_NO_OVERLAP");   // testing / A-B aid: blocking exchange in front of every stage kernel
    h->overlap = !(no && atoi(no));
    if (const char *mr = getenv("P2DE_OVERLAP_MIN_ROWS")) h->overlap_min_rows = std::max(3, atoi(mr));
    int lo = 0, hi = 0;
    CU(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(h, cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
    for (cudaEvent_t *e : {&h->ev_rows, &h->ev_all, &h->ev_halo, &h->ev_dt}) CU(h, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  // y-stripes: rank r owns element rows [r*Ky, (r+1)*Ky) of the global mesh; the stripe below /
  // above is rank -/+ 1, wrapping around when the global mesh is periodic in y.
  const bool per_y = h->topo.periodic_y != 0;
  h->has_lo = rank > 0 || per_y;
  h->has_hi = rank < nranks - 1 || per_y;
  h->rank_lo = (rank - 1 + nranks) % nranks;
  h->rank_hi = (rank + 1) % nranks;
  h->topo.ghost_lo = h->has_lo; h->topo.ghost_hi = h->has_hi;
  if (h->overlap) if (int rc = p2p_setup(h)) return rc;
  return P2DE_OK;
}

}  // extern "C"
