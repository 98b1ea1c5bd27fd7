// flow.cu -- the handle behind include/taxila_gpu.h: device storage, halo exchange, step loop,
// host <-> device layout conversion, and the extern "C" entry points.
//
// There is no CPU fallback anywhere in this file: every entry point that computes launches CUDA
// kernels on the handle's device and fails with a non-zero code if it cannot.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../include/taxila_gpu.h"
#include "flow.h"
#include "setup_kernels.cuh"
#include "bc_kernels.cuh"
#include "lag_schedule.h"

using namespace txg;

// PETSc error codes (petscerror.h) -- the reference's convention, lbm_error.F90:30-45
enum {
  TXG_ERR_MEM = 55,
  TXG_ERR_SUP = 56,
  TXG_ERR_ORDER = 58,
  TXG_ERR_ARG_WRONG = 62,
  TXG_ERR_ARG_OUTOFRANGE = 63,
  TXG_ERR_ARG_NULL = 85,
  TXG_ERR_LIB = 76
};

// ------------------------------------------------------------------ NCCL, bound at run time
// The single-GPU path needs no NCCL at all; multi-GPU binds libnccl.so.2 on first use so that a host
// program that already loaded NCCL (torch.distributed, an MPI+NCCL Fortran driver) shares it.
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat64 = 8 };
enum { ncclMax = 2 };
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
#define TXG_SYM(field, name)                                  \
  field = (decltype(field))dlsym(lib, name);                  \
  if (!field) {                                               \
    err = std::string("libnccl lacks symbol ") + name;        \
    return false;                                             \
  }
    TXG_SYM(GetUniqueId, "ncclGetUniqueId")
    TXG_SYM(CommInitRank, "ncclCommInitRank")
    TXG_SYM(CommDestroy, "ncclCommDestroy")
    TXG_SYM(Send, "ncclSend")
    TXG_SYM(Recv, "ncclRecv")
    TXG_SYM(AllReduce, "ncclAllReduce")
    TXG_SYM(GroupStart, "ncclGroupStart")
    TXG_SYM(GroupEnd, "ncclGroupEnd")
    TXG_SYM(GetErrorString, "ncclGetErrorString")
#undef TXG_SYM
    return true;
  }
};
NcclApi g_nccl;
std::string g_create_error;
}  // namespace

// ------------------------------------------------------------------ the handle
struct KernelTimer {
  const char *name = "";  // a string literal of this file: the pointers txg_kernel_times hands out stay valid
  double ms = 0.;
  int64_t launches = 0;
  int64_t per_graph[2] = {0, 0};  // launches one replay of the two-step graph of parity p adds
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct txg_flow {
  txg_config cfg;
  Grid g;
  Phys p;
  KernelSet ks;
  int device = 0;
  int num_sms = 0;
  int S = 0, Q = 0, D = 0, R = 1;
  cudaStream_t s_main = nullptr, s_comm = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_step0 = nullptr, ev_step1 = nullptr;
  double *f[2] = {nullptr, nullptr};  // f[cur] holds the populations fi(m,n,X); collide pushes into f[cur^1]
  int cur = 0;
  double *rho = nullptr;       // stencil field (rho, or psi with an EOS), R ghost planes
  double *rho_true = nullptr;  // true density (only allocated with an EOS; else == rho)
  double *u0 = nullptr;        // initial per-component velocity [S][D][nnodes], optional
  double *gw = nullptr;
  uint8_t *cls = nullptr;
  uint32_t *nbmask = nullptr, *ffmask = nullptr;
  // sparse storage (kernels.cuh): node -> position map, position -> node list, per-position masks and
  // wall records; plane_off[zz] = position of the first fluid node of extended plane zz (NZl+2Rz+1 entries)
  uint32_t *P = nullptr, *list = nullptr, *lmask = nullptr, *nbr = nullptr;
  uint32_t *nbr_all = nullptr;  // [Q-1][fs] every lattice neighbour: the fused path; nbr (centres only) and Fbuf: the split path
  int pf_blocks = 0;            // L2 prefetch distance of the fused kernel in blocks (TXG_PF)
  bool fused = false;           // forces + collide in one kernel (order-4 stencil; TXG_SPLIT=1 forces the split path)
  // one-pass step (opt-in, TXG_LAG=1; lag_schedule.h): the fused kernel also sums the next step's densities
  // out of L2, band by band behind the collision front.  rho_next receives them; the buffers swap every step.
  bool lag_wanted = false, lag = false;
  int lag_rows = 64, lag_planes = 2, lag_mpos = 512;  // TXG_LAG_ROWS / TXG_LAG_PLANES / TXG_LAG_MPOS
  // rho tiles in shared memory (opt-in, TXG_RHOTILE=1): window starts per block of the fused kernel (k_build_rtab)
  bool tile_wanted = false, tile = false;
  uint32_t *rtab = nullptr;
  uint32_t *rtab_lag = nullptr;  // the same per block of the one-pass launch (TXG_LAG=1 TXG_RHOTILE=1)
  // band blocks (band_kernel.cuh; TXG_BAND=0 switches them off): bit rows instead of the adjacency table, density windows in shared memory
  bool band_wanted = false, band = false;  // (measured slower than the table kernel: profiles/r2c_band_results.txt)
  // staged form (stage_kernel.cuh; TXG_STAGE=0 switches it off): the default K2 of the order-4 step where it applies --
  // 2.5-3.5 % faster than the table kernel at 512^3 (profiles/r2d_stage_results.txt)
  bool stage_wanted = true, stage = false;
  bool forces_tile_on = true;  // TXG_FORCES_TILE=0: the map-walking k_forces for the wide stencils
  int tile_march = 16;         // planes a block of k_forces_tile marches over (TXG_TILE_MARCH)
  // orders 8, 10 without face BCs: k_step_tile = forces + collide + push in one kernel.  Opt-in (TXG_WIDE_FUSED=1): 23.7 ms
  // against 8.4 + 8.5 ms for k_forces_tile + k_collide at 512^3 (the box fill runs at the collision's 16 warps per SM)
  bool wide_fused = false;
  int stage_lb = 0;  // positions per block of the staged kernel
  int stage_pf = 0;  // its L2 prefetch distance in blocks (TXG_STAGE_PF)
  bool stage_clc = false;  // k_step_stage_clc: blocks take over the next block of the grid (TXG_STAGE_CLC)
  unsigned char *adjc = nullptr;  // compressed adjacency records of the staged kernel (TXG_STAGE_ADJC; stage_kernel.cuh AdjcGeom)
  bool stage_adjc = false;
  int stage_pg = 0;  // TXG_STAGE_PG: the staged kernel prefetches the gathers of a warp's next item into L1
  int stage_v = 0;         // block shape of the staged kernel: index into KernelSet::stage_warps (TXG_STAGE_WARPS = 4, 6, 12)
  // Two consecutive steps (buffer parity p -> p) of the default path as ONE CUDA graph: a thin z-slab (strong scaling)
  // or a small box spends its time in launch gaps and NCCL call overhead, not in kernels.  Built lazily after two eager
  // steps (the halo staging buffers exist by then), dropped at every walls upload; TXG_GRAPH=0 switches it off.
  cudaGraphExec_t step_graph[2] = {nullptr, nullptr};
  int64_t graph_launches[2] = {0, 0};
  bool graph_wanted = true, graph_failed = false;
  int eager_steps = 0;
  uint32_t *adjm = nullptr;        // [Q][fs] the adjacency rows and, as row Q-1, the mask row: ONE tensor for the staged kernel
  CUtensorMap tm_f[2], tm_adj;     // the two population buffers and adjm as 2-D tensors
  int band_lb = 1024;                 // positions per block (TXG_BAND_LB)
  int band_prefetch = 1;              // L2 prefetch of the block's next chunk (TXG_BAND_PF)
  BitrowEntry *bitrows = nullptr;
  uint32_t *rowend = nullptr, *xrow = nullptr;
  BandBlock *band_blocks = nullptr;
  BandParams band_params;
  std::vector<int> band_plane_block0;  // [NZl + 1] first block of every owned plane
  int band_smem = 0;
  // pull form of the band step (band_kernel.cuh; TXG_PULL=0 keeps the push): the population buffer holds the COLLIDED
  // populations g of the last step at their own nodes (state_g); the reference's fi = stream + bounce-back of g is
  // formed by the gathers of the next step, or by materialise() for every path that wants fi itself
  bool pull_wanted = false, pull = false, state_g = false;
  double *rho_next = nullptr;
  LagRowDev *lag_rows_dev = nullptr;  // schedule rows (the M blocks read them)
  LagCRow *lag_crows_dev = nullptr;   // their C parts, copied into the kernel's constant table before every launch
  unsigned *lag_done = nullptr;       // [rows] C blocks finished, zeroed before every launch
  LagMeta lag_meta;
  unsigned lag_nrows = 0, lag_grid_x = 0;
  double *wallrec = nullptr;    // [S*D + D][fs]
  double *Fbuf = nullptr;       // [S*D][fs] forces of the current step (k_forces -> k_collide)
  double *halo_recv = nullptr;  // NCCL staging: [2 faces][S][NCROSS][fluid nodes of the boundary plane]
  size_t halo_recv_doubles = 0;
  std::vector<long long> plane_off;
  long long nstore = 0;
  long long alloc_nstore = -1;  // fluid-node count the position-indexed arrays are allocated for
  int *counters = nullptr;  // [0] bad wall codes, [1] fluid nodes next to 900-902 walls
  double *staging = nullptr;
  size_t staging_bytes = 0;
  double *f_old = nullptr;  // DistributionCalcDeltaNorm
  bool have_old = false;
  unsigned long long *norm_bits = nullptr;
  // exported copies (lazily allocated)
  double *x_rho = nullptr, *x_u = nullptr, *x_F = nullptr, *x_rhot = nullptr, *x_prs = nullptr, *x_velt = nullptr;
  bool walls_set = false, state_set = false, rho_current = false;
  // external face BCs (lbm_bc.F90): bc_mode = some face is DIRICHLET / NEUMANN / VELOCITY.
  // The step then runs in the reference's own order (collide first, FlowApplyBCs last) on the split
  // kernels, and Fbuf holds the forces of FlowCalcRhoForces between steps.
  bool bc_mode = false, forces_current = false;
  // face BCs on the fused K2 (order 4; TXG_SPLIT=1: off): every node is collided by k_step_fused with forces re-formed from
  // the stored densities; the nodes of the BC faces -- whose populations BCApply changed after those densities were summed --
  // are then collided AGAIN by k_collide<FACE> with the forces FlowCalcRhoForces stored for them (same push targets: the
  // second result replaces the first)
  bool bc_fused = false;
  LatticeTab lt;
  FaceDesc faces[6];
  bool face_here[6] = {false, false, false, false, false, false};  // this rank holds the face
  ReflectPairs reflect[6];  // (n <- p) lists of the BC_REFLECTING faces
  bool has_reflecting = false;
  bool spec_any = false;  // some rank of the run has free-slip contacts: every rank joins exchange_parked
  std::vector<int> bc_order;                                       // faces in BCApply's execution order
  double *bc_vals[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool pressure_outlet[6] = {false, false, false, false, false, false};  // flow%bc_flags(b) .eq. BC_PRESSURE_OUTLET
  double outlet_pressure[6] = {0., 0., 0., 0., 0., 0.};                  // flow%bc_data(1,b)
  // free-slip walls (900-902): slots rewritten after every push (bc_kernels.cuh)
  uint32_t *spec_dst = nullptr, *spec_src = nullptr;
  double *spec_tmp = nullptr;
  long long spec_n = 0;
  int phase = 0;  // position in the six-procedure sequence of LBMRun2
  // multi-GPU
  ncclComm_t comm = nullptr;
  int up = -1, down = -1;  // z-neighbour ranks (or -1)
  // measurement
  bool timing = false;
  std::deque<KernelTimer> timers;
  std::vector<cudaEvent_t> ev_pool;  // recycled timing events
  int timers_pending = 0;
  float last_ms = 0.f;
  int64_t last_launches = 0, launches = 0;
  std::string err;
};

#define TXG_FAIL(h, code, ...)                      \
  do {                                              \
    char buf_[512];                                 \
    snprintf(buf_, sizeof buf_, __VA_ARGS__);       \
    if (h) (h)->err = buf_; else g_create_error = buf_; \
    return (code);                                  \
  } while (0)

#define TXG_CUDA(h, call)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      TXG_FAIL(h, TXG_ERR_LIB, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
               cudaGetErrorString(e_));                                                         \
  } while (0)

#define TXG_NCCL(h, call)                                                                          \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != 0)                                                                                   \
      TXG_FAIL(h, TXG_ERR_LIB, "NCCL error %d at %s:%d: %s", r_, __FILE__, __LINE__, g_nccl.GetErrorString(r_)); \
  } while (0)

#define TXG_TRY(expr)        \
  do {                       \
    int rc_ = (expr);        \
    if (rc_) return rc_;     \
  } while (0)

static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ------------------------------------------------------------------ kernel timing
static void drain_timers(txg_flow *h);
static void release_lag_table(txg_flow *h);
static void drop_step_graphs(txg_flow *h);
static int check_eos(txg_flow *h);
static KernelTimer *timer_for(txg_flow *h, const char *name) {
  for (auto &t : h->timers)
    if (t.name == name || strcmp(t.name, name) == 0) return &t;
  h->timers.push_back(KernelTimer());  // (a deque: the entries of the other kernels stay where they are)
  h->timers.back().name = name;
  return &h->timers.back();
}
// timing events are recycled: a long timed run holds at most TIMER_DRAIN_AT pairs
constexpr int TIMER_DRAIN_AT = 2048;
static cudaEvent_t timer_event(txg_flow *h) {
  if (!h->ev_pool.empty()) {
    cudaEvent_t e = h->ev_pool.back();
    h->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return e;
}
struct ScopedKernel {
  txg_flow *h;
  KernelTimer *t = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  // count = false: a timed span that is not one of this library's kernel launches (the NCCL exchanges)
  ScopedKernel(txg_flow *h_, const char *name, cudaStream_t s_, bool count = true) : h(h_), s(s_) {
    if (count) h->launches++;
    t = timer_for(h, name);
    if (h->timing) {
      a = timer_event(h);
      b = timer_event(h);
      if (a && b) {
        cudaEventRecord(a, s);
      } else {  // out of events: this launch is counted, not timed
        if (a) h->ev_pool.push_back(a);
        if (b) h->ev_pool.push_back(b);
        a = b = nullptr;
      }
    }
    t->launches++;
  }
  ~ScopedKernel() {
    if (a && b) {
      cudaEventRecord(b, s);
      t->pending.emplace_back(a, b);
      if (++h->timers_pending >= TIMER_DRAIN_AT) drain_timers(h);
    }
  }
};
static void drain_timers(txg_flow *h) {
  for (auto &t : h->timers) {
    for (auto &pr : t.pending) {
      float ms = 0.f;
      cudaEventSynchronize(pr.second);
      cudaEventElapsedTime(&ms, pr.first, pr.second);
      t.ms += ms;
      h->ev_pool.push_back(pr.first);
      h->ev_pool.push_back(pr.second);
    }
    t.pending.clear();
  }
  h->timers_pending = 0;
}

// ------------------------------------------------------------------ set-up
static int select_kernels(txg_flow *h) {
  const txg_config &c = h->cfg;
  const bool mrt = c.relaxation_mode == TXG_RELAXATION_MODE_MRT;
  bool ok = false;
  if (c.discretization == TXG_D3Q19_DISCRETIZATION) {
    if (h->S == 1) ok = kernel_set_d3q19_s1(mrt, c.isotropy_order, &h->ks);
    if (h->S == 2) ok = kernel_set_d3q19_s2(mrt, c.isotropy_order, &h->ks);
    if (h->S == 3) ok = kernel_set_d3q19_s3(mrt, c.isotropy_order, &h->ks);
    if (h->S == 4) ok = kernel_set_d3q19_s4(mrt, c.isotropy_order, &h->ks);
    if (h->S == 5) ok = kernel_set_d3q19_s5(mrt, c.isotropy_order, &h->ks);
  } else {
    if (h->S == 1) ok = kernel_set_d2q9_s1(mrt, c.isotropy_order, &h->ks);
    if (h->S == 2) ok = kernel_set_d2q9_s2(mrt, c.isotropy_order, &h->ks);
    if (h->S == 3) ok = kernel_set_d2q9_s3(mrt, c.isotropy_order, &h->ks);
    if (h->S == 4) ok = kernel_set_d2q9_s4(mrt, c.isotropy_order, &h->ks);
    if (h->S == 5) ok = kernel_set_d2q9_s5(mrt, c.isotropy_order, &h->ks);
  }
  if (!ok)
    TXG_FAIL(h, TXG_ERR_SUP,
             "no device kernels for discretization %d, ncomponents %d, isotropy order %d (built: D3Q19 order 4/8, "
             "D2Q9 order 4/8/10, 1-5 components; D3 order 10 is an LBMError in the reference too)",
             c.discretization, h->S, c.isotropy_order);
  return 0;
}

static void fill_phys(txg_flow *h) {
  const txg_config &c = h->cfg;
  Phys &p = h->p;
  memset(&p, 0, sizeof p);
  const bool d3 = c.discretization == TXG_D3Q19_DISCRETIZATION;
  for (int m = 0; m < h->S; ++m) {
    const bool mrt = c.relaxation_mode == TXG_RELAXATION_MODE_MRT;
    p.inv_tau[m] = 1.0 / c.tau[m];
    const double s_c = mrt ? c.s_c[m] : 1.0 / c.tau[m];  // lbm_relaxation.F90:133
    const double rates[7] = {c.s_c[m], c.s_e[m], c.s_e2[m], c.s_q[m], c.s_nu[m], c.s_pi[m], c.s_m[m]};
    for (int r = 0; r < h->Q; ++r) {
      const int which = d3 ? D3Q19::rate_of(r) : D2Q9::rate_of(r);
      const int norm = d3 ? D3Q19::Mnorm(r) : D2Q9::Mnorm(r);
      p.mrt_rate[m][r] = rates[which] / (double)norm;
    }
    p.mm[m] = c.mm[m];
    p.d_k[m] = 1. - 2. / (3. * c.mm[m]);  // lbm_component.F90:158
    p.mmot[m] = c.mm[m] * s_c;            // lbm_flow.F90:519-521
    for (int k = 0; k < h->S; ++k) p.gf[m][k] = c.gf[m][k];
    p.eos_kind[m] = c.use_nonideal_eos ? c.eos_type[m] : TXG_EOS_DENSITY;
    p.eos_rho0[m] = c.eos_rho0[m];
    p.eos_psi0[m] = c.eos_psi0[m];
    p.pr_a[m] = c.eos_pr_a[m];
    p.pr_b[m] = c.eos_pr_b[m];
    p.pr_R[m] = c.eos_pr_R[m];
    p.pr_T[m] = c.eos_pr_T[m];
    {  // alpha of EOSApply_PR (lbm_eos.F90:331-332); 0.37464, 1.54226, 0.26992 are default-real literals there
      const double om = c.eos_pr_omega[m];
      const double k = (double)0.37464f + (double)1.54226f * om - (double)0.26992f * (om * om);
      const double t = 1. + k * (1. - sqrt(c.eos_pr_T[m] / c.eos_pr_Tc[m]));
      p.pr_alpha[m] = t * t;
    }
    p.pr_c0g[m] = 6.0 * c.gf[m][m];  // dist%disc%c_0 * g_mm (lbm_flow.F90:799)
  }
  for (int d = 0; d < 3; ++d) p.gvt[d] = c.gvt[d];
  p.nminerals = c.nminerals;
  p.fluidfluid = c.fluidfluid_forces;
  p.fluidsolid = c.fluidsolid_forces;
  p.body = c.body_forces;
  p.eos = c.use_nonideal_eos;
  p.gw = h->gw;
  p.eos_bad = h->counters + 3;
}

static int validate(const txg_config *c) {
  txg_flow *h = nullptr;
  if (!c) TXG_FAIL(h, TXG_ERR_ARG_NULL, "txg_create: null config");
  if (c->struct_bytes != (int32_t)sizeof(txg_config))
    TXG_FAIL(h, TXG_ERR_ARG_WRONG, "txg_config.struct_bytes = %d but the library was built with %zu", c->struct_bytes,
             sizeof(txg_config));
  const bool d3 = c->discretization == TXG_D3Q19_DISCRETIZATION, d2 = c->discretization == TXG_D2Q9_DISCRETIZATION;
  if (!d3 && !d2) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "invalid discretization %d", c->discretization);
  if (c->ndims != (d3 ? 3 : 2)) TXG_FAIL(h, TXG_ERR_ARG_WRONG, "ndims %d does not match the discretization", c->ndims);
  if (c->ncomponents < 1 || c->ncomponents > TXG_NMAX_COMPONENTS)
    TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "ncomponents %d out of range", c->ncomponents);
  if (c->NX < 1 || c->NY < 1 || c->NZ < 1) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "invalid box %d x %d x %d", c->NX, c->NY, c->NZ);
  if (d2 && c->NZ != 1) TXG_FAIL(h, TXG_ERR_ARG_WRONG, "NZ must be 1 for D2Q9");
  if (c->relaxation_mode != TXG_RELAXATION_MODE_SRT && c->relaxation_mode != TXG_RELAXATION_MODE_MRT)
    TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "invalid relaxation mode in LBM");  // lbm_relaxation.F90:166
  const int Rneed = stencil_radius(c->isotropy_order);
  if (c->stencil_size_rho < Rneed || c->stencil_size_rho > 3)
    TXG_FAIL(h, TXG_ERR_ARG_WRONG, "stencil_size_rho %d too small for isotropy order %d (needs %d)", c->stencil_size_rho,
             c->isotropy_order, Rneed);
  if ((long long)c->NX * c->NY * (c->ndims == 3 ? c->zl + 2 * c->stencil_size_rho : 1) >= (1ll << 31))
    TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "slab of %d x %d x %d nodes exceeds the 31-bit node index of the fluid list", c->NX, c->NY, c->zl);
  if (c->NX <= 2 * c->stencil_size_rho || c->NY <= 2 * c->stencil_size_rho)
    TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "box too small for the stencil");
  if (c->nminerals < 1 || c->nminerals > TXG_MAX_MINERALS) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "nminerals %d out of range", c->nminerals);
  if (c->nranks < 1 || c->rank < 0 || c->rank >= c->nranks) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "invalid rank %d of %d", c->rank, c->nranks);
  if (d2 && c->nranks != 1) TXG_FAIL(h, TXG_ERR_SUP, "D2Q9 runs on one rank (z-slab decomposition is 3-D only)");
  if (c->zs < 0 || c->zl < 1 || c->zs + c->zl > c->NZ) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "invalid z-slab [%d, %d) of %d", c->zs, c->zs + c->zl, c->NZ);
  if (c->nranks == 1 && (c->zs != 0 || c->zl != c->NZ)) TXG_FAIL(h, TXG_ERR_ARG_WRONG, "a single rank must own the whole box");
  if (c->nranks > 1 && c->zl < c->stencil_size_rho) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "z-slab thinner than the stencil");
  // one rank, periodic z: the ghost planes are copies of the R opposite owned planes
  if (c->ndims == 3 && c->nranks == 1 && c->periodic[2] && c->zl < c->stencil_size_rho)
    TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "periodic z with NZ = %d thinner than the stencil (%d planes)", c->zl, c->stencil_size_rho);
  for (int b = 0; b < 6; ++b) {
    const int fl = c->bc_flags[b];
    if (fl < TXG_BC_NULL || fl > TXG_BC_VELOCITY)
      TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "bc_flags[%d] = %d is not one of BC_NULL .. BC_VELOCITY (lbm_definitions.h:29-35)", b, fl);
    if (b >= 2 * c->ndims && fl != TXG_BC_NULL && fl != TXG_BC_PERIODIC) TXG_FAIL(h, TXG_ERR_ARG_WRONG, "bc_flags[%d] set on a 2-D box", b);
    if (fl >= TXG_BC_REFLECTING && b < 2 * c->ndims && c->periodic[b / 2])
      TXG_FAIL(h, TXG_ERR_ARG_WRONG, "Multiple BCs provided for boundary %d: periodic and bc_flags = %d (lbm_flow.F90:1069-1071)", b, fl);
  }
  // BCApplyReflectingD2 on XM reads ci(p, Z_DIRECTION) of a two-column array (lbm_bc.F90:1001: out of bounds in the
  // reference itself): there is no defined behaviour to reproduce
  if (c->ndims == 2 && c->bc_flags[TXG_BOUNDARY_XM] == TXG_BC_REFLECTING)
    TXG_FAIL(h, TXG_ERR_SUP, "BC_REFLECTING on the xm face of a 2-D box indexes ci(:, Z_DIRECTION) out of bounds in the reference (lbm_bc.F90:1001)");
  for (int m = 0; m < c->ncomponents; ++m) {
    if (!(c->tau[m] > 0.) || !(c->mm[m] > 0.)) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "component %d: tau and mm must be positive", m + 1);
    if (c->use_nonideal_eos && (c->eos_type[m] < TXG_EOS_DENSITY || c->eos_type[m] > TXG_EOS_THERMO))
      TXG_FAIL(h, 1, "Invalid EOS type");  // lbm_eos.F90:167
    if (c->use_nonideal_eos && c->eos_type[m] == TXG_EOS_PR && !(c->eos_pr_Tc[m] > 0.))
      TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "component %d: EOS_PR needs eos_pr_Tc > 0 (EOSSetFromOptions_PR, lbm_eos.F90:287)", m + 1);
  }
  return 0;
}

// ------------------------------------------------------------------ external face BCs: set-up
// lattice tables for the run-time surface kernels
template <class L>
static void fill_lattice_tab(LatticeTab &lt) {
  lt.Q = L::Q;
  lt.D = L::D;
  for (int n = 0; n < 19; ++n) {
    for (int d = 0; d < 3; ++d) lt.c[n][d] = (n < L::Q && d < L::D) ? L::c(n, d) : 0;
    lt.w[n] = n < L::Q ? L::w(n) : 0.;
  }
}

static void setup_faces(txg_flow *h) {
  const txg_config &c = h->cfg;
  if (c.discretization == TXG_D3Q19_DISCRETIZATION)
    fill_lattice_tab<D3Q19>(h->lt);
  else
    fill_lattice_tab<D2Q9>(h->lt);
  const int D = h->D;
  const int NZl = D == 3 ? c.zl : 1;
  const int N[3] = {c.NX, c.NY, NZl};
  for (int b = 0; b < 2 * D; ++b) {
    FaceDesc &f = h->faces[b];
    memset(&f, 0, sizeof f);
    f.axis = b / 2;
    f.sign = (b % 2 == 0) ? 1 : -1;
    f.coord = (b % 2 == 0) ? 0 : N[f.axis] - 1;
    f.t1 = f.axis == 0 ? 1 : 0;
    f.t2 = f.axis == 2 ? 1 : 2;
    f.n1 = N[f.t1];
    f.n2 = D == 3 ? N[f.t2] : 1;
    f.type = c.bc_flags[b];
    // x and y faces exist on every z-slab; zm on the first, zp on the last (info%zs.eq.1, info%ze.eq.NZ)
    h->face_here[b] = f.axis < 2 || (b == TXG_BOUNDARY_ZM ? c.zs == 0 : c.zs + c.zl == c.NZ);
  }
  // BCApplyReflectingD3 / D2 (lbm_bc.F90:825-1073): for every incoming direction n (ascending) and every p (ascending)
  // that passes the face's own test, fi(n) = fi(p).  The tests are the mirror rule on every face except XM in 3-D,
  // which compares ci(n, X) with -ci(p, Z) (:849) -- restated as written.
  h->has_reflecting = false;
  for (int b = 0; b < 2 * D; ++b) {
    ReflectPairs &rp = h->reflect[b];
    rp.count = 0;
    if (c.bc_flags[b] != TXG_BC_REFLECTING) continue;
    h->has_reflecting = true;
    const FaceDesc &f = h->faces[b];
    const auto &ci = h->lt.c;
    auto match = [&](int n, int p) -> bool {
      if (D == 3) {
        if (b == 0) return ci[n][1] == ci[p][1] && ci[n][2] == ci[p][2] && ci[n][0] == -ci[p][2];
        if (b == 1) return ci[n][1] == ci[p][1] && ci[n][2] == ci[p][2] && ci[n][0] == -ci[p][0];
        if (b < 4) return ci[n][0] == ci[p][0] && ci[n][2] == ci[p][2] && ci[n][1] == -ci[p][1];
        return ci[n][0] == ci[p][0] && ci[n][1] == ci[p][1] && ci[n][2] == -ci[p][2];
      }
      if (b == 1) return ci[n][1] == ci[p][1] && ci[n][0] == -ci[p][0];
      if (b == 2 || b == 3) return ci[n][0] == ci[p][0] && ci[n][1] == -ci[p][1];
      return false;
    };
    for (int n = 1; n < h->Q; ++n) {
      if (f.sign * ci[n][f.axis] <= 0) continue;
      for (int p = 1; p < h->Q; ++p)
        if (match(n, p) && rp.count < 64) {
          rp.n[rp.count] = (unsigned char)n;
          rp.p[rp.count] = (unsigned char)p;
          ++rp.count;
        }
    }
  }
  // BCApply (lbm_bc.F90:781-807): every BC type once, in the order of its first face; inside a type
  // the faces in the order xm, xp, ym, yp, zm, zp
  h->bc_order.clear();
  bool done[16] = {false};
  done[TXG_BC_NULL] = done[TXG_BC_PERIODIC] = true;
  for (int side = 0; side < 2 * D; ++side) {
    const int type = c.bc_flags[side];
    if (done[type]) continue;
    done[type] = true;
    for (int b = 0; b < 2 * D; ++b)
      if (c.bc_flags[b] == type) h->bc_order.push_back(b);
  }
}

extern "C" int txg_config_defaults(txg_config *c) {
  if (!c) return TXG_ERR_ARG_NULL;
  memset(c, 0, sizeof *c);
  c->struct_bytes = (int32_t)sizeof *c;
  c->ndims = 3;
  c->discretization = TXG_D3Q19_DISCRETIZATION;
  c->ncomponents = 1;
  c->NX = c->NY = c->NZ = 1;
  c->zs = 0;
  c->zl = 1;
  c->stencil_size_rho = 1;
  c->relaxation_mode = TXG_RELAXATION_MODE_SRT;
  c->isotropy_order = 4;
  c->nminerals = 1;
  c->nranks = 1;
  for (int m = 0; m < TXG_NMAX_COMPONENTS; ++m) {
    c->tau[m] = c->s_c[m] = c->s_e[m] = c->s_e2[m] = c->s_q[m] = c->s_nu[m] = c->s_pi[m] = c->s_m[m] = 1.0;
    c->mm[m] = 1.0;
    c->eos_rho0[m] = 1.0;
    c->eos_psi0[m] = 1.0;
    c->eos_type[m] = TXG_EOS_DENSITY;
    // EOSSetFromOptions_PR defaults (lbm_eos.F90:282-316); 0.0778, 0.45724, 0.9, 0.344 are default-real literals
    c->eos_pr_a[m] = 2.0 / 49.;
    c->eos_pr_b[m] = 2.0 / 21.;
    c->eos_pr_R[m] = 1.0;
    c->eos_pr_Tc[m] = c->eos_pr_a[m] / c->eos_pr_b[m] * (double)0.0778f / (double)0.45724f / c->eos_pr_R[m];
    c->eos_pr_T[m] = (double)0.9f * c->eos_pr_Tc[m];
    c->eos_pr_omega[m] = (double)0.344f;
  }
  return 0;
}

extern "C" const char *txg_last_error(txg_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int alloc_zero(txg_flow *h, void **p, size_t bytes) {
  TXG_CUDA(h, cudaMalloc(p, bytes));
  TXG_CUDA(h, cudaMemsetAsync(*p, 0, bytes, h->s_main));
  return 0;
}

extern "C" int txg_destroy(txg_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  drain_timers(h);
  drop_step_graphs(h);
  release_lag_table(h);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  h->ev_pool.clear();
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  void *ptrs[] = {h->f[0], h->f[1], h->rho, h->rho_true != h->rho ? h->rho_true : nullptr, h->u0, h->gw, h->cls,
                  h->nbmask, h->ffmask, h->P, h->list, h->lmask, h->nbr, h->nbr_all, h->wallrec, h->halo_recv, h->counters, h->Fbuf, h->staging, h->f_old, h->norm_bits, h->x_rho, h->x_u, h->x_F,
                  h->x_rhot, h->x_prs, h->x_velt, h->spec_dst, h->spec_src, h->spec_tmp, h->bc_vals[0], h->bc_vals[1],
                  h->bc_vals[2], h->bc_vals[3], h->bc_vals[4], h->bc_vals[5], h->rho_next, h->lag_rows_dev, h->lag_crows_dev, h->lag_done, h->rtab, h->rtab_lag,
                  h->bitrows, h->rowend, h->xrow, h->band_blocks, h->adjm, h->adjc};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (h->ev_a) cudaEventDestroy(h->ev_a);
  if (h->ev_b) cudaEventDestroy(h->ev_b);
  if (h->ev_step0) cudaEventDestroy(h->ev_step0);
  if (h->ev_step1) cudaEventDestroy(h->ev_step1);
  if (h->s_main) cudaStreamDestroy(h->s_main);
  if (h->s_comm) cudaStreamDestroy(h->s_comm);

  delete h;
  return 0;
}

extern "C" int txg_create(txg_handle *out, const txg_config *cfg, int device) {
  if (!out) {
    g_create_error = "txg_create: null handle pointer";
    return TXG_ERR_ARG_NULL;
  }
  *out = nullptr;
  TXG_TRY(validate(cfg));
  int ndev = 0;
  {
    txg_flow *h = nullptr;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      TXG_FAIL(h, TXG_ERR_LIB, "no CUDA device available (%s); this library has no CPU path",
               e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "device %d out of range (%d devices)", device, ndev);
  }
  txg_flow *h = new txg_flow();
  h->cfg = *cfg;
  h->device = device;
  h->S = cfg->ncomponents;
  h->D = cfg->ndims;
  h->Q = cfg->discretization == TXG_D3Q19_DISCRETIZATION ? 19 : 9;
  h->R = cfg->stencil_size_rho;
  int rc = 0;
  auto fail = [&](int code) {
    g_create_error = h->err;
    txg_destroy(h);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) {
    h->err = "cudaSetDevice failed";
    return fail(TXG_ERR_LIB);
  }
  if ((rc = select_kernels(h))) return fail(rc);
  if (h->ks.set_forces_tile_smem && h->ks.set_forces_tile_smem() != 0) {
    cudaGetLastError();
    h->ks.forces_tile = nullptr;  // (a box that does not fit the device's shared memory: keep k_forces)
  }
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
      h->err = "cudaGetDeviceProperties failed";
      return fail(TXG_ERR_LIB);
    }
    h->num_sms = prop.multiProcessorCount;
    // measured at 512^3 porous, k_step_fused ms: 0: 11.21, 148: 10.83, 222: 10.84, 296: 10.89, 444: 11.01, 592: 11.15, 888: 11.99
    h->pf_blocks = h->num_sms;
    if (const char *pf = getenv("TXG_PF")) h->pf_blocks = atoi(pf);
    if (const char *rt = getenv("TXG_RHOTILE")) h->tile_wanted = rt[0] == '1';
    if (const char *lg = getenv("TXG_LAG")) h->lag_wanted = lg[0] == '1';
    if (const char *v = getenv("TXG_LAG_ROWS")) h->lag_rows = atoi(v);
    if (const char *v = getenv("TXG_LAG_PLANES")) h->lag_planes = atoi(v);
    if (const char *v = getenv("TXG_LAG_MPOS")) h->lag_mpos = atoi(v);
    if (const char *v = getenv("TXG_BAND")) h->band_wanted = v[0] != '0';
    if (const char *v = getenv("TXG_STAGE")) h->stage_wanted = v[0] != '0';
    if (const char *v = getenv("TXG_GRAPH")) h->graph_wanted = v[0] != '0';
    if (const char *v = getenv("TXG_FORCES_TILE")) h->forces_tile_on = v[0] != '0';
    if (const char *v = getenv("TXG_TILE_MARCH")) h->tile_march = std::max(1, atoi(v));
    if (const char *v = getenv("TXG_BAND_LB")) h->band_lb = std::max(16, atoi(v));
    if (const char *v = getenv("TXG_BAND_PF")) h->band_prefetch = atoi(v);
    if (const char *v = getenv("TXG_PULL")) h->pull_wanted = v[0] != '0';
    const char *sp = getenv("TXG_SPLIT");
    for (int b = 0; b < 2 * cfg->ndims; ++b) h->bc_mode = h->bc_mode || cfg->bc_flags[b] >= TXG_BC_REFLECTING;
    // face BCs act between the forces and the collision: they need the split kernels and the force buffer
    h->fused = h->ks.step_fused != nullptr && !(sp && sp[0] == '1') && !h->bc_mode;
    h->bc_fused = h->ks.step_fused != nullptr && !(sp && sp[0] == '1') && h->bc_mode;
    const char *wf = getenv("TXG_WIDE_FUSED");
    h->wide_fused = wf && wf[0] == '1' && h->ks.step_tile != nullptr && h->ks.forces_tile != nullptr && h->forces_tile_on && !(sp && sp[0] == '1') &&
                    !h->bc_mode;
  }
  Grid &g = h->g;
  g.NX = cfg->NX;
  g.NY = cfg->NY;
  g.NZl = cfg->ndims == 3 ? cfg->zl : 1;
  g.R = h->R;
  g.Rz = cfg->ndims == 3 ? h->R : 0;
  g.perx = cfg->periodic[0];
  g.pery = cfg->periodic[1];
  g.plane = (long long)g.NX * g.NY;
  g.nnodes = (long long)g.NZl * g.plane;
  g.nE = (long long)(g.NZl + 2 * g.Rz) * g.plane;
  g.fs = 0;
  g.own0 = g.own1 = 0;
  g.P = g.list = nullptr;
  g.cnx = g.NX + 2 * g.R;
  g.cny = g.NY + 2 * g.R;
  auto body = [&]() -> int {
    TXG_CUDA(h, cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
    // the communication stream at the highest priority: the block scheduler then places the few blocks of an NCCL or
    // unpack kernel ahead of the waiting blocks of the interior kernel; at equal priority they only start when the interior
    // grid has been dispatched to the end, and the halo is not overlapped at all (measured: halo_f span = interior K2 time)
    {
      int prio_low = 0, prio_high = 0;
      TXG_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
      TXG_CUDA(h, cudaStreamCreateWithPriority(&h->s_comm, cudaStreamNonBlocking, prio_high));
    }

    TXG_CUDA(h, cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    TXG_CUDA(h, cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    TXG_CUDA(h, cudaEventCreate(&h->ev_step0));
    TXG_CUDA(h, cudaEventCreate(&h->ev_step1));
    // the population / density arrays are sized by the fluid-node count: allocated in txg_set_walls
    TXG_TRY(alloc_zero(h, (void **)&h->cls, (size_t)(g.NZl + 2 * g.Rz) * g.cny * g.cnx));
    TXG_TRY(alloc_zero(h, (void **)&h->nbmask, (size_t)g.nnodes * sizeof(uint32_t)));
    if (h->ks.ff_words) TXG_TRY(alloc_zero(h, (void **)&h->ffmask, (size_t)h->ks.ff_words * g.nnodes * sizeof(uint32_t)));
    TXG_TRY(alloc_zero(h, (void **)&h->counters, 8 * sizeof(int)));  // [4]: blocks of k_step_fused_tile that gave up waiting
    TXG_TRY(alloc_zero(h, (void **)&h->norm_bits, sizeof(unsigned long long)));
    std::vector<double> gw((size_t)cfg->nminerals * h->S);
    for (int k = 0; k < cfg->nminerals; ++k)
      for (int m = 0; m < h->S; ++m) gw[(size_t)k * h->S + m] = cfg->gw[k][m];
    TXG_CUDA(h, cudaMalloc((void **)&h->gw, gw.size() * sizeof(double)));
    TXG_CUDA(h, cudaMemcpy(h->gw, gw.data(), gw.size() * sizeof(double), cudaMemcpyHostToDevice));
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
    return 0;
  };
  if ((rc = body())) return fail(rc);
  fill_phys(h);
  setup_faces(h);
  // z neighbours of the slab ring
  const int nr = cfg->nranks, r = cfg->rank;
  if (cfg->ndims == 3) {
    h->up = r + 1 < nr ? r + 1 : (cfg->periodic[2] ? 0 : -1);
    h->down = r > 0 ? r - 1 : (cfg->periodic[2] ? nr - 1 : -1);
  }
  *out = h;
  return 0;
}

extern "C" int txg_nccl_unique_id(unsigned char id_out[128]) {
  txg_flow *h = nullptr;
  if (!g_nccl.load()) TXG_FAIL(h, TXG_ERR_LIB, "%s", g_nccl.err.c_str());
  ncclUniqueId id;
  TXG_NCCL(h, g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return 0;
}

extern "C" int txg_comm_init(txg_handle h, const unsigned char id_in[128]) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (h->cfg.nranks == 1) return 0;
  if (!g_nccl.load()) TXG_FAIL(h, TXG_ERR_LIB, "%s", g_nccl.err.c_str());
  TXG_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(id.internal, id_in, 128);
  TXG_NCCL(h, g_nccl.CommInitRank(&h->comm, h->cfg.nranks, id, h->cfg.rank));
  return 0;
}

// ------------------------------------------------------------------ halo exchange along z
// One ghost plane of f each side; R ghost planes of rho each side.  `n_up`/`n_down` are the chunk
// lists (offset into the buffer of the plane to send / the ghost plane to fill).
struct Chunk {
  long long send_off, recv_off;
  long long send_count, recv_count;
};

static int exchange(txg_flow *h, double *buf, const std::vector<Chunk> &to_up, const std::vector<Chunk> &to_down,
                    cudaStream_t s) {
  // to_up: my top planes -> up neighbour's bottom ghost; I receive my bottom ghost from `down`.
  // to_down: my bottom planes -> down neighbour's top ghost; I receive my top ghost from `up`.
  const int nr = h->cfg.nranks;
  if (nr == 1) {
    if (h->up < 0) return 0;  // not periodic in z: ghosts are never read
    // periodic single rank: my own top planes are my bottom ghost and vice versa
    for (const std::vector<Chunk> *v : {&to_up, &to_down})
      for (const Chunk &c : *v) {
        if (c.send_count != c.recv_count) TXG_FAIL(h, TXG_ERR_ARG_WRONG, "periodic z: ghost planes do not mirror the boundary planes");
        if (c.send_count)
          TXG_CUDA(h, cudaMemcpyAsync(buf + c.recv_off, buf + c.send_off, c.send_count * sizeof(double), cudaMemcpyDeviceToDevice, s));
      }
    // (device-to-device copies, not kernels: not counted in `launches`)
    return 0;
  }
  if (!h->comm) TXG_FAIL(h, TXG_ERR_ORDER, "nranks > 1 but txg_comm_init was not called");
  TXG_NCCL(h, g_nccl.GroupStart());
  for (const Chunk &c : to_up) {
    if (h->up >= 0 && c.send_count) TXG_NCCL(h, g_nccl.Send(buf + c.send_off, (size_t)c.send_count, ncclFloat64, h->up, h->comm, s));
    if (h->down >= 0 && c.recv_count) TXG_NCCL(h, g_nccl.Recv(buf + c.recv_off, (size_t)c.recv_count, ncclFloat64, h->down, h->comm, s));
  }
  for (const Chunk &c : to_down) {
    if (h->down >= 0 && c.send_count) TXG_NCCL(h, g_nccl.Send(buf + c.send_off, (size_t)c.send_count, ncclFloat64, h->down, h->comm, s));
    if (h->up >= 0 && c.recv_count) TXG_NCCL(h, g_nccl.Recv(buf + c.recv_off, (size_t)c.recv_count, ncclFloat64, h->up, h->comm, s));
  }
  TXG_NCCL(h, g_nccl.GroupEnd());
  return 0;
}

// Populations pushed across a z face sit in this slab's ghost-plane positions; they belong in the
// neighbour's boundary plane.  Single rank, periodic z: unpack straight from the own opposite
// ghost plane.  Several ranks: send each ghost plane's crossing directions (one contiguous run of
// positions per (m, n)), receive the neighbours' into the staging buffer, unpack from there.  The
// unpack is masked (k_halo_unpack): a slot whose source node is solid keeps the bounce-back value
// its own node wrote.
static int exchange_f(txg_flow *h, double *buf, cudaStream_t s) {
  if (h->D != 3) return 0;
  ScopedKernel span(h, "halo_f", s, false);
  const Grid &g = h->g;
  const int nr = h->cfg.nranks, Rz = g.Rz;
  const std::vector<long long> &po = h->plane_off;
  // extended planes: bottom ghost Rz-1, bottom owned Rz, top owned Rz+NZl-1, top ghost Rz+NZl
  const long long gb0 = po[Rz - 1], ob0 = po[Rz], ob1 = po[Rz + 1];
  const long long ot0 = po[Rz + g.NZl - 1], gt0 = po[Rz + g.NZl], gt1 = po[Rz + g.NZl + 1];
  const long long nbot = ob1 - ob0, ntop = gt0 - ot0;  // fluid nodes of the boundary planes
  if (nr == 1) {
    if (h->up < 0) return 0;  // not periodic in z: the ghost planes are solid (class 255), nothing is pushed
    // the bottom plane takes the c_z = +1 pushes that left through the top ghost plane, and vice versa
    if (gt1 - gt0 != nbot || ob0 - gb0 != ntop) TXG_FAIL(h, TXG_ERR_ARG_WRONG, "periodic z: ghost planes do not mirror the boundary planes");
    if (nbot) h->ks.halo_unpack<<<blocks_for(nbot, 256), 256, 0, s>>>(g, buf, buf, gt0, 0, h->lmask, ob0, nbot, 1);
    if (ntop) h->ks.halo_unpack<<<blocks_for(ntop, 256), 256, 0, s>>>(g, buf, buf, gb0, 0, h->lmask, ot0, ntop, 0);
    TXG_CUDA(h, cudaGetLastError());
    h->launches += 2;
    return 0;
  }
  if (!h->comm) TXG_FAIL(h, TXG_ERR_ORDER, "nranks > 1 but txg_comm_init was not called");
  const int NC = D3Q19::NCROSS;
  const long long nsend_up = gt1 - gt0, nsend_down = ob0 - gb0;  // fluid nodes of my ghost planes
  // staging: [received from down | received from up | packed for up | packed for down], S * NC rows each
  const size_t need = (size_t)h->S * NC * (size_t)(nbot + ntop + nsend_up + nsend_down);
  if (h->halo_recv_doubles < need) {
    if (h->halo_recv) cudaFree(h->halo_recv);
    h->halo_recv = nullptr;
    TXG_CUDA(h, cudaMalloc((void **)&h->halo_recv, std::max<size_t>(need, 1) * sizeof(double)));
    h->halo_recv_doubles = need;
  }
  const size_t rows = (size_t)h->S * NC;
  double *from_down = h->halo_recv, *from_up = from_down + rows * (size_t)nbot;
  double *to_up = from_up + rows * (size_t)ntop, *to_down = to_up + rows * (size_t)nsend_up;
  // the crossing rows of each ghost plane packed into one message per neighbour (k_halo_pack), received packed
  if (h->up >= 0 && nsend_up) h->ks.halo_pack<<<blocks_for(nsend_up, 256), 256, 0, s>>>(g, buf, to_up, gt0, nsend_up, 1);
  if (h->down >= 0 && nsend_down) h->ks.halo_pack<<<blocks_for(nsend_down, 256), 256, 0, s>>>(g, buf, to_down, gb0, nsend_down, 0);
  TXG_CUDA(h, cudaGetLastError());
  h->launches += 2;
  TXG_NCCL(h, g_nccl.GroupStart());
  if (h->up >= 0 && nsend_up) TXG_NCCL(h, g_nccl.Send(to_up, rows * (size_t)nsend_up, ncclFloat64, h->up, h->comm, s));
  if (h->down >= 0 && nbot) TXG_NCCL(h, g_nccl.Recv(from_down, rows * (size_t)nbot, ncclFloat64, h->down, h->comm, s));
  if (h->down >= 0 && nsend_down) TXG_NCCL(h, g_nccl.Send(to_down, rows * (size_t)nsend_down, ncclFloat64, h->down, h->comm, s));
  if (h->up >= 0 && ntop) TXG_NCCL(h, g_nccl.Recv(from_up, rows * (size_t)ntop, ncclFloat64, h->up, h->comm, s));
  TXG_NCCL(h, g_nccl.GroupEnd());
  if (h->down >= 0 && nbot) h->ks.halo_unpack<<<blocks_for(nbot, 256), 256, 0, s>>>(g, buf, from_down, 0, 1, h->lmask, ob0, nbot, 1);
  if (h->up >= 0 && ntop) h->ks.halo_unpack<<<blocks_for(ntop, 256), 256, 0, s>>>(g, buf, from_up, 0, 1, h->lmask, ot0, ntop, 0);
  TXG_CUDA(h, cudaGetLastError());
  h->launches += 2;
  return 0;
}

// rho (psi) halo: my top R owned planes fill the up neighbour's bottom ghost planes and vice versa;
// both are contiguous runs of positions with matching fluid-node counts.
static int exchange_rho(txg_flow *h, double *buf, cudaStream_t s) {
  if (h->D != 3) return 0;
  ScopedKernel span(h, "halo_rho", s, false);
  const Grid &g = h->g;
  const int Rz = g.Rz;
  const std::vector<long long> &po = h->plane_off;
  std::vector<Chunk> upv, downv;
  for (int m = 0; m < h->S; ++m) {
    const long long base = (long long)m * g.fs;
    // my top R owned planes [NZl, NZl+R) (extended index) -> neighbour's bottom ghost [0, R)
    upv.push_back({base + po[g.NZl], base + po[0], po[g.NZl + Rz] - po[g.NZl], po[Rz] - po[0]});
    // my bottom R owned planes [R, 2R) -> neighbour's top ghost [NZl+R, NZl+2R)
    downv.push_back({base + po[Rz], base + po[g.NZl + Rz], po[2 * Rz] - po[Rz], po[g.NZl + 2 * Rz] - po[g.NZl + Rz]});
  }
  return exchange(h, buf, upv, downv, s);
}

// Pull form: the collided populations that will cross a z face are read by the neighbour slab out of ITS ghost plane:
// my top owned plane's c_z > 0 rows fill the up neighbour's bottom ghost plane, my bottom owned plane's c_z < 0 rows the
// down neighbour's top ghost plane -- one contiguous run of positions per (component, direction), received in place
// (ghost and owned plane hold the same fluid nodes), no unpack kernel.
static int exchange_g(txg_flow *h, double *buf, cudaStream_t s) {
  if (h->D != 3) return 0;
  const Grid &g = h->g;
  const int Rz = g.Rz;
  const std::vector<long long> &po = h->plane_off;
  const long long gb0 = po[Rz - 1], ob0 = po[Rz], ob1 = po[Rz + 1];
  const long long ot0 = po[Rz + g.NZl - 1], gt0 = po[Rz + g.NZl], gt1 = po[Rz + g.NZl + 1];
  std::vector<Chunk> upv, downv;
  for (int m = 0; m < h->S; ++m)
    for (int n = 1; n < h->Q; ++n) {
      const int cz = D3Q19::c(n, 2);
      const long long blk = (long long)(m * h->Q + n) * g.fs;
      if (cz > 0) upv.push_back({blk + ot0, blk + gb0, gt0 - ot0, ob0 - gb0});
      if (cz < 0) downv.push_back({blk + ob0, blk + gt0, ob1 - ob0, gt1 - gt0});
    }
  return exchange(h, buf, upv, downv, s);
}

// ------------------------------------------------------------------ host <-> device layout conversion
static int ensure_staging(txg_flow *h, size_t bytes) {
  if (h->staging_bytes >= bytes) return 0;
  if (h->staging) cudaFree(h->staging);
  h->staging = nullptr;
  h->staging_bytes = 0;
  TXG_CUDA(h, cudaMalloc((void **)&h->staging, bytes));
  h->staging_bytes = bytes;
  return 0;
}

// host array [gz][gy][gx][K*S] (ghost width gw in x,y; gwz in z) -> device SoA (dense over the owned
// nodes, or position-indexed over the fluid nodes)
static int import_field(txg_flow *h, const double *host, int gw, int gwz, int K, double *dst, int dense) {
  const Grid &g = h->g;
  const int dof = h->S * K;
  const size_t plane_elems = (size_t)(g.NX + 2 * gw) * (g.NY + 2 * gw) * dof;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)g.NZl, ((size_t)256 << 20) / (plane_elems * 8)));
  TXG_TRY(ensure_staging(h, plane_elems * 8 * chunk));
  for (int z0 = 0; z0 < g.NZl; z0 += chunk) {
    const int nz = std::min(chunk, g.NZl - z0);
    TXG_CUDA(h, cudaMemcpyAsync(h->staging, host + (size_t)(z0 + gwz) * plane_elems, plane_elems * 8 * nz,
                                cudaMemcpyHostToDevice, h->s_main));
    const long long total = (long long)nz * g.plane * dof;
    k_import_aos<<<blocks_for(total, 256), 256, 0, h->s_main>>>(g, h->staging, dst, gw, h->S, K, dense, z0, nz);
    TXG_CUDA(h, cudaGetLastError());
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  }
  return 0;
}

static int export_field(txg_flow *h, double *host, int gw, int gwz, int K, int S, const double *src, int dense) {
  const Grid &g = h->g;
  const int dof = S * K;
  const size_t plane_elems = (size_t)(g.NX + 2 * gw) * (g.NY + 2 * gw) * dof;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)g.NZl, ((size_t)256 << 20) / (plane_elems * 8)));
  TXG_TRY(ensure_staging(h, plane_elems * 8 * chunk));
  for (int z0 = 0; z0 < g.NZl; z0 += chunk) {
    const int nz = std::min(chunk, g.NZl - z0);
    if (gw > 0)  // keep the caller's ghost entries: round-trip the planes through the staging buffer
      TXG_CUDA(h, cudaMemcpyAsync(h->staging, host + (size_t)(z0 + gwz) * plane_elems, plane_elems * 8 * nz,
                                  cudaMemcpyHostToDevice, h->s_main));
    const long long total = (long long)nz * g.plane * dof;
    k_export_aos<<<blocks_for(total, 256), 256, 0, h->s_main>>>(g, h->staging, src, gw, S, K, dense, z0, nz);
    TXG_CUDA(h, cudaGetLastError());
    TXG_CUDA(h, cudaMemcpyAsync(host + (size_t)(z0 + gwz) * plane_elems, h->staging, plane_elems * 8 * nz,
                                cudaMemcpyDeviceToHost, h->s_main));
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  }
  return 0;
}

// ------------------------------------------------------------------ sparse storage
// Positions of the fluid nodes of the extended slab: count per 256-slot chunk of each plane, scan the
// chunk counts on the host ((NZl+2Rz) * plane/256 integers), fill P and list with a ballot rank
// inside each chunk.  Then size and allocate every position-indexed array.
// device arrays sized by the fluid-node count: freed and re-allocated only when that count changes (a
// second walls upload with the same geometry -- the reference calls WallsSetValues once, a driver that
// re-initialises a run does not pay for 40 GB of cudaFree / cudaMalloc)
static void free_storage(txg_flow *h) {
  for (void **q : {(void **)&h->P, (void **)&h->list, (void **)&h->lmask, (void **)&h->nbr, (void **)&h->nbr_all, (void **)&h->wallrec, (void **)&h->Fbuf, (void **)&h->f[0],
                   (void **)&h->f[1], (void **)&h->rho, (void **)&h->f_old}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  if (h->rho_true && h->cfg.use_nonideal_eos) cudaFree(h->rho_true);
  h->rho_true = nullptr;
  for (void **q : {(void **)&h->rho_next, (void **)&h->lag_rows_dev, (void **)&h->lag_crows_dev, (void **)&h->lag_done, (void **)&h->rtab_lag}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  h->lag = false;
  h->alloc_nstore = -1;
}
// zeroed array of `bytes`: the existing allocation when the storage is being re-used
static int fresh_zero(txg_flow *h, void **p, size_t bytes) {
  if (!*p) TXG_CUDA(h, cudaMalloc(p, bytes));
  TXG_CUDA(h, cudaMemsetAsync(*p, 0, bytes, h->s_main));
  return 0;
}

// entries behind the last component of the density arrays: the bulk-copy windows of the band blocks (<= BAND_CAP_MAX
// doubles) and of k_step_fused_tile may overrun the last position
constexpr int BAND_CAP_MAX = 4096;
constexpr size_t RHO_PAD = BAND_CAP_MAX;

static int build_storage(txg_flow *h) {
  Grid &g = h->g;
  h->have_old = false;
  h->state_set = false;
  g.P = g.list = nullptr;
  const int nzE = g.NZl + 2 * g.Rz;
  const int bpp = (int)((g.plane + 255) / 256);  // chunks per plane
  const long long nchunks = (long long)bpp * nzE;
  unsigned *d_cnt = nullptr;
  TXG_CUDA(h, cudaMalloc((void **)&d_cnt, (size_t)nchunks * sizeof(unsigned)));
  k_count_fluid<<<(unsigned)nchunks, 256, 0, h->s_main>>>(g, h->cls, bpp, d_cnt);
  TXG_CUDA(h, cudaGetLastError());
  std::vector<unsigned> cnt((size_t)nchunks);
  TXG_CUDA(h, cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)nchunks * sizeof(unsigned), cudaMemcpyDeviceToHost, h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  h->plane_off.assign((size_t)nzE + 1, 0);
  long long run = 0;
  for (long long c = 0; c < nchunks; ++c) {
    if (c % bpp == 0) h->plane_off[(size_t)(c / bpp)] = run;
    const unsigned k = cnt[(size_t)c];
    cnt[(size_t)c] = (unsigned)run;
    run += k;
  }
  h->plane_off[(size_t)nzE] = run;
  h->nstore = run;
  g.own0 = h->plane_off[(size_t)g.Rz];
  g.own1 = h->plane_off[(size_t)(g.Rz + g.NZl)];
  // >= nstore + 1 (a solid neighbour maps to the next position), padded so that the aligned
  // chunks of the bulk-copy kernel (40, 64 or 128 positions) never run past the end of a row
  g.fs = ((run + 128 + 127) / 128) * 128;
  if ((long long)h->Q * g.fs >= (1ll << 32)) {
    cudaFree(d_cnt);
    TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "%lld fluid nodes in the slab: the %d populations of one component exceed the 32-bit element index; use more z-slabs", run, h->Q);
  }
  if (h->alloc_nstore != run) {
    free_storage(h);
    h->alloc_nstore = run;
  }
  if (run != g.nE) {
    TXG_CUDA(h, cudaMemcpyAsync(d_cnt, cnt.data(), (size_t)nchunks * sizeof(unsigned), cudaMemcpyHostToDevice, h->s_main));
    if (!h->P) TXG_CUDA(h, cudaMalloc((void **)&h->P, (size_t)(g.nE + 1) * sizeof(uint32_t)));
    if (!h->list) TXG_CUDA(h, cudaMalloc((void **)&h->list, (size_t)std::max<long long>(run, 1) * sizeof(uint32_t)));
    k_fill_fluid<<<(unsigned)nchunks, 256, 0, h->s_main>>>(g, h->cls, bpp, d_cnt, h->P, h->list);
    TXG_CUDA(h, cudaGetLastError());
    const uint32_t total = (uint32_t)run;
    TXG_CUDA(h, cudaMemcpyAsync(h->P + g.nE, &total, sizeof total, cudaMemcpyHostToDevice, h->s_main));
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
    g.P = h->P;
    g.list = h->list;
  }
  cudaFree(d_cnt);
  // position-indexed arrays
  const size_t fbytes = (size_t)h->S * h->Q * g.fs * sizeof(double);
  TXG_TRY(fresh_zero(h, (void **)&h->f[0], fbytes));
  TXG_TRY(fresh_zero(h, (void **)&h->f[1], fbytes));
  h->cur = 0;
  // (+ 256 entries: the bulk-copy windows of k_step_fused_tile may overrun the last position of the last component)
  TXG_TRY(fresh_zero(h, (void **)&h->rho, ((size_t)h->S * g.fs + RHO_PAD) * sizeof(double)));
  if (h->cfg.use_nonideal_eos) {
    if (h->rho_true == h->rho) h->rho_true = nullptr;
    TXG_TRY(fresh_zero(h, (void **)&h->rho_true, ((size_t)h->S * g.fs + RHO_PAD) * sizeof(double)));
  } else {
    h->rho_true = h->rho;
  }
  if (!h->fused) TXG_TRY(fresh_zero(h, (void **)&h->Fbuf, (size_t)h->S * h->D * g.fs * sizeof(double)));
  TXG_TRY(fresh_zero(h, (void **)&h->lmask, (size_t)g.fs * sizeof(uint32_t)));
  const long long nown = g.own1 - g.own0;
  int nrec = 0;
  if (h->fused || h->bc_fused) TXG_TRY(fresh_zero(h, (void **)&h->nbr_all, (size_t)(h->Q - 1) * g.fs * sizeof(uint32_t)));
  if (!h->fused) TXG_TRY(fresh_zero(h, (void **)&h->nbr, (size_t)h->ks.ncen * g.fs * sizeof(uint32_t)));
  if (nown) {
    if (h->fused || h->bc_fused) h->ks.build_nbr_all<<<blocks_for(nown, 128), 128, 0, h->s_main>>>(g, h->nbr_all);
    if (!h->fused) h->ks.build_nbr<<<blocks_for(nown, 128), 128, 0, h->s_main>>>(g, h->nbr);
    TXG_CUDA(h, cudaGetLastError());
    TXG_CUDA(h, cudaMemsetAsync(h->counters + 2, 0, sizeof(int), h->s_main));
    k_gather_mask<<<blocks_for(nown, 256), 256, 0, h->s_main>>>(g, h->nbmask, h->list, h->lmask, h->counters + 2);
    TXG_CUDA(h, cudaGetLastError());
    TXG_CUDA(h, cudaMemcpyAsync(&nrec, h->counters + 2, sizeof nrec, cudaMemcpyDeviceToHost, h->s_main));
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  }
  if (nrec) {
    const size_t nk = (size_t)(h->S * h->D + h->D);
    TXG_TRY(fresh_zero(h, (void **)&h->wallrec, nk * (size_t)g.fs * sizeof(double)));
    h->ks.build_wallrec<<<blocks_for(nown, 128), 128, 0, h->s_main>>>(g, h->p, h->cls, h->lmask, h->ffmask, h->wallrec);
    TXG_CUDA(h, cudaGetLastError());
  }
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  return 0;
}

// (first, count) of the positions of owned planes [z0, z0 + nz)
static inline void plane_range(const txg_flow *h, int z0, int nz, long long *first, long long *count) {
  *first = h->plane_off[(size_t)(h->g.Rz + z0)];
  *count = h->plane_off[(size_t)(h->g.Rz + z0 + nz)] - *first;
}

// ------------------------------------------------------------------ free-slip walls (900-902)
// The slots the mirrors rewrite after every push (specular_table.h has the derivation), rebuilt at every
// walls upload.  Restrictions, each refused with PETSC_ERR_SUP and none silently different from the
// reference: one rank only (source and target of a reflection may sit in different z-slabs), and every
// population reflected off a free-slip wall must land on a fluid node.
static int build_specular(txg_flow *h, int contacts) {
  for (void **q : {(void **)&h->spec_dst, (void **)&h->spec_src, (void **)&h->spec_tmp}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  h->spec_n = 0;
  if (!contacts) return 0;
  const Grid &g = h->g;
  // (several ranks: the table of a slab reads the parked populations of the neighbour slabs' boundary planes out of its
  //  ghost planes, which exchange_parked fills after every push)
  const long long ncls = (long long)(g.NZl + 2 * g.Rz) * g.cny * g.cnx;
  std::vector<uint8_t> cls((size_t)ncls);
  TXG_CUDA(h, cudaMemcpy(cls.data(), h->cls, (size_t)ncls, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> P;
  if (g.P) {
    P.resize((size_t)g.nE + 1);
    TXG_CUDA(h, cudaMemcpy(P.data(), g.P, P.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  // z: one rank wraps inside its own slab; several ranks look into the ghost planes (the neighbour's nodes)
  const int per[3] = {g.perx, g.pery, h->D == 3 && h->cfg.nranks == 1 ? h->cfg.periodic[2] : 0};
  SpecularTable t;
  build_specular_table(h->lt, g.NX, g.NY, g.NZl, g.R, g.Rz, per, cls.data(), g.P ? P.data() : nullptr, g.fs, t);
  if (t.parked)
    TXG_FAIL(h, TXG_ERR_SUP,
             "free-slip walls (900-902): %lld reflected populations would land on a solid node (walls meeting in a corner, "
             "or an obstacle touching the wall); the reference parks them in wall-node storage, which the device does not have",
             t.parked);
  h->spec_n = (long long)t.dst.size();
  if (!h->spec_n) return 0;
  TXG_CUDA(h, cudaMalloc((void **)&h->spec_dst, t.dst.size() * sizeof(uint32_t)));
  TXG_CUDA(h, cudaMalloc((void **)&h->spec_src, t.src.size() * sizeof(uint32_t)));
  TXG_CUDA(h, cudaMalloc((void **)&h->spec_tmp, t.dst.size() * h->S * sizeof(double)));
  TXG_CUDA(h, cudaMemcpy(h->spec_dst, t.dst.data(), t.dst.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  TXG_CUDA(h, cudaMemcpy(h->spec_src, t.src.data(), t.src.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return 0;
}

// Free-slip walls on several ranks: a mirror of axis x or y sends the population that left S along c_n to T = S +
// tangential(c_n), which lies in the next z-plane when c_n,z != 0 -- possibly in the next slab.  The push parked that
// population in slot (opp(n), S); the table of T's slab reads it from its ghost-plane position of S.  So after every push
// each slab's ghost planes take the neighbours' boundary-plane rows of the directions pointing AWAY from this slab (the
// ghost rows pointing away are free: their own pushes were sent on by exchange_f): my top owned plane's c_z < 0 rows fill
// the up neighbour's bottom ghost plane, my bottom owned plane's c_z > 0 rows the down neighbour's top ghost plane.
// (DistributionBouncebackD3, lbm_distribution_function.F90:669-784; the reference gets there through the ghosted fi.)
static int exchange_parked(txg_flow *h, double *buf, cudaStream_t s) {
  if (h->D != 3 || h->cfg.nranks == 1 || !h->spec_any) return 0;
  const Grid &g = h->g;
  const int Rz = g.Rz;
  const std::vector<long long> &po = h->plane_off;
  const long long gb0 = po[Rz - 1], ob0 = po[Rz], ob1 = po[Rz + 1];
  const long long ot0 = po[Rz + g.NZl - 1], gt0 = po[Rz + g.NZl], gt1 = po[Rz + g.NZl + 1];
  std::vector<Chunk> upv, downv;
  for (int m = 0; m < h->S; ++m)
    for (int n = 1; n < h->Q; ++n) {
      const int cz = D3Q19::c(n, 2);
      const long long blk = (long long)(m * h->Q + n) * g.fs;
      if (cz < 0) upv.push_back({blk + ot0, blk + gb0, gt0 - ot0, ob0 - gb0});
      if (cz > 0) downv.push_back({blk + ob0, blk + gt0, ob1 - ob0, gt1 - gt0});
    }
  return exchange(h, buf, upv, downv, s);
}

// rewrite the free-slip slots of a freshly pushed buffer (after the z halo)
static int apply_specular(txg_flow *h, double *f, cudaStream_t s) {
  TXG_TRY(exchange_parked(h, f, s));
  if (!h->spec_n) return 0;
  const long long n = h->spec_n * h->S, stride = (long long)h->Q * h->g.fs;
  {
    ScopedKernel sk(h, "k_specular_gather", s);
    k_specular_gather<<<blocks_for(n, 256), 256, 0, s>>>(f, h->spec_src, h->spec_tmp, h->spec_n, h->S, stride);
  }
  {
    ScopedKernel sk(h, "k_specular_scatter", s);
    k_specular_scatter<<<blocks_for(n, 256), 256, 0, s>>>(f, h->spec_dst, h->spec_tmp, h->spec_n, h->S, stride);
  }
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ walls
// ------------------------------------------------------------------ one-pass step (opt-in, TXG_LAG=1)
// Block schedule of k_step_fused_lag for the current geometry (lag_schedule.h), rebuilt at every walls upload.
// Boxes the schedule does not cover keep the two-kernel step; nothing is refused.
// window starts of k_step_fused_tile for the current geometry (opt-in), rebuilt at every walls upload
static int build_rtab(txg_flow *h) {
  if (h->rtab) cudaFree(h->rtab);
  h->rtab = nullptr;
  h->tile = false;
  const Grid &g = h->g;
  if (!h->tile_wanted || !h->fused || !h->ks.step_fused_tile || h->ks.fused_threads != 128 || g.own1 <= g.own0) return 0;
  const int PB = 4 * h->ks.npw;
  const long long nblocks = (g.own1 - g.own0 + PB - 1) / PB;
  TXG_CUDA(h, cudaMalloc((void **)&h->rtab, (size_t)nblocks * h->ks.rtab_groups * sizeof(uint32_t)));
  h->ks.build_rtab<<<blocks_for(nblocks * h->ks.rtab_groups, 128), 128, 0, h->s_main>>>(g, h->nbr_all, PB, nblocks, h->rtab);
  TXG_CUDA(h, cudaGetLastError());
  TXG_CUDA(h, cudaMemsetAsync(h->counters + 4, 0, sizeof(int), h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  h->tile = true;
  return 0;
}

// The C rows of the schedule live in ONE __constant__ table per device (c_lag_rows), uploaded before every launch on the
// handle's stream: two handles of one device would overwrite each other's table under a running kernel.  The first
// handle that enables the one-pass step owns the table until it is destroyed or loses eligibility; others keep the
// two-kernel step.
static txg_flow *g_lag_owner[64] = {};
static void release_lag_table(txg_flow *h) {
  if (h->device >= 0 && h->device < 64 && g_lag_owner[h->device] == h) g_lag_owner[h->device] = nullptr;
}

static int build_lag(txg_flow *h) {
  static_assert(sizeof(LagRow) == sizeof(LagRowDev), "host / device schedule row layout");
  release_lag_table(h);
  for (void **q : {(void **)&h->lag_rows_dev, (void **)&h->lag_crows_dev, (void **)&h->lag_done, (void **)&h->rtab_lag}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  h->lag = false;
  h->lag_nrows = h->lag_grid_x = 0;
  const Grid &g = h->g;
  // (several ranks: the z halos of the one-pass step run on the main stream, not overlapped with the interior yet)
  if (!h->lag_wanted || !h->fused || !h->ks.step_fused_lag || h->ks.fused_threads != 128 || h->D != 3 || h->p.eos || h->spec_n ||
      h->bc_mode)
    return 0;
  if (h->device < 0 || h->device >= 64 || (g_lag_owner[h->device] && g_lag_owner[h->device] != h)) return 0;
  const int nzE = g.NZl + 2 * g.Rz;
  std::vector<uint32_t> row_off((size_t)nzE * g.NY + 1);
  if (g.P) {
    // P[(zz*NY + y)*NX] = position of the first fluid node at or after the start of row (zz, y); P[nE] = nstore
    TXG_CUDA(h, cudaMemcpy2D(row_off.data(), sizeof(uint32_t), g.P, (size_t)g.NX * sizeof(uint32_t), sizeof(uint32_t),
                             row_off.size(), cudaMemcpyDeviceToHost));
  } else {
    for (size_t i = 0; i < row_off.size(); ++i) row_off[i] = (uint32_t)(i * (size_t)g.NX);
  }
  const int PB = 4 * h->ks.npw;
  const int MB = std::max(PB, h->lag_mpos / PB * PB);
  const LagSchedule sc = build_lag_schedule(g.NY, g.NZl, g.Rz, g.pery, row_off.data(), PB, MB, h->lag_rows, h->lag_planes, LAG_MAX_ROWS);
  if (!sc.ok || sc.rows.size() > (size_t)LAG_MAX_ROWS || sc.nbands > 16 || (unsigned long long)sc.rows.size() * sc.grid_x >= (1ull << 31)) return 0;
  TXG_CUDA(h, cudaMalloc((void **)&h->lag_rows_dev, sc.rows.size() * sizeof(LagRow)));
  TXG_CUDA(h, cudaMalloc((void **)&h->lag_done, (sc.rows.size() + 1) * sizeof(unsigned)));  // + 1: the gave-up counter
  TXG_CUDA(h, cudaMemset(h->lag_done, 0, (sc.rows.size() + 1) * sizeof(unsigned)));
  TXG_CUDA(h, cudaMemcpy(h->lag_rows_dev, sc.rows.data(), sc.rows.size() * sizeof(LagRow), cudaMemcpyHostToDevice));
  std::vector<LagCRow> crows(sc.rows.size());
  for (size_t i = 0; i < sc.rows.size(); ++i) crows[i] = LagCRow{sc.rows[i].cfirst, sc.rows[i].ccount};
  TXG_CUDA(h, cudaMalloc((void **)&h->lag_crows_dev, crows.size() * sizeof(LagCRow)));
  TXG_CUDA(h, cudaMemcpy(h->lag_crows_dev, crows.data(), crows.size() * sizeof(LagCRow), cudaMemcpyHostToDevice));
  TXG_TRY(fresh_zero(h, (void **)&h->rho_next, ((size_t)h->S * g.fs + RHO_PAD) * sizeof(double)));
  if (h->tile_wanted && h->ks.step_fused_lag_tile) {
    // density tiles in the C blocks too: window starts per block of this launch
    const long long nblk = (long long)sc.rows.size() * sc.grid_x;
    TXG_CUDA(h, cudaMalloc((void **)&h->rtab_lag, (size_t)nblk * h->ks.rtab_groups * sizeof(uint32_t)));
    h->ks.build_rtab_lag<<<blocks_for(nblk * h->ks.rtab_groups, 128), 128, 0, h->s_main>>>(
        g, h->nbr_all, PB, reinterpret_cast<const uint32_t *>(h->lag_crows_dev), (long long)sc.rows.size(), (int)sc.grid_x, h->rtab_lag);
    TXG_CUDA(h, cudaGetLastError());
    TXG_CUDA(h, cudaMemsetAsync(h->counters + 4, 0, sizeof(int), h->s_main));
  }
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  h->lag_meta.rows_per_band = sc.rows_per_band;
  h->lag_meta.lag = sc.lag;
  h->lag_meta.MB = sc.MB;
  h->lag_meta.nrows = (int)sc.rows.size();
  h->lag_meta.row_blocks = (int)sc.grid_x;
  for (int b = 0; b < 16; ++b)
    for (int k = 0; k < 3; ++k) h->lag_meta.depbands[b][k] = b < sc.nbands ? sc.depbands[b][k] : -1;
  h->lag_nrows = (unsigned)sc.rows.size();
  h->lag_grid_x = sc.grid_x;
  h->lag = true;
  g_lag_owner[h->device] = h;
  return 0;
}

// ------------------------------------------------------------------ band blocks (band_kernel.cuh)
// Bit rows, xrow and the block table of k_step_band for the current geometry, rebuilt at every walls upload.  A block
// owns up to band_lb consecutive positions of one owned plane; its three density windows (planes z-1, z, z+1) start at
// the first position of row y0 - 1 and must reach the end of row y1 + 1 (y0 .. y1 = the rows the block touches).  The
// window length `cap` is the longest such run over all blocks, bounded by what two blocks per SM can hold in shared
// memory; a stencil neighbour outside its window (the wrapped rows of a periodic y, rows too long for the budget) is
// read from global memory -- same value, so the table is a performance matter only.
static int build_band(txg_flow *h) {
  for (void **q : {(void **)&h->bitrows, (void **)&h->rowend, (void **)&h->xrow, (void **)&h->band_blocks}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  h->band = false;
  h->pull = false;
  h->state_g = false;
  const Grid &g = h->g;
  if (!h->band_wanted || !h->fused || !h->ks.step_band || g.own1 <= g.own0) return 0;
  const int rpp = g.NY + 2, nzE = g.NZl + 2 * g.Rz, NW = (g.NX + 31) / 32;
  if (g.NX > (1 << BITROW_XBITS) || (long long)nzE * rpp >= (long long)BITROW_MAX_ROWS || g.fs > (long long)BITROW_POSMASK) return 0;
  // row offsets on the host: row_off[zz * NY + y] = position of the first fluid node at or after the start of the row
  std::vector<uint32_t> row_off((size_t)nzE * g.NY + 1);
  if (g.P) {
    TXG_CUDA(h, cudaMemcpy2D(row_off.data(), sizeof(uint32_t), g.P, (size_t)g.NX * sizeof(uint32_t), sizeof(uint32_t), row_off.size(),
                             cudaMemcpyDeviceToHost));
  } else {
    for (size_t i = 0; i < row_off.size(); ++i) row_off[i] = (uint32_t)(i * (size_t)g.NX);
  }
  const int NP = h->ks.band_windows;
  // shared memory: two blocks per SM
  int dev_smem = 0;
  TXG_CUDA(h, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  const int per_block = std::min(dev_smem, (int)(((size_t)dev_smem + 1024) / (512 / h->ks.band_threads) - 2048));
  int cap_max = std::min(BAND_CAP_MAX, per_block / (NP * h->S * 8)) & ~1;
  if (cap_max < 64) return 0;
  std::vector<BandBlock> blocks;
  h->band_plane_block0.assign((size_t)g.NZl + 1, 0);
  const int LB = h->band_lb;
  uint32_t need_max = 0;
  for (int z = 0; z < g.NZl; ++z) {
    h->band_plane_block0[(size_t)z] = (int)blocks.size();
    const int zz = z + g.Rz;
    const uint32_t *ro = row_off.data() + (size_t)zz * g.NY;  // NY + 1 entries: ro[NY] = start of the next plane
    const uint32_t p0 = ro[0], p1 = ro[g.NY];
    int y = 0;
    for (uint32_t first = p0; first < p1; first += (uint32_t)LB) {
      const uint32_t count = std::min<uint32_t>((uint32_t)LB, p1 - first), last = first + count - 1;
      while (ro[y + 1] <= first) ++y;  // row of the first position (rows without fluid nodes are skipped)
      const int y0 = y;
      int y1 = y0;
      while (ro[y1 + 1] <= last) ++y1;
      BandBlock b{first, count, {0u, 0u, 0u}, {0u, 0u, 0u}};
      const int ys = std::max(y0 - 1, 0), ye = std::min(y1 + 1, g.NY - 1);
      for (int pl = 0; pl < NP; ++pl) {
        const int zp = NP == 3 ? zz + pl - 1 : zz;
        const uint32_t *rp = row_off.data() + (size_t)zp * g.NY;
        b.win[pl] = rp[ys] & ~1u;
        b.len[pl] = std::min<uint32_t>((uint32_t)cap_max, (rp[ye + 1] - b.win[pl] + 1u) & ~1u);
        need_max = std::max(need_max, b.len[pl]);
      }
      blocks.push_back(b);
    }
  }
  h->band_plane_block0[(size_t)g.NZl] = (int)blocks.size();
  if (blocks.empty()) return 0;
  const int cap = std::max(64, (int)need_max);
  h->band_smem = NP * h->S * cap * 8;
  TXG_CUDA(h, (cudaError_t)h->ks.set_band_smem(h->band_smem));
  TXG_CUDA(h, cudaMalloc((void **)&h->bitrows, (size_t)nzE * rpp * NW * sizeof(BitrowEntry)));
  TXG_CUDA(h, cudaMalloc((void **)&h->rowend, (size_t)nzE * rpp * sizeof(uint32_t)));
  TXG_CUDA(h, cudaMalloc((void **)&h->xrow, (size_t)g.fs * sizeof(uint32_t)));
  TXG_CUDA(h, cudaMemsetAsync(h->xrow, 0, (size_t)g.fs * sizeof(uint32_t), h->s_main));
  TXG_CUDA(h, cudaMalloc((void **)&h->band_blocks, blocks.size() * sizeof(BandBlock)));
  TXG_CUDA(h, cudaMemcpyAsync(h->band_blocks, blocks.data(), blocks.size() * sizeof(BandBlock), cudaMemcpyHostToDevice, h->s_main));
  k_build_bitrows<<<blocks_for((long long)nzE * rpp * NW, 128), 128, 0, h->s_main>>>(g, h->cls, NW, h->bitrows, h->rowend);
  TXG_CUDA(h, cudaGetLastError());
  k_build_xrow<<<blocks_for(h->nstore, 256), 256, 0, h->s_main>>>(g, 0, h->nstore, h->xrow);
  TXG_CUDA(h, cudaGetLastError());
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));  // (`blocks` is a local)
  h->band_params.rows = h->bitrows;
  h->band_params.rowend = h->rowend;
  h->band_params.xrow = h->xrow;
  h->band_params.blocks = h->band_blocks;
  h->band_params.NW = NW;
  h->band_params.rows_per_plane = rpp;
  h->band_params.cap = cap;
  h->band = true;
  // pull form: plain bounce-back walls only (free-slip tables and face BCs work on pushed populations)
  h->pull = h->pull_wanted && h->ks.step_band_pull && !h->spec_n && !h->bc_mode && !h->lag && !h->tile;
  if (h->pull) TXG_CUDA(h, (cudaError_t)h->ks.set_band_pull_smem(h->band_smem));
  return 0;
}

// ------------------------------------------------------------------ staged K2 (stage_kernel.cuh, opt-in)
// Tensor descriptors of the arrays k_step_stage streams: [S * Q][fs] doubles (both population buffers) and [Q][fs] words
// (adjacency rows + mask row, copied into one array), box = one warp item.  cuTensorMapEncodeTiled is a driver entry
// point; it is fetched through the runtime so that the library keeps linking against cudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int build_stage_tensors(txg_flow *h) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      h->stage = false;  // (an old driver: keep the table kernel)
      return 0;
    }
    encode = (EncodeTiledFn)fn;
  }
  const Grid &g = h->g;
  const size_t row = (size_t)g.fs * sizeof(uint32_t);
  TXG_CUDA(h, cudaMalloc((void **)&h->adjm, (size_t)h->Q * row));
  TXG_CUDA(h, cudaMemcpyAsync(h->adjm, h->nbr_all, (size_t)(h->Q - 1) * row, cudaMemcpyDeviceToDevice, h->s_main));
  TXG_CUDA(h, cudaMemcpyAsync(h->adjm + (size_t)(h->Q - 1) * g.fs, h->lmask, row, cudaMemcpyDeviceToDevice, h->s_main));
  auto make = [&](CUtensorMap *tm, CUtensorMapDataType type, void *base, int esize, int rows) -> bool {
    const cuuint64_t dims[2] = {(cuuint64_t)g.fs, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)g.fs * (cuuint64_t)esize};
    const cuuint32_t box[2] = {(cuuint32_t)h->ks.stage_item, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return encode(tm, type, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!make(&h->tm_f[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, h->f[0], 8, h->ks.stage_rows_f) ||
      !make(&h->tm_f[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, h->f[1], 8, h->ks.stage_rows_f) ||
      !make(&h->tm_adj, CU_TENSOR_MAP_DATA_TYPE_UINT32, h->adjm, 4, h->ks.stage_rows_a))
    TXG_FAIL(h, TXG_ERR_LIB, "cuTensorMapEncodeTiled failed for the staged kernel (fs = %lld)", (long long)g.fs);
  return 0;
}

extern "C" int txg_set_walls(txg_handle h, const double *walls_rg) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!walls_rg) TXG_FAIL(h, TXG_ERR_ARG_NULL, "txg_set_walls: null array");
  drop_step_graphs(h);
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  const long long n = (long long)(g.NZl + 2 * g.Rz) * g.cny * g.cnx;
  // classify in chunks through the staging buffer
  const long long chunk = std::min<long long>(n, (long long)32 << 20);
  TXG_TRY(ensure_staging(h, (size_t)chunk * 8));
  TXG_CUDA(h, cudaMemsetAsync(h->counters, 0, 4 * sizeof(int), h->s_main));
  for (long long o = 0; o < n; o += chunk) {
    const long long c = std::min(chunk, n - o);
    TXG_CUDA(h, cudaMemcpyAsync(h->staging, walls_rg + o, (size_t)c * 8, cudaMemcpyHostToDevice, h->s_main));
    k_classify<<<blocks_for(c, 256), 256, 0, h->s_main>>>(h->staging, h->cls + o, c, h->counters);
    TXG_CUDA(h, cudaGetLastError());
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  }
  h->ks.build_masks<<<blocks_for(g.nnodes, 128), 128, 0, h->s_main>>>(g, h->cls, h->nbmask, h->ffmask, h->counters + 1);
  TXG_CUDA(h, cudaGetLastError());
  int counters[4];
  TXG_CUDA(h, cudaMemcpyAsync(counters, h->counters, sizeof counters, cudaMemcpyDeviceToHost, h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  if (counters[0]) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "walls array holds %d negative or NaN codes", counters[0]);
  TXG_TRY(build_storage(h));
  // free-slip walls anywhere in the run?  (every rank must join the halo of the parked populations or none)
  h->spec_any = counters[1] > 0;
  if (h->cfg.nranks > 1) {
    if (!h->comm) {
      if (counters[1]) TXG_FAIL(h, TXG_ERR_ORDER, "free-slip walls (900-902) on %d ranks: call txg_comm_init before txg_set_walls", h->cfg.nranks);
    } else {
      int *flag = h->counters + 6;
      const int mine = counters[1] > 0 ? 1 : 0;
      TXG_CUDA(h, cudaMemcpyAsync(flag, &mine, sizeof mine, cudaMemcpyHostToDevice, h->s_main));
      TXG_NCCL(h, g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMax, h->comm, h->s_main));
      int any = 0;
      TXG_CUDA(h, cudaMemcpyAsync(&any, flag, sizeof any, cudaMemcpyDeviceToHost, h->s_main));
      TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
      h->spec_any = any != 0;
    }
  }
  TXG_TRY(build_specular(h, counters[1]));
  TXG_TRY(build_lag(h));
  TXG_TRY(build_rtab(h));
  TXG_TRY(build_band(h));
  // nodes of BC_REFLECTING faces keep the density from before BCApply (MASK_STALE, kernels.cuh): mark them, then unmark
  // the nodes the faces of BCUpdateRho (Dirichlet / Neumann / velocity) share with them
  if (h->has_reflecting)
    for (int pass = 1; pass >= 0; --pass)
      for (int b = 0; b < 2 * h->D; ++b) {
        const FaceDesc &fd = h->faces[b];
        if (!h->face_here[b] || (pass ? fd.type != TXG_BC_REFLECTING : fd.type < TXG_BC_DIRICHLET)) continue;
        k_bc_mark_stale<<<blocks_for((long long)fd.n1 * fd.n2, 128), 128, 0, h->s_main>>>(h->g, fd, h->nbmask, h->lmask, pass);
        TXG_CUDA(h, cudaGetLastError());
        h->launches++;
      }
  // staged form of K2: whenever the fused kernel applies and no opt-in experiment replaces it
  // (S = 3: 10 positions per item; its word rows start off 16-byte boundaries and the copies never completed on the device)
  h->stage = h->stage_wanted && h->fused && h->ks.step_stage[0] && !h->tile && !h->band && !h->lag && h->ks.npw % 4 == 0;
  if (h->adjm) cudaFree(h->adjm);
  h->adjm = nullptr;
  if (h->stage) TXG_TRY(build_stage_tensors(h));
  if (h->stage) {
    h->stage_v = 0;
    if (const char *v = getenv("TXG_STAGE_WARPS")) {
      const int w = atoi(v);
      for (int i = 0; i < 3; ++i)
        if (h->ks.stage_warps[i] == w) h->stage_v = i;
    }
    // shared-memory carve-out: what the resident blocks of this shape need (+ 1 KB per block the driver reserves, + the static
    // barriers), the rest of the 256 KB stays L1 for the density gathers
    int carve = (h->ks.stage_blocks_v[h->stage_v] * (h->ks.stage_smem_v[h->stage_v] + 2048) * 100 + 228 * 1024 - 1) / (228 * 1024);
    carve = std::min(100, std::max(carve, 1));
    if (const char *v = getenv("TXG_STAGE_CARVE")) carve = atoi(v);
    TXG_CUDA(h, (cudaError_t)h->ks.set_stage_attrs(h->stage_v, carve));
    int rounds = 2;
    if (const char *v = getenv("TXG_STAGE_ROUNDS")) rounds = std::max(1, atoi(v));
    h->stage_lb = rounds * h->ks.stage_warps[h->stage_v] * h->ks.npw;
    h->stage_pf = 0;
    if (const char *v = getenv("TXG_STAGE_PF")) h->stage_pf = atoi(v);  // > 0: tensor prefetch, < 0: plain prefetch lines
    h->stage_clc = false;
    if (const char *v = getenv("TXG_STAGE_CLC")) h->stage_clc = v[0] != '0';
    h->stage_pg = 0;
    if (const char *v = getenv("TXG_STAGE_PG")) h->stage_pg = atoi(v);
    h->stage_adjc = false;
    if (const char *v = getenv("TXG_STAGE_ADJC")) h->stage_adjc = v[0] != '0';
    h->stage_adjc = h->stage_adjc && h->stage_v == 0 && !h->stage_clc;
    if (h->adjc) cudaFree(h->adjc);
    h->adjc = nullptr;
    if (h->stage_adjc) {
      // one record per item of ITEM positions; items are aligned on absolute multiples of ITEM (= the warp items of the kernel)
      const long long item0 = h->g.own0 / h->ks.stage_item, item1 = (h->g.own1 + h->ks.stage_item - 1) / h->ks.stage_item;
      TXG_CUDA(h, cudaMalloc((void **)&h->adjc, (size_t)(h->g.fs / h->ks.stage_item + 1) * h->ks.adjc_rec_bytes));
      if (item1 > item0) {
        h->ks.build_adjc<<<blocks_for(item1 - item0, 128), 128, 0, h->s_main>>>(h->g, h->nbr_all, h->lmask, h->adjc, item0, item1 - item0);
        TXG_CUDA(h, cudaGetLastError());
      }
      TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
    }
  }
  h->walls_set = true;
  return 0;
}

extern "C" int txg_get_node_class(txg_handle h, uint8_t *out) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!out) TXG_FAIL(h, TXG_ERR_ARG_NULL, "null array");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  TXG_CUDA(h, cudaMemcpy(out, h->cls, (size_t)(g.NZl + 2 * g.Rz) * g.cny * g.cnx, cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------ state in
// blocks of 128 threads = 4 warps of npw (fluid node, all components) items
static inline unsigned hot_blocks(const txg_flow *h, long long count) {
  const long long warps = (count + h->ks.npw - 1) / h->ks.npw;
  return (unsigned)((warps + 3) / 4);
}
static int run_moments(txg_flow *h, int z0, int nz, cudaStream_t s) {
  if (nz <= 0) return 0;
  long long first, count;
  plane_range(h, z0, nz, &first, &count);
  if (count == 0) return 0;
  if (h->state_g) {  // collided populations in the buffer: gather, then sum
    ScopedKernel sk(h, "k_moments_pull", s);
    h->ks.moments_pull<<<hot_blocks(h, count), 128, 0, s>>>(h->g, h->p, h->band_params, h->f[h->cur], nullptr, h->rho, h->rho_true, first, count);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  ScopedKernel sk(h, "k_moments", s);
  h->ks.moments<<<hot_blocks(h, count), 128, 0, s>>>(h->g, h->p, h->f[h->cur], h->rho, h->rho_true, first, count, count, 0);
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}
// the planes [za, za + nz) and [zb, zb + nz) (za + nz <= zb) in ONE launch: the boundary planes of a slab are a few hundred
// blocks each, and two launches of that size cost more in launch gaps and tails than in work
static int run_moments_pair(txg_flow *h, int za, int zb, int nz, cudaStream_t s) {
  long long fa, ca, fb, cb;
  plane_range(h, za, nz, &fa, &ca);
  plane_range(h, zb, nz, &fb, &cb);
  if (h->state_g || ca == 0 || cb == 0 || za + nz > zb) {
    TXG_TRY(run_moments(h, za, nz, s));
    return run_moments(h, zb, nz, s);
  }
  ScopedKernel sk(h, "k_moments", s);
  h->ks.moments_pair<<<hot_blocks(h, ca + cb), 128, 0, s>>>(h->g, h->p, h->f[h->cur], h->rho, h->rho_true, fa, ca + cb, ca, fb - (fa + ca));
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}

// The reference's fi (streamed, bounced-back populations) in the current buffer: a no-op unless the pull form left the
// collided populations there.  Every path that reads or hands out fi calls this first.
static int materialise(txg_flow *h) {
  if (!h->state_g) return 0;
  const Grid &g = h->g;
  const long long count = g.own1 - g.own0;
  if (count > 0) {
    ScopedKernel sk(h, "k_pull_stream", h->s_main);
    h->ks.pull_stream<<<hot_blocks(h, count), 128, 0, h->s_main>>>(g, h->p, h->band_params, h->f[h->cur], h->f[h->cur ^ 1], nullptr, nullptr, g.own0, count);
    TXG_CUDA(h, cudaGetLastError());
  }
  h->cur ^= 1;
  h->state_g = false;
  return 0;
}
static int run_forces(txg_flow *h, int z0, int nz, cudaStream_t s) {
  if (nz <= 0) return 0;
  long long first, count;
  plane_range(h, z0, nz, &first, &count);
  if (count == 0) return 0;
  if (h->fused || h->wide_fused) return 0;  // the collide launch forms the forces itself
  if (h->ks.forces_tile && h->forces_tile_on) {  // wide stencils: psi staged as dense tiles in shared memory
    ScopedKernel sk(h, "k_forces_tile", s);
    // a block marches over zc planes (ring of psi planes in shared memory): long marches amortise the ring's prologue,
    // short ones keep enough blocks for a thin slab
    const int zc = h->D == 3 ? std::max(1, std::min(h->tile_march, nz)) : 1;
    const dim3 grid((unsigned)((h->g.NX + h->ks.forces_tile_tx - 1) / h->ks.forces_tile_tx),
                    (unsigned)((h->g.NY + h->ks.forces_tile_ty - 1) / h->ks.forces_tile_ty), (unsigned)((nz + zc - 1) / zc));
    h->ks.forces_tile<<<grid, 256, (size_t)h->ks.forces_tile_smem, s>>>(h->g, h->p, h->rho, h->rho_true, h->lmask, h->ffmask, h->wallrec, h->Fbuf, z0,
                                                                        nz, zc);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  ScopedKernel sk(h, "k_forces", s);
  h->ks.forces<<<hot_blocks(h, count), 128, 0, s>>>(h->g, h->p, h->rho, h->rho_true, h->lmask, h->nbr, h->ffmask, h->wallrec,
                                                     h->Fbuf, first, count, FaceDesc(), nullptr);
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}
static int run_collide(txg_flow *h, int z0, int nz, cudaStream_t s) {
  if (nz <= 0) return 0;
  long long first, count;
  plane_range(h, z0, nz, &first, &count);
  if (count == 0) return 0;
  if (h->fused && h->band && !h->tile) {
    ScopedKernel sk(h, h->pull ? "k_step_band_pull" : "k_step_band", s);
    const int b0 = h->band_plane_block0[(size_t)z0], b1 = h->band_plane_block0[(size_t)(z0 + nz)];
    if (h->pull)
      h->ks.step_band_pull<<<(unsigned)(b1 - b0), h->ks.band_threads, (size_t)h->band_smem, s>>>(h->g, h->p, h->band_params, h->f[h->cur], h->f[h->cur ^ 1],
                                                                                               h->rho, h->wallrec, b0, h->band_prefetch, h->state_g ? 0 : 1,
                                                                                               h->counters + 4);
    else
      h->ks.step_band<<<(unsigned)(b1 - b0), h->ks.band_threads, (size_t)h->band_smem, s>>>(h->g, h->p, h->band_params, h->f[h->cur], h->f[h->cur ^ 1], h->rho,
                                                                                          h->wallrec, b0, h->band_prefetch, 0, h->counters + 4);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  if (h->fused && h->stage && !h->tile) {
    // short blocks: TXG_STAGE_ROUNDS (default 2) rounds of the block's warps, aligned on absolute multiples of their length
    const long long LB = h->stage_lb, blk0 = first / LB, nblk = (first + count - 1) / LB - blk0 + 1;
    if (h->stage_clc) {
      ScopedKernel sk(h, "k_step_stage_clc", s);
      h->ks.step_stage_clc[h->stage_v]<<<(unsigned)nblk, 32 * h->ks.stage_warps[h->stage_v], (size_t)h->ks.stage_smem_v[h->stage_v], s>>>(h->g, h->p, h->tm_f[h->cur], h->tm_adj, h->f[h->cur ^ 1],
                                                                                                 h->rho, h->wallrec, first, count, blk0, (int)LB, h->stage_pg);
      TXG_CUDA(h, cudaGetLastError());
      return 0;
    }
    ScopedKernel sk(h, "k_step_stage", s);
    if (h->stage_adjc)
      h->ks.step_stage_adjc<<<(unsigned)nblk, 32 * h->ks.stage_warps[0], (size_t)h->ks.stage_smem_adjc, s>>>(
          h->g, h->p, h->tm_f[h->cur], h->tm_adj, h->f[h->cur ^ 1], h->rho, h->wallrec, first, count, blk0, (int)LB, 0, h->f[h->cur], h->adjm, h->adjc,
          h->nbr_all, h->stage_pg);
    else
      h->ks.step_stage[h->stage_v]<<<(unsigned)nblk, 32 * h->ks.stage_warps[h->stage_v], (size_t)h->ks.stage_smem_v[h->stage_v], s>>>(
          h->g, h->p, h->tm_f[h->cur], h->tm_adj, h->f[h->cur ^ 1], h->rho, h->wallrec, first, count, blk0, (int)LB, h->stage_pf, h->f[h->cur], h->adjm,
          nullptr, nullptr, h->stage_pg);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  if (h->fused && h->tile && first == h->g.own0) {  // (window starts are per block counted from own0: whole-slab launches only)
    ScopedKernel sk(h, "k_step_fused_tile", s);
    const long long warps = (count + h->ks.npw - 1) / h->ks.npw;
    h->ks.step_fused_tile<<<(unsigned)((warps + 3) / 4), 128, 0, s>>>(h->g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->rho, h->lmask,
                                                                     h->nbr_all, h->wallrec, h->rtab, h->counters + 4, first, count,
                                                                     h->pf_blocks);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  if (h->fused) {
    ScopedKernel sk(h, "k_step_fused", s);
    const int wpb = h->ks.fused_threads / 32;  // warps per block
    const long long warps = (count + h->ks.npw - 1) / h->ks.npw;
    h->ks.step_fused<<<(unsigned)((warps + wpb - 1) / wpb), h->ks.fused_threads, 0, s>>>(h->g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->rho, h->lmask, h->nbr_all,
                                                           h->wallrec, first, count,
                                                           h->ks.fused_threads == 128 ? h->pf_blocks : 0, count, 0);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  if (h->wide_fused) {
    ScopedKernel sk(h, "k_step_tile", s);
    const int zc = h->D == 3 ? std::max(1, std::min(h->tile_march, nz)) : 1;
    const dim3 grid((unsigned)((h->g.NX + h->ks.forces_tile_tx - 1) / h->ks.forces_tile_tx),
                    (unsigned)((h->g.NY + h->ks.forces_tile_ty - 1) / h->ks.forces_tile_ty), (unsigned)((nz + zc - 1) / zc));
    h->ks.step_tile<<<grid, 256, (size_t)h->ks.forces_tile_smem, s>>>(h->g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->rho, h->rho_true, h->lmask, h->nbr,
                                                                      h->ffmask, h->wallrec, z0, nz, zc);
    TXG_CUDA(h, cudaGetLastError());
    return 0;
  }
  ScopedKernel sk(h, "k_collide", s);
  h->ks.collide<<<hot_blocks(h, count), 128, 0, s>>>(h->g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->Fbuf, h->lmask, h->nbr,
                                                      first, count, h->has_reflecting ? h->rho_true : nullptr, FaceDesc(), nullptr);
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}
// planes [za, za + 1) and [zb, zb + 1) in one launch of the default K2 (other forms: two launches)
static int run_collide_pair(txg_flow *h, int za, int zb, cudaStream_t s) {
  long long fa, ca, fb, cb;
  plane_range(h, za, 1, &fa, &ca);
  plane_range(h, zb, 1, &fb, &cb);
  // (the staged form too: both kernels compute the same bits, and one table-kernel launch over two single planes beats two launches)
  if (!h->fused || h->band || h->tile || !h->ks.step_fused_pair || ca == 0 || cb == 0 || za >= zb) {
    TXG_TRY(run_collide(h, za, 1, s));
    return run_collide(h, zb, 1, s);
  }
  ScopedKernel sk(h, "k_step_fused", s);
  const int wpb = h->ks.fused_threads / 32;
  const long long warps = (ca + cb + h->ks.npw - 1) / h->ks.npw;
  h->ks.step_fused_pair<<<(unsigned)((warps + wpb - 1) / wpb), h->ks.fused_threads, 0, s>>>(h->g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->rho, h->lmask,
                                                                                          h->nbr_all, h->wallrec, fa, ca + cb, 0, ca, fb - (fa + ca));
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}

// the populations of the owned nodes are in f[cur]: with push storage that is the whole state
static int state_ready(txg_flow *h) {
  h->state_set = true;
  h->state_g = false;  // the buffer holds streamed populations
  h->rho_current = false;
  h->forces_current = false;
  return 0;
}

extern "C" int txg_set_rho_u(txg_handle h, const double *rho_rg, const double *u_g) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!rho_rg) TXG_FAIL(h, TXG_ERR_ARG_NULL, "txg_set_rho_u: null rho");
  if (!h->walls_set) TXG_FAIL(h, TXG_ERR_ORDER, "txg_set_rho_u before txg_set_walls (the walls size the device storage)");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  TXG_TRY(import_field(h, rho_rg, g.R, g.Rz, 1, h->rho_true, 0));
  if (u_g) {
    if (!h->u0) TXG_CUDA(h, cudaMalloc((void **)&h->u0, (size_t)h->S * h->D * g.nnodes * sizeof(double)));
    TXG_TRY(import_field(h, u_g, 1, h->D == 3 ? 1 : 0, h->D, h->u0, 1));
  } else if (h->u0) {
    cudaFree(h->u0);
    h->u0 = nullptr;
  }
  return 0;
}

extern "C" int txg_set_fi(txg_handle h, const double *fi_g) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!fi_g) TXG_FAIL(h, TXG_ERR_ARG_NULL, "txg_set_fi: null fi");
  if (!h->walls_set) TXG_FAIL(h, TXG_ERR_ORDER, "txg_set_fi before txg_set_walls");
  TXG_CUDA(h, cudaSetDevice(h->device));
  TXG_TRY(import_field(h, fi_g, 1, h->D == 3 ? 1 : 0, h->Q, h->f[h->cur], 0));
  return state_ready(h);
}

// psi = EOS(rho) over the owned planes of rho_true -> rho (stencil field); identity without an EOS
__global__ void k_eos_field(Grid g, Phys p, int S, const double *__restrict__ rho_true, double *__restrict__ psi) {
  const long long pos = g.own0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= g.own1) return;
  for (int m = 0; m < S; ++m) {
    const long long o = m * g.fs + pos;
    psi[o] = eos_psi(p, m, rho_true[o]);
  }
}

extern "C" int txg_fi_init(txg_handle h) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!h->walls_set) TXG_FAIL(h, TXG_ERR_ORDER, "txg_fi_init before txg_set_walls");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  if (h->cfg.use_nonideal_eos) {
    k_eos_field<<<blocks_for(std::max<long long>(g.own1 - g.own0, 1), 256), 256, 0, h->s_main>>>(g, h->p, h->S, h->rho_true, h->rho);
    TXG_CUDA(h, cudaGetLastError());
    h->launches++;
  }
  TXG_TRY(exchange_rho(h, h->rho, h->s_main));
  if (h->fused) {
    const long long nown = g.own1 - g.own0;
    if (nown) {
      ScopedKernel sk(h, "k_fi_init_fused", h->s_main);
      h->ks.fi_init_fused<<<hot_blocks(h, nown), 128, 0, h->s_main>>>(g, h->p, h->f[h->cur], h->rho, h->rho_true, h->u0, h->lmask,
                                                                      h->nbr_all, h->wallrec, g.own0, nown);
      TXG_CUDA(h, cudaGetLastError());
    }
  } else {
    ScopedKernel sk(h, "k_fi_init", h->s_main);
    h->ks.fi_init<<<blocks_for(g.nnodes, 128), 128, 0, h->s_main>>>(g, h->p, h->f[h->cur], h->rho, h->rho_true, h->u0,
                                                                     h->nbmask, h->ffmask, h->cls, 0, g.NZl);
    TXG_CUDA(h, cudaGetLastError());
  }
  return state_ready(h);
}

// ------------------------------------------------------------------ the step
// One reference time step = collide, communicate fi, stream, bounce-back, density, forces, flux,
// common velocity (lbm.F90:286-361).  On the device: K1 moments (density) with the rho halo, K2a
// forces (every gather of the step), then K2b (momentum + velocity + collision + push-streaming with
// bounce-back) with the halo of the pushed populations.  Boundary planes run first so that their halos travel on the communication
// stream while the interior computes.
// One step as ONE hot launch (opt-in): collide + push every owned plane and, behind the collision front, sum the
// new densities of planes 1 .. NZl-2 out of L2; the two boundary planes are summed by k_moments once the z halo has
// delivered their crossing populations.  On entry rho (+ halo) belongs to f[cur]; on exit again.
static int one_step_lag(txg_flow *h) {
  const Grid &g = h->g;
  cudaStream_t sm = h->s_main;
  if (!h->rho_current) {
    TXG_TRY(run_moments(h, 0, g.NZl, sm));
    TXG_TRY(exchange_rho(h, h->rho, sm));
  }
  TXG_CUDA(h, cudaMemsetAsync(h->lag_done, 0, (size_t)h->lag_nrows * sizeof(unsigned), sm));
  // the schedule rows live in a constant table of the kernel's module: another handle may have used it last
  TXG_CUDA(h, (cudaError_t)h->ks.upload_lag_rows(h->lag_crows_dev, (size_t)h->lag_nrows * sizeof(LagCRow), sm));
  {
    ScopedKernel sk(h, "k_step_fused_lag", sm);
    (h->rtab_lag ? h->ks.step_fused_lag_tile : h->ks.step_fused_lag)<<<h->lag_grid_x * h->lag_nrows, 128, 0, sm>>>(g, h->p, h->lag_meta, h->f[h->cur], h->f[h->cur ^ 1], h->rho,
                                                                             h->rho_next, h->lmask, h->nbr_all, h->wallrec, h->lag_rows_dev,
                                                                             h->lag_done, h->lag_done + h->lag_nrows, h->rtab_lag, h->counters + 4, h->pf_blocks);
    TXG_CUDA(h, cudaGetLastError());
  }
  TXG_TRY(exchange_f(h, h->f[h->cur ^ 1], sm));
  h->cur ^= 1;
  std::swap(h->rho, h->rho_next);
  h->rho_true = h->rho;  // no EOS on this path
  TXG_TRY(run_moments(h, 0, 1, sm));
  TXG_TRY(run_moments(h, g.NZl - 1, 1, sm));
  TXG_TRY(exchange_rho(h, h->rho, sm));
  h->rho_current = true;
  return 0;
}

static int one_step(txg_flow *h) {
  const Grid &g = h->g;
  // (free-slip walls: their slots are rewritten after the whole push and its halos, on one stream)
  const bool split = h->cfg.nranks > 1 && g.NZl >= 4 * g.R + 2 && !h->spec_any;
  cudaStream_t sm = h->s_main, sc = h->s_comm;
  if (!split) {
    TXG_TRY(run_moments(h, 0, g.NZl, sm));
    TXG_TRY(exchange_rho(h, h->rho, sm));
    TXG_TRY(run_forces(h, 0, g.NZl, sm));
    TXG_TRY(run_collide(h, 0, g.NZl, sm));
    if (h->pull) {
      TXG_TRY(exchange_g(h, h->f[h->cur ^ 1], sm));
    } else {
      TXG_TRY(exchange_f(h, h->f[h->cur ^ 1], sm));
      TXG_TRY(apply_specular(h, h->f[h->cur ^ 1], sm));
    }
    h->cur ^= 1;
    h->state_g = h->pull;
    return 0;
  }
  const int R = g.R;
  // K1: bottom and top R planes, then their halo on the comm stream; interior K1 and the interior
  // forces (which need no halo) meanwhile.  (Running K1 and the forces of different sub-slabs on two
  // streams was tried and gains nothing: a later grid only gets SMs at the tail of an earlier one.)
  TXG_TRY(run_moments_pair(h, 0, g.NZl - R, R, sm));
  TXG_CUDA(h, cudaEventRecord(h->ev_a, sm));
  TXG_CUDA(h, cudaStreamWaitEvent(sc, h->ev_a, 0));
  TXG_TRY(exchange_rho(h, h->rho, sc));
  TXG_CUDA(h, cudaEventRecord(h->ev_b, sc));
  TXG_TRY(run_moments(h, R, g.NZl - 2 * R, sm));
  TXG_TRY(run_forces(h, R, g.NZl - 2 * R, sm));
  TXG_CUDA(h, cudaStreamWaitEvent(sm, h->ev_b, 0));
  // boundary planes: forces (they need the rho halo), collide, halo of the new populations on the
  // comm stream; interior collide meanwhile
  TXG_TRY(run_forces(h, 0, R, sm));
  TXG_TRY(run_forces(h, g.NZl - R, R, sm));
  TXG_TRY(run_collide_pair(h, 0, g.NZl - 1, sm));
  TXG_CUDA(h, cudaEventRecord(h->ev_a, sm));
  TXG_CUDA(h, cudaStreamWaitEvent(sc, h->ev_a, 0));
  if (h->pull)
    TXG_TRY(exchange_g(h, h->f[h->cur ^ 1], sc));
  else
    TXG_TRY(exchange_f(h, h->f[h->cur ^ 1], sc));
  TXG_CUDA(h, cudaEventRecord(h->ev_b, sc));
  TXG_TRY(run_collide(h, 1, g.NZl - 2, sm));
  TXG_CUDA(h, cudaStreamWaitEvent(sm, h->ev_b, 0));
  h->cur ^= 1;
  h->state_g = h->pull;
  return 0;
}

// ------------------------------------------------------------------ the step with external face BCs
// FlowCalcRhoForces (lbm_flow.F90:445-456): density, BCApplyDirichletToRho, density halo, forces -> Fbuf.
// `dirichlet` = false is FlowUpdateMoments (:466-478), which forms the same moments without the override.
static int bc_moments_forces(txg_flow *h, bool dirichlet) {
  const Grid &g = h->g;
  cudaStream_t sm = h->s_main;
  TXG_TRY(run_moments(h, 0, g.NZl, sm));
  if (dirichlet)
    for (int b = 0; b < 2 * h->D; ++b) {
      const FaceDesc &fd = h->faces[b];
      if (fd.type != TXG_BC_DIRICHLET || !h->face_here[b]) continue;
      if (!h->bc_vals[b]) TXG_FAIL(h, TXG_ERR_ORDER, "boundary %d is BC_DIRICHLET but txg_set_bc_values was not called for it", b);
      ScopedKernel sk(h, "k_bc_dirichlet_rho", sm);
      k_bc_dirichlet_rho<<<blocks_for((long long)fd.n1 * fd.n2, 128), 128, 0, sm>>>(g, h->p, h->S, fd, h->bc_vals[b], h->S * h->D, h->rho,
                                                                                  h->rho_true, h->nbmask);
      TXG_CUDA(h, cudaGetLastError());
    }
  TXG_TRY(exchange_rho(h, h->rho, sm));
  if (h->bc_fused) {
    // only BCApply and the second collision of the face nodes read stored forces: form them on the BC faces alone
    for (int b = 0; b < 2 * h->D; ++b) {
      const FaceDesc &fd = h->faces[b];
      if (!h->face_here[b] || fd.type < TXG_BC_REFLECTING) continue;
      const long long n = (long long)fd.n1 * fd.n2;
      ScopedKernel sk(h, "k_forces_face", sm);
      h->ks.forces_face<<<hot_blocks(h, n), 128, 0, sm>>>(g, h->p, h->rho, h->rho_true, h->lmask, h->nbr, h->ffmask, h->wallrec, h->Fbuf, 0, n, fd,
                                                          h->nbmask);
      TXG_CUDA(h, cudaGetLastError());
    }
  } else {
    TXG_TRY(run_forces(h, 0, g.NZl, sm));
  }
  h->forces_current = true;
  return 0;
}

// BCApply (lbm_bc.F90:781-807) on the populations of f[cur]; BCUpdateRho (:436-611) needs no launch:
// every consumer of the density of a face node (collision, export) sums the populations itself.
static int bc_apply(txg_flow *h) {
  const Grid &g = h->g;
  cudaStream_t sm = h->s_main;
  for (int b : h->bc_order) {
    const FaceDesc &fd = h->faces[b];
    if (!h->face_here[b]) continue;
    if (fd.type == TXG_BC_REFLECTING) {
      ScopedKernel sk(h, "k_bc_reflect", sm);
      k_bc_reflect<<<blocks_for((long long)fd.n1 * fd.n2, 128), 128, 0, sm>>>(g, h->Q, h->S, fd, h->reflect[b], h->f[h->cur], h->nbmask);
      TXG_CUDA(h, cudaGetLastError());
      continue;
    }
    if (!h->bc_vals[b])
      TXG_FAIL(h, TXG_ERR_ORDER, "boundary %d has bc_flags = %d but txg_set_bc_values was not called for it", b, fd.type);
    ScopedKernel sk(h, "k_bc_apply", sm);
    k_bc_apply<<<blocks_for((long long)fd.n1 * fd.n2, 128), 128, 0, sm>>>(g, h->lt, h->S, fd, h->bc_vals[b], h->f[h->cur], h->Fbuf, h->nbmask);
    TXG_CUDA(h, cudaGetLastError());
  }
  return 0;
}

// One time step in the reference's own order (lbm.F90:286-361): FlowCollision + communicate + stream +
// bounce-back (k_collide with the z halo and the free-slip slots), then FlowApplyBCs.  Fbuf carries the
// forces from one step's FlowApplyBCs to the next step's collision, as flow%forces does.
static int one_step_bc(txg_flow *h) {
  const Grid &g = h->g;
  cudaStream_t sm = h->s_main;
  // FlowUpdateBCPressureOutlet (top of FlowApplyBCs, only when ncomponents /= 1, lbm_flow.F90:1965-1972):
  // it reads the densities the previous step left, i.e. the sums of the populations still in f[cur]
  if (h->S == 2)
    for (int b = 0; b < 2 * h->D; ++b) {
      if (!h->pressure_outlet[b] || !h->face_here[b]) continue;
      const FaceDesc &fd = h->faces[b];
      if (!h->bc_vals[b]) TXG_FAIL(h, TXG_ERR_ORDER, "boundary %d is a pressure outlet but txg_set_bc_values was not called for it", b);
      ScopedKernel sk(h, "k_bc_pressure_outlet", sm);
      k_bc_pressure_outlet<<<blocks_for((long long)fd.n1 * fd.n2, 128), 128, 0, sm>>>(g, h->Q, fd, h->bc_vals[b], h->S * h->D, h->f[h->cur],
                                                                                    h->nbmask, h->outlet_pressure[b], h->cfg.gf[1][0]);
      TXG_CUDA(h, cudaGetLastError());
    }
  if (h->bc_fused) {
    long long first, count;
    plane_range(h, 0, g.NZl, &first, &count);
    if (count) {
      ScopedKernel sk(h, "k_step_fused", sm);
      const int wpb = h->ks.fused_threads / 32;
      const long long warps = (count + h->ks.npw - 1) / h->ks.npw;
      h->ks.step_fused<<<(unsigned)((warps + wpb - 1) / wpb), h->ks.fused_threads, 0, sm>>>(h->g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->rho, h->lmask,
                                                                                          h->nbr_all, h->wallrec, first, count,
                                                                                          h->ks.fused_threads == 128 ? h->pf_blocks : 0, count, 0);
      TXG_CUDA(h, cudaGetLastError());
    }
    for (int b = 0; b < 2 * h->D; ++b) {
      const FaceDesc &fd = h->faces[b];
      if (!h->face_here[b] || fd.type < TXG_BC_REFLECTING) continue;
      const long long n = (long long)fd.n1 * fd.n2;
      ScopedKernel sk(h, "k_collide_face", sm);
      h->ks.collide_face<<<hot_blocks(h, n), 128, 0, sm>>>(g, h->p, h->f[h->cur], h->f[h->cur ^ 1], h->Fbuf, h->lmask, h->nbr, 0, n,
                                                           h->has_reflecting ? h->rho_true : nullptr, fd, h->nbmask);
      TXG_CUDA(h, cudaGetLastError());
    }
  } else {
    TXG_TRY(run_collide(h, 0, g.NZl, sm));
  }
  TXG_TRY(exchange_f(h, h->f[h->cur ^ 1], sm));
  TXG_TRY(apply_specular(h, h->f[h->cur ^ 1], sm));
  h->cur ^= 1;
  TXG_TRY(bc_moments_forces(h, true));
  TXG_TRY(bc_apply(h));
  return 0;
}

// FlowParseBC's BC_PRESSURE_OUTLET (lbm_flow.F90:1170-1189) with the checks of FlowUpdateBCPressureOutlet (:2000-2005)
// and FlowUpdateDensityFromPressure (:2260)
extern "C" int txg_set_bc_pressure_outlet(txg_handle h, int boundary, double pressure) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (boundary < 0 || boundary >= 2 * h->D) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "boundary %d out of range", boundary);
  if (h->cfg.bc_flags[boundary] != TXG_BC_DIRICHLET)
    TXG_FAIL(h, TXG_ERR_ARG_WRONG, "a pressure outlet is a BC_DIRICHLET face (lbm_flow.F90:1178): bc_flags[%d] = %d", boundary, h->cfg.bc_flags[boundary]);
  if (h->S > 2) TXG_FAIL(h, 1, "Invalid number of components for flow");
  if (h->S == 2 && (fabs(h->cfg.gf[0][0]) > (double)1.e-10f || h->cfg.use_nonideal_eos))
    TXG_FAIL(h, 1, "Pressure outlet not implemented for non-ideal EOS or g11 or g22 /= 0");
  if (boundary >= 4 && h->g.NZl < 2) TXG_FAIL(h, TXG_ERR_SUP, "a z pressure outlet needs a slab of at least two planes");
  h->pressure_outlet[boundary] = true;
  h->outlet_pressure[boundary] = pressure;
  return 0;
}

// BCSetValues (lbm_bc.F90:215-228)
extern "C" int txg_set_bc_values(txg_handle h, int boundary, const double *vals) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (boundary < 0 || boundary >= 2 * h->D) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "boundary %d out of range", boundary);
  if (!h->face_here[boundary]) return 0;  // a z face of another slab
  if (!vals) TXG_FAIL(h, TXG_ERR_ARG_NULL, "txg_set_bc_values: null array");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const FaceDesc &fd = h->faces[boundary];
  const size_t bytes = (size_t)fd.n1 * fd.n2 * h->S * h->D * sizeof(double);
  if (!h->bc_vals[boundary]) TXG_CUDA(h, cudaMalloc((void **)&h->bc_vals[boundary], bytes));
  TXG_CUDA(h, cudaMemcpyAsync(h->bc_vals[boundary], vals, bytes, cudaMemcpyHostToDevice, h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  return 0;
}

static void drop_step_graphs(txg_flow *h) {
  for (int p = 0; p < 2; ++p) {
    if (h->step_graph[p]) cudaGraphExecDestroy(h->step_graph[p]);
    h->step_graph[p] = nullptr;
  }
  h->eager_steps = 0;
  h->graph_failed = false;
}

// capture one_step twice (parity p -> p ^ 1 -> p) on the main stream; the communication stream joins the capture through
// the events of the split schedule.  Nothing executes here.
static int build_step_graph(txg_flow *h, int p) {
  const int64_t l0 = h->launches;
  const int cur0 = h->cur;
  std::deque<KernelTimer> counts = h->timers;  // (the per-kernel launch counters move during the capture too)
  if (cudaStreamBeginCapture(h->s_main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    h->graph_failed = true;
    return 0;
  }
  int rc = one_step(h);
  if (!rc) rc = one_step(h);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->s_main, &g);
  h->graph_launches[p] = h->launches - l0;
  h->launches = l0;
  h->cur = cur0;
  // remember what one replay adds to every kernel's launch counter, then put the counters back
  for (size_t i = 0; i < h->timers.size(); ++i) {
    const int64_t before = i < counts.size() ? counts[i].launches : 0;
    h->timers[i].per_graph[p] = h->timers[i].launches - before;
    h->timers[i].launches = before;
  }
  if (rc || e != cudaSuccess || !g) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    h->graph_failed = true;  // keep the eager step
    return rc;
  }
  if (cudaGraphInstantiate(&h->step_graph[p], g, 0) != cudaSuccess) {
    cudaGetLastError();
    h->step_graph[p] = nullptr;
    h->graph_failed = true;
  }
  cudaGraphDestroy(g);
  return 0;
}

extern "C" int txg_step(txg_handle h, int nsteps) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!h->state_set) TXG_FAIL(h, TXG_ERR_ORDER, "txg_step before txg_fi_init / txg_set_fi");
  if (nsteps < 0) TXG_FAIL(h, TXG_ERR_ARG_OUTOFRANGE, "nsteps %d", nsteps);
  TXG_CUDA(h, cudaSetDevice(h->device));
  const int64_t l0 = h->launches;
  TXG_CUDA(h, cudaEventRecord(h->ev_step0, h->s_main));
  if (h->bc_mode) {
    if (nsteps > 0 && !h->forces_current) TXG_TRY(bc_moments_forces(h, false));  // FlowUpdateMoments of LBMInit2 (lbm.F90:238)
    for (int i = 0; i < nsteps; ++i) TXG_TRY(one_step_bc(h));
  } else if (h->lag) {
    for (int i = 0; i < nsteps; ++i) TXG_TRY(one_step_lag(h));
    // a density block that gave up waiting left wrong densities behind: fail here, not at the next export
    if (nsteps > 0) TXG_TRY(check_eos(h));
  } else {
    int i = 0;
    // two warm-up steps eagerly, then pairs of steps as graph replays (not while kernels are being timed one by one)
    for (; i < nsteps && h->eager_steps < 2; ++i, ++h->eager_steps) TXG_TRY(one_step(h));
    // (one rank only: replaying a graph that holds NCCL nodes costs 0.8 ms of host time per step, measured on 2 B200)
    const bool graphs = h->graph_wanted && !h->graph_failed && !h->timing && !h->pull && !h->state_g && h->cfg.nranks == 1;
    if (graphs && nsteps - i >= 2) {
      const int p = h->cur;
      if (!h->step_graph[p]) TXG_TRY(build_step_graph(h, p));
      if (h->step_graph[p]) {
        for (; nsteps - i >= 2; i += 2) {
          TXG_CUDA(h, cudaGraphLaunch(h->step_graph[p], h->s_main));
          h->launches += h->graph_launches[p];
          for (auto &t : h->timers) t.launches += t.per_graph[p];
        }
      }
    }
    for (; i < nsteps; ++i) TXG_TRY(one_step(h));
  }
  TXG_CUDA(h, cudaEventRecord(h->ev_step1, h->s_main));
  h->last_launches = h->launches - l0;
  if (!h->lag) h->rho_current = false;  // (the one-pass step leaves rho and its halo current)
  return 0;
}

// the six reference procedures: only the order is checked; the device step runs at the last one
static int phase_call(txg_flow *h, int expect, const char *name) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (h->phase != expect)
    TXG_FAIL(h, TXG_ERR_ORDER, "%s (procedure %d of 6) called out of the LBMRun2 order (lbm.F90:286-361): procedure %d is next", name,
             expect + 1, h->phase + 1);
  h->phase = (expect + 1) % 6;
  return 0;
}
extern "C" int txg_collision(txg_handle h) { return phase_call(h, 0, "FlowCollision"); }
extern "C" int txg_communicate_fi(txg_handle h) { return phase_call(h, 1, "DistributionCommunicateFi"); }
extern "C" int txg_stream(txg_handle h) { return phase_call(h, 2, "FlowStream"); }
extern "C" int txg_bounceback(txg_handle h) { return phase_call(h, 3, "FlowBounceback"); }
extern "C" int txg_apply_bcs(txg_handle h) { return phase_call(h, 4, "FlowApplyBCs"); }
extern "C" int txg_update_flux(txg_handle h) {
  TXG_TRY(phase_call(h, 5, "FlowUpdateFlux"));
  return txg_step(h, 1);
}

static int check_eos(txg_flow *h);

// ------------------------------------------------------------------ state out
// refresh rho (+halo) from the current populations
static int refresh_rho(txg_flow *h) {
  TXG_TRY(materialise(h));
  if (h->bc_mode) return h->forces_current ? 0 : bc_moments_forces(h, false);  // exports read f and Fbuf
  if (h->rho_current) return 0;
  TXG_TRY(run_moments(h, 0, h->g.NZl, h->s_main));
  TXG_TRY(exchange_rho(h, h->rho, h->s_main));
  h->rho_current = true;
  return 0;
}

extern "C" int txg_update_moments(txg_handle h) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!h->state_set) TXG_FAIL(h, TXG_ERR_ORDER, "txg_update_moments before txg_fi_init / txg_set_fi");
  TXG_CUDA(h, cudaSetDevice(h->device));
  h->rho_current = false;
  if (h->bc_mode) return bc_moments_forces(h, false);
  return refresh_rho(h);
}

// bit b: box face b carries an external BC and lies on this rank
static int bc_faces_here(const txg_flow *h) {
  int bits = 0;
  for (int b = 0; b < 2 * h->D; ++b)
    if (h->face_here[b] && h->faces[b].type >= TXG_BC_REFLECTING) bits |= 1 << b;
  return bits;
}

static int run_export(txg_flow *h, double *rho_o, double *u_o, double *F_o, double *rhot, double *prs, double *velt) {
  const Grid &g = h->g;
  ScopedKernel sk(h, "k_export", h->s_main);
  h->ks.export_state<<<blocks_for(g.nnodes, 128), 128, 0, h->s_main>>>(g, h->p, h->f[h->cur], h->rho, h->nbmask, h->ffmask, h->cls,
                                                                        h->bc_mode ? h->Fbuf : nullptr, h->has_reflecting ? h->rho_true : nullptr,
                                                                        h->bc_fused ? bc_faces_here(h) : 0, rho_o, u_o,
                                                                        F_o, rhot, prs, velt,
                                                                        h->cfg.null_pressure, 0, g.NZl);
  TXG_CUDA(h, cudaGetLastError());
  return 0;
}

static int ensure(txg_flow *h, double **p, size_t n) {
  if (*p) return 0;
  TXG_CUDA(h, cudaMalloc((void **)p, n * sizeof(double)));
  return 0;
}

extern "C" int txg_get_fi(txg_handle h, double *fi_g) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!fi_g) TXG_FAIL(h, TXG_ERR_ARG_NULL, "null array");
  if (!h->state_set) TXG_FAIL(h, TXG_ERR_ORDER, "no state on the device yet");
  TXG_CUDA(h, cudaSetDevice(h->device));
  TXG_TRY(materialise(h));
  TXG_TRY(check_eos(h));
  return export_field(h, fi_g, 1, h->D == 3 ? 1 : 0, h->Q, h->S, h->f[h->cur], 0);
}

extern "C" int txg_get_state(txg_handle h, double *rho_rg, double *u_g, double *forces_g) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!h->state_set) TXG_FAIL(h, TXG_ERR_ORDER, "no state on the device yet");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  TXG_TRY(refresh_rho(h));
  TXG_TRY(check_eos(h));
  const size_t n = (size_t)g.nnodes;
  if (rho_rg) TXG_TRY(ensure(h, &h->x_rho, n * h->S));
  if (u_g) TXG_TRY(ensure(h, &h->x_u, n * h->S * h->D));
  if (forces_g) TXG_TRY(ensure(h, &h->x_F, n * h->S * h->D));
  TXG_TRY(run_export(h, rho_rg ? h->x_rho : nullptr, u_g ? h->x_u : nullptr, forces_g ? h->x_F : nullptr, nullptr, nullptr, nullptr));
  const int gz1 = h->D == 3 ? 1 : 0;
  if (rho_rg) TXG_TRY(export_field(h, rho_rg, g.R, g.Rz, 1, h->S, h->x_rho, 1));
  if (u_g) TXG_TRY(export_field(h, u_g, 1, gz1, h->D, h->S, h->x_u, 1));
  if (forces_g) TXG_TRY(export_field(h, forces_g, 1, gz1, h->D, h->S, h->x_F, 1));
  return 0;
}

extern "C" int txg_get_diagnostics(txg_handle h, double *rhot, double *prs, double *velt) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!h->state_set) TXG_FAIL(h, TXG_ERR_ORDER, "no state on the device yet");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  TXG_TRY(refresh_rho(h));
  TXG_TRY(check_eos(h));  // (the reference stops the run in EOSApply, ierr 58; a failed device step must not hand out fields)
  const size_t n = (size_t)g.nnodes;
  if (rhot) TXG_TRY(ensure(h, &h->x_rhot, n));
  if (prs) TXG_TRY(ensure(h, &h->x_prs, n));
  if (velt) TXG_TRY(ensure(h, &h->x_velt, n * h->D));
  const bool fast = h->fused && h->ks.export_diag_fused && h->nbr_all && h->lmask && !h->band && !h->pull && !h->lag && !h->bc_mode && !h->bc_fused &&
                    !h->has_reflecting && getenv("TXG_EXPORT_GENERIC") == nullptr;
  if (fast) {
    // fused path: the fluid nodes by one lane per (node, component) over the position-indexed rows, the solid nodes by a fill
    ScopedKernel sk(h, "k_export_diag_fused", h->s_main);
    k_export_fill_solid<<<blocks_for(g.nnodes, 256), 256, 0, h->s_main>>>(g, h->nbmask, h->D, rhot ? h->x_rhot : nullptr, prs ? h->x_prs : nullptr,
                                                                          velt ? h->x_velt : nullptr, h->cfg.null_pressure);
    TXG_CUDA(h, cudaGetLastError());
    const long long nown = g.own1 - g.own0;
    if (nown) {
      h->ks.export_diag_fused<<<hot_blocks(h, nown), 128, 0, h->s_main>>>(g, h->p, h->f[h->cur], h->rho, h->lmask, h->nbr_all, h->wallrec, g.own0, nown,
                                                                          rhot ? h->x_rhot : nullptr, prs ? h->x_prs : nullptr,
                                                                          velt ? h->x_velt : nullptr);
      TXG_CUDA(h, cudaGetLastError());
    }
  } else {
    TXG_TRY(run_export(h, nullptr, nullptr, nullptr, rhot ? h->x_rhot : nullptr, prs ? h->x_prs : nullptr, velt ? h->x_velt : nullptr));
  }
  // the export kernels write all three fields in the host arrays' own (natural, owned-only) layout: three plain copies
  if (rhot) TXG_CUDA(h, cudaMemcpyAsync(rhot, h->x_rhot, n * 8, cudaMemcpyDeviceToHost, h->s_main));
  if (prs) TXG_CUDA(h, cudaMemcpyAsync(prs, h->x_prs, n * 8, cudaMemcpyDeviceToHost, h->s_main));
  if (velt) TXG_CUDA(h, cudaMemcpyAsync(velt, h->x_velt, n * h->D * 8, cudaMemcpyDeviceToHost, h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  return 0;
}

extern "C" int txg_delta_norm(txg_handle h, double *norm) {
  if (!h) return TXG_ERR_ARG_NULL;
  if (!norm) TXG_FAIL(h, TXG_ERR_ARG_NULL, "null norm");
  if (!h->state_set) TXG_FAIL(h, TXG_ERR_ORDER, "no state on the device yet");
  TXG_CUDA(h, cudaSetDevice(h->device));
  const Grid &g = h->g;
  TXG_TRY(materialise(h));
  TXG_TRY(check_eos(h));
  const long long n = (long long)h->S * h->Q * g.fs;
  if (!h->f_old) {
    TXG_CUDA(h, cudaMalloc((void **)&h->f_old, (size_t)n * 8));
    TXG_CUDA(h, cudaMemsetAsync(h->f_old, 0, (size_t)n * 8, h->s_main));
  }
  TXG_CUDA(h, cudaMemsetAsync(h->norm_bits, 0, sizeof(unsigned long long), h->s_main));
  // ghost-plane positions hold pushes in transit: compare the owned positions only, per (m,n) block
  const long long nown = g.own1 - g.own0;
  for (int b = 0; b < h->S * h->Q && nown; ++b) {
    const long long off = (long long)b * g.fs + g.own0;
    k_delta_norm<<<blocks_for(nown, 256), 256, 0, h->s_main>>>(h->f[h->cur] + off, h->f_old + off, nown, h->norm_bits);
  }
  TXG_CUDA(h, cudaGetLastError());
  h->launches += h->S * h->Q;
  // VecNorm(NORM_INFINITY) is a reduction over all ranks (lbm_distribution_function.F90:822); the
  // maximum of non-negative doubles is the maximum of their bit patterns, here taken as doubles
  if (h->cfg.nranks > 1) {
    if (!h->comm) TXG_FAIL(h, TXG_ERR_ORDER, "nranks > 1 but txg_comm_init was not called");
    TXG_NCCL(h, g_nccl.AllReduce(h->norm_bits, h->norm_bits, 1, ncclFloat64, ncclMax, h->comm, h->s_main));
  }
  unsigned long long bits = 0;
  TXG_CUDA(h, cudaMemcpyAsync(&bits, h->norm_bits, sizeof bits, cudaMemcpyDeviceToHost, h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  double v;
  memcpy(&v, &bits, sizeof v);
  *norm = h->have_old ? v : 1.e99;  // lbm_distribution_function.F90:818
  h->have_old = true;
  return 0;
}

// EOSApply_PR stops the run when its inner square root goes negative (lbm_eos.F90:337-341); the device
// kernels count such values and the next synchronising call reports them
static int check_eos(txg_flow *h) {
  if (h->lag && h->lag_done) {
    // the one-pass step gives up waiting after ~1 s instead of hanging the device (block dispatch out of linear order)
    unsigned gave_up = 0;
    TXG_CUDA(h, cudaMemcpyAsync(&gave_up, h->lag_done + h->lag_nrows, sizeof gave_up, cudaMemcpyDeviceToHost, h->s_main));
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
    if (gave_up) TXG_FAIL(h, TXG_ERR_LIB, "one-pass step (TXG_LAG=1): %u density blocks gave up waiting for their collision rows; results are invalid", gave_up);
  }
  if (h->tile || h->rtab_lag || h->band) {
    int gave_up = 0;
    TXG_CUDA(h, cudaMemcpyAsync(&gave_up, h->counters + 4, sizeof gave_up, cudaMemcpyDeviceToHost, h->s_main));
    TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
    if (gave_up) TXG_FAIL(h, TXG_ERR_LIB, "%d threads gave up waiting for the bulk copies of their density windows; results are invalid", gave_up);
  }
  bool pr = false;
  for (int m = 0; m < h->S; ++m) pr = pr || (h->cfg.use_nonideal_eos && h->cfg.eos_type[m] == TXG_EOS_PR);
  if (!pr) return 0;
  int bad = 0;
  TXG_CUDA(h, cudaMemcpyAsync(&bad, h->counters + 3, sizeof bad, cudaMemcpyDeviceToHost, h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  if (bad) TXG_FAIL(h, TXG_ERR_ORDER, "PR EOS inner sqrt went negative (%d values)", bad);  // ierr = 58 in the reference
  return 0;
}

extern "C" int txg_synchronize(txg_handle h) {
  if (!h) return TXG_ERR_ARG_NULL;
  TXG_CUDA(h, cudaSetDevice(h->device));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_main));
  TXG_CUDA(h, cudaStreamSynchronize(h->s_comm));
  return check_eos(h);
}

// ------------------------------------------------------------------ measurement hooks
extern "C" int txg_last_step_ms(txg_handle h, float *ms, int64_t *launches) {
  if (!h) return TXG_ERR_ARG_NULL;
  TXG_CUDA(h, cudaSetDevice(h->device));
  TXG_CUDA(h, cudaEventSynchronize(h->ev_step1));
  float t = 0.f;
  TXG_CUDA(h, cudaEventElapsedTime(&t, h->ev_step0, h->ev_step1));
  if (ms) *ms = t;
  if (launches) *launches = h->last_launches;
  return 0;
}
extern "C" int txg_enable_kernel_timing(txg_handle h, int on) {
  if (!h) return TXG_ERR_ARG_NULL;
  h->timing = on != 0;
  return 0;
}
extern "C" int txg_reset_kernel_times(txg_handle h) {
  if (!h) return TXG_ERR_ARG_NULL;
  drain_timers(h);
  for (auto &t : h->timers) {
    t.ms = 0.;
    t.launches = 0;
  }
  return 0;
}
extern "C" int txg_kernel_times(txg_handle h, int cap, const char **names, double *ms, int64_t *launches, int *n) {
  if (!h) return TXG_ERR_ARG_NULL;
  TXG_CUDA(h, cudaSetDevice(h->device));
  drain_timers(h);
  int k = 0;
  for (auto &t : h->timers) {
    if (k >= cap) break;
    if (names) names[k] = t.name;
    if (ms) ms[k] = t.ms;
    if (launches) launches[k] = t.launches;
    ++k;
  }
  if (n) *n = k;
  return 0;
}
