// Internal declarations of libqgcm_b200.so (sm_100a).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qgcm_b200.h"

#define QG_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " at " \
                               __FILE__ ":" + std::to_string(__LINE__));                    \
  } while (0)

// every kernel launch goes through this macro: counts it and, when profiling is on,
// brackets it with CUDA events on the launching stream
#define QG_LAUNCH(m, name, grid, block, smem, kern, ...)            \
  do {                                                              \
    qg::prof_begin(m, name);                                        \
    kern<<<grid, block, smem, (m)->stream>>>(__VA_ARGS__);          \
    qg::launch_check(m, name);                                      \
    qg::prof_end(m);                                                \
    (m)->launches++;                                                \
  } while (0)

// the same for a kernel that may start while the previous kernel of the stream still runs (programmatic
// dependent launch): the previous kernel executes griddepcontrol.launch_dependents, this one executes
// griddepcontrol.wait before it touches anything the previous one writes
#define QG_LAUNCH_PDL(m, name, grid, block, smem, kern, ...)                                   \
  do {                                                                                         \
    qg::prof_begin(m, name);                                                                   \
    cudaLaunchConfig_t lc_ = {};                                                               \
    lc_.gridDim = dim3(grid); lc_.blockDim = dim3(block); lc_.dynamicSmemBytes = smem;         \
    lc_.stream = (m)->stream;                                                                  \
    cudaLaunchAttribute at_[1];                                                                \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
    at_[0].val.programmaticStreamSerializationAllowed = 1;                                     \
    lc_.attrs = at_; lc_.numAttrs = 1;                                                         \
    QG_CUDA(cudaLaunchKernelEx(&lc_, kern, __VA_ARGS__));                                      \
    qg::launch_check(m, name);                                                                 \
    qg::prof_end(m);                                                                           \
    (m)->launches++;                                                                           \
  } while (0)

namespace qg {

constexpr int NLMAX = QGCM_NLMAX;
constexpr int TRI_L = 32;  // rows per chunk of the partitioned tridiagonal solve

// Grid description shared by ocean and atmosphere kernels.  Fields are stored
// x-fastest with a common row pitch `ld` (multiple of 16 doubles = 128 B) for both the
// p grid (nxp x nyp) and the T grid (nxt x nyt); a layer is ld*nyp doubles.
struct Grid {
  int nxt, nyt, nxp, nyp, nl;
  int ld;           // row pitch in doubles
  size_t lsz;       // layer stride in doubles = ld*nyp
  int cyclic;       // x-periodic (ocean option; atmosphere always)
  double dx, dxm2, hdxm1, rdxf0, norm;  // dx, 1/dx^2, 0.5/dx, 1/(dx f0), 1/(nxt*nyt)
  double xl, yl;
  double tdt;       // 2*dt
  // y-slab partition (multi-GPU): this grid holds global p rows [jg0, jg0+nyp) of nyp_g, of
  // which local rows [own0, own1) are owned (the rest are halo copies); single GPU:
  // jg0 = 0, own0 = 0, own1 = nyp = nyp_g.  norm, yl stay the global values.
  int jg0, nyp_g, own0, own1;
  __host__ __device__ bool wall_s() const { return jg0 == 0; }
  __host__ __device__ bool wall_n() const { return jg0 + nyp == nyp_g; }
};
constexpr int HALO = 3;   // halo rows kept on either side of a slab (del-6th of pom needs 3)

// Device-side constants small enough to pass by value to kernels
struct LayerConsts {
  double h[NLMAX], gp[NLMAX], ah2[NLMAX], ah4[NLMAX];
  double amat[NLMAX * NLMAX];   // ld = nl
  double ctl2m[NLMAX * NLMAX], ctm2l[NLMAX * NLMAX], rdm2[NLMAX];
};

// Peer-memory transport of the y-slab ranks (one process per GPU, mailboxes mapped into each
// other's address space with CUDA IPC).  The exchanges of an ocean step are done by the kernels
// that produce or consume the data: they store into the peers' mailboxes over NVLink, publish an
// epoch flag and spin on the flags the peers publish -- no collective library call on the step
// stream.  Every exchange has two slots (epoch parity): a rank can only be one exchange of a
// kind ahead of a peer, because finishing exchange e+1 needs that peer's flag e+1, which it
// publishes after it has consumed exchange e.
struct PeerCtx {
  int n, rank;                  // n == 0: transport not set up (NCCL or loopback carries the exchange)
  int fglen;                    // doubles per rank in the slab-row mailbox (nl * 2 * ld)
  int halolen;                  // doubles per side in the halo mailbox (PEER_HALO_ROWS * ld)
  double *box[8];               // mailbox of every rank (own allocation for self)
  unsigned long long epoch;     // of this exchange; flags only ever increase
};
constexpr int PEER_VEC = 16;         // doubles per rank in an all-reduce
constexpr int PEER_HALO_ROWS = 64;   // field-layer rows a halo exchange can carry per side
// mailbox layout, in doubles: all-reduce vectors [2][n][PEER_VEC], flags (all-reduce [n],
// slab rows [n], halo [2]), slab rows [2][n][fglen], halo rows [2][2][halolen]
__host__ __device__ inline size_t peer_off_vec(int n, int slot, int src) { return (size_t)(slot * n + src) * PEER_VEC; }
__host__ __device__ inline size_t peer_off_flagv(int n) { return (size_t)2 * n * PEER_VEC; }
__host__ __device__ inline size_t peer_off_flagf(int n) { return peer_off_flagv(n) + n; }
__host__ __device__ inline size_t peer_off_flagh(int n) { return peer_off_flagf(n) + n; }
__host__ __device__ inline size_t peer_off_fg(int n, int fglen, int slot, int src) {
  return ((peer_off_flagh(n) + 2 + 15) / 16) * 16 + (size_t)(slot * n + src) * fglen;
}
__host__ __device__ inline size_t peer_off_halo(int n, int fglen, int halolen, int slot, int side) {
  return peer_off_fg(n, fglen, 2, 0) + (size_t)(slot * 2 + side) * halolen;
}
__host__ __device__ inline size_t peer_box_doubles(int n, int fglen, int halolen) { return peer_off_halo(n, fglen, halolen, 2, 0); }

// tuning override read at launch time (scripts/march_sweep.py); the defaults are the measured best
inline int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// Rows per march of a marching kernel (k_qgstep2, k_oml_march).  Blocks that do not fit the
// resident capacity of the GPU run as a second wave at full cost (measured: oml at 1 km, 480
// blocks on 444 slots 0.46 ms, 440 blocks 0.33 ms), so the number of marches is the largest
// that fills a whole number of waves, with the fewest waves whose marches stay under `cap` rows.
//   bx: blocks per march row, resident: blocks the GPU holds at once, lo: shortest useful march
//   extra: marches outside the count (the two short boundary marches of k_qgstep2) that share the waves
inline int pick_march_rows(int nrows, int bx, int resident, int lo, int cap, int extra = 0) {
  for (int w = 1; w <= 64; ++w) {
    const int c = std::max(1, (w * resident) / bx - extra);
    const int rows = (nrows + c - 1) / c;
    if (rows <= cap) return std::max(lo, rows);
  }
  return cap;
}

// points per direction of the sub-sampled output of ocnc_out / atnc_out / qocdiag_out:
// min(mod(n,nsk),1) + (n - mod(n,nsk))/nsk  (src/nc_subs.F:869-876)
inline int sub_count(int n, int nsk) {
  const int mwk = n % nsk;
  return std::min(mwk, 1) + (n - mwk) / nsk;
}

// Plan for the batched x-transform + partitioned y-tridiagonal Helmholtz solver
struct HelmPlan {
  int kind;          // 0: DST-I rows (box), 1: real FFT rows (periodic)
  int n;             // real transform length (= nxt)
  int m;             // complex length n/2
  int nrad;
  int radix[8];
  int twoff[8];     // per-pass twiddle table offsets
  int nmodes;        // batch (number of vertical modes)
  int ld, nyp, nxp;
  int row0;          // first solved local row (1 on a single GPU; own0 (+1 at the southern wall) in a slab)
  int nrows;         // solved rows: the interior rows this rank owns
  int nchunk;        // ceil(nrows / TRI_L)
  int lastlen;       // rows in the last chunk
  int nk;            // number of wavenumber columns solved (n-1 box, n periodic)
  int koff;          // first column offset within a row (1 box, 0 periodic)
  double a;          // off-diagonal 1/dy^2
  double ftnorm;
  size_t smem_bytes;
  double2 *wm = nullptr;     // exp(-2 pi i k/m), k<m
  double2 *wn = nullptr;     // exp(-2 pi i k/n), k<=m
  double *sintw = nullptr;   // 2 sin(k pi/n), k<m
  double *bcoef = nullptr;   // [nmodes][ld] diagonal b per mode and column
  double *binv = nullptr;    // [nmodes][TRI_L][ld]  forward-elimination reciprocals
  double *vl = nullptr;      // [nmodes][TRI_L][ld]  left spike of a full chunk
  double *vll = nullptr;     // [nmodes][TRI_L][ld]  left spike of the last chunk
  double *pt = nullptr;      // [nmodes][nchunk][ld] block-Thomas pivots of the interface system
  double *fg = nullptr;      // [nmodes][2][nchunk][ld] first/last local values -> interface rhs
  double *yx = nullptr;      // [nmodes][2][nchunk][ld] neighbour values per chunk (yprev, xnext)
  // fast three-pass DST path (helmholtz.cu, k_dst3); fast = R3 of the plan (16, 15, R3) or 0
  int fast = 0, fast_grid = 0, fast_attr = 0;
  double2 *s1base = nullptr, *tw2 = nullptr, *tw3base = nullptr, *wnbase = nullptr;
  double c1[16], s1c[16];
  double2 wnr[16];
  // y-slab coupling (nranks > 1): every slab is one more level of the same partition
  int nranks = 1, rank = 0, wall_s = 1, wall_n = 1;
  int slab_rows[16];           // solved rows of every slab
  double *slab_send = nullptr; // [nmodes][2][ld] first/last rows of this slab's local solution
  double *slab_ae = nullptr;   // [nmodes][nranks][2][ld] left-spike first/last values (alpha, eps) of every slab
  double *slab_fg = nullptr;   // [nranks][nmodes][2][ld] all-gathered first/last rows of the slab-local solutions
  double *slab_yx = nullptr;   // [nmodes][2][ld] true neighbour rows of this slab (Y of the slab below, X of the one above)
  bool walls_dirty = true;     // someone wrote the wall rows of the work array (homsol, qgcm_helmholtz): zero them after the solve
  bool slab_pushed = false;    // helm_solve_a already delivered the slab rows to the peers' mailboxes
  PeerCtx slab_peer = {};      // peer-memory transport: the gathered rows arrive in the mailbox, k_slab_solve waits for them
  int *slab_err = nullptr;
  double *wsum = nullptr;    // [ld] box: 2 cot(k pi / 2n) at odd sine wavenumbers k (x-sum of a sine series), else 0
  double *spec = nullptr;    // [nmodes][nspec] per-block shares of the modal area integrals (k_tri3)
  int nspec = 0;
  double *rowsum = nullptr;  // [nmodes][nyp]  xintp row sums of the solution
  double *ayrow = nullptr;   // [nmodes][2]    periodic: line sums of rows 2 and nyp-1
};

// Box ocean with the fast DST plan: ocinvq runs as forward transform of the vorticity layers ->
// tridiagonal kernels that project layers <-> modes in spectral space -> constraint algebra ->
// inverse transform that adds the homogeneous solutions and writes the pressure layers
// (helmholtz.cu, k_dst3 / k_tri3); this is what those kernels need beyond the plan
struct FusedInv {
  const double *q;          // vorticity layers [nl][nyp][ld]
  double *pnew;             // new pressure layers (the lagged-p buffer)
  const double *yrel;
  double beta, f0;
  const double *ddyn;       // null over a flat bottom
  const double *hom;        // ochom [nl-1][nyp][ld]
  const double *coef;       // device hclco[nl-1], written by k_inv_scalars before the inverse transform
  double ctl2m[NLMAX * NLMAX], ctm2l[NLMAX * NLMAX];
};

// Scratch and tables of the coupled forcing xforc (atmos.cu); built on first use
struct XfPlan {
  bool ready = false;
  int nxf = 0, nyf = 0, ldf = 0;         // ocean-resolution atmosphere p grid (nxpaor x nypaor) and its pitch
  double *u1 = nullptr, *v1 = nullptr;   // layer-1 geostrophic velocity at atmosphere p points [nypa][ld]
  double *taux = nullptr, *tauy = nullptr;   // stress on the fine grid [nyf][ldf]
  double *stb = nullptr;                 // bicubic weights [5][16][(ndxr+1)^2]: general, u-south, v-south, u-north, v-north
  int *iam = nullptr, *iap = nullptr, *jam = nullptr, *jap = nullptr;   // bilint subscripts (0-based)
  double *wpx = nullptr, *wmx = nullptr, *wpy = nullptr, *wmy = nullptr;
  double *fsp_o = nullptr, *fsp_a = nullptr;   // fsprim at ocean / atmosphere T rows
  double *part = nullptr;                // per-block partial sums
  int npart = 0;
};

struct Model;

}  // namespace qg

struct qgcm_model {
  qgcm_config cfg;
  int flags;
  bool ocean_only, atmos_only, cyclic, sb_hflux, nb_hflux, tau_udiff;
  bool has_ocean, has_atmos;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  qg::Grid go, ga;                 // ocean, atmosphere grids
  qg::LayerConsts lo, la;
  double fnot, beta;
  double dto, dta;
  double rrcpoc, rrcpat, raoro;
  // per-row latitude arrays on device
  double *yporel = nullptr, *yparel = nullptr, *ytorel = nullptr, *ytarel = nullptr;
  std::vector<double> h_ypo, h_ypa;
  // every named device field: pointer, logical (nx, ny, nl)
  struct Field {
    double *d = nullptr;
    int nx = 0, ny = 0, nl = 1, ld = 0;   // ld == 0: dense 1-D/2-D small array
    int nyg = 0, joff = 0, o0 = 0, o1 = 0;   // y-slab: global rows, global index of local row 0, owned local rows
    size_t lsz = 0;                       // device layer stride (gridded fields)
    size_t elems = 0;                     // device elements allocated
  };
  std::map<std::string, Field> fields;
  // leapfrog pointer rotation: logical name -> current buffer (the map entries are
  // swapped, host never sees the rotation)
  qgcm_scalars *d_scal = nullptr;        // device-resident scalar state
  double *d_coef = nullptr;              // small device scratch for step coefficients
  double *d_red = nullptr;               // reduction scratch
  double *d_red_a = nullptr;             // the atmosphere mixed layer's (= d_red in ocean-only models)
  cudaStream_t at_stream = nullptr;      // coupled cycles: the atmosphere steps' branch (api.cu cycle_body)
  cudaEvent_t ev_atfork = nullptr, ev_atjoin = nullptr;
  size_t red_elems = 0;
  qg::HelmPlan hpo, hpa;
  qg::XfPlan xf;
  // multi-GPU: rank layout, communicator (NCCL, or in-process peers sharing one stream) and
  // the small device vector that carries all-reduce payloads
  int nranks = 1, rank = 0;
  void *nccl = nullptr;
  std::vector<qgcm_model *> peers;       // loopback group (all ranks in this process), empty otherwise
  double *d_cv = nullptr;                // [32] reduction payload
  unsigned int *d_ticket = nullptr;      // last-block-done counters
  double *d_val = nullptr;               // valids: per-block partials
  // running sums of src/timavge.F (timavg.cu): contribution counts; packing buffer of qgcm_get_field_sub
  int nsumat = 0, nsumoc = 0, nsum_ocavg = 0;
  double *d_pack = nullptr;
  size_t pack_elems = 0;
  double *d_mon = nullptr, *d_monf = nullptr;   // qgcm_monnc_ocean: per-row sums; four scratch fields
  size_t mon_elems = 0;
  double *d_mona = nullptr, *d_monaf = nullptr;  // qgcm_monnc_atmos: the same for the atmosphere grid
  size_t mona_elems = 0;
  bool shared_stream = false;            // loopback ranks > 0 borrow rank 0's stream
  // peer-memory transport (slab.cu): own mailbox, peers' mapped mailboxes, exchange counters
  double *mailbox = nullptr;
  std::vector<void *> peer_maps;         // cudaIpcOpenMemHandle results to close
  qg::PeerCtx peer = {};
  bool use_peer = false;                 // qgcm_comm_init_peer / qgcm_comm_transport: exchanges go through the mailboxes
  unsigned long long epoch_vec = 0, epoch_fg = 0, epoch_halo = 0;
  unsigned int *d_ticket2 = nullptr;     // [0] slab-row push, [1] halo push
  int *d_peer_err = nullptr;             // device error block of the peer transport (layout at peer_wait)
  int *h_peer_err = nullptr;             // host-mapped copy of its flag: read on every call, no stream sync needed
  // qgcm_set_field_async: copy stream, shadow buffers, names waiting for qgcm_commit_fields
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_step = nullptr;
  std::map<std::string, double *> shadow;
  std::vector<std::string> pending;      // each name at most once: one buffer swap per commit
  bool ddynoc_flat = true, ddynat_flat = true;   // the topography field is identically zero (its device buffer starts zeroed)
  double *wrk_o = nullptr, *wrk_a = nullptr;   // modal work arrays [nl][nyp][ld]
  double *xfo = nullptr, *xfa = nullptr, *sstnew = nullptr, *astnew = nullptr, *hmnew = nullptr;
  std::vector<void *> allocs;
  // per model (= per device): resident-block counts of the marching kernels; the dynamic
  // shared-memory attribute is (re)set with them, so two models on two devices in one
  // process each configure their own device
  int qg_resident = 0, oml_resident = 0;
  // second stream for small launches that are independent of a big one running beside them (the
  // wall-warp launches of the vorticity step); forked from and joined to `stream` with events, so
  // stream order -- and a stream capture -- sees them as part of the step
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // per-kernel event timing (qgcm_profile)
  bool prof = false;
  struct ProfRec { const char *name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;

  // CUDA graphs of whole timesteps (api.cu, run_graphed): the 15 launches of an ocean step and the
  // ~20 of an atmosphere step are latency bound on small decks.  The leapfrog rotates buffer
  // pointers on the host (q and p with period 2, the mixed layers with period 3), so a graph is
  // cached per (kind of step, pointer state) together with the pointer state it leaves behind.
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    std::vector<double *> after;   // pointer state after the step (fields in map order, then sstnew, astnew, hmnew)
    int64_t launches = 0;
  };
  std::map<std::string, StepGraph> graphs;
  int eager_steps[4] = {0, 0, 0, 0};     // steps of each kind run before graphs are used (lazy set-up happens in them)

  double *F(const char *name) { return fields.at(name).d; }
  void swapf(const char *a, const char *b) { std::swap(fields.at(a).d, fields.at(b).d); }
};

#ifdef __CUDACC__
namespace qg {
// Error block of the peer transport, 8 ints in device memory (qgcm_model::d_peer_err):
//   [0]    sticky error flag, read by every later wait
//   [2,3]  64-bit address of a host-mapped copy of the flag (the host reads it on every call
//          without synchronising the stream), 0 when absent
//   [4,5]  64-bit give-up time of a wait in GPU clocks (qgcm_comm_peer_timeout, default 120 s)
// spin until *flag >= epoch.  A lost peer must not hang the device for ever: the wait gives up
// after the configured time and raises the flag (device copy and host-mapped copy); once it is
// set every later wait returns at once, the step finishes with whatever the mailbox holds, and
// the next library call of any kind on that model reports the error (slab.cu, check_peer_err).
__device__ __forceinline__ void peer_wait(const volatile unsigned long long *flag, unsigned long long epoch, int *err) {
  const long long t0 = clock64();
  unsigned int spins = 0;
  while (*flag < epoch) {
    if ((++spins & 1023u) == 0) {
      if (*reinterpret_cast<volatile int *>(err)) break;
      const long long limit = *reinterpret_cast<volatile long long *>(err + 4);
      if (clock64() - t0 > limit) {
        *reinterpret_cast<volatile int *>(err) = 1;
        int *host = *reinterpret_cast<int *volatile *>(err + 2);
        if (host) *reinterpret_cast<volatile int *>(host) = 1;
        __threadfence_system();
        break;
      }
    }
  }
}
// All-reduce (sum, rank order) of my[0..n) over the slab ranks, executed by one whole block of
// >= n_ranks*n threads: peer stores, system fence, epoch flags, wait, local sum.  out may alias my.
__device__ __forceinline__ void peer_allreduce_block(const PeerCtx &c, const double *my, int n, double *out, int *err) {
  const int slot = (int)(c.epoch & 1ull);
  for (int idx = threadIdx.x; idx < c.n * n; idx += blockDim.x) {
    const int r = idx / n, i = idx - r * n;
    c.box[r][peer_off_vec(c.n, slot, c.rank) + i] = my[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < c.n) {
    reinterpret_cast<volatile unsigned long long *>(c.box[threadIdx.x] + peer_off_flagv(c.n))[c.rank] = c.epoch;
    peer_wait(reinterpret_cast<const volatile unsigned long long *>(c.box[c.rank] + peer_off_flagv(c.n)) + threadIdx.x, c.epoch, err);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < n) {
    const volatile double *mine = c.box[c.rank];
    double s = 0.0;
    for (int r = 0; r < c.n; ++r) s += mine[peer_off_vec(c.n, slot, r) + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}
// deterministic sum of v[first..last) by one block of 256 threads (fixed strided partials,
// fixed-order tree); every thread returns the total.  red: 8 doubles of shared memory.
__device__ __forceinline__ double block256_range_sum(const double *v, int first, int last, double *red) {
  double s = 0.0;
  for (int i = first + threadIdx.x; i < last; i += 256) s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}
}  // namespace qg
#endif

namespace qg {

void *dalloc(qgcm_model *m, size_t bytes);
void prof_begin(qgcm_model *m, const char *name);
void prof_end(qgcm_model *m);
void launch_check(qgcm_model *m, const char *name);   // names the kernel whose launch was rejected

// helmholtz.cu
void helm_plan_create(qgcm_model *m, HelmPlan &hp, const Grid &g, int kind, const double *rdm2, int nmodes);
void helm_set_diag(qgcm_model *m, HelmPlan &hp, const double *b_host_per_mode);  // [nmodes][n]
void helm_solve(qgcm_model *m, HelmPlan &hp, double *wrk, int nmodes);

// ocean.cu / atmos.cu / misc.cu entry points (host launchers)
void launch_qgostep(qgcm_model *m);
void launch_ocinvq(qgcm_model *m);
void launch_oml(qgcm_model *m);
void launch_ocqbdy(qgcm_model *m, double *q, const double *p);
void launch_qcomp(qgcm_model *m, bool ocean, double *q, const double *p);
void launch_xforc_ocean_ekman(qgcm_model *m);
void launch_tlavg_ocean(qgcm_model *m);
void launch_constr(qgcm_model *m);
void launch_homsol(qgcm_model *m);
void launch_qgastep(qgcm_model *m);
void launch_atinvq(qgcm_model *m);
void launch_aml(qgcm_model *m);
void launch_atqzbd(qgcm_model *m, double *q, const double *p);
void launch_tlavg_atmos(qgcm_model *m);
void launch_xforc(qgcm_model *m);
void launch_valids(qgcm_model *m, qgcm_valids_report *rep);
// timavg.cu
void add_field(qgcm_model *m, const char *name, int nx, int ny, int nl, int ld, size_t lsz = 0, const Grid *slab = nullptr);
void launch_tavini(qgcm_model *m);
void launch_tavocn(qgcm_model *m);
void launch_tavatm(qgcm_model *m);
void launch_avg_ocn_k247(qgcm_model *m);
void field_sub_size(qgcm_model *m, const char *name, int nsk, int64_t *n);
void get_field_sub(qgcm_model *m, const char *name, int nsk, double *host, int64_t n);
// monitor.cu
void launch_monnc_ocean(qgcm_model *m, qgcm_monitor_ocean *rep);
void launch_monnc_atmos(qgcm_model *m, qgcm_monitor_atmos *rep);
// qocdiag.cu
void qocdiag_size(qgcm_model *m, int nsk, int64_t *n);
void launch_qocdiag(qgcm_model *m, int nsk, double *host, int64_t n);

// slab.cu: y-slab multi-GPU drivers.  `ms` is the set of ranks this process drives: one model
// with an NCCL communicator, or every rank of an in-process loopback group.
typedef std::vector<qgcm_model *> Ranks;
void slab_bounds(int nyp_global, int nranks, int rank, int *p0, int *p1);
void comm_allreduce_cv(const Ranks &ms, size_t off, int n);
void comm_allreduce_host(const Ranks &ms, std::vector<std::vector<double>> &vals);
void comm_halo(const Ranks &ms, const std::vector<const char *> &fields);
void nccl_unique_id(void *out128);
void nccl_init(qgcm_model *m, const void *id128);
void nccl_destroy(qgcm_model *m);
void group_create(qgcm_model **models, int n);
void peer_export(qgcm_model *m, void *handle64);
void peer_init(qgcm_model *m, const void *handles, int n);
void peer_close(qgcm_model *m);
void set_transport(qgcm_model *m, int kind);
bool peer_active(const qgcm_model *m);    // this model's exchanges go through the peer mailboxes
void check_peer_err(qgcm_model *m);       // throws once a peer exchange has timed out (host-side flag, no sync)
void peer_set_timeout(qgcm_model *m, double seconds);
PeerCtx peer_next_vec(qgcm_model *m);     // context of the next all-reduce (advances the epoch); n == 0 when not active
Ranks ranks_of(qgcm_model *m);
void slab_ocean_step(const Ranks &ms);
void slab_constr(const Ranks &ms);
void slab_homsol(const Ranks &ms);
void slab_qcomp_ocean(const Ranks &ms);
void slab_tlavg_ocean(const Ranks &ms);
// pieces of the single-GPU launchers that the slab drivers interleave with communication
void oml_phase_a(qgcm_model *m);                 // up to the local partial sums (d_cv[0..2])
void oml_phase_b(qgcm_model *m);                 // entoc and its local integral (d_cv[3])
void ocinvq_phase_a(qgcm_model *m);              // l2m, forward transform, local chunk solve, slab first/last rows
void ocinvq_phase_b(qgcm_model *m);              // slab coupling, final rows, inverse transform, local integrals (d_cv[4..])
void ocinvq_phase_c(qgcm_model *m);              // constraint algebra, mode->layer
void homsol_box_a(qgcm_model *m);
void homsol_box_b(qgcm_model *m, std::vector<double> &share);
void homsol_box_c(qgcm_model *m, const std::vector<double> &aipohs);
void constr_ocean_share(qgcm_model *m, std::vector<double> &v);
void constr_ocean_store(qgcm_model *m, const std::vector<double> &v);
void helm_clean_walls(qgcm_model *md, HelmPlan &hp, double *wrk, int nmodes);
void helm_solve_a(qgcm_model *m, HelmPlan &hp, double *wrk, int nmodes, const FusedInv *fz = nullptr);
void helm_solve_b(qgcm_model *m, HelmPlan &hp, double *wrk, int nmodes, const FusedInv *fz = nullptr);
bool helm_can_fuse(const qgcm_model *m, const HelmPlan &hp, int nl);   // box ocean, fast DST plan, 3 layers
void helm_fused_inverse(qgcm_model *m, HelmPlan &hp, double *wrk, int nl, const FusedInv &fz);

}  // namespace qg
