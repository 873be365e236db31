// y-slab multi-GPU layer for the ocean-only box decks (NAtl 2 km / 1 km).
//
// Each rank holds global p rows [p0 - HALO, p1 + HALO) of the domain as an ordinary local
// grid, so every stencil kernel runs unchanged: it treats the slab edges as walls, which
// spoils at most HALO rows next to an artificial edge -- exactly the halo rows, which are
// then overwritten with the neighbour's owned rows.  Per ocean step the ranks exchange
//   * one all-reduce of 3 doubles inside oml (xfosum and two monitors),
//   * one all-gather of 2 rows per mode for the Helmholtz solve (the slab is one more level
//     of the chunk partition, see helmholtz.cu; no transpose, no halo for the transforms),
//   * one all-reduce of 1+nl doubles (entrainment integral, modal integrals) before the
//     constraint algebra, which every rank then evaluates redundantly,
//   * one halo exchange of HALO rows of po, qo and sst with each neighbour.
// Transport is NCCL (one process per GPU; the library dlopen()s libnccl.so.2 so that it loads
// on machines without it), or an in-process loopback group: all ranks in one process on one
// device and one stream, driven in lockstep -- that is how the single-GPU test box and the
// CPU-side reasoning exercise the N > 1 path.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "qgcm_internal.h"

namespace qg {

void slab_bounds(int nyp_global, int nranks, int rank, int *p0, int *p1) {
  const int base = nyp_global / nranks, rem = nyp_global % nranks;
  *p0 = rank * base + std::min(rank, rem);
  *p1 = *p0 + base + (rank < rem ? 1 : 0);
}

// ---------------------------------------------------------------- NCCL through dlopen
namespace {
typedef struct { char internal[128]; } nccl_uid;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(nccl_uid *) = nullptr;
  int (*CommInitRank)(void **, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;   // ncclFloat64, ncclSum (nccl.h)

NcclApi &nccl() {
  static NcclApi a;
  if (a.h) return a;
  a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.h) throw std::runtime_error(std::string("NCCL is not available: ") + dlerror());
  auto sym = [&](const char *n) {
    void *p = dlsym(a.h, n);
    if (!p) throw std::runtime_error(std::string("libnccl lacks ") + n);
    return p;
  };
  a.GetUniqueId = (int (*)(nccl_uid *))sym("ncclGetUniqueId");
  a.CommInitRank = (int (*)(void **, int, nccl_uid, int))sym("ncclCommInitRank");
  a.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
  a.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclAllReduce");
  a.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))sym("ncclAllGather");
  a.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))sym("ncclSend");
  a.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))sym("ncclRecv");
  a.GroupStart = (int (*)())sym("ncclGroupStart");
  a.GroupEnd = (int (*)())sym("ncclGroupEnd");
  a.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
  return a;
}
void nccl_ok(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + nccl().GetErrorString(rc));
}
}  // namespace

void nccl_unique_id(void *out128) {
  nccl_uid id;
  nccl_ok(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, &id, sizeof(id));
}
void nccl_init(qgcm_model *m, const void *id128) {
  if (m->nranks < 2) throw std::runtime_error("qgcm_comm_init_nccl: model was created with nranks = 1");
  nccl_uid id;
  std::memcpy(&id, id128, sizeof(id));
  QG_CUDA(cudaSetDevice(m->cfg.device));
  nccl_ok(nccl().CommInitRank(&m->nccl, m->nranks, id, m->rank), "ncclCommInitRank");
}
void nccl_destroy(qgcm_model *m) {
  if (m->nccl) nccl().CommDestroy(m->nccl);
  m->nccl = nullptr;
}

// ---------------------------------------------------------------- peer-memory transport (CUDA IPC)
// One process per GPU.  Every rank allocates a mailbox, hands out its CUDA IPC handle
// (qgcm_peer_handle), the host program gathers the handles (MPI_Allgather in the Fortran driver,
// torch.distributed in bench.py) and every rank maps the others' mailboxes
// (qgcm_comm_init_peer).  From then on the exchanges of a step are stores into the peers'
// mailboxes plus epoch flags, issued by the kernels of the step themselves.
static int peer_halolen(const qgcm_model *m) { return PEER_HALO_ROWS * m->go.ld; }

void peer_export(qgcm_model *m, void *handle64) {
  if (m->nranks < 2) throw std::runtime_error("qgcm_peer_handle: model was created with nranks = 1");
  if (m->nranks > 8) throw std::runtime_error("qgcm_peer_handle: at most 8 ranks");
  if (!m->mailbox) {
    const int fglen = m->go.nl * 2 * m->hpo.ld;
    const size_t n = peer_box_doubles(m->nranks, fglen, peer_halolen(m));
    m->mailbox = (double *)dalloc(m, sizeof(double) * n);
    m->d_ticket2 = (unsigned int *)dalloc(m, sizeof(unsigned int) * 4);
    m->d_peer_err = (int *)dalloc(m, sizeof(int) * 8);
    QG_CUDA(cudaMemset(m->mailbox, 0, sizeof(double) * n));
    QG_CUDA(cudaMemset(m->d_ticket2, 0, sizeof(unsigned int) * 4));
    QG_CUDA(cudaHostAlloc((void **)&m->h_peer_err, sizeof(int) * 4, cudaHostAllocMapped));
    m->h_peer_err[0] = 0;
    int *dev_view = nullptr;
    QG_CUDA(cudaHostGetDevicePointer((void **)&dev_view, m->h_peer_err, 0));
    long long blk[4] = {0, (long long)(uintptr_t)dev_view, 0, 0};
    QG_CUDA(cudaMemcpy(m->d_peer_err, blk, sizeof(blk), cudaMemcpyHostToDevice));
    const char *ts = std::getenv("QGCM_PEER_TIMEOUT_S");
    peer_set_timeout(m, (ts && *ts) ? std::atof(ts) : 120.0);
    QG_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  QG_CUDA(cudaIpcGetMemHandle(&h, m->mailbox));
  static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
  std::memcpy(handle64, &h, sizeof(h));
}

// handles: nranks x 64 bytes, in rank order (every rank's qgcm_peer_handle)
void peer_init(qgcm_model *m, const void *handles, int n) {
  if (n != m->nranks || n > 8) throw std::runtime_error("qgcm_comm_init_peer: handle count differs from nranks (<= 8)");
  if (!m->mailbox) throw std::runtime_error("qgcm_comm_init_peer: call qgcm_peer_handle first");
  if (m->peer.n) throw std::runtime_error("qgcm_comm_init_peer: already initialised");
  QG_CUDA(cudaSetDevice(m->cfg.device));
  PeerCtx c = {};
  c.n = n; c.rank = m->rank; c.fglen = m->go.nl * 2 * m->hpo.ld; c.halolen = peer_halolen(m); c.epoch = 0;
  for (int r = 0; r < n; ++r) {
    if (r == m->rank) { c.box[r] = m->mailbox; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char *)handles + (size_t)r * 64, sizeof(h));
    void *p = nullptr;
    QG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    m->peer_maps.push_back(p);
    c.box[r] = (double *)p;
  }
  m->peer = c;
  m->use_peer = true;
}

void peer_close(qgcm_model *m) {
  for (void *p : m->peer_maps) cudaIpcCloseMemHandle(p);
  m->peer_maps.clear();
  m->peer.n = 0;
  m->use_peer = false;
}

// kind 0: NCCL, 1: peer mailboxes.  Every rank must switch between the same two steps.
void set_transport(qgcm_model *m, int kind) {
  if (kind == 0) {
    if (!m->nccl) throw std::runtime_error("qgcm_comm_transport: no NCCL communicator (qgcm_comm_init_nccl)");
    m->use_peer = false;
  } else if (kind == 1) {
    if (!m->peer.n) throw std::runtime_error("qgcm_comm_transport: no peer mailboxes (qgcm_comm_init_peer)");
    m->use_peer = true;
  } else {
    throw std::runtime_error("qgcm_comm_transport: kind is 0 (NCCL) or 1 (peer memory)");
  }
}

bool peer_active(const qgcm_model *m) { return m->peers.empty() && m->use_peer && m->peer.n > 0; }

PeerCtx peer_next_vec(qgcm_model *m) {
  PeerCtx c = m->peer;
  if (!peer_active(m)) { c.n = 0; return c; }
  c.epoch = ++m->epoch_vec;
  return c;
}

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerCtx c, double *v, int n, int *err) {
  peer_allreduce_block(c, v, n, v, err);
}

// this rank's slab rows (hpo.slab_send) into every rank's mailbox; the block that finishes last
// publishes the epoch.  grid (blocks per peer, nranks).  The consumer (k_slab_solve) waits for
// the flags of all ranks and reads the rows from its own mailbox.
__global__ void __launch_bounds__(256) k_slab_push(PeerCtx c, const double *send, unsigned int *ticket) {
  __shared__ bool last;
  const int slot = (int)(c.epoch & 1ull);
  double2 *dst = reinterpret_cast<double2 *>(c.box[blockIdx.y] + peer_off_fg(c.n, c.fglen, slot, c.rank));
  const double2 *src = reinterpret_cast<const double2 *>(send);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.fglen / 2; i += gridDim.x * blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x == 0) *ticket = 0u;
  if ((int)threadIdx.x < c.n)
    reinterpret_cast<volatile unsigned long long *>(c.box[threadIdx.x] + peer_off_flagf(c.n))[c.rank] = c.epoch;
}

// Halo exchange through the mailboxes.  One block per (row, side): side 0 carries my top owned
// rows up (they land in the "from below" half of rank+1's mailbox), side 1 my bottom owned rows
// down.  The last block to finish publishes the epoch to both neighbours; then every block waits
// for the neighbour on its side and copies that neighbour's rows from the own mailbox into the
// halo rows of the field.  Staging (instead of storing into the neighbour's field) matters: the
// neighbour may still be inside its own step, writing those very rows.
struct HaloArgs {
  PeerCtx c;
  int nfl, ld, nrows;          // field layers, row pitch, HALO
  double *base[24];            // first row of every field layer
  int ny[24];                  // rows the slab holds of that layer (T fields: one less than p fields)
  int own0, own1;              // owned p rows [own0, own1) of the slab
  unsigned int *ticket;
  int *err;
};
constexpr int HALO_XS = 4;     // blocks per row
__device__ __forceinline__ void halo_copy(double2 *dst, const double2 *src, int i0, int i1, bool vol) {
  // three 16-byte loads in flight per thread before the first store
  for (int i = i0 + (int)threadIdx.x; i < i1; i += 3 * 256) {
    double2 v[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int j = i + u * 256;
      if (j < i1) {
        if (vol)
          asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(src + j));
        else
          v[u] = src[j];
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int j = i + u * 256;
      if (j < i1) dst[j] = v[u];
    }
  }
}
__global__ void __launch_bounds__(256) k_halo_peer(HaloArgs a) {
  __shared__ bool last;
  const PeerCtx &c = a.c;
  const int slot = (int)(c.epoch & 1ull);
  const int xs = blockIdx.x % HALO_XS, rowid = blockIdx.x / HALO_XS;
  const int fl = rowid / a.nrows, r = rowid - fl * a.nrows, side = blockIdx.y;
  const int n2 = a.ld / 2, per = (n2 + HALO_XS - 1) / HALO_XS, i0 = xs * per, i1 = min(n2, i0 + per);
  const bool has_dn = c.rank > 0, has_up = c.rank + 1 < c.n;
  // ---- push
  if (side == 0 && has_up) {
    const double2 *src = reinterpret_cast<const double2 *>(a.base[fl] + (size_t)(a.own1 - a.nrows + r) * a.ld);
    double2 *dst = reinterpret_cast<double2 *>(c.box[c.rank + 1] + peer_off_halo(c.n, c.fglen, c.halolen, slot, 0) +
                                               (size_t)(fl * a.nrows + r) * a.ld);
    halo_copy(dst, src, i0, i1, false);
  }
  if (side == 1 && has_dn) {
    const double2 *src = reinterpret_cast<const double2 *>(a.base[fl] + (size_t)min(a.own0 + r, a.ny[fl] - 1) * a.ld);
    double2 *dst = reinterpret_cast<double2 *>(c.box[c.rank - 1] + peer_off_halo(c.n, c.fglen, c.halolen, slot, 1) +
                                               (size_t)(fl * a.nrows + r) * a.ld);
    halo_copy(dst, src, i0, i1, false);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(a.ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x == 0) {
      *a.ticket = 0u;
      if (has_up) reinterpret_cast<volatile unsigned long long *>(c.box[c.rank + 1] + peer_off_flagh(c.n))[0] = c.epoch;
      if (has_dn) reinterpret_cast<volatile unsigned long long *>(c.box[c.rank - 1] + peer_off_flagh(c.n))[1] = c.epoch;
    }
  }
  // ---- pull: side 0 = rows from below into my bottom halo, side 1 = rows from above into my top halo
  if ((side == 0 && !has_dn) || (side == 1 && !has_up)) return;
  if (threadIdx.x == 0)
    peer_wait(reinterpret_cast<const volatile unsigned long long *>(c.box[c.rank] + peer_off_flagh(c.n)) + side, c.epoch, a.err);
  __syncthreads();
  __threadfence_system();
  const int row = side == 0 ? r : a.own1 + r;
  if (row >= a.ny[fl]) return;          // T fields hold one halo row less above the owned rows
  const double2 *src = reinterpret_cast<const double2 *>(c.box[c.rank] + peer_off_halo(c.n, c.fglen, c.halolen, slot, side) +
                                                         (size_t)(fl * a.nrows + r) * a.ld);
  halo_copy(reinterpret_cast<double2 *>(a.base[fl] + (size_t)row * a.ld), src, i0, i1, true);
}

// The device raises the flag in host-mapped memory as well, so this is a plain host load: it is
// made on every step and on every call that hands state back to the host (qgcm_get_field,
// qgcm_get_scalars, qgcm_sync ...), after their stream synchronisation where they have one.
void check_peer_err(qgcm_model *m) {
  if (!m->h_peer_err) return;
  if (*reinterpret_cast<volatile int *>(m->h_peer_err))
    throw std::runtime_error("y-slab exchange timed out waiting for a peer rank (qgcm_comm_peer_timeout); the slab state is "
                             "no longer valid: destroy the models of this partition and restart from the last restart file");
}

// give-up time of a mailbox wait.  Ranks must enter each ocean step within this time of each
// other; a host that stalls one rank for longer (restart or netCDF output on one rank) should
// put a host barrier in front of the next step or raise the limit.
void peer_set_timeout(qgcm_model *m, double seconds) {
  if (!m->d_peer_err) throw std::runtime_error("qgcm_comm_peer_timeout: call qgcm_peer_handle first");
  if (!(seconds > 0.0)) throw std::runtime_error("qgcm_comm_peer_timeout: the time must be positive");
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, m->cfg.device);
  const long long clocks = (long long)std::min(seconds * 1e3 * (double)khz, 9.0e18);
  QG_CUDA(cudaMemcpy(reinterpret_cast<long long *>(m->d_peer_err) + 2, &clocks, sizeof(clocks), cudaMemcpyHostToDevice));
}

Ranks ranks_of(qgcm_model *m) {
  if (!m->peers.empty()) return m->peers;
  return Ranks{m};
}

static void check_comm(const Ranks &ms) {
  qgcm_model *m = ms[0];
  if (m->nranks == 1) return;
  if (ms.size() == 1 && !m->nccl && !peer_active(m))
    throw std::runtime_error("y-slab model has no communicator: call qgcm_comm_init_nccl, qgcm_comm_init_peer or qgcm_group_create first");
  if (ms.size() > 1 && (int)ms.size() != m->nranks) throw std::runtime_error("loopback group does not hold every rank");
}

// ---------------------------------------------------------------- collectives
struct PtrList { double *p[8]; int n; };
// fixed rank order: every rank ends with bit-identical sums
__global__ void k_loop_allreduce(PtrList l, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int r = 0; r < l.n; ++r) s += l.p[r][i];
  for (int r = 0; r < l.n; ++r) l.p[r][i] = s;
}

// sum ms[*]->d_cv[off .. off+n) over the ranks, in place
void comm_allreduce_cv(const Ranks &ms, size_t off, int n) {
  check_comm(ms);
  qgcm_model *m0 = ms[0];
  if (m0->nranks == 1) return;
  if (peer_active(m0)) {
    if (n > PEER_VEC) throw std::runtime_error("comm_allreduce_cv: payload too large for the mailbox");
    QG_LAUNCH(m0, "k_peer_allreduce", 1, 256, 0, k_peer_allreduce, peer_next_vec(m0), m0->d_cv + off, n, m0->d_peer_err);
    return;
  }
  if (ms.size() == 1) {
    nccl_ok(nccl().AllReduce(m0->d_cv + off, m0->d_cv + off, (size_t)n, NCCL_DOUBLE, NCCL_SUM, m0->nccl, m0->stream), "ncclAllReduce");
    return;
  }
  PtrList l;
  l.n = (int)ms.size();
  for (int r = 0; r < l.n; ++r) l.p[r] = ms[r]->d_cv + off;
  QG_LAUNCH(m0, "k_loop_allreduce", (n + 63) / 64, 64, 0, k_loop_allreduce, l, n);
}

// host-side values (initialisation procedures): vals[r] is rank r's share; all become the sum
void comm_allreduce_host(const Ranks &ms, std::vector<std::vector<double>> &vals) {
  check_comm(ms);
  qgcm_model *m0 = ms[0];
  if (m0->nranks == 1) return;
  const size_t n = vals[0].size();
  if (n > (size_t)PEER_VEC) throw std::runtime_error("comm_allreduce_host: payload too large");
  for (size_t r = 0; r < ms.size(); ++r)
    QG_CUDA(cudaMemcpyAsync(ms[r]->d_cv + 8, vals[r].data(), sizeof(double) * n, cudaMemcpyHostToDevice, ms[r]->stream));
  comm_allreduce_cv(ms, 8, (int)n);
  for (size_t r = 0; r < ms.size(); ++r) {
    QG_CUDA(cudaMemcpyAsync(vals[r].data(), ms[r]->d_cv + 8, sizeof(double) * n, cudaMemcpyDeviceToHost, ms[r]->stream));
    QG_CUDA(cudaStreamSynchronize(ms[r]->stream));
  }
}

// hpo.slab_send (n doubles per rank) -> hpo.slab_fg (rank-major) on every rank
static void comm_allgather_slab(const Ranks &ms) {
  check_comm(ms);
  qgcm_model *m0 = ms[0];
  const size_t n = (size_t)m0->go.nl * 2 * m0->hpo.ld;
  if (peer_active(m0)) {
    // the rows go straight into every rank's mailbox; k_slab_solve waits for them there
    if (m0->hpo.slab_pushed) return;      // k_tri_reduced has delivered them already
    PeerCtx c = m0->peer;
    c.epoch = ++m0->epoch_fg;
    m0->hpo.slab_peer = c;
    m0->hpo.slab_err = m0->d_peer_err;
    QG_LAUNCH(m0, "k_slab_push", dim3(16, c.n), 256, 0, k_slab_push, c, m0->hpo.slab_send, m0->d_ticket2);
    return;
  }
  m0->hpo.slab_peer.n = 0;
  if (ms.size() == 1) {
    nccl_ok(nccl().AllGather(m0->hpo.slab_send, m0->hpo.slab_fg, n, NCCL_DOUBLE, m0->nccl, m0->stream), "ncclAllGather");
    return;
  }
  for (size_t d = 0; d < ms.size(); ++d)
    for (size_t s = 0; s < ms.size(); ++s)
      QG_CUDA(cudaMemcpyAsync(ms[d]->hpo.slab_fg + s * n, ms[s]->hpo.slab_send, sizeof(double) * n, cudaMemcpyDeviceToDevice,
                              m0->stream));
}

// HALO owned rows next to each inner slab edge replace the neighbour's halo rows
void comm_halo(const Ranks &ms, const std::vector<const char *> &names) {
  check_comm(ms);
  qgcm_model *m0 = ms[0];
  if (m0->nranks == 1) return;
  if (peer_active(m0)) {
    const Grid &g = m0->go;
    HaloArgs a;
    a.c = m0->peer;
    a.c.epoch = ++m0->epoch_halo;
    a.nfl = 0; a.ld = g.ld; a.nrows = HALO; a.own0 = g.own0; a.own1 = g.own1;
    for (const char *nm : names) {
      qgcm_model::Field &f = m0->fields.at(nm);
      if (f.ld != g.ld) throw std::runtime_error("comm_halo: field pitch differs from the grid pitch");
      for (int k = 0; k < f.nl; ++k) {
        if (a.nfl >= 24 || (a.nfl + 1) * HALO > PEER_HALO_ROWS) throw std::runtime_error("comm_halo: too many field layers for the mailbox");
        a.base[a.nfl] = f.d + (size_t)k * f.lsz;
        a.ny[a.nfl] = f.ny;
        ++a.nfl;
      }
    }
    a.ticket = m0->d_ticket2 + 1;
    a.err = m0->d_peer_err;
    QG_LAUNCH(m0, "k_halo_peer", dim3(a.nfl * HALO * HALO_XS, 2), 256, 0, k_halo_peer, a);
    return;
  }
  const bool loop = ms.size() > 1;
  if (!loop) nccl_ok(nccl().GroupStart(), "ncclGroupStart");
  for (qgcm_model *m : ms) {
    const Grid &g = m->go;
    for (const char *nm : names) {
      qgcm_model::Field &f = m->fields.at(nm);
      for (int k = 0; k < f.nl; ++k) {
        double *base = f.d + (size_t)k * f.lsz;
        if (m->rank + 1 < m->nranks) {
          // upward: my top owned rows -> bottom halo of rank+1; and its bottom owned rows -> my top halo
          const int ntop = f.ny - g.own1;                     // HALO for p fields, HALO-1 for T fields
          double *send = base + (size_t)(g.own1 - HALO) * f.ld;
          double *recv = base + (size_t)g.own1 * f.ld;
          if (loop) {
            qgcm_model *up = ms[m->rank + 1];
            double *ub = up->fields.at(nm).d + (size_t)k * up->fields.at(nm).lsz;
            QG_CUDA(cudaMemcpyAsync(ub, send, sizeof(double) * HALO * f.ld, cudaMemcpyDeviceToDevice, m0->stream));
            QG_CUDA(cudaMemcpyAsync(recv, ub + (size_t)up->go.own0 * f.ld, sizeof(double) * ntop * f.ld, cudaMemcpyDeviceToDevice,
                                    m0->stream));
          } else {
            nccl_ok(nccl().Send(send, (size_t)HALO * f.ld, NCCL_DOUBLE, m->rank + 1, m->nccl, m->stream), "ncclSend");
            nccl_ok(nccl().Recv(recv, (size_t)ntop * f.ld, NCCL_DOUBLE, m->rank + 1, m->nccl, m->stream), "ncclRecv");
          }
        }
        if (!loop && m->rank > 0) {
          // downward partner of the exchange above (the loopback branch already did both directions)
          const int ntop_below = (f.ny == g.nyp) ? HALO : HALO - 1;
          nccl_ok(nccl().Send(base + (size_t)g.own0 * f.ld, (size_t)ntop_below * f.ld, NCCL_DOUBLE, m->rank - 1, m->nccl, m->stream),
                  "ncclSend");
          nccl_ok(nccl().Recv(base, (size_t)HALO * f.ld, NCCL_DOUBLE, m->rank - 1, m->nccl, m->stream), "ncclRecv");
        }
      }
    }
  }
  if (!loop) nccl_ok(nccl().GroupEnd(), "ncclGroupEnd");
}

// ---------------------------------------------------------------- drivers
// oml + qgostep + ocinvq + ocqbdy (src/q-gcm.F:1229-1249) over y-slabs
void slab_ocean_step(const Ranks &ms) {
  check_comm(ms);
  // with the peer-memory transport the kernels of the step carry the exchanges themselves:
  // k_oml_reduce and k_inv_partials all-reduce their sums, k_slab_push/k_slab_solve move the
  // slab rows, k_halo_peer the halo rows; otherwise NCCL (or the loopback copies) sit between
  const bool peer = peer_active(ms[0]);
  for (qgcm_model *m : ms) oml_phase_a(m);
  if (!peer) comm_allreduce_cv(ms, 0, 3);
  for (qgcm_model *m : ms) {
    oml_phase_b(m);
    launch_qgostep(m);
    ocinvq_phase_a(m);
  }
  comm_allgather_slab(ms);
  for (qgcm_model *m : ms) ocinvq_phase_b(m);
  if (!peer) comm_allreduce_cv(ms, 3, 1 + ms[0]->go.nl);
  for (qgcm_model *m : ms) {
    ocinvq_phase_c(m);
    launch_ocqbdy(m, m->F("qo"), m->F("po"));
  }
  comm_halo(ms, {"po", "qo", "sst"});
  if (peer) check_peer_err(ms[0]);      // host-mapped flag: no synchronisation, checked every step
}

void slab_constr(const Ranks &ms) {
  check_comm(ms);
  std::vector<std::vector<double>> v(ms.size());
  for (size_t r = 0; r < ms.size(); ++r) constr_ocean_share(ms[r], v[r]);
  comm_allreduce_host(ms, v);
  for (size_t r = 0; r < ms.size(); ++r) constr_ocean_store(ms[r], v[r]);
}

void slab_homsol(const Ranks &ms) {
  check_comm(ms);
  for (qgcm_model *m : ms) homsol_box_a(m);
  comm_allgather_slab(ms);
  std::vector<std::vector<double>> v(ms.size());
  for (size_t r = 0; r < ms.size(); ++r) homsol_box_b(ms[r], v[r]);
  comm_allreduce_host(ms, v);
  for (size_t r = 0; r < ms.size(); ++r) homsol_box_c(ms[r], v[r]);
}

// q from p on both time levels (src/q-gcm.F:719-732); the 5-point stencil spoils one row at an
// artificial edge, so the halos of q are refreshed afterwards
void slab_qcomp_ocean(const Ranks &ms) {
  check_comm(ms);
  for (qgcm_model *m : ms) {
    launch_qcomp(m, true, m->F("qo"), m->F("po"));
    launch_qcomp(m, true, m->F("qom"), m->F("pom"));
    launch_ocqbdy(m, m->F("qo"), m->F("po"));
    launch_ocqbdy(m, m->F("qom"), m->F("pom"));
  }
  comm_halo(ms, {"qo", "qom"});
}

void slab_tlavg_ocean(const Ranks &ms) {
  for (qgcm_model *m : ms) launch_tlavg_ocean(m);
}

// in-process loopback group: every rank on one device and one stream
void group_create(qgcm_model **models, int n) {
  if (n < 2 || n > 8) throw std::runtime_error("qgcm_group_create: 2..8 ranks");
  Ranks ms(models, models + n);
  for (int r = 0; r < n; ++r) {
    if (!ms[r] || ms[r]->nranks != n || ms[r]->rank != r)
      throw std::runtime_error("qgcm_group_create: models must be ranks 0..n-1 of an n-rank partition, in order");
    if (ms[r]->cfg.device != ms[0]->cfg.device) throw std::runtime_error("qgcm_group_create: loopback ranks share one device");
  }
  for (int r = 0; r < n; ++r) {
    QG_CUDA(cudaStreamSynchronize(ms[r]->stream));
    if (r > 0) {
      QG_CUDA(cudaStreamDestroy(ms[r]->stream));
      ms[r]->stream = ms[0]->stream;
      ms[r]->shared_stream = true;
    }
    ms[r]->peers = ms;
  }
}

}  // namespace qg
