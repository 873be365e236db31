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

Ranks ranks_of(qgcm_model *m) {
  if (!m->peers.empty()) return m->peers;
  return Ranks{m};
}

static void check_comm(const Ranks &ms) {
  qgcm_model *m = ms[0];
  if (m->nranks == 1) return;
  if (ms.size() == 1 && !m->nccl)
    throw std::runtime_error("y-slab model has no communicator: call qgcm_comm_init_nccl or qgcm_group_create first");
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
  if (n > 24) throw std::runtime_error("comm_allreduce_host: payload too large");
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
  for (qgcm_model *m : ms) oml_phase_a(m);
  comm_allreduce_cv(ms, 0, 3);
  for (qgcm_model *m : ms) {
    oml_phase_b(m);
    launch_qgostep(m);
    ocinvq_phase_a(m);
  }
  comm_allgather_slab(ms);
  for (qgcm_model *m : ms) ocinvq_phase_b(m);
  comm_allreduce_cv(ms, 3, 1 + ms[0]->go.nl);
  for (qgcm_model *m : ms) {
    ocinvq_phase_c(m);
    launch_ocqbdy(m, m->F("qo"), m->F("po"));
  }
  comm_halo(ms, {"po", "qo", "sst"});
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
