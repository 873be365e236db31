// C ABI of libqgcm_b200.so (include/qgcm_b200.h): lifetime, state transfer and the
// main-loop procedures of src/q-gcm.F:1222-1269, :1328-1407.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "qgcm_internal.h"

static thread_local std::string g_err;

#define QG_TRY(...)                   \
  try {                               \
    __VA_ARGS__;                      \
    return 0;                         \
  } catch (const std::exception &e) { \
    g_err = e.what();                 \
    return 1;                         \
  }

namespace qg {

void *dalloc(qgcm_model *m, size_t bytes) {
  void *p = nullptr;
  QG_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
  // zero-fill, complete before anything else touches the buffer: a memset on the legacy stream is
  // asynchronous and not ordered with the model's non-blocking stream (under GPU contention it
  // landed after the first kernel that filled a table).  Allocation is start-up work.
  QG_CUDA(cudaMemset(p, 0, bytes ? bytes : 8));
  QG_CUDA(cudaDeviceSynchronize());
  m->allocs.push_back(p);
  return p;
}

static cudaEvent_t prof_event(qgcm_model *m) {
  if (!m->prof_pool.empty()) {
    cudaEvent_t e = m->prof_pool.back();
    m->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  QG_CUDA(cudaEventCreate(&e));
  return e;
}
void prof_begin(qgcm_model *m, const char *name) {
  if (!m->prof) return;
  qgcm_model::ProfRec r;
  r.name = name;
  r.e0 = prof_event(m);
  r.e1 = prof_event(m);
  QG_CUDA(cudaEventRecord(r.e0, m->stream));
  m->prof_recs.push_back(r);
}
void prof_end(qgcm_model *m) {
  if (!m->prof) return;
  QG_CUDA(cudaEventRecord(m->prof_recs.back().e1, m->stream));
}

void launch_check(qgcm_model *m, const char *name) {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) {
    static const bool dbg = getenv("QGCM_DEBUG_SYNC") != nullptr;   // serialise to pin asynchronous faults
    if (!dbg) return;
    e = cudaStreamSynchronize(m->stream);
    if (e == cudaSuccess) return;
  }
  cudaGetLastError();
  throw std::runtime_error(std::string("kernel ") + name + ": " + cudaGetErrorString(e));
}

void add_field(qgcm_model *m, const char *name, int nx, int ny, int nl, int ld, size_t lsz, const Grid *slab) {
  qgcm_model::Field f;
  f.nx = nx; f.ny = ny; f.nl = nl; f.ld = ld; f.lsz = lsz;
  f.nyg = ny; f.joff = 0; f.o0 = 0; f.o1 = ny;
  if (slab && ld) {   // gridded field of a y-slab: host arrays are global, the device holds local rows
    f.nyg = slab->nyp_g - (slab->nyp - ny);   // p fields: nyp_g, T fields: nyp_g - 1
    f.joff = slab->jg0;
    f.o0 = slab->own0;
    f.o1 = std::min(slab->own1, ny);
  }
  // gridded fields (ld != 0) use the grid's p-row layer stride lsz for T and p arrays
  // alike; dense (ld == 0) arrays are nx*ny*nl
  f.elems = ld ? lsz * nl : (size_t)nx * ny * nl;
  f.d = (double *)dalloc(m, sizeof(double) * f.elems);
  m->fields[name] = f;
}

static void make_grid(Grid &g, int nxt, int nyt, int nl, int cyclic, double dx, double fnot, double dt) {
  g.nxt = nxt; g.nyt = nyt; g.nxp = nxt + 1; g.nyp = nyt + 1; g.nl = nl;
  g.ld = ((g.nxp + 15) / 16) * 16;
  g.lsz = (size_t)g.ld * g.nyp;
  g.cyclic = cyclic;
  g.dx = dx; g.dxm2 = 1.0 / (dx * dx); g.hdxm1 = 0.5 / dx; g.rdxf0 = 1.0 / (dx * fnot);
  g.norm = 1.0 / ((double)nxt * nyt);
  g.xl = nxt * dx; g.yl = nyt * dx;
  g.tdt = 2.0 * dt;
  g.jg0 = 0; g.nyp_g = g.nyp; g.own0 = 0; g.own1 = g.nyp;
}

// y-slab of the global grid for (nranks, rank): owned p rows plus HALO rows on the inner sides
static void make_slab(Grid &g, int nranks, int rank) {
  const Grid full = g;
  int p0, p1;
  slab_bounds(full.nyp, nranks, rank, &p0, &p1);
  const int lo = std::max(0, p0 - HALO), hi = std::min(full.nyp, p1 + HALO);
  g.nyp = hi - lo; g.nyt = g.nyp - 1;
  g.lsz = (size_t)g.ld * g.nyp;
  g.jg0 = lo; g.nyp_g = full.nyp; g.own0 = p0 - lo; g.own1 = p1 - lo;
  // norm, yl keep the global values
}

static double *upload(qgcm_model *m, const std::vector<double> &v) {
  double *d = (double *)dalloc(m, sizeof(double) * v.size());
  QG_CUDA(cudaMemcpy(d, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
  return d;
}

static qgcm_model *create(const qgcm_config *cfg) {
  if (!cfg) throw std::runtime_error("qgcm_create: null config");
  if (cfg->abi_version != QGCM_ABI_VERSION || cfg->struct_bytes != (int)sizeof(qgcm_config))
    throw std::runtime_error("qgcm_create: qgcm_config ABI mismatch (version/size)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw std::runtime_error("qgcm_create: no CUDA device; libqgcm_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) throw std::runtime_error("qgcm_create: bad device ordinal");
  QG_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  QG_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) throw std::runtime_error("qgcm_create: kernels are built for sm_100a (Blackwell) only");
  if (cfg->nlo < 2 || cfg->nlo > QGCM_NLMAX || cfg->nla < 2 || cfg->nla > QGCM_NLMAX)
    throw std::runtime_error("qgcm_create: layer count out of range");
  qgcm_model *m = new qgcm_model();
  try {
    m->cfg = *cfg;
    m->flags = cfg->flags;
    m->ocean_only = cfg->flags & QGCM_OCEAN_ONLY;
    m->atmos_only = cfg->flags & QGCM_ATMOS_ONLY;
    m->cyclic = cfg->flags & QGCM_CYCLIC_OCEAN;
    m->sb_hflux = cfg->flags & QGCM_SB_HFLUX;
    m->nb_hflux = cfg->flags & QGCM_NB_HFLUX;
    m->tau_udiff = cfg->flags & QGCM_TAU_UDIFF;
    m->has_ocean = !m->atmos_only;
    m->has_atmos = !m->ocean_only;
    QG_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    m->fnot = cfg->fnot; m->beta = cfg->beta;
    // src/q-gcm.F:377-441
    const double dxa = cfg->ndxr * cfg->dxo;
    m->dta = cfg->dta; m->dto = cfg->nstr * cfg->dta;
    m->rrcpat = 1.0 / (cfg->rhoat * cfg->cpat);
    m->rrcpoc = 1.0 / (cfg->rhooc * cfg->cpoc);
    m->raoro = cfg->rhoat / cfg->rhooc;
    make_grid(m->go, cfg->nxto, cfg->nyto, cfg->nlo, m->cyclic, cfg->dxo, cfg->fnot, m->dto);
    m->nranks = std::max(1, (int)cfg->nranks);
    m->rank = m->nranks > 1 ? cfg->rank : 0;
    if (m->nranks > 1) {
      if (!m->ocean_only || m->cyclic)
        throw std::runtime_error("qgcm_create: y-slab partitioning covers the ocean-only box decks (NAtl); "
                                 "coupled and channel decks run on one GPU");
      if (m->rank < 0 || m->rank >= m->nranks) throw std::runtime_error("qgcm_create: bad rank");
      if ((cfg->nyto + 1) / m->nranks < 4 * HALO) throw std::runtime_error("qgcm_create: slabs thinner than 4 halo widths");
      make_slab(m->go, m->nranks, m->rank);
    }
    make_grid(m->ga, cfg->nxta, cfg->nyta, cfg->nla, 1, dxa, cfg->fnot, m->dta);
    const double yla = cfg->nyta * dxa;
    for (int k = 0; k < NLMAX; ++k) {
      m->lo.h[k] = cfg->hoc[k]; m->lo.gp[k] = cfg->gpoc[k]; m->lo.ah2[k] = cfg->ah2oc[k]; m->lo.ah4[k] = cfg->ah4oc[k];
      m->lo.rdm2[k] = cfg->rdm2oc[k];
      m->la.h[k] = cfg->hat[k]; m->la.gp[k] = cfg->gpat[k]; m->la.ah2[k] = 0.0; m->la.ah4[k] = cfg->ah4at[k];
      m->la.rdm2[k] = cfg->rdm2at[k];
    }
    for (int i = 0; i < NLMAX * NLMAX; ++i) {
      m->lo.amat[i] = cfg->amatoc[i]; m->lo.ctl2m[i] = cfg->ctl2moc[i]; m->lo.ctm2l[i] = cfg->ctm2loc[i];
      m->la.amat[i] = cfg->amatat[i]; m->la.ctl2m[i] = cfg->ctl2mat[i]; m->la.ctm2l[i] = cfg->ctm2lat[i];
    }
    m->d_scal = (qgcm_scalars *)dalloc(m, sizeof(qgcm_scalars));
    m->d_coef = (double *)dalloc(m, sizeof(double) * 128);
    m->d_cv = (double *)dalloc(m, sizeof(double) * 32);
    m->d_ticket = (unsigned int *)dalloc(m, sizeof(unsigned int) * 4);
    size_t red = 0;
    if (m->has_ocean) {
      const Grid &g = m->go;
      std::vector<double> ypo(g.nyp), yporel(g.nyp), ytorel(g.nyt);
      for (int j = 1; j <= g.nyp; ++j) {
        ypo[j - 1] = (cfg->ny1 - 1) * dxa + (g.jg0 + j - 1) * g.dx;
        yporel[j - 1] = ypo[j - 1] - 0.5 * yla;
      }
      for (int j = 1; j <= g.nyt; ++j) ytorel[j - 1] = (ypo[j - 1] + 0.5 * g.dx) - 0.5 * yla;
      m->h_ypo = ypo;
      m->yporel = upload(m, yporel);
      m->ytorel = upload(m, ytorel);
      for (const char *n : {"po", "pom", "qo", "qom"}) add_field(m, n, g.nxp, g.nyp, g.nl, g.ld, g.lsz, &g);
      for (const char *n : {"wekpo", "entoc", "ddynoc", "tauxo", "tauyo"}) add_field(m, n, g.nxp, g.nyp, 1, g.ld, g.lsz, &g);
      for (const char *n : {"sst", "sstm", "wekto", "fnetoc"}) add_field(m, n, g.nxt, g.nyt, 1, g.ld, g.lsz, &g);
      add_field(m, "sstbar", g.nyp_g - 1, 1, 1, 0);
      if (m->cyclic) {
        add_field(m, "pch1oc", g.nyp, g.nl - 1, 1, 0);
        add_field(m, "pch2oc", g.nyp, g.nl - 1, 1, 0);
        add_field(m, "pbhoc", g.nyp, 1, 1, 0);
      } else {
        add_field(m, "ochom", g.nxp, g.nyp, g.nl - 1, g.ld, g.lsz, &g);
      }
      m->wrk_o = (double *)dalloc(m, sizeof(double) * g.lsz * g.nl);
      m->xfo = (double *)dalloc(m, sizeof(double) * g.lsz);
      m->sstnew = (double *)dalloc(m, sizeof(double) * g.lsz);
      helm_plan_create(m, m->hpo, g, m->cyclic ? 1 : 0, cfg->rdm2oc, g.nl);
      const size_t nb = (size_t)((g.nxt + 63) / 64) * ((g.nyt + 7) / 8);   // >= the oml tile count (64 x 12 tiles)
      red = std::max(red, 3 * nb + 4 * (size_t)g.nyp);
    }
    if (m->has_atmos) {
      const Grid &g = m->ga;
      std::vector<double> ypa(g.nyp), yparel(g.nyp), ytarel(g.nyt);
      for (int j = 1; j <= g.nyp; ++j) {
        ypa[j - 1] = (j - 1) * g.dx;
        yparel[j - 1] = ypa[j - 1] - 0.5 * yla;
      }
      for (int j = 1; j <= g.nyt; ++j) ytarel[j - 1] = (ypa[j - 1] + 0.5 * g.dx) - 0.5 * yla;
      m->h_ypa = ypa;
      m->yparel = upload(m, yparel);
      m->ytarel = upload(m, ytarel);
      for (const char *n : {"pa", "pam", "qa", "qam"}) add_field(m, n, g.nxp, g.nyp, g.nl, g.ld, g.lsz);
      for (const char *n : {"wekpa", "entat", "ddynat", "dtopat", "tauxa", "tauya"}) add_field(m, n, g.nxp, g.nyp, 1, g.ld, g.lsz);
      for (const char *n : {"ast", "astm", "hmixa", "hmixam", "wekta", "fnetat", "xc1ast"}) add_field(m, n, g.nxt, g.nyt, 1, g.ld, g.lsz);
      add_field(m, "uekat", g.nxp, g.nyt, 1, g.ld, g.lsz);
      add_field(m, "vekat", g.nxt, g.nyp, 1, g.ld, g.lsz);
      add_field(m, "astbar", g.nyt, 1, 1, 0);
      add_field(m, "pch1at", g.nyp, g.nl - 1, 1, 0);
      add_field(m, "pch2at", g.nyp, g.nl - 1, 1, 0);
      add_field(m, "pbhat", g.nyp, 1, 1, 0);
      m->wrk_a = (double *)dalloc(m, sizeof(double) * g.lsz * g.nl);
      m->xfa = (double *)dalloc(m, sizeof(double) * g.lsz);
      m->astnew = (double *)dalloc(m, sizeof(double) * g.lsz);
      m->hmnew = (double *)dalloc(m, sizeof(double) * g.lsz);
      helm_plan_create(m, m->hpa, g, 1, cfg->rdm2at, g.nl);
      const size_t nb = (size_t)((g.nxt + 63) / 64) * ((g.nyt + 3) / 4);   // aml tiles are 64 x 4
      red = std::max(red, 3 * nb + 4 * (size_t)g.nyp);
    }
    m->red_elems = red + 512;
    m->d_red = (double *)dalloc(m, sizeof(double) * m->red_elems);
    // the atmosphere's mixed layer has its own scratch: a coupled cycle runs the ocean step beside the
    // atmosphere steps (run_cycle below)
    m->d_red_a = (cfg->flags & QGCM_OCEAN_ONLY) ? m->d_red : (double *)dalloc(m, sizeof(double) * m->red_elems);
    QG_CUDA(cudaStreamSynchronize(m->stream));
  } catch (...) {
    for (void *p : m->allocs) cudaFree(p);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    throw;
  }
  return m;
}

static qgcm_model::Field &lookup(qgcm_model *m, const char *name, int64_t n) {
  auto it = m->fields.find(name);
  if (it == m->fields.end()) throw std::runtime_error(std::string("unknown field '") + name + "'");
  qgcm_model::Field &f = it->second;
  if (n >= 0 && n != (int64_t)f.nx * f.nyg * f.nl)
    throw std::runtime_error(std::string("field '") + name + "': element count mismatch");
  return f;
}

static void copy_field(qgcm_model *m, qgcm_model::Field &f, double *host, bool to_device, bool sync = true) {
  if (f.ld == 0) {
    const size_t bytes = sizeof(double) * (size_t)f.nx * f.ny * f.nl;
    if (to_device) QG_CUDA(cudaMemcpyAsync(f.d, host, bytes, cudaMemcpyHostToDevice, m->stream));
    else QG_CUDA(cudaMemcpyAsync(host, f.d, bytes, cudaMemcpyDeviceToHost, m->stream));
  } else {
    const size_t lsz = f.lsz;   // device layer stride
    // host arrays are the reference's global arrays; a slab uploads its local rows (owned +
    // halo) and downloads the rows it owns, leaving the rest of the host array untouched
    for (int k = 0; k < f.nl; ++k) {
      double *d = f.d + (size_t)k * lsz;
      double *h = host + (size_t)k * f.nx * f.nyg + (size_t)f.joff * f.nx;
      if (to_device)
        QG_CUDA(cudaMemcpy2DAsync(d, sizeof(double) * f.ld, h, sizeof(double) * f.nx, sizeof(double) * f.nx, f.ny,
                                  cudaMemcpyHostToDevice, m->stream));
      else
        QG_CUDA(cudaMemcpy2DAsync(h + (size_t)f.o0 * f.nx, sizeof(double) * f.nx, d + (size_t)f.o0 * f.ld, sizeof(double) * f.ld,
                                  sizeof(double) * f.nx, f.o1 - f.o0, cudaMemcpyDeviceToHost, m->stream));
    }
  }
  if (!sync) return;
  QG_CUDA(cudaStreamSynchronize(m->stream));
  if (!to_device) check_peer_err(m);      // never hand back state a timed-out exchange has spoilt
}

static void invalidate_graphs(qgcm_model *m);

static void note_topography(qgcm_model *m, const char *name, const double *host, int64_t n) {
  const bool oc = std::strcmp(name, "ddynoc") == 0, at = std::strcmp(name, "ddynat") == 0;
  if (oc || at) {      // flat bottom / no orography: remember it (invert.cu skips the field)
    bool flat = true;
    for (int64_t i = 0; i < n && flat; ++i) flat = (host[i] == 0.0);
    (oc ? m->ddynoc_flat : m->ddynat_flat) = flat;
  }
}

// n fields, one synchronisation: with page-locked host arrays (qgcm_host_register) the copies are
// back-to-back DMAs on the model's stream
static void copy_fields(qgcm_model *m, int n, const char *const *names, double *const *hosts, const int64_t *counts, bool to_device) {
  if (n < 0 || (n > 0 && (!names || !hosts || !counts))) throw std::runtime_error("qgcm_get_fields/qgcm_set_fields: bad argument list");
  for (int i = 0; i < n; ++i) lookup(m, names[i], counts[i]);      // validate everything before the first byte moves
  for (int i = 0; i < n; ++i) {
    copy_field(m, lookup(m, names[i], counts[i]), hosts[i], to_device, false);
    if (to_device) {
      note_topography(m, names[i], hosts[i], counts[i]);
      if (std::strncmp(names[i], "ddyn", 4) == 0) invalidate_graphs(m);
    }
  }
  QG_CUDA(cudaStreamSynchronize(m->stream));
  if (!to_device) check_peer_err(m);
}

void launch_xforc(qgcm_model *m);
void launch_aml(qgcm_model *m);

// Asynchronous upload of a gridded field from PINNED host memory: the copy runs on a second
// stream into a shadow buffer while the model keeps stepping on the current one; the swap
// happens at qgcm_commit_fields.
static void set_field_async(qgcm_model *m, const char *name, const double *host, int64_t n) {
  qgcm_model::Field &f = lookup(m, name, n);
  if (!f.ld) throw std::runtime_error("qgcm_set_field_async: gridded fields only");
  static const char *ok[] = {"tauxo", "tauyo", "fnetoc", "ddynoc", "tauxa", "tauya", "fnetat", "ddynat", "dtopat"};
  bool allowed = false;
  for (const char *o : ok) allowed = allowed || std::strcmp(o, name) == 0;
  if (!allowed) throw std::runtime_error("qgcm_set_field_async: only forcing fields the step never writes or rotates");
  if (!m->copy_stream) {
    QG_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    QG_CUDA(cudaEventCreateWithFlags(&m->ev_copy, cudaEventDisableTiming));
    QG_CUDA(cudaEventCreateWithFlags(&m->ev_step, cudaEventDisableTiming));
  }
  double *&shadow = m->shadow[name];
  if (!shadow) shadow = (double *)dalloc(m, sizeof(double) * f.elems);
  if (m->pending.empty()) {
    // the shadow buffers were the live ones until the last commit: kernels enqueued before
    // it may still read them
    QG_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_step, 0));
  }
  for (int k = 0; k < f.nl; ++k) {
    const double *h = host + (size_t)k * f.nx * f.nyg + (size_t)f.joff * f.nx;
    QG_CUDA(cudaMemcpy2DAsync(shadow + (size_t)k * f.lsz, sizeof(double) * f.ld, h, sizeof(double) * f.nx, sizeof(double) * f.nx,
                              f.ny, cudaMemcpyHostToDevice, m->copy_stream));
  }
  // a field uploaded twice before one commit overwrites its shadow buffer (same copy stream, in
  // order) and is swapped exactly once
  if (std::find(m->pending.begin(), m->pending.end(), name) == m->pending.end()) m->pending.push_back(name);
  if (std::strcmp(name, "ddynoc") == 0) { m->ddynoc_flat = false; invalidate_graphs(m); }     // contents unknown until inspected: read the field
  if (std::strcmp(name, "ddynat") == 0) { m->ddynat_flat = false; invalidate_graphs(m); }
}
static void commit_fields(qgcm_model *m) {
  if (m->pending.empty()) return;
  QG_CUDA(cudaEventRecord(m->ev_copy, m->copy_stream));
  QG_CUDA(cudaStreamWaitEvent(m->stream, m->ev_copy, 0));
  for (const std::string &nm : m->pending) std::swap(m->fields.at(nm).d, m->shadow.at(nm));
  m->pending.clear();
  QG_CUDA(cudaEventRecord(m->ev_step, m->stream));   // everything that reads the retired buffers is before this point
}

static void ocean_step_eager(qgcm_model *m) {
  launch_oml(m);
  launch_qgostep(m);
  launch_ocinvq(m);
  launch_ocqbdy(m, m->F("qo"), m->F("po"));
}
static void atmos_step_eager(qgcm_model *m) {
  launch_aml(m);
  launch_qgastep(m);
  launch_atinvq(m);
  launch_atqzbd(m, m->F("qa"), m->F("pa"));
}

// ---- whole steps as CUDA graphs (single GPU) ----
static std::vector<double *> pointer_state(qgcm_model *m) {
  std::vector<double *> v;
  for (auto &kv : m->fields) v.push_back(kv.second.d);
  v.push_back(m->sstnew); v.push_back(m->astnew); v.push_back(m->hmnew);
  return v;
}
static void set_pointer_state(qgcm_model *m, const std::vector<double *> &v) {
  size_t i = 0;
  for (auto &kv : m->fields) kv.second.d = v[i++];
  m->sstnew = v[i++]; m->astnew = v[i++]; m->hmnew = v[i++];
}
// a captured step bakes in everything the launch code read on the host: drop the graphs whenever
// such state may have changed outside a step (tables, flags, communicators)
static void invalidate_graphs(qgcm_model *m) {
  for (auto &kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
}
// kind: 0 atmosphere step, 1 xforc + ocean step + atmosphere step (coupled), 2 ocean step (ocean only),
// 3 a whole coupled cycle (cycle_body)
template <class Body>
static void run_graphed(qgcm_model *m, int kind, Body body) {
  static const bool off = env_int("QGCM_GRAPH", 1) == 0;
  // the first steps run eagerly: lazily built plans and function attributes are set up in them
  if (off || m->prof || m->nranks > 1 || m->eager_steps[kind] < 2) {
    m->eager_steps[kind]++;
    body();
    return;
  }
  const std::vector<double *> before = pointer_state(m);
  std::string key(1, (char)('0' + kind));
  key.append(reinterpret_cast<const char *>(before.data()), before.size() * sizeof(double *));
  auto it = m->graphs.find(key);
  if (it == m->graphs.end()) {
    const int64_t l0 = m->launches;
    cudaGraph_t g = nullptr;
    QG_CUDA(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
    try {
      body();
    } catch (...) {
      cudaStreamEndCapture(m->stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    QG_CUDA(cudaStreamEndCapture(m->stream, &g));
    qgcm_model::StepGraph sg;
    QG_CUDA(cudaGraphInstantiate(&sg.exec, g, 0));
    QG_CUDA(cudaGraphDestroy(g));
    sg.after = pointer_state(m);
    sg.launches = m->launches - l0;
    m->launches = l0;
    it = m->graphs.emplace(key, sg).first;
  }
  QG_CUDA(cudaGraphLaunch(it->second.exec, m->stream));
  set_pointer_state(m, it->second.after);
  m->launches += it->second.launches;
}

// One coupled cycle, nt = first .. first + nstr - 1 with mod(first, nstr) = 1 (src/q-gcm.F:1220-1269): xforc, the
// ocean step and nstr atmosphere steps.  After xforc the ocean step and the atmosphere steps touch disjoint
// state (the ocean reads tauxo, tauyo, wekto, wekpo, fnetoc and its own fields; the atmosphere reads wekta,
// wekpa, fnetat, tauxa, tauya and its own; the scalars of the two live in separate members of qgcm_scalars), and
// nothing reads the other side before the next xforc.  So the atmosphere steps are a second branch of the cycle's
// graph: their small kernels (385 x 97 points in the double-gyre deck) fill the SMs the ocean kernels' last
// waves leave idle instead of queueing behind them.
static void cycle_body(qgcm_model *m) {
  if (!m->at_stream) {
    // (a higher stream priority for this branch measured no different: 0.667 ms per dg_coupled cycle either way)
    QG_CUDA(cudaStreamCreateWithFlags(&m->at_stream, cudaStreamNonBlocking));
    QG_CUDA(cudaEventCreateWithFlags(&m->ev_atfork, cudaEventDisableTiming));
    QG_CUDA(cudaEventCreateWithFlags(&m->ev_atjoin, cudaEventDisableTiming));
  }
  launch_xforc(m);
  cudaStream_t main = m->stream;
  const bool split = !m->prof && env_int("QGCM_CYCLE_FORK", 1) != 0;      // profiled cycles stay on one stream
  if (split) {
    QG_CUDA(cudaEventRecord(m->ev_atfork, main));
    QG_CUDA(cudaStreamWaitEvent(m->at_stream, m->ev_atfork, 0));
    m->stream = m->at_stream;      // every launch goes through (m)->stream
  }
  try {
    for (int i = 0; i < m->cfg.nstr; ++i) atmos_step_eager(m);
  } catch (...) {
    m->stream = main;
    throw;
  }
  m->stream = main;
  ocean_step_eager(m);
  if (split) {
    QG_CUDA(cudaEventRecord(m->ev_atjoin, m->at_stream));
    QG_CUDA(cudaStreamWaitEvent(main, m->ev_atjoin, 0));
  }
}

static void ocean_step(qgcm_model *m) {
  if (m->nranks > 1) { slab_ocean_step(ranks_of(m)); return; }
  run_graphed(m, 2, [m] { ocean_step_eager(m); });
}
static void atmos_step(qgcm_model *m) { run_graphed(m, 0, [m] { atmos_step_eager(m); }); }

}  // namespace qg

using namespace qg;

extern "C" {

const char *qgcm_last_error(void) { return g_err.c_str(); }
int qgcm_abi_version(void) { return QGCM_ABI_VERSION; }

int qgcm_create(const qgcm_config *cfg, qgcm_model **out) { QG_TRY(*out = create(cfg)); }

int qgcm_destroy(qgcm_model *m) {
  if (!m) return 0;
  cudaSetDevice(m->cfg.device);
  cudaStreamSynchronize(m->stream);
  // a member of an in-process loopback group leaves: the group is dissolved, and the ranks that
  // borrowed this model's stream get one of their own (they may be destroyed in any order)
  for (qgcm_model *p : m->peers) {
    if (p == m) continue;
    p->peers.clear();
    if (p->shared_stream && p->stream == m->stream && !m->shared_stream) {
      p->stream = nullptr;
      cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
      p->shared_stream = false;
    }
  }
  m->peers.clear();
  peer_close(m);
  for (auto &kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);
  for (void *p : m->allocs) cudaFree(p);
  if (m->h_peer_err) cudaFreeHost(m->h_peer_err);
  if (m->nccl) { try { nccl_destroy(m); } catch (...) {} }
  if (m->copy_stream) { cudaStreamSynchronize(m->copy_stream); cudaStreamDestroy(m->copy_stream); cudaEventDestroy(m->ev_copy); cudaEventDestroy(m->ev_step); }
  if (m->at_stream) { cudaStreamSynchronize(m->at_stream); cudaStreamDestroy(m->at_stream); cudaEventDestroy(m->ev_atfork); cudaEventDestroy(m->ev_atjoin); }
  if (m->side_stream) { cudaStreamSynchronize(m->side_stream); cudaStreamDestroy(m->side_stream); cudaEventDestroy(m->ev_fork); cudaEventDestroy(m->ev_join); }
  if (!m->shared_stream) cudaStreamDestroy(m->stream);
  delete m;
  return 0;
}

int qgcm_field_size(qgcm_model *m, const char *name, int64_t *n) {
  QG_TRY(qgcm_model::Field &f = lookup(m, name, -1); *n = (int64_t)f.nx * f.nyg * f.nl);
}
int qgcm_set_field(qgcm_model *m, const char *name, const double *host, int64_t n) {
  QG_TRY({
    copy_field(m, lookup(m, name, n), const_cast<double *>(host), true);
    note_topography(m, name, host, n);
    if (std::strncmp(name, "ddyn", 4) == 0) invalidate_graphs(m);      // the flat-bottom flag is baked into captured steps
  });
}
int qgcm_host_register(void *host, int64_t bytes) {
  QG_TRY({
    if (!host || bytes <= 0) throw std::runtime_error("qgcm_host_register: bad range");
    QG_CUDA(cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable));
  });
}
int qgcm_host_unregister(void *host) { QG_TRY(QG_CUDA(cudaHostUnregister(host))); }
int qgcm_get_fields(qgcm_model *m, int32_t n, const char *const *names, double *const *hosts, const int64_t *counts) {
  QG_TRY(copy_fields(m, n, names, hosts, counts, false));
}
int qgcm_set_fields(qgcm_model *m, int32_t n, const char *const *names, const double *const *hosts, const int64_t *counts) {
  QG_TRY(copy_fields(m, n, names, const_cast<double *const *>(hosts), counts, true));
}
int qgcm_get_field(qgcm_model *m, const char *name, double *host, int64_t n) {
  QG_TRY(copy_field(m, lookup(m, name, n), host, false));
}
int qgcm_set_field_async(qgcm_model *m, const char *name, const double *host, int64_t n) { QG_TRY(set_field_async(m, name, host, n)); }
int qgcm_commit_fields(qgcm_model *m) { QG_TRY(commit_fields(m)); }
int qgcm_set_scalars(qgcm_model *m, const qgcm_scalars *s) {
  QG_TRY(QG_CUDA(cudaMemcpyAsync(m->d_scal, s, sizeof(*s), cudaMemcpyHostToDevice, m->stream));
         QG_CUDA(cudaStreamSynchronize(m->stream)));
}
int qgcm_get_scalars(qgcm_model *m, qgcm_scalars *s) {
  QG_TRY(QG_CUDA(cudaMemcpyAsync(s, m->d_scal, sizeof(*s), cudaMemcpyDeviceToHost, m->stream));
         QG_CUDA(cudaStreamSynchronize(m->stream));
         check_peer_err(m));
}
int qgcm_sync(qgcm_model *m) {
  QG_TRY({
    QG_CUDA(cudaStreamSynchronize(m->stream));
    check_peer_err(m);      // a peer-memory exchange gave up waiting for another rank
  });
}

int qgcm_constr(qgcm_model *m) { QG_TRY(invalidate_graphs(m); launch_constr(m)); }
int qgcm_homsol(qgcm_model *m) { QG_TRY(invalidate_graphs(m); launch_homsol(m)); }
int qgcm_qcomp_ocean(qgcm_model *m) {
  QG_TRY(if (m->nranks > 1) { slab_qcomp_ocean(ranks_of(m)); } else {
         launch_qcomp(m, true, m->F("qo"), m->F("po")); launch_qcomp(m, true, m->F("qom"), m->F("pom"));
         launch_ocqbdy(m, m->F("qo"), m->F("po")); launch_ocqbdy(m, m->F("qom"), m->F("pom")); });
}
int qgcm_qcomp_atmos(qgcm_model *m) {
  QG_TRY(launch_qcomp(m, false, m->F("qa"), m->F("pa")); launch_qcomp(m, false, m->F("qam"), m->F("pam"));
         launch_atqzbd(m, m->F("qa"), m->F("pa")); launch_atqzbd(m, m->F("qam"), m->F("pam")));
}

int qgcm_helmholtz(qgcm_model *m, int which, double *wrk, const double *b) {
  QG_TRY({
    const bool atmos = which != 0;
    if (atmos ? !m->has_atmos : !m->has_ocean) throw std::runtime_error("qgcm_helmholtz: grid not present");
    invalidate_graphs(m);
    const Grid &g = atmos ? m->ga : m->go;
    HelmPlan &hp = atmos ? m->hpa : m->hpo;
    const LayerConsts &lc = atmos ? m->la : m->lo;
    double *dw = atmos ? m->wrk_a : m->wrk_o;
    std::vector<double> bb((size_t)g.nl * hp.n);
    for (int q = 0; q < g.nl; ++q) std::memcpy(&bb[(size_t)q * hp.n], b, sizeof(double) * hp.n);
    helm_set_diag(m, hp, bb.data());
    QG_CUDA(cudaMemcpy2DAsync(dw, sizeof(double) * g.ld, wrk, sizeof(double) * g.nxp, sizeof(double) * g.nxp, g.nyp,
                              cudaMemcpyHostToDevice, m->stream));
    hp.walls_dirty = true;
    helm_solve(m, hp, dw, 1);
    QG_CUDA(cudaMemcpy2DAsync(wrk, sizeof(double) * g.nxp, dw, sizeof(double) * g.ld, sizeof(double) * g.nxp, g.nyp,
                              cudaMemcpyDeviceToHost, m->stream));
    QG_CUDA(cudaStreamSynchronize(m->stream));
    // restore the per-mode operators used by ocinvq/atinvq
    std::vector<double> bd2(hp.n);
    const double PI = 3.14159265358979324, TWOPI = 6.28318530717958648, a = hp.a;
    if (hp.kind == 1) {
      for (int i = 2; i <= hp.n / 2; ++i) {
        int i1 = 2 * i - 1;
        bd2[i1 - 2] = -2.0 * a + 2.0 * g.dxm2 * (cos((i - 1) * TWOPI / hp.n) - 1.0);
        bd2[i1 - 1] = bd2[i1 - 2];
      }
      bd2[0] = -2.0 * a;
      bd2[hp.n - 1] = -2.0 * a - 4.0 * g.dxm2;
    } else {
      for (int i = 2; i <= hp.n; ++i) bd2[i - 2] = -2.0 * a + 2.0 * g.dxm2 * (cos((i - 1) * PI / hp.n) - 1.0);
      bd2[hp.n - 1] = 0.0;
    }
    for (int q = 0; q < g.nl; ++q)
      for (int i = 0; i < hp.n; ++i) bb[(size_t)q * hp.n + i] = bd2[i] - lc.rdm2[q];
    helm_set_diag(m, hp, bb.data());
  });
}

int qgcm_xforc(qgcm_model *m) { QG_TRY(launch_xforc(m)); }
int qgcm_oml(qgcm_model *m) { QG_TRY(launch_oml(m)); }
int qgcm_qgostep(qgcm_model *m) { QG_TRY(launch_qgostep(m)); }
int qgcm_ocinvq(qgcm_model *m) { QG_TRY(launch_ocinvq(m)); }
int qgcm_ocqbdy(qgcm_model *m) { QG_TRY(launch_ocqbdy(m, m->F("qo"), m->F("po"))); }
int qgcm_aml(qgcm_model *m) { QG_TRY(launch_aml(m)); }
int qgcm_qgastep(qgcm_model *m) { QG_TRY(launch_qgastep(m)); }
int qgcm_atinvq(qgcm_model *m) { QG_TRY(launch_atinvq(m)); }
int qgcm_atqzbd(qgcm_model *m) { QG_TRY(launch_atqzbd(m, m->F("qa"), m->F("pa"))); }
int qgcm_tlavg_ocean(qgcm_model *m) { QG_TRY(slab_tlavg_ocean(ranks_of(m))); }
int qgcm_tlavg_atmos(qgcm_model *m) { QG_TRY(launch_tlavg_atmos(m)); }
int qgcm_ocean_step(qgcm_model *m) { QG_TRY(ocean_step(m)); }
int qgcm_atmos_step(qgcm_model *m) { QG_TRY(atmos_step(m)); }

// src/q-gcm.F:1220-1408.  nstr == 1: mod(nt,1).eq.1 never holds in the reference, so the
// shipped NAtl 1 km deck never steps its ocean; here the ocean steps
// on every nt in that case.
int qgcm_run(qgcm_model *m, int64_t nt_first, int64_t nt_last) {
  QG_TRY({
    const int nstr = m->cfg.nstr;
    for (int64_t nt = nt_first; nt <= nt_last; ++nt) {
      const bool ocstep = (nstr == 1) ? true : (nt % nstr == 1);
      const bool coupled1 = m->has_atmos && m->has_ocean && m->nranks == 1;
      // a whole cycle in one graph when no averaging falls inside it (tlavg_ocean follows the nt of an ocean
      // step on its 1-in-25 cadence, tlavg_atmos every 100th nt; the k247 accumulator wants po between steps)
      bool whole = ocstep && coupled1 && nstr > 1 && nt + nstr - 1 <= nt_last && !(m->flags & QGCM_OCNC_AVG_K247) &&
                   (nt - 1) % (25 * (int64_t)nstr) != 0;
      for (int64_t k = nt; whole && k < nt + nstr; ++k)
        if ((k - 1) % 100 == 0) whole = false;
      if (whole) {
        run_graphed(m, 3, [m] { cycle_body(m); });
        nt += nstr - 1;
        continue;
      }
      if (ocstep && coupled1) {
        // coupled: xforc + ocean step + this nt's atmosphere step as one graph (67 launches)
        run_graphed(m, 1, [m] { launch_xforc(m); ocean_step_eager(m); atmos_step_eager(m); });
        if (m->flags & QGCM_OCNC_AVG_K247) launch_avg_ocn_k247(m);      // src/q-gcm.F:1250-1252 (po is final after the ocean step)
      } else {
        if (ocstep) {
          if (m->has_atmos) launch_xforc(m);
          if (m->has_ocean) {
            ocean_step(m);
            if (m->flags & QGCM_OCNC_AVG_K247)      // src/q-gcm.F:1250-1252; every rank of a loopback group
              for (qgcm_model *r : ranks_of(m)) launch_avg_ocn_k247(r);
          }
        }
        if (m->has_atmos) atmos_step(m);
      }
      if (m->has_ocean && ((nt - 1) % (25 * (int64_t)nstr) == 0)) slab_tlavg_ocean(ranks_of(m));
      if (m->has_atmos && ((nt - 1) % 100 == 0)) launch_tlavg_atmos(m);
    }
  });
}

// ---- y-slab multi-GPU (ocean-only box decks) ----
int qgcm_slab_bounds(int32_t nyp_global, int32_t nranks, int32_t rank, int32_t *jp0, int32_t *nyp_own) {
  QG_TRY({
    if (nranks < 1 || rank < 0 || rank >= nranks || nyp_global < nranks) throw std::runtime_error("qgcm_slab_bounds: bad arguments");
    int p0, p1;
    slab_bounds(nyp_global, nranks, rank, &p0, &p1);
    *jp0 = p0;
    *nyp_own = p1 - p0;
  });
}
int qgcm_nccl_unique_id(void *id128) { QG_TRY(nccl_unique_id(id128)); }
int qgcm_comm_init_nccl(qgcm_model *m, const void *id128) { QG_TRY(nccl_init(m, id128)); }
int qgcm_group_create(qgcm_model **models, int32_t n) { QG_TRY(group_create(models, n)); }
int qgcm_peer_handle(qgcm_model *m, void *handle64) { QG_TRY(peer_export(m, handle64)); }
int qgcm_comm_init_peer(qgcm_model *m, const void *handles, int32_t n) { QG_TRY(peer_init(m, handles, n)); }
int qgcm_comm_transport(qgcm_model *m, int32_t kind) { QG_TRY(set_transport(m, kind)); }
int qgcm_comm_peer_timeout(qgcm_model *m, double seconds) { QG_TRY(peer_set_timeout(m, seconds)); }
int qgcm_comm_close_peer(qgcm_model *m) {
  QG_TRY({
    QG_CUDA(cudaStreamSynchronize(m->stream));
    peer_close(m);
  });
}

int qgcm_valids(qgcm_model *m, qgcm_valids_report *rep) { QG_TRY(launch_valids(m, rep)); }

int qgcm_tavini(qgcm_model *m) { QG_TRY(launch_tavini(m)); }
int qgcm_tavocn(qgcm_model *m) { QG_TRY(launch_tavocn(m)); }
int qgcm_tavatm(qgcm_model *m) { QG_TRY(launch_tavatm(m)); }
int qgcm_avg_ocn_k247(qgcm_model *m) { QG_TRY(launch_avg_ocn_k247(m)); }
int qgcm_tav_counts(qgcm_model *m, int32_t *nsumat, int32_t *nsumoc, int32_t *nsum_ocavg) {
  QG_TRY(*nsumat = m->nsumat; *nsumoc = m->nsumoc; *nsum_ocavg = m->nsum_ocavg);
}
int qgcm_field_sub_size(qgcm_model *m, const char *name, int32_t nsk, int64_t *n) { QG_TRY(field_sub_size(m, name, nsk, n)); }
int qgcm_get_field_sub(qgcm_model *m, const char *name, int32_t nsk, double *host, int64_t n) {
  QG_TRY(get_field_sub(m, name, nsk, host, n));
}

int qgcm_monnc_ocean(qgcm_model *m, qgcm_monitor_ocean *rep) { QG_TRY(launch_monnc_ocean(m, rep)); }
int qgcm_monnc_atmos(qgcm_model *m, qgcm_monitor_atmos *rep) { QG_TRY(launch_monnc_atmos(m, rep)); }
int qgcm_qocdiag_size(qgcm_model *m, int32_t nsko, int64_t *n) { QG_TRY(qocdiag_size(m, nsko, n)); }
int qgcm_qocdiag(qgcm_model *m, int32_t nsko, double *host, int64_t n) { QG_TRY(launch_qocdiag(m, nsko, host, n)); }

int64_t qgcm_launch_count(qgcm_model *m) { return m ? m->launches : 0; }

int qgcm_profile(qgcm_model *m, int enable) {
  QG_TRY({
    QG_CUDA(cudaStreamSynchronize(m->stream));
    for (auto &r : m->prof_recs) { m->prof_pool.push_back(r.e0); m->prof_pool.push_back(r.e1); }
    m->prof_recs.clear();
    m->prof = enable != 0;      // profiled steps are launched kernel by kernel (run_graphed)
  });
}

int qgcm_profile_report(qgcm_model *m, char *buf, int64_t nbuf) {
  QG_TRY({
    QG_CUDA(cudaStreamSynchronize(m->stream));
    std::map<std::string, std::pair<double, long>> acc;
    std::vector<std::string> order;
    for (auto &r : m->prof_recs) {
      float ms = 0.f;
      QG_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
      if (!acc.count(r.name)) order.push_back(r.name);
      acc[r.name].first += ms;
      acc[r.name].second += 1;
    }
    std::string out;
    for (auto &n : order) {
      char line[256];
      snprintf(line, sizeof(line), "%s %ld %.6f\n", n.c_str(), acc[n].second, acc[n].first);
      out += line;
    }
    if ((int64_t)out.size() + 1 > nbuf) throw std::runtime_error("qgcm_profile_report: buffer too small");
    std::memcpy(buf, out.c_str(), out.size() + 1);
  });
}
void *qgcm_stream(qgcm_model *m) { return m ? (void *)m->stream : nullptr; }

}  // extern "C"
