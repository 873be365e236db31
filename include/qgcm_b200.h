/*
 * qgcm_b200.h -- C ABI of libqgcm_b200.so, the B200-native replacement for the
 * per-timestep hot path of Q-GCM v1.5.0 (reference fork jinkakei/q-gcm).
 *
 * The reference has no FFI: its seam is the set of argument-less Fortran module
 * procedures called from the main loop (src/q-gcm.F:1222-1269) plus the inline
 * time-level averaging block (src/q-gcm.F:1328-1407), all acting on module-global
 * static arrays.  Each entry point below replaces exactly one of those procedures;
 * the Fortran side keeps the same subroutine names and forwards through the
 * ISO_C_BINDING interface module shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is then
 *     available from qgcm_last_error() (reference convention is print + stop,
 *     src/nc_subs.F:84-112, src/ocisubs.F:361-365).
 *   - all host arrays are Fortran column-major, unpadded, double precision, exactly
 *     as declared in the reference data modules (src/ocstate_data.F:39-42,
 *     src/intrfac_data.F:39-48, src/ochomog_data.F:44-69, ...).
 *   - the library owns the device mirrors; host arrays are stale between
 *     qgcm_get_field calls.  Step calls are asynchronous on one CUDA stream per GPU;
 *     qgcm_get_field / qgcm_get_scalars / qgcm_sync synchronise.
 *   - there is no CPU fallback: qgcm_create fails if no sm_100 device is present.
 */
#ifndef QGCM_B200_H
#define QGCM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QGCM_NLMAX 9          /* nlmax in src/eigmode.f:86 */
#define QGCM_ABI_VERSION 1

/* cpp macros of src/make.config:11-45 become run-time flags.  Variants of the reference that no
 * shipped deck uses and this library does NOT implement -- each fails loudly instead of computing
 * something else: -Datmos_only inside xforc (qgcm_xforc returns an error; the other atmosphere
 * procedures work), the fork's -Dsponge_layer_k247 relaxation term of qgostep
 * (src/qgosubs.F:203-205; there is no flag for it, so a deck that needs it cannot be expressed),
 * and y-slab partitions (nranks > 1) of channel or coupled decks (qgcm_create returns an error:
 * they fit one GPU). */
enum {
  QGCM_OCEAN_ONLY   = 1 << 0,   /* -Docean_only   */
  QGCM_ATMOS_ONLY   = 1 << 1,   /* -Datmos_only   */
  QGCM_CYCLIC_OCEAN = 1 << 2,   /* -Dcyclic_ocean */
  QGCM_SB_HFLUX     = 1 << 3,   /* -Dsb_hflux     */
  QGCM_NB_HFLUX     = 1 << 4,   /* -Dnb_hflux     */
  QGCM_TAU_UDIFF    = 1 << 5,   /* -Dtau_udiff    */
  QGCM_OCNC_AVG_K247 = 1 << 6   /* -Docnc_avg_k247: qgcm_run accumulates po after every ocean step */
};

/*
 * Everything the hot path needs that the Fortran main program has computed by
 * src/q-gcm.F:976 (after eigmod, radiat, topset).  Compile-time PARAMETERs of
 * src/parameters_data.F:41-119 and the positional values of src/in_param.f:31-142
 * arrive here as plain run-time fields.  Matrices are Fortran column-major with
 * leading dimension nlo (ocean) / nla (atmosphere), packed at the array start.
 */
typedef struct qgcm_config {
  int32_t abi_version;        /* must be QGCM_ABI_VERSION */
  int32_t struct_bytes;       /* sizeof(qgcm_config) as seen by the caller */
  int32_t flags;              /* QGCM_* bit mask */
  int32_t device;             /* CUDA device ordinal for this process */
  /* src/parameters_data.F:41-88 */
  int32_t nxto, nyto, nlo;    /* ocean T cells and layers; nxpo=nxto+1, nypo=nyto+1 */
  int32_t nxta, nyta, nla;    /* atmosphere T cells and layers */
  int32_t ndxr, nx1, ny1;     /* dxa/dxo, ocean start indices in the atmos grid */
  int32_t nstr;               /* dto/dta */
  /* y-slab partition of the ocean (multi-GPU).  rank r owns p rows
   * [jp0, jp0+nyp_loc); single GPU: nranks=1. */
  int32_t nranks, rank;
  int32_t reserved_i[4];
  /* src/parameters_data.F:96-99 */
  double fnot, beta;
  /* src/in_param.f:31-142 */
  double dxo, dta;
  double delek, cdat, rhoat, rhooc, cpat, cpoc;
  double bccoat, bccooc, xcexp, ycexp;
  double xlamda, hmoc, st2d, st4d;
  double hmat, hmamin, ahmd, at2d, at4d, hmadmp;
  /* outputs of radiat, src/radsubs.f:544-560 and src/radiate_data.F:34-38 */
  double tsbdy, tnbdy, fspco;
  double Bmup, B1down, Cmup, C1down, D0up, Dmup, Dmdown, bface, cface, dface;
  double Aup[QGCM_NLMAX * QGCM_NLMAX], Adown[QGCM_NLMAX * QGCM_NLMAX];
  double Bup[QGCM_NLMAX], Cup[QGCM_NLMAX], Dup[QGCM_NLMAX];
  double rbetat[QGCM_NLMAX], aface[QGCM_NLMAX];
  /* layer vectors, src/occonst_data.F:36-44, src/atconst_data.F:36-44 */
  double hoc[QGCM_NLMAX], gpoc[QGCM_NLMAX], ah2oc[QGCM_NLMAX], ah4oc[QGCM_NLMAX];
  double toc[QGCM_NLMAX];
  double hat[QGCM_NLMAX], gpat[QGCM_NLMAX], ah4at[QGCM_NLMAX], tat[QGCM_NLMAX];
  /* outputs of eigmod, src/eigmode.f:386-428 */
  double amatoc[QGCM_NLMAX * QGCM_NLMAX];
  double ctl2moc[QGCM_NLMAX * QGCM_NLMAX], ctm2loc[QGCM_NLMAX * QGCM_NLMAX];
  double rdm2oc[QGCM_NLMAX];
  double amatat[QGCM_NLMAX * QGCM_NLMAX];
  double ctl2mat[QGCM_NLMAX * QGCM_NLMAX], ctm2lat[QGCM_NLMAX * QGCM_NLMAX];
  double rdm2at[QGCM_NLMAX];
  double reserved_d[8];
} qgcm_config;

/*
 * Host mirror of the scalar state that the hot path mutates:
 * constraint variables (src/ochomog_data.F:57-69, src/athomog_data.F:43-55) and the
 * monitor scalars the step routines write as side effects (src/monitor_data.F:41-61).
 */
typedef struct qgcm_scalars {
  /* ocean constraints */
  double xon[QGCM_NLMAX], dpioc[QGCM_NLMAX], dpiocp[QGCM_NLMAX];
  double ocncs[QGCM_NLMAX], ocncn[QGCM_NLMAX], ocncsp[QGCM_NLMAX], ocncnp[QGCM_NLMAX];
  double enisoc[QGCM_NLMAX], eninoc[QGCM_NLMAX];
  double ajisoc[QGCM_NLMAX], ajinoc[QGCM_NLMAX];
  double ap3soc[QGCM_NLMAX], ap3noc[QGCM_NLMAX], ap5soc[QGCM_NLMAX], ap5noc[QGCM_NLMAX];
  double txisoc, txinoc, bdrins, bdrinn;
  /* ocean homogeneous-solution constants (homsol, src/conhoms.F:386-640) */
  double aipohs[QGCM_NLMAX];
  double cdiffo[QGCM_NLMAX * QGCM_NLMAX], cdhoc[QGCM_NLMAX * QGCM_NLMAX];
  double hc1soc[QGCM_NLMAX], hc2soc[QGCM_NLMAX], hc1noc[QGCM_NLMAX], hc2noc[QGCM_NLMAX];
  double aipcho[QGCM_NLMAX], hbsioc, aipbho;
  /* ocean monitors */
  double cfraoc, centoc;
  double ttmads, vfmads, ttmdfs, ttmadn, vfmadn, ttmdfn;
  double ermaso[QGCM_NLMAX], emfroc[QGCM_NLMAX];
  double xinhom_oc[QGCM_NLMAX];       /* last xinhom(m) of ocinvq (debug/monitor) */
  /* atmosphere constraints */
  double xan[QGCM_NLMAX], dpiat[QGCM_NLMAX], dpiatp[QGCM_NLMAX];
  double atmcs[QGCM_NLMAX], atmcn[QGCM_NLMAX], atmcsp[QGCM_NLMAX], atmcnp[QGCM_NLMAX];
  double enisat[QGCM_NLMAX], eninat[QGCM_NLMAX];
  double ajisat[QGCM_NLMAX], ajinat[QGCM_NLMAX];
  double ap5sat[QGCM_NLMAX], ap5nat[QGCM_NLMAX];
  double txisat, txinat;
  double hc1sat[QGCM_NLMAX], hc2sat[QGCM_NLMAX], hc1nat[QGCM_NLMAX], hc2nat[QGCM_NLMAX];
  double aipcha[QGCM_NLMAX], hbsiat, aipbha;
  /* atmosphere monitors */
  double cfraat, centat;
  double ermasa[QGCM_NLMAX], emfrat[QGCM_NLMAX];
  double xinhom_at[QGCM_NLMAX];
  double arlaav, slhfav, oradav, arocav;
  double reserved_d[16];
} qgcm_scalars;

typedef struct qgcm_model qgcm_model;   /* opaque */

/* ---- lifetime ---------------------------------------------------------------- */

/* Allocates device mirrors of the module storage of src/ocstate_data.F:39-55,
 * src/atstate_data.F:37-40, src/intrfac_data.F:39-48, src/ochomog_data.F:44-69,
 * builds the Helmholtz coefficient tables that src/q-gcm.F:929-973 computes
 * (bd2oc/bd2at, aoc/aat) and the transform plans replacing dsinti/drffti. */
int qgcm_create(const qgcm_config *cfg, qgcm_model **out);
int qgcm_destroy(qgcm_model *m);
const char *qgcm_last_error(void);
int qgcm_abi_version(void);

/* ---- state transfer ----------------------------------------------------------- */

/* name is the reference's Fortran variable name, lower case: "po","pom","qo","qom",
 * "sst","sstm","wekto","wekpo","entoc","tauxo","tauyo","fnetoc","ddynoc","ochom",
 * "pch1oc","pch2oc","pbhoc", and the atmosphere twins "pa","pam","qa","qam","ast",
 * "astm","hmixa","hmixam","wekta","wekpa","entat","tauxa","tauya","fnetat","ddynat",
 * "dtopat","xc1ast","uekat","vekat","sstbar","astbar","pch1at","pch2at","pbhat".
 * n is the element count and must match the Fortran declaration. */
int qgcm_set_field(qgcm_model *m, const char *name, const double *host, int64_t n);
int qgcm_get_field(qgcm_model *m, const char *name, double *host, int64_t n);
int qgcm_field_size(qgcm_model *m, const char *name, int64_t *n);
/* Overlapped upload for forcing that the host supplies while the model runs (the ocean-only
 * reference reads tauxo/tauyo/fnetoc once, src/q-gcm.F:790-808; a host that updates them every
 * step uses this pair).  `host` must be page-locked and must not be modified until the next
 * synchronising call AFTER the commit (qgcm_sync, qgcm_get_field, qgcm_get_scalars) has
 * returned: qgcm_commit_fields only orders the step stream behind the copy, it does not wait
 * for the DMA on the host.  The copy runs on a second stream into a shadow buffer;
 * qgcm_commit_fields makes the step stream wait for it and switches the named fields over
 * (a field uploaded more than once before a commit keeps the last upload). */
int qgcm_set_field_async(qgcm_model *m, const char *name, const double *host, int64_t n);
int qgcm_commit_fields(qgcm_model *m);
/* Restart and output I/O at PCIe speed (resave_nc writes po, pom, sst, sstm and the atmosphere's
 * pa, pam, ast, astm, hmixa, hmixam, src/nc_subs.F:1331-1360; restart_nc reads them back,
 * :1923-1943).  The Fortran state arrays are static module storage that lives as long as the
 * process: qgcm_host_register page-locks such an array once (cudaHostRegister), after which every
 * qgcm_set_field / qgcm_get_field on it is a DMA straight from / into the array.
 * qgcm_get_fields / qgcm_set_fields move n fields with one synchronisation at the end instead of
 * one per field (names[i], hosts[i], counts[i] as in qgcm_get_field); with registered arrays the
 * whole transfer is one stream of back-to-back DMAs.  Unregistered (pageable) arrays work too,
 * at the driver's staging speed. */
int qgcm_host_register(void *host, int64_t bytes);
int qgcm_host_unregister(void *host);
int qgcm_get_fields(qgcm_model *m, int32_t n, const char *const *names, double *const *hosts, const int64_t *counts);
int qgcm_set_fields(qgcm_model *m, int32_t n, const char *const *names, const double *const *hosts, const int64_t *counts);
int qgcm_set_scalars(qgcm_model *m, const qgcm_scalars *s);
int qgcm_get_scalars(qgcm_model *m, qgcm_scalars *s);
int qgcm_sync(qgcm_model *m);

/* ---- initialisation-time procedures that use the device solver ---------------- */

/* constr, src/conhoms.F:44-314: dpioc, dpiocp (and ocncs.. / atmcs.. ) from p. */
int qgcm_constr(qgcm_model *m);
/* homsol, src/conhoms.F:318-818: ochom/aipohs/cdiffo/cdhoc (box) or
 * pch1oc/pch2oc/pbhoc/hc* (cyclic), and the atmosphere twins. */
int qgcm_homsol(qgcm_model *m);
/* q from p for both time levels: qcomp + ocqbdy (+merqcy), src/q-gcm.F:719-732,
 * and qcomp + atqzbd + merqcy for the atmosphere, src/q-gcm.F:738-749. */
int qgcm_qcomp_ocean(qgcm_model *m);
int qgcm_qcomp_atmos(qgcm_model *m);
/* hsbxoc / hscyoc / hscyat, src/ocisubs.F:415-618, src/atisubs.F:301-395, on a
 * host array wrk(nxp,nyp) with coefficient vector b(nxt); which: 0 ocean, 1 atmos. */
int qgcm_helmholtz(qgcm_model *m, int which, double *wrk, const double *b);

/* ---- the per-timestep procedures (src/q-gcm.F:1222-1269) ---------------------- */

int qgcm_xforc(qgcm_model *m);      /* src/xfosubs.F:52-858   */
int qgcm_oml(qgcm_model *m);        /* src/omlsubs.F:47-236   */
int qgcm_qgostep(qgcm_model *m);    /* src/qgosubs.F:45-221   */
int qgcm_ocinvq(qgcm_model *m);     /* src/ocisubs.F:64-407   */
int qgcm_ocqbdy(qgcm_model *m);     /* src/vorsubs.F:245-388, call ocqbdy(qo,po) */
int qgcm_aml(qgcm_model *m);        /* src/amlsubs.F:47-238   */
int qgcm_qgastep(qgcm_model *m);    /* src/qgasubs.F:45-148   */
int qgcm_atinvq(qgcm_model *m);     /* src/atisubs.F:60-293   */
int qgcm_atqzbd(qgcm_model *m);     /* src/vorsubs.F:396-480, call atqzbd(qa,pa) */
int qgcm_tlavg_ocean(qgcm_model *m);  /* src/q-gcm.F:1328-1366 */
int qgcm_tlavg_atmos(qgcm_model *m);  /* src/q-gcm.F:1370-1407 */

/* oml + qgostep + ocinvq + ocqbdy, the body of src/q-gcm.F:1229-1249 */
int qgcm_ocean_step(qgcm_model *m);
/* aml + qgastep + atinvq + atqzbd, src/q-gcm.F:1257-1269 */
int qgcm_atmos_step(qgcm_model *m);
/* The loop body of src/q-gcm.F:1220-1408 for nt = nt_first..nt_last inclusive:
 * ocean step when mod(nt,nstr)==1 (every step when nstr==1),
 * atmosphere step unless ocean_only, time-level averaging on its cadence.
 * Coupled models: a whole cycle (xforc, the ocean step, nstr atmosphere steps) inside the range runs as
 * one CUDA graph in which the atmosphere steps are a branch beside the ocean step (after xforc the two
 * touch disjoint state); the results are bit-identical with the call-by-call order, and every other entry
 * point sees the state only after both branches have joined. */
int qgcm_run(qgcm_model *m, int64_t nt_first, int64_t nt_last);

/* ---- y-slab multi-GPU (new: the reference is single-node OpenMP over j, src/qgosubs.F:173-184)
 *
 * The ocean-only box decks (NAtl) partition into contiguous slabs of p rows, one per GPU, the
 * same way the reference's OpenMP loops partition j.  cfg.nranks / cfg.rank select the slab;
 * qgcm_set_field / qgcm_get_field still take the reference's GLOBAL host arrays: a slab
 * uploads the rows it holds and downloads the rows it owns (the other rows of the host
 * array are left untouched, so calling qgcm_get_field for every rank on one buffer
 * assembles the global field).  With a communicator in place qgcm_constr, qgcm_homsol,
 * qgcm_qcomp_ocean, qgcm_ocean_step, qgcm_tlavg_ocean and qgcm_run act on the partition;
 * the per-procedure calls qgcm_oml / qgcm_ocinvq are single-GPU only. */

/* p rows [*jp0, *jp0 + *nyp_own) of nyp_global owned by `rank`; pure host arithmetic */
int qgcm_slab_bounds(int32_t nyp_global, int32_t nranks, int32_t rank, int32_t *jp0, int32_t *nyp_own);
/* one process per GPU: rank 0 fills a 128-byte NCCL id, the host program broadcasts it
 * (MPI_Bcast in the Fortran driver, torch.distributed in bench.py), every rank joins */
int qgcm_nccl_unique_id(void *id128);
int qgcm_comm_init_nccl(qgcm_model *m, const void *id128);
/* Peer-memory transport (one process per GPU, all on one NVLink/NVSwitch node, <= 8 ranks):
 * every rank exports the 64-byte CUDA IPC handle of its mailbox, the host program gathers the
 * handles in rank order (MPI_Allgather / torch.distributed), every rank maps them.  Afterwards
 * the exchanges of qgcm_ocean_step are stores into the peers' mailboxes plus epoch flags,
 * issued by the kernels of the step themselves (no collective call on the step stream); the
 * initialisation procedures use the same mailboxes, so NCCL is optional.  All ranks must call
 * the partition procedures in the same order and enter each within the peer time-out of the
 * others (120 s unless QGCM_PEER_TIMEOUT_S or qgcm_comm_peer_timeout says otherwise; a host
 * that stalls one rank for longer -- restart or netCDF output on one rank -- puts an
 * MPI_Barrier in front of the next step): a rank that waits longer gives up, its slab is then
 * invalid, and every later call on that model, qgcm_get_field and qgcm_get_scalars included,
 * fails with that error instead of handing back spoilt state (the flag lives in host-mapped
 * memory, so the check costs no synchronisation and is made on every step).
 * qgcm_comm_transport switches a model that has both between NCCL (0) and peer memory (1);
 * every rank must switch at the same point of the call sequence. */
int qgcm_peer_handle(qgcm_model *m, void *handle64);
int qgcm_comm_init_peer(qgcm_model *m, const void *handles, int32_t n);
int qgcm_comm_transport(qgcm_model *m, int32_t kind);
/* give-up time of a mailbox wait in seconds (after qgcm_peer_handle; every rank the same) */
int qgcm_comm_peer_timeout(qgcm_model *m, double seconds);
/* Tear-down order (CUDA IPC rule: an exported buffer must not be freed while another process
 * still maps it): every rank calls qgcm_comm_close_peer (unmaps the others' mailboxes), the
 * host program synchronises the ranks (MPI_Barrier), then every rank calls qgcm_destroy. */
int qgcm_comm_close_peer(qgcm_model *m);
/* all ranks in one process on one device (tests on a single GPU): models[r] must be rank r
 * of an n-rank partition; afterwards a partition call on any member steps every rank */
int qgcm_group_create(qgcm_model **models, int32_t n);

/* ---- device-side validity scan ---------------------------------
 *
 * valids, src/valsubs.F:43-630: extreme-value scan of po, qo, sst, wekto (and pa, qa, ast,
 * wekta, tauxa, tauya) against the reference's fixed thresholds (:77-81) and the perturbed
 * layer-thickness check of the ocean (:380-524, thkmin = 100 m, critpc = 20 %), evaluated on
 * the device so that the every-0.25-day check (src/q-gcm.F:1278) does not download 2 GB.
 * solnok = 0 means the reference would dump and stop.  The neighbourhood print-outs of
 * scan2D/scan3D stay on the Fortran side (they run on downloaded fields after a failure).
 * On a y-slab model the report covers the rows this rank owns; min/max and the hfbad
 * percentages of the ranks combine by min/max/sum. */
typedef struct qgcm_valids_report {
  int32_t solnok, reserved;
  double patmin, patmax, qatmin, qatmax, astmin, astmax, wtamin, wtamax, txamin, txamax, tyamin, tyamax;
  double pocmin, pocmax, qocmin, qocmax, sstmin, sstmax, wtomin, wtomax;
  double hfmint, hfmaxt, hfmini, hfmaxi, hfminb, hfmaxb;   /* full layer thickness: top, intermediate, bottom */
  double hfbad[QGCM_NLMAX];                                /* % of the area thinner than thkmin, per layer */
} qgcm_valids_report;
int qgcm_valids(qgcm_model *m, qgcm_valids_report *rep);

/* ---- device-side running sums and packed output ------------
 *
 * tavini / tavatm / tavocn, src/timavge.F:108-617, and the fork's avg_ocn_k247 (:624-660,
 * called after every ocean step, src/q-gcm.F:1250-1252): the sums stay in HBM under the
 * reference's array names -- "txocav","tyocav","wpocav","wtocav","fmocav","sstav","uufo",
 * "tufo","utufo","vvfo","tvfo","vtvfo","pocav","qocav","po_avg" and the atmosphere twins
 * "txatav","tyatav","wtatav","fmatav","astav","uufa","tufa","utufa","vvfa","tvfa","vtvfa",
 * "patav","qatav" (shapes as declared at src/timavge.F:46-85) -- and are read with
 * qgcm_get_field when tavout (src/timavge.F:667) needs them, so that the daily accumulation
 * (src/q-gcm.F:1477-1482) does not download the state.  The sums are allocated on first use;
 * qgcm_tavini zeroes them and the contribution counts nsumat, nsumoc, nsum_ocavg. */
int qgcm_tavini(qgcm_model *m);
int qgcm_tavatm(qgcm_model *m);
int qgcm_tavocn(qgcm_model *m);
int qgcm_avg_ocn_k247(qgcm_model *m);
int qgcm_tav_counts(qgcm_model *m, int32_t *nsumat, int32_t *nsumoc, int32_t *nsum_ocavg);
/* The sub-sampled vector that ocnc_out / atnc_out hand to nf_put_vara_double
 * (src/nc_subs.F:869-890, :906-917, :1110-1130): host(i,j,k) = field(1+(i-1)*nsk, 1+(j-1)*nsk, k)
 * with iw = min(mod(nx,nsk),1) + (nx-mod(nx,nsk))/nsk points per direction, packed on the
 * device so that an output step moves 1/nsk^2 of the field over PCIe.  Any gridded field
 * name of qgcm_get_field; n must equal the count qgcm_field_sub_size returns. */
int qgcm_field_sub_size(qgcm_model *m, const char *name, int32_t nsk, int64_t *n);
int qgcm_get_field_sub(qgcm_model *m, const char *name, int32_t nsk, double *host, int64_t n);

/* qocdiag_out, src/qocdiag.F:303-683 (-Dqoc_diag, e.g. examples/double_gyre_ocean_only; called
 * between oml and qgostep at output steps, src/q-gcm.F:1237-1239): the ocean vorticity tendency
 * and its component terms, evaluated on the device at the sub-sampled points only.
 * host(ipwk, jpwk, nlo, 5), the last index in the reference's output order dqdt, qotjac, qt2dif,
 * qt4dif, qotent; each (ipwk, jpwk) plane is the wrk vector of one nf_put_vara_double call
 * (:616-672).  n must equal the count qgcm_qocdiag_size returns (5*ipwk*jpwk*nlo). */
int qgcm_qocdiag_size(qgcm_model *m, int32_t nsko, int64_t *n);
int qgcm_qocdiag(qgcm_model *m, int32_t nsko, double *host, int64_t n);

/* monnc_comp, ocean section with couroc, src/monitor_diag.F:480-840, :1450-1925 (called every dgnday, src/q-gcm.F:1442):
 * Ekman-velocity and entrainment means, interface displacement moments, wind work, per-layer
 * kinetic energy, its tendency and the del-sqd/del-4th dissipation integrals (with the one-sided
 * boundary Laplacians of del4bx/del4ch, :900-1120, and the weighted area integral genint,
 * :1160-1210), jet position, stream-function extrema, layer transports, bottom drag and the
 * mixed-layer heat diagnostics -- the values monnc_comp stores in the module `monitor`
 * (src/monitor_data.F:41-61) for monnc_out, computed on the device from the resident state.
 * ocjpos is the reference's 1-based T-row index.  The call includes couroc (:1450-1925), which
 * monnc_comp invokes at :826.  On a y-slab partition this is a partition call (like qgcm_constr): every rank sums the
 * rows it owns, the shares are added across the ranks and every rank returns the same report. */
typedef struct qgcm_monitor_ocean {
  double wetmoc, watmoc, wepmoc, wapmoc, entmoc, enamoc;
  double etamoc[QGCM_NLMAX], et2moc[QGCM_NLMAX], ddtpeoc[QGCM_NLMAX], pkenoc, utauoc;
  double ocjval[QGCM_NLMAX];
  int32_t ocjpos[QGCM_NLMAX], reserved_i;
  double pavgoc[QGCM_NLMAX], qavgoc[QGCM_NLMAX], ah2doc[QGCM_NLMAX], ah4doc[QGCM_NLMAX];
  double kealoc[QGCM_NLMAX], ddtkeoc[QGCM_NLMAX], osfmin[QGCM_NLMAX], osfmax[QGCM_NLMAX], occirc[QGCM_NLMAX];
  double btdgoc, sstmin, sstmax, hfmloc, tmlmoc, occtot;
  /* couroc, src/monitor_diag.F:1450-1925: velocity extrema at the cell faces and the largest
   * Courant number, mixed layer (geostrophic + Ekman) and every QG layer */
  double umminoc, ummaxoc, vmminoc, vmmaxoc, cnmloc;
  double ugminoc[QGCM_NLMAX], ugmaxoc[QGCM_NLMAX], vgminoc[QGCM_NLMAX], vgmaxoc[QGCM_NLMAX], cnqgoc[QGCM_NLMAX];
} qgcm_monitor_ocean;
int qgcm_monnc_ocean(qgcm_model *m, qgcm_monitor_ocean *rep);
/* monnc_comp, atmosphere section with courat, src/monitor_diag.F:186-478, :1215-1445.  The
 * reference's own slips are reproduced: vkedot integrates the stale workspace attwk3, which
 * holds del-sqd of the lagged v (:391, :404), and the atmosphere has no del-sqd dissipation
 * term.  davgat is the mean of dtopat (src/topsubs.F:429-430), formed from the device field.
 * atstpos is the 1-based T-row index. */
typedef struct qgcm_monitor_atmos {
  double wetmat, watmat, wepmat, wapmat;
  double entmat[QGCM_NLMAX], enamat[QGCM_NLMAX], etamat[QGCM_NLMAX], et2mat[QGCM_NLMAX], ddtpeat[QGCM_NLMAX], pkenat[QGCM_NLMAX];
  double utauat;
  double atstval[QGCM_NLMAX];
  int32_t atstpos[QGCM_NLMAX], reserved_i;
  double pavgat[QGCM_NLMAX], qavgat[QGCM_NLMAX], ah4dat[QGCM_NLMAX], kealat[QGCM_NLMAX], ddtkeat[QGCM_NLMAX];
  double tmlmat, hmlmat, hcmlat, astmin, astmax, tmaooc, olrtop, davgat;
  double umminat, ummaxat, vmminat, vmmaxat, cnmlat;
  double ugminat[QGCM_NLMAX], ugmaxat[QGCM_NLMAX], vgminat[QGCM_NLMAX], vgmaxat[QGCM_NLMAX], cnqgat[QGCM_NLMAX];
} qgcm_monitor_atmos;
int qgcm_monnc_atmos(qgcm_model *m, qgcm_monitor_atmos *rep);

/* ---- instrumentation ---------------------------------------------------------- */

/* number of kernels launched by this model since creation */
int64_t qgcm_launch_count(qgcm_model *m);
/* per-kernel timing with CUDA events on the launching stream: enable (1) / disable (0)
 * clears the records; the report is one text line per kernel: "name launches total_ms" */
int qgcm_profile(qgcm_model *m, int enable);
int qgcm_profile_report(qgcm_model *m, char *buf, int64_t nbuf);
/* cudaStream_t the model launches on, as an opaque pointer (for event timing) */
void *qgcm_stream(qgcm_model *m);

#ifdef __cplusplus
}
#endif
#endif /* QGCM_B200_H */
