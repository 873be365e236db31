// Internal declarations of libqgcm_b200.so (sm_100a).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
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
};

// Device-side constants small enough to pass by value to kernels
struct LayerConsts {
  double h[NLMAX], gp[NLMAX], ah2[NLMAX], ah4[NLMAX];
  double amat[NLMAX * NLMAX];   // ld = nl
  double ctl2m[NLMAX * NLMAX], ctm2l[NLMAX * NLMAX], rdm2[NLMAX];
};

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
  int nrows;         // interior rows nyp-2
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
  double *rowsum = nullptr;  // [nmodes][nyp]  xintp row sums of the solution
  double *ayrow = nullptr;   // [nmodes][2]    periodic: line sums of rows 2 and nyp-1
};

// Scratch and tables of the coupled forcing xforc (atmos.cu); built on first use
struct XfPlan {
  bool ready = false;
  int nxf = 0, nyf = 0, ldf = 0;         // ocean-resolution atmosphere p grid (nxpaor x nypaor) and its pitch
  double *u1 = nullptr, *v1 = nullptr;   // layer-1 geostrophic velocity at atmosphere p points [nypa][ld]
  double *taux = nullptr, *tauy = nullptr;   // stress on the fine grid [nyf][ldf]
  double *stb = nullptr;                 // bicubic weights [5][(ndxr+1)^2][16]: general, u-south, v-south, u-north, v-north
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
    size_t lsz = 0;                       // device layer stride (gridded fields)
    size_t elems = 0;                     // device elements allocated
  };
  std::map<std::string, Field> fields;
  // leapfrog pointer rotation: logical name -> current buffer (the map entries are
  // swapped, host never sees the rotation)
  qgcm_scalars *d_scal = nullptr;        // device-resident scalar state
  double *d_coef = nullptr;              // small device scratch for step coefficients
  double *d_red = nullptr;               // reduction scratch
  size_t red_elems = 0;
  qg::HelmPlan hpo, hpa;
  qg::XfPlan xf;
  double *wrk_o = nullptr, *wrk_a = nullptr;   // modal work arrays [nl][nyp][ld]
  double *xfo = nullptr, *xfa = nullptr, *sstnew = nullptr, *astnew = nullptr, *hmnew = nullptr;
  std::vector<void *> allocs;
  // per-kernel event timing (qgcm_profile)
  bool prof = false;
  struct ProfRec { const char *name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;

  double *F(const char *name) { return fields.at(name).d; }
  void swapf(const char *a, const char *b) { std::swap(fields.at(a).d, fields.at(b).d); }
};

#ifdef __CUDACC__
namespace qg {
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

}  // namespace qg
