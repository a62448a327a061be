// jc_internal.cuh -- shared definitions of the sm_100a angular-C_ell pipeline (not part of the ABI).
//
// Pipeline per chunk of cosmologies (all FP64, one stream, no host sync):
//   K1 jc_setup_kernel    per-cosmology tables: chi(a) [256], D(a) [128], EH scalars, sigma8 norm,
//                         halofit S(R) [256], k_nl/n_eff/C and Takahashi coefficients at the 513
//                         Limber nodes                                  (one CTA per cosmology)
//   K2 jc_tracer_kernel   radial kernels R_i(a_n): lensing efficiency on the [257 x 512] z' grid
//                         against cosmology-independent n(z') tables, NC / NLA / m-bias
//   K3 jc_power_kernel    V[n,l] = w_n P(k=(l+1/2)/chi_n, a_n) dchi/da / chi_n^2 / c^2
//                         (Eisenstein-Hu + halofit, fully in registers)
//   K4 jc_contract_kernel C[(i,j),l] = e_i(l) e_j(l) sum_n R_i[n] R_j[n] V[n,l]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/jc_b200.h"

#define JC_NA 513        // Limber nodes (angular_cl.py:96)
#define JC_NA_PAD 520    // padded per-node stride (64-byte multiple)
#define JC_NCHI 256      // chi table (background.py:199)
#define JC_NGROW 128     // growth table (background.py:443)
#define JC_NLENS 257     // z' nodes of the lensing-efficiency Simpson rule (probes.py:51)
#define JC_NLENS_COLS 512  // Limber nodes with chi>0 (node 512 is a=1: kernel == 0)
#define JC_NHFK 257      // halofit ln k nodes (power.py:111)
#define JC_NHFR 256      // halofit ln R nodes (power.py:93)
#define JC_NROMB 129     // Romberg nodes, divmax=7 (power.py:77)
#define JC_MAX_CHUNK 4096
#define JC_JVP_MAX_GROUP 4  // tangent directions one JVP pass can carry (DualN<4>: the setup kernel's tables fill an SM's shared memory)

#define JC_C_LIGHT 299792.458      // constants.py:9
#define JC_RH 2997.92458           // constants.py:15
#define JC_H0 100.0                // constants.py:21
#define JC_TCMB 2.726              // constants.py:12
#define JC_C1_RHOCRIT (5.0 * 1e-14 * (2.7750 * 1e11))  // constants.py:24,27 ; probes.py:121
#define JC_STERADIAN_TO_ARCMIN2 11818102.86004228      // redshift.py:10
#define JC_TWO_PI_SQ 19.739208802178716                // 2 pi^2

// Device-side view of the plan: cosmology-independent tables (all device pointers).
struct JcDevPlan {
  int T, P, L, Lpad;
  int nonlinear;             // JC_PK_*: 0 linear, 1 halofit takahashi2012, 2 halofit smith2003
  int transfer;              // JC_TF_*: 1 Eisenstein-Hu with wiggles, 2 no-wiggle fit
  int TS;                    // tracer stride of the node-major R table (>= T, TS mod 16 in {4,12})
  int n_src;                 // number of weak-lensing tracers
  double zmax;               // Limber zmax (max over probes), angular_cl.py:63
  double lens_zmax;          // zmax of the WL probe(s), probes.py:24
  // chi table quadrature: points p = 2i (node i), 2i+1 (midpoint of interval i)
  const double* chi_pt_a;    // [511]
  const double* chi_pt_lna;  // [511]
  const double* chi_h6;      // [255]  h_i / 6
  int growth, ncp;           // JC_GROWTH_*; doubles per cosmology / tangent row (8 or 9)
  int reserved_sms;          // SMs occupied by a concurrent kernel (gather pusher): persistent kernels size their grid without them
  int grid_mode, grid_na;    // grid plan (jc_grid_plan_create): nodes = caller's scale factors, "ell + 1/2" = caller's k
  // growth table quadrature: points p = 2i (node), 2i+1 (a_i + h_i/2; JC_GROWTH_GAMMA: ln a_i + h_i/2 with h in ln a)
  const double* gr_pt_a;     // [255]
  const double* gr_pt_lna;   // [255]
  const double* gr_h;        // [127]
  // Limber nodes
  const double* limb_a;      // [513]
  const double* limb_lna;    // [513]
  const double* limb_z;      // [513]
  const double* limb_w;      // [513] Simpson weights * dx/3
  const double* limb_chi_t;  // [513] interpolation weight in the chi table
  const double* limb_gr_t;   // [513] interpolation weight in the growth table
  const uint16_t* limb_chi_ix;  // [513] i0 | i1<<8
  const uint16_t* limb_gr_ix;   // [513]
  // sigma8 Romberg functional
  const double* romb_k;      // [129]
  const double* romb_lnk;    // [129]
  const double* romb_f;      // [129] w_i (b-a) k (k W(8k))^2 / (2 pi^2)
  const double* romb_w;      // [129] w_i (b-a)
  // halofit grids
  const double* hf_k;        // [257]
  const double* hf_lnk;      // [257]
  const double* hf_wk;       // [257] Simpson weights * dx/3
  const double* hf_r;        // [256]
  const double* hf_logr;     // [256]
  // lensing-efficiency grid, [257][512] (row m = z' node, column n = Limber node)
  const double* lens_t;      // interpolation weight of a'=1/(1+z') in the chi table
  const uint16_t* lens_ix;   // i0 | i1<<8
  const double* lens_nw;     // [n_src][257][512]  simpson_w[m]/(3*256) * n_s(z'(m,n))
  // tracers
  const double* nz_node;     // [520][TS] normalised n_i(z_n), node-major
  const double* bias_node;   // [520][TS] cosmology-independent part of b_i(z_n) (NC) / b_IA (WL)
  const int* tr_kind;        // [T]
  const int* tr_inv_growth;  // [T] bias multiplies 1/D(a)
  const int* tr_ia;          // [T]
  const int* tr_src;         // [T] index into lens_nw or -1
  const int* src_tracer;     // [n_src] tracer index of each lensing source
  const int* fin_idx;        // [n_fin] tracers the finish kernel evaluates at every node (NC, delta planes, IA sources)
  int n_fin;
  const int* tr_delta_ix;    // [T] delta_nz source plane: chi-table bracket i0 | i1<<8 of a_s, else -1
  const double* tr_delta_t;  // [T] its interpolation weight
  const double* tr_m1;       // [T] 1 + m
  // ell
  double lnl_min, lnl_max;   // min / max of ln(ell + 1/2): the ln k range of the tabulated transfer function
  double lnl_step;           // spacing of ln(ell + 1/2) when it is uniform (np.logspace in ell + 1/2 ... ), else 0
  const double* ell;         // [L]
  const double* ellp5;       // [L] ell + 0.5
  const double* lnellp5;     // [L]
  const double* ellfac;      // [L] WL ell factor (probes.py:73)
  const double* ell108;      // [L] (l+1/2)^1.08
  const double* ell14;       // [L] (l+1/2)^1.4
  const double* ellm3;       // [L] (l+1/2)^-3
  const double* covnorm;     // [L] (2l+1) gradient(l)   (angular_cl.py:139, without f_sky)
  const double* math_tab;    // [JCM_TAB_DOUBLES] exp2 / log tables of jc_math.cuh (JCM_TAB_*)
  // pairs
  const uint8_t* pair_i;     // [P]
  const uint8_t* pair_j;     // [P]
  // contraction order of the persistent TMA kernel: pairs sorted by the first 12-node stage their product can be
  // non-zero at (number-count kernels vanish where n(z) does), in tiles of 8
  const uint8_t* cpair_i;    // [Ppad8] tracer i of sorted pair q
  const uint8_t* cpair_j;    // [Ppad8]
  const uint16_t* cpair_out; // [Ppad8] output row (index into the reference's pair order)
  const uint8_t* ctile_lo;   // [Ppad8 / 8] first stage of the tile's range
  const uint8_t* ctile_hi;   // [Ppad8 / 8] last stage (lo > hi: nothing to do, the rows are zero)
};

#define JC_PROF_SLOTS 2048
struct JcProf {
  int enabled;
  int used;                                   // chunk passes recorded
  cudaEvent_t ev[JC_PROF_SLOTS][JC_N_STAGES + 1];
  int launches[JC_PROF_SLOTS][JC_N_STAGES];
};

struct jc_plan {
  int device;
  JcProf* prof;
  jc_problem problem;
  JcDevPlan d;
  void* dev_blob;       // single allocation holding every table
  size_t dev_blob_bytes;
  double noise[JC_MAX_TRACERS];
  // arena for the host entry point
  void* arena_ws;
  size_t arena_ws_bytes;
  double* arena_cosmo;
  size_t arena_cosmo_bytes;
  double* arena_cl[2];
  size_t arena_cl_bytes;
  cudaStream_t s_compute, s_copy;
  cudaEvent_t ev_done[2], ev_copied[2];
};

// Function attributes (dynamic shared-memory opt-in) belong to the device's context: run the statement(s) the first time the call site is reached on
// the CURRENT device (`flag`: one bit per device ordinal, racing callers repeat an idempotent call).
#define JC_ONCE_PER_DEVICE(flag, ...)                                           \
  do {                                                                          \
    int jc_d_ = 0;                                                              \
    cudaGetDevice(&jc_d_);                                                      \
    const unsigned long long jc_b_ = 1ull << (jc_d_ & 63);                      \
    if (!(__atomic_load_n(&(flag), __ATOMIC_ACQUIRE) & jc_b_)) {                \
      __VA_ARGS__;                                                              \
      __atomic_fetch_or(&(flag), jc_b_, __ATOMIC_RELEASE);                      \
    }                                                                           \
  } while (0)

// Entry points that own a device (plan creation / destruction, the host-buffer calls) switch to the plan's device and put the caller's
// current device back on every return path.
struct JcDeviceGuard {
  int prev = -1;
  cudaError_t status;
  explicit JcDeviceGuard(int device) {
    cudaGetDevice(&prev);
    status = cudaSetDevice(device);
  }
  ~JcDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  JcDeviceGuard(const JcDeviceGuard&) = delete;
  JcDeviceGuard& operator=(const JcDeviceGuard&) = delete;
};

void jc_set_cuda_error(cudaError_t e, const char* where);
extern int g_jc_power_exact;      // jc_set_option("power_exact"): exact-formula power kernel everywhere
extern int g_contract_cfg;        // jc_set_option("contract_kernel")
extern int g_jc_lens_mma;         // jc_set_option("lens_mma"): K2a launches of >= 8 sources on the DMMA lens kernel
extern int g_jc_jvp_adjoint;      // jc_set_option("jvp_adjoint"): 1 = reverse sweep of the point function in K3 for >= 3 directions (default)
extern int g_jc_jvp_group;        // jc_set_option("jvp_group"): tangent directions carried per JVP pass (1..JC_JVP_MAX_GROUP)
extern double g_jc_contract_eps;  // jc_set_option("contract_eps"): support threshold of the contraction, read at plan creation
typedef int (*jc_slice_cb)(void* ctx, int64_t first_row, int64_t rows);
int jc_run_pipeline(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, double* cl_dev, void* ws_dev,
                    size_t ws_bytes, cudaStream_t s, int64_t chunk_cap, int64_t slice, jc_slice_cb cb, void* ctx,
                    int reserved_sms);
void jc_math_table(double* out288);  // host: tables of the table-driven exp / log (jc_math.cuh)
int jc_pipeline_init();  // one-time function attributes (dynamic shared memory opt-in)

struct Ws {  // resolved workspace pointers for one chunk of cosmologies
  double* chitab;  // [chunk][256]
  double* gtab;    // [chunk][128]
  double* scal;    // [chunk][32]
  double* stab;    // [chunk][256]
  double* node;    // [chunk][JC_NODE_FIELDS][JC_NA_PAD]
  double* rker;    // [chunk][JC_NA_PAD][TS]   node-major radial kernels R_i(a_n)
  double* vtab;    // [chunk][513][Lpad]
  double* ellpow;  // [chunk][Lpad]  (l+1/2)^(3+n_s)
  ptrdiff_t doff;  // JVP passes: offset (doubles) from a value to its first tangent plane (plane k at (k + 1) * doff); 0 otherwise
};

#ifdef __CUDACC__
__device__ __forceinline__ double* node_ptr(const Ws& ws, int c, int field) {
  return ws.node + ((size_t)c * JC_NODE_FIELDS + field) * JC_NA_PAD;
}
// 1/x for finite positive normal x: MUFU.RCP64H seed + 2 Newton steps (<= ~1.5 ulp), no special cases.
__device__ __forceinline__ double jc_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
#endif

// per-stage launchers (one translation unit per kernel)
void jc_launch_setup(const JcDevPlan& pl, const double* cosmo, const Ws& ws, int chunk, cudaStream_t s);
void jc_launch_contract_1cta(const JcDevPlan& pl, const Ws& ws, double* cl, int chunk, cudaStream_t s);
void jc_launch_sigmasqr(const JcDevPlan& pl, const Ws& ws, const double* cosmo, int chunk, const double* R_dev, int n_R, double* out,
                        cudaStream_t s);
void jc_launch_transfer(const JcDevPlan& pl, const Ws& ws, int chunk, double* tk, cudaStream_t s);
// JVP variants: the same kernels instantiated on DualN<ntan> (value + ntan tangents per workspace entry, 1 <= ntan <=
// JC_JVP_MAX_GROUP; tangent plane k of a table at offset (k + 1) * ws.doff); `tangent` = ntan directions [ntan][ncp] in
// parameter space per entry
void jc_launch_setup_jvp(const JcDevPlan& pl, const double* cosmo, const double* tangent, const Ws& ws, int chunk, int kdiv,
                         int ntan, cudaStream_t s);
// skip_static: the pass's value tables are not read by anyone (later tangent groups of the reverse-sweep path) -- the kernels
// return at once when none of the pass's directions can move the tracer kernels (JC_SCAL_MOVES_R)
int jc_launch_tracers_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s, int skip_static = 0);
void jc_launch_finish_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s, int skip_static = 0);
void jc_launch_power_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s);
void jc_launch_contract_jvp(const JcDevPlan& pl, const Ws& ws, double* dcl, int64_t dcl_cosmo_stride, int chunk,
                            cudaStream_t s);
// K3 for all directions of a Jacobian at once: value + ntan directional derivatives from one reverse sweep of the point function
// (jc_power_adj.cu).  Direction k's tangent plane of every table sits at jc_jvp_plane(k) * ws.doff: the planes written by the
// grouped K1 / K2 passes when group j = k / JC_JVP_MAX_GROUP runs on the workspace shifted by j * (JC_JVP_MAX_GROUP + 1) planes.
#define JC_JVP_ADJ_MAX 8
#define JC_JVP_FUSED_MAX 512  // B*K up to which a small batch runs as B*K one-direction entries in one (latency-bound) pass
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int jc_jvp_plane(int k) { return (k / JC_JVP_MAX_GROUP) * (JC_JVP_MAX_GROUP + 1) + (k % JC_JVP_MAX_GROUP) + 1; }
inline int jc_jvp_planes(int ntan) { return jc_jvp_plane(ntan - 1) + 1; }  // planes a workspace for ntan directions holds
bool jc_power_adj_supported(const JcDevPlan& pl, int ntan);
void jc_launch_power_adj(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s);
int jc_setup_init();
int jc_launch_tracers(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s);  // lensing; returns #launches
void jc_launch_finish(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s);
void jc_launch_power(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s);
void jc_launch_contract(const JcDevPlan& pl, const Ws& ws, double* cl, int chunk, cudaStream_t s);
int jc_contract_init();

#define JC_CUDA_TRY(expr)                          \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) {                       \
      jc_set_cuda_error(_e, #expr);                \
      return JC_ERR_CUDA;                          \
    }                                              \
  } while (0)
