/*
 * jc_b200.h -- C ABI of the B200-native angular-power-spectrum path.
 *
 * This is the drop-in boundary for jax_cosmo's `angular_cl` / `gaussian_cl_covariance_and_mean`
 * hot path.  The reference has NO native interface for this path (pure Python on JAX; SURVEY.md
 * section 8b) -- its boundary is the Python call
 *     jax_cosmo/angular_cl.py:49   angular_cl(cosmo, ell, probes, transfer_fn, nonlinear_fn)
 *     jax_cosmo/angular_cl.py:101  noise_cl(ell, probes)
 *     jax_cosmo/angular_cl.py:120  gaussian_cl_covariance(ell, probes, cl_signal, cl_noise, f_sky, sparse)
 *     jax_cosmo/angular_cl.py:166  gaussian_cl_covariance_and_mean(...)
 * so the entry points below are what an XLA-FFI custom call (`jax.ffi`) or a ctypes stub for those
 * four functions binds (INTEGRATION.md shows both bindings).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types, no exceptions across the boundary;
 *   - every function returns JC_OK (0) or a negative jc_status; nothing is launched on error;
 *   - `*_dev` pointers are DEVICE pointers owned by the caller; `*_host` are host pointers;
 *   - device entry points are stream-ordered and asynchronous (`stream` is a cudaStream_t passed
 *     as void*), never allocate or free device memory, and keep no mutable global state, so they
 *     are re-entrant and CUDA-graph capturable;
 *   - all floating point is IEEE binary64 ("jax_enable_x64" semantics of the reference).
 *
 * Layouts (row-major, last index fastest)
 *   cosmo  [B, 8]      Omega_c, Omega_b, h, n_s, sigma8, Omega_k, w0, wa
 *                      (= Cosmology.tree_flatten order, jax_cosmo/core.py:99-108); [B, 9] with the
 *                      growth index gamma appended when jc_problem.growth == JC_GROWTH_GAMMA
 *                      (core.py:104-105); tangent rows have the same width
 *   ell    [L]
 *   cl     [B, P, L]   P = T(T+1)/2 tracer pairs (i<=j), row-major upper triangle
 *                      (jax_cosmo/angular_cl.py:15-25); one [P, L] slab == the reference's output
 *   cov    [P, P, L]   jax_cosmo.sparse block layout (jax_cosmo/angular_cl.py:156-157)
 */
#ifndef JC_B200_H
#define JC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JC_ABI_VERSION 2
#define JC_MAX_TRACERS 32
#define JC_MAX_SHIFTS 4
#define JC_N_COSMO_PARAMS 8     /* row width with JC_GROWTH_ODE; +1 (gamma) with JC_GROWTH_GAMMA */
#define JC_MAX_COSMO_PARAMS 9
#define JC_N_LIMBER_NODES 513 /* simps(..., 512) in a, jax_cosmo/angular_cl.py:96 */

typedef enum jc_status {
  JC_OK = 0,
  JC_ERR_INVALID = -1,     /* malformed argument (NULL, n<=0, bad enum, ...)                  */
  JC_ERR_UNSUPPORTED = -2, /* valid in the reference but outside this path (raise, no fallback) */
  JC_ERR_WORKSPACE = -3,   /* workspace smaller than jc_workspace_bytes(plan, 1)              */
  JC_ERR_CUDA = -4,        /* a CUDA runtime call failed; see jc_last_cuda_error()            */
  JC_ERR_NO_DEVICE = -5    /* no CUDA device / not an sm_100 device                           */
} jc_status;

/* n(z) families: jax_cosmo/redshift.py */
enum {
  JC_NZ_SMAIL = 1, /* redshift.py:61-77   z^a exp(-(z/z0)^b), params = {a, b, z0}                        */
  JC_NZ_FU = 2,    /* redshift.py:80-105  (z^a + z^(ab)) / (z^b + c), params = {a, b, c}                  */
  JC_NZ_DELTA = 3, /* redshift.py:108-123 source plane at params[0]; weak lensing without IA only
                      (probes.py:53-64; the reference raises for density / NLA kernels, probes.py:82,107) */
  JC_NZ_KDE = 4    /* redshift.py:126-156 Gaussian KDE of a catalogue: kde_z / kde_w / kde_n / kde_bw     */
};
/* bias families: jax_cosmo/bias.py */
enum {
  JC_BIAS_NONE = 0,
  JC_BIAS_CONSTANT = 1,       /* bias.py:10-22  params = {b}          */
  JC_BIAS_INVERSE_GROWTH = 2, /* bias.py:25-39  params = {b}  b/D(a)  */
  JC_BIAS_DES_Y1_IA = 3       /* bias.py:42-57  params = {A, eta, z0} */
};
enum {
  JC_TRACER_WEAK_LENSING = 1, /* probes.py:132-223 */
  JC_TRACER_NUMBER_COUNTS = 2 /* probes.py:226-281 */
};
enum {
  JC_PK_LINEAR = 0,               /* power.py:81-83   nonlinear_fn=power.linear  */
  JC_PK_HALOFIT_TAKAHASHI2012 = 1, /* power.py:144-262 nonlinear_fn=power.halofit */
  JC_PK_HALOFIT_SMITH2003 = 2      /* power.py:182-198,239-242  partial(power.halofit, prescription="smith2003") */
};
enum {
  JC_TF_EISENSTEIN_HU_OSC = 1,     /* transfer.py:10-156, type="eisenhu_osc" (default) */
  JC_TF_EISENSTEIN_HU_NOWIGGLE = 2 /* transfer.py:99-105, partial(Eisenstein_Hu, type="eisenhu") */
};

enum {
  JC_GROWTH_ODE = 0,  /* background.py:443-488  linear growth ODE (cosmologies with gamma=None)         */
  JC_GROWTH_GAMMA = 1 /* background.py:515-582  f = Omega_m(a)^gamma, D = exp(int f dln a), D(1) = 1   */
};

/* One redshift bin.  `shifts` is the chain of systematic_shift wrappers (redshift.py:159-171),
 * outermost first: pz_fn(z) = parent.pz_fn(clip(z - shift, 0)).  `zmax` is the n(z)'s own
 * normalisation range (redshift.py:16,29-30). */
typedef struct jc_nz {
  int32_t family;
  int32_t n_shifts;
  double params[4];
  double shifts[JC_MAX_SHIFTS];
  double gals_per_arcmin2;
  double zmax;
  const double* kde_z; /* JC_NZ_KDE: HOST arrays [kde_n] (catalogue redshifts, weights); read during   */
  const double* kde_w; /*            jc_plan_create only                                                */
  int64_t kde_n;
  double kde_bw;       /* bandwidth (config["bw"])                                                      */
} jc_nz;

typedef struct jc_bias {
  int32_t family;
  int32_t reserved;
  double params[3];
} jc_bias;

/* One tracer = one redshift bin of one probe, in probe-list order then bin order. */
typedef struct jc_tracer {
  int32_t kind;       /* JC_TRACER_*                                                        */
  int32_t ia_enabled; /* WL only: add the NLA kernel (probes.py:201-203) with `bias` as b_IA */
  jc_nz nz;
  jc_bias bias;      /* NC: galaxy bias; WL: IA bias when ia_enabled                        */
  double m_bias;     /* WL multiplicative bias m: kernel *= (1+m), probes.py:205-207        */
  double sigma_e;    /* WL shape noise, probes.py:210-223                                   */
  double probe_zmax; /* max zmax over the bins of the owning probe (probes.py:24,179-186)   */
} jc_tracer;

typedef struct jc_problem {
  int32_t abi_version; /* must be JC_ABI_VERSION */
  int32_t n_tracers;
  int32_t transfer;  /* JC_TF_*  */
  int32_t nonlinear; /* JC_PK_*  */
  int32_t growth;    /* JC_GROWTH_*: selects the growth-factor table and the cosmology row width */
  int32_t reserved;
  jc_tracer tracers[JC_MAX_TRACERS];
} jc_problem;

/* Offsets (in doubles) of the per-stage tables inside a workspace for n_cosmo cosmologies
 * processed as ONE chunk.  Exposed so that tests can check every stage against the oracle. */
typedef struct jc_ws_layout {
  int64_t chunk;       /* cosmologies per pass for the given workspace size                  */
  int64_t node_stride; /* padded length of a per-node array (>= 513)                         */
  int64_t ell_stride;  /* padded n_ell of a V row                                            */
  int64_t tracer_stride; /* padded T of an R row                                             */
  int64_t chitab;      /* [chunk, 256]  chi(a) table, background.py:223-236                  */
  int64_t gtab;        /* [chunk, 128]  D(a)/D(1) table, background.py:461-481               */
  int64_t scal;        /* [chunk, 32]   per-cosmology scalars (EH constants, pknorm, ...)    */
  int64_t stab;        /* [chunk, 256]  halofit S(R) table (sigma^2(R, a) = D(a)^2 S(R))     */
  int64_t node;        /* [chunk, JC_NODE_FIELDS, node_stride] per-Limber-node arrays        */
  int64_t rker;        /* [chunk, node_stride, tracer_stride] radial kernels R_i(a_n), node-major */
  int64_t vtab;        /* [chunk, 513, ell_stride] V[n, l] = w_n P(k_ln, a_n) dchi/da/chi^2/c^2 */
  int64_t ellpow;      /* [chunk, ell_stride]  (l+1/2)^(3+n_s)                                */
  int64_t total;       /* doubles                                                            */
} jc_ws_layout;

/* fields of the per-node table (index into ws.node) */
enum {
  JC_NODE_CHI = 0,    /* chi(a_n) >= 0                                  */
  JC_NODE_INVCHIC,    /* 1 / max(chi, 1)                                */
  JC_NODE_LNCHIC,     /* ln max(chi, 1)                                 */
  JC_NODE_GEOM,       /* w_n dchi/da / max(chi^2, 1) / c^2              */
  JC_NODE_GROWTH,     /* D(a_n) in [0, 1]                               */
  JC_NODE_HUBBLE,     /* H(a_n) = 100 sqrt(E^2)                         */
  JC_NODE_AMP,        /* D^2 pknorm / (2 pi^2)                          */
  JC_NODE_RNL,        /* 1 / k_nl                                       */
  JC_NODE_LNKNL,      /* ln k_nl                                        */
  JC_NODE_NEFF,       /* n_eff                                          */
  JC_NODE_CURV,       /* C                                              */
  JC_NODE_AN,         /* a_n                                            */
  JC_NODE_BN,         /* b_n                                            */
  JC_NODE_LNCF,       /* ln(c_n f3)                                     */
  JC_NODE_P3,         /* 3 - gamma_n                                    */
  JC_NODE_ALPHA,      /* alpha_n                                        */
  JC_NODE_BETA,       /* beta_n                                         */
  JC_NODE_NU,         /* nu_n                                           */
  JC_NODE_E1,         /* 3 f1                                           */
  JC_NODE_E2,         /* f2                                             */
  JC_NODE_NQ108,      /* (13.41 k_eq max(chi,1))^-1.08 : q^1.08 = (l+1/2)^1.08 * this   */
  JC_NODE_NSILK,      /* (k_silk max(chi,1))^-1.4      : (k/k_silk)^1.4 = (l+1/2)^1.4 * this */
  JC_NODE_NAMP,       /* max(chi,1)^-(3+n_s) D^2 pknorm/(2 pi^2) : Delta^2_L = (l+1/2)^(3+n_s) T^2 * this */
  JC_NODE_GK,         /* GEOM * 2 pi^2 max(chi,1)^3    : V = Delta^2 (l+1/2)^-3 * this */
  JC_NODE_MU,         /* mu_n (0 for takahashi2012, power.py:223)       */
  JC_NODE_FIELDS
};
/* fields of the per-cosmology scalar block (index into ws.scal) */
enum {
  JC_SCAL_LN13KEQ = 0, /* ln(13.41 k_eq)      transfer.py:116   */
  JC_SCAL_INV13KEQ,    /* 1 / (13.41 k_eq)                      */
  JC_SCAL_BETA_C,      /* transfer.py:110-111                   */
  JC_SCAL_C14_ALPHA_C, /* 14.2 / alpha_c      transfer.py:118   */
  JC_SCAL_SH_D,        /* sound horizon       transfer.py:72-77 */
  JC_SCAL_LNKSILK,     /* ln k_silk           transfer.py:79-85 */
  JC_SCAL_ALPHA_B,     /* transfer.py:131                       */
  JC_SCAL_BETA_B,      /* transfer.py:136                       */
  JC_SCAL_BETA_NODE,   /* transfer.py:133                       */
  JC_SCAL_FB,
  JC_SCAL_FC,
  JC_SCAL_NS,
  JC_SCAL_PKNORM,      /* sigma8^2 / sigmasqr(8)  power.py:47   */
  JC_SCAL_SIGMASQR8,   /* raw sigmasqr(cosmo, 8)  power.py:56-78 */
  JC_SCAL_OMEGA_M,
  JC_SCAL_ALPHA_GAMMA, /* no-wiggle fit  transfer.py:87-91    */
  JC_SCAL_OMH_T27,     /* Omega_m h / (tcmb/2.7)^2: q = k / (this * gamma shape), transfer.py:92-100 */
  JC_SCAL_MOVES_R,     /* forward-mode passes only: tangent plane k holds 1 when direction k can change the tracer kernels
                          R_i(a) (a component along Omega_c, Omega_b, Omega_k, w0, wa or gamma), 0 when dR = 0 identically
                          (h, n_s, sigma8 reach C_ell through P(k) alone); the value plane holds 0 */
  JC_SCAL_FIELDS = 32
};

typedef struct jc_plan jc_plan; /* opaque: cosmology-independent device tables for one problem */

/* Build the plan for (problem, ell) on CUDA device `device`: validates the problem, uploads the
 * quadrature grids / interpolation brackets / Romberg weights and runs the one-time n(z) kernels.
 * Synchronous; allocates device memory owned by the plan. */
int jc_plan_create(const jc_problem* problem, const double* ell_host, int32_t n_ell,
                   int32_t device, jc_plan** plan_out);
void jc_plan_destroy(jc_plan* plan);

int32_t jc_plan_n_tracers(const jc_plan* plan);
int32_t jc_plan_n_cls(const jc_plan* plan); /* P = T(T+1)/2 */
int32_t jc_plan_n_ell(const jc_plan* plan);
int32_t jc_plan_n_cosmo_params(const jc_plan* plan); /* width of a cosmology / tangent row: 8 or 9 */

/* Workspace: recommended size for n_cosmo cosmologies (processed in chunks of at most
 * JC_MAX_CHUNK) and the table offsets for a given size.  Any ws_bytes >= jc_workspace_bytes(plan,
 * 1, ..) is accepted by jc_angular_cl_f64; smaller chunks just mean more passes. */
int jc_workspace_bytes(const jc_plan* plan, int64_t n_cosmo, size_t* bytes_out);
int jc_workspace_layout(const jc_plan* plan, size_t ws_bytes, jc_ws_layout* layout_out);

/* angular_cl for a batch of cosmologies (replaces jax_cosmo/angular_cl.py:49-98; B == 1 is the
 * reference call).  cosmo_dev [B,8], cl_dev [B,P,L], ws_dev scratch.  Asynchronous on `stream`. */
int jc_angular_cl_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo,
                      double* cl_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* Forward-mode derivatives in the same pipeline (BASELINE config 4; what jax.jacfwd of the reference's
 * angular_cl returns, notebook docs/notebooks/jax-cosmo-intro.ipynb:989): for each of the n_tangents
 * directions tangents_dev[k, 0..7] in parameter space (same order as a cosmo row), dcl[b, k] is the
 * directional derivative of cl[b] along it -- the derivative of the discretised program with
 * interpolation / root indices and clip branches frozen at the evaluation point.  cl_dev may be NULL.
 * cosmo_dev [B,8], tangents_dev [K,8], cl_dev [B,P,L], dcl_dev [B,K,P,L].
 * Throughput path (jc_workspace_bytes_jvp_group(plan, B, K) sizes the workspace):
 *   K1 / K2 run on value + g <= 4 tangents per pass (DualN<g> kernels: the exp / log / reciprocals are computed once per
 *   group; the 7 wCDM parameters run as 4 + 3);
 *   K3 for 3 <= K <= 8 directions (Eisenstein-Hu with wiggles) takes ONE reverse sweep of the (ell, node) point function and
 *   forms every directional derivative from it with one multiply-add per input (jc_power_adj.cu); otherwise K3 runs on
 *   DualN<g> as well;
 *   K4: one tangent contraction per direction -- two products, one for a direction that cannot move the tracer kernels
 *   (no component along Omega_c, Omega_b, Omega_k, w0, wa, gamma: dR = 0 identically; JC_SCAL_MOVES_R).  Order such
 *   directions LAST: in the reverse-sweep path a second tangent group made only of them skips its K2 pass as well.
 *   With a smaller workspace (jc_workspace_bytes_jvp() = 2 x jc_workspace_bytes() is the minimum) the tangent groups shrink,
 *   down to one direction per pass.  jc_set_option("jvp_group", 1..4) caps g, jc_set_option("jvp_adjoint", 0) disables the
 *   reverse sweep (A/B partners of the tests).
 * Small batches: when the workspace has room for B*K <= 512 entries (ask jc_workspace_bytes_jvp(plan, B*K)), all K
 * directions of every cosmology run in ONE pass over B*K one-direction entries instead of latency-bound passes (a Jacobian
 * at one cosmology: 0.58 ms instead of 2.5 ms).  The modes differ by rounding only (<= 1e-12 of a spectrum's largest derivative). */
int jc_workspace_bytes_jvp(const jc_plan* plan, int64_t n_cosmo, size_t* bytes_out);
int jc_workspace_bytes_jvp_group(const jc_plan* plan, int64_t n_cosmo, int32_t n_tangents, size_t* bytes_out);
int jc_angular_cl_jvp_f64(const jc_plan* plan, const double* cosmo_dev, const double* tangents_dev,
                          int32_t n_tangents, int64_t n_cosmo, double* cl_dev, double* dcl_dev,
                          void* ws_dev, size_t ws_bytes, void* stream);

/* Same call with HOST buffers: copies cosmo in, runs, copies cl out, synchronises.  Uses a
 * device arena owned by the plan (grown on first use).  This is what the Python drop-in calls. */
int jc_angular_cl_host_f64(jc_plan* plan, const double* cosmo_host, int64_t n_cosmo,
                           double* cl_host);

/* Per-tracer noise (replaces probes.py:210-223,274-281 / angular_cl.py:101-117): noise_host[T];
 * the [P,L] noise_cl has noise[i] on auto pairs (i,i) and 0 elsewhere. */
int jc_noise_f64(const jc_plan* plan, double* noise_host);

/* Gaussian covariance in the sparse block layout (replaces angular_cl.py:120-163, sparse=True):
 * cov[(ij),(mn),l] = (C_im C_jn + C_in C_jm) / ((2l+1) gradient(l) f_sky),  C = signal + noise.
 * cl_dev [B,P,L] signal, noise_dev [T], cov_dev [B,P,P,L].  Uses the plan's ell. */
int jc_gaussian_cov_f64(const jc_plan* plan, const double* cl_dev, const double* noise_dev,
                        int64_t n_cosmo, double f_sky, double* cov_dev, void* stream);

/* Gaussian log-likelihood on the sparse block covariance (replaces jax_cosmo/likelihood.py:9-61 with
 * sparse.inv / sparse.slogdet, jax_cosmo/sparse.py:295-366; BASELINE config 3):
 *   loglike[b] = -0.5 * (r^T C^-1 r - log det C),  r = mu[b] - data   (the reference's sign convention;
 *   without the log-determinant term when include_logdet == 0).
 * data_dev [N] (data_stride = 0) or [B, N] (data_stride = N), mu_dev [B, N], N = P*L in cls-major order;
 * cov_dev [B, P, P, L] with SPD slices cov[b, :, :, l]; loglike_dev [B]; scratch_dev [B, L, 2] doubles.
 * JC_ERR_UNSUPPORTED if a slice does not fit shared memory (P > 238). */
int jc_gaussian_loglike_f64(const double* data_dev, int64_t data_stride, const double* mu_dev,
                            const double* cov_dev, int64_t n_cosmo, int32_t P, int32_t L,
                            int32_t include_logdet, double* loglike_dev, double* scratch_dev, void* stream);

/* BASELINE config 3 end to end on the device, without forming the covariance: the Gaussian log-likelihood of a data
 * vector under the Gaussian covariance OF THE MODEL SPECTRA, i.e. the reference's
 *     mu, cov = gaussian_cl_covariance_and_mean(cosmo, ell, probes, f_sky=f_sky, sparse=True)     angular_cl.py:166-196
 *     lnL = gaussian_log_likelihood(data, mu, cov, include_logdet)                                likelihood.py:9-61
 * for a batch of cosmologies whose signal spectra cl_dev [B, P, L] are already on the device (jc_angular_cl_f64).
 * cov[(ij),(mn),l] = (C_im C_jn + C_in C_jm) / nu_l is the operator S -> C S C on symmetric T x T matrices in the basis
 * of the P unique pairs, so r^T cov^-1 r = nu/2 tr(C^-1 R C^-1 R) and log det cov = (T+1) log det C + T log 2 - P log nu
 * per ell (C = signal + noise, R = residual matrix, nu = (2l+1) gradient(l) f_sky): O(T^3) per slice instead of O(P^3),
 * 8 P bytes read instead of 8 P^2 (csrc/jc_cl_loglike.cu).  Equal to the two-call form in exact arithmetic.
 * data_dev [P*L] (data_stride = 0) or [B, P*L] (data_stride = P*L), noise_dev [T] (jc_noise_f64), loglike_dev [B],
 * scratch_dev [B, L, 2] doubles.  dcl_dev (may be NULL; needs include_logdet = 1): [B, P, L] cotangent d lnL / d cl[b,p,l]
 * INCLUDING the dependence of the covariance and of its determinant on the spectra -- contracted with the forward-mode
 * Jacobian by jc_vjp_f64 it is the gradient of the full likelihood, what jax.grad(likelihood) returns for the
 * reference's README example (README.md:17-27).  n_cosmo <= 65535 per call. */
int jc_gaussian_cl_loglike_f64(const jc_plan* plan, const double* cl_dev, const double* data_dev, int64_t data_stride,
                               const double* noise_dev, int64_t n_cosmo, double f_sky, int32_t include_logdet,
                               double* loglike_dev, double* dcl_dev, double* scratch_dev, void* stream);

/* Fisher matrix F[b] = J^T C^-1 J on the sparse block covariance (the reference's recipe
 * sparse.dot(dmu.T, sparse.inv(cov), dmu), docs/notebooks/jax-cosmo-intro.ipynb cell 51; pairs with
 * jc_angular_cl_jvp_f64): jac_dev [B, K, P*L] (one row per parameter, cls-major), cov_dev [B, P, P, L],
 * fisher_dev [B, K, K], scratch_dev [B, L, K*K + 1] doubles; K <= 16. */
int jc_fisher_f64(const double* jac_dev, const double* cov_dev, int64_t n_cosmo, int32_t n_params,
                  int32_t P, int32_t L, double* fisher_dev, double* scratch_dev, void* stream);

/* Reverse mode: grad_dev[b, k] = sum_n jac_dev[b, k, n] * cot_dev[b, n] -- the vector-Jacobian product
 * J^T g that jax.grad / jax.vjp of the reference deliver for a scalar function of the C_ell
 * (README.md:24 of the reference), from the forward-mode Jacobian of jc_angular_cl_jvp_f64: with
 * K <= 9 parameters, K tangent passes cost less than one reverse sweep through the Limber pipeline.
 * jac_dev [B, K, N] (N = P*L), cot_dev [B, N] (cot_stride = N) or [N] shared (cot_stride = 0), grad_dev [B, K].
 * One CTA per (b, k), fixed reduction order: bitwise reproducible. */
int jc_vjp_f64(const double* jac_dev, const double* cot_dev, int64_t cot_stride, int64_t n_cosmo,
               int32_t n_params, int64_t N, double* grad_dev, void* stream);

/* Stand-alone background / matter-power functions of the reference on a caller-chosen grid:
 *   radial_comoving_distance, transverse_comoving_distance, growth_factor, H     background.py:199-344,371-398,129-143
 *   linear_matter_power, nonlinear_matter_power (halofit)                       power.py:14-53,144-272
 * A grid plan runs the path's own setup and power kernels with the n_a <= 512 scale factors in place of the Limber
 * nodes and the n_k wavenumbers [h/Mpc] in place of (ell + 1/2) / chi: same tables, same interpolation rule, same
 * halofit root.  jc_angular_cl_* reject grid plans and jc_grid_eval_f64 rejects ordinary plans (JC_ERR_INVALID).
 * Outputs (device, any may be NULL): pk [B, n_a, n_k] in (Mpc/h)^3 (linear if the plan's nonlinear == JC_PK_LINEAR),
 * chi / chi_transverse [B, n_a] in Mpc/h, growth [B, n_a] (D(1) = 1), hubble [B, n_a] in km/s/(Mpc/h),
 * transfer [B, n_k]: the Eisenstein-Hu T(k) of the plan's fit (transfer.py:10-156),
 * kernels [B, T, n_a]: the probes' radial kernels WeakLensing.kernel / NumberCounts.kernel (probes.py:188-281) without
 * the weak-lensing ell factor (probes.py:73), for plans made by jc_grid_plan_create_probes (NULL otherwise).
 * Workspace: jc_workspace_bytes(plan, B). */
int jc_grid_plan_create(int32_t transfer, int32_t nonlinear, int32_t growth, const double* k_host, int32_t n_k,
                        const double* a_host, int32_t n_a, int32_t device, jc_plan** plan_out);
int jc_grid_eval_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, double* pk_dev, double* chi_dev,
                     double* chi_transverse_dev, double* growth_dev, double* hubble_dev, double* transfer_dev,
                     double* kernels_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* The other public functions of background.py on a grid plan's scale factors (n_a <= 512):
 * aux_dev [B, JC_BG_FIELDS, n_a], rows JC_BG_*:
 *   growth_rate  dlnD/dlna: ODE plans -> a D'/D of the growth solution, interpolated like growth_factor
 *                (background.py:478-483, 491-512); gamma plans -> Omega_m(a)^gamma (background.py:551-584)
 *   Omega_m_a, Omega_de_a (background.py:145-196), dchioverda [Mpc/h] (background.py:270-294), w, f_de (background.py:25-90) */
enum { JC_BG_GROWTH_RATE = 0, JC_BG_OMEGA_M_A, JC_BG_OMEGA_DE_A, JC_BG_DCHIOVERDA, JC_BG_W, JC_BG_F_DE, JC_BG_FIELDS };
int jc_grid_background_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, double* aux_dev, void* ws_dev,
                           size_t ws_bytes, void* stream);

/* background.a_of_chi (background.py:245-267): scale factor at comoving distance chi [Mpc/h] by the reference's interp()
 * on the cosmology's (decreasing) 256-point chi table, neighbour rule as written in scipy/interpolate.py:25-37.
 * chi_dev [n_chi] is shared by all cosmologies, a_dev [B, n_chi].  Any plan (its growth mode fixes the row width). */
int jc_a_of_chi_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, const double* chi_dev, int64_t n_chi,
                    double* a_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* power.sigmasqr (power.py:56-78): sigma^2(R) of the UNNORMALISED spectrum T(k)^2 k^n_s (primordial_matter_power,
 * power.py:14-18) with a top-hat window, Romberg over 129 points (divmax = 7) between log10(kmin) and log10(kmax) with
 * k = e^x (the reference's limits as written).  R_dev [n_R] in Mpc/h (shared by all cosmologies), out_dev [B, n_R].  Only the reference's default
 * kmin = 1e-4, kmax = 1e3 are tabulated: other limits return JC_ERR_UNSUPPORTED.  The plan's transfer fit is used. */
int jc_sigmasqr_f64(const jc_plan* plan, const double* cosmo_dev, int64_t n_cosmo, const double* R_dev, int32_t n_R,
                    double kmin, double kmax, double* out_dev, void* ws_dev, size_t ws_bytes, void* stream);
int jc_grid_plan_create_probes(const jc_problem* problem, const double* a_host, int32_t n_a, int32_t device,
                               jc_plan** plan_out);

/* redshift_distribution.__call__ (redshift.py:27-31): the normalised n(z) = pz_fn(z) / simps(pz_fn, 0, zmax, 256) of
 * one bin (smail / fu / kde, under its systematic_shift chain), evaluated by the device functions the plan tables are
 * built from.  HOST pointers, synchronous.  JC_ERR_UNSUPPORTED for delta_nz (not a distribution). */
int jc_nz_eval_f64(const jc_nz* nz, const double* z_host, int64_t n, double* out_host);

/* Diagnostics: run a subset of the pipeline stages of one chunk (n_cosmo <= the workspace's chunk) on `stream`.
 * stage_mask bits: 0 setup, 1 lensing efficiency, 2 tracer finish, 3 power, 4 contraction (persistent TMA kernel),
 * 5 contraction (8-warp cp.async kernel, one CTA per SM).  Used by scripts/overlap_probe.py. */
int jc_debug_stages_f64(const jc_plan* plan, int32_t stage_mask, const double* cosmo_dev, int64_t n_cosmo, double* cl_dev,
                        void* ws_dev, size_t ws_bytes, void* stream);

/* jax_cosmo.sparse on the device (sparse.py): a block matrix of [ny, nx] diagonal blocks of size n is S[ny, nx, n].
 * jc_sparse_bmm_f64: C[i,k,l] = sum_j A[i,j,l] * B[j,k,l] for i < I, j < J, k < K, l < L, every operand addressed
 * by element strides (in doubles; a stride of 0 broadcasts) -- one entry point behind sparse.dot's seven
 * combinations (sparse.py:72-292): sparse @ vec, sparse @ dense, vec @ sparse, dense @ sparse, sparse @ sparse and
 * dense @ sparse @ dense (two calls, the second with L = 1).  Fixed summation order over j. */
int jc_sparse_bmm_f64(const double* A_dev, int64_t sAi, int64_t sAj, int64_t sAl, const double* B_dev, int64_t sBj,
                      int64_t sBk, int64_t sBl, double* C_dev, int64_t sCi, int64_t sCk, int64_t sCl, int32_t I,
                      int32_t J, int32_t K, int32_t L, void* stream);
/* sparse.inv / sparse.slogdet / sparse.det (sparse.py:296-389) for a square sparse matrix S[P, P, L]: per diagonal
 * position l the inverse of S[:, :, l] and sign / log|det| by Gauss-Jordan elimination with partial pivoting (what
 * np.linalg.inv / slogdet do per slice in the reference).  inv_dev [P, P, L], sign_dev [L], logdet_dev [L] (any of
 * the three may be NULL); scratch_dev [L, P, 2P] doubles.  The reference's slogdet of the whole matrix is
 * (prod_l sign[l], sum_l logdet[l]). */
int jc_sparse_inv_f64(const double* sparse_dev, int32_t P, int32_t L, double* inv_dev, double* sign_dev,
                      double* logdet_dev, double* scratch_dev, void* stream);

/* ---- multi-GPU: the final gather of the per-rank [B/G, P, L] blocks over NVLink (the path's one exchange) --------
 * The reference has no multi-device code (no pmap / shard_map / psum); a batch of cosmologies is the
 * data-parallel axis this path adds, and a gather of the shards is what `jax.vmap(angular_cl)` over a sharded batch
 * would end with.  One rank per GPU.  Every rank owns a FULL-size result buffer [rows_total, P, L] in device memory
 * (jc_gather_create allocates it and exports a CUDA IPC handle); after the handles have been exchanged
 * (jc_gather_connect_ipc: other processes of the box; jc_gather_connect_local: other devices of this process) a rank
 * computes its rows in sub-chunks straight into its own buffer and the copy engines push each finished sub-chunk into
 * the same rows of every peer's buffer while the next sub-chunk computes (cudaMemcpyAsync over NVLink peer memory on
 * side streams -- no SM is taken from the FP64 kernels).  On return `stream` is ordered after this rank's OUTGOING
 * pushes; the rows the peers write are complete once every rank has got there: close the step with a stream-ordered
 * cross-rank barrier (e.g. a one-element NCCL all-reduce).  Results are bitwise those of jc_angular_cl_f64. */
#define JC_IPC_HANDLE_BYTES 64
#define JC_MAX_RANKS 16
typedef struct jc_gather jc_gather;
/* push_sms: 0 = copy-engine transport (cudaMemcpyAsync per slice and peer); > 0 = a persistent pusher kernel on that many
 * SMs stores every finished slice to all peers with the destinations interleaved (uniform all-to-all NVLink traffic; needed
 * at 8 GPUs, where many medium-sized copy-engine copies lose 40 % of the link rate); the contraction then leaves those SMs
 * out of its persistent grid. */
int jc_gather_create(int32_t rank, int32_t world, int32_t device, size_t bytes, int32_t push_sms, jc_gather** gather_out,
                     unsigned char* ipc_handle_out /* [JC_IPC_HANDLE_BYTES] or NULL */);
int jc_gather_status(jc_gather* gather, int32_t* pusher_aborted_out); /* synchronous; 1: the pusher timed out waiting */
void* jc_gather_buffer(const jc_gather* gather); /* this rank's device buffer */
int jc_gather_connect_ipc(jc_gather* gather, const unsigned char* ipc_handles /* [world][JC_IPC_HANDLE_BYTES] */);
int jc_gather_connect_local(jc_gather* gather, void* const* buffers /* [world] */, const int32_t* devices /* [world] */);
int jc_gather_destroy(jc_gather* gather);
/* cosmo_dev [n_cosmo, 8|9] = this rank's rows, which land in rows [row_offset, row_offset + n_cosmo) of every buffer;
 * sub_chunk = cosmologies per compute chunk of K1..K3 (< 1: as many as the workspace holds), push_rows = cosmologies per
 * contraction launch + push (< 1: the whole chunk); ws as for jc_angular_cl_f64.  equal_shards != 0: every rank calls with
 * the same n_cosmo / sub_chunk / push_rows (B divisible by the number of ranks) -- the copy-engine transport then
 * runs the pushes of a slice in lockstep on all ranks (a flag barrier per slice by stream memory operations; every
 * rank MUST make the call, a missing one blocks the others' copy streams). */
int jc_angular_cl_gather_f64(const jc_plan* plan, jc_gather* gather, const double* cosmo_dev, int64_t n_cosmo,
                             int64_t row_offset, int64_t sub_chunk, int64_t push_rows, int32_t equal_shards, void* ws_dev,
                             size_t ws_bytes, void* stream);
/* the exchange alone: push rows [row_offset, row_offset + rows) (row_bytes each) of the local buffer to every peer */
int jc_gather_push_f64(jc_gather* gather, size_t row_bytes, int64_t row_offset, int64_t rows, void* stream);

/* Per-stage device timing (CUDA events recorded on the launch stream between the stages of
 * jc_angular_cl_f64).  Stages: 0 setup, 1 lensing efficiency, 2 tracer finish, 3 power, 4 pair
 * contraction.  While enabled the plan is not re-entrant.  jc_profile_read synchronises the
 * recorded events, returns the summed milliseconds and kernel-launch counts per stage since the
 * last read, and resets the counters. */
#define JC_N_STAGES 5
int jc_profile_enable(jc_plan* plan, int32_t enable);
int jc_profile_read(jc_plan* plan, double* stage_ms, int64_t* stage_launches);

/* FP64 roofline probe: runs an FMA-only kernel (mode 0: DFMA chains, mode 1: DMMA m8n8k4,
 * mode 2: both interleaved in every warp -- tells whether they share a datapath) on the
 * current device for ~`seconds` and returns TFLOP/s.  Used by bench.py for the roofline
 * denominator, which MEASURED_PEAKS.json does not hold for FP64. */
int jc_fp64_peak_tflops(int32_t mode, double seconds, double* tflops_out);

/* Self-test hook for the kernels' own FP64 elementary functions (csrc/jc_math.cuh):
 * y[i] = fn(x[i]) on device arrays; fn: 0 exp, 1 log, 2 sin, 3 x^(-1/3), 4 1/x, 5 table-driven exp,
 * 6 table-driven log (x >= 1).  Synchronises the stream. */
int jc_debug_math_f64(int32_t fn, const double* x_dev, double* y_dev, int64_t n, void* stream);

/* Process-wide options (diagnostics / A-B runs; defaults are the hot path):
 *   "power_exact"   0 | 1   1: the power kernel evaluates the Eisenstein-Hu formula at every (ell, node) point instead of
 *                           interpolating the per-cosmology T(k) table (default 0; also env JC_POWER_EXACT);
 *   "contract_eps"  >= 0    relative threshold below which a number-counts n(z_n) counts as zero when the contraction's
 *                           node ranges are planned; 0 = exact zeros only (bitwise the full sum); read by jc_plan_create
 *                           (default 1e-20; also env JC_CONTRACT_EPS);
 *   "contract_kernel" 0..3  0: persistent TMA contraction where it applies (>= 17 pair tiles), 3: the cp.async kernel
 *                           everywhere (all node stages, the reference's pair order) -- the A/B partner of the tests;
 *   "jvp_group"     1..4    tangent directions carried per pass of jc_angular_cl_jvp_f64 (default 4; also env JC_JVP_GROUP);
 *   "jvp_adjoint"   0 | 1   1: K3 of jc_angular_cl_jvp_f64 by one reverse sweep for 3..8 directions (default 1; env JC_JVP_ADJOINT);
 *   "lens_mma"      0 | 1   1: lensing-efficiency launches of >= 8 sources run on the FP64 tensor-core kernel (env JC_LENS_MMA). */
int jc_set_option(const char* name, double value);
int jc_get_option(const char* name, double* value_out);

const char* jc_status_string(int status);
const char* jc_last_cuda_error(void);
int32_t jc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JC_B200_H */
