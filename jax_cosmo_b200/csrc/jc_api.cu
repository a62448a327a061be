// jc_api.cu -- remaining C-ABI entry points: host-buffer call, Gaussian covariance, FP64 roofline
// probe, status strings.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "jc_internal.cuh"
#include "jc_math.cuh"

static thread_local char g_cuda_err[512] = "";

void jc_set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

extern "C" const char* jc_last_cuda_error(void) { return g_cuda_err; }

// ---- process-wide options (jc_set_option) -------------------------------------------------------------------------
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static double env_double(const char* name, double dflt) { const char* e = getenv(name); return e ? atof(e) : dflt; }
int g_jc_power_exact = env_int("JC_POWER_EXACT", 0);
double g_jc_contract_eps = env_double("JC_CONTRACT_EPS", 1e-20);
int g_jc_jvp_group = env_int("JC_JVP_GROUP", JC_JVP_MAX_GROUP);
int g_jc_jvp_adjoint = env_int("JC_JVP_ADJOINT", 1);
int g_jc_lens_mma = env_int("JC_LENS_MMA", 0);

extern "C" int jc_set_option(const char* name, double value) {
  if (!name) return JC_ERR_INVALID;
  if (!strcmp(name, "power_exact")) { g_jc_power_exact = value != 0.0; return JC_OK; }
  if (!strcmp(name, "contract_eps")) { if (!(value >= -1.0) || value > 1e-6) return JC_ERR_INVALID; g_jc_contract_eps = value; return JC_OK; }
  if (!strcmp(name, "contract_kernel")) { const int v = (int)value; if (v < 0 || v > 3) return JC_ERR_INVALID; g_contract_cfg = v; return JC_OK; }
  if (!strcmp(name, "jvp_group")) { const int v = (int)value; if (v < 1 || v > JC_JVP_MAX_GROUP) return JC_ERR_INVALID; g_jc_jvp_group = v; return JC_OK; }
  if (!strcmp(name, "jvp_adjoint")) { g_jc_jvp_adjoint = value != 0.0; return JC_OK; }
  if (!strcmp(name, "lens_mma")) { g_jc_lens_mma = value != 0.0; return JC_OK; }
  return JC_ERR_INVALID;
}
extern "C" int jc_get_option(const char* name, double* value_out) {
  if (!name || !value_out) return JC_ERR_INVALID;
  if (!strcmp(name, "power_exact")) { *value_out = g_jc_power_exact; return JC_OK; }
  if (!strcmp(name, "contract_eps")) { *value_out = g_jc_contract_eps; return JC_OK; }
  if (!strcmp(name, "contract_kernel")) { *value_out = g_contract_cfg < 0 ? 0 : g_contract_cfg; return JC_OK; }
  if (!strcmp(name, "jvp_group")) { *value_out = g_jc_jvp_group; return JC_OK; }
  if (!strcmp(name, "jvp_adjoint")) { *value_out = g_jc_jvp_adjoint; return JC_OK; }
  if (!strcmp(name, "lens_mma")) { *value_out = g_jc_lens_mma; return JC_OK; }
  return JC_ERR_INVALID;
}
extern "C" int32_t jc_abi_version(void) { return JC_ABI_VERSION; }

extern "C" const char* jc_status_string(int status) {
  switch (status) {
    case JC_OK: return "ok";
    case JC_ERR_INVALID: return "invalid argument";
    case JC_ERR_UNSUPPORTED: return "configuration not supported by the B200 path (no fallback)";
    case JC_ERR_WORKSPACE: return "workspace too small";
    case JC_ERR_CUDA: return "CUDA runtime error";
    case JC_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry: H2D cosmologies, chunked compute on one stream, D2H of each finished chunk on
// a second stream (double-buffered), so PCIe traffic overlaps the FP64 kernels when the host
// buffers are pinned.
// ---------------------------------------------------------------------------------------------
// 2 cosmologies per SM and chunk: the D2H stream is the bottleneck (PCIe, ~54 GB/s), so the exposed part is the
// first chunk's compute; measured e2e at config 5: 1024 -> 6.15e9, 592 -> 6.26e9, 444 -> 6.34e9, 296 -> 6.40e9 C_ell/s
#define JC_HOST_CHUNK 296

static int ensure(void** p, size_t* have, size_t need) {
  if (*have >= need) return JC_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  JC_CUDA_TRY(cudaMalloc(p, need));
  *have = need;
  return JC_OK;
}

extern "C" int jc_angular_cl_host_f64(jc_plan* plan, const double* cosmo_host, int64_t n_cosmo,
                                      double* cl_host) {
  if (!plan || !cosmo_host || !cl_host || n_cosmo < 1) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);
  JC_CUDA_TRY(guard.status);
  if (!plan->s_compute) {
    JC_CUDA_TRY(cudaStreamCreateWithFlags(&plan->s_compute, cudaStreamNonBlocking));
    JC_CUDA_TRY(cudaStreamCreateWithFlags(&plan->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      JC_CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_done[i], cudaEventDisableTiming));
      JC_CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_copied[i], cudaEventDisableTiming));
    }
  }
  static int64_t host_chunk = 0;  // idempotent; JC_HOST_CHUNK env = tuning knob (profiles/r01_tuning.md)
  if (!host_chunk) {
    const char* e = getenv("JC_HOST_CHUNK");
    const int64_t v = e ? atoll(e) : 0;
    host_chunk = v > 0 ? v : JC_HOST_CHUNK;
  }
  const int64_t chunk = n_cosmo < host_chunk ? n_cosmo : host_chunk;
  const size_t pl_elems = (size_t)plan->d.P * plan->d.L;
  size_t ws_need = 0;
  int st = jc_workspace_bytes(plan, chunk, &ws_need);
  if (st != JC_OK) return st;
  if ((st = ensure(&plan->arena_ws, &plan->arena_ws_bytes, ws_need)) != JC_OK) return st;
  if ((st = ensure((void**)&plan->arena_cosmo, &plan->arena_cosmo_bytes,
                   (size_t)n_cosmo * plan->d.ncp * sizeof(double))) != JC_OK) return st;
  size_t cl_need = (size_t)chunk * pl_elems * sizeof(double);
  if (plan->arena_cl_bytes < cl_need) {
    for (int i = 0; i < 2; ++i) {
      if (plan->arena_cl[i]) cudaFree(plan->arena_cl[i]);
      plan->arena_cl[i] = nullptr;
    }
    plan->arena_cl_bytes = 0;
    for (int i = 0; i < 2; ++i) JC_CUDA_TRY(cudaMalloc((void**)&plan->arena_cl[i], cl_need));
    plan->arena_cl_bytes = cl_need;
  }
  JC_CUDA_TRY(cudaMemcpyAsync(plan->arena_cosmo, cosmo_host, (size_t)n_cosmo * plan->d.ncp * sizeof(double),
                              cudaMemcpyHostToDevice, plan->s_compute));
  int k = 0;
  for (int64_t c0 = 0; c0 < n_cosmo; c0 += chunk, ++k) {
    const int b = k & 1;
    const int64_t nc = (n_cosmo - c0) < chunk ? (n_cosmo - c0) : chunk;
    if (k >= 2) JC_CUDA_TRY(cudaStreamWaitEvent(plan->s_compute, plan->ev_copied[b], 0));
    st = jc_angular_cl_f64(plan, plan->arena_cosmo + c0 * plan->d.ncp, nc, plan->arena_cl[b],
                           plan->arena_ws, plan->arena_ws_bytes, plan->s_compute);
    if (st != JC_OK) return st;
    JC_CUDA_TRY(cudaEventRecord(plan->ev_done[b], plan->s_compute));
    JC_CUDA_TRY(cudaStreamWaitEvent(plan->s_copy, plan->ev_done[b], 0));
    JC_CUDA_TRY(cudaMemcpyAsync(cl_host + (size_t)c0 * pl_elems, plan->arena_cl[b], (size_t)nc * pl_elems * sizeof(double),
                                cudaMemcpyDeviceToHost, plan->s_copy));
    JC_CUDA_TRY(cudaEventRecord(plan->ev_copied[b], plan->s_copy));
  }
  JC_CUDA_TRY(cudaStreamSynchronize(plan->s_compute));
  JC_CUDA_TRY(cudaStreamSynchronize(plan->s_copy));
  return JC_OK;
}

// ---------------------------------------------------------------------------------------------
// Gaussian covariance, sparse block layout [P,P,L] (angular_cl.py:120-163); HBM-write bound: 8 P^2 L bytes
// out (35.3 MB per cosmology at 10+10 bins) for 8 P L bytes in.
// One CTA per (cosmology, row pair p = (i, j)): the 2 T spectra the row needs -- C_im and C_jn for all m, n, noise
// already on the autos -- are staged in shared memory (2 T L doubles, 32 KB), then the CTA streams the P x L outputs of
// the row, ell fastest: 4 conflict-free LDS, 2 FMA and one coalesced 8-byte store per element.  (The first version
// recomputed four pair indices and gathered four values from L2 per element: 1.03 TB/s, 16 % of the HBM peak.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pair_index(int i, int j, int T) {  // angular_cl.py:34-38
  if (i > j) { int t = i; i = j; j = t; }
  return i * T - (i * (i - 1)) / 2 + (j - i);
}

__global__ void __launch_bounds__(256) jc_cov_kernel(JcDevPlan pl, const double* __restrict__ cl,
                                                     const double* __restrict__ noise, double f_sky,
                                                     double* __restrict__ cov) {
  extern __shared__ __align__(16) double s_cov[];
  const int L = pl.L, P = pl.P, T = pl.T;
  const int p = blockIdx.x;
  const size_t c = blockIdx.y;
  const int i = pl.pair_i[p], j = pl.pair_j[p];
  double* sI = s_cov;                  // [T][L]  C_im + delta_im noise_i   (cl_obs = signal + noise, angular_cl.py:112-115,135)
  double* sJ = sI + (size_t)T * L;     // [T][L]  C_jn + delta_jn noise_j
  double* sN = sJ + (size_t)T * L;     // [L]     1 / ((2l+1) gradient(l) f_sky)   (angular_cl.py:139)
  const double* C = cl + c * (size_t)P * L;
  for (int q = threadIdx.x; q < T * L; q += 256) {
    const int m = q / L, l = q - m * L;
    sI[q] = C[(size_t)pair_index(i, m, T) * L + l] + (i == m ? noise[i] : 0.0);
    sJ[q] = C[(size_t)pair_index(j, m, T) * L + l] + (j == m ? noise[j] : 0.0);
  }
  for (int l = threadIdx.x; l < L; l += 256) sN[l] = 1.0 / (pl.covnorm[l] * f_sky);
  __syncthreads();
  double* out = cov + (c * P + p) * (size_t)P * L;
  // thread = fixed ell lane(s), loop over the row's P column pairs (m, n)
  for (int q0 = threadIdx.x; q0 < P * L; q0 += 256) {
    const int q = q0 / L, l = q0 - q * L;
    const int m = pl.pair_i[q], n = pl.pair_j[q];
    const double v = (sI[m * L + l] * sJ[n * L + l] + sI[n * L + l] * sJ[m * L + l]) * sN[l];  // angular_cl.py:146-147
    out[q0] = v;
  }
}

extern "C" int jc_gaussian_cov_f64(const jc_plan* plan, const double* cl_dev, const double* noise_dev,
                                   int64_t n_cosmo, double f_sky, double* cov_dev, void* stream) {
  if (!plan || !cl_dev || !noise_dev || !cov_dev || n_cosmo < 1) return JC_ERR_INVALID;
  if (plan->d.L < 2) return JC_ERR_INVALID;  // np.gradient needs >= 2 points
  if (n_cosmo > 65535) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);
  const size_t smem = ((size_t)2 * plan->d.T * plan->d.L + plan->d.L) * sizeof(double);
  if (smem > 200 * 1024) return JC_ERR_UNSUPPORTED;
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, cudaFuncSetAttribute(jc_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  jc_cov_kernel<<<dim3(plan->d.P, (unsigned)n_cosmo), 256, smem, (cudaStream_t)stream>>>(plan->d, cl_dev, noise_dev, f_sky, cov_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

// ---------------------------------------------------------------------------------------------
// device-math self test
// ---------------------------------------------------------------------------------------------
__global__ void jc_debug_math_kernel(int fn, const double* __restrict__ x, double* __restrict__ y, int64_t n,
                                     const double* __restrict__ tab_g) {
  __shared__ __align__(16) double tab[JCM_TAB_DOUBLES];
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += blockDim.x) tab[i] = tab_g[i];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r;
  switch (fn) {
    case 0: r = jcm_exp(v); break;
    case 1: r = jcm_log(v); break;
    case 2: r = jcm_sin(v); break;
    case 3: r = jcm_rcbrt(v); break;
    case 4: r = jcm_rcp(v); break;
    case 5: r = jcm_exp_t(v, tab); break;
    default: r = jcm_log_t(v, tab); break;
  }
  y[i] = r;
}

extern "C" int jc_debug_math_f64(int32_t fn, const double* x_dev, double* y_dev, int64_t n, void* stream) {
  if (!x_dev || !y_dev || n < 1 || fn < 0 || fn > 6) return JC_ERR_INVALID;
  double tab_h[JCM_TAB_DOUBLES];
  jc_math_table(tab_h);
  double* tab_d = nullptr;
  JC_CUDA_TRY(cudaMalloc(&tab_d, sizeof(tab_h)));
  JC_CUDA_TRY(cudaMemcpy(tab_d, tab_h, sizeof(tab_h), cudaMemcpyHostToDevice));
  jc_debug_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fn, x_dev, y_dev, n, tab_d);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(tab_d);
  if (e != cudaSuccess) { jc_set_cuda_error(e, "jc_debug_math_kernel"); return JC_ERR_CUDA; }
  return JC_OK;
}

// ---------------------------------------------------------------------------------------------
// per-stage profiling
// ---------------------------------------------------------------------------------------------
extern "C" int jc_profile_enable(jc_plan* plan, int32_t enable) {
  if (!plan) return JC_ERR_INVALID;
  JcDeviceGuard guard(plan->device);
  JC_CUDA_TRY(guard.status);
  if (enable && !plan->prof) {
    plan->prof = new JcProf();
    memset(plan->prof, 0, sizeof(JcProf));
    for (int i = 0; i < JC_PROF_SLOTS; ++i)
      for (int j = 0; j <= JC_N_STAGES; ++j) JC_CUDA_TRY(cudaEventCreate(&plan->prof->ev[i][j]));
  }
  if (plan->prof) { plan->prof->enabled = enable ? 1 : 0; plan->prof->used = 0; }
  return JC_OK;
}

extern "C" int jc_profile_read(jc_plan* plan, double* stage_ms, int64_t* stage_launches) {
  if (!plan || !plan->prof || !stage_ms || !stage_launches) return JC_ERR_INVALID;
  JcProf* p = plan->prof;
  for (int j = 0; j < JC_N_STAGES; ++j) { stage_ms[j] = 0.0; stage_launches[j] = 0; }
  for (int i = 0; i < p->used; ++i) {
    JC_CUDA_TRY(cudaEventSynchronize(p->ev[i][JC_N_STAGES]));
    for (int j = 0; j < JC_N_STAGES; ++j) {
      float ms = 0.f;
      JC_CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[i][j], p->ev[i][j + 1]));
      stage_ms[j] += ms;
      stage_launches[j] += p->launches[i][j];
    }
  }
  p->used = 0;
  return JC_OK;
}

// ---------------------------------------------------------------------------------------------
// FP64 roofline probe
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) jc_dfma_probe_kernel(double* out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
  const double m = 1.0000000001, b = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;  // keep the chain alive
}

// NACC independent accumulator tiles per warp, 16 DMMA per iteration (NACC = 4) or NACC per iteration
template <int NACC>
__global__ void __launch_bounds__(512) jc_dmma_probe_kernel(double* out, int iters, double seed) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c0[i] = c1[i] = 0.0;
  double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16 / NACC + (16 % NACC ? 1 : 0); ++u) {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// The contraction's register tile: 2 A fragments x 8 B fragments -> 16 accumulator tiles, 16 DMMA per iteration.
// SRC 0: fragments stay in registers; SRC 1: every iteration reloads them from shared memory (4 + 8 LDS.64)
// and forms the A fragments as products (2 DMUL), like jc_contract's k-step.
template <int SRC>
__global__ void __launch_bounds__(512) jc_dmma_tile_probe_kernel(double* out, int iters, double seed) {
  __shared__ double sh[512 * 3];
  for (int i = threadIdx.x; i < 512 * 3; i += blockDim.x) sh[i] = seed + i * 1e-9;
  __syncthreads();
  double c0[2][8], c1[2][8];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) c0[m][n] = c1[m][n] = 0.0;
  double a[2], b[8];
  a[0] = seed + threadIdx.x * 1e-9; a[1] = a[0] + 1.0;
#pragma unroll
  for (int n = 0; n < 8; ++n) b[n] = 1.0 + n + threadIdx.x * 1e-12;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    if (SRC == 1) {
      const double* r = sh + ((it & 15) * 20 + (lane & 3) * 20 * 16) % 512;
      const double* v = sh + 512 + ((it & 15) * 60 + (lane & 3) * 60 + (lane >> 2)) % 448;
#pragma unroll
      for (int m = 0; m < 2; ++m) a[m] = r[(lane >> 2) + m * 8] * r[(lane >> 3) + 4 + m];
#pragma unroll
      for (int n = 0; n < 8; ++n) b[n] = v[n * 8];
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int m = 0; m < 2; ++m)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c0[m][n]), "+d"(c1[m][n]) : "d"(a[m]), "d"(b[n]));
  }
  double s = 0.0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n) s += c0[m][n] + c1[m][n];
  if (s == 12345.678) out[0] = s;
}

// mode 2: every warp interleaves 16 DFMA with 4 DMMA per iteration (128 FMA-lanes each side):
// if the two share one datapath the rate stays at the single-pipe peak, otherwise it adds up.
__global__ void __launch_bounds__(256) jc_mixed_probe_kernel(double* out, int iters, double seed) {
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = seed + threadIdx.x * 1e-9 + i;
  double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-12;
  const double m = 1.0000000001, bb = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fma(f[j], m, bb);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  if (s == 12345.678) out[0] = s;
}

extern "C" int jc_fp64_peak_tflops(int32_t mode, double seconds, double* tflops_out) {
  // mode & 15: 0 DFMA, 1 DMMA, 2 both interleaved.  DMMA only: (mode >> 4) & 15 = warps per SM sub-partition
  // (0 = full occupancy), mode >> 8 = 1 selects 16 independent accumulator tiles per warp instead of 4.
  const int wps = (mode >> 4) & 15, wide = mode >> 8;
  mode &= 15;
  if (!tflops_out || mode < 0 || mode > 2 || wps > 4 || wide > 3 || ((wps || wide) && mode != 1)) return JC_ERR_INVALID;
  int dev = 0, sms = 0;
  JC_CUDA_TRY(cudaGetDevice(&dev));
  JC_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d_out = nullptr;
  JC_CUDA_TRY(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  JC_CUDA_TRY(cudaEventCreate(&e0));
  JC_CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = wps ? sms : sms * 8, threads = wps ? 128 * wps : 256, iters = 4096;
  // flops per launch
  const double f_dfma = (double)blocks * threads * iters * 64.0 * 2.0;
  const double f_dmma = (double)blocks * (threads / 32) * iters * 16.0 * (8 * 8 * 4 * 2.0);
  // mode 2: per iteration and warp 16 DMMA (16*512 flop) + 128 DFMA instructions (128*64 flop)
  double flops = mode == 0 ? f_dfma : (mode == 1 ? f_dmma : f_dmma + (double)blocks * threads * iters * 128.0 * 2.0);
  auto launch = [&]() {
    if (mode == 0) jc_dfma_probe_kernel<<<blocks, threads>>>(d_out, iters, 1.0);
    else if (mode == 1 && wide == 3) jc_dmma_tile_probe_kernel<1><<<blocks, threads>>>(d_out, iters, 1.0);
    else if (mode == 1 && wide == 2) jc_dmma_tile_probe_kernel<0><<<blocks, threads>>>(d_out, iters, 1.0);
    else if (mode == 1 && wide) jc_dmma_probe_kernel<16><<<blocks, threads>>>(d_out, iters, 1.0);
    else if (mode == 1) jc_dmma_probe_kernel<4><<<blocks, threads>>>(d_out, iters, 1.0);
    else jc_mixed_probe_kernel<<<blocks, threads>>>(d_out, iters, 1.0);
  };
  launch();  // warm-up
  JC_CUDA_TRY(cudaDeviceSynchronize());
  auto t0 = std::chrono::steady_clock::now();
  double total_ms = 0.0, total_flops = 0.0;
  do {
    JC_CUDA_TRY(cudaEventRecord(e0));
    for (int i = 0; i < 4; ++i) launch();
    JC_CUDA_TRY(cudaEventRecord(e1));
    JC_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    JC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    total_ms += ms;
    total_flops += 4.0 * flops;
  } while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds);
  *tflops_out = total_flops / (total_ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return JC_OK;
}
