// jc_loglike.cu -- Gaussian log-likelihood and Fisher matrix on the sparse block covariance layout
// [P, P, L] (replaces likelihood.py:9-61 with sparse.inv / sparse.slogdet, sparse.py:295-366, and the
// notebook's Fisher recipe sparse.dot(dmu.T, sparse.inv(cov), dmu), jax-cosmo-intro.ipynb cell 51;
// BASELINE configs 3 and 4).
//
// The covariance is "a diagonal of blocks": for each ell the [P x P] slice C_l = cov[:, :, l] couples the
// P spectra, different ell are independent.  Hence for any vectors u, v of length P*L (cls-major)
//     u^T C^-1 v = sum_l u_l^T C_l^-1 v_l ,     log det C = sum_l log det C_l ,
// which is what the reference's sparse.inv (batched inverse over ell) and Schur-recursion slogdet compute.
// One CTA per (cosmology, ell): the slice is gathered into shared memory as a packed lower triangle
// (P(P+1)/2 doubles: 177 KB at P = 210) and factorised in place (right-looking Cholesky; Gaussian C_l
// covariances are SPD); the n_rhs right-hand sides are forward-substituted on the fly (Y = L^-1 U), and the
// slice contributes the Gram matrix Y^T Y (= U_l^T C_l^-1 U_l) and logdet_l = 2 sum log L_kk.  A second tiny
// kernel sums the per-ell partials in a fixed order (deterministic).
//   log-likelihood: one right-hand side r = mu - data, chi2 = Gram[0][0];
//   Fisher matrix:  the K rows of the Jacobian, F = Gram.
#include "jc_internal.cuh"

#define JC_MAX_RHS 16

namespace {

__device__ __forceinline__ int tri(int i, int j) { return (i * (i + 1) >> 1) + j; }  // i >= j

// rhs [B, n_rhs, P*L]; sub (optional) [*, P*L] with stride sub_stride is subtracted from every rhs;
// partial [B, L, n_rhs*n_rhs + 1]
__global__ void __launch_bounds__(512) jc_slice_solve_kernel(const double* __restrict__ cov,
                                                             const double* __restrict__ rhs, int n_rhs,
                                                             const double* __restrict__ sub, int64_t sub_stride,
                                                             int P, int L, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* A = sm;                               // packed lower triangle
  double* R = sm + ((size_t)P * (P + 1) >> 1);  // [n_rhs][P] right-hand sides -> solutions
  __shared__ double s_piv;
  const int l = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const size_t N = (size_t)P * L;
  const double* C = cov + (size_t)b * P * P * L + l;
  for (int i = warp; i < P; i += nwarps)
    for (int j = lane; j <= i; j += 32) A[tri(i, j)] = C[((size_t)i * P + j) * L];
  for (int q = tid; q < n_rhs * P; q += nthr) {
    const int j = q / P, i = q - j * P;
    double v = rhs[((size_t)b * n_rhs + j) * N + (size_t)i * L + l];
    if (sub) v -= sub[(size_t)b * sub_stride + (size_t)i * L + l];
    R[q] = v;
  }
  __syncthreads();

  double logdet = 0.0;  // accumulated by thread 0
  for (int k = 0; k < P; ++k) {
    if (tid == 0) {
      const double d = sqrt(A[tri(k, k)]);
      A[tri(k, k)] = d;
      s_piv = 1.0 / d;
      logdet += log(d);
    }
    __syncthreads();
    const double inv = s_piv;
    if (tid < n_rhs) R[tid * P + k] *= inv;  // y_k of each forward substitution (updates from columns < k are done)
    __syncthreads();
    for (int i = k + 1 + tid; i < P; i += nthr) {
      const double lik = A[tri(i, k)] * inv;
      A[tri(i, k)] = lik;
      for (int j = 0; j < n_rhs; ++j) R[j * P + i] -= lik * R[j * P + k];
    }
    __syncthreads();
    for (int i = k + 1 + warp; i < P; i += nwarps) {  // trailing update of the lower triangle
      const double lik = A[tri(i, k)];
      double* row = A + tri(i, 0);
      for (int j = k + 1 + lane; j <= i; j += 32) row[j] -= lik * A[tri(j, k)];
    }
    __syncthreads();
  }
  double* out = partial + ((size_t)b * L + l) * (n_rhs * n_rhs + 1);
  for (int pq = warp; pq < n_rhs * n_rhs; pq += nwarps) {  // Gram matrix Y^T Y
    const int a = pq / n_rhs, c = pq - a * n_rhs;
    double s = 0.0;
    for (int i = lane; i < P; i += 32) s += R[a * P + i] * R[c * P + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[pq] = s;
  }
  if (tid == 0) out[n_rhs * n_rhs] = 2.0 * logdet;
}

// mode 0: out[b] = -0.5 * (chi2 - logdet)   mode 1: out[b] = -0.5 * chi2   (likelihood.py:57-61, reference's sign)
// mode 2: out[b, :] = Gram (Fisher matrix)
__global__ void jc_slice_reduce_kernel(const double* __restrict__ partial, int64_t B, int L, int n_rhs, int mode,
                                       double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = mode == 2 ? n_rhs * n_rhs : 1;
  if (idx >= B * per) return;
  const int64_t b = idx / per;
  const int e = (int)(idx - b * per);
  const int stride = n_rhs * n_rhs + 1;
  double s = 0.0, logdet = 0.0;
  for (int l = 0; l < L; ++l) {
    s += partial[(b * L + l) * stride + e];
    logdet += partial[(b * L + l) * stride + n_rhs * n_rhs];
  }
  out[idx] = mode == 2 ? s : (mode == 0 ? -0.5 * (s - logdet) : -0.5 * s);
}

int launch(const double* cov, const double* rhs, int n_rhs, const double* sub, int64_t sub_stride, int64_t B, int P,
           int L, int mode, double* out, double* scratch, cudaStream_t s) {
  if (B > 65535 || n_rhs < 1 || n_rhs > JC_MAX_RHS) return JC_ERR_INVALID;
  const size_t smem = ((size_t)P * (P + 1) / 2 + (size_t)n_rhs * P) * sizeof(double);
  if (smem > 225 * 1024) return JC_ERR_UNSUPPORTED;  // slice does not fit one SM's shared memory
  JC_CUDA_TRY(cudaFuncSetAttribute(jc_slice_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = P >= 128 ? 512 : (P >= 32 ? 256 : 64);
  jc_slice_solve_kernel<<<dim3(L, (unsigned)B), threads, smem, s>>>(cov, rhs, n_rhs, sub, sub_stride, P, L, scratch);
  const int64_t n_out = B * (mode == 2 ? n_rhs * n_rhs : 1);
  jc_slice_reduce_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, s>>>(scratch, B, L, n_rhs, mode, out);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

}  // namespace

extern "C" int jc_gaussian_loglike_f64(const double* data_dev, int64_t data_stride, const double* mu_dev,
                                       const double* cov_dev, int64_t n_cosmo, int32_t P, int32_t L,
                                       int32_t include_logdet, double* loglike_dev, double* scratch_dev,
                                       void* stream) {
  if (!data_dev || !mu_dev || !cov_dev || !loglike_dev || !scratch_dev || n_cosmo < 1 || P < 1 || L < 1)
    return JC_ERR_INVALID;
  return launch(cov_dev, mu_dev, 1, data_dev, data_stride, n_cosmo, P, L, include_logdet ? 0 : 1, loglike_dev,
                scratch_dev, (cudaStream_t)stream);
}

// grad[b, k] = <jac[b, k, :], cot[b, :]>: one CTA per (k, b), 8 independent partial sums per thread, fixed
// shuffle / shared-memory reduction tree (deterministic)
__global__ void __launch_bounds__(256) jc_vjp_kernel(const double* __restrict__ jac, const double* __restrict__ cot,
                                                     int64_t cot_stride, int K, int64_t N, double* __restrict__ grad) {
  const int k = blockIdx.x;
  const int64_t b = blockIdx.y;
  const double* jr = jac + ((size_t)b * K + k) * N;
  const double* cr = cot + (size_t)b * cot_stride;
  double acc = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += 256) acc = fma(jr[n], cr[n], acc);
  __shared__ double part[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w];
    grad[(size_t)b * K + k] = s;
  }
}

extern "C" int jc_vjp_f64(const double* jac_dev, const double* cot_dev, int64_t cot_stride, int64_t n_cosmo,
                          int32_t n_params, int64_t N, double* grad_dev, void* stream) {
  if (!jac_dev || !cot_dev || !grad_dev || n_cosmo < 1 || n_cosmo > 65535 || n_params < 1 || N < 1 ||
      (cot_stride != 0 && cot_stride != N))
    return JC_ERR_INVALID;
  jc_vjp_kernel<<<dim3(n_params, (unsigned)n_cosmo), 256, 0, (cudaStream_t)stream>>>(jac_dev, cot_dev, cot_stride,
                                                                                    n_params, N, grad_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_fisher_f64(const double* jac_dev, const double* cov_dev, int64_t n_cosmo, int32_t n_params,
                             int32_t P, int32_t L, double* fisher_dev, double* scratch_dev, void* stream) {
  if (!jac_dev || !cov_dev || !fisher_dev || !scratch_dev || n_cosmo < 1 || P < 1 || L < 1) return JC_ERR_INVALID;
  return launch(cov_dev, jac_dev, n_params, nullptr, 0, n_cosmo, P, L, 2, fisher_dev, scratch_dev, (cudaStream_t)stream);
}
