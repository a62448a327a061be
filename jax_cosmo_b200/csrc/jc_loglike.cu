// jc_loglike.cu -- Gaussian log-likelihood on the sparse block covariance layout [P, P, L]
// (replaces likelihood.py:9-61 with sparse.inv / sparse.slogdet, sparse.py:295-366; BASELINE config 3).
//
// The covariance is "a diagonal of blocks": for each ell the [P x P] slice C_l = cov[:, :, l] couples the
// P spectra, different ell are independent.  Hence
//     r^T C^-1 r = sum_l r_l^T C_l^-1 r_l ,     log det C = sum_l log det C_l ,
// which is what the reference's sparse.inv (batched inverse over ell) and Schur-recursion slogdet compute.
// One CTA per (cosmology, ell): the slice is gathered into shared memory as a packed lower triangle
// (P(P+1)/2 doubles: 177 KB at P = 210), factorised in place (right-looking Cholesky; Gaussian C_l
// covariances are SPD), r_l is forward-substituted, and chi2_l = |L^-1 r_l|^2, logdet_l = 2 sum log L_kk.
// A second tiny kernel sums the per-ell partials in a fixed order (deterministic).
#include "jc_internal.cuh"

namespace {

__device__ __forceinline__ int tri(int i, int j) { return (i * (i + 1) >> 1) + j; }  // i >= j

__global__ void __launch_bounds__(512) jc_loglike_slice_kernel(const double* __restrict__ cov,
                                                               const double* __restrict__ mu,
                                                               const double* __restrict__ data, int64_t data_stride,
                                                               int P, int L, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* A = sm;                            // packed lower triangle
  double* r = sm + ((size_t)P * (P + 1) >> 1);  // residual / solution
  __shared__ double s_piv;
  __shared__ double s_red[32];
  const int l = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const double* C = cov + (size_t)b * P * P * L + l;
  for (int i = warp; i < P; i += nwarps)
    for (int j = lane; j <= i; j += 32) A[tri(i, j)] = C[((size_t)i * P + j) * L];
  for (int i = tid; i < P; i += nthr)
    r[i] = mu[(size_t)b * P * L + (size_t)i * L + l] - data[(size_t)b * data_stride + (size_t)i * L + l];  // r = mu - data
  __syncthreads();

  double logdet = 0.0;  // accumulated by thread 0
  for (int k = 0; k < P; ++k) {
    if (tid == 0) {
      const double d = sqrt(A[tri(k, k)]);
      A[tri(k, k)] = d;
      s_piv = 1.0 / d;
      logdet += log(d);
      r[k] *= s_piv;  // y_k of the forward substitution (all updates of r[k] from columns < k are done)
    }
    __syncthreads();
    const double inv = s_piv, yk = r[k];
    for (int i = k + 1 + tid; i < P; i += nthr) {
      const double lik = A[tri(i, k)] * inv;
      A[tri(i, k)] = lik;
      r[i] -= lik * yk;
    }
    __syncthreads();
    for (int i = k + 1 + warp; i < P; i += nwarps) {  // trailing update of the lower triangle
      const double lik = A[tri(i, k)];
      double* row = A + tri(i, 0);
      for (int j = k + 1 + lane; j <= i; j += 32) row[j] -= lik * A[tri(j, k)];
    }
    __syncthreads();
  }
  double chi2 = 0.0;
  for (int i = tid; i < P; i += nthr) chi2 += r[i] * r[i];
  for (int o = 16; o > 0; o >>= 1) chi2 += __shfl_xor_sync(0xffffffffu, chi2, o);
  if (lane == 0) s_red[warp] = chi2;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += s_red[w];
    partial[((size_t)b * L + l) * 2 + 0] = s;
    partial[((size_t)b * L + l) * 2 + 1] = 2.0 * logdet;
  }
}

__global__ void jc_loglike_reduce_kernel(const double* __restrict__ partial, int64_t B, int L, int include_logdet,
                                         double* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double chi2 = 0.0, logdet = 0.0;
  for (int l = 0; l < L; ++l) {
    chi2 += partial[(b * L + l) * 2];
    logdet += partial[(b * L + l) * 2 + 1];
  }
  out[b] = include_logdet ? -0.5 * (chi2 - logdet) : -0.5 * chi2;  // likelihood.py:57-61 (sign as in the reference)
}

}  // namespace

extern "C" int jc_gaussian_loglike_f64(const double* data_dev, int64_t data_stride, const double* mu_dev,
                                       const double* cov_dev, int64_t n_cosmo, int32_t P, int32_t L,
                                       int32_t include_logdet, double* loglike_dev, double* scratch_dev,
                                       void* stream) {
  if (!data_dev || !mu_dev || !cov_dev || !loglike_dev || !scratch_dev || n_cosmo < 1 || P < 1 || L < 1)
    return JC_ERR_INVALID;
  if (n_cosmo > 65535) return JC_ERR_INVALID;
  const size_t smem = ((size_t)P * (P + 1) / 2 + P) * sizeof(double);
  if (smem > 225 * 1024) return JC_ERR_UNSUPPORTED;  // slice does not fit one SM's shared memory (P > 238)
  cudaStream_t s = (cudaStream_t)stream;
  JC_CUDA_TRY(cudaFuncSetAttribute(jc_loglike_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = P >= 128 ? 512 : (P >= 32 ? 256 : 64);
  jc_loglike_slice_kernel<<<dim3(L, (unsigned)n_cosmo), threads, smem, s>>>(cov_dev, mu_dev, data_dev, data_stride, P, L,
                                                                           scratch_dev);
  jc_loglike_reduce_kernel<<<(unsigned)((n_cosmo + 127) / 128), 128, 0, s>>>(scratch_dev, n_cosmo, L, include_logdet,
                                                                            loglike_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}
