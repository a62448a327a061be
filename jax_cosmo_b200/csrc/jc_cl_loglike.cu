// jc_cl_loglike.cu -- Gaussian log-likelihood of a data vector of C_ell under the Gaussian covariance of the model
// spectra, WITHOUT forming the covariance (BASELINE config 3 end to end on the device):
//
//   reference:  mu, cov = gaussian_cl_covariance_and_mean(cosmo, ell, probes, sparse=True)      angular_cl.py:120-196
//               lnL = gaussian_log_likelihood(data, mu, cov)                                     likelihood.py:9-61
//
// The reference builds cov[(ij),(mn),l] = (C_im C_jn + C_in C_jm) / nu_l (35 MB per cosmology at 10+10 bins), inverts 100
// [210 x 210] slices and runs a Schur-recursion slogdet.  That covariance is the operator S -> C S C on symmetric T x T
// matrices written in the basis of the P = T(T+1)/2 unique pairs: with G = diag(1 for i<j, 1/2 for i=j)
//     nu Cov = A G^-1,   A x = vech(C S C)   =>   r^T Cov^-1 r = nu/2 tr(C^-1 R C^-1 R) = nu/2 || L^-1 R L^-T ||_F^2
//     log det Cov = (T+1) log det C + T log 2 - P log nu            (det A = det(C)^(T+1), det G = 2^-T)
// with C = signal + noise (T x T, per ell), R the symmetric matrix of residuals mu - data and C = L L^T.  A slice costs
// O(T^3) = 8e3 flop instead of O(P^3) = 3e6 and reads 8 P bytes instead of 8 P^2: the likelihood of a cosmology
// becomes a 2 % epilogue of its spectra.  Identical in exact arithmetic to the reference's expression; tests hold it to
// the reference-generated golden (tests/golden/likelihood_*.npz) and to the explicit-covariance kernel (jc_loglike.cu).
//
// One warp per (cosmology, ell) slice, lane = matrix row; the CTA stages 8 consecutive ell of the [P, L] spectra so
// that global reads are 64-byte segments.  Optionally the same pass returns the cotangent d lnL / d cl[p, l] INCLUDING
// the dependence of the covariance and of its determinant on the spectra,
//     dlnL/dS = L^-T [ -1/2 (nu X - nu X^2 - (T+1) I) ] L^-1,   X = L^-1 R L^-T,
// which jc_vjp_f64 contracts with the forward-mode Jacobian: the gradient of the full likelihood that
// jax.grad(likelihood) returns in the reference's README (README.md:17-27).
#include "jc_internal.cuh"

namespace {

constexpr int LT = 8;  // ell per CTA (one warp each)

__device__ __forceinline__ int pair_idx(int i, int j, int T) {  // angular_cl.py:34-38, i <= j
  return i * T - (i * (i - 1)) / 2 + (j - i);
}

template <bool GRAD>
__global__ void __launch_bounds__(LT * 32) jc_cl_loglike_kernel(JcDevPlan pl, const double* __restrict__ cl,
                                                                const double* __restrict__ data, int64_t data_stride,
                                                                const double* __restrict__ noise, double f_sky,
                                                                double* __restrict__ partial /* [B, L, 2] */,
                                                                double* __restrict__ dcl /* [B, P, L] or null */) {
  extern __shared__ __align__(16) double sm[];
  const int T = pl.T, P = pl.P, L = pl.L, TP = T + 1;
  const int b = blockIdx.y, l0 = blockIdx.x * LT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* s_cl = sm;                       // [P][LT] signal
  double* s_r = s_cl + (size_t)P * LT;     // [P][LT] residual mu - data; later the cotangent tile
  double* mats = s_r + (size_t)P * LT;
  const int nmat = GRAD ? 3 : 2;
  double* Cm = mats + (size_t)warp * nmat * T * TP;  // [T][T+1] C -> L (lower triangle)
  double* Xm = Cm + (size_t)T * TP;                  // [T][T+1] R -> X
  double* Mm = Xm + (size_t)T * TP;                  // [T][T+1] (GRAD) M -> dlnL/dS
  const double* clb = cl + (size_t)b * P * L;
  const double* db = data + (size_t)b * data_stride;
  for (int q = threadIdx.x; q < P * LT; q += LT * 32) {
    const int p = q / LT, w = q - p * LT, l = l0 + w;
    double s = 0.0, d = 0.0;
    if (l < L) { s = clb[(size_t)p * L + l]; d = db[(size_t)p * L + l]; }
    s_cl[q] = s;
    s_r[q] = s - d;
  }
  __syncthreads();
  const int l = l0 + warp;
  const bool live = l < L;  // warps past the last ell idle through the CTA barriers below
  const int i = lane;
  const bool row = live && i < T;
  double chi2 = 0.0, logdet = 0.0, nu = 1.0;
  if (live) {
    nu = pl.covnorm[l] * f_sky;  // (2l+1) gradient(l) f_sky, angular_cl.py:139
    if (row) {
      for (int j = 0; j < T; ++j) {
        const int p = i <= j ? pair_idx(i, j, T) : pair_idx(j, i, T);
        Cm[i * TP + j] = s_cl[p * LT + warp] + (i == j ? noise[i] : 0.0);
        Xm[i * TP + j] = s_r[p * LT + warp];
      }
    }
    __syncwarp();
    // Cholesky C = L L^T (right-looking, lower triangle in place); an indefinite C gives NaN, as the reference's inverse would.
    // The slice is latency bound (one warp, dependent steps): no division, square root or logarithm inside the loops --
    // 1/L_kk by rsqrt is kept on the diagonal (the substitutions multiply by it), log det from a running product.
    double ld = 0.0, prod = 1.0;
    for (int k = 0; k < T; ++k) {
      const double ckk = Cm[k * TP + k];
      const double inv = rsqrt(ckk);  // 1 / L_kk
      prod *= ckk;                    // = L_kk^2
      if ((k & 7) == 7) { ld += log(prod); prod = 1.0; }
      __syncwarp();
      if (row && i > k) Cm[i * TP + k] *= inv;
      if (i == k) Cm[k * TP + k] = inv;
      __syncwarp();
      if (row && i > k) {
        const double lik = Cm[i * TP + k];
        for (int j = k + 1; j <= i; ++j) Cm[i * TP + j] -= lik * Cm[j * TP + k];
      }
      __syncwarp();
    }
    ld += log(prod);  // sum_k log L_kk^2
    logdet = (T + 1) * ld + T * 0.6931471805599453 - P * log(nu);
    // Y^T = R L^-T: lane j forward-substitutes column j of R (= its own row, R is symmetric) in place
    if (row) {
      for (int r = 0; r < T; ++r) {
        double s0 = Xm[i * TP + r], s1 = 0.0;
        int m = 0;
        for (; m + 1 < r; m += 2) {
          s0 -= Cm[r * TP + m] * Xm[i * TP + m];
          s1 -= Cm[r * TP + m + 1] * Xm[i * TP + m + 1];
        }
        if (m < r) s0 -= Cm[r * TP + m] * Xm[i * TP + m];
        Xm[i * TP + r] = (s0 + s1) * Cm[r * TP + r];
      }
    }
    __syncwarp();
    // X = L^-1 Y^T: lane c forward-substitutes column c in place; X is symmetric
    double ss = 0.0;
    if (row) {
      for (int r = 0; r < T; ++r) {
        double s0 = Xm[r * TP + i], s1 = 0.0;
        int m = 0;
        for (; m + 1 < r; m += 2) {
          s0 -= Cm[r * TP + m] * Xm[m * TP + i];
          s1 -= Cm[r * TP + m + 1] * Xm[(m + 1) * TP + i];
        }
        if (m < r) s0 -= Cm[r * TP + m] * Xm[m * TP + i];
        const double x = (s0 + s1) * Cm[r * TP + r];
        Xm[r * TP + i] = x;
        ss += x * x;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    chi2 = 0.5 * nu * ss;
    __syncwarp();
    if (GRAD) {
      // M = -1/2 (nu X - nu X^2 - (T+1) I); lane i forms row i
      if (row) {
        for (int j = 0; j < T; ++j) {
          double x2 = 0.0;
          for (int m = 0; m < T; ++m) x2 += Xm[i * TP + m] * Xm[m * TP + j];
          Mm[i * TP + j] = -0.5 * (nu * (Xm[i * TP + j] - x2) - (i == j ? (double)(T + 1) : 0.0));
        }
      }
      __syncwarp();
      // Q = L^-T M: lane c back-substitutes column c in place
      if (row) {
        for (int r = T - 1; r >= 0; --r) {
          double s = Mm[r * TP + i];
          for (int m = r + 1; m < T; ++m) s -= Cm[m * TP + r] * Mm[m * TP + i];
          Mm[r * TP + i] = s * Cm[r * TP + r];  // the diagonal holds 1 / L_rr
        }
      }
      __syncwarp();
      // G = Q L^-1 (symmetric) = L^-T Q^T: lane c back-substitutes row c of Q in place
      if (row) {
        for (int r = T - 1; r >= 0; --r) {
          double s = Mm[i * TP + r];
          for (int m = r + 1; m < T; ++m) s -= Cm[m * TP + r] * Mm[i * TP + m];
          Mm[i * TP + r] = s * Cm[r * TP + r];
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      partial[((size_t)b * L + l) * 2 + 0] = chi2;
      partial[((size_t)b * L + l) * 2 + 1] = logdet;
    }
  }
  if (GRAD) {
    __syncthreads();  // every warp has read its residual column of s_r: reuse the tile for the cotangent
    if (row) {
      for (int j = i; j < T; ++j)  // d lnL / d cl[(i,j)] : both S_ij and S_ji move with an off-diagonal spectrum
        s_r[pair_idx(i, j, T) * LT + warp] = (i == j ? 1.0 : 2.0) * Mm[i * TP + j];
    }
    __syncthreads();
    double* ob = dcl + (size_t)b * P * L;
    for (int q = threadIdx.x; q < P * LT; q += LT * 32) {
      const int p = q / LT, w = q - p * LT;
      if (l0 + w < L) ob[(size_t)p * L + l0 + w] = s_r[q];
    }
  }
}

// loglike[b] = -1/2 (sum_l chi2 - sum_l logdet), fixed summation order
__global__ void jc_cl_loglike_sum_kernel(const double* __restrict__ partial, int64_t B, int L, int include_logdet,
                                         double* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double c = 0.0, d = 0.0;
  for (int l = 0; l < L; ++l) {
    c += partial[((size_t)b * L + l) * 2 + 0];
    d += partial[((size_t)b * L + l) * 2 + 1];
  }
  out[b] = include_logdet ? -0.5 * (c - d) : -0.5 * c;
}

}  // namespace

extern "C" int jc_gaussian_cl_loglike_f64(const jc_plan* plan, const double* cl_dev, const double* data_dev,
                                          int64_t data_stride, const double* noise_dev, int64_t n_cosmo, double f_sky,
                                          int32_t include_logdet, double* loglike_dev, double* dcl_dev,
                                          double* scratch_dev, void* stream) {
  if (!plan || plan->d.grid_mode || !cl_dev || !data_dev || !noise_dev || !loglike_dev || !scratch_dev || n_cosmo < 1)
    return JC_ERR_INVALID;
  const JcDevPlan& pl = plan->d;
  if (pl.L < 2 || pl.T > 32 || !(f_sky > 0.0)) return JC_ERR_INVALID;  // np.gradient needs >= 2 ell (angular_cl.py:139)
  if (data_stride != 0 && data_stride != (int64_t)pl.P * pl.L) return JC_ERR_INVALID;
  if (dcl_dev && !include_logdet) return JC_ERR_UNSUPPORTED;  // the cotangent is that of the full likelihood
  if (n_cosmo > 65535) return JC_ERR_INVALID;                 // grid.y; callers batch (35 us per 1000 cosmologies)
  JcDeviceGuard guard(plan->device);
  JC_CUDA_TRY(guard.status);
  cudaStream_t s = (cudaStream_t)stream;
  const int nmat = dcl_dev ? 3 : 2;
  const size_t smem = ((size_t)2 * pl.P * LT + (size_t)LT * nmat * pl.T * (pl.T + 1)) * sizeof(double);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, {
    cudaFuncSetAttribute(jc_cl_loglike_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(jc_cl_loglike_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (smem > 200 * 1024) return JC_ERR_UNSUPPORTED;
  const dim3 grid((pl.L + LT - 1) / LT, (unsigned)n_cosmo);
  if (dcl_dev)
    jc_cl_loglike_kernel<true><<<grid, LT * 32, smem, s>>>(pl, cl_dev, data_dev, data_stride, noise_dev, f_sky, scratch_dev, dcl_dev);
  else
    jc_cl_loglike_kernel<false><<<grid, LT * 32, smem, s>>>(pl, cl_dev, data_dev, data_stride, noise_dev, f_sky, scratch_dev, nullptr);
  jc_cl_loglike_sum_kernel<<<(unsigned)((n_cosmo + 127) / 128), 128, 0, s>>>(scratch_dev, n_cosmo, pl.L, include_logdet, loglike_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}
