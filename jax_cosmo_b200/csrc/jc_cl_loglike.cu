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
// One warp per (cosmology, ell) slice, matrices in registers (lane = row / column, shuffles for the rest); the CTA
// stages 8 consecutive ell of the [P, L] spectra so that global reads are 64-byte segments.  Optionally the same pass returns the cotangent d lnL / d cl[p, l] INCLUDING
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

// One warp per (cosmology, ell) slice, TM = T rounded up to 8 / 16 / 24 / 32, every loop fully unrolled.  Lane i keeps
// ROW i of C (Cholesky) and COLUMN i of the right-hand sides (substitutions) in registers; what a step needs from the
// other lanes -- column k of L -- is written once to a column-major shared-memory tile and read back by all lanes as
// broadcast LDS.128 (two entries per instruction).  History: matrices in shared memory with one dependent load-FMA-store
// per step: 66 k cycles per slice, 45 % of the config-3 step; registers + one 64-bit shuffle per L entry: shuffle bound
// (2 SHFL per entry, one warp-shuffle per clock and SM), 36 %.  All eliminations are right-looking / column oriented so
// that the T^2/2 FMAs of a phase are independent; rows and columns beyond T are identity / zero padding.
template <int TM, bool GRAD>
__global__ void __launch_bounds__(LT * 32) jc_cl_loglike_kernel(JcDevPlan pl, const double* __restrict__ cl,
                                                                const double* __restrict__ data, int64_t data_stride,
                                                                const double* __restrict__ noise, double f_sky,
                                                                double* __restrict__ partial /* [B, L, 2] */,
                                                                double* __restrict__ dcl /* [B, P, L] or null */) {
  extern __shared__ __align__(16) double sm[];
  const int T = pl.T, P = pl.P, L = pl.L;
  constexpr int TP = TM + 2;   // column stride of the L tile (even: LDS.128 pairs stay 16-byte aligned)
  constexpr int TQ = TM + 1;   // row stride of the transpose tile
  constexpr int PER_WARP = TM * TP + TM + TM * TQ;
  const int b = blockIdx.y, l0 = blockIdx.x * LT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* s_cl = sm;                    // [P][LT] signal
  double* s_r = s_cl + (size_t)P * LT;  // [P][LT] residual mu - data; later the cotangent tile
  double* Lc = s_r + (size_t)P * LT + (size_t)warp * PER_WARP;  // [TM][TP] column-major: Lc[k * TP + j] = L_jk (j > k)
  double* invd = Lc + TM * TP;                                  // [TM] 1 / L_kk
  double* St = invd + TM;                                       // [TM][TQ] transpose tile
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
  const bool rowl = i < TM;
  double g[GRAD ? TM : 1];  // (GRAD) column i of d lnL / dS at the end
  if (live) {
    const double nu = pl.covnorm[l] * f_sky;  // (2l+1) gradient(l) f_sky, angular_cl.py:139
    double x[TM];
    double logdet;
    {
      double c[TM];
#pragma unroll
      for (int j = 0; j < TM; ++j) {
        const bool in = i < T && j < T;
        const int p = in ? (i <= j ? pair_idx(i, j, T) : pair_idx(j, i, T)) : 0;
        const double sv = s_cl[p * LT + warp], rv = s_r[p * LT + warp];
        c[j] = in ? sv + (i == j ? noise[i] : 0.0) : (i == j ? 1.0 : 0.0);
        x[j] = in ? rv : 0.0;
      }
      // Cholesky C = L L^T, right-looking.  Step k: every lane publishes its (unscaled) entry of column k, all lanes read the
      // pivot and the entries below it; with t = C_ik / C_kk the update is C_ij -= t C_jk.  An indefinite C gives NaN, as the
      // reference's inverse would.  log det from the running product of the pivots.
      double prod = 1.0, ld = 0.0;
#pragma unroll
      for (int k = 0; k < TM; ++k) {
        if (rowl) Lc[k * TP + i] = c[k];
        __syncwarp();
        const double dkk = Lc[k * TP + k];
        const double inv = rsqrt(dkk);
        prod *= dkk;
        if ((k & 7) == 7) { ld += log(prod); prod = 1.0; }
        if (i == 0) invd[k] = inv;
        const double t = c[k] * (inv * inv);
        if (((k + 1) & 1) && k + 1 < TM) c[k + 1] = fma(-t, Lc[k * TP + k + 1], c[k + 1]);  // odd first row: one LDS.64
#pragma unroll
        for (int j = (k + 2) & ~1; j < TM; j += 2) {  // TM is even: aligned pairs to the end
          const double2 pr = *reinterpret_cast<const double2*>(Lc + k * TP + j);
          c[j] = fma(-t, pr.x, c[j]);
          c[j + 1] = fma(-t, pr.y, c[j + 1]);
        }
      }
      logdet = (T + 1) * (ld + log(prod)) + T * 0.6931471805599453 - P * log(nu);
    }
    // scale the stored columns: Lc[k][i] = L_ik = C_ik^(k) / L_kk (lane i owns row i of the tile)
    __syncwarp();
    if (rowl) {
#pragma unroll
      for (int k = 0; k < TM; ++k)
        if (k < i) Lc[k * TP + i] *= invd[k];
    }
    __syncwarp();
    // forward substitution L y = v (lane i: its own right-hand side), column oriented
    auto forward = [&](double (&v)[TM]) {
#pragma unroll
      for (int m = 0; m < TM; ++m) {
        v[m] *= invd[m];
        if (((m + 1) & 1) && m + 1 < TM) v[m + 1] = fma(-Lc[m * TP + m + 1], v[m], v[m + 1]);
#pragma unroll
        for (int r = (m + 2) & ~1; r < TM; r += 2) {
          const double2 pr = *reinterpret_cast<const double2*>(Lc + m * TP + r);
          v[r] = fma(-pr.x, v[m], v[r]);
          v[r + 1] = fma(-pr.y, v[m], v[r + 1]);
        }
      }
    };
    auto backward = [&](double (&v)[TM]) {  // L^T q = v: q_m final, then eliminated from the rows above with L_mr = Lc[r][m]
#pragma unroll
      for (int m = TM - 1; m >= 0; --m) {
        v[m] *= invd[m];
#pragma unroll
        for (int r = 0; r < m; ++r) v[r] = fma(-Lc[r * TP + m], v[m], v[r]);
      }
    };
    auto transpose = [&](double (&v)[TM]) {  // lane i: v[r] = M[r][i]  ->  v[r] = M[i][r]
      __syncwarp();
      if (rowl) {
#pragma unroll
        for (int r = 0; r < TM; ++r) St[r * TQ + i] = v[r];
      }
      __syncwarp();
      if (rowl) {
#pragma unroll
        for (int r = 0; r < TM; ++r) v[r] = St[i * TQ + r];
      }
    };
    forward(x);    // column i of Y = L^-1 R
    transpose(x);  // column i of Y^T
    forward(x);    // column i of X = L^-1 R L^-T (symmetric)
    double ss = 0.0;
#pragma unroll
    for (int r = 0; r < TM; ++r) ss = fma(x[r], x[r], ss);
    if (!rowl) ss = 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
      partial[((size_t)b * L + l) * 2 + 0] = 0.5 * nu * ss;
      partial[((size_t)b * L + l) * 2 + 1] = logdet;
    }
    if constexpr (GRAD) {
      // column i of M = -1/2 (nu X - nu X^2 - (T+1) I) inside the T x T block, 0 in the padding; X through the tile
      __syncwarp();
      if (rowl) {
#pragma unroll
        for (int r = 0; r < TM; ++r) St[r * TQ + i] = x[r];  // St[r][i] = X[r][i]
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        double x2 = 0.0;
#pragma unroll
        for (int m = 0; m < TM; ++m) x2 = fma(St[r * TQ + m], x[m], x2);  // sum_m X[r][m] X[m][i]
        g[r] = (i < T && r < T) ? -0.5 * (nu * (x[r] - x2) - (i == r ? (double)(T + 1) : 0.0)) : 0.0;
      }
      backward(g);   // column i of Q = L^-T M
      transpose(g);  // column i of Q^T
      backward(g);   // column i of dlnL/dS = L^-T M L^-1 (symmetric)
    }
  }
  if constexpr (GRAD) {
    __syncthreads();  // every warp has read its residual column of s_r: reuse the tile for the cotangent
    if (live && i < T) {
#pragma unroll
      for (int r = 0; r < TM; ++r)  // d lnL / d cl[(r,i)], r <= i: both S_ri and S_ir move with an off-diagonal spectrum
        if (r <= i && r < T) s_r[pair_idx(r, i, T) * LT + warp] = (r == i ? 1.0 : 2.0) * g[r];
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
  const int TM = (pl.T + 7) & ~7;
  const size_t smem = ((size_t)2 * pl.P * LT + (size_t)LT * (TM * (TM + 2) + TM + TM * (TM + 1))) * sizeof(double);
  if (smem > 200 * 1024) return JC_ERR_UNSUPPORTED;
  const dim3 grid((pl.L + LT - 1) / LT, (unsigned)n_cosmo);
#define JC_LL_LAUNCH(TM_, GRAD_)                                                                                              \
  do {                                                                                                                        \
    static unsigned long long done_ = 0;                                                                                      \
    JC_ONCE_PER_DEVICE(done_, cudaFuncSetAttribute(jc_cl_loglike_kernel<TM_, GRAD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                   200 * 1024));                                                              \
    jc_cl_loglike_kernel<TM_, GRAD_><<<grid, LT * 32, smem, s>>>(pl, cl_dev, data_dev, data_stride, noise_dev, f_sky, scratch_dev, \
                                                                 dcl_dev);                                                    \
  } while (0)
  if (dcl_dev) {
    switch (TM) { case 8: JC_LL_LAUNCH(8, true); break; case 16: JC_LL_LAUNCH(16, true); break;
                  case 24: JC_LL_LAUNCH(24, true); break; default: JC_LL_LAUNCH(32, true); break; }
  } else {
    switch (TM) { case 8: JC_LL_LAUNCH(8, false); break; case 16: JC_LL_LAUNCH(16, false); break;
                  case 24: JC_LL_LAUNCH(24, false); break; default: JC_LL_LAUNCH(32, false); break; }
  }
#undef JC_LL_LAUNCH
  jc_cl_loglike_sum_kernel<<<(unsigned)((n_cosmo + 127) / 128), 128, 0, s>>>(scratch_dev, n_cosmo, pl.L, include_logdet, loglike_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}
