// jc_sparse.cu -- stand-alone linear algebra on the jax_cosmo.sparse block layout (sparse.py): a matrix of
// [ny, nx] diagonal blocks of size n is stored as S[ny, nx, n]; every product and the inverse decouple into n
// independent small dense problems, one per diagonal position l (= one per ell for the covariance).
//
//   jc_sparse_bmm_f64   C[i,k,l] = sum_j A[i,j,l] B[j,k,l] with arbitrary element strides: covers sparse.dot's
//                       sparse @ vec / dense / sparse, vec / dense @ sparse and the bilinear form (sparse.py:72-292)
//   jc_sparse_inv_f64   per-l inverse and log-determinant by Gauss-Jordan with partial pivoting, the algorithm
//                       class of np.linalg.inv / slogdet that the reference calls per slice (sparse.py:296-389)
//
// The fused consumers (likelihood, Fisher, gradient: jc_loglike.cu) stay the fast path for SPD covariances; these
// kernels complete the module for general (non-symmetric, rectangular in bmm) inputs.
#include "jc_internal.cuh"

namespace {

// one thread per output element (l fastest: coalesced for unit l-stride), fixed summation order over j
__global__ void __launch_bounds__(256) jc_sparse_bmm_kernel(const double* __restrict__ A, int64_t sAi, int64_t sAj, int64_t sAl,
                                                            const double* __restrict__ B, int64_t sBj, int64_t sBk, int64_t sBl,
                                                            double* __restrict__ C, int64_t sCi, int64_t sCk, int64_t sCl,
                                                            int I, int J, int K, int L) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)I * K * L) return;
  const int l = (int)(idx % L);
  const int64_t r = idx / L;
  const int k = (int)(r % K), i = (int)(r / K);
  const double* a = A + i * sAi + l * sAl;
  const double* b = B + k * sBk + l * sBl;
  double acc = 0.0;
  for (int j = 0; j < J; ++j) acc = fma(a[j * sAj], b[j * sBj], acc);
  C[i * sCi + k * sCk + l * sCl] = acc;
}

// One CTA per diagonal position l.  W = [A_l | I] (P x 2P, row-major) lives in the caller's scratch (L2 resident:
// 0.7 MB per slice at P = 210); per pivot column: block arg-max of |W[r][col]| over r >= col (ties -> lowest row,
// as LAPACK's idamax), row swap, scale, rank-1 elimination of every other row.
__global__ void __launch_bounds__(512) jc_sparse_inv_kernel(const double* __restrict__ S, int P, int L,
                                                            double* __restrict__ inv, double* __restrict__ sign_out,
                                                            double* __restrict__ logdet_out, double* __restrict__ scratch) {
  const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int W2 = 2 * P;
  double* W = scratch + (size_t)l * P * W2;
  __shared__ double s_val[16];
  __shared__ int s_row[16];
  __shared__ int s_piv;
  __shared__ double s_pivval;
  for (int e = tid; e < P * W2; e += nt) {
    const int r = e / W2, c = e - r * W2;
    W[e] = c < P ? S[((size_t)r * P + c) * L + l] : (c - P == r ? 1.0 : 0.0);
  }
  __syncthreads();
  double sign = 1.0, logdet = 0.0;
  for (int col = 0; col < P; ++col) {
    // ---- pivot search ----
    double best = -1.0;
    int brow = P;
    for (int r = col + tid; r < P; r += nt) {
      const double v = fabs(W[(size_t)r * W2 + col]);
      if (v > best) { best = v; brow = r; }  // rows visited in increasing order per thread: first maximum
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
      if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_row[tid >> 5] = brow; }
    __syncthreads();
    if (tid == 0) {
      double b = s_val[0];
      int br = s_row[0];
      for (int w = 1; w < (nt >> 5); ++w)
        if (s_val[w] > b || (s_val[w] == b && s_row[w] < br)) { b = s_val[w]; br = s_row[w]; }
      s_piv = br < P ? br : col;
      s_pivval = W[(size_t)s_piv * W2 + col];
    }
    __syncthreads();
    const int piv = s_piv;
    const double pv = s_pivval;
    if (piv != col) sign = -sign;
    if (pv < 0.0) sign = -sign;
    if (pv == 0.0) sign = 0.0;
    logdet += log(fabs(pv));
    // ---- swap rows col <-> piv and scale the pivot row ----
    const double ipv = 1.0 / pv;
    for (int c = tid; c < W2; c += nt) {
      const double a = W[(size_t)piv * W2 + c], b = W[(size_t)col * W2 + c];
      W[(size_t)piv * W2 + c] = b;
      W[(size_t)col * W2 + c] = a * ipv;
    }
    __syncthreads();
    // ---- eliminate column col from every other row (columns <= col of the A part are already final) ----
    const int c0 = col + 1, nc = W2 - c0;
    for (int e = tid; e < P * nc; e += nt) {
      const int r = e / nc, c = c0 + (e - r * nc);
      if (r == col) continue;
      const double f = W[(size_t)r * W2 + col];
      W[(size_t)r * W2 + c] = fma(-f, W[(size_t)col * W2 + c], W[(size_t)r * W2 + c]);
    }
    __syncthreads();
  }
  if (inv)
    for (int e = tid; e < P * P; e += nt) {
      const int r = e / P, c = e - r * P;
      inv[((size_t)r * P + c) * L + l] = W[(size_t)r * W2 + P + c];
    }
  if (tid == 0) {
    if (sign_out) sign_out[l] = sign;
    if (logdet_out) logdet_out[l] = logdet;
  }
}

}  // namespace

extern "C" int jc_sparse_bmm_f64(const double* A, int64_t sAi, int64_t sAj, int64_t sAl, const double* B, int64_t sBj,
                                 int64_t sBk, int64_t sBl, double* C, int64_t sCi, int64_t sCk, int64_t sCl, int32_t I,
                                 int32_t J, int32_t K, int32_t L, void* stream) {
  if (!A || !B || !C || I < 1 || J < 1 || K < 1 || L < 1) return JC_ERR_INVALID;
  const int64_t total = (int64_t)I * K * L;
  const int64_t blocks = (total + 255) / 256;
  if (blocks > 0x7fffffff) return JC_ERR_INVALID;
  jc_sparse_bmm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, sAi, sAj, sAl, B, sBj, sBk, sBl, C, sCi,
                                                                           sCk, sCl, I, J, K, L);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}

extern "C" int jc_sparse_inv_f64(const double* sparse_dev, int32_t P, int32_t L, double* inv_dev, double* sign_dev,
                                 double* logdet_dev, double* scratch_dev, void* stream) {
  if (!sparse_dev || !scratch_dev || P < 1 || L < 1 || (!inv_dev && !sign_dev && !logdet_dev)) return JC_ERR_INVALID;
  jc_sparse_inv_kernel<<<L, 512, 0, (cudaStream_t)stream>>>(sparse_dev, P, L, inv_dev, sign_dev, logdet_dev, scratch_dev);
  JC_CUDA_TRY(cudaGetLastError());
  return JC_OK;
}
