// jc_tracers.cu -- K2: radial tracer kernels R_i(a_n), node-major [n][TS].
//
// K2a  lensing efficiency  q_s(z_n) = int_{z_n}^{zmax} n_s(z') max(chi'-chi_n,0)/max(chi',1) dz'
//      (probes.py:44-51).  The z' grid (linspace(z_n, zmax, 257) per node), its chi-table brackets and
//      the normalised n_s(z') are cosmology independent (plan tables, L2 resident); per cosmology only
//      the 256-entry chi table differs.  CTA = 128 nodes x 16 cosmologies: warp = (32 nodes, 4
//      cosmologies); the 4 warps that share a node block read the same table lines (L1 hits), so L2
//      traffic per cosmology is 1/16 of the table set.  g is computed once for all sources.
// K2b  finish: WL = (q (1+z) chi 3 H0^2 Om/(2c) + NLA) (1+m)   probes.py:51,71-74,102-129,201-207
//              NC = n_i(z) b_i(z) H(a)                           probes.py:77-99
#include "jc_internal.cuh"

namespace {

constexpr int LENS_NCOS = 4;     // cosmologies per thread
constexpr int LENS_CGROUPS = 4;  // cosmology groups per CTA
constexpr int LENS_CTA_COSMO = LENS_NCOS * LENS_CGROUPS;
constexpr int LENS_NODES = 128;  // nodes per CTA

template <int NS>
__global__ void __launch_bounds__(LENS_NODES * LENS_CGROUPS)
jc_lens_kernel(JcDevPlan pl, Ws ws, int n_cosmo, int s0) {
  __shared__ double s_chit[LENS_CTA_COSMO][JC_NCHI];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * LENS_NODES + (warp & 3) * 32 + lane;  // node
  const int cg = warp >> 2;                                         // cosmology group in CTA
  const int cta_c0 = blockIdx.y * LENS_CTA_COSMO;
  for (int i = threadIdx.x; i < LENS_CTA_COSMO * JC_NCHI; i += blockDim.x) {
    int cc = min(cta_c0 + i / JC_NCHI, n_cosmo - 1);
    s_chit[i / JC_NCHI][i % JC_NCHI] = ws.chitab[(size_t)cc * JC_NCHI + (i % JC_NCHI)];
  }
  const int c0 = cta_c0 + cg * LENS_NCOS;
  double chin[LENS_NCOS];
#pragma unroll
  for (int c = 0; c < LENS_NCOS; ++c) chin[c] = node_ptr(ws, min(c0 + c, n_cosmo - 1), JC_NODE_CHI)[n];
  __syncthreads();
  const double(*chit)[JC_NCHI] = s_chit + cg * LENS_NCOS;
  const size_t NL = (size_t)JC_NLENS * JC_NLENS_COLS;
  // sources s0..s0+NS-1 are consecutive [257][512] slabs (the launcher guarantees s0+NS <= n_src)
  const double* nw0 = pl.lens_nw + (size_t)s0 * NL + n;
  double acc[LENS_NCOS][NS];
#pragma unroll
  for (int c = 0; c < LENS_NCOS; ++c)
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[c][s] = 0.0;

#pragma unroll 2
  for (int m = 0; m < JC_NLENS; ++m) {
    const size_t o = (size_t)m * JC_NLENS_COLS;
    const double t = __ldg(pl.lens_t + o + n);
    const int ix = __ldg(pl.lens_ix + o + n);
    const int i0 = ix & 255, i1 = ix >> 8;
    double wv[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) wv[s] = __ldg(nw0 + o + s * NL);
#pragma unroll
    for (int c = 0; c < LENS_NCOS; ++c) {
      const double f0 = chit[c][i0], f1 = chit[c][i1];
      const double chip = fmax(fma(f1 - f0, t, f0), 0.0);                 // background.py:242
      const double g = fmax(chip - chin[c], 0.0) * jc_rcp(fmax(chip, 1.0));  // probes.py:49
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[c][s] = fma(wv[s], g, acc[c][s]);
    }
  }
  const double dz = pl.lens_zmax - pl.limb_z[n];  // simps: dx * N (probes.py:51)
#pragma unroll
  for (int c = 0; c < LENS_NCOS; ++c) {
    if (c0 + c >= n_cosmo) break;
    double* row = ws.rker + ((size_t)(c0 + c) * JC_NA_PAD + n) * pl.TS;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      row[pl.src_tracer[s0 + s]] = acc[c][s] * dz;
    }
  }
}

__global__ void __launch_bounds__(256) jc_tracer_finish_kernel(JcDevPlan pl, Ws ws) {
  const int c = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= pl.T * JC_NA) return;
  const int n = idx / pl.T, t = idx - n * pl.T;  // tracer fastest: contiguous writes of R[n][:]
  const double Om = ws.scal[(size_t)c * JC_SCAL_FIELDS + JC_SCAL_OMEGA_M];
  const double H = node_ptr(ws, c, JC_NODE_HUBBLE)[n];
  const double D = node_ptr(ws, c, JC_NODE_GROWTH)[n];
  double* out = ws.rker + ((size_t)c * JC_NA_PAD + n) * pl.TS + t;
  const double nz = pl.nz_node[(size_t)n * pl.TS + t];
  double b = pl.bias_node[(size_t)n * pl.TS + t];
  if (pl.tr_inv_growth[t]) b = b / D;  // bias.py:37-39
  double r;
  if (pl.tr_kind[t] == JC_TRACER_WEAK_LENSING) {
    const double chi = node_ptr(ws, c, JC_NODE_CHI)[n];
    const double q = (n < JC_NLENS_COLS) ? *out : 0.0;  // node 512 is a=1: chi=0, kernel = 0
    r = q * (1.0 + pl.limb_z[n]) * chi * (3.0 * JC_H0 * JC_H0 * Om / 2.0 / JC_C_LIGHT);
    if (pl.tr_ia[t]) r += nz * b * H * (-(JC_C1_RHOCRIT)*Om / D);  // probes.py:119-123
    r *= pl.tr_m1[t];
  } else {
    r = nz * b * H;
  }
  *out = r;
}

template <int NS>
void launch_lens(const JcDevPlan& pl, const Ws& ws, int chunk, int s0, cudaStream_t st) {
  dim3 grid(JC_NLENS_COLS / LENS_NODES, (chunk + LENS_CTA_COSMO - 1) / LENS_CTA_COSMO);
  jc_lens_kernel<NS><<<grid, LENS_NODES * LENS_CGROUPS, 0, st>>>(pl, ws, chunk, s0);
}

}  // namespace

int jc_launch_tracers(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  int n_launch = 0;
  for (int s0 = 0; s0 < pl.n_src; ++n_launch) {
    const int rem = pl.n_src - s0;
    if (rem >= 10) { launch_lens<10>(pl, ws, chunk, s0, s); s0 += 10; }
    else if (rem >= 8) { launch_lens<8>(pl, ws, chunk, s0, s); s0 += 8; }
    else if (rem >= 6) { launch_lens<6>(pl, ws, chunk, s0, s); s0 += 6; }
    else if (rem >= 5) { launch_lens<5>(pl, ws, chunk, s0, s); s0 += 5; }
    else if (rem == 4) { launch_lens<4>(pl, ws, chunk, s0, s); s0 += 4; }
    else if (rem == 3) { launch_lens<3>(pl, ws, chunk, s0, s); s0 += 3; }
    else if (rem == 2) { launch_lens<2>(pl, ws, chunk, s0, s); s0 += 2; }
    else { launch_lens<1>(pl, ws, chunk, s0, s); s0 += 1; }
  }
  return n_launch;
}

void jc_launch_finish(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  jc_tracer_finish_kernel<<<dim3((pl.T * JC_NA + 255) / 256, chunk), 256, 0, s>>>(pl, ws);
}
