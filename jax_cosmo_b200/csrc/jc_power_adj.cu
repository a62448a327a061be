// jc_power_adj.cu -- K3 of the forward-mode path for several directions at once: V and its NT directional derivatives from ONE
// reverse sweep of the point function (jc_power_point.cuh) instead of NT forward-mode tangents.
//
// ncu on the DualN<4> / DualN<3> passes of the exact power kernel (profiles/r02_ncu_summary.md, section 4): 49 % of a 7-parameter
// Jacobian batch, 71 % of the FP64 pipe at ~225 FP64 instructions per point and direction -- arithmetic, not latency.  The
// point function has 28 inputs (18 per-node fields, (l+1/2)^(3+n_s), 9 Eisenstein-Hu constants) whose tangents K1 / K2 have
// already tabulated for every direction; the reverse sweep costs ~2 x the value and each direction one multiply-add per input.
//
// Tangent planes: direction k of a table lives at offset jvp_plane(k) * ws.doff from its value -- the planes the grouped K1 / K2
// passes wrote (group j = k / 4 occupies planes 5 j .. 5 j + 4, its plane 5 j being a second copy of the values).
#include "jc_internal.cuh"
#include "jc_math.cuh"
#define JC_B200_H_FIELDS 1
enum { JCP_NODE_INVCHIC = JC_NODE_INVCHIC, JCP_NODE_LNCHIC = JC_NODE_LNCHIC, JCP_NODE_RNL = JC_NODE_RNL, JCP_NODE_LNKNL = JC_NODE_LNKNL,
       JCP_NODE_AN = JC_NODE_AN, JCP_NODE_BN = JC_NODE_BN, JCP_NODE_LNCF = JC_NODE_LNCF, JCP_NODE_P3 = JC_NODE_P3,
       JCP_NODE_ALPHA = JC_NODE_ALPHA, JCP_NODE_BETA = JC_NODE_BETA, JCP_NODE_NU = JC_NODE_NU, JCP_NODE_E1 = JC_NODE_E1,
       JCP_NODE_E2 = JC_NODE_E2, JCP_NODE_NQ108 = JC_NODE_NQ108, JCP_NODE_NSILK = JC_NODE_NSILK, JCP_NODE_NAMP = JC_NODE_NAMP,
       JCP_NODE_GK = JC_NODE_GK, JCP_NODE_MU = JC_NODE_MU };
enum { JCP_SCAL_INV13KEQ = JC_SCAL_INV13KEQ, JCP_SCAL_BETA_C = JC_SCAL_BETA_C, JCP_SCAL_C14_ALPHA_C = JC_SCAL_C14_ALPHA_C,
       JCP_SCAL_SH_D = JC_SCAL_SH_D, JCP_SCAL_ALPHA_B = JC_SCAL_ALPHA_B, JCP_SCAL_BETA_B = JC_SCAL_BETA_B,
       JCP_SCAL_BETA_NODE = JC_SCAL_BETA_NODE, JCP_SCAL_FB = JC_SCAL_FB, JCP_SCAL_FC = JC_SCAL_FC };
#include "jc_power_point.cuh"

namespace {

constexpr int ADJ_NPT = 8;        // Limber nodes per thread
constexpr int ADJ_SCAL = 12;      // scal fields 0..11 cover every Eisenstein-Hu constant the point function reads

struct DevMath {
  const double* tab;
  __device__ __forceinline__ double log(double x) const { return jcm_log_t(x, tab); }       // arguments >= 1 on this path
  __device__ __forceinline__ double exp(double x) const { return jcm_exp_t<true>(x, tab); }
  __device__ __forceinline__ double expb(double x) const { return jcm_exp_t<false>(x, tab); }
  __device__ __forceinline__ double sin(double x) const { return jcm_sin(x); }
  __device__ __forceinline__ double rcbrt(double x) const { return jcm_rcbrt(x); }
  __device__ __forceinline__ double rcp(double x) const { return jcm_rcp(x); }
};

// The sweep's gradient, input by input, in registers (the field numbers are compile-time constants after inlining: entries the
// sweep never writes cost nothing).  The dot products with the tangent tables run AFTER the sweep: their 18 x NT + 9 x NT + NT
// loads are then independent of any arithmetic and go out back to back -- with the multiply-adds inside the sweep every input
// put a load latency in front of NT dependent FMAs (measured: 25 % of the FP64 pipe at 8 warps per SM).
struct GradRegs {
  double gn[JC_NODE_FIELDS], gs[ADJ_SCAL], ge;
  __device__ __forceinline__ void node(int f, double g) { gn[f] = g; }
  __device__ __forceinline__ void scal(int f, double g) { gs[f] = g; }
  __device__ __forceinline__ void ell(double g) { ge = g; }
};

template <int NT>
__global__ void __launch_bounds__(256, NT <= 3 ? 2 : 1) jc_power_adj_kernel(JcDevPlan pl, Ws ws, unsigned inv_L) {
  constexpr int NGRP = (JC_NA + ADJ_NPT - 1) / ADJ_NPT;
  __shared__ __align__(16) double s_tab[JCM_TAB_DOUBLES];
  __shared__ double s_ds[NT * ADJ_SCAL];
  const int c = blockIdx.y;
  const ptrdiff_t doff = ws.doff;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += 256) s_tab[i] = pl.math_tab[i];
  if (threadIdx.x < NT * ADJ_SCAL) {
    const int k = threadIdx.x / ADJ_SCAL, f = threadIdx.x - k * ADJ_SCAL;
    s_ds[threadIdx.x] = scp[f + jc_jvp_plane(k) * doff];
  }
  __syncthreads();
  const DevMath m{s_tab};
  const bool halofit = pl.nonlinear != 0, smith = pl.nonlinear == JC_PK_HALOFIT_SMITH2003;
  JcPointIn<double> in;
  in.inv13keq = scp[JC_SCAL_INV13KEQ]; in.beta_c = scp[JC_SCAL_BETA_C]; in.c14ac = scp[JC_SCAL_C14_ALPHA_C];
  in.shd = scp[JC_SCAL_SH_D]; in.alpha_b = scp[JC_SCAL_ALPHA_B]; in.beta_b = scp[JC_SCAL_BETA_B];
  in.beta_node = scp[JC_SCAL_BETA_NODE]; in.fb = scp[JC_SCAL_FB]; in.fc = scp[JC_SCAL_FC];
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;

  for (unsigned idx = blockIdx.x * 256 + threadIdx.x; idx < (unsigned)(NGRP * pl.L); idx += gridDim.x * 256) {
    const int grp = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
    const int l = (int)idx - grp * pl.L;
    in.lnl = pl.lnellp5[l]; in.lp5 = pl.ellp5[l]; in.l108 = pl.ell108[l]; in.l14 = pl.ell14[l]; in.lm3 = pl.ellm3[l];
    const double* lpp = ws.ellpow + (size_t)c * pl.Lpad + l;
    in.lpns = lpp[0];
    double dlp[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) dlp[k] = lpp[jc_jvp_plane(k) * doff];
    const int n0 = grp * ADJ_NPT;
    const int n1 = min(n0 + ADJ_NPT, JC_NA);
    double* vout = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;
#pragma unroll 1
    for (int n = n0; n < n1; ++n) {
      const double* ndn = nd + n;
#define NODE(f) ndn[(size_t)(f) * JC_NA_PAD]
      in.invchic = NODE(JC_NODE_INVCHIC); in.lnchic = NODE(JC_NODE_LNCHIC); in.nq108 = NODE(JC_NODE_NQ108);
      in.nsilk = NODE(JC_NODE_NSILK); in.namp = NODE(JC_NODE_NAMP); in.gk = NODE(JC_NODE_GK);
      if (halofit) {
        in.rnl = NODE(JC_NODE_RNL); in.lnknl = NODE(JC_NODE_LNKNL); in.beta = NODE(JC_NODE_BETA); in.alpha = NODE(JC_NODE_ALPHA);
        in.e1 = NODE(JC_NODE_E1); in.e2 = NODE(JC_NODE_E2); in.p3 = NODE(JC_NODE_P3); in.lncf = NODE(JC_NODE_LNCF);
        in.an = NODE(JC_NODE_AN); in.nu = NODE(JC_NODE_NU); in.bn = NODE(JC_NODE_BN);
        in.mu = smith ? NODE(JC_NODE_MU) : 0.0;
      }
#undef NODE
      GradRegs G;
      const double V = jc_point_adjoint(in, m, halofit, smith, G);
      double a[NT];
#pragma unroll
      for (int k = 0; k < NT; ++k) a[k] = G.ge * dlp[k];
      constexpr int NF_LIN[7] = {JC_NODE_INVCHIC, JC_NODE_NQ108, JC_NODE_NSILK, JC_NODE_NAMP, JC_NODE_GK, 0, 0};
      constexpr int NF_HALO[12] = {JC_NODE_LNCHIC, JC_NODE_RNL, JC_NODE_LNKNL, JC_NODE_BETA, JC_NODE_ALPHA, JC_NODE_E1, JC_NODE_E2,
                                   JC_NODE_P3, JC_NODE_LNCF, JC_NODE_AN, JC_NODE_NU, JC_NODE_BN};
      constexpr int SF[9] = {JC_SCAL_INV13KEQ, JC_SCAL_BETA_C, JC_SCAL_C14_ALPHA_C, JC_SCAL_SH_D, JC_SCAL_ALPHA_B, JC_SCAL_BETA_B,
                             JC_SCAL_BETA_NODE, JC_SCAL_FB, JC_SCAL_FC};
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int k = 0; k < NT; ++k) a[k] = fma(G.gn[NF_LIN[i]], ndn[(size_t)NF_LIN[i] * JC_NA_PAD + jc_jvp_plane(k) * doff], a[k]);
      if (halofit) {
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
          for (int k = 0; k < NT; ++k) a[k] = fma(G.gn[NF_HALO[i]], ndn[(size_t)NF_HALO[i] * JC_NA_PAD + jc_jvp_plane(k) * doff], a[k]);
        if (smith) {
#pragma unroll
          for (int k = 0; k < NT; ++k) a[k] = fma(G.gn[JC_NODE_MU], ndn[(size_t)JC_NODE_MU * JC_NA_PAD + jc_jvp_plane(k) * doff], a[k]);
        }
      }
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int k = 0; k < NT; ++k) a[k] = fma(G.gs[SF[i]], s_ds[k * ADJ_SCAL + SF[i]], a[k]);
      double* vp = vout + (size_t)n * pl.Lpad;
      vp[0] = V;
#pragma unroll
      for (int k = 0; k < NT; ++k) vp[jc_jvp_plane(k) * doff] = a[k];
    }
  }
}

template <int NT>
void launch_adj(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  constexpr int NGRP = (JC_NA + ADJ_NPT - 1) / ADJ_NPT;
  const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
  const int full = (NGRP * pl.L + 255) / 256;
  jc_power_adj_kernel<NT><<<dim3(full < 8 ? full : 8, chunk), 256, 0, s>>>(pl, ws, inv_L);
}

}  // namespace

bool jc_power_adj_supported(const JcDevPlan& pl, int ntan) {
  return ntan >= 3 && ntan <= JC_JVP_ADJ_MAX && pl.transfer != JC_TF_EISENSTEIN_HU_NOWIGGLE && !pl.grid_mode;
}

void jc_launch_power_adj(const JcDevPlan& pl, const Ws& ws, int chunk, int ntan, cudaStream_t s) {
  switch (ntan) {
    case 3: launch_adj<3>(pl, ws, chunk, s); break;
    case 4: launch_adj<4>(pl, ws, chunk, s); break;
    case 5: launch_adj<5>(pl, ws, chunk, s); break;
    case 6: launch_adj<6>(pl, ws, chunk, s); break;
    case 7: launch_adj<7>(pl, ws, chunk, s); break;
    default: launch_adj<8>(pl, ws, chunk, s); break;
  }
}
