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

constexpr int ADJ_NPT = 8;                  // Limber nodes per thread (one group)
constexpr int ADJ_GB = 5;                   // groups per CTA
constexpr int ADJ_NN = ADJ_GB * ADJ_NPT;    // nodes whose tables a CTA stages in shared memory
constexpr int ADJ_NF = 18;                  // node fields the point function reads
constexpr int ADJ_SCAL = 12;                // scal fields 0..11 cover every Eisenstein-Hu constant the point function reads

// compact order of the staged node fields: 5 of the linear part, 12 of halofit, mu (smith2003)
__constant__ int ADJ_FIELD[ADJ_NF] = {JC_NODE_INVCHIC, JC_NODE_NQ108, JC_NODE_NSILK, JC_NODE_NAMP, JC_NODE_GK,
                                      JC_NODE_LNCHIC, JC_NODE_RNL, JC_NODE_LNKNL, JC_NODE_BETA, JC_NODE_ALPHA, JC_NODE_E1,
                                      JC_NODE_E2, JC_NODE_P3, JC_NODE_LNCF, JC_NODE_AN, JC_NODE_NU, JC_NODE_BN, JC_NODE_MU};
enum { C_INVCHIC = 0, C_NQ108, C_NSILK, C_NAMP, C_GK, C_LNCHIC, C_RNL, C_LNKNL, C_BETA, C_ALPHA, C_E1, C_E2, C_P3, C_LNCF, C_AN,
       C_NU, C_BN, C_MU };

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
// sweep never writes cost nothing).  The dot products with the tangent tables run AFTER the sweep: their loads are then
// independent of any arithmetic and go out back to back.
struct GradRegs {
  double gn[JC_NODE_FIELDS], gs[ADJ_SCAL], ge;
  __device__ __forceinline__ void node(int f, double g) { gn[f] = g; }
  __device__ __forceinline__ void scal(int f, double g) { gs[f] = g; }
  __device__ __forceinline__ void ell(double g) { ge = g; }
};

// CTA = (cosmology, ADJ_GB groups of 8 nodes); thread = (group, ell) walking its group's nodes.  The node fields of the CTA's 40
// nodes and their NT tangents are staged once in shared memory as [field][plane][node] (plane 0 = value): the tangent planes of
// a table are workspace-chunks apart, so a point's 18 x NT tangent loads were 32-byte sectors of their own -- 25 % of them L2
// round trips at two warps per scheduler (ncu on the first version: long scoreboard 5.5 warps per issue, 23 % of the FP64
// pipe).  Staged, they are coalesced along the node index when loaded and warp-broadcast LDS when used.
template <int NT>
__global__ void __launch_bounds__(256, 1) jc_power_adj_kernel(JcDevPlan pl, Ws ws, unsigned inv_L) {
  constexpr int NGRP = (JC_NA + ADJ_NPT - 1) / ADJ_NPT;
  constexpr int NP = NT + 1;
  extern __shared__ __align__(16) double adj_smem[];
  double* s_tab = adj_smem;                          // [JCM_TAB_DOUBLES]
  double* s_ds = s_tab + JCM_TAB_DOUBLES;            // [NT][ADJ_SCAL]
  double* s_nd = s_ds + NT * ADJ_SCAL;               // [ADJ_NF][NP][ADJ_NN]
  const int c = blockIdx.y;
  const int g0 = blockIdx.x * ADJ_GB, n_a = g0 * ADJ_NPT;
  const int nn_cnt = min(ADJ_NN, JC_NA - n_a);
  const ptrdiff_t doff = ws.doff;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD + n_a;
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += 256) s_tab[i] = pl.math_tab[i];
  if (threadIdx.x < NT * ADJ_SCAL) {
    const int k = threadIdx.x / ADJ_SCAL, f = threadIdx.x - k * ADJ_SCAL;
    s_ds[threadIdx.x] = scp[f + jc_jvp_plane(k) * doff];
  }
  // staging: 8 loads in flight per thread before the first store (ncu: with load -> store per element the stores waited on the
  // long scoreboard for 17 % of the kernel's time)
  constexpr int STAGE_TOTAL = ADJ_NF * NP * ADJ_NN, STAGE_U = 8;
  for (int i0 = threadIdx.x; i0 < STAGE_TOTAL; i0 += 256 * STAGE_U) {
    double v[STAGE_U];
#pragma unroll
    for (int u = 0; u < STAGE_U; ++u) {
      const int i = i0 + u * 256;
      const int row = i / ADJ_NN, nn = i - row * ADJ_NN;
      const int f = row / NP, p = row - f * NP;
      const ptrdiff_t off = p ? (ptrdiff_t)jc_jvp_plane(p - 1) * doff : 0;
      v[u] = (i < STAGE_TOTAL && nn < nn_cnt) ? nd[(size_t)ADJ_FIELD[f] * JC_NA_PAD + nn + off] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < STAGE_U; ++u)
      if (i0 + u * 256 < STAGE_TOTAL) s_nd[i0 + u * 256] = v[u];
  }
  __syncthreads();
  const DevMath m{s_tab};
  const bool halofit = pl.nonlinear != 0, smith = pl.nonlinear == JC_PK_HALOFIT_SMITH2003;
  JcPointIn<double> in;
  in.inv13keq = scp[JC_SCAL_INV13KEQ]; in.beta_c = scp[JC_SCAL_BETA_C]; in.c14ac = scp[JC_SCAL_C14_ALPHA_C];
  in.shd = scp[JC_SCAL_SH_D]; in.alpha_b = scp[JC_SCAL_ALPHA_B]; in.beta_b = scp[JC_SCAL_BETA_B];
  in.beta_node = scp[JC_SCAL_BETA_NODE]; in.fb = scp[JC_SCAL_FB]; in.fc = scp[JC_SCAL_FC];
  const int ng = min(ADJ_GB, NGRP - g0);

  for (unsigned idx = threadIdx.x; idx < (unsigned)(ng * pl.L); idx += 256) {
    const int gl = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
    const int l = (int)idx - gl * pl.L;
    in.lnl = pl.lnellp5[l]; in.lp5 = pl.ellp5[l]; in.l108 = pl.ell108[l]; in.l14 = pl.ell14[l]; in.lm3 = pl.ellm3[l];
    const double* lpp = ws.ellpow + (size_t)c * pl.Lpad + l;
    in.lpns = lpp[0];
    double dlp[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) dlp[k] = lpp[jc_jvp_plane(k) * doff];
    const int nn0 = gl * ADJ_NPT;
    const int nn1 = min(nn0 + ADJ_NPT, nn_cnt);
    double* vout = ws.vtab + ((size_t)c * JC_NA + n_a) * pl.Lpad + l;
#pragma unroll 1
    for (int nn = nn0; nn < nn1; ++nn) {
      const double* sn = s_nd + nn;
#define SV(cf) sn[(cf) * NP * ADJ_NN]                    /* value of compact field cf at this node */
#define ST(cf, k) sn[((cf) * NP + 1 + (k)) * ADJ_NN]     /* its tangent along direction k */
      in.invchic = SV(C_INVCHIC); in.lnchic = SV(C_LNCHIC); in.nq108 = SV(C_NQ108); in.nsilk = SV(C_NSILK);
      in.namp = SV(C_NAMP); in.gk = SV(C_GK);
      if (halofit) {
        in.rnl = SV(C_RNL); in.lnknl = SV(C_LNKNL); in.beta = SV(C_BETA); in.alpha = SV(C_ALPHA); in.e1 = SV(C_E1);
        in.e2 = SV(C_E2); in.p3 = SV(C_P3); in.lncf = SV(C_LNCF); in.an = SV(C_AN); in.nu = SV(C_NU); in.bn = SV(C_BN);
        in.mu = smith ? SV(C_MU) : 0.0;
      }
      GradRegs G;
      const double V = jc_point_adjoint(in, m, halofit, smith, G);
      double a[NT];
#pragma unroll
      for (int k = 0; k < NT; ++k) a[k] = G.ge * dlp[k];
      constexpr int NODE_OF[ADJ_NF] = {JC_NODE_INVCHIC, JC_NODE_NQ108, JC_NODE_NSILK, JC_NODE_NAMP, JC_NODE_GK,
                                       JC_NODE_LNCHIC, JC_NODE_RNL, JC_NODE_LNKNL, JC_NODE_BETA, JC_NODE_ALPHA, JC_NODE_E1,
                                       JC_NODE_E2, JC_NODE_P3, JC_NODE_LNCF, JC_NODE_AN, JC_NODE_NU, JC_NODE_BN, JC_NODE_MU};
      constexpr int SF[9] = {JC_SCAL_INV13KEQ, JC_SCAL_BETA_C, JC_SCAL_C14_ALPHA_C, JC_SCAL_SH_D, JC_SCAL_ALPHA_B, JC_SCAL_BETA_B,
                             JC_SCAL_BETA_NODE, JC_SCAL_FB, JC_SCAL_FC};
#pragma unroll
      for (int cf = 0; cf < 5; ++cf)
#pragma unroll
        for (int k = 0; k < NT; ++k) a[k] = fma(G.gn[NODE_OF[cf]], ST(cf, k), a[k]);
      if (halofit) {
#pragma unroll
        for (int cf = 5; cf < 17; ++cf)
#pragma unroll
          for (int k = 0; k < NT; ++k) a[k] = fma(G.gn[NODE_OF[cf]], ST(cf, k), a[k]);
        if (smith) {
#pragma unroll
          for (int k = 0; k < NT; ++k) a[k] = fma(G.gn[JC_NODE_MU], ST(C_MU, k), a[k]);
        }
      }
#undef SV
#undef ST
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int k = 0; k < NT; ++k) a[k] = fma(G.gs[SF[i]], s_ds[k * ADJ_SCAL + SF[i]], a[k]);
      double* vp = vout + (size_t)nn * pl.Lpad;
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
  const size_t smem = (size_t)(JCM_TAB_DOUBLES + NT * ADJ_SCAL + ADJ_NF * (NT + 1) * ADJ_NN) * sizeof(double);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, cudaFuncSetAttribute(jc_power_adj_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jc_power_adj_kernel<NT><<<dim3((NGRP + ADJ_GB - 1) / ADJ_GB, chunk), 256, smem, s>>>(pl, ws, inv_L);
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
