// jc_power.cu -- K3: V[n,l] = geom_n * P(k = (l+1/2)/max(chi_n,1), a_n) for every (node, ell).
//
// A thread owns one ell and NPT consecutive Limber nodes (l fastest across lanes, so a warp shares its
// node constants and V rows are written coalesced); the ell-side table entries, the per-cosmology EH
// constants and the polynomial coefficients are loaded once per thread and reused over the NPT nodes.
// Eisenstein-Hu "eisenhu_osc" transfer (transfer.py:113-153), linear power (power.py:49-52) and
// halofit takahashi2012 (power.py:246-262) are evaluated entirely in registers.  Compared with the
// reference's expression tree the arithmetic is re-associated only (<= a few ulp):
//   * every fixed-exponent power law of k = (l+1/2)/chi_c is separable: k^p = (l+1/2)^p * chi_c^-p,
//     so q^1.08, (k/k_silk)^1.4, k^(3+n_s) and k^-3 are one multiply of a per-ell (plan / ws.ellpow)
//     and a per-node (setup kernel) table entry;
//   * (1+D2L)^beta * exp(-(y/4+y^2/8)) is one exp -- 5 exp, 3 log, 1 rcbrt, 1 sin remain per point;
//   * the ~14 divisions of T(k) and of Delta^2_Q + Delta^2_H are merged into two reciprocals;
//   * exp/log/sin/rcbrt come from jc_math.cuh (coefficients as constant-bank DFMA operands).
#include <cstdlib>

#include "jc_internal.cuh"
#include "jc_math.cuh"

namespace {


struct PowerK {
  double e1, c699, c142, c386, c18, inv54, inv52, eighth, quarter;
};
static __constant__ PowerK PK = {2.718281828459045 /* np.exp(1.0) */, 69.9, 14.2, 386.0, 1.8,
                                 1.0 / 5.4, 1.0 / 5.2, 0.125, 0.25};

// NPT: Limber nodes per thread; MINB: CTAs per SM the register allocation must allow.
// TAB: table-driven exp / log (tables staged in shared memory once per CTA; the CTA then strides over
// the (node group, ell) index space of its cosmology so that the staging is amortised).
template <int NPT, int MINB, bool TAB>
__global__ void __launch_bounds__(256, MINB) jc_power_kernel(JcDevPlan pl, Ws ws, unsigned inv_L) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
  __shared__ __align__(16) double s_tab[TAB ? JCM_TAB_DOUBLES : 2];
  if (TAB) {
    for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += 256) s_tab[i] = pl.math_tab[i];
    __syncthreads();
  }
#define EXPF(x) (TAB ? jcm_exp_t((x), s_tab) : jcm_exp(x))
#define LOGF(x) (TAB ? jcm_log_t((x), s_tab) : jcm_log(x))
  const int c = blockIdx.y;
  const double* sc = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  for (unsigned idx = blockIdx.x * 256 + threadIdx.x; idx < (unsigned)(NGRP * pl.L); idx += gridDim.x * 256) {
  // idx / L: multiply-high by inv_L = ceil(2^32 / L) is exact while 513 L^2 < 2^32 (inv_L = 0 otherwise)
  const int grp = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
  const int l = (int)idx - grp * pl.L;
  // ell side
  const double lnl = pl.lnellp5[l], lp5 = pl.ellp5[l], l108 = pl.ell108[l], l14 = pl.ell14[l];
  const double lm3 = pl.ellm3[l], lpns = ws.ellpow[(size_t)c * pl.Lpad + l];
  // cosmology side (transfer.py:47-136)
  const double inv13keq = sc[JC_SCAL_INV13KEQ], c14ac = sc[JC_SCAL_C14_ALPHA_C], b18 = PK.c18 * sc[JC_SCAL_BETA_C];
  const double shd = sc[JC_SCAL_SH_D], alpha_b = sc[JC_SCAL_ALPHA_B], fb = sc[JC_SCAL_FB], fc = sc[JC_SCAL_FC];
  const double bnode = sc[JC_SCAL_BETA_NODE], bb = sc[JC_SCAL_BETA_B];
  const double bnode3 = bnode * bnode * bnode, bb3 = bb * bb * bb;
  const int n0 = grp * NPT;
  const int n1 = min(n0 + NPT, JC_NA);
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;
  double* vout = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;
#define NODE(f) nd[(f)*JC_NA_PAD + n]

#pragma unroll 1
  for (int n = n0; n < n1; ++n) {
    const double lnk = lnl - NODE(JC_NODE_LNCHIC);
    const double k = lp5 * NODE(JC_NODE_INVCHIC);  // angular_cl.py:73
    // ---- Eisenstein & Hu (transfer.py:113-153) --------------------------------------------------
    const double q = k * inv13keq;
    const double q2 = q * q;
    const double W = fma(PK.c699, l108 * NODE(JC_NODE_NQ108), JCK.one);  // 1 + 69.9 q^1.08
    const double U1 = fma(PK.c142, W, PK.c386);                          // C(alpha=1) W
    const double U2 = fma(c14ac, W, PK.c386);                            // C(alpha_c) W
    const double L1 = LOGF(fma(b18, q, PK.e1));
    const double L2 = LOGF(fma(PK.c18, q, PK.e1));
    const double L1W = L1 * W, L2W = L2 * W;
    const double N1 = fma(U1, q2, L1W);  // T~(k,1,beta_c)       = L1W / N1
    const double N2 = fma(U2, q2, L1W);  // T~(k,alpha_c,beta_c) = L1W / N2
    const double N3 = fma(U1, q2, L2W);  // T~(k,1,1)            = L2W / N3
    const double ks = k * shd;
    const double x54 = ks * PK.inv54;
    const double x54_2 = x54 * x54;
    const double Fm1 = x54_2 * x54_2;  // f = 1/(1+Fm1)
    // Tc = f T1 + (1-f) T2 = L1W (N2 + Fm1 N1) / ((1+Fm1) N1 N2)
    const double numC = L1W * fma(Fm1, N1, N2);
    const double denC = (JCK.one + Fm1) * (N1 * N2);
    const double ks2 = ks * ks, ks3 = ks2 * ks;
    const double arg = ks2 * jcm_rcbrt(ks3 + bnode3);  // k s~ = ks^2 / cbrt(ks^3 + beta_node^3)
    const double x52 = ks * PK.inv52;
    const double X52 = fma(x52, x52, JCK.one);
    const double BB = ks3 + bb3;  // 1/(1+(beta_b/ks)^3) = ks^3 / BB
    const double silk = EXPF(-(l14 * NODE(JC_NODE_NSILK)));  // exp(-(k/k_silk)^1.4)
    // Tb = [T3/X52 + alpha_b ks^3/BB silk] sin(arg)/arg
    const double N3X = N3 * X52;
    const double numB = fma(L2W, BB, alpha_b * ks3 * silk * N3X) * jcm_sin(arg);
    const double denB = N3X * BB * arg;
    const double Tk = fma(fb * numB, denC, fc * numC * denB) * jcm_rcp(denB * denC);
    // ---- Delta^2_L = k^3 P_lin / (2 pi^2)  (power.py:49-52, :250) -------------------------------------
    const double d2l = lpns * NODE(JC_NODE_NAMP) * (Tk * Tk);
    double d2;
    if (pl.nonlinear) {  // halofit, takahashi2012 (power.py:246-262)
      const double y = k * NODE(JC_NODE_RNL);
      const double lny = lnk - NODE(JC_NODE_LNKNL);
      const double y2 = y * y;
      // Delta^2_Q = D2L (1+D2L)^beta / (1 + alpha D2L) exp(-(y/4 + y^2/8))
      const double Nq = d2l * EXPF(fma(NODE(JC_NODE_BETA), LOGF(JCK.one + d2l), -fma(y2, PK.eighth, PK.quarter * y)));
      const double Dq = fma(NODE(JC_NODE_ALPHA), d2l, JCK.one);
      const double ye1 = EXPF(NODE(JC_NODE_E1) * lny);
      const double ye2 = EXPF(NODE(JC_NODE_E2) * lny);
      const double cfy = EXPF(NODE(JC_NODE_P3) * (NODE(JC_NODE_LNCF) + lny));
      const double Nh = NODE(JC_NODE_AN) * ye1 * y2;
      const double Dh = (fma(NODE(JC_NODE_BN), ye2, JCK.one) + cfy) * (y2 + NODE(JC_NODE_NU));
      d2 = fma(Nq, Dh, Nh * Dq) * jcm_rcp(Dq * Dh);  // Delta^2_Q + Delta^2_H
    } else {
      d2 = d2l;
    }
    // P = 2 pi^2 / k^3 * Delta^2 ;  V = P * geom = Delta^2 (l+1/2)^-3 * [geom 2 pi^2 chi_c^3]
    vout[(size_t)n * pl.Lpad] = d2 * lm3 * NODE(JC_NODE_GK);
  }
  }  // idx
#undef NODE
#undef EXPF
#undef LOGF
}

// split: CTAs per cosmology (each strides over the index space); 0 = one CTA per 256 indices
template <int NPT, int MINB, bool TAB>
void launch_power_cfg(const JcDevPlan& pl, const Ws& ws, int chunk, int split, cudaStream_t s) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
  const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
  const int full = (NGRP * pl.L + 255) / 256;
  const int gx = split > 0 && split < full ? split : full;
  jc_power_kernel<NPT, MINB, TAB><<<dim3(gx, chunk), 256, 0, s>>>(pl, ws, inv_L);
}

}  // namespace

void jc_launch_power(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  static int cfg = -1;
  if (cfg < 0) { const char* e = getenv("JC_POWER_CFG"); cfg = e ? atoi(e) : 0; }  // tuning knob
  switch (cfg) {
    case 1: launch_power_cfg<4, 1, false>(pl, ws, chunk, 0, s); break;  // polynomial exp/log, 1 CTA / 256 indices
    case 2: launch_power_cfg<4, 1, true>(pl, ws, chunk, 4, s); break;
    case 3: launch_power_cfg<4, 1, true>(pl, ws, chunk, 16, s); break;
    case 4: launch_power_cfg<4, 1, true>(pl, ws, chunk, 8, s); break;   // unconstrained registers
    case 5: launch_power_cfg<2, 1, true>(pl, ws, chunk, 8, s); break;
    case 6: launch_power_cfg<1, 6, true>(pl, ws, chunk, 8, s); break;   // one node per thread, 40 registers
    // fastest (profiles/r01_tuning.md): table-driven exp/log, 64 registers, 8 CTAs per cosmology
    default: launch_power_cfg<4, 4, true>(pl, ws, chunk, 8, s); break;
  }
}
