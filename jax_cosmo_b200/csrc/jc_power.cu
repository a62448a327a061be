// jc_power.cu -- K3: V[n,l] = geom_n * P(k = (l+1/2)/max(chi_n,1), a_n) for every (node, ell).
//
// A thread owns one ell and NPT (= 16) consecutive Limber nodes (l fastest across lanes, so a warp shares its
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
//   * exp/log are table driven (jc_math.cuh: 2^n T[j] p4(r), 256 entries per octave; {c_j, -ln c_j} + log1p degree 7), the tables
//     are staged in shared memory once per CTA and the CTA strides over its cosmology's index space;
//     sin/rcbrt/rcp and all polynomial coefficients are constant-bank DFMA operands.
// Template on the scalar type: double (hot path) or Dual (JVP pass, jc_dual.cuh).
#include <cstdlib>

#include "jc_internal.cuh"
#include "jc_dual.cuh"

namespace {

struct PowerK {
  double e1, c699, c142, c386, c18, inv54, inv52, eighth, quarter, c043, two_e1, c625, c731;
};
static __constant__ PowerK PK = {2.718281828459045 /* np.exp(1.0) */, 69.9, 14.2, 386.0, 1.8,
                                 1.0 / 5.4, 1.0 / 5.2, 0.125, 0.25, 0.43, 2.0 * 2.718281828459045, 62.5, 731.0};

// Per-cosmology constants of the transfer function (transfer.py:47-136), loaded once per thread.
template <class T> struct EhK {
  T inv13keq, c14ac, b18, shd, alpha_b, fb, fc, bnode3, bb3, alpha_g, omh_t27;
};
template <class T> __device__ __forceinline__ EhK<T> eh_load(const double* scp, ptrdiff_t doff) {
  auto SC = [&](int f) { return JxMem<T>::ld(scp + f, doff); };
  EhK<T> E;
  E.inv13keq = SC(JC_SCAL_INV13KEQ); E.c14ac = SC(JC_SCAL_C14_ALPHA_C); E.b18 = PK.c18 * SC(JC_SCAL_BETA_C);
  E.shd = SC(JC_SCAL_SH_D); E.alpha_b = SC(JC_SCAL_ALPHA_B); E.fb = SC(JC_SCAL_FB); E.fc = SC(JC_SCAL_FC);
  const T bnode = SC(JC_SCAL_BETA_NODE), bb = SC(JC_SCAL_BETA_B);
  E.bnode3 = bnode * bnode * bnode; E.bb3 = bb * bb * bb;
  E.alpha_g = SC(JC_SCAL_ALPHA_GAMMA); E.omh_t27 = SC(JC_SCAL_OMH_T27);  // no-wiggle fit only
  return E;
}

// T(k): Eisenstein & Hu with wiggles (transfer.py:113-153) or the no-wiggle fit (transfer.py:92-105).
// q108 = q^1.08 and ks14 = (k / k_silk)^1.4 come from the caller (separable tables in the point kernel, exp in the
// table build); the ~14 divisions are merged into one reciprocal.
template <class T, bool NOWIG>
__device__ __forceinline__ T eh_point(const EhK<T>& E, T k, T q108, T ks14, const double* __restrict__ s_tab) {
  if constexpr (NOWIG) {
    const T ks43 = PK.c043 * (k * E.shd);
    const T k2 = ks43 * ks43;
    const T q = k * jx_rcp(E.omh_t27 * (E.alpha_g + (JCK.one - E.alpha_g) * jx_rcp(k2 * k2 + JCK.one)));
    const T L = jx_log_t(PK.c18 * q + PK.two_e1, s_tab);
    const T Wn = PK.c625 * q + JCK.one;                  // C = 14.2 + 731/Wn
    const T LW = L * Wn;
    return LW * jx_rcp(LW + (PK.c142 * Wn + PK.c731) * (q * q));
  } else {
    const T q = k * E.inv13keq;
    const T q2 = q * q;
    const T W = PK.c699 * q108 + JCK.one;  // 1 + 69.9 q^1.08
    const T U1 = PK.c142 * W + PK.c386;    // C(alpha=1) W
    const T U2 = E.c14ac * W + PK.c386;    // C(alpha_c) W
    const T L1 = jx_log_t(E.b18 * q + PK.e1, s_tab);
    const T L2 = jx_log_t(PK.c18 * q + PK.e1, s_tab);
    const T L1W = L1 * W, L2W = L2 * W;
    const T N1 = U1 * q2 + L1W;  // T~(k,1,beta_c)       = L1W / N1
    const T N2 = U2 * q2 + L1W;  // T~(k,alpha_c,beta_c) = L1W / N2
    const T N3 = U1 * q2 + L2W;  // T~(k,1,1)            = L2W / N3
    const T ks = k * E.shd;
    const T x54 = ks * PK.inv54;
    const T x54_2 = x54 * x54;
    const T Fm1 = x54_2 * x54_2;  // f = 1/(1+Fm1)
    // Tc = f T1 + (1-f) T2 = L1W (N2 + Fm1 N1) / ((1+Fm1) N1 N2)
    const T numC = L1W * (Fm1 * N1 + N2);
    const T denC = (JCK.one + Fm1) * (N1 * N2);
    const T ks2 = ks * ks, ks3 = ks2 * ks;
    const T arg = ks2 * jx_rcbrt(ks3 + E.bnode3);  // k s~ = ks^2 / cbrt(ks^3 + beta_node^3)
    const T x52 = ks * PK.inv52;
    const T X52 = x52 * x52 + JCK.one;
    const T BB = ks3 + E.bb3;  // 1/(1+(beta_b/ks)^3) = ks^3 / BB
    const T silk = jx_exp_t(-ks14, s_tab);  // exp(-(k/k_silk)^1.4)
    // Tb = [T3/X52 + alpha_b ks^3/BB silk] sin(arg)/arg
    const T N3X = N3 * X52;
    const T numB = (L2W * BB + E.alpha_b * ks3 * silk * N3X) * jx_sin(arg);
    const T denB = N3X * BB * arg;
    return ((E.fb * numB) * denC + E.fc * numC * denB) * jx_rcp(denB * denC);
  }
}

// NPT: Limber nodes per thread; MINB: CTAs per SM the register allocation must allow; NOWIG: the
// no-wiggle Eisenstein-Hu fit (transfer.py:99-105) instead of the default "eisenhu_osc".
template <class T, int NPT, int MINB, bool NOWIG>
__global__ void __launch_bounds__(256, MINB) jc_power_kernel(JcDevPlan pl, Ws ws, unsigned inv_L) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
  __shared__ __align__(16) double s_tab[JCM_TAB_DOUBLES];
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += 256) s_tab[i] = pl.math_tab[i];
  __syncthreads();
  const ptrdiff_t doff = ws.doff;
  const int c = blockIdx.y;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  const EhK<T> E = eh_load<T>(scp, doff);  // cosmology side (transfer.py:47-136)
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;
#define NODE(f) JxMem<T>::ld(nd + (f)*JC_NA_PAD + n, doff)

  for (unsigned idx = blockIdx.x * 256 + threadIdx.x; idx < (unsigned)(NGRP * pl.L); idx += gridDim.x * 256) {
    // idx / L: multiply-high by inv_L = ceil(2^32 / L) is exact while 513 L^2 < 2^32 (inv_L = 0 otherwise)
    const int grp = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
    const int l = (int)idx - grp * pl.L;
    // ell side
    const double lnl = pl.lnellp5[l], lp5 = pl.ellp5[l], l108 = pl.ell108[l], l14 = pl.ell14[l], lm3 = pl.ellm3[l];
    const T lpns = JxMem<T>::ld(ws.ellpow + (size_t)c * pl.Lpad + l, doff);
    const int n0 = grp * NPT;
    const int n1 = min(n0 + NPT, JC_NA);
    double* vout = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;

#pragma unroll 1
    for (int n = n0; n < n1; ++n) {
      const T lnk = lnl - NODE(JC_NODE_LNCHIC);
      const T k = lp5 * NODE(JC_NODE_INVCHIC);  // angular_cl.py:73
      const T Tk = eh_point<T, NOWIG>(E, k, NOWIG ? T(0.0) : T(l108 * NODE(JC_NODE_NQ108)), NOWIG ? T(0.0) : T(l14 * NODE(JC_NODE_NSILK)), s_tab);
      // ---- Delta^2_L = k^3 P_lin / (2 pi^2)  (power.py:49-52, :250) -----------------------------------
      const T d2l = lpns * NODE(JC_NODE_NAMP) * (Tk * Tk);
      T d2;
      if (pl.nonlinear) {  // halofit, takahashi2012 (power.py:246-262)
        const T y = k * NODE(JC_NODE_RNL);
        const T lny = lnk - NODE(JC_NODE_LNKNL);
        const T y2 = y * y;
        // Delta^2_Q = D2L (1+D2L)^beta / (1 + alpha D2L) exp(-(y/4 + y^2/8))
        const T Nq = d2l * jx_exp_t(NODE(JC_NODE_BETA) * jx_log_t(JCK.one + d2l, s_tab) - (y2 * PK.eighth + PK.quarter * y), s_tab);
        const T Dq = NODE(JC_NODE_ALPHA) * d2l + JCK.one;
        // y^(3 f1), y^(f2), c f3 y^(3-gamma): |ln y| < 30 and exponents of order one -- bounded arguments, no clamp
        const T ye1 = jx_exp_tb(NODE(JC_NODE_E1) * lny, s_tab);
        const T ye2 = jx_exp_tb(NODE(JC_NODE_E2) * lny, s_tab);
        const T cfy = jx_exp_tb(NODE(JC_NODE_P3) * (NODE(JC_NODE_LNCF) + lny), s_tab);
        const T Nh = NODE(JC_NODE_AN) * ye1 * y2;
        T ynu = y2 + NODE(JC_NODE_NU);  // 1/(1 + mu/y + nu/y^2) = y^2 / (y^2 + mu y + nu)
        if (pl.nonlinear == JC_PK_HALOFIT_SMITH2003) ynu = ynu + NODE(JC_NODE_MU) * y;  // mu = 0 in takahashi2012
        const T Dh = (NODE(JC_NODE_BN) * ye2 + JCK.one + cfy) * ynu;
        d2 = (Nq * Dh + Nh * Dq) * jx_rcp(Dq * Dh);  // Delta^2_Q + Delta^2_H
      } else {
        d2 = d2l;
      }
      // P = 2 pi^2 / k^3 * Delta^2 ;  V = P * geom = Delta^2 (l+1/2)^-3 * [geom 2 pi^2 chi_c^3]
      JxMem<T>::st(vout + (size_t)n * pl.Lpad, doff, d2 * lm3 * NODE(JC_NODE_GK));
    }
  }
#undef NODE
}


// ---------------------------------------------------------------------------------------------------------------------
// K3 with the transfer function tabulated per cosmology (the hot-path default for double, non-grid plans with >= 32 ell).
//
// T(k) depends on k only, and the 51,300 (ell, node) points of a cosmology sample it densely: ~80 of the exact kernel's
// 186 FP64 instructions per point re-evaluate the same smooth function.  Here one CTA (512 threads, two per SM) owns a
// cosmology: it evaluates T(k) with the exact formula at NT nodes uniform in ln k over the cosmology's own range
// [ln(l_min+1/2) - ln chi_max, ln(l_max+1/2)] (spacing h = range / (NT - 7) = 0.0023 at the bench configuration),
// derives the node slopes h T'(k_j) by 4th-order central differences, keeps {T, h T'} pairs in shared memory (96 KB) and
// evaluates every point by cubic Hermite interpolation (2 LDS.128 + 11 FP64 instead of 2 log, 1 exp, 1 sin, 1 rcbrt and
// ~45 multiply-adds).  The halofit part uses the reduced-degree exp / log of jc_math.cuh.
// Accuracy: the interpolation error is largest where the baryon wiggles are densest in ln k (k ~ 0.4 h/Mpc: 8e-10 on a single V entry);
// on C_ell, an integral over 513 nodes, it is <= 6e-11 (oracle experiment over the config-5 box, profiles/r02_power_tab.md) against the
// path's 1e-6 bar.  The exact kernel above stays the one behind grid plans, JVP passes, L < 32 and JC_POWER_EXACT=1,
// and a cosmology whose range would need h > 0.004 takes the exact formula inside this kernel (warp-uniform branch).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TAB_NT = 6144;           // table nodes (3 pad nodes on either side of the range)
constexpr int TAB_THREADS = 512;
constexpr double TAB_H_MAX = 0.004;

// The (ell, node) points of one cosmology for the CTA of jc_power_tab_kernel.  TAB: T(k) by cubic Hermite interpolation
// in the shared-memory table; !TAB: the exact formula (cosmologies whose ln k range is too wide for the table) -- a
// separate instantiation so that the EH constants are not live in the tabulated loop.
template <int NPT, bool NOWIG, bool TAB>
__device__ __forceinline__ void tab_points(const JcDevPlan& pl, const Ws& ws, int c, unsigned inv_L, const EhK<double>& E,
                                           const double* __restrict__ nd, const double2* __restrict__ tk,
                                           const double* __restrict__ s_tab, double x0, double inv_h) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
#define NODE(f) nd[(f)*JC_NA_PAD + n]
  for (unsigned idx = threadIdx.x; idx < (unsigned)(NGRP * pl.L); idx += TAB_THREADS) {
    const int grp = inv_L ? (int)__umulhi(idx, inv_L) : (int)(idx / (unsigned)pl.L);
    const int l = (int)idx - grp * pl.L;
    const double lnl = pl.lnellp5[l], lp5 = pl.ellp5[l], lm3 = pl.ellm3[l];
    const double lpns = ws.ellpow[(size_t)c * pl.Lpad + l];
    const int n0 = grp * NPT;
    const int n1 = min(n0 + NPT, JC_NA);
    double* vout = ws.vtab + (size_t)c * JC_NA * pl.Lpad + l;
#pragma unroll 1
    for (int n = n0; n < n1; ++n) {
      const double lnk = lnl - NODE(JC_NODE_LNCHIC);
      const double k = lp5 * NODE(JC_NODE_INVCHIC);  // angular_cl.py:73
      double Tk;
      if constexpr (TAB) {
        const double u = (lnk - x0) * inv_h;
        int i = __double2int_rz(u);
        i = max(2, min(i, TAB_NT - 4));
        const double t = u - (double)i;
        const double2 a = tk[i], b = tk[i + 1];
        const double D = b.x - a.x;
        const double c3 = (a.y + b.y) - (D + D);  // cubic Hermite in t: f0 + t (d0 + t (c2 + t c3))
        const double c2 = (D - a.y) - c3;
        Tk = fma(t, fma(t, fma(t, c3, c2), a.y), a.x);
      } else {
        Tk = eh_point<double, NOWIG>(E, k, NOWIG ? 0.0 : pl.ell108[l] * NODE(JC_NODE_NQ108), NOWIG ? 0.0 : pl.ell14[l] * NODE(JC_NODE_NSILK), s_tab);
      }
      const double d2l = lpns * NODE(JC_NODE_NAMP) * (Tk * Tk);
      double d2;
      if (pl.nonlinear) {  // halofit (power.py:246-262), same association as the exact kernel
        const double y = k * NODE(JC_NODE_RNL);
        const double lny = lnk - NODE(JC_NODE_LNKNL);
        const double y2 = y * y;
        const double Nq = d2l * jcm_exp_t3(NODE(JC_NODE_BETA) * jcm_log_t4(JCK.one + d2l, s_tab) - (y2 * PK.eighth + PK.quarter * y), s_tab);
        const double Dq = NODE(JC_NODE_ALPHA) * d2l + JCK.one;
        const double ye1 = jcm_exp_t3<false>(NODE(JC_NODE_E1) * lny, s_tab);
        const double ye2 = jcm_exp_t3<false>(NODE(JC_NODE_E2) * lny, s_tab);
        const double cfy = jcm_exp_t3<false>(NODE(JC_NODE_P3) * (NODE(JC_NODE_LNCF) + lny), s_tab);
        const double Nh = NODE(JC_NODE_AN) * ye1 * y2;
        double ynu = y2 + NODE(JC_NODE_NU);
        if (pl.nonlinear == JC_PK_HALOFIT_SMITH2003) ynu = ynu + NODE(JC_NODE_MU) * y;
        const double Dh = (NODE(JC_NODE_BN) * ye2 + JCK.one + cfy) * ynu;
        d2 = (Nq * Dh + Nh * Dq) * jcm_rcp(Dq * Dh);
      } else {
        d2 = d2l;
      }
      vout[(size_t)n * pl.Lpad] = d2 * lm3 * NODE(JC_NODE_GK);
    }
  }
#undef NODE
}

template <int NPT, bool NOWIG>
__global__ void __launch_bounds__(TAB_THREADS, 2) jc_power_tab_kernel(JcDevPlan pl, Ws ws, unsigned inv_L, double lnl_min,
                                                                       double lnl_max) {
  extern __shared__ __align__(16) double smem_p[];
  double2* tk = reinterpret_cast<double2*>(smem_p);  // [TAB_NT] {T(k_j), h dT/dlnk(k_j)}
  double* s_tab = smem_p + 2 * TAB_NT;
  for (int i = threadIdx.x; i < JCM_TAB_DOUBLES; i += TAB_THREADS) s_tab[i] = pl.math_tab[i];
  const int c = blockIdx.x;
  const double* scp = ws.scal + (size_t)c * JC_SCAL_FIELDS;
  const EhK<double> E = eh_load<double>(scp, 0);
  const double* nd = ws.node + (size_t)c * JC_NODE_FIELDS * JC_NA_PAD;
  // ln k range of this cosmology: chi decreases with the node index, node 0 = a_min holds the largest chi
  const double lo = lnl_min - nd[JC_NODE_LNCHIC * JC_NA_PAD + 0], hi = lnl_max - nd[JC_NODE_LNCHIC * JC_NA_PAD + JC_NA - 1];
  const double h = (hi - lo) * (1.0 / (TAB_NT - 7));
  const bool use_tab = h <= TAB_H_MAX;
  const double x0 = lo - 3.0 * h, inv_h = 1.0 / h;
  __syncthreads();
  if (use_tab) {
    const double ln13keq = scp[JC_SCAL_LN13KEQ], lnksilk = scp[JC_SCAL_LNKSILK];
    for (int j = threadIdx.x; j < TAB_NT; j += TAB_THREADS) {
      const double lnk = fma((double)j, h, x0);
      const double k = jcm_exp_t(lnk, s_tab);
      double q108 = 0.0, ks14 = 0.0;
      if (!NOWIG) {
        q108 = jcm_exp_t(1.08 * (lnk - ln13keq), s_tab);
        ks14 = jcm_exp_t(1.4 * (lnk - lnksilk), s_tab);
      }
      tk[j].x = eh_point<double, NOWIG>(E, k, q108, ks14, s_tab);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < TAB_NT; j += TAB_THREADS) {
      double d = 0.0;
      if (j >= 2 && j < TAB_NT - 2)
        d = (8.0 * (tk[j + 1].x - tk[j - 1].x) - (tk[j + 2].x - tk[j - 2].x)) * (1.0 / 12.0);
      tk[j].y = d;
    }
    __syncthreads();
  }
  if (use_tab) tab_points<NPT, NOWIG, true>(pl, ws, c, inv_L, E, nd, tk, s_tab, x0, inv_h);
  else tab_points<NPT, NOWIG, false>(pl, ws, c, inv_L, E, nd, tk, s_tab, x0, inv_h);
}

template <int NPT>
void launch_power_tab(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
  const size_t smem = (size_t)(2 * TAB_NT + JCM_TAB_DOUBLES) * sizeof(double);
  static unsigned long long attr_done = 0;
  JC_ONCE_PER_DEVICE(attr_done, {
    cudaFuncSetAttribute(jc_power_tab_kernel<NPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(jc_power_tab_kernel<NPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (pl.transfer == JC_TF_EISENSTEIN_HU_NOWIGGLE)
    jc_power_tab_kernel<NPT, true><<<chunk, TAB_THREADS, smem, s>>>(pl, ws, inv_L, pl.lnl_min, pl.lnl_max);
  else
    jc_power_tab_kernel<NPT, false><<<chunk, TAB_THREADS, smem, s>>>(pl, ws, inv_L, pl.lnl_min, pl.lnl_max);
}

// split: CTAs per cosmology (each strides over the index space)
template <class T, int NPT, int MINB>
void launch_power_cfg(const JcDevPlan& pl, const Ws& ws, int chunk, int split, cudaStream_t s) {
  constexpr int NGRP = (JC_NA + NPT - 1) / NPT;
  const unsigned inv_L = (pl.L >= 2 && pl.L <= 2048) ? (unsigned)((0x100000000ull + pl.L - 1) / pl.L) : 0u;
  const int full = (NGRP * pl.L + 255) / 256;
  const int gx = split < full ? split : full;
  if (pl.transfer == JC_TF_EISENSTEIN_HU_NOWIGGLE)
    jc_power_kernel<T, NPT, MINB, true><<<dim3(gx, chunk), 256, 0, s>>>(pl, ws, inv_L);
  else
    jc_power_kernel<T, NPT, MINB, false><<<dim3(gx, chunk), 256, 0, s>>>(pl, ws, inv_L);
}

}  // namespace

void jc_launch_power(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  static int cfg = -1, tab_npt = 0;
  if (cfg < 0) {  // tuning knobs
    const char* e = getenv("JC_POWER_CFG");
    const char* t = getenv("JC_POWER_TAB_NPT");
    tab_npt = t ? atoi(t) : 0;
    cfg = e ? atoi(e) : 0;
  }
  // jc_set_option("power_exact", 1) / JC_POWER_EXACT=1: the exact-formula kernel everywhere (A/B runs, stage tests)
  if (!g_jc_power_exact && !pl.grid_mode && pl.L >= 32 && ws.doff == 0) {  // tabulated transfer function (see jc_power_tab_kernel)
    switch (tab_npt) {
      case 4: launch_power_tab<4>(pl, ws, chunk, s); break;
      case 16: launch_power_tab<16>(pl, ws, chunk, s); break;
      default: launch_power_tab<8>(pl, ws, chunk, s); break;
    }
    return;
  }
  switch (cfg) {
    case 1: launch_power_cfg<double, 4, 1>(pl, ws, chunk, 8, s); break;  // unconstrained registers
    case 2: launch_power_cfg<double, 4, 3>(pl, ws, chunk, 8, s); break;  // 80 registers, 3 CTAs / SM
    case 3: launch_power_cfg<double, 4, 2>(pl, ws, chunk, 8, s); break;  // 128 registers, 2 CTAs / SM
    case 4: launch_power_cfg<double, 2, 3>(pl, ws, chunk, 8, s); break;
    case 5: launch_power_cfg<double, 8, 3>(pl, ws, chunk, 8, s); break;
    case 6: launch_power_cfg<double, 8, 4>(pl, ws, chunk, 8, s); break;   // 8 nodes per thread at 64 registers
    case 7: launch_power_cfg<double, 8, 4>(pl, ws, chunk, 4, s); break;
    case 8: launch_power_cfg<double, 16, 4>(pl, ws, chunk, 4, s); break;
    case 9: launch_power_cfg<double, 4, 4>(pl, ws, chunk, 16, s); break;
    case 10: launch_power_cfg<double, 32, 4>(pl, ws, chunk, 2, s); break;
    case 11: launch_power_cfg<double, 32, 4>(pl, ws, chunk, 4, s); break;
    case 12: launch_power_cfg<double, 64, 4>(pl, ws, chunk, 2, s); break;
    case 13: launch_power_cfg<double, 16, 4>(pl, ws, chunk, 2, s); break;
    case 14: launch_power_cfg<double, 16, 4>(pl, ws, chunk, 8, s); break;
    case 15: launch_power_cfg<double, 4, 4>(pl, ws, chunk, 8, s); break;  // the round's earlier default: 6.34 ms
    case 16: launch_power_cfg<double, 16, 3>(pl, ws, chunk, 8, s); break;  // 80 registers, no spills
    case 17: launch_power_cfg<double, 16, 5>(pl, ws, chunk, 8, s); break;  // 48 registers
    // fastest (profiles/r01_tuning.md): 16 nodes per thread (the ell-side loads and index arithmetic amortise over 16
    // points), 64 registers, 8 CTAs per cosmology: 6.13 ms
    default: launch_power_cfg<double, 16, 4>(pl, ws, chunk, 8, s); break;
  }
}

void jc_launch_power_jvp(const JcDevPlan& pl, const Ws& ws, int chunk, cudaStream_t s) {
  static int cfg = -1;
  if (cfg < 0) { const char* e = getenv("JC_POWER_JVP_CFG"); cfg = e ? atoi(e) : 0; }  // tuning knob
  switch (cfg) {
    case 1: launch_power_cfg<Dual, 1, 2>(pl, ws, chunk, 16, s); break;
    case 2: launch_power_cfg<Dual, 4, 2>(pl, ws, chunk, 8, s); break;
    case 3: launch_power_cfg<Dual, 4, 1>(pl, ws, chunk, 8, s); break;
    case 4: launch_power_cfg<Dual, 8, 2>(pl, ws, chunk, 8, s); break;
    case 5: launch_power_cfg<Dual, 16, 2>(pl, ws, chunk, 8, s); break;
    case 6: launch_power_cfg<Dual, 8, 3>(pl, ws, chunk, 8, s); break;
    case 7: launch_power_cfg<Dual, 16, 3>(pl, ws, chunk, 8, s); break;
    case 8: launch_power_cfg<Dual, 16, 4>(pl, ws, chunk, 8, s); break;
    case 9: launch_power_cfg<Dual, 1, 1>(pl, ws, chunk, 16, s); break;  // the round's earlier default
    // 16 nodes per thread, 3 CTAs / SM: 7-tangent batch JVP 38.6 -> 31.0 ms per 1024 cosmologies (scripts/jvp_throughput.py)
    default: launch_power_cfg<Dual, 16, 3>(pl, ws, chunk, 8, s); break;
  }
}
